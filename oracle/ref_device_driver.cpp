// TEST INFRASTRUCTURE (oracle/). Not product code, never shipped, never timed as the product.
//
// Recorder around the UNMODIFIED reference DEVICE-RUN path:
// emcSORSolver (equilibrium + non-equilibrium), calcEFieldAtGridPts, generateInitialParticles,
// emcNGPScheme::assignToMesh, emcSimulationResults::updateCurrentParticleConcentrations,
// emcBasicParticleHandler::driftScatterParticles and handleOhmicContacts, driven in the order of
// emcSimulation::calcEquilibriumCharacteristics / performEMCStep (include/emcSimulation.hpp:139-193).
// Built like ref_bulk_driver (overlay emcUtil.hpp with the recording RNG, -fno-access-control, no
// OpenMP) into oracle/_ref/.  Every intermediate grid and ensemble of every step is dumped.
//
// Plug-in variants (--scheme ngp|cic|nec|vwd, --electron emc|vwd, --surface-ymin-const p, --surface-ymax-mom h):
// emcNGPScheme / emcCICScheme / emcNECScheme / examples/mosfet2D/NECSchemeVWD.hpp, emcElectron /
// examples/mosfet2D/electronVWD.hpp, emcConstantSurfaceScatterMechanism / emcMomentumDependentSurfaceScatterMechanism.
//
// Scenario: a 2-D silicon bar (resistor2D.cpp:73-115 with its sizes as options) with ohmic contacts
// on the full XMIN / XMAX faces, optionally a gate on a YMIN segment and a second doping region (so
// that the Robin boundary term, region look-ups and Coulomb tables per region are exercised).
#ifndef DEVICE_DIM
#define DEVICE_DIM 2
#endif
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <SiliconFunctions.hpp> // -I $(REF)/examples

#include <PMSchemes/emcCICScheme.hpp>
#include <PMSchemes/emcNECScheme.hpp>
#include <PMSchemes/emcNGPScheme.hpp>
#include <ParticleHandler/emcBasicParticleHandler.hpp>
#include <SurfaceScatterMechanisms/emcConstantSurfaceScatterMechanism.hpp>
#include <SurfaceScatterMechanisms/emcMomentumDependentSurfaceScatterMechanism.hpp>
#include <ParticleType/emcElectron.hpp>
#include <PoissonSolver/emcSORSolver.hpp>
#include <emcDevice.hpp>
#include <emcGrainScatterMechanism.hpp>
#include <emcSimulationParameter.hpp>
#include <emcSimulationResults.hpp>

#if DEVICE_DIM == 2
#include <mosfet2D/NECSchemeVWD.hpp> // -I $(REF)/examples
#endif
#include <mosfet2D/electronVWD.hpp>

// -DDEVICE_DIM=3 builds the 3-D recorder (box lx x ly x lz, the contacts cover whole faces / a strip of YMIN over the full
// depth); the NEC schemes interpolate forces in 2-D only and exist in the 2-D build alone
using T = double;
const SizeType Dim = DEVICE_DIM;
using DeviceType = emcDevice<T, Dim>;
using Grid = emcGrid<T, Dim>;

struct Blob {
  std::ofstream os;
  explicit Blob(const std::string &path) : os(path, std::ios::binary) {}
  void put(const std::string &name, char dtype, const void *data, const std::vector<std::uint64_t> &dims,
           size_t elemSize) {
    std::uint32_t nl = name.size();
    os.write((const char *)&nl, 4);
    os.write(name.data(), nl);
    os.write(&dtype, 1);
    std::uint32_t nd = dims.size();
    os.write((const char *)&nd, 4);
    std::uint64_t n = 1;
    for (auto d : dims) {
      os.write((const char *)&d, 8);
      n *= d;
    }
    os.write((const char *)data, n * elemSize);
  }
  void f64(const std::string &n, const std::vector<double> &v, std::vector<std::uint64_t> dims = {}) {
    if (dims.empty())
      dims = {v.size()};
    put(n, 'd', v.data(), dims, 8);
  }
  void i64(const std::string &n, const std::vector<std::int64_t> &v, std::vector<std::uint64_t> dims = {}) {
    if (dims.empty())
      dims = {v.size()};
    put(n, 'q', v.data(), dims, 8);
  }
  void u64(const std::string &n, const std::vector<std::uint64_t> &v) { put(n, 'Q', v.data(), {v.size()}, 8); }
  template <class G> void grid(const std::string &n, const G &g) {
    std::vector<double> v(g.begin(), g.end());
    auto e = g.getExtent();
    std::vector<std::uint64_t> dims;
    for (size_t d = e.size(); d-- > 0;)
      dims.push_back(e[d]); // x fastest
    f64(n, v, dims);
  }
};

template <class Handler> static void dumpEnsemble(Blob &b, const std::string &p, Handler &h) {
  const auto &parts = h.particles[0];
  const auto &pos = h.positionsParticles[0];
  const size_t n = parts.size();
  std::vector<double> k, x, e, tau, g;
  std::vector<std::int64_t> idx;
  for (size_t i = 0; i < n; i++) {
    k.insert(k.end(), parts[i].k.begin(), parts[i].k.end());
    x.insert(x.end(), pos[i].begin(), pos[i].end());
    e.push_back(parts[i].energy);
    tau.push_back(parts[i].tau);
    g.push_back(parts[i].grainTau);
    idx.insert(idx.end(), {(std::int64_t)parts[i].valley, (std::int64_t)parts[i].subValley,
                           (std::int64_t)parts[i].region});
  }
  b.f64(p + "k", k, {n, 3});
  b.f64(p + "pos", x, {n, Dim});
  b.f64(p + "energy", e);
  b.f64(p + "tau", tau);
  b.f64(p + "label", g);
  b.i64(p + "idx", idx, {n, 3});
}

struct Options {
  double lz = 6e-8, hz = 2e-8; // 3-D build only
  double lx = 2e-7, ly = 1e-7, hx = 1e-8, hy = 2.5e-8, width = 1e-6, doping = 1e22, doping2 = 0, voltage = 0.05,
         dt = 1e-15, acc = 1e-4, omega = 1.8, emax = 4.0, gateVoltage = 0.5, surfYminConst = -1, surfYmaxMom = -1, grainRate = 0,
         grainProb = 0.5;
  int steps = 10, levels = 1000, gate = 0;
  int snapEvery = 1; // > 1: the grids and ensembles only of every snapEvery-th step and of the last one (long chained runs);
                     // the per-contact counters and the draw marks of EVERY step are always kept
  unsigned long seed = 5;
  std::string out = "device.blob", scheme = "ngp", electron = "emc";
};

template <class Electron> std::unique_ptr<Electron> makeElectron(const Options &o);
template <> std::unique_ptr<emcElectron<T, DeviceType>> makeElectron(const Options &o) {
  return std::make_unique<emcElectron<T, DeviceType>>(o.levels, o.emax, false);
}
template <> std::unique_ptr<electronVWD<T, DeviceType>> makeElectron(const Options &o) {
  return std::make_unique<electronVWD<T, DeviceType>>(o.levels, o.emax);
}

template <class PMScheme, class Electron> int run(const Options &o) {
  using Handler = emcBasicParticleHandler<T, DeviceType, PMScheme>;
  using Solver = emcSORSolver<T, DeviceType, Handler>;
  const double lx = o.lx, ly = o.ly, hx = o.hx, hy = o.hy, width = o.width, doping = o.doping, doping2 = o.doping2,
               voltage = o.voltage, dt = o.dt, acc = o.acc, omega = o.omega, emax = o.emax, gateVoltage = o.gateVoltage;
  const int steps = o.steps, levels = o.levels, gate = o.gate;
  const unsigned long seed = o.seed;
  const std::string out = o.out;
  std::vector<std::uint64_t> draws;
  RecordingRNG::sink() = &draws;
  std::streambuf *oldBuf = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());

#if DEVICE_DIM == 2
  DeviceType device{Silicon::getSiliconMaterial<T>(), {lx, ly}, {hx, hy}};
  device.setDeviceWidth(width);
  device.addConstantDopingRegion({0, 0}, {lx, ly}, doping);
  std::vector<int> regions = {0};
  if (doping2 != 0) {
    device.addConstantDopingRegion({lx / 2, 0}, {lx, ly}, doping2);
    regions.push_back(1);
  }
  device.addOhmicContact(emcBoundaryPos::XMAX, 0, {0}, {ly});
  device.addOhmicContact(emcBoundaryPos::XMIN, voltage, {0}, {ly});
  if (gate)
    device.addGateContact(emcBoundaryPos::YMIN, gateVoltage, {lx / 3}, {2 * lx / 3}, 3.9, 1.2e-9, 1.15 / 2);
#else
  const double lz = o.lz, hz = o.hz;
  DeviceType device{Silicon::getSiliconMaterial<T>(), {lx, ly, lz}, {hx, hy, hz}};
  device.addConstantDopingRegion({0, 0, 0}, {lx, ly, lz}, doping);
  std::vector<int> regions = {0};
  if (doping2 != 0) {
    device.addConstantDopingRegion({lx / 2, 0, 0}, {lx, ly, lz}, doping2);
    regions.push_back(1);
  }
  device.addOhmicContact(emcBoundaryPos::XMAX, 0, {0, 0}, {ly, lz});
  device.addOhmicContact(emcBoundaryPos::XMIN, voltage, {0, 0}, {ly, lz});
  if (gate)
    device.addGateContact(emcBoundaryPos::YMIN, gateVoltage, {lx / 3, 0}, {2 * lx / 3, lz}, 3.9, 1.2e-9, 1.15 / 2);
#endif

  Solver solver(device, acc, omega);
  PMScheme pmScheme;
  emcSimulationParameter<T, DeviceType> param;
  param.setTimes(steps * dt, dt, 0);
  param.setNrStepsForFinalAvg(0);
  auto electrons = makeElectron<Electron>(o);
  Silicon::addXValley(electrons);
  Silicon::addAcousticScattering(0, electrons, device, regions);
  Silicon::addZeroOrderInterValleyScattering(0, electrons, device, regions);
  Silicon::addFirstOrderInterValleyScattering(0, electrons, device, regions);
  Silicon::addCoulombScattering(0, electrons, device, regions);
  if (o.surfYminConst >= 0)
    electrons->setSurfaceScatterMechanism(
        emcBoundaryPos::YMIN,
        std::make_unique<emcConstantSurfaceScatterMechanism<T, DeviceType>>(o.surfYminConst, device.getMaxPos()));
  if (o.surfYmaxMom >= 0)
    electrons->setSurfaceScatterMechanism(
        emcBoundaryPos::YMAX,
        std::make_unique<emcMomentumDependentSurfaceScatterMechanism<T, DeviceType>>(o.surfYmaxMom, device.getMaxPos()));
  if (o.grainRate > 0)
    electrons->setGrainScatterMechanism(std::make_unique<emcGrainScatterMechanism<T>>(o.grainProb, o.grainRate));
  param.addParticleType(std::move(electrons));
  Handler handler(device, pmScheme, param.particleTypes, param.nrCarriersPerPart, seed);
  emcSimulationResults<T, DeviceType> results(device, param);
  Blob blob(out);

  const auto extent = device.getGridExtent();
  const size_t nx = extent[0], ny = extent[1];
  // ---- static device description -----------------------------------------------------------
  {
    blob.grid("doping_norm", device.getDopingProfile().getDoping(true));
    std::vector<std::int64_t> region, ohmic, reservoir, contactIdx, face, gateIdx;
    DeviceType::SizeVec c;
    auto &surf = device.getSurface();
    for (c.fill(0); !device.isEndCoord(c); device.advanceCoord(c)) {
      region.push_back(device.getDopingProfile().getDopingRegionIdx(c));
      ohmic.push_back(surf.isOhmicContact(c));
      reservoir.push_back(surf.isReservoirContact(c));
      contactIdx.push_back(surf.getContactIdx(c));
      face.push_back((std::int64_t)toUnderlying(surf.getBoundaryPos(c)));
      for (int f = 0; f < 2 * (int)Dim; f++) { // contact index per face this cell lies on (-1: none / not on the face)
        auto bp = static_cast<emcBoundaryPos>(f);
        std::int64_t id = -2;
        if (surf.isOnBoundary(c, bp))
          id = surf.getContactIdx(surf.getCoordBoundary(c, bp), bp);
        gateIdx.push_back(id);
      }
    }
    std::vector<std::uint64_t> cellDims;
    for (size_t d = Dim; d-- > 0;)
      cellDims.push_back(extent[d]);
    auto withFaces = cellDims;
    withFaces.push_back(2 * Dim);
    blob.i64("region", region, cellDims);
    blob.i64("is_ohmic", ohmic, cellDims);
    blob.i64("is_reservoir", reservoir, cellDims);
    blob.i64("contact_idx", contactIdx, cellDims);
    blob.i64("first_face", face, cellDims);
    blob.i64("face_contact", gateIdx, withFaces);
    std::vector<double> contacts; // type, voltage, epsOx, thickness, barrier
    for (size_t i = 0; i < surf.getNrContacts(); i++) {
      const bool isGate = surf.getContactType(i) == emcContactType::GATE;
      contacts.insert(contacts.end(),
                      {(double)toUnderlying(surf.getContactType(i)), surf.getContactVoltage(i),
                       isGate ? surf.getContactFurtherParameter(i, 0) : 0., isGate ? surf.getContactFurtherParameter(i, 1) : 0.,
                       isGate ? surf.getContactFurtherParameter(i, 2) : 0.});
    }
    blob.f64("contacts", contacts, {surf.getNrContacts(), 5});
    blob.grid("expected_at_contact", handler.expNrPart[0]);
    blob.f64("solver_consts", {solver.accuracy, solver.omega, solver.hFactor[0], solver.hFactor[1], solver.hFactorSum,
                               solver.hProduct, solver.h[0], solver.h[1]});
    std::vector<double> gf(solver.gateFactor.begin(), solver.gateFactor.end()),
        gv(solver.gateVoltage.begin(), solver.gateVoltage.end()),
        gb(solver.gateBarrierHeight.begin(), solver.gateBarrierHeight.end());
    blob.f64("gate_factor", gf);
    blob.f64("gate_voltage", gv);
    blob.f64("gate_barrier", gb);
    blob.f64("device_consts", {device.getThermalVoltage(), device.getDebyeLength(), device.getCellVolume(),
                               device.getMaterial().getNi(), hx, hy, lx, ly, width});
    auto &sh = param.particleTypes[0]->scatterHandler;
    for (auto &[key, tables] : sh.scatterTables) {
      std::vector<double> flat;
      for (auto &t : tables)
        flat.insert(flat.end(), t.begin(), t.end());
      std::string sfx = "_v" + std::to_string(std::get<0>(key)) + "_r" + std::to_string(std::get<1>(key));
      blob.f64("cum" + sfx, flat, {tables.size(), (std::uint64_t)levels});
      blob.f64("tau" + sfx, {sh.tau.at(key)});
    }
  }
  // ---- calcEquilibriumCharacteristics (emcSimulation.hpp:139-146) --------------------------
  blob.grid("pot_guess", results.currPot);
  solver.calcEquilibriumPotential(results.currPot, device);
  blob.grid("pot_eq", results.currPot);
  pmScheme.calcEField(results.eField, results.currPot, device);
  blob.grid("ex_eq", results.eField[0]);
  blob.grid("ey_eq", results.eField[1]);
  if (Dim > 2)
    blob.grid("ez_eq", results.eField[Dim - 1]);
  handler.generateInitialParticles(results.currPot);
  blob.u64("draws_init_count", {draws.size()});
  dumpEnsemble(blob, "init_", handler);
  results.nrPart[0].fill(0);
  handler.assignParticlesToMesh(0, results.nrPart[0]);
  results.updateCurrentParticleConcentrations(device);
  blob.grid("count_eq", results.nrPart[0]);
  blob.grid("conc_eq", results.currConc[0]);

  // ---- performEMCStep, non-FMM branch (emcSimulation.hpp:177-192) -----------------------------
  std::vector<std::uint64_t> drawMarks; // per step: before drift, after drift, after contacts
  bool resetBC = true;
  std::vector<std::int64_t> removedAll, injectedAll, sizeAll; // every step: per contact, ensemble size after the contacts
  for (int s = 0; s < steps; s++) {
    const std::string p = "s" + std::to_string(s) + "_";
    const bool snap = o.snapEvery <= 1 || s % o.snapEvery == 0 || s == steps - 1;
    solver.calcNonEquilibriumPotential(results.currPot, device, results.currConc[0], resetBC);
    resetBC = false; // emcSimulation: true only for step 0 (:107, :118-120)
    pmScheme.calcEField(results.eField, results.currPot, device);
    if (snap) {
      blob.grid(p + "pot", results.currPot);
      blob.grid(p + "ex", results.eField[0]);
      blob.grid(p + "ey", results.eField[1]);
      if (Dim > 2)
        blob.grid(p + "ez", results.eField[Dim - 1]);
    }
    // label the particles through the (dynamically inert) grain clock so that removals can be traced -- unless a grain
    // mechanism is set: then the clock is live and is recorded as it is
    if (!(o.grainRate > 0))
      for (size_t i = 0; i < handler.particles[0].size(); i++)
        handler.particles[0][i].grainTau = 1000. + i;
    if (snap)
      dumpEnsemble(blob, p + "pre_", handler);
    drawMarks.push_back(draws.size());
    auto nrRem = handler.driftScatterParticles(dt, results.eField);
    drawMarks.push_back(draws.size());
    if (snap)
      dumpEnsemble(blob, p + "drift_", handler);
    std::vector<std::int64_t> rem(nrRem[0].begin(), nrRem[0].end());
    if (snap)
      blob.i64(p + "removed_per_contact", rem);
    removedAll.insert(removedAll.end(), rem.begin(), rem.end());
    auto nrInj = handler.handleOhmicContacts();
    drawMarks.push_back(draws.size());
    if (snap)
      dumpEnsemble(blob, p + "post_", handler);
    std::vector<std::int64_t> inj(nrInj[0].begin(), nrInj[0].end());
    if (snap)
      blob.i64(p + "net_injected_per_contact", inj);
    injectedAll.insert(injectedAll.end(), inj.begin(), inj.end());
    sizeAll.push_back((std::int64_t)handler.getNrParticles(0));
    results.nrPart[0].fill(0);
    handler.assignParticlesToMesh(0, results.nrPart[0]);
    results.updateCurrentParticleConcentrations(device);
    if (snap) {
      blob.grid(p + "count", results.nrPart[0]);
      blob.grid(p + "conc", results.currConc[0]);
    }
  }
  if (o.snapEvery > 1) {
    const std::uint64_t nC = removedAll.size() / steps;
    blob.i64("removed_all", removedAll, {(std::uint64_t)steps, nC});
    blob.i64("net_injected_all", injectedAll, {(std::uint64_t)steps, nC});
    blob.i64("size_all", sizeAll, {(std::uint64_t)steps});
  }
  blob.u64("draw_marks", drawMarks);
  blob.u64("draws", draws);
  blob.f64("params", {lx, ly, hx, hy, width, doping, doping2, voltage, dt, acc, omega, (double)steps, (double)levels,
                      emax, (double)gate, gateVoltage, (double)seed, o.surfYminConst, o.surfYmaxMom});
  std::cout.rdbuf(oldBuf);
  std::cout << "ref_device_driver: grid " << nx << "x" << ny << ", " << handler.getNrParticles(0) << " particles after "
            << steps << " steps, " << draws.size() << " draws -> " << out << "\n";
  return 0;
}

int main(int argc, char **argv) {
  Options o;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--lx") o.lx = std::stod(v);
    else if (k == "--ly") o.ly = std::stod(v);
    else if (k == "--hx") o.hx = std::stod(v);
    else if (k == "--hy") o.hy = std::stod(v);
    else if (k == "--width") o.width = std::stod(v);
    else if (k == "--lz") o.lz = std::stod(v);
    else if (k == "--hz") o.hz = std::stod(v);
    else if (k == "--doping") o.doping = std::stod(v);
    else if (k == "--doping2") o.doping2 = std::stod(v); // right half of the bar, second region
    else if (k == "--voltage") o.voltage = std::stod(v);
    else if (k == "--dt") o.dt = std::stod(v);
    else if (k == "--acc") o.acc = std::stod(v);
    else if (k == "--omega") o.omega = std::stod(v);
    else if (k == "--steps") o.steps = std::stoi(v);
    else if (k == "--snap-every") o.snapEvery = std::stoi(v);
    else if (k == "--levels") o.levels = std::stoi(v);
    else if (k == "--emax") o.emax = std::stod(v);
    else if (k == "--gate") o.gate = std::stoi(v); // 1: gate contact on the middle third of YMIN
    else if (k == "--gate-voltage") o.gateVoltage = std::stod(v);
    else if (k == "--seed") o.seed = std::stoul(v);
    else if (k == "--out") o.out = v;
    else if (k == "--scheme") o.scheme = v;
    else if (k == "--electron") o.electron = v;
    else if (k == "--surface-ymin-const") o.surfYminConst = std::stod(v); // specularity parameter
    else if (k == "--surface-ymax-mom") o.surfYmaxMom = std::stod(v);     // rms roughness height [m]
    else if (k == "--grain-rate") o.grainRate = std::stod(v);             // emcGrainScatterMechanism, [1/s]
    else if (k == "--grain-prob") o.grainProb = std::stod(v);
    else {
      std::cerr << "unknown option " << k << "\n";
      return 2;
    }
  }
  using E = emcElectron<T, DeviceType>;
  using V = electronVWD<T, DeviceType>;
  const bool vwd = o.electron == "vwd";
  if (o.scheme == "ngp") return vwd ? run<emcNGPScheme<T, DeviceType>, V>(o) : run<emcNGPScheme<T, DeviceType>, E>(o);
  if (o.scheme == "cic") return vwd ? run<emcCICScheme<T, DeviceType>, V>(o) : run<emcCICScheme<T, DeviceType>, E>(o);
#if DEVICE_DIM == 2
  if (o.scheme == "nec") return vwd ? run<emcNECScheme<T, DeviceType>, V>(o) : run<emcNECScheme<T, DeviceType>, E>(o);
  if (o.scheme == "vwd") return vwd ? run<emcNECSchemeVWD<T, DeviceType>, V>(o) : run<emcNECSchemeVWD<T, DeviceType>, E>(o);
#endif
  std::cerr << "unknown scheme " << o.scheme << "\n";
  return 2;
}
