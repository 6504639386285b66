// TEST INFRASTRUCTURE (oracle/). Not product code, never shipped, never timed as the product.
//
// Recorder around the UNMODIFIED reference classes as examples/hotCarrierMHP/hotCarrierMHP.cpp uses them for its
// data-parallel part: TWO moved species -- emcElectron and emcHole, one non-parabolic isotropic valley
// each, mono-energetic photo-excited start, no field -- coupled through SHARED LO phonon bath(s) and a shared plasmon
// screening object; per step (hotCarrierMHP.cpp:655-705, with carrier-carrier scattering, recombination, energy-selective
// contacts and band filling switched off):  moveParticles (electrons, then holes) -> bath update -> screening from the
// live density and temperature of both species -> reinitScatterTables of both.  Material presets and the mechanism
// construction follow hotCarrierMHP.cpp:166-470 (MAPbI3).  Built like ref_ga2o3_driver into oracle/_ref/.
//
//   --polar hot|screened_hot|screened_eq   --qresolved 0|1   --screening 0|1   --acoustic-bath 0|1
#include <cmath>
#include <cstdint>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <ParticleType/emcElectron.hpp>
#include <ParticleType/emcHole.hpp>
#include <ScatterMechanisms/emcHotPhononFroehlichMechanism.hpp>
#include <ScatterMechanisms/emcScreenedFroehlichInteraction.hpp>
#include <ValleyTypes/emcNonParabolicIsotropValley.hpp>
#include <emcDevice.hpp>
#include <emcPhononBath.hpp>
#include <emcPlasmonScreening.hpp>

#include <basicBulkParticleHandler.hpp> // -I $(REF)/examples/bulkSimulation

using T = double;
using DeviceType = emcDevice<T, 3>;
using Handler = basicBulkParticleHandler<T, DeviceType>;
using TypeMap = Handler::MapIdxToParticleTypes;

struct Blob {
  std::ofstream os;
  explicit Blob(const std::string &path) : os(path, std::ios::binary) {}
  void put(const std::string &name, char dtype, const void *data, const std::vector<std::uint64_t> &dims, size_t elemSize) {
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
};

static void dumpEnsemble(Blob &b, const std::string &p, Handler &h, size_t type) {
  const auto &parts = h.particles[type];
  const auto &pos = h.positionsParticles[type];
  const size_t n = parts.size();
  std::vector<double> k, x, e, tau, g;
  std::vector<std::int64_t> idx;
  for (size_t i = 0; i < n; i++) {
    k.insert(k.end(), parts[i].k.begin(), parts[i].k.end());
    x.insert(x.end(), pos[i].begin(), pos[i].end());
    e.push_back(parts[i].energy);
    tau.push_back(parts[i].tau);
    g.push_back(parts[i].grainTau);
    idx.insert(idx.end(), {(std::int64_t)parts[i].valley, (std::int64_t)parts[i].subValley, (std::int64_t)parts[i].region});
  }
  b.f64(p + "k", k, {n, 3});
  b.f64(p + "pos", x, {n, 3});
  b.f64(p + "energy", e);
  b.f64(p + "tau", tau);
  b.f64(p + "graintau", g);
  b.i64(p + "idx", idx, {n, 3});
}

static void dumpTables(Blob &blob, const std::string &p, emcParticleType<T, DeviceType> &type, int levels) {
  auto &sh = type.scatterHandler;
  for (auto &[key, tables] : sh.scatterTables) {
    std::vector<double> flat;
    for (auto &t : tables)
      flat.insert(flat.end(), t.begin(), t.end());
    blob.f64(p + "tab_cum", flat, {tables.size(), (std::uint64_t)levels});
    blob.f64(p + "tab_tau", {sh.tau.at(key)});
  }
}

int main(int argc, char **argv) {
  std::string out = "mhp.blob", polar = "screened_hot";
  // MAPbI3 preset of hotCarrierMHP.cpp:190 and its defaults
  double box = 8e-8, density = 1e24, dt = 5e-15, temperature = 300, tauLO = 0.6e-12, tauAc = 30e-12, emax = 4.0, epsHi = 5.0,
         epsLo = 33.5, massE = 0.20, massH = 0.25, hwLO = 0.0115, gap = 1.60, photon = 3.1, dq = 5e7, alphaE = -1, alphaH = -1;
  int steps = 60, levels = 1000, screening = 1, qresolved = 1, acousticBath = 1, bins = 40;
  unsigned long seed = 1;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--out") out = v;
    else if (k == "--polar") polar = v;
    else if (k == "--box") box = std::stod(v);
    else if (k == "--density") density = std::stod(v);
    else if (k == "--dt") dt = std::stod(v);
    else if (k == "--steps") steps = std::stoi(v);
    else if (k == "--levels") levels = std::stoi(v);
    else if (k == "--emax") emax = std::stod(v);
    else if (k == "--screening") screening = std::stoi(v);
    else if (k == "--qresolved") qresolved = std::stoi(v);
    else if (k == "--acoustic-bath") acousticBath = std::stoi(v);
    else if (k == "--bins") bins = std::stoi(v);
    else if (k == "--dq") dq = std::stod(v);
    else if (k == "--alpha-e") alphaE = std::stod(v);
    else if (k == "--alpha-h") alphaH = std::stod(v);
    else if (k == "--seed") seed = std::stoul(v);
    else {
      std::cerr << "unknown option " << k << "\n";
      return 2;
    }
  }
  // two-band Kane estimate per species (hotCarrierMHP.cpp:262-273)
  if (alphaE < 0) alphaE = (1. / gap) * (1. - massE) * (1. - massE);
  if (alphaH < 0) alphaH = (1. / gap) * (1. - massH) * (1. - massH);
  std::vector<std::uint64_t> draws;
  RecordingRNG::sink() = &draws;
  std::streambuf *oldBuf = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());

  const T Vsim = box * box * box;
  const std::array<T, 3> maxPos = {box, box, box};
  const std::array<T, 3> spacing = {box / 10, box / 10, box / 10};
  emcMaterial<T> material(epsLo, 4000., 1e10, 2000., gap);
  const T excess = photon - material.getBandGap();
  const T fh = massH / (massE + massH), fe = massE / (massE + massH);
  const T Ee = fh * excess, Eh = fe * excess;
  DeviceType device{material, maxPos, spacing, temperature};
  device.addConstantDopingRegion({0, 0, 0}, maxPos, density);
  TypeMap types;
  types[0] = std::make_unique<emcElectron<T, DeviceType>>(levels, emax, false, Ee);
  types[0]->addValley(std::make_unique<emcNonParabolicIsotropValley<T>>(massE, constants::me, 1, alphaE));
  types[1] = std::make_unique<emcHole<T, DeviceType>>(levels, emax, false, Eh);
  types[1]->addValley(std::make_unique<emcNonParabolicIsotropValley<T>>(massH, constants::me, 1, alphaH));
  const bool hot = polar == "hot" || polar == "screened_hot";
  auto screen = std::make_shared<emcPlasmonScreening<T>>(epsHi, screening != 0);
  std::vector<std::shared_ptr<emcPhononBath<T>>> baths;
  if (hot)
    baths.push_back(std::make_shared<emcPhononBath<T>>(bins, dq, tauLO, hwLO, temperature, Vsim, acousticBath != 0, hwLO / 2., tauAc,
                                                       0., 0.00781, tauAc));
  for (int p = 0; p < 2; p++) {
    const T mass = p == 0 ? massE : massH;
    const std::string tag = std::string(p == 0 ? "MHP-e" : "MHP-h") + "0";
    if (polar == "screened_hot") {
      types[p]->addScatterMechanism({0}, std::make_unique<emcScreenedHotPhononFroehlichAbsorption3D<T>>(
                                             0, hwLO, mass, epsHi, epsLo, baths[0], screen, qresolved != 0, tag));
      types[p]->addScatterMechanism({0}, std::make_unique<emcScreenedHotPhononFroehlichEmission3D<T>>(
                                             0, hwLO, mass, epsHi, epsLo, baths[0], screen, qresolved != 0, tag));
    } else if (polar == "hot") {
      types[p]->addScatterMechanism({0}, std::make_unique<emcHotPhononFroehlichAbsorption3D<T>>(0, hwLO, mass, epsHi, epsLo, baths[0], tag));
      types[p]->addScatterMechanism({0}, std::make_unique<emcHotPhononFroehlichEmission3D<T>>(0, hwLO, mass, epsHi, epsLo, baths[0], tag));
    } else if (polar == "screened_eq") {
      types[p]->addScatterMechanism({0}, std::make_unique<emcScreenedFroehlichAbsorption3D<T>>(0, hwLO, mass, epsHi, epsLo, temperature,
                                                                                              screen, tag));
      types[p]->addScatterMechanism({0}, std::make_unique<emcScreenedFroehlichEmission3D<T>>(0, hwLO, mass, epsHi, epsLo, temperature,
                                                                                            screen, tag));
    } else {
      std::cerr << "unknown --polar\n";
      return 2;
    }
  }
  std::array<T, 3> noField = {0., 0., 0.};
  Handler handler(device, types, noField, 0., seed);
  Blob blob(out);
  dumpTables(blob, "init_e_", *types[0], levels);
  dumpTables(blob, "init_h_", *types[1], levels);
  handler.generateInitialParticles();
  blob.u64("draws_init_count", {draws.size()});
  dumpEnsemble(blob, "init_e_", handler, 0);
  dumpEnsemble(blob, "init_h_", handler, 1);

  // hotCarrierMHP.cpp:563-582
  auto updateScreening = [&]() {
    if (!screening)
      return;
    T qs2 = 0;
    for (SizeType p = 0; p < 2; ++p) {
      const SizeType nrPart = handler.getNrParticles(p);
      if (nrPart == 0)
        continue;
      const T nS = static_cast<T>(nrPart) / Vsim;
      const T tS = T(2) * handler.getAvgEnergy(p)[0] * T(constants::q) / (T(3) * T(constants::kB)); // getMBTemp
      if (tS <= T(0))
        continue;
      qs2 += nS * constants::q * constants::q / (epsHi * constants::eps0 * constants::kB * tS);
    }
    screen->setQs2(qs2);
    for (auto &bath : baths)
      bath->setScreeningQ2(qs2);
  };
  updateScreening(); // hotCarrierMHP.cpp:632-638: screening of the photo-excited ensemble before the first step
  if (screening) {
    types[0]->reinitScatterTables();
    types[1]->reinitScatterTables();
  }
  dumpTables(blob, "start_e_", *types[0], levels);
  dumpTables(blob, "start_h_", *types[1], levels);

  std::vector<double> obs, meanNq, qs2Series, tauSeries, counts;
  std::vector<std::uint64_t> drawCount;
  for (int step = 1; step <= steps; step++) {
    handler.moveParticles(dt);
    drawCount.push_back(draws.size());
    for (int p = 0; p < 2; p++) {
      obs.push_back(handler.getAvgEnergy(p)[0]);
      obs.push_back(handler.getAvgDriftVelocity(p)[0]);
    }
    for (auto &b : baths) {
      counts.insert(counts.end(), b->nEm.begin(), b->nEm.end());
      counts.insert(counts.end(), b->nAbs.begin(), b->nAbs.end());
      b->update(dt);
    }
    updateScreening();
    if (hot || screening) {
      types[0]->reinitScatterTables();
      types[1]->reinitScatterTables();
    }
    qs2Series.push_back(screen->getQs2());
    for (auto &b : baths)
      meanNq.push_back(b->getMeanNq());
    tauSeries.push_back(types[0]->getTau(0, 0));
    tauSeries.push_back(types[1]->getTau(0, 0));
  }
  dumpEnsemble(blob, "final_e_", handler, 0);
  dumpEnsemble(blob, "final_h_", handler, 1);
  dumpTables(blob, "final_e_", *types[0], levels);
  dumpTables(blob, "final_h_", *types[1], levels);
  const std::uint64_t nB = baths.size(), nBins = hot ? baths[0]->nrBins : 0;
  blob.f64("obs", obs, {(std::uint64_t)steps, 2, 2});
  blob.f64("qs2", qs2Series);
  blob.f64("tau_series", tauSeries, {(std::uint64_t)steps, 2});
  if (hot) {
    blob.f64("mean_nq", meanNq, {(std::uint64_t)steps, nB});
    blob.f64("bath_counts", counts, {(std::uint64_t)steps, nB, 2, nBins});
    std::vector<double> nq;
    for (auto &b : baths)
      nq.insert(nq.end(), b->Nq.begin(), b->Nq.end());
    blob.f64("final_nq", nq, {nB, nBins});
  }
  blob.u64("draw_count_after_step", drawCount);
  blob.u64("draws", draws);
  blob.f64("params", {box, density, dt, temperature, tauLO, tauAc, emax, epsHi, epsLo, massE, massH, hwLO, gap, photon, dq, alphaE,
                      alphaH, Ee, Eh, (double)steps, (double)levels, (double)screening, (double)qresolved, (double)acousticBath,
                      (double)bins, (double)seed, (double)handler.getNrParticles(0), (double)handler.getNrParticles(1)});
  std::cout.rdbuf(oldBuf);
  std::cout << "ref_mhp_driver: " << handler.getNrParticles(0) << " electrons, " << handler.getNrParticles(1) << " holes, " << steps
            << " steps, " << draws.size() << " draws -> " << out << "\n";
  return 0;
}
