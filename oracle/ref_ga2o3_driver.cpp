// TEST INFRASTRUCTURE (oracle/). Not product code, never shipped, never timed as the product.
//
// Recorder around the UNMODIFIED reference hot-phonon bulk path:
// examples/hotPhononGa2O3/Ga2O3Functions.hpp (material, Gamma valley, acoustic + non-polar optical + polar optical
// mechanisms, phonon baths, plasmon screening) driven exactly like runOneField() of
// examples/hotPhononGa2O3/hotPhononGa2O3.cpp:152-365 (moveParticles -> observables -> screening update -> bath
// update -> reinitScatterTables), without the Pauli option.  Built like ref_bulk_driver (overlay emcUtil.hpp with the
// recording RNG, -fno-access-control, no OpenMP) into oracle/_ref/.
//
//   --polar eq|hot|screened_eq|screened_hot   emcFroehlich*3D / emcHotPhononFroehlich*3D /
//                                              emcScreenedFroehlich*3D / emcScreenedHotPhononFroehlich*3D
//   --multimode 0|1  --screening 0|1  --qresolved 0|1  --qres-angle 0|1  --acoustic-bath 0|1  --impurity 0|1
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
#include <emcDevice.hpp>

#include <basicBulkParticleHandler.hpp> // -I $(REF)/examples/bulkSimulation
#include <Ga2O3Functions.hpp>           // -I $(REF)/examples/hotPhononGa2O3

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

static void dumpEnsemble(Blob &b, const std::string &p, Handler &h) {
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
    idx.insert(idx.end(), {(std::int64_t)parts[i].valley, (std::int64_t)parts[i].subValley, (std::int64_t)parts[i].region});
  }
  b.f64(p + "k", k, {n, 3});
  b.f64(p + "pos", x, {n, 3});
  b.f64(p + "energy", e);
  b.f64(p + "tau", tau);
  b.f64(p + "graintau", g);
  b.i64(p + "idx", idx, {n, 3});
}

static void dumpTables(Blob &blob, const std::string &p, emcElectron<T, DeviceType> &type, int levels) {
  auto &sh = type.scatterHandler;
  for (auto &[key, tables] : sh.scatterTables) {
    std::vector<double> flat;
    for (auto &t : tables)
      flat.insert(flat.end(), t.begin(), t.end());
    std::string sfx = "_v" + std::to_string(std::get<0>(key)) + "_r" + std::to_string(std::get<1>(key));
    blob.f64(p + "cum" + sfx, flat, {tables.size(), (std::uint64_t)levels});
    blob.f64(p + "tau" + sfx, {sh.tau.at(key)});
  }
}

int main(int argc, char **argv) {
  std::string out = "ga2o3.blob", polar = "screened_hot";
  double box = 3e-7, doping = 1e23, field = 2e7, dt = 1e-16, temperature = 300, tauLO = Ga2O3::tauLODefault,
         tauAc = Ga2O3::tauAcDefault, emax = 5.0, screenEps = Ga2O3::epsLo;
  int steps = 200, levels = 2000, multimode = 0, screening = 0, qresolved = 0, qresAngle = 1, acousticBath = 1, impurity = 0,
      reinitEvery = 1;
  unsigned long seed = 1;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--out") out = v;
    else if (k == "--polar") polar = v;
    else if (k == "--box") box = std::stod(v);
    else if (k == "--doping") doping = std::stod(v);
    else if (k == "--field") field = std::stod(v);
    else if (k == "--dt") dt = std::stod(v);
    else if (k == "--temperature") temperature = std::stod(v);
    else if (k == "--tau-lo") tauLO = std::stod(v);
    else if (k == "--tau-ac") tauAc = std::stod(v);
    else if (k == "--emax") emax = std::stod(v);
    else if (k == "--steps") steps = std::stoi(v);
    else if (k == "--levels") levels = std::stoi(v);
    else if (k == "--multimode") multimode = std::stoi(v);
    else if (k == "--screening") screening = std::stoi(v);
    else if (k == "--qresolved") qresolved = std::stoi(v);
    else if (k == "--qres-angle") qresAngle = std::stoi(v);
    else if (k == "--acoustic-bath") acousticBath = std::stoi(v);
    else if (k == "--impurity") impurity = std::stoi(v);
    else if (k == "--reinit-every") reinitEvery = std::stoi(v);
    else if (k == "--seed") seed = std::stoul(v);
    else {
      std::cerr << "unknown option " << k << "\n";
      return 2;
    }
  }
  std::vector<std::uint64_t> draws;
  RecordingRNG::sink() = &draws;
  std::streambuf *oldBuf = std::cout.rdbuf();
  std::ostringstream sink;
  std::cout.rdbuf(sink.rdbuf());

  const std::array<T, 3> maxPos = {box, box, box};
  const T h = box / 2.;
  const T Vsim = box * box * box;
  DeviceType device{Ga2O3::getGa2O3Material<T>(), maxPos, {h, h, h}, temperature};
  device.addConstantDopingRegion({0, 0, 0}, maxPos, doping);
  TypeMap types;
  types[0] = std::make_unique<emcElectron<T, DeviceType>>(levels, emax, false);
  Ga2O3::addGammaValley(types[0], Ga2O3::alpha);
  Ga2O3::addAcousticScattering(0, types[0], device, {0}, Ga2O3::sigmaAc);
  Ga2O3::addNonPolarOpticalScattering(0, types[0], device, {0});
  if (impurity)
    Ga2O3::addImpurityScattering(0, types[0], device, {0});
  const std::vector<T> modeEnergy = multimode ? Ga2O3::modeEnergyMulti : std::vector<T>{Ga2O3::phononEnergyPOP};
  const std::vector<T> modeEpsLo = multimode ? Ga2O3::modeEpsLoMulti() : std::vector<T>{Ga2O3::epsLo};
  auto screen = Ga2O3::makeScreening(screening != 0, screenEps);
  screen->update(doping, temperature);
  std::vector<std::shared_ptr<Ga2O3::PhononBath>> baths;
  const bool hot = polar == "hot" || polar == "screened_hot";
  if (hot) {
    baths = Ga2O3::makePhononBaths(modeEnergy, tauLO, temperature, Vsim, acousticBath != 0, tauAc);
    for (auto &b : baths)
      b->setScreeningQ2(screen->getQs2());
  }
  if (polar == "screened_hot")
    Ga2O3::addScreenedHotPolarOpticalScattering(0, types[0], {0}, modeEnergy, modeEpsLo, baths, screen, qresolved != 0,
                                                qresAngle != 0);
  else if (polar == "screened_eq")
    Ga2O3::addScreenedEquilibriumPolarOpticalScattering(0, types[0], {0}, temperature, modeEnergy, modeEpsLo, screen);
  else if (polar == "hot")
    Ga2O3::addHotPolarOpticalScattering(0, types[0], {0}, modeEnergy, baths);
  else if (polar == "eq")
    for (auto e : modeEnergy)
      Ga2O3::addEquilibriumPolarOpticalScattering(0, types[0], {0}, temperature, e);
  else {
    std::cerr << "unknown --polar\n";
    return 2;
  }

  Handler handler(device, types, {-1, 0, 0});
  handler.setSeed(seed);
  handler.resetAppliedFieldStrength(field);
  Blob blob(out);
  auto &electron = static_cast<emcElectron<T, DeviceType> &>(*types[0]);
  dumpTables(blob, "init_", electron, levels);
  handler.generateInitialParticles();
  blob.u64("draws_init_count", {draws.size()});
  dumpEnsemble(blob, "init_", handler);

  std::vector<double> obs, meanNq, qs2, tauSeries, tAc;
  std::vector<double> counts; // [steps][baths][2][bins] before update
  std::vector<std::uint64_t> drawCount;
  for (int step = 1; step <= steps; step++) {
    handler.moveParticles(dt);
    drawCount.push_back(draws.size());
    const T v = handler.getAvgDriftVelocity(0)[0];
    const T e = handler.getAvgEnergy(0)[0];
    obs.push_back(e);
    obs.push_back(v);
    bool stale = false;
    if (screening) {
      const T Te = 2. * e * constants::q / (3. * constants::kB);
      screen->update(doping, Te);
      for (auto &b : baths)
        b->setScreeningQ2(screen->getQs2());
      stale = true;
    }
    qs2.push_back(screen->getQs2());
    if (hot) {
      for (auto &b : baths) {
        counts.insert(counts.end(), b->nEm.begin(), b->nEm.end());
        counts.insert(counts.end(), b->nAbs.begin(), b->nAbs.end());
        b->update(dt);
      }
      stale = true;
    }
    if (stale && step % reinitEvery == 0)
      types[0]->reinitScatterTables();
    for (auto &b : baths)
      meanNq.push_back(b->getMeanNq());
    tAc.push_back(hot ? baths[0]->getAcousticTemp() : temperature);
    tauSeries.push_back(types[0]->getTau(0, 0));
  }
  dumpEnsemble(blob, "final_", handler);
  dumpTables(blob, "final_", electron, levels);
  const std::uint64_t nB = baths.size(), nBins = hot ? baths[0]->nrBins : 0;
  blob.f64("obs", obs, {(std::uint64_t)steps, 2});
  blob.f64("qs2", qs2);
  blob.f64("tau_series", tauSeries);
  blob.f64("t_acoustic", tAc);
  if (hot) {
    blob.f64("mean_nq", meanNq, {(std::uint64_t)steps, nB});
    blob.f64("bath_counts", counts, {(std::uint64_t)steps, nB, 2, nBins});
    std::vector<double> nq, cw, cwn;
    for (auto &b : baths) {
      nq.insert(nq.end(), b->Nq.begin(), b->Nq.end());
      cw.insert(cw.end(), b->cumW.begin(), b->cumW.end());
      cwn.insert(cwn.end(), b->cumWN.begin(), b->cumWN.end());
    }
    blob.f64("final_nq", nq, {nB, nBins});
    blob.f64("final_cumw", cw, {nB, nBins + 1});
    blob.f64("final_cumwn", cwn, {nB, nBins + 1});
  }
  blob.u64("draw_count_after_step", drawCount);
  blob.u64("draws", draws);
  blob.f64("mode_energy", modeEnergy);
  blob.f64("mode_eps_lo", modeEpsLo);
  blob.f64("params", {box, doping, field, dt, temperature, tauLO, tauAc, emax, (double)steps, (double)levels, (double)multimode,
                      (double)screening, (double)qresolved, (double)qresAngle, (double)acousticBath, (double)impurity,
                      (double)reinitEvery, (double)seed, screenEps, (double)handler.getNrParticles(0)});
  std::cout.rdbuf(oldBuf);
  std::cout << "ref_ga2o3_driver: " << handler.getNrParticles(0) << " particles, " << steps << " steps, " << draws.size()
            << " draws -> " << out << "\n";
  return 0;
}
