// TEST / BASELINE INFRASTRUCTURE (oracle/).  Timing driver around the UNMODIFIED
// reference bulk path: basicBulkParticleHandler::moveParticles followed by the three
// observable passes, exactly as examples/bulkSimulation/bulkSimulation.cpp:150-157
// drives them, with the reference's OpenMP parallelisation.  Built only where
// /root/reference exists (oracle/Makefile -> oracle/_ref/ref_bulk_bench); the binary
// travels to the GPU box and is what `bench.py --impl reference` and the
// cpu_baseline leg execute.  Prints one JSON object.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

#include <SiliconFunctions.hpp> // -I $(REF)/examples
#include <basicBulkParticleHandler.hpp>

#include <ParticleType/emcElectron.hpp>
#include <emcDevice.hpp>

using DeviceType = emcDevice<double, 3>;
using Handler = basicBulkParticleHandler<double, DeviceType>;

int main(int argc, char **argv) {
  long nTarget = 100000;
  int steps = 100, warmup = 20, threads = 1;
  double field = 1e6, dt = 1e-16;
  unsigned long seed = 12345;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string k = argv[i], v = argv[i + 1];
    if (k == "--particles") nTarget = std::stol(v);
    else if (k == "--steps") steps = std::stoi(v);
    else if (k == "--warmup") warmup = std::stoi(v);
    else if (k == "--threads") threads = std::stoi(v);
    else if (k == "--field") field = std::stod(v);
    else if (k == "--dt") dt = std::stod(v);
    else if (k == "--seed") seed = std::stoul(v);
  }
#ifdef _OPENMP
  omp_set_num_threads(threads);
#else
  threads = 1;
#endif
  // doping 1e23 m^-3 as shipped (bulkSimulation.cpp:33); box chosen for the particle count
  const double doping = 1e23;
  const double box = std::cbrt(nTarget / doping);
  const int cells = 5;
  const double h = box / cells;
  std::streambuf *old = std::cout.rdbuf();
  std::ostringstream sink; // the reference prints table info; keep stdout = one JSON line
  std::cout.rdbuf(sink.rdbuf());
  DeviceType device{Silicon::getSiliconMaterial<double>(), {box, box, box}, {h, h, h}, 300};
  device.addConstantDopingRegion({0, 0, 0}, {box, box, box}, doping);
  Handler::MapIdxToParticleTypes types;
  types[0] = std::make_unique<emcElectron<double, DeviceType>>(1000, 1., false);
  Silicon::addXValley(types[0]);
  Silicon::addAcousticScattering(0, types[0], device, {0});
  Silicon::addZeroOrderInterValleyScattering(0, types[0], device, {0});
  Silicon::addFirstOrderInterValleyScattering(0, types[0], device, {0});
  Handler handler(device, types, {-1, 0, 0}, field, seed);
  handler.generateInitialParticles();
  std::cout.rdbuf(old);
  const long n = handler.getNrParticles(0);

  double sink2 = 0;
  auto oneStep = [&](double &tMove, double &tObs) {
    auto t0 = std::chrono::steady_clock::now();
    handler.moveParticles(dt);
    auto t1 = std::chrono::steady_clock::now();
    auto e = handler.getAvgEnergy(0);
    auto v = handler.getAvgDriftVelocity(0);
    auto o = handler.getValleyOccupationProbability(0);
    auto t2 = std::chrono::steady_clock::now();
    sink2 += e[0] + v[0] + o[0];
    tMove += std::chrono::duration<double>(t1 - t0).count();
    tObs += std::chrono::duration<double>(t2 - t1).count();
  };
  double wm = 0, wo = 0;
  for (int s = 0; s < warmup; s++) oneStep(wm, wo);
  double tMove = 0, tObs = 0;
  for (int s = 0; s < steps; s++) oneStep(tMove, tObs);
  auto e = handler.getAvgEnergy(0);
  auto v = handler.getAvgDriftVelocity(0);
  std::printf("{\"particles\": %ld, \"steps\": %d, \"warmup\": %d, \"threads\": %d, \"move_s\": %.6f, "
              "\"obs_s\": %.6f, \"psteps_per_s\": %.6e, \"psteps_per_s_move_only\": %.6e, "
              "\"avg_energy\": %.8e, \"avg_drift_velocity\": %.8e, \"checksum\": %.6e}\n",
              n, steps, warmup, threads, tMove, tObs, (double)n * steps / (tMove + tObs),
              (double)n * steps / tMove, e[0], v[0], sink2);
  return 0;
}
