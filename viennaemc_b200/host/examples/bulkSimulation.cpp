// Bulk silicon under a uniform field on the GPU-resident drop-in API -- the scenario of
// the reference's examples/bulkSimulation/bulkSimulation.cpp (12 500 electrons, 10 kV/cm,
// 40 000 steps of 0.1 fs by default) with its compile-time constants as options, and the
// same three result files (time, value per valley):
//   <prefix>AvgEnergy.txt  <prefix>AvgDriftVelocity.txt  <prefix>valleyOccupation.txt
//
//   bulkSimulation [--particles N] [--field V/m] [--steps K] [--dt s] [--seed S]
//                  [--steps-per-launch L] [--prefix name] [--temperature K] [--doping 1/m3]
//                  [--grain-rate 1/s --grain-prob p]   (grain-boundary scattering, emcGrainScatterMechanism)
//                  [--lookahead N]   time steps a moveParticles(dt) call runs ahead on the device (default 16, 1 = none)
//                  [--print-at S]    write the ensemble after step S (handler.print, "<prefix>Electrons<S>.txt")
//                  [--velocities 1|3] after every step one line with v.E_dir (1) or v (3) of every particle
//                                    (handler.printDriftVelocities / printVelocities -> "<prefix>Velocities.txt", the input of
//                                    examples/singleLayerMoS2/calcMobilityFromVACF.py)
//
// --steps-per-launch > 1 uses the handler's fused entry point (several time steps per kernel
// launch, particle state kept in registers in between); 1 is the reference's call pattern
// (moveParticles(dt) followed by the three observable getters).
#include <chrono>
#include <cmath>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>

#include <ParticleType/emcElectron.hpp>
#include <basicBulkParticleHandler.hpp>
#include <emcDevice.hpp>

#include "SiliconModel.hpp"

using NumType = double;
using DeviceType = emcDevice<NumType, 3>;
using ParticleHandler = basicBulkParticleHandler<NumType, DeviceType>;

int main(int argc, char **argv) {
  double particles = 12500, field = 1e6, dt = 1e-16, temperature = 300, doping = 1e23, grainRate = 0, grainProb = 0.5;
  long steps = 40000, stepsPerLaunch = 1, lookahead = 0, printAt = -1, velocities = 0;
  unsigned long seed = 0;
  std::string prefix = "bulkSimulation";
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string key = argv[i], val = argv[i + 1];
    if (key == "--particles") particles = std::stod(val);
    else if (key == "--field") field = std::stod(val);
    else if (key == "--steps") steps = std::stol(val);
    else if (key == "--dt") dt = std::stod(val);
    else if (key == "--seed") seed = std::stoul(val);
    else if (key == "--steps-per-launch") stepsPerLaunch = std::stol(val);
    else if (key == "--prefix") prefix = val;
    else if (key == "--lookahead") lookahead = std::stol(val);
    else if (key == "--print-at") printAt = std::stol(val);
    else if (key == "--velocities") velocities = std::stol(val);
    else if (key == "--temperature") temperature = std::stod(val);
    else if (key == "--doping") doping = std::stod(val);
    else if (key == "--grain-rate") grainRate = std::stod(val);
    else if (key == "--grain-prob") grainProb = std::stod(val);
    else {
      std::cerr << "unknown option " << key << "\n";
      return 2;
    }
  }
  // cubic box holding the requested number of electrons at the given doping, 5 cells per edge
  const NumType edge = std::cbrt(particles / doping);
  const std::array<NumType, 3> maxPos = {edge, edge, edge}, spacing = {edge / 5, edge / 5, edge / 5};

  DeviceType device{SiliconModel::material<NumType>(), maxPos, spacing, temperature};
  device.addConstantDopingRegion({0, 0, 0}, maxPos, doping);

  ParticleHandler::MapIdxToParticleTypes particleTypes;
  particleTypes[0] = std::make_unique<emcElectron<NumType, DeviceType>>(1000, 1., false);
  SiliconModel::addXValley<NumType>(particleTypes[0]);
  SiliconModel::addScattering<NumType>(particleTypes[0], device, {0},
                                       SiliconModel::ACOUSTIC | SiliconModel::ZERO_ORDER | SiliconModel::FIRST_ORDER);

  if (grainRate > 0)
    particleTypes[0]->setGrainScatterMechanism(std::make_unique<emcGrainScatterMechanism<NumType>>(grainProb, grainRate));

  ParticleHandler handler(device, particleTypes, {-1, 0, 0}, field, seed);
  if (lookahead > 0)
    handler.setLookahead(lookahead);
  std::cout << "Creating Particles...\n";
  handler.generateInitialParticles();
  handler.printNrParticles();

  std::vector<std::vector<NumType>> avgEnergy(steps + 1), avgDriftVel(steps + 1), valleyOcc(steps + 1);
  avgEnergy[0] = handler.getAvgEnergy(0);
  avgDriftVel[0] = handler.getAvgDriftVelocity(0);
  valleyOcc[0] = handler.getValleyOccupationProbability(0);

  std::ofstream velFileOut;
  if (velocities)
    velFileOut.open(prefix + "Velocities.txt");
  std::cout << "Starting Simulation ...\n";
  const auto start = std::chrono::high_resolution_clock::now();
  if (stepsPerLaunch <= 1) {
    for (long s = 1; s <= steps; s++) {
      handler.moveParticles(dt);
      avgEnergy[s] = handler.getAvgEnergy(0);
      avgDriftVel[s] = handler.getAvgDriftVelocity(0);
      valleyOcc[s] = handler.getValleyOccupationProbability(0);
      if (s == printAt)
        handler.print(prefix, std::to_string(s));
      if (velocities == 1)
        handler.printDriftVelocities(velFileOut);
      else if (velocities == 3)
        handler.printVelocities(velFileOut);
    }
  } else {
    std::vector<double> series;
    handler.moveParticles(dt, steps, stepsPerLaunch, 0, series);
    const SizeType nV = particleTypes[0]->getNrValleys();
    const double total = handler.getNrParticles(0);
    for (long s = 1; s <= steps; s++) {
      for (SizeType v = 0; v < nV; v++) {
        const double *o = &series[((s - 1) * nV + v) * 3];
        avgEnergy[s].push_back(o[2] ? o[0] / o[2] : 0.);
        avgDriftVel[s].push_back(o[2] ? o[1] / o[2] : 0.);
        valleyOcc[s].push_back(o[2] / total);
      }
    }
  }
  const double seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - start).count();
  std::cout << "Wall time: " << seconds << " s  (" << handler.getNrParticles(0) * double(steps) / seconds
            << " particle-steps/s)\n";

  std::ofstream energyFile(prefix + "AvgEnergy.txt"), velFile(prefix + "AvgDriftVelocity.txt"),
      occFile(prefix + "valleyOccupation.txt");
  for (long s = 0; s <= steps; s++) {
    energyFile << s * dt << " " << avgEnergy[s] << "\n";
    velFile << s * dt << " " << avgDriftVel[s] << "\n";
    occFile << s * dt << " " << valleyOcc[s] << "\n";
  }
  handler.deleteParticles();
  return 0;
}
