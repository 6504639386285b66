// n-doped silicon bar between two ohmic contacts, self-consistent (NGP charge assignment + nonlinear
// SOR Poisson every step) on the GPU-resident drop-in API -- the scenario of the reference's
// examples/resistor2D/resistor2D.cpp (1 um x 1 um, 101 x 21 grid, 1e22 m^-3, 50 mV, 50 000 steps of
// 1 fs by default) with its compile-time constants as options and the same result files.
//
//   resistor2D [--voltage V] [--steps K] [--transient K] [--avg K] [--dt s] [--seed S] [--doping 1/m3]
//              [--lx m] [--ly m] [--hx m] [--hy m] [--width m] [--carriers-per-particle n]
//              [--poisson-interval n] [--red-black 0|1] [--progress K] [--prefix name] [--grain-rate 1/s --grain-prob p]
//
// Prints the terminal currents, the particle-steps per second of the Monte Carlo loop and the mean
// number of SOR sweeps per step.
#include <chrono>
#include <iostream>
#include <memory>
#include <string>

#include <PMSchemes/emcNGPScheme.hpp>
#include <ParticleHandler/emcBasicParticleHandler.hpp>
#include <ParticleType/emcElectron.hpp>
#include <PoissonSolver/emcSORSolver.hpp>
#include <emcGrainScatterMechanism.hpp>
#include <emcSimulation.hpp>

#include "SiliconModel.hpp"

using NumType = double;
using DeviceType = emcDevice<NumType, 2>;
using PMScheme = emcNGPScheme<NumType, DeviceType>;
using ParticleHandler = emcBasicParticleHandler<NumType, DeviceType, PMScheme>;
using PoissonSolver = emcSORSolver<NumType, DeviceType, ParticleHandler>;
using SimulationType = emcSimulation<NumType, DeviceType, PoissonSolver, ParticleHandler, PMScheme>;

int main(int argc, char **argv) {
  double voltage = 0.05, dt = 1e-15, doping = 1e22, lx = 1e-6, ly = 1e-6, hx = 1e-8, hy = 5e-8, width = 1e-6, grainRate = 0,
         grainProb = 1;
  long steps = 50000, transient = 20000, avg = 20000, carriers = 1, poissonInterval = 1, progress = 5000, redBlack = 1;
  unsigned long seed = 0;
  bool seeded = false;
  std::string prefix = "resistor";
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string key = argv[i], val = argv[i + 1];
    if (key == "--voltage") voltage = std::stod(val);
    else if (key == "--steps") steps = std::stol(val);
    else if (key == "--transient") transient = std::stol(val);
    else if (key == "--avg") avg = std::stol(val);
    else if (key == "--dt") dt = std::stod(val);
    else if (key == "--seed") seed = std::stoul(val), seeded = true;
    else if (key == "--doping") doping = std::stod(val);
    else if (key == "--lx") lx = std::stod(val);
    else if (key == "--ly") ly = std::stod(val);
    else if (key == "--hx") hx = std::stod(val);
    else if (key == "--hy") hy = std::stod(val);
    else if (key == "--width") width = std::stod(val);
    else if (key == "--carriers-per-particle") carriers = std::stol(val);
    else if (key == "--poisson-interval") poissonInterval = std::stol(val);
    else if (key == "--progress") progress = std::stol(val);
    else if (key == "--red-black") redBlack = std::stol(val);
    else if (key == "--prefix") prefix = val;
    else if (key == "--grain-rate") grainRate = std::stod(val);
    else if (key == "--grain-prob") grainProb = std::stod(val);
    else {
      std::cerr << "unknown option " << key << "\n";
      return 2;
    }
  }
  transient = std::min(transient, steps);
  avg = std::min(avg, steps);

  DeviceType device{SiliconModel::material<NumType>(), {lx, ly}, {hx, hy}};
  device.setDeviceWidth(width);
  device.addConstantDopingRegion({0, 0}, {lx, ly}, doping);
  device.addOhmicContact(emcBoundaryPos::XMAX, 0, {0}, {ly});
  device.addOhmicContact(emcBoundaryPos::XMIN, voltage, {0}, {ly});

  PoissonSolver solver(device, 1e-4, 1.8);
  solver.setRedBlackOrdering(redBlack != 0);
  PMScheme pmScheme;
  emcSimulationParameter<NumType, DeviceType> param;
  param.setTimes((steps - 0.5) * dt, dt, transient == 0 ? 0. : (transient - 0.5) * dt);
  param.setNrCarriersPerPart(carriers);
  param.setNamePrefix(prefix);
  param.setNrStepsBetweenShowProgress(progress);
  param.setNrStepsForFinalAvg(avg);
  if (seeded)
    param.setSeed(seed);

  auto electrons = std::make_unique<emcElectron<NumType, DeviceType>>(1000, 4, false);
  SiliconModel::addXValley<NumType>(electrons);
  using namespace SiliconModel;
  addScattering<NumType>(electrons, device, {0}, ACOUSTIC | ZERO_ORDER | FIRST_ORDER | COULOMB);
  if (grainRate > 0)
    electrons->setGrainScatterMechanism(std::make_unique<emcGrainScatterMechanism<NumType>>(grainProb, grainRate));
  param.addParticleType(std::move(electrons));

  SimulationType simulation(param, device, solver, pmScheme);
  simulation.setPoissonInterval(poissonInterval);
  const auto start = std::chrono::steady_clock::now();
  simulation.execute();
  const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
  const double n = simulation.getParticleHandler().getNrParticles(0);
  std::cout << "current XMAX contact: " << simulation.getAvgCurrent(0, 0) << " A, XMIN contact: " << simulation.getAvgCurrent(0, 1)
            << " A\n"
            << "wall time " << seconds << " s, " << steps << " steps, " << n << " particles at the end, "
            << n * steps / seconds << " particle-steps/s (Monte Carlo loop alone: " << simulation.getLoopSeconds() << " s, "
            << n * steps / simulation.getLoopSeconds() << " particle-steps/s), " << double(simulation.getTotalNrSorSweeps()) / steps
            << " SOR sweeps per step\n";
  return 0;
}
