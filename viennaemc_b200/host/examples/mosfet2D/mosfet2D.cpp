// n-channel MOSFET (the device of ViennaWD's example 1), self-consistent on the GPU-resident drop-in API -- the
// scenario of the reference's examples/mosfet2D/mosfet2D.cpp (125 nm x 100 nm, 1 nm grid, source / drain 5e25 m^-3,
// p-type bulk, 1.2 nm gate oxide, Vd = Vg = 1 V, 66 667 steps of 0.15 fs by default) with its compile-time constants as
// options and the same result files.  NEC-VWD particle-mesh scheme, electronVWD particle type, optional rough
// Si/SiO2 interface (constant specularity).
//
//   mosfet2D [--vd V] [--vg V] [--steps K] [--transient K] [--avg K] [--dt s] [--seed S] [--roughness p]
//            [--red-black 0|1] [--poisson-interval n] [--progress K] [--prefix name]
#include <algorithm>
#include <chrono>
#include <iostream>
#include <memory>
#include <numeric>
#include <string>

#include <ParticleHandler/emcBasicParticleHandler.hpp>
#include <PoissonSolver/emcSORSolver.hpp>
#include <SurfaceScatterMechanisms/emcConstantSurfaceScatterMechanism.hpp>
#include <emcDevice.hpp>
#include <emcSimulation.hpp>

#include "../SiliconModel.hpp"
#include "NECSchemeVWD.hpp"
#include "electronVWD.hpp"

using NumType = double;
using DeviceType = emcDevice<NumType, 2>;
using PMScheme = emcNECSchemeVWD<NumType, DeviceType>;
using ParticleHandler = emcBasicParticleHandler<NumType, DeviceType, PMScheme>;
using PoissonSolver = emcSORSolver<NumType, DeviceType, ParticleHandler>;
using SimulationType = emcSimulation<NumType, DeviceType, PoissonSolver, ParticleHandler, PMScheme>;

// potential as ViennaWD writes it: mid-gap reference, sign flipped
NumType adaptPotential(const NumType &pot, const DeviceType &device) {
  return device.getMaterial().getBandGap() / 2 - pot * device.getThermalVoltage();
}

int main(int argc, char **argv) {
  double vd = 1., vg = 1., dt = 1.5e-16, roughness = -1;
  long steps = 66667, transient = 33334, avg = 6667, poissonInterval = 1, progress = 1000, redBlack = 1;
  unsigned long seed = 0;
  bool seeded = false;
  std::string prefix = "mosfet";
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string key = argv[i], val = argv[i + 1];
    if (key == "--vd") vd = std::stod(val);
    else if (key == "--vg") vg = std::stod(val);
    else if (key == "--steps") steps = std::stol(val);
    else if (key == "--transient") transient = std::stol(val);
    else if (key == "--avg") avg = std::stol(val);
    else if (key == "--dt") dt = std::stod(val);
    else if (key == "--seed") seed = std::stoul(val), seeded = true;
    else if (key == "--roughness") roughness = std::stod(val); // specularity of the Si/SiO2 interface, < 0: smooth
    else if (key == "--red-black") redBlack = std::stol(val);
    else if (key == "--poisson-interval") poissonInterval = std::stol(val);
    else if (key == "--progress") progress = std::stol(val);
    else if (key == "--prefix") prefix = val;
    else {
      std::cerr << "unknown option " << key << "\n";
      return 2;
    }
  }
  transient = std::min(transient, steps);
  avg = std::min(avg, steps);

  DeviceType device{SiliconModel::material<NumType>(), {125e-9, 100e-9}, {1e-9, 1e-9}};
  device.setDeviceWidth(1e-6);
  device.addConstantDopingRegion({0, 30e-9}, {125e-9, 100e-9}, -5e23); // bulk
  device.addConstantDopingRegion({0, 0}, {51e-9, 30e-9}, 5e25);        // source
  device.addConstantDopingRegion({51e-9, 0}, {75e-9, 30e-9}, -5e24);   // channel
  device.addConstantDopingRegion({75e-9, 0}, {125e-9, 30e-9}, 5e25);   // drain
  device.addOhmicContact(emcBoundaryPos::YMAX, 0, {0}, {125e-9});      // substrate
  device.addOhmicContact(emcBoundaryPos::YMIN, 0, {0}, {51e-9});       // source
  device.addGateContact(emcBoundaryPos::YMIN, vg, {51e-9}, {75e-9}, 3.9, 1.2e-9, device.getMaterial().getBandGap() / 2.);
  device.addOhmicContact(emcBoundaryPos::YMIN, vd, {75e-9}, {125e-9}); // drain

  PoissonSolver solver(device, 1e-4, 1.8);
  solver.setRedBlackOrdering(redBlack != 0);
  PMScheme pmScheme;
  emcSimulationParameter<NumType, DeviceType> param;
  param.setTimes((steps - 0.5) * dt, dt, transient == 0 ? 0. : (transient - 0.5) * dt);
  param.setAdaptPotentialForWriteFunction(adaptPotential);
  param.setNamePrefix(prefix);
  param.setNrStepsBetweenShowProgress(progress);
  param.setNrStepsForFinalAvg(avg);
  if (seeded)
    param.setSeed(seed);

  auto electron = std::make_unique<electronVWD<NumType, DeviceType>>();
  SiliconModel::addXValley<NumType>(electron);
  std::vector<int> regions(device.getDopingProfile().getNrDopingRegions());
  std::iota(regions.begin(), regions.end(), 0);
  using namespace SiliconModel;
  addScattering<NumType>(electron, device, regions, ACOUSTIC | COULOMB | ZERO_ORDER, /*coulombSecond=*/true);
  if (roughness >= 0)
    electron->setSurfaceScatterMechanism(
        emcBoundaryPos::YMIN,
        std::make_unique<emcConstantSurfaceScatterMechanism<NumType, DeviceType>>(roughness, device.getMaxPos()));
  param.addParticleType(std::move(electron));

  SimulationType simulation(param, device, solver, pmScheme);
  simulation.setPoissonInterval(poissonInterval);
  const auto start = std::chrono::steady_clock::now();
  simulation.execute();
  const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
  const double n = simulation.getParticleHandler().getNrParticles(0);
  std::cout << "currents: substrate " << simulation.getAvgCurrent(0, 0) << " A, source " << simulation.getAvgCurrent(0, 1)
            << " A, gate " << simulation.getAvgCurrent(0, 2) << " A, drain " << simulation.getAvgCurrent(0, 3) << " A\n"
            << "wall time " << seconds << " s, " << steps << " steps, " << n << " particles at the end, "
            << n * steps / seconds << " particle-steps/s (Monte Carlo loop alone: " << simulation.getLoopSeconds() << " s, "
            << n * steps / simulation.getLoopSeconds() << " particle-steps/s), " << double(simulation.getTotalNrSorSweeps()) / steps
            << " SOR sweeps per step\n";
  return 0;
}
