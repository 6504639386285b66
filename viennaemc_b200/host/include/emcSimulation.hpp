// Self-consistent ensemble Monte Carlo run of a device -- the whole time-step loop is GPU resident.
// Interface mirrored: reference include/emcSimulation.hpp (template parameters and static checks
// :28-47, ctor :83-94, execute :106-127, setChannelCurrentRegion :65-70, getAvgDriftCurrent :72-74,
// setPoissonInterval :80, checkDopingProfile :231-254); same console output and result files.
//
// What differs from the reference is where the work happens.  One step of the reference
// (performEMCStep :160-193) is   Poisson -> field -> drift/scatter -> contacts -> charge assignment ->
// concentration,   each a host loop over particles or grid points with the grids passed from object to
// object.  Here these are the device kernels behind emcgpu_device_run_averaging (include/emcgpu.h):
// ensemble, potential, concentration, field and the running sums stay in GPU memory for the whole run;
// only the per-step contact counters (a few ints) come back, in chunks of steps.  The solver and the
// PM scheme objects the user passes in keep their role as plug-ins: the solver supplies accuracy and
// relaxation factor and is attached to the handler's context, the PM scheme selects the kernel variant.
#ifndef EMC_SIMULATION_HPP
#define EMC_SIMULATION_HPP

#include <algorithm>
#include <chrono>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

#include <PMSchemes/emcAbstractPMScheme.hpp>
#include <ParticleHandler/emcAbstractParticleHandler.hpp>
#include <PoissonSolver/emcAbstractSolver.hpp>
#include <emcDevice.hpp>
#include <emcGpuBinding.hpp>
#include <emcOutput.hpp>
#include <emcParticleInitialization.hpp>
#include <emcSimulationParameter.hpp>
#include <emcSimulationResults.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType, class PoissonSolver, class ParticleHandler, class PMScheme> class emcSimulation {
  static const SizeType Dim = DeviceType::Dimension;
  static_assert(std::is_same<DeviceType, emcDevice<T, Dim>>::value, "DeviceType in emcSimulation is not of the right type.");
  static_assert(std::is_base_of<emcAbstractSolver<T, DeviceType, ParticleHandler>, PoissonSolver>::value,
                "PoissonSolver in emcSimulation is not of the right type.");
  static_assert(std::is_base_of<emcAbstractPMScheme<T, DeviceType>, PMScheme>::value,
                "PMScheme in emcSimulation is not of the right type.");
  static_assert(std::is_base_of<emcAbstractParticleHandler<T, DeviceType, PMScheme, Dim>, ParticleHandler>::value,
                "ParticleHandler in emcSimulation is not of the right type.");

  emcSimulationParameter<T, DeviceType> &param;
  const DeviceType &device;
  PoissonSolver &solver;
  PMScheme pmScheme;
  ParticleHandler particleHandler;
  emcSimulationResults<T, DeviceType> results;

  bool trackChannelCurrent = false;
  T channelX0 = 0, channelX1 = 0, channelLength = 1;
  T driftCurrentSum = 0;
  SizeType driftCurrentCount = 0;
  SizeType poissonInterval = 1;
  long long totalSorSweeps = 0;
  double loopSeconds = 0; // wall time of the Monte Carlo loop alone (without equilibrium set-up and file output)

  emcgpu_ctx *ctx() { return particleHandler.gpuContext(); }
  void fetch(int grid, emcGrid<T, Dim> &into) {
    emcgpu::require(ctx(), emcgpu_device_get_grid(ctx(), grid, into.raw()), "emcgpu_device_get_grid");
  }
  // host mirrors of the device-resident grids (for the file writers)
  void fetchCurrentGrids() {
    const SizeType t = particleHandler.gpuParticleType();
    fetch(EMCGPU_GRID_POTENTIAL, results.currPot);
    fetch(EMCGPU_GRID_CONCENTRATION, results.currConc[t]);
    fetch(EMCGPU_GRID_COUNT, results.nrPart[t]);
    for (SizeType d = 0; d < Dim; d++)
      fetch(EMCGPU_GRID_EFIELD_X + static_cast<int>(d), results.eField[d]);
  }

public:
  emcSimulation() = delete;
  emcSimulation(emcSimulationParameter<T, DeviceType> &inParam, DeviceType &inDevice, PoissonSolver &inSolver,
                PMScheme inPMScheme)
      : param(inParam), device(inDevice), solver(inSolver), pmScheme(inPMScheme),
        particleHandler(device, pmScheme, param.particleTypes, param.nrCarriersPerPart, param.seedRNG),
        results(device, param) {
    param.check();
    param.print();
    checkDopingProfile();
    solver.attach(ctx()); // the solver works on the grids of the handler's context from now on
  }

  // low-noise (Ramo-Shockley) current tally over the slab x0 <= x <= x1, every non-transient step
  void setChannelCurrentRegion(T x0, T x1, T length) {
    channelX0 = x0;
    channelX1 = x1;
    channelLength = length;
    trackChannelCurrent = true;
  }
  T getAvgDriftCurrent() const { return driftCurrentCount ? driftCurrentSum / driftCurrentCount : 0; }
  // frozen-field sub-cycling: Poisson every n-th step only
  void setPoissonInterval(SizeType n) { poissonInterval = n < 1 ? 1 : n; }

  // --- additive accessors (tests, benchmarks) ---
  ParticleHandler &getParticleHandler() { return particleHandler; }
  const emcSimulationResults<T, DeviceType> &getResults() const { return results; }
  T getAvgCurrent(SizeType idxType, SizeType idxContact) const { return results.getAvgCurrent(idxType, idxContact); }
  long long getTotalNrSorSweeps() const { return totalSorSweeps; }
  double getLoopSeconds() const { return loopSeconds; }

  void execute() {
    const SizeType totalSteps = param.getNrSteps();
    const SizeType nrTransient = param.getNrTransientSteps();
    const SizeType firstAvgStep = totalSteps - param.nrStepsForFinalAvg;
    const SizeType gpuType = particleHandler.gpuParticleType();
    const SizeType nC = device.getSurface().getNrContacts();

    std::cout << "Equilibrium Characteristics ..." << std::endl;
    calcEquilibriumCharacteristics();
    writeCurrentResultsToFiles("Eq");

    std::cout << "Monte Carlo Procedure ..." << std::endl;
    emcgpu::require(ctx(), emcgpu_set_option(ctx(), "poisson_interval", static_cast<int64_t>(poissonInterval)),
                    "emcgpu_set_option");
    emcgpu::require(ctx(), emcgpu_set_option(ctx(), "sor_order", solver.getRedBlackOrdering() ? 1 : 0), "emcgpu_set_option");
    const SizeType progress = std::max<SizeType>(1, param.nrStepsBetweenShowProgress);
    std::vector<int32_t> counters, sweeps;
    auto rem = zeroCounter(), inj = zeroCounter();
    SizeType step = 0;
    const auto loopStart = std::chrono::steady_clock::now();
    while (step < totalSteps) {
      // a chunk ends after a step whose index is a multiple of the progress interval (where the reference
      // reports progress, :118-121); with the channel-current tally on, every non-transient step is its own chunk
      SizeType end = std::min(totalSteps, (step / progress) * progress + (step % progress == 0 ? 1 : progress + 1));
      if (trackChannelCurrent && end > nrTransient)
        end = step < nrTransient ? nrTransient : step + 1;
      const SizeType n = end - step;
      const SizeType nAverage = end > firstAvgStep ? end - std::max(step, firstAvgStep) : 0;
      counters.assign(n * 2 * std::max<SizeType>(1, nC), 0);
      sweeps.assign(n, 0);
      // Dirichlet values are re-imposed in the very first step only (resetBC is dropped after step 0, :119)
      emcgpu::require(ctx(),
                      emcgpu_device_run_averaging(ctx(), param.stepTime, static_cast<int>(n), static_cast<int>(nAverage),
                                                  solver.getAccuracy(), solver.getOmega(), step == 0 ? 1 : 0, counters.data(),
                                                  sweeps.data()),
                      "emcgpu_device_run_averaging");
      particleHandler.sumOverRanks(counters); // sharded run: the terminal currents count the particles of all ranks
      for (SizeType s = 0; s < n; s++) {
        totalSorSweeps += sweeps[s];
        if (step + s < nrTransient)
          continue;
        for (SizeType c = 0; c < nC; c++) {
          rem[gpuType][c] = counters[(s * 2 + 0) * nC + c];
          inj[gpuType][c] = counters[(s * 2 + 1) * nC + c];
        }
        results.updateCurrent(rem, inj);
      }
      results.nrAvgSteps += nAverage;
      if (trackChannelCurrent && end > nrTransient) {
        for (SizeType idxType = 0; idxType < param.particleTypes.size(); idxType++)
          driftCurrentSum += particleHandler.getChannelDriftCurrent(idxType, channelX0, channelX1, channelLength);
        driftCurrentCount++;
      }
      step = end;
      if ((step - 1) % progress == 0)
        showProgress(step - 1);
    }
    loopSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - loopStart).count();
    showProgress(totalSteps);
    fetchCurrentGrids();
    fetch(EMCGPU_GRID_SUM_POTENTIAL, results.avgPot);
    fetch(EMCGPU_GRID_SUM_CONCENTRATION, results.avgConc[gpuType]);
    if (particleHandler.isSharded()) {
      // every rank solved the same Poisson problem on the same (all-reduced) charge: the grids must agree bit for bit
      const bool samePot = particleHandler.identicalOnAllRanks(results.currPot.raw(), results.currPot.getSize());
      const bool sameConc = particleHandler.identicalOnAllRanks(results.avgConc[gpuType].raw(), results.avgConc[gpuType].getSize());
      const bool sameCount = particleHandler.identicalOnAllRanks(results.nrPart[gpuType].raw(), results.nrPart[gpuType].getSize());
      const bool same = samePot && sameConc && sameCount;
      if (!same)
        emcMessage::getInstance()
            .addError(std::string("sharded run: replicated grids differ between the ranks (potential ") + (samePot ? "same" : "DIFFERS") +
                      ", averaged concentration " + (sameConc ? "same" : "DIFFERS") + ", carriers per grid point " +
                      (sameCount ? "same" : "DIFFERS") + ").")
            .print();
      std::cout << "Sharded run: " << particleHandler.shardWorldSize() << " ranks, potential and averaged concentration "
                << "identical on all ranks, " << particleHandler.allReduceCalls() << " all-reduces ("
                << particleHandler.allReduceBytes() << " bytes) on this rank" << std::endl;
    }
    if (particleHandler.isShardRoot())
      results.writeFinalResults(device);
    particleHandler.print(param.namePrefix, "Final");
  }

private:
  std::vector<std::vector<int>> zeroCounter() const {
    return std::vector<std::vector<int>>(param.particleTypes.size(),
                                         std::vector<int>(device.getSurface().getNrContacts(), 0));
  }

  // equilibrium potential and field, initial ensemble drawn from them, its charge on the grid (:139-146)
  void calcEquilibriumCharacteristics() {
    solver.calcEquilibriumPotential(results.currPot, device);
    emcgpu::require(ctx(), emcgpu_device_efield(ctx()), "emcgpu_device_efield");
    particleHandler.generateInitialParticles(results.currPot);
    emcgpu::require(ctx(), emcgpu_device_assign(ctx()), "emcgpu_device_assign");
    emcgpu::require(ctx(), emcgpu_device_concentration(ctx()), "emcgpu_device_concentration");
    fetchCurrentGrids();
    particleHandler.printNrParticles();
  }

  void showProgress(SizeType nrStep) {
    std::cout << "\tNr. Iteration: \t\t" << nrStep << " / " << param.getNrSteps() << std::endl;
  }
  void writeCurrentResultsToFiles(std::string nameSuffix) {
    if (particleHandler.isShardRoot()) // the grids are replicated: rank 0 writes them
      results.writeCurrentResults(nameSuffix, device);
    particleHandler.print(param.namePrefix, nameSuffix);
  }

  // every grid point needs a doping region
  void checkDopingProfile() const {
    typename DeviceType::SizeVec coord;
    for (coord.fill(0); !device.isEndCoord(coord); device.advanceCoord(coord)) {
      if (device.getDopingProfile().getDopingRegionIdx(coord) != -1)
        continue;
      std::string where;
      for (auto c : coord)
        where += std::to_string(c) + " ";
      emcMessage::getInstance()
          .addError("Doping Regions must be added to every discrete grid point in the simulated device. Found missing "
                    "doping region at coordinate (" + where + ").")
          .print();
    }
  }
};

#endif
