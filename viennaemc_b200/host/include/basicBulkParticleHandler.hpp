// Bulk (periodic box, uniform field) particle handler -- GPU resident.
//
// Drop-in for the handler of the reference's bulk example,
// examples/bulkSimulation/basicBulkParticleHandler.hpp: same template signature,
// same public interface (ctors :91-120, setSeed :125-131, resetAppliedFieldStrength
// :134-137, generateInitialParticles :143-159, getNrParticles, getAppliedField,
// printNrParticles, moveParticles :181-225, print :227-248, printDriftVelocities /
// printVelocities :251-285, getValleyOccupationProbability :289-300, getAvgEnergy
// :304-322, getAvgDriftVelocity :326-347, deleteParticles), so the reference's
// bulkSimulation.cpp main() compiles against it unchanged.  (The include guard
// below is deliberately the reference's: pre-including this header makes the
// example's `#include "basicBulkParticleHandler.hpp"` a no-op.)
//
// What differs behind the interface:
//   * the ensemble lives in GPU memory as SoA streams (one emcgpu context per moved
//     particle type); generateInitialParticles() creates it on the host with the
//     reference's draw sequence (same seed -> same initial ensemble) and uploads it;
//   * moveParticles(dt) is a launch of the bulk step kernels, which also reduce the
//     per-valley observables of every step; getAvgEnergy / getAvgDriftVelocity /
//     getValleyOccupationProbability return those without another pass;
//   * LOOK-AHEAD: the reference's driver loop calls moveParticles(dt) and the three getAvg* once per time step
//     (bulkSimulation.cpp:150-157).  One step per call would pin the GPU to one launch and one host round trip per
//     step, so a call runs `lookahead` steps at once (emcgpu_bulk_step_ahead, default 16, setLookahead()) and the next
//     lookahead-1 calls -- and their getAvg* -- are served from the series that launch delivered.  Whatever needs the
//     ensemble of the driver's current step (print, printVelocities, a new field, seed, table rebuild, ...) first
//     rewinds to the state before the launch and repeats exactly the steps served so far: the random numbers are keyed
//     by particle id and step, so this reproduces them bit for bit.  Types with phonon baths (per-step host feedback)
//     keep one step per call; the grain clocks of a grain mechanism ride along (the look-ahead copy has its own);
//   * SEVERAL GPUs: started once per GPU with EMCGPU_SHARD=1 and the launcher variables RANK, WORLD_SIZE,
//     LOCAL_RANK plus EMCNCCL_ID_FILE (like the device-run handler, ParticleHandler/emcBasicParticleHandler.hpp), every
//     process creates the same initial ensemble (the seed of rank 0) and keeps a contiguous block of it on its GPU.  The
//     Philox streams are keyed by the particle's position in the WHOLE ensemble, so every particle moves exactly as in a
//     one-GPU run of the same seed.  What the driver reads -- the per-valley sums behind getAvg* / occupation, the event
//     counters of the phonon baths -- is summed over the ranks (ncclAllReduce through libemcnccl: once per look-ahead
//     window, or once per step with phonon baths), so baths, screening and rebuilt rate tables are identical on all ranks.
//     getNrParticles() is the size of the whole ensemble; print() / printVelocities() write the block of this rank
//     ("<...>.rank<r>.txt"); start every rank in a directory of its own if the driver writes files;
//   * random numbers in the step are counter-based Philox streams keyed by particle
//     id and step (not one mt19937_64 per OpenMP thread), seeded from the handler seed;
//   * nothing is moved on the CPU: without a CUDA device, or with a mechanism /
//     valley that has no device implementation, construction fails with an error.
#ifndef BASIC_BULK_PARTICLE_HANDLER_HPP
#define BASIC_BULK_PARTICLE_HANDLER_HPP

#include <array>
#include <chrono>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <ParticleType/emcParticleType.hpp>
#include <detail/emcBulkEnsembleBuilder.hpp>
#include <emcGpuBinding.hpp>
#include <emcPhononBath.hpp>
#include <emcnccl.h>
#include <emcGrid.hpp>
#include <emcParticleInitialization.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType, SizeType Dim = DeviceType::Dimension> struct basicBulkParticleHandler {
  static_assert(Dim == 3, "the bulk handler simulates a 3-D periodic box");
  typedef emcParticleType<T, DeviceType> ParticleType;
  typedef typename DeviceType::SizeVec SizeVec;
  typedef typename DeviceType::ValueVec ValueVec;
  typedef std::map<SizeType, std::unique_ptr<ParticleType>> MapIdxToParticleTypes;

private:
  typedef emcdetail::HostEnsemble HostEnsemble;
  struct TypeState {
    emcgpu_ctx *ctx = nullptr;
    HostEnsemble staging;
    SizeType nrParticles = 0;
    bool uploaded = false;
    std::vector<double> lastObs; // [valley][3] sums of the last step (or of the resting ensemble)
    bool obsValid = false;
    bool grain = false;                           // the type carries a grain mechanism: clocks live on the device
    SizeType tableVersion = 0;                    // version of the scatter tables on the device
    std::vector<emcPhononBath<T> *> phononBaths;  // baths fed by the polar-optical mechanisms of the type
    // look-ahead: per-step sums of the steps the device has already done, and how many of them the driver has asked for
    std::vector<double> ahead;
    SizeType aheadN = 0, aheadServed = 0;
    T aheadDt = 0;
    // per-particle velocities of the steps of the current window (printDriftVelocities / printVelocities): recorded by
    // the step kernels once a driver has asked for them (components: 0 = not yet, 1 = v.E_dir, 3 = v), [step][particle][c]
    int recordVel = 0;
    std::vector<double> velAhead;
    SizeType velSteps = 0; // steps of the current window velAhead holds
    // several GPUs: this process holds the particles [shardFirst, shardFirst + nrLocal) of the ensemble (one GPU: all)
    SizeType shardFirst = 0, nrLocal = 0;
  };

  DeviceType &device;
  MapIdxToParticleTypes &idxTypeToPartType;
  ValueVec appliedFieldDir;
  ValueVec appliedField;
  T fieldStrength = 0;
  mutable std::map<SizeType, TypeState> state;
  SizeType lookahead = defaultLookahead();
  emcRNG hostRng; // particle creation only (the reference's rngs[0])
  unsigned long stepSeed = 0;
  int mathMode = EMCGPU_MATH_FAST;
  int shardRank = 0, shardWorld = 1;
  emcnccl_comm *comm = nullptr;

  static int envInt(const char *name, int fallback) {
    const char *e = std::getenv(name);
    return (e && *e) ? std::atoi(e) : fallback;
  }
  int cudaDeviceOrdinal() const { return envInt("EMCGPU_DEVICE", shardWorld > 1 ? envInt("LOCAL_RANK", shardRank) : 0); }
  // in-place sum over the ranks of a small host array (no-op on one GPU)
  void sumOverRanks(std::vector<double> &v) const {
    if (comm && !v.empty() && emcnccl_allreduce_sum_host_f64(comm, v.data(), static_cast<int64_t>(v.size())) != 0)
      emcMessage::getInstance().addError(std::string("sharded run: all-reduce failed: ") + emcnccl_last_error()).print();
  }
  static SizeType defaultLookahead() {
    const char *e = std::getenv("EMCGPU_LOOKAHEAD");
    const long v = e ? std::atol(e) : 16;
    return v < 1 ? 1 : static_cast<SizeType>(v);
  }

  // The device is ahead of the driver (steps done but not yet asked for): back to the ensemble of the driver's current
  // step.  Rewind to the state before the look-ahead launch and repeat the steps served so far -- same particle ids, same
  // step indices, same tables (a rebuilt table set has not been uploaded yet at this point), hence the same trajectories.
  void materialize(SizeType idxType) const {
    auto &st = state.at(idxType);
    if (st.ctx && st.aheadServed < st.aheadN) {
      if (st.nrLocal > 0)
        emcgpu::require(st.ctx, emcgpu_bulk_rewind(st.ctx), "emcgpu_bulk_rewind");
      if (st.aheadServed > 0 && st.nrLocal > 0) {
        emcgpu::require(st.ctx, emcgpu_bulk_record_velocities(st.ctx, 0, nullptr, 0), "emcgpu_bulk_record_velocities");
        std::vector<double> again(st.aheadServed * idxTypeToPartType.at(idxType)->getNrValleys() * 3);
        emcgpu::require(st.ctx,
                        emcgpu_bulk_step(st.ctx, st.aheadDt, static_cast<int>(st.aheadServed), static_cast<int>(st.aheadServed),
                                         again.data()),
                        "emcgpu_bulk_step");
      }
    }
    st.aheadN = st.aheadServed = 0;
    st.velSteps = 0;
  }
  void materializeAll() const {
    for (auto &[idxType, st] : state) {
      (void)st;
      materialize(idxType);
    }
  }

  void configure(SizeType idxType) {
    auto &st = state[idxType];
    materialize(idxType);
    const auto box = device.getMaxPos();
    const double b[3] = {box[0], box[1], box[2]}, d[3] = {appliedFieldDir[0], appliedFieldDir[1], appliedFieldDir[2]};
    emcgpu::require(st.ctx,
                    emcgpu_bulk_configure(st.ctx, b, d, fieldStrength, idxTypeToPartType[idxType]->getCharge(), mathMode),
                    "emcgpu_bulk_configure");
  }

  void upload(SizeType idxType) {
    auto &st = state[idxType];
    if (st.uploaded)
      return;
    // contiguous blocks whose sizes differ by at most one (viennaemc_b200/sharding.py: shard_range); the particle id that
    // keys the Philox streams is the position in the whole ensemble
    const SizeType base = st.staging.size() / shardWorld, rem = st.staging.size() % shardWorld;
    st.shardFirst = shardRank * base + std::min<SizeType>(shardRank, rem);
    st.nrLocal = base + (static_cast<SizeType>(shardRank) < rem ? 1 : 0);
    const double *ptrs[EMCGPU_N_STREAMS];
    for (int s = 0; s < EMCGPU_N_STREAMS; s++)
      ptrs[s] = st.staging.stream[s].data() + st.shardFirst;
    emcgpu::require(st.ctx,
                    emcgpu_set_ensemble(st.ctx, static_cast<int64_t>(st.nrLocal), st.nrLocal ? ptrs : nullptr,
                                        st.nrLocal ? st.staging.packed.data() + st.shardFirst : nullptr,
                                        static_cast<int64_t>(st.shardFirst)),
                    "emcgpu_set_ensemble");
    if (st.grain && st.nrLocal)
      emcgpu::require(st.ctx, emcgpu_set_grain_clock(st.ctx, st.staging.grainTau.data() + st.shardFirst), "emcgpu_set_grain_clock");
    emcgpu::require(st.ctx, emcgpu_rng_philox(st.ctx, stepSeed), "emcgpu_rng_philox");
    emcgpu::require(st.ctx, emcgpu_set_step_index(st.ctx, 1), "emcgpu_set_step_index");
    st.uploaded = true;
    st.obsValid = false;
  }

  void download(SizeType idxType, HostEnsemble &out) const {
    materialize(idxType);
    const auto &st = state.at(idxType);
    const SizeType n = st.nrLocal; // the block of this rank
    double *ptrs[EMCGPU_N_STREAMS];
    for (int s = 0; s < EMCGPU_N_STREAMS; s++) {
      out.stream[s].resize(n);
      ptrs[s] = out.stream[s].data();
    }
    out.packed.resize(n);
    if (n)
      emcgpu::require(st.ctx, emcgpu_get_ensemble(st.ctx, ptrs, out.packed.data()), "emcgpu_get_ensemble");
  }

  // tables rebuilt on the host since the last upload (reinitScatterTables, e.g. after a phonon-bath update): upload
  // them again; the prefix sums of the baths follow the baths in any case (the q-resolved angle reads them)
  void refreshModel(SizeType idxType) {
    auto &st = state[idxType];
    auto &type = *idxTypeToPartType[idxType];
    if (type.scatterHandler.getTableVersion() != st.tableVersion) {
      emcgpu::uploadParticleType(st.ctx, type);
      st.tableVersion = type.scatterHandler.getTableVersion();
    } else {
      emcgpu::uploadPhononBaths(st.ctx, st.phononBaths);
    }
  }

  const std::vector<double> &observables(SizeType idxType) {
    auto &st = state.at(idxType);
    if (!st.obsValid) {
      materialize(idxType);
      upload(idxType);
      st.lastObs.assign(idxTypeToPartType[idxType]->getNrValleys() * 3, 0.);
      if (st.nrLocal)
        emcgpu::require(st.ctx, emcgpu_bulk_observables(st.ctx, st.lastObs.data()), "emcgpu_bulk_observables");
      sumOverRanks(st.lastObs);
      st.obsValid = true;
    }
    return st.lastObs;
  }

public:
  basicBulkParticleHandler() = delete;
  basicBulkParticleHandler(const basicBulkParticleHandler &) = delete;

  basicBulkParticleHandler(DeviceType &inDevice, MapIdxToParticleTypes &inTypes, const ValueVec &inFieldDirection)
      : basicBulkParticleHandler(inDevice, inTypes, inFieldDirection, 0) {}

  // inSeed == 0: seed from the clock (as the reference does)
  basicBulkParticleHandler(DeviceType &inDevice, MapIdxToParticleTypes &inTypes, const ValueVec &inFieldDirection,
                           T inFieldStrength, long unsigned int inSeed = 0)
      : device(inDevice), idxTypeToPartType(inTypes), appliedFieldDir(inFieldDirection), fieldStrength(inFieldStrength) {
    normalize(appliedFieldDir);
    appliedField = scale(appliedFieldDir, inFieldStrength);
    // (EMCGPU_SEED replaces the clock of an unseeded handler: reproducible runs of an unmodified main())
    const char *envSeed = std::getenv("EMCGPU_SEED");
    const unsigned long seed =
        inSeed != 0 ? inSeed
        : (envSeed && *envSeed)
            ? std::strtoul(envSeed, nullptr, 10)
            : static_cast<unsigned long>(std::chrono::high_resolution_clock::now().time_since_epoch().count());
    unsigned long commonSeed = seed;
    if (envInt("EMCGPU_SHARD", 0)) {
      shardWorld = std::max(1, envInt("WORLD_SIZE", 1));
      shardRank = envInt("RANK", 0);
    }
    if (shardWorld > 1) {
      const char *idFile = std::getenv("EMCNCCL_ID_FILE");
      if (!idFile || emcnccl_init_from_file(idFile, shardRank, shardWorld, cudaDeviceOrdinal(), 120., &comm) != 0)
        emcMessage::getInstance()
            .addError(std::string("sharded run: cannot join the NCCL communicator (EMCNCCL_ID_FILE must name a file all "
                                  "ranks can reach): ") + emcnccl_last_error())
            .print();
      // every rank creates the same ensemble and draws the same step streams: the seed of rank 0, in 16-bit pieces (exact)
      std::vector<double> pieces(4, 0.);
      if (shardRank == 0)
        for (int i = 0; i < 4; i++)
          pieces[i] = static_cast<double>((static_cast<unsigned long long>(seed) >> (16 * i)) & 0xffffull);
      sumOverRanks(pieces);
      unsigned long long joined = 0;
      for (int i = 0; i < 4; i++)
        joined |= static_cast<unsigned long long>(pieces[i]) << (16 * i);
      commonSeed = static_cast<unsigned long>(joined);
    }
    hostRng.seed(commonSeed);
    stepSeed = commonSeed;
    for (const auto &[idxType, type] : idxTypeToPartType) {
      auto &st = state[idxType];
      if (!type->isMoved())
        continue;
      type->initScatterTables(); // host, exactly as the reference
      int rc = emcgpu_create(cudaDeviceOrdinal(), &st.ctx);
      if (rc != EMCGPU_OK)
        emcMessage::getInstance()
            .addError(std::string("cannot create the GPU context for ") + type->getName() + ": " +
                      emcgpu_last_error(nullptr))
            .print();
      emcgpu::uploadParticleType(st.ctx, *type);
      st.grain = emcgpu::uploadGrainMechanism(st.ctx, *type);
      st.tableVersion = type->scatterHandler.getTableVersion();
      st.phononBaths = emcgpu::collectPhononBaths(*type);
      configure(idxType);
    }
  }

  ~basicBulkParticleHandler() {
    for (auto &[idxType, st] : state) {
      (void)idxType;
      if (st.ctx)
        emcgpu_destroy(st.ctx);
    }
    if (comm)
      emcnccl_destroy(comm);
  }

  // EMCGPU_MATH_EXACT reproduces the reference's rounding operation by operation (replay parity);
  // the default EMCGPU_MATH_FAST agrees with it to ~1e-15 per step
  void setMathMode(int mode) {
    mathMode = mode;
    for (const auto &[idxType, type] : idxTypeToPartType)
      if (type->isMoved())
        configure(idxType);
  }
  bool isSharded() const { return shardWorld > 1; }
  int shardRankOf() const { return shardRank; }
  int shardWorldSize() const { return shardWorld; }
  // the particles of a type this rank holds (getNrParticles(): the whole ensemble)
  SizeType getNrParticlesOfThisRank(SizeType idxType) const {
    const auto &st = state.at(idxType);
    return st.uploaded ? st.nrLocal : st.nrParticles;
  }
  // the C-ABI context of a particle type (multi-GPU drivers, tests)
  emcgpu_ctx *getGpuContext(SizeType idxType) {
    materialize(idxType);
    return state.at(idxType).ctx;
  }
  // time steps a moveParticles(dt) call runs at once on the device (1 = one launch per call); see the header comment
  void setLookahead(SizeType nSteps) {
    materializeAll();
    lookahead = nSteps < 1 ? 1 : nSteps;
  }
  SizeType getLookahead() const { return lookahead; }

  void setSeed(SizeType inSeed) {
    materializeAll();
    hostRng.seed(inSeed);
    stepSeed = inSeed;
    for (auto &[idxType, st] : state)
      if (st.ctx && st.uploaded)
        emcgpu::require(st.ctx, emcgpu_rng_philox(st.ctx, stepSeed), "emcgpu_rng_philox");
  }

  void resetAppliedFieldStrength(T inAppliedFieldStrength) {
    fieldStrength = inAppliedFieldStrength;
    appliedField = scale(appliedFieldDir, inAppliedFieldStrength);
    for (const auto &[idxType, type] : idxTypeToPartType)
      if (type->isMoved())
        configure(idxType);
  }

  // cell by cell in storage order: floor(n) particles plus one more with probability frac(n)
  void generateInitialParticles() {
    for (const auto &[idxType, type] : idxTypeToPartType) {
      auto &st = state[idxType];
      st.aheadN = st.aheadServed = 0;
      emcdetail::generateBulkEnsemble(st.staging, *type, device, hostRng);
      st.nrParticles = st.staging.size();
      st.uploaded = false;
      st.obsValid = false;
      if (type->isMoved())
        upload(idxType);
    }
  }

  SizeType getNrParticles(SizeType idxType) const { return state.at(idxType).nrParticles; }
  ValueVec getAppliedField() const { return appliedField; }

  void printNrParticles() const {
    for (const auto &[idxType, type] : idxTypeToPartType)
      std::cout << "\t" << state.at(idxType).nrParticles << " " << type->getName() << "\n";
  }

  // one time step of every moved particle type: free flights, scattering, periodic wrap
  void moveParticles(T tStep) {
    for (const auto &[idxType, type] : idxTypeToPartType) {
      if (!type->isMoved())
        continue;
      auto &st = state[idxType];
      if (st.nrParticles == 0)
        continue;
      const SizeType nObs = type->getNrValleys() * 3;
      // this step is already done on the device: hand out its sums
      if (st.aheadServed < st.aheadN && tStep == st.aheadDt && type->scatterHandler.getTableVersion() == st.tableVersion) {
        st.lastObs.assign(st.ahead.begin() + st.aheadServed * nObs, st.ahead.begin() + (st.aheadServed + 1) * nObs);
        st.aheadServed++;
        st.obsValid = true;
        continue;
      }
      materialize(idxType);
      upload(idxType);
      refreshModel(idxType);
      SizeType nAhead = st.phononBaths.empty() ? lookahead : 1;
      st.velSteps = 0;
      const bool here = st.nrLocal > 0; // (a rank of a sharded run may hold none of a tiny ensemble: it only joins the sums)
      if (st.recordVel) { // the window's velocities come back with it (at most 1 GB of them per window)
        const SizeType perStep = st.nrLocal * st.recordVel;
        nAhead = std::max<SizeType>(1, std::min<SizeType>(nAhead, (SizeType(1) << 27) / std::max<SizeType>(1, perStep)));
        st.velAhead.resize(std::max<SizeType>(1, nAhead * perStep));
        emcgpu::require(st.ctx, emcgpu_bulk_record_velocities(st.ctx, st.recordVel, st.velAhead.data(), static_cast<int64_t>(nAhead)),
                        "emcgpu_bulk_record_velocities");
        st.velSteps = nAhead;
      } else {
        emcgpu::require(st.ctx, emcgpu_bulk_record_velocities(st.ctx, 0, nullptr, 0), "emcgpu_bulk_record_velocities");
      }
      if (nAhead > 1) {
        st.ahead.assign(nAhead * nObs, 0.);
        if (here)
          emcgpu::require(st.ctx,
                          emcgpu_bulk_step_ahead(st.ctx, tStep, static_cast<int>(nAhead), static_cast<int>(nAhead), st.ahead.data()),
                          "emcgpu_bulk_step_ahead");
        sumOverRanks(st.ahead); // one all-reduce per look-ahead window
        st.aheadN = nAhead;
        st.aheadServed = 1;
        st.aheadDt = tStep;
        st.lastObs.assign(st.ahead.begin(), st.ahead.begin() + nObs);
      } else {
        st.lastObs.assign(nObs, 0.);
        if (here)
          emcgpu::require(st.ctx, emcgpu_bulk_step(st.ctx, tStep, 1, 1, st.lastObs.data()), "emcgpu_bulk_step");
        sumOverRanks(st.lastObs);
        // recordEmission / recordAbsorption of this step, of all ranks
        emcgpu::collectPhononCounts(st.ctx, st.phononBaths, [this](std::vector<double> &v) { sumOverRanks(v); }, here);
      }
      st.obsValid = true;
    }
  }

  // nSteps time steps in one call; series receives per step and valley {sum E, sum v.E_dir, count}
  // (additive to the reference interface: lets a driver keep the observables of a whole run on the
  // device side and fuse several steps per kernel launch)
  void moveParticles(T tStep, SizeType nSteps, SizeType stepsPerLaunch, SizeType idxType, std::vector<double> &series) {
    auto &st = state.at(idxType);
    materialize(idxType);
    upload(idxType);
    refreshModel(idxType);
    const SizeType nV = idxTypeToPartType[idxType]->getNrValleys();
    series.assign(nSteps * nV * 3, 0.);
    if (st.nrLocal > 0)
      emcgpu::require(st.ctx,
                      emcgpu_bulk_step(st.ctx, tStep, static_cast<int>(nSteps), static_cast<int>(stepsPerLaunch), series.data()),
                      "emcgpu_bulk_step");
    sumOverRanks(series);
    emcgpu::collectPhononCounts(st.ctx, st.phononBaths, [this](std::vector<double> &v) { sumOverRanks(v); }, st.nrLocal > 0);
    st.lastObs.assign(series.end() - nV * 3, series.end());
    st.obsValid = true;
  }

  // band filling (reference :380-470) serialises the particle loop: not on the GPU path, and never run on the CPU here
  template <class PauliExclusion> void moveParticleTypeWithBandFilling(T, SizeType, PauliExclusion &) {
    emcMessage::getInstance()
        .addError("moveParticleTypeWithBandFilling: Pauli exclusion makes the particle loop sequential; it has no GPU "
                  "implementation and there is no CPU fallback.")
        .print();
  }

  // Pairwise / host-side ensemble edits of the hot-carrier example (reference :472-640: carrierCarrierScatter,
  // interCarrierScatter, recombine, extractCarriers).  They walk the host ensemble particle by particle, in pairs, in
  // sequence -- not part of the data-parallel particle loop and never run on the CPU here: rejected with their names.
  template <class Scatter> void carrierCarrierScatter(Scatter &, SizeType, T) { rejectHostStep(Scatter::name(), "carrierCarrierScatter"); }
  template <class Scatter> void interCarrierScatter(Scatter &, SizeType, SizeType, T) { rejectHostStep(Scatter::name(), "interCarrierScatter"); }
  template <class Recombination> void recombine(Recombination &, SizeType, SizeType, T) { rejectHostStep(Recombination::name(), "recombine"); }
  template <class Contact> void extractCarriers(Contact &, SizeType, T) { rejectHostStep(Contact::name(), "extractCarriers"); }

  // "<prefix><TypeName><suffix>.txt": box extent, then per particle: index, position[, k, energy, sub-valley, valley]
  void print(std::string namePrefix, std::string nameSuffix) const {
    for (const auto &[idxType, type] : idxTypeToPartType) {
      HostEnsemble h;
      const auto &st = state.at(idxType);
      const HostEnsemble *src = &st.staging;
      SizeType n = st.nrParticles, first = 0;
      if (st.uploaded) { // the block of this rank (one GPU: everything), numbered as in the whole ensemble
        download(idxType, h);
        src = &h;
        n = st.nrLocal;
        first = st.shardFirst;
      }
      std::ofstream os(namePrefix + type->getName() + nameSuffix + (comm ? ".rank" + std::to_string(shardRank) : std::string()) +
                       ".txt");
      os << device.getMaxPos() << "\n";
      for (SizeType i = 0; i < n; i++) {
        os << first + i << " " << src->stream[EMCGPU_X][i] << " " << src->stream[EMCGPU_Y][i] << " " << src->stream[EMCGPU_Z][i];
        if (type->isMoved())
          os << " " << src->stream[EMCGPU_KX][i] << " " << src->stream[EMCGPU_KY][i] << " " << src->stream[EMCGPU_KZ][i]
             << " " << src->stream[EMCGPU_ENERGY][i] << " " << ((src->packed[i] >> 8) & 0xffu) << " "
             << (src->packed[i] & 0xffu);
        if (i + 1 < n)
          os << "\n";
      }
    }
  }

  // one line per moved type: v.E_dir of every particle (for velocity autocorrelation post-processing)
  void printDriftVelocities(std::ofstream &os) const { printVelocityLines(os, true); }
  // one line per moved type: the three velocity components of every particle
  void printVelocities(std::ofstream &os) const { printVelocityLines(os, false); }

  std::vector<T> getValleyOccupationProbability(SizeType idxType) {
    const auto &obs = observables(idxType);
    const SizeType nV = idxTypeToPartType[idxType]->getNrValleys();
    const T total = static_cast<T>(state.at(idxType).nrParticles);
    std::vector<T> occ(nV, 0.);
    for (SizeType v = 0; v < nV; v++)
      occ[v] = obs[3 * v + 2] / total;
    return occ;
  }
  std::vector<T> getAvgEnergy(SizeType idxType) { return perValleyMean(idxType, 0); }
  std::vector<T> getAvgDriftVelocity(SizeType idxType) { return perValleyMean(idxType, 1); }

  void deleteParticles() {
    for (auto &[idxType, st] : state) {
      (void)idxType;
      st.aheadN = st.aheadServed = 0;
      st.staging.clear();
      st.nrParticles = 0;
      st.uploaded = false;
      st.obsValid = false;
      if (st.ctx)
        emcgpu::require(st.ctx, emcgpu_set_ensemble(st.ctx, 0, nullptr, nullptr, 0), "emcgpu_set_ensemble");
    }
  }

private:
  static void rejectHostStep(const char *cls, const char *method) {
    emcMessage::getInstance()
        .addError(std::string("basicBulkParticleHandler::") + method + ": " + cls +
                  " edits the host ensemble pairwise / sequentially; it has no GPU implementation and there is no CPU "
                  "fallback (run the example with this mechanism switched off).")
        .print();
  }
  std::vector<T> perValleyMean(SizeType idxType, int which) {
    const auto &obs = observables(idxType);
    const SizeType nV = idxTypeToPartType[idxType]->getNrValleys();
    std::vector<T> mean(nV, 0.);
    for (SizeType v = 0; v < nV; v++)
      if (obs[3 * v + 2] != 0)
        mean[v] = obs[3 * v + which] / obs[3 * v + 2];
    return mean;
  }

  void printVelocityLines(std::ofstream &os, bool projected) const {
    for (const auto &[idxType, type] : idxTypeToPartType) {
      if (!type->isMoved())
        continue;
      HostEnsemble h;
      auto &st = state.at(idxType);
      const int comps = projected ? 1 : 3;
      // the step kernels recorded the velocities of the driver's current step: one line from the record
      const SizeType cur = st.aheadN ? st.aheadServed : (st.velSteps ? 1 : 0); // 1-based step of the window
      if (st.recordVel == comps && cur >= 1 && cur <= st.velSteps && st.obsValid) {
        const double *v = st.velAhead.data() + (cur - 1) * st.nrLocal * comps;
        for (SizeType i = 0; i < st.nrLocal; i++) {
          if (projected)
            os << v[i];
          else
            os << std::array<T, 3>{v[3 * i], v[3 * i + 1], v[3 * i + 2]};
          if (i + 1 < st.nrLocal)
            os << " ";
        }
        os << std::endl;
        continue;
      }
      st.recordVel = comps; // from the next launch on the kernels record them
      const HostEnsemble *src = &st.staging;
      SizeType n = st.nrParticles;
      if (st.uploaded) { // the block of this rank (one GPU: everything)
        download(idxType, h);
        src = &h;
        n = st.nrLocal;
      }
      for (SizeType i = 0; i < n; i++) {
        const std::array<T, 3> k = {src->stream[EMCGPU_KX][i], src->stream[EMCGPU_KY][i], src->stream[EMCGPU_KZ][i]};
        const auto *valley = type->getValley(src->packed[i] & 0xffu);
        const auto vel = valley->getVelocity(k, src->stream[EMCGPU_ENERGY][i], (src->packed[i] >> 8) & 0xffu);
        if (projected)
          os << innerProduct(vel, appliedFieldDir);
        else
          os << vel;
        if (i + 1 < n)
          os << " ";
      }
      os << std::endl;
    }
  }
};

#endif
