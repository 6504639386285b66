// Particle handler of a device run -- GPU resident.
// Interface mirrored: reference include/ParticleHandler/emcBasicParticleHandler.hpp (ctor :52-60,
// getNrParticles, printNrParticles, driftScatterParticles :76-145, assignParticlesToMesh :148-153,
// handleOhmicContacts :158-192, print :194-216, getChannelDriftCurrent :224-237) and the initial
// ensemble of emcAbstractParticleHandler::generateInitialParticles (:133-148).
//
// The ensemble of the (one) moved particle type lives in GPU memory behind a C-ABI context; every
// public member maps to one emcgpu_device_* call.  emcSimulation does not go through these per-call
// members in its step loop: it drives the whole step (Poisson, field, particles, contacts, charge
// assignment) device resident through gpuContext().  Initial particles are created on the host with
// the reference's draw order (same seed -> same ensemble); contact injection and the step itself draw
// from counter-based Philox streams.
#ifndef EMC_BASIC_PARTICLE_HANDLER_HPP
#define EMC_BASIC_PARTICLE_HANDLER_HPP

#include <cstdlib>
#include <fstream>

#include <ParticleHandler/emcAbstractParticleHandler.hpp>
#include <detail/emcBulkEnsembleBuilder.hpp>
#include <detail/emcDeviceFlatten.hpp>
#include <emcGpuBinding.hpp>
#include <emcnccl.h>

// SEVERAL GPUs: started once per GPU with EMCGPU_SHARD=1 and the usual launcher variables RANK, WORLD_SIZE,
// LOCAL_RANK (torchrun --no-python, mpirun, a shell loop) plus EMCNCCL_ID_FILE (a path all ranks can reach), every process
// creates the same initial ensemble (same seed), keeps its block of the particles and joins one NCCL communicator
// (libemcnccl).  Potential, field and concentration are replicated; per step the library all-reduces the reservoir share
// table and the carriers-per-grid-point grid over NVLink (emcgpu_device_set_sharding).  Counters and ensemble sizes the
// host sees are sums over the ranks; grid and current files are written by rank 0, particle files per rank.
template <class T, class DeviceType, class PMScheme, SizeType Dim = DeviceType::Dimension>
class emcBasicParticleHandler : public emcAbstractParticleHandler<T, DeviceType, PMScheme, Dim> {
  typedef emcAbstractParticleHandler<T, DeviceType, PMScheme, Dim> Base;

public:
  typedef typename Base::SizeVec SizeVec;
  typedef typename Base::ValueVec ValueVec;
  typedef typename Base::MapIdxToParticleTypes MapIdxToParticleTypes;
  typedef typename Base::NettoParticleCounter NettoParticleCounter;

private:
  emcgpu_ctx *ctx = nullptr;
  SizeType gpuType = 0; // index of the particle type that lives on the GPU
  emcRNG hostRng;       // particle creation only (the reference's rngs[0])
  SizeType seed;
  emcdetail::HostEnsemble staging;
  bool uploaded = false;
  bool grain = false; // the type carries a grain mechanism: its clocks live on the device
  // several GPUs: this process holds the particles [shardFirst, shardFirst + shardCount) of the initial ensemble
  int shardRank = 0, shardWorld = 1;
  emcnccl_comm *comm = nullptr;
  SizeType shardFirst = 0, shardCount = 0;

  static int envInt(const char *name, int def) {
    const char *e = std::getenv(name);
    return (e && *e) ? std::atoi(e) : def;
  }

  SizeType nrOnGpu() const { return uploaded ? static_cast<SizeType>(emcgpu_ensemble_size(ctx)) : staging.size(); }

  void upload() {
    if (uploaded)
      return;
    // contiguous blocks whose sizes differ by at most one (viennaemc_b200/sharding.py: shard_range)
    const SizeType base = staging.size() / shardWorld, rem = staging.size() % shardWorld;
    shardFirst = shardRank * base + std::min<SizeType>(shardRank, rem);
    shardCount = base + (static_cast<SizeType>(shardRank) < rem ? 1 : 0);
    const double *ptrs[EMCGPU_N_STREAMS];
    for (int s = 0; s < EMCGPU_N_STREAMS; s++)
      ptrs[s] = staging.stream[s].data() + shardFirst;
    // particle ids (they key the Philox streams) of different ranks never meet: rank << 40
    emcgpu::require(ctx,
                    emcgpu_set_ensemble(ctx, static_cast<int64_t>(shardCount), ptrs, staging.packed.data() + shardFirst,
                                        static_cast<int64_t>(shardRank) << 40),
                    "emcgpu_set_ensemble");
    if (grain && shardCount)
      emcgpu::require(ctx, emcgpu_set_grain_clock(ctx, staging.grainTau.data() + shardFirst), "emcgpu_set_grain_clock");
    emcgpu::require(ctx, emcgpu_rng_philox(ctx, seed), "emcgpu_rng_philox");
    emcgpu::require(ctx, emcgpu_set_step_index(ctx, 1), "emcgpu_set_step_index");
    // head room for injected particles: avoids re-allocations during the run
    emcgpu::require(ctx, emcgpu_device_reserve(ctx, static_cast<int64_t>(shardCount * 1.25) + 1024),
                    "emcgpu_device_reserve");
    uploaded = true;
  }
  void download(emcdetail::HostEnsemble &h) const {
    const SizeType n = nrOnGpu();
    double *ptrs[EMCGPU_N_STREAMS];
    for (int s = 0; s < EMCGPU_N_STREAMS; s++) {
      h.stream[s].resize(n);
      ptrs[s] = h.stream[s].data();
    }
    h.packed.resize(n);
    if (n)
      emcgpu::require(ctx, emcgpu_get_ensemble(ctx, ptrs, h.packed.data()), "emcgpu_get_ensemble");
  }

public:
  emcBasicParticleHandler() = delete;
  emcBasicParticleHandler(const emcBasicParticleHandler &) = delete;
  emcBasicParticleHandler(const DeviceType &inDevice, PMScheme &inPMScheme, MapIdxToParticleTypes &inTypes,
                          SizeType inNrCarriersPerParticle, SizeType inSeed)
      : Base(inDevice, inPMScheme, inNrCarriersPerParticle, inTypes), hostRng(inSeed), seed(inSeed) {
    if (inPMScheme.deviceSchemeId() < 1)
      emcMessage::getInstance()
          .addError("The particle-mesh scheme has no device implementation (deviceSchemeId() == 0); it cannot run on "
                    "the GPU path and there is no CPU fallback.")
          .print();
    SizeType moved = 0;
    for (const auto &[idxType, type] : this->idxTypeToPartType) {
      if (!type->isMoved())
        continue;
      gpuType = idxType;
      moved++;
      type->initScatterTables(); // host, exactly as the reference (emcAbstractParticleHandler.hpp:99-101)
    }
    if (moved != 1)
      emcMessage::getInstance()
          .addError("The GPU particle handler moves exactly one particle type (found " + std::to_string(moved) + ").")
          .print();
    if (envInt("EMCGPU_SHARD", 0)) {
      shardWorld = std::max(1, envInt("WORLD_SIZE", 1));
      shardRank = envInt("RANK", 0);
    }
    const int ordinal = envInt("EMCGPU_DEVICE", shardWorld > 1 ? envInt("LOCAL_RANK", shardRank) : 0);
    if (emcgpu_create(ordinal, &ctx) != EMCGPU_OK)
      emcMessage::getInstance()
          .addError(std::string("cannot create the GPU context: ") + emcgpu_last_error(nullptr))
          .print();
    if (shardWorld > 1) {
      const char *idFile = std::getenv("EMCNCCL_ID_FILE");
      if (!idFile || emcnccl_init_from_file(idFile, shardRank, shardWorld, ordinal, 120., &comm) != 0)
        emcMessage::getInstance()
            .addError(std::string("sharded run: cannot join the NCCL communicator (EMCNCCL_ID_FILE must name a file all "
                                  "ranks can reach): ") + emcnccl_last_error())
            .print();
    }
    auto &type = *this->idxTypeToPartType.at(gpuType);
    emcgpu::uploadParticleType(ctx, type);
    grain = emcgpu::uploadGrainMechanism(ctx, type);
    emcdetail::FlatDevice<T, Dim> flat(this->device);
    flat.desc.pmScheme = inPMScheme.deviceSchemeId() - 1;
    emcgpu::require(ctx,
                    emcgpu_device_configure(ctx, &flat.desc, type.getCharge(), static_cast<double>(this->nrCarriersPerPart),
                                            this->expNrPart[gpuType].raw(), EMCGPU_MATH_FAST),
                    "emcgpu_device_configure");
    if (comm)
      emcgpu::require(ctx, emcgpu_device_set_sharding(ctx, shardRank, shardWorld, emcnccl_allreduce_sum_f64, comm),
                      "emcgpu_device_set_sharding");
    // plug-ins of the particle type that act inside the device kernels: wall mechanisms, creation rules at contacts
    for (SizeType face = 0; face < 2 * Dim; face++) {
      const auto *wall = type.scatterHandler.getSurfaceScatterMechanism(face);
      if (!wall)
        continue;
      if (wall->deviceSurfaceKind() < 0)
        emcMessage::getInstance()
            .addError("The surface scatter mechanism of face " + std::to_string(face) + " of " + type.getName() +
                      " has no device implementation (deviceSurfaceKind() < 0); it cannot run on the GPU path and there "
                      "is no CPU fallback.")
            .print();
      emcgpu::require(ctx, emcgpu_device_set_surface(ctx, static_cast<int>(face), wall->deviceSurfaceKind(), wall->deviceSurfaceParameter()),
                      "emcgpu_device_set_surface");
    }
    if (type.isInjected()) {
      if (type.deviceParticleKind() < 0)
        emcMessage::getInstance()
            .addError(type.getName() + " is injected at contacts but has no device creation rule (deviceParticleKind() < 0); "
                      "it cannot run on the GPU path and there is no CPU fallback.")
            .print();
      emcgpu::require(ctx, emcgpu_device_set_particle_kind(ctx, type.deviceParticleKind()), "emcgpu_device_set_particle_kind");
    }
  }
  ~emcBasicParticleHandler() override {
    if (ctx)
      emcgpu_destroy(ctx);
    if (comm)
      emcnccl_destroy(comm);
  }

  // --- several GPUs ---
  bool isSharded() const { return shardWorld > 1; }
  int shardRankOf() const { return shardRank; }
  int shardWorldSize() const { return shardWorld; }
  bool isShardRoot() const { return shardRank == 0; }
  // in-place sum over the ranks of a small host array (no-op on one GPU)
  void sumOverRanks(std::vector<double> &v) const {
    if (comm && !v.empty() && emcnccl_allreduce_sum_host_f64(comm, v.data(), static_cast<int64_t>(v.size())) != 0)
      emcMessage::getInstance().addError(std::string("sharded run: all-reduce failed: ") + emcnccl_last_error()).print();
  }
  void sumOverRanks(std::vector<int32_t> &v) const {
    if (!comm)
      return;
    std::vector<double> d(v.begin(), v.end());
    sumOverRanks(d);
    for (SizeType i = 0; i < v.size(); i++)
      v[i] = static_cast<int32_t>(d[i]);
  }
  // true if `data` is bit for bit the same on every rank (the replicated grids must be): the ranks compare 16-bit pieces
  // of a checksum through sum and sum of squares -- sum(p^2) * W == sum(p)^2 holds only if all p are equal (integers, exact)
  bool identicalOnAllRanks(const double *data, SizeType n) const {
    if (!comm)
      return true;
    uint64_t h = 1469598103934665603ull;
    const unsigned char *b = reinterpret_cast<const unsigned char *>(data);
    for (SizeType i = 0; i < n * sizeof(double); i++)
      h = (h ^ b[i]) * 1099511628211ull;
    std::vector<double> v(8);
    for (int k = 0; k < 4; k++) {
      const double p = static_cast<double>((h >> (16 * k)) & 0xffffu);
      v[k] = p;
      v[4 + k] = p * p;
    }
    sumOverRanks(v);
    for (int k = 0; k < 4; k++)
      if (v[4 + k] * shardWorld != v[k] * v[k])
        return false;
    return true;
  }
  int64_t allReduceCalls() const { return comm ? emcnccl_calls(comm) : 0; }
  int64_t allReduceBytes() const { return comm ? emcnccl_bytes(comm) : 0; }

  // the context that holds the ensemble and the grids of the run (emcSimulation, emcSORSolver::attach, tests)
  emcgpu_ctx *gpuContext() { return ctx; }
  SizeType gpuParticleType() const { return gpuType; }

  bool calcsPartPartInteraction() const override { return false; }
  // particles of the whole simulation (summed over the ranks of a sharded run)
  SizeType getNrParticles(SizeType idxType) const {
    if (idxType != gpuType)
      return 0;
    if (!comm || !uploaded)
      return nrOnGpu();
    std::vector<double> n(1, static_cast<double>(nrOnGpu()));
    sumOverRanks(n);
    return static_cast<SizeType>(n[0]);
  }
  SizeType getNrParticlesOfThisRank(SizeType idxType) const { return idxType == gpuType ? nrOnGpu() : 0; }
  void printNrParticles() const override {
    for (const auto &[idxType, type] : this->idxTypeToPartType)
      std::cout << "\t" << getNrParticles(idxType) << " " << type->getName() << "\n";
  }

  // cells in storage order; while (n >= 1) { add; n -= carriers per particle }; one more with probability n
  void generateInitialParticles(const emcGrid<T, Dim> &potential) override {
    auto &type = *this->idxTypeToPartType.at(gpuType);
    std::uniform_real_distribution<T> uniform(0., 1.);
    SizeVec coord;
    for (coord.fill(0); !this->device.isEndCoord(coord); this->device.advanceCoord(coord)) {
      auto toCreate = type.getInitialNrParticles(coord, this->device, potential);
      while (toCreate >= 1) {
        emcdetail::appendParticle(staging, type, this->device, coord, hostRng);
        toCreate -= this->nrCarriersPerPart;
      }
      if (uniform(hostRng) < toCreate)
        emcdetail::appendParticle(staging, type, this->device, coord, hostRng);
    }
    uploaded = false;
    upload();
  }

  NettoParticleCounter driftScatterParticles(T tStep, std::vector<emcGrid<T, Dim>> &eField) override {
    upload();
    for (SizeType d = 0; d < Dim; d++)
      emcgpu::require(ctx, emcgpu_device_set_grid(ctx, EMCGPU_GRID_EFIELD_X + static_cast<int>(d), eField[d].raw()),
                      "emcgpu_device_set_grid");
    auto netto = this->initNettoParticleCounter();
    std::vector<int32_t> removed(std::max<SizeType>(1, netto[gpuType].size()), 0);
    emcgpu::require(ctx, emcgpu_device_step(ctx, tStep, removed.data()), "emcgpu_device_step");
    for (SizeType c = 0; c < netto[gpuType].size(); c++)
      netto[gpuType][c] = removed[c];
    return netto;
  }

  void assignParticlesToMesh(SizeType idxType, emcGrid<T, Dim> &gridNrParticles) override {
    if (idxType != gpuType)
      return;
    upload();
    emcgpu::require(ctx, emcgpu_device_assign(ctx), "emcgpu_device_assign");
    std::vector<double> count(gridNrParticles.getSize());
    emcgpu::require(ctx, emcgpu_device_get_grid(ctx, EMCGPU_GRID_COUNT, count.data()), "emcgpu_device_get_grid");
    SizeType i = 0;
    for (auto &v : gridNrParticles)
      v += count[i++];
  }

  NettoParticleCounter handleOhmicContacts() override {
    upload();
    auto netto = this->initNettoParticleCounter();
    std::vector<int32_t> net(std::max<SizeType>(1, netto[gpuType].size()), 0);
    emcgpu::require(ctx, emcgpu_device_contacts(ctx, net.data(), nullptr, 0), "emcgpu_device_contacts");
    for (SizeType c = 0; c < netto[gpuType].size(); c++)
      netto[gpuType][c] = net[c];
    return netto;
  }

  // "<prefix><TypeName><suffix>.txt": box extent, then per particle: index, position, k, energy, sub-valley, valley, tau
  void print(std::string namePrefix, std::string nameSuffix) override {
    for (const auto &[idxType, type] : this->idxTypeToPartType) {
      // sharded run: every rank writes the particles it holds, "<...>.rank<r>.txt"
      std::ofstream os(namePrefix + type->getName() + nameSuffix + (comm ? ".rank" + std::to_string(shardRank) : std::string()) + ".txt");
      os << this->device.getMaxPos() << "\n";
      if (idxType != gpuType)
        continue;
      emcdetail::HostEnsemble h;
      const emcdetail::HostEnsemble *src = &staging;
      if (uploaded) {
        download(h);
        src = &h;
      }
      const SizeType n = src->size();
      for (SizeType i = 0; i < n; i++) {
        os << i << " " << src->stream[EMCGPU_X][i] << " " << src->stream[EMCGPU_Y][i];
        if (Dim > 2)
          os << " " << src->stream[EMCGPU_Z][i];
        os << " " << src->stream[EMCGPU_KX][i] << " " << src->stream[EMCGPU_KY][i] << " " << src->stream[EMCGPU_KZ][i] << " "
           << src->stream[EMCGPU_ENERGY][i] << " " << ((src->packed[i] >> 8) & 0xffu) << " " << (src->packed[i] & 0xffu)
           << " " << src->stream[EMCGPU_TAU][i];
        if (i + 1 < n)
          os << "\n";
      }
    }
  }

  // Ramo-Shockley estimate q * carriers * sum(v_x) / L over the particles with x0 <= x <= x1
  T getChannelDriftCurrent(SizeType idxType, T x0, T x1, T channelLength) const {
    if (idxType != gpuType)
      return 0;
    emcdetail::HostEnsemble h;
    download(h);
    const auto &type = *this->idxTypeToPartType.at(idxType);
    double sumVx = 0;
    for (SizeType i = 0; i < h.size(); i++) {
      if (h.stream[EMCGPU_X][i] < x0 || h.stream[EMCGPU_X][i] > x1)
        continue;
      const std::array<T, 3> k = {h.stream[EMCGPU_KX][i], h.stream[EMCGPU_KY][i], h.stream[EMCGPU_KZ][i]};
      sumVx += type.getValley(h.packed[i] & 0xffu)->getVelocity(k, h.stream[EMCGPU_ENERGY][i], (h.packed[i] >> 8) & 0xffu)[0];
    }
    std::vector<double> total(1, sumVx);
    sumOverRanks(total);
    return constants::q * this->nrCarriersPerPart * total[0] / channelLength;
  }
};

#endif
