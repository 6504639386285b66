// K1: the bulk time-step kernel.  One launch advances every particle of the
// shard by `nSteps` consecutive time steps (state stays in registers between
// them) and accumulates the per-step observables.
//
// Replaces, per particle and step (reference tree):
//   examples/bulkSimulation/basicBulkParticleHandler.hpp:195-213  moveParticles body
//   include/emcParticleDrift.hpp:12-36                            drift()
//   include/emcScatterHandler.hpp:148-170, :237-244               selection
//   include/ParticleType/emcParticleType.hpp:187-189              getNewTau
//   examples/bulkSimulation/basicBulkParticleHandler.hpp:289-347  the three observable passes
//
// Memory behaviour: each particle is one coalesced read and one coalesced
// write of 8 fp64 streams + one u32 stream (136 B per launch, independent of
// nSteps); the cumulative rate tables are staged into shared memory with one
// TMA bulk copy per CTA; the model constants live in shared memory; the
// random numbers are counter-based (0 B).
#pragma once
#include "emc_device.cuh"

namespace emc {

struct BulkParams {
  double *stream[EMCGPU_N_STREAMS];
  uint32_t *packed;
  int64_t n;
  int64_t idBase;
  const DevModel *model;
  const double *tables; // [set][level][stride]
  const DevMech *mechs;
  int32_t nMechTotal;
  int32_t tablesInSmem;
  double dt;
  int32_t nSteps;
  int64_t step0;
  Vec3 box, force, dir;
  // rng
  uint64_t seed;
  const uint64_t *draws;
  const int64_t *offsets;
  uint32_t *cursor;
  // outputs
  double *obs; // [nSteps][nValleys][3], accumulated with atomics
  long long *events;
  long long evCap;
  unsigned long long *evCount;
  int *status;
  BathView baths; // phonon baths of polar-optical mechanisms (counts == nullptr: none)
  // grain boundaries (emcGrainScatterMechanism): second free-flight clock per particle, nullptr = no grain mechanism
  double *grainTau;
  double *grainOut; // K1d out of place: where the flight kernel leaves the clocks
  double grainProb, grainTau0;
  // launch-uniform flight constants per valley (host-built: emcgpu.cu buildFlightConsts), read from the constant bank
  FlightConst fc[EMCGPU_MAX_VALLEYS];
  // split step (emc_bulk_split.cuh): per-particle byte "step at which the particle left the flight kernel" (0xFF: finished),
  // claim counter of the event kernel
  uint8_t *frozen;
  unsigned *claim;
  int32_t eventClaim; // particles per claim of a warp of the event kernel (multiple of 256)
  // flight kernel out of place (emcgpu_bulk_step_ahead: the input ensemble stays as it was): nullptr = in place
  double *streamOut[EMCGPU_N_STREAMS];
  uint32_t *packedOut;
  // per-particle velocities of every step (printDriftVelocities / printVelocities, basicBulkParticleHandler.hpp:251-285):
  // [nSteps][n][velComponents], velComponents = 1 (v.Ê) or 3 (v); nullptr = not recorded.  General step kernel only.
  double *velOut;
  int32_t velComponents;
};

// emcGrainScatterMechanism::scatterParticle (include/emcGrainScatterMechanism.hpp:40-77) + emcParticleType::getNewGrainTau
// (emcParticleType.hpp:191-193): with probability (1 - transmission) the particle is reflected into the opposite hemisphere
// about its k, otherwise transmitted into the same one; elastic; then a new exponential clock.  Returns the new clock.
template <int RNG_MODE> __device__ __forceinline__ double grainEvent(const BulkParams &P, Particle &p, Rng &rng) {
  double rnd;
  if (uniform01(rng.raw<RNG_MODE>()) > P.grainProb) {
    rnd = uniform01(rng.raw<RNG_MODE>());
    if (rnd < 0.5) rnd += 0.5;
  } else {
    rnd = uniform01(rng.raw<RNG_MODE>());
    if (rnd > 0.5) rnd -= 0.5;
  }
  p.k = randomDirectionWrtK<true>(p.k, 1.0 - 2.0 * rnd, uniform01(rng.raw<RNG_MODE>()));
  return __dmul_rn(-log(uniformLog(rng.raw<RNG_MODE>())), P.grainTau0);
}

constexpr int kBulkThreads = 256;
#ifndef EMC_STREAM_THREADS
#define EMC_STREAM_THREADS 256
#endif
constexpr int kStreamThreads = EMC_STREAM_THREADS;
constexpr int kMaxStepsPerLaunch = 64;

// ---- TMA bulk copy global -> shared, completion on an mbarrier ------------
__device__ __forceinline__ uint32_t smemAddr(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbarInit(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbarTryWait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n"
               ".reg .pred p;\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
               "selp.u32 %0, 1, 0, p;\n"
               "}"
               : "=r"(ok)
               : "r"(smemAddr(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
  while (!mbarTryWait(bar, parity)) {
  }
}
__device__ __forceinline__ void mbarArrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void tmaBulkLoad(void *dstSmem, const void *srcGlobal, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemAddr(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}
// same with an L2 cache policy (streaming data: evict first)
__device__ __forceinline__ uint64_t l2EvictFirstPolicy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tmaBulkLoadHint(void *dstSmem, const void *srcGlobal, uint32_t bytes, uint64_t *bar,
                                                uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smemAddr(dstSmem)),
      "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)), "l"(policy)
      : "memory");
}
// TMA bulk copy shared -> global, completion tracked by the issuing thread's bulk groups
__device__ __forceinline__ void tmaBulkStore(void *dstGlobal, const void *srcSmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(smemAddr(srcSmem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tmaCommitGroup() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tmaWaitGroupRead() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void tmaWaitGroup() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fenceProxyAsyncShared() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ double warpSum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Shared-memory layout of the step kernels (host and device agree through this).
struct BulkSmem {
  size_t model, obs, mechs, fastV, fast, queue, tables, total;
  __host__ __device__ BulkSmem(int nObsDoubles, int nValleys, int nMechTotal, int64_t tableDoubles, bool tablesInSmem,
                               int queueWords) {
    auto up = [](size_t x) { return (x + 15) & ~size_t(15); };
    size_t off = 0;
    model = off; off = up(off + sizeof(DevModel));
    obs = off; off = up(off + (size_t)nObsDoubles * sizeof(double));
    mechs = off; off = up(off + (size_t)nMechTotal * sizeof(DevMech));
    fastV = off; off = up(off + (size_t)nValleys * sizeof(FastValley));
    fast = off; off = up(off + (size_t)nValleys * EMCGPU_MAX_SUBVALLEYS * sizeof(FastSub));
    queue = off; off = up(off + (size_t)queueWords * sizeof(uint32_t));
    tables = off; if (tablesInSmem) off = up(off + (size_t)tableDoubles * sizeof(double));
    total = off;
  }
};

struct CtaState {
  const DevModel *model;
  double *obs;
  const DevMech *mechs;
  const FastValley *fastV;
  const FastSub *fast;
  uint32_t *queue;
  const double *tables;
};

// Kernel prologue: TMA bulk copy of the cumulative tables into shared memory,
// cooperative copy of the model, construction of the fast-path constants.
__device__ __forceinline__ CtaState stageCta(const BulkParams &P, unsigned char *smemRaw, uint64_t *tableBar,
                                             int nObsDoubles, int queueWords) {
  const int nV = P.model->nValleys;
  const BulkSmem L(nObsDoubles, nV, P.nMechTotal, P.model->tableDoubles, P.tablesInSmem != 0, queueWords);
  DevModel *sModel = reinterpret_cast<DevModel *>(smemRaw + L.model);
  double *sObs = reinterpret_cast<double *>(smemRaw + L.obs);
  DevMech *sMechs = reinterpret_cast<DevMech *>(smemRaw + L.mechs);
  FastValley *sFastV = reinterpret_cast<FastValley *>(smemRaw + L.fastV);
  FastSub *sFast = reinterpret_cast<FastSub *>(smemRaw + L.fast);
  double *sTables = reinterpret_cast<double *>(smemRaw + L.tables);
  const int tid = threadIdx.x;
  const uint32_t tableBytes = (uint32_t)(P.model->tableDoubles * sizeof(double));
  if (P.tablesInSmem) {
    if (tid == 0) {
      mbarInit(tableBar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbarExpectTx(tableBar, tableBytes);
      const uint32_t chunk = 32768u; // each a multiple of 16 B
      for (uint32_t o = 0; o < tableBytes; o += chunk) {
        const uint32_t b = min(chunk, tableBytes - o);
        tmaBulkLoad(reinterpret_cast<unsigned char *>(sTables) + o,
                    reinterpret_cast<const unsigned char *>(P.tables) + o, b, tableBar);
      }
    }
  }
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(P.model);
    uint32_t *dst = reinterpret_cast<uint32_t *>(sModel);
    for (int i = tid; i < (int)(sizeof(DevModel) / 4); i += blockDim.x) dst[i] = src[i];
    const uint32_t *msrc = reinterpret_cast<const uint32_t *>(P.mechs);
    uint32_t *mdst = reinterpret_cast<uint32_t *>(sMechs);
    for (int i = tid; i < (int)(P.nMechTotal * sizeof(DevMech) / 4); i += blockDim.x) mdst[i] = msrc[i];
    for (int i = tid; i < nObsDoubles; i += blockDim.x) sObs[i] = 0.0;
  }
  __syncthreads();
  for (int i = tid; i < nV * EMCGPU_MAX_SUBVALLEYS; i += blockDim.x) {
    const DevValley &v = sModel->valleys[i / EMCGPU_MAX_SUBVALLEYS];
    const int s = i % EMCGPU_MAX_SUBVALLEYS;
    buildFastSub(v, s < v.deg ? s : 0, P.force, P.dir, P.dt, sFast[i]);
    if (s == 0) sFastV[i / EMCGPU_MAX_SUBVALLEYS].f = P.fc[i / EMCGPU_MAX_SUBVALLEYS];
  }
  __syncthreads();
  if (P.tablesInSmem) mbarWait(tableBar, 0);
  CtaState c;
  c.model = sModel;
  c.obs = sObs;
  c.mechs = sMechs;
  c.fastV = sFastV;
  c.fast = sFast;
  c.queue = reinterpret_cast<uint32_t *>(smemRaw + L.queue);
  c.tables = P.tablesInSmem ? sTables : P.tables;
  return c;
}

// One free flight of duration t followed by the periodic wrap (drift() emcParticleDrift.hpp:12-36 + the wrap of
// basicBulkParticleHandler.hpp:600-613).  FAST arithmetic in a valley with signed-permutation rotations: the division-free
// flight of emc_device.cuh (leaves r = 1/S of the new state in rInv); otherwise the reference's operation order.
template <bool EXACT>
__device__ __forceinline__ void flightAndWrap(const CtaState &C, const BulkParams &P, Particle &p, double t, double &rInv) {
  if constexpr (!EXACT) {
    const FlightConst &f = C.fastV[p.valley].f;
    if (f.diag) {
      const FastSub &fs = C.fast[p.valley * EMCGPU_MAX_SUBVALLEYS + p.sub];
      FlightAux o;
      flightCore(fs.a[0], fs.a[1], fs.a[2], f.Fh[0] * t, f.Fh[1] * t, f.Fh[2] * t, f.KP * t, f.c2a, p.k.x, p.k.y, p.k.z,
                 p.pos.x, p.pos.y, p.pos.z, o);
      if (mayNeedWrap(p.pos.x, (uint32_t)__double2hiint(P.box.x)) || mayNeedWrap(p.pos.y, (uint32_t)__double2hiint(P.box.y)) ||
          mayNeedWrap(p.pos.z, (uint32_t)__double2hiint(P.box.z))) {
        p.pos.x = wrapExact(p.pos.x, P.box.x);
        p.pos.y = wrapExact(p.pos.y, P.box.y);
        p.pos.z = wrapExact(p.pos.z, P.box.z);
      }
      p.energy = flightEnergy(f.fE, o);
      rInv = o.r;
      return;
    }
  }
  drift<EXACT, 3>(C.model->valleys[p.valley], p, t, P.force);
  p.pos.x = wrap1<EXACT>(p.pos.x, P.box.x);
  p.pos.y = wrap1<EXACT>(p.pos.y, P.box.y);
  p.pos.z = wrap1<EXACT>(p.pos.z, P.box.z);
}

// The scattering part of a time step (A.5 of SURVEY.md): entered with the
// particle already drifted to its first scattering time; tRem = dt - tau_old.
template <bool EXACT, int RNG_MODE>
__device__ __forceinline__ void scatterLoop(const CtaState &C, const BulkParams &P, Particle &p, Rng &rng,
                                            int64_t particleId, int64_t step, double tRem, double &rInv) {
  using A = Arith<EXACT>;
  const DevModel &model = *C.model;
  while (tRem > 0.0) {
    const int set = (p.region < kMaxRegions) ? model.setOf[p.valley][p.region] : -1;
    double tauTab = model.defaultTau;
    if (set >= 0) {
      const DevTableSet &ts = model.sets[set];
      const int lvl = energyLevel(p.energy, model.dE, model.nLevels);
      const double r = uniform01(rng.raw<RNG_MODE>());
      const double *row = C.tables + ts.tabOffset + (int64_t)lvl * ts.stride;
      const int m = selectMechanism(row, ts.nMech, r);
      int mechId = -1;
      if (m >= 0) {
        const DevMech &mech = C.mechs[ts.mechOffset + m];
        mechId = mech.mechId;
        sampleFinalState<EXACT, RNG_MODE>(model, mech, p, rng, P.baths);
      }
      if (P.evCap > 0) {
        const unsigned long long e = atomicAdd(P.evCount, 1ull);
        if ((long long)e < P.evCap) {
          long long *dst = P.events + 4 * e;
          dst[0] = step;
          dst[1] = particleId;
          dst[2] = m;
          dst[3] = mechId;
        }
      }
    }
    // getNewTau with the (possibly new) valley and the unchanged region
    // (emcParticleType.hpp:187-189)
    {
      const int set2 = (p.region < kMaxRegions) ? model.setOf[p.valley][p.region] : -1;
      if (set2 >= 0) tauTab = model.sets[set2].tau;
    }
    const double newTau = A::mul(-log(uniformLog(rng.raw<RNG_MODE>())), tauTab);
    p.tau = A::add(p.tau, newTau);
    flightAndWrap<EXACT>(C, P, p, fmin(tRem, newTau), rInv);
    tRem = A::sub(tRem, newTau);
  }
}

// One full time step of one particle, any case (A.5 of SURVEY.md).  Returns
// v.Ê of the final state for the drift-velocity observable.
template <bool EXACT, int RNG_MODE>
__device__ __forceinline__ double bulkParticleStep(const CtaState &C, const BulkParams &P, Particle &p, Rng &rng,
                                                   int64_t particleId, int64_t step) {
  using A = Arith<EXACT>;
  const double dt = P.dt;
  if constexpr (!EXACT) {
    if (p.tau >= dt) // no scattering in this step
      return fastStep(C.fast[p.valley * EMCGPU_MAX_SUBVALLEYS + p.sub], C.fastV[p.valley], dt, P.box, p.k.x, p.k.y,
                      p.k.z, p.energy, p.tau, p.pos.x, p.pos.y, p.pos.z);
  }
  double rInv = 0.0;
  flightAndWrap<EXACT>(C, P, p, fmin(p.tau, dt), rInv);
  scatterLoop<EXACT, RNG_MODE>(C, P, p, rng, particleId, step, A::sub(dt, p.tau), rInv);
  p.tau = A::sub(p.tau, dt);
  if constexpr (!EXACT) {
    const FlightConst &f = C.fastV[p.valley].f;
    if (f.diag) // the step always ends with a flight: rInv belongs to the final state
      return flightVelocity(f, C.fast[p.valley * EMCGPU_MAX_SUBVALLEYS + p.sub], p.k.x, p.k.y, p.k.z, rInv);
  }
  return driftVelocity<EXACT>(C.model->valleys[p.valley], p.sub, p.k, p.energy, P.dir);
}

// Out-of-line copy of the complete step for the rare paths of the pipelined
// kernel (keeps their register needs away from the streaming loop).
template <bool EXACT, int RNG_MODE>
__device__ __noinline__ double bulkParticleStepNI(const CtaState &C, const BulkParams &P, Particle &p, Rng &rng,
                                                  int64_t particleIndex) {
  const uint64_t id = (uint64_t)(P.idBase + particleIndex);
  rng.k0 = (uint32_t)P.seed;
  rng.k1 = (uint32_t)(P.seed >> 32);
  rng.idLo = (uint32_t)id;
  rng.idHi = (uint32_t)(id >> 32);
  rng.status = P.status;
  rng.n = 0;
  rng.step = (uint32_t)P.step0;
  if constexpr (RNG_MODE == RNG_REPLAY) {
    rng.stream = P.draws + P.offsets[particleIndex] + P.cursor[particleIndex];
    rng.streamEnd = P.draws + P.offsets[particleIndex + 1];
  }
  const double vd = bulkParticleStep<EXACT, RNG_MODE>(C, P, p, rng, P.idBase + particleIndex, P.step0);
  if constexpr (RNG_MODE == RNG_REPLAY)
    P.cursor[particleIndex] = (uint32_t)(rng.stream - (P.draws + P.offsets[particleIndex]));
  return vd;
}

// Per-valley partial sums of one warp into the CTA's shared accumulators
// (basicBulkParticleHandler.hpp:289-347).  Whole warp must call.
__device__ __forceinline__ void accumulateObsWarp(double *o, int nV, bool live, int valley, double e, double vd) {
  const int lane = threadIdx.x & 31;
  for (int v = 0; v < nV; v++) {
    const bool mine = live && valley == v;
    const unsigned cnt = __popc(__ballot_sync(0xffffffffu, mine));
    if (cnt == 0) continue; // warp-uniform
    const double se = warpSum(mine ? e : 0.0), sv = warpSum(mine ? vd : 0.0);
    if (lane == 0) {
      atomicAdd(o + 3 * v + 0, se);
      atomicAdd(o + 3 * v + 1, sv);
      atomicAdd(o + 3 * v + 2, (double)cnt);
    }
  }
}

__device__ __forceinline__ void loadParticle(const BulkParams &P, int64_t i, Particle &p, Rng &rng) {
  p.k.x = P.stream[EMCGPU_KX][i];
  p.k.y = P.stream[EMCGPU_KY][i];
  p.k.z = P.stream[EMCGPU_KZ][i];
  p.energy = 0.0; // never read: every step starts with a drift, which recomputes E from k (emcParticleDrift.hpp:25)
  p.tau = P.stream[EMCGPU_TAU][i];
  p.pos.x = P.stream[EMCGPU_X][i];
  p.pos.y = P.stream[EMCGPU_Y][i];
  p.pos.z = P.stream[EMCGPU_Z][i];
  const uint32_t w = P.packed[i];
  p.valley = w & 0xffu;
  p.sub = (w >> 8) & 0xffu;
  p.region = w >> 16;
  const uint64_t id = (uint64_t)(P.idBase + i);
  rng.k0 = (uint32_t)P.seed;
  rng.k1 = (uint32_t)(P.seed >> 32);
  rng.idLo = (uint32_t)id;
  rng.idHi = (uint32_t)(id >> 32);
  rng.status = P.status;
  rng.n = 0;
}
template <int RNG_MODE> __device__ __forceinline__ void attachReplay(const BulkParams &P, int64_t i, Rng &rng) {
  if constexpr (RNG_MODE == RNG_REPLAY) {
    rng.stream = P.draws + P.offsets[i] + P.cursor[i];
    rng.streamEnd = P.draws + P.offsets[i + 1];
  }
}
__device__ __forceinline__ void storeParticleState(const BulkParams &P, int64_t i, const Particle &p) {
  P.stream[EMCGPU_KX][i] = p.k.x;
  P.stream[EMCGPU_KY][i] = p.k.y;
  P.stream[EMCGPU_KZ][i] = p.k.z;
  P.stream[EMCGPU_ENERGY][i] = p.energy;
  P.stream[EMCGPU_TAU][i] = p.tau;
  P.stream[EMCGPU_X][i] = p.pos.x;
  P.stream[EMCGPU_Y][i] = p.pos.y;
  P.stream[EMCGPU_Z][i] = p.pos.z;
  P.packed[i] = (uint32_t)p.valley | ((uint32_t)p.sub << 8) | ((uint32_t)p.region << 16);
}
template <int RNG_MODE>
__device__ __forceinline__ void storeParticle(const BulkParams &P, int64_t i, const Particle &p, const Rng &rng) {
  P.stream[EMCGPU_KX][i] = p.k.x;
  P.stream[EMCGPU_KY][i] = p.k.y;
  P.stream[EMCGPU_KZ][i] = p.k.z;
  P.stream[EMCGPU_ENERGY][i] = p.energy;
  P.stream[EMCGPU_TAU][i] = p.tau;
  P.stream[EMCGPU_X][i] = p.pos.x;
  P.stream[EMCGPU_Y][i] = p.pos.y;
  P.stream[EMCGPU_Z][i] = p.pos.z;
  P.packed[i] = (uint32_t)p.valley | ((uint32_t)p.sub << 8) | ((uint32_t)p.region << 16);
  if constexpr (RNG_MODE == RNG_REPLAY) P.cursor[i] = (uint32_t)(rng.stream - (P.draws + P.offsets[i]));
}

// ---------------------------------------------------------------------------
// K1a: ONE time step per launch, streaming.  Every lane moves VEC consecutive
// particles per iteration with 16/32-byte vector loads and stores.  Particles
// that scatter in this step (tau < dt, ~1 % in Si at dt = 1e-16 s) are not
// processed in place -- that would make the whole warp walk the long event
// path at 1/32 lane utilisation -- but queued (index only) in a per-warp
// shared-memory queue and processed 32 at a time by the full warp.
template <int VEC> struct VecIO;
template <> struct VecIO<1> {
  static __device__ __forceinline__ void ld(const double *p, double (&v)[1]) { v[0] = __ldcs(p); }
  static __device__ __forceinline__ void st(double *p, const double (&v)[1]) { __stcs(p, v[0]); }
  static __device__ __forceinline__ void ldw(const uint32_t *p, uint32_t (&v)[1]) { v[0] = __ldcs(p); }
  static __device__ __forceinline__ void stw(uint32_t *p, const uint32_t (&v)[1]) { __stcs(p, v[0]); }
};
template <> struct VecIO<2> {
  static __device__ __forceinline__ void ld(const double *p, double (&v)[2]) {
    const double2 t = __ldcs(reinterpret_cast<const double2 *>(p));
    v[0] = t.x; v[1] = t.y;
  }
  static __device__ __forceinline__ void st(double *p, const double (&v)[2]) {
    __stcs(reinterpret_cast<double2 *>(p), make_double2(v[0], v[1]));
  }
  static __device__ __forceinline__ void ldw(const uint32_t *p, uint32_t (&v)[2]) {
    const uint2 t = __ldcs(reinterpret_cast<const uint2 *>(p));
    v[0] = t.x; v[1] = t.y;
  }
  static __device__ __forceinline__ void stw(uint32_t *p, const uint32_t (&v)[2]) {
    __stcs(reinterpret_cast<uint2 *>(p), make_uint2(v[0], v[1]));
  }
};
template <> struct VecIO<4> {
  // 256-bit global accesses (sm_100+)
  static __device__ __forceinline__ void ld(const double *p, double (&v)[4]) {
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
  }
  static __device__ __forceinline__ void st(double *p, const double (&v)[4]) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3])
                 : "memory");
  }
  static __device__ __forceinline__ void ldw(const uint32_t *p, uint32_t (&v)[4]) {
    const uint4 t = __ldcs(reinterpret_cast<const uint4 *>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void stw(uint32_t *p, const uint32_t (&v)[4]) {
    __stcs(reinterpret_cast<uint4 *>(p), make_uint4(v[0], v[1], v[2], v[3]));
  }
};



template <bool EXACT, int RNG_MODE>
__device__ __noinline__ void processQueued(const CtaState &C, const BulkParams &P, uint32_t idx, bool active) {
  Particle p;
  Rng rng;
  double e = 0.0, vd = 0.0;
  p.valley = 0;
  if (active) {
    loadParticle(P, idx, p, rng);
    attachReplay<RNG_MODE>(P, idx, rng);
    rng.step = (uint32_t)P.step0;
    vd = bulkParticleStep<EXACT, RNG_MODE>(C, P, p, rng, P.idBase + idx, P.step0);
    e = p.energy;
    storeParticle<RNG_MODE>(P, idx, p, rng);
  }
  __syncwarp();
  accumulateObsWarp(C.obs, C.model->nValleys, active, p.valley, e, vd);
}

template <bool EXACT, int RNG_MODE, int VEC>
__global__ void __launch_bounds__(kStreamThreads, 2) bulkStreamKernel(const __grid_constant__ BulkParams P) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint64_t tableBar;
  constexpr int QW = 32 + 32 * VEC;
  const int nV = P.model->nValleys;
  const CtaState C = stageCta(P, smemRaw, &tableBar, nV * 3, (kStreamThreads / 32) * QW);
  const DevModel &model = *C.model;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *q = C.queue + warp * QW;
  int qn = 0; // warp-uniform
  const bool single = nV == 1;
  double accE = 0.0, accV = 0.0;
  unsigned accN = 0;
  const double dt = P.dt;

  const int64_t nGroups = P.n / VEC;
  const int64_t nRounded = (nGroups + 31) & ~int64_t(31);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + tid; g < nRounded; g += stride) {
    const bool live = g < nGroups;
    const int64_t i0 = g * VEC;
    double s[EMCGPU_N_STREAMS][VEC];
    uint32_t w[VEC];
    bool ev[VEC];
    double eOut[VEC], vOut[VEC];
#pragma unroll
    for (int j = 0; j < VEC; j++) ev[j] = false;
    if (live) {
#pragma unroll
      for (int c = 0; c < EMCGPU_N_STREAMS; c++) {
        if (c == EMCGPU_ENERGY) { // write-only stream, see loadParticle
#pragma unroll
          for (int j = 0; j < VEC; j++) s[c][j] = 0.0;
        } else {
          VecIO<VEC>::ld(P.stream[c] + i0, s[c]);
        }
      }
      VecIO<VEC>::ldw(P.packed + i0, w);
#pragma unroll
      for (int j = 0; j < VEC; j++) {
        eOut[j] = 0.0;
        vOut[j] = 0.0;
        if (s[EMCGPU_TAU][j] >= dt) {
          const int valley = w[j] & 0xffu, sub = (w[j] >> 8) & 0xffu;
          if constexpr (EXACT) {
            Particle p;
            p.k = Vec3{s[EMCGPU_KX][j], s[EMCGPU_KY][j], s[EMCGPU_KZ][j]};
            p.energy = s[EMCGPU_ENERGY][j];
            p.tau = s[EMCGPU_TAU][j];
            p.pos = Vec3{s[EMCGPU_X][j], s[EMCGPU_Y][j], s[EMCGPU_Z][j]};
            p.valley = valley;
            p.sub = sub;
            const DevValley &v = model.valleys[valley];
            drift<true, 3>(v, p, dt, P.force);
            s[EMCGPU_KX][j] = p.k.x;
            s[EMCGPU_KY][j] = p.k.y;
            s[EMCGPU_KZ][j] = p.k.z;
            s[EMCGPU_ENERGY][j] = p.energy;
            s[EMCGPU_TAU][j] = __dsub_rn(p.tau, dt);
            s[EMCGPU_X][j] = wrap1<true>(p.pos.x, P.box.x);
            s[EMCGPU_Y][j] = wrap1<true>(p.pos.y, P.box.y);
            s[EMCGPU_Z][j] = wrap1<true>(p.pos.z, P.box.z);
            vOut[j] = driftVelocity<true>(v, sub, p.k, p.energy, P.dir);
          } else {
            vOut[j] = fastStep(C.fast[valley * EMCGPU_MAX_SUBVALLEYS + sub], C.fastV[valley], dt, P.box,
                               s[EMCGPU_KX][j], s[EMCGPU_KY][j], s[EMCGPU_KZ][j], s[EMCGPU_ENERGY][j],
                               s[EMCGPU_TAU][j], s[EMCGPU_X][j], s[EMCGPU_Y][j], s[EMCGPU_Z][j]);
          }
          eOut[j] = s[EMCGPU_ENERGY][j];
          if (single) {
            accE += eOut[j];
            accV += vOut[j];
            accN++;
          }
        } else {
          ev[j] = true;
        }
      }
      // scattering particles are written back unchanged here and rewritten by
      // processQueued (same warp, ordered by __syncwarp)
#pragma unroll
      for (int c = 0; c < EMCGPU_N_STREAMS; c++) VecIO<VEC>::st(P.stream[c] + i0, s[c]);
    }
    if (!single) {
#pragma unroll
      for (int j = 0; j < VEC; j++)
        accumulateObsWarp(C.obs, nV, live && !ev[j], live ? (int)(w[j] & 0xffu) : 0, live ? eOut[j] : 0.0,
                          live ? vOut[j] : 0.0);
    }
#pragma unroll
    for (int j = 0; j < VEC; j++) {
      const unsigned mask = __ballot_sync(0xffffffffu, ev[j]);
      if (ev[j]) q[qn + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(i0 + j);
      qn += __popc(mask);
    }
    __syncwarp();
    while (qn >= 32) {
      qn -= 32;
      const uint32_t idx = q[qn + lane];
      __syncwarp();
      processQueued<EXACT, RNG_MODE>(C, P, idx, true);
    }
  }
  // the n % VEC particles past the last whole group
  if (blockIdx.x == 0 && warp == 0) {
    const int tail = (int)(P.n - nGroups * VEC);
    if (lane < tail) q[qn + lane] = (uint32_t)(nGroups * VEC + lane);
    qn += tail;
  }
  __syncwarp();
  while (qn > 0) {
    const int take = min(32, qn);
    qn -= take;
    const bool active = lane < take;
    const uint32_t idx = active ? q[qn + lane] : 0u;
    __syncwarp();
    processQueued<EXACT, RNG_MODE>(C, P, idx, active);
  }
  if (single) {
    const double se = warpSum(accE), sv = warpSum(accV);
    const unsigned cnt = __reduce_add_sync(0xffffffffu, accN);
    if (lane == 0 && cnt) {
      atomicAdd(C.obs + 0, se);
      atomicAdd(C.obs + 1, sv);
      atomicAdd(C.obs + 2, (double)cnt);
    }
  }
  __syncthreads();
  for (int j = tid; j < nV * 3; j += blockDim.x) {
    const double v = C.obs[j];
    if (v != 0.0) atomicAdd(P.obs + j, v);
  }
}

// ---------------------------------------------------------------------------
// K1a/TMA: ONE time step per launch as a warp-specialised TMA pipeline, one
// persistent CTA per SM.
//
//   loader warp (1 lane)       tile t -> ring stage t % S:  9 cp.async.bulk loads
//                              (one per SoA stream, L2 evict-first) completing on
//                              full[stage], as soon as the stage's previous
//                              tenant has been stored (freed[stage]).
//   storer warp (1 lane)       when the consumers are done with a stage
//                              (done[stage]): 9 cp.async.bulk stores back to the
//                              same global addresses; publishes how many tiles
//                              are complete in global memory (storesDone).
//   10 consumer warps          claim 64-particle sub-tiles dynamically, update
//                              them in place in shared memory (fastStep), arrive
//                              on done[stage].
//   scattering particles       (tau < dt) are copied -- state and all -- into a
//                              CTA-wide shared-memory event queue and left
//                              untouched in the tile.  Whenever 32 are queued a
//                              consumer warp claims the batch, runs the complete
//                              scattering step with all lanes busy and writes the
//                              result straight to global memory once the store of
//                              the tile they came from has completed (so the two
//                              writes cannot be reordered; the lines are still in
//                              L2, so no extra DRAM traffic).  Sub-tiles in which
//                              >= 25 % of the particles scatter (dt >~ tau
//                              regimes) are processed in place instead.
//
// Loads run S-1 tiles ahead of the arithmetic independent of occupancy and
// register pressure; dynamic sub-tile claiming keeps the pipeline moving while
// some warp is busy with an event batch.
#ifndef EMC_TMA_CONSUMER_WARPS
#define EMC_TMA_CONSUMER_WARPS 10
#endif
constexpr int kTmaConsumerWarps = EMC_TMA_CONSUMER_WARPS;
constexpr int kTmaThreads = (kTmaConsumerWarps + 2) * 32; // + loader warp + storer warp
constexpr int kTile = 256; // particles per tile
constexpr int kSubTile = 64;
constexpr int kSubsPerTile = kTile / kSubTile;
constexpr int kTileBytes = kTile * (EMCGPU_N_STREAMS * 8 + 4);
constexpr int kQueueCap = 256;
constexpr int kMaxStages = 16;
constexpr int kMinStages = 3;
// bytes the loader brings in per tile: every stream except the energy, which
// the step never reads (each step starts with a drift that recomputes E from k)
constexpr int kTileLoadBytes = kTile * ((EMCGPU_N_STREAMS - 1) * 8 + 4);
constexpr int kDenseEvents = kSubTile / 4;

constexpr int kFreeLag = 0;  // bulk stores whose shared-memory reads may still be pending
constexpr int kStoreLag = 6; // bulk stores allowed in flight before their completion is awaited

struct TmaControl {
  uint64_t full[kMaxStages];
  uint64_t done[kMaxStages];
  uint64_t freed[kMaxStages];
  uint64_t tableBar;
  unsigned nextSub;
  unsigned qTail, qHead;
  unsigned storesDone;  // number of this CTA's tiles whose store is complete
  unsigned loadsIssued; // number of this CTA's tiles whose loads have been issued
  unsigned dirty[kMaxStages]; // the packed index words of the stage were modified (dense sub-tiles)
};
struct EventQueue {
  double f[EMCGPU_N_STREAMS][kQueueCap];
  uint32_t w[kQueueCap];
  uint32_t idx[kQueueCap];
  uint32_t tile[kQueueCap];
  uint32_t flag[kQueueCap];
};
constexpr int kTmaQueueWords = (int)((sizeof(TmaControl) + sizeof(EventQueue) + 3) / 4);

__host__ __device__ inline size_t tmaRingOffset(const BulkSmem &L) { return (L.total + 127) & ~size_t(127); }

__device__ __forceinline__ void storeCursor(const BulkParams &P, int64_t i, const Rng &rng) {
  P.cursor[i] = (uint32_t)(rng.stream - (P.draws + P.offsets[i]));
}

// Scattering step of up to 32 queued particles, all lanes busy.
template <bool EXACT, int RNG_MODE>
__device__ __noinline__ void processEventBatch(const CtaState &C, const BulkParams &P, TmaControl *ctl,
                                               EventQueue *Q, unsigned head, int count) {
  const int lane = threadIdx.x & 31;
  const bool active = lane < count;
  Particle p;
  Rng rng;
  double e = 0.0, vd = 0.0;
  uint32_t idx = 0, tile = 0;
  p.valley = 0;
  if (active) {
    const unsigned pos = head + lane, slot = pos % kQueueCap;
    const uint32_t epoch = pos / kQueueCap + 1;
    while (*reinterpret_cast<const volatile uint32_t *>(&Q->flag[slot]) != 2 * epoch - 1) { // full, this epoch
    }
    __threadfence_block();
    p.k = Vec3{Q->f[EMCGPU_KX][slot], Q->f[EMCGPU_KY][slot], Q->f[EMCGPU_KZ][slot]};
    p.energy = Q->f[EMCGPU_ENERGY][slot];
    p.tau = Q->f[EMCGPU_TAU][slot];
    p.pos = Vec3{Q->f[EMCGPU_X][slot], Q->f[EMCGPU_Y][slot], Q->f[EMCGPU_Z][slot]};
    const uint32_t w = Q->w[slot];
    idx = Q->idx[slot];
    tile = Q->tile[slot];
    __threadfence_block();
    *const_cast<volatile uint32_t *>(&Q->flag[slot]) = 2 * epoch; // consumed: the slot may be refilled
    p.valley = w & 0xffu;
    p.sub = (w >> 8) & 0xffu;
    p.region = w >> 16;
    vd = bulkParticleStepNI<EXACT, RNG_MODE>(C, P, p, rng, idx);
    e = p.energy;
  }
  __syncwarp();
  // the tile this particle sits in is written back unchanged by the storer warp;
  // our write must come after that store has completed
  const unsigned newest = __reduce_max_sync(0xffffffffu, tile);
  while (*reinterpret_cast<volatile unsigned *>(&ctl->storesDone) <= newest) {
  }
  __threadfence_block();
  if (active) storeParticleState(P, idx, p);
  accumulateObsWarp(C.obs, C.model->nValleys, active, p.valley, e, vd);
}

template <bool EXACT, int RNG_MODE>
__global__ void __launch_bounds__(kTmaThreads, 1) bulkTmaKernel(const __grid_constant__ BulkParams P, const int stages) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int nV = P.model->nValleys;
  const BulkSmem L(nV * 3, nV, P.nMechTotal, P.model->tableDoubles, P.tablesInSmem != 0, kTmaQueueWords);
  TmaControl *ctl = reinterpret_cast<TmaControl *>(smemRaw + L.queue);
  EventQueue *Q = reinterpret_cast<EventQueue *>(smemRaw + L.queue + sizeof(TmaControl));
  unsigned char *ring = smemRaw + tmaRingOffset(L);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < stages; s++) {
      mbarInit(&ctl->full[s], 1);
      mbarInit(&ctl->done[s], kSubsPerTile);
      mbarInit(&ctl->freed[s], 1);
    }
    ctl->nextSub = 0;
    ctl->qTail = 0;
    ctl->qHead = 0;
    ctl->storesDone = 0;
    ctl->loadsIssued = 0;
    for (int s = 0; s < stages; s++) ctl->dirty[s] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fenceProxyAsyncShared();
  }
  for (int i = tid; i < kQueueCap; i += blockDim.x) Q->flag[i] = 0;
  const CtaState C = stageCta(P, smemRaw, &ctl->tableBar, nV * 3, kTmaQueueWords); // fences + __syncthreads inside
  const DevModel &model = *C.model;

  const int64_t nTiles = P.n / kTile;
  const int myTiles = blockIdx.x < nTiles ? (int)((nTiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
  const double dt = P.dt;
  const bool single = nV == 1;
  double accE = 0.0, accV = 0.0;
  unsigned accN = 0;

  if (warp == kTmaConsumerWarps) {
    // ----------------------------- loader warp -----------------------------
    if (lane == 0) {
      const uint64_t pol = l2EvictFirstPolicy();
      for (int t = 0; t < myTiles; t++) {
        const int st = t % stages;
        if (t >= stages) mbarWait(&ctl->freed[st], (t / stages - 1) & 1); // the previous tenant has been stored
        const int64_t i0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * kTile;
        unsigned char *dst = ring + (size_t)st * kTileBytes;
        mbarExpectTx(&ctl->full[st], kTileLoadBytes);
#pragma unroll
        for (int c = 0; c < EMCGPU_N_STREAMS; c++)
          if (c != EMCGPU_ENERGY)
            tmaBulkLoadHint(dst + c * kTile * 8, P.stream[c] + i0, kTile * 8, &ctl->full[st], pol);
        tmaBulkLoadHint(dst + EMCGPU_N_STREAMS * kTile * 8, P.packed + i0, kTile * 4, &ctl->full[st], pol);
        // full[st] is now in the phase of tile t: consumers may wait on its parity
        *reinterpret_cast<volatile unsigned *>(&ctl->loadsIssued) = (unsigned)(t + 1);
      }
    }
  } else if (warp == kTmaConsumerWarps + 1) {
    // ----------------------------- storer warp -----------------------------
    if (lane == 0) {
      for (int t = 0; t < myTiles; t++) {
        const int st = t % stages;
        mbarWait(&ctl->done[st], (t / stages) & 1);
        const int64_t i0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * kTile;
        const unsigned char *src = ring + (size_t)st * kTileBytes;
#pragma unroll
        for (int c = 0; c < EMCGPU_N_STREAMS; c++) tmaBulkStore(P.stream[c] + i0, src + c * kTile * 8, kTile * 8);
        // valley / sub-valley / region only change in scattering events: the event
        // path writes them itself, except for sub-tiles processed in place
        if (*reinterpret_cast<volatile unsigned *>(&ctl->dirty[st])) {
          tmaBulkStore(P.packed + i0, src + EMCGPU_N_STREAMS * kTile * 8, kTile * 4);
          *reinterpret_cast<volatile unsigned *>(&ctl->dirty[st]) = 0;
        }
        tmaCommitGroup();
        // stores stay in flight: a stage is handed back to the loader once the
        // store issued kFreeLag tiles ago has finished reading shared memory
        tmaWaitGroupRead<kFreeLag>();
        if (t >= kFreeLag) mbarArrive(&ctl->freed[(t - kFreeLag) % stages]);
        tmaWaitGroup<kStoreLag>(); // all but the kStoreLag newest stores are complete in global memory
        if (t + 1 > kStoreLag) {
          __threadfence_block();
          *reinterpret_cast<volatile unsigned *>(&ctl->storesDone) = (unsigned)(t + 1 - kStoreLag);
        }
      }
      tmaWaitGroup<0>();
      __threadfence_block();
      *reinterpret_cast<volatile unsigned *>(&ctl->storesDone) = (unsigned)myTiles + 1u;
    }
  } else {
    // --------------------------- consumer warps ---------------------------
    const unsigned ltMask = (1u << lane) - 1u;
    for (;;) {
      unsigned q = 0;
      if (lane == 0) q = atomicAdd(&ctl->nextSub, 1u);
      q = __shfl_sync(0xffffffffu, q, 0);
      const int t = (int)(q / kSubsPerTile), sub = (int)(q % kSubsPerTile);
      if (t >= myTiles) break;
      const int st = t % stages;
      while (*reinterpret_cast<volatile unsigned *>(&ctl->loadsIssued) <= (unsigned)t) {
      }
      mbarWait(&ctl->full[st], (t / stages) & 1);
      const uint32_t tileBase = smemAddr(ring + (size_t)st * kTileBytes);
      const int j = sub * kSubTile + 2 * lane;
      const int64_t i0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * kTile + j;
      double s[EMCGPU_N_STREAMS][2];
      uint32_t w[2];
#pragma unroll
      for (int c = 0; c < EMCGPU_N_STREAMS; c++) {
        if (c == EMCGPU_ENERGY) {
          s[c][0] = s[c][1] = 0.0;
        } else {
          asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];"
                       : "=d"(s[c][0]), "=d"(s[c][1])
                       : "r"(tileBase + c * kTile * 8 + j * 8));
        }
      }
      asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];"
                   : "=r"(w[0]), "=r"(w[1])
                   : "r"(tileBase + EMCGPU_N_STREAMS * kTile * 8 + j * 4));
      bool ev[2];
      double eOut[2], vOut[2];
#pragma unroll
      for (int jj = 0; jj < 2; jj++) {
        ev[jj] = false;
        eOut[jj] = 0.0;
        vOut[jj] = 0.0;
        if (s[EMCGPU_TAU][jj] >= dt) {
          const int valley = w[jj] & 0xffu, sv = (w[jj] >> 8) & 0xffu;
          if constexpr (EXACT) {
            Particle p;
            p.k = Vec3{s[EMCGPU_KX][jj], s[EMCGPU_KY][jj], s[EMCGPU_KZ][jj]};
            p.energy = s[EMCGPU_ENERGY][jj];
            p.tau = s[EMCGPU_TAU][jj];
            p.pos = Vec3{s[EMCGPU_X][jj], s[EMCGPU_Y][jj], s[EMCGPU_Z][jj]};
            p.valley = valley;
            p.sub = sv;
            const DevValley &v = model.valleys[valley];
            drift<true, 3>(v, p, dt, P.force);
            s[EMCGPU_KX][jj] = p.k.x;
            s[EMCGPU_KY][jj] = p.k.y;
            s[EMCGPU_KZ][jj] = p.k.z;
            s[EMCGPU_ENERGY][jj] = p.energy;
            s[EMCGPU_TAU][jj] = __dsub_rn(p.tau, dt);
            s[EMCGPU_X][jj] = wrap1<true>(p.pos.x, P.box.x);
            s[EMCGPU_Y][jj] = wrap1<true>(p.pos.y, P.box.y);
            s[EMCGPU_Z][jj] = wrap1<true>(p.pos.z, P.box.z);
            vOut[jj] = driftVelocity<true>(v, sv, p.k, p.energy, P.dir);
          } else {
            vOut[jj] = fastStep(C.fast[valley * EMCGPU_MAX_SUBVALLEYS + sv], C.fastV[valley], dt, P.box,
                                s[EMCGPU_KX][jj], s[EMCGPU_KY][jj], s[EMCGPU_KZ][jj], s[EMCGPU_ENERGY][jj],
                                s[EMCGPU_TAU][jj], s[EMCGPU_X][jj], s[EMCGPU_Y][jj], s[EMCGPU_Z][jj]);
          }
          eOut[jj] = s[EMCGPU_ENERGY][jj];
          if (single) {
            accE += eOut[jj];
            accV += vOut[jj];
            accN++;
          }
        } else {
          ev[jj] = true;
        }
      }
      const unsigned m0 = __ballot_sync(0xffffffffu, ev[0]), m1 = __ballot_sync(0xffffffffu, ev[1]);
      const int nEv = __popc(m0) + __popc(m1);
      const bool dense = nEv >= kDenseEvents;
      if (dense) {
        // dt >~ tau regime: most lanes scatter, process them right here
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          bool did = false;
          int valley = 0;
          if (ev[jj]) {
            Particle p;
            Rng rng;
            p.k = Vec3{s[EMCGPU_KX][jj], s[EMCGPU_KY][jj], s[EMCGPU_KZ][jj]};
            p.energy = s[EMCGPU_ENERGY][jj];
            p.tau = s[EMCGPU_TAU][jj];
            p.pos = Vec3{s[EMCGPU_X][jj], s[EMCGPU_Y][jj], s[EMCGPU_Z][jj]};
            p.valley = w[jj] & 0xffu;
            p.sub = (w[jj] >> 8) & 0xffu;
            p.region = w[jj] >> 16;
            vOut[jj] = bulkParticleStepNI<EXACT, RNG_MODE>(C, P, p, rng, i0 + jj);
            s[EMCGPU_KX][jj] = p.k.x;
            s[EMCGPU_KY][jj] = p.k.y;
            s[EMCGPU_KZ][jj] = p.k.z;
            s[EMCGPU_ENERGY][jj] = p.energy;
            s[EMCGPU_TAU][jj] = p.tau;
            s[EMCGPU_X][jj] = p.pos.x;
            s[EMCGPU_Y][jj] = p.pos.y;
            s[EMCGPU_Z][jj] = p.pos.z;
            w[jj] = (uint32_t)p.valley | ((uint32_t)p.sub << 8) | ((uint32_t)p.region << 16);
            eOut[jj] = p.energy;
            valley = p.valley;
            did = true;
          }
          __syncwarp();
          accumulateObsWarp(C.obs, nV, did, valley, eOut[jj], vOut[jj]);
          ev[jj] = false;
        }
      }
      if (!single) {
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          const bool fastDone = !dense ? !ev[jj] : !((jj ? m1 : m0) >> lane & 1u);
          accumulateObsWarp(C.obs, nV, fastDone, (int)(w[jj] & 0xffu), eOut[jj], vOut[jj]);
        }
      }
#pragma unroll
      for (int c = 0; c < EMCGPU_N_STREAMS; c++)
        asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(tileBase + c * kTile * 8 + j * 8), "d"(s[c][0]),
                     "d"(s[c][1])
                     : "memory");
      if (dense) {
        asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(tileBase + EMCGPU_N_STREAMS * kTile * 8 + j * 4),
                     "r"(w[0]), "r"(w[1])
                     : "memory");
        *reinterpret_cast<volatile unsigned *>(&ctl->dirty[st]) = 1u;
      }
      fenceProxyAsyncShared();
      __syncwarp();
      if (lane == 0) mbarArrive(&ctl->done[st]);
      if (!dense && nEv > 0) {
        // copy the scattering particles into the CTA-wide event queue
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(&ctl->qTail, (unsigned)nEv);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
          if (ev[jj]) {
            const unsigned pos = base + (jj ? __popc(m0) : 0) + __popc((jj ? m1 : m0) & ltMask);
            const unsigned slot = pos % kQueueCap;
            const uint32_t epoch = pos / kQueueCap + 1;
            while (*reinterpret_cast<volatile uint32_t *>(&Q->flag[slot]) != 2 * (epoch - 1)) { // previous tenant read
            }
#pragma unroll
            for (int c = 0; c < EMCGPU_N_STREAMS; c++) Q->f[c][slot] = s[c][jj];
            Q->w[slot] = w[jj];
            Q->idx[slot] = (uint32_t)(i0 + jj);
            Q->tile[slot] = (uint32_t)t;
            __threadfence_block();
            *reinterpret_cast<volatile uint32_t *>(&Q->flag[slot]) = 2 * epoch - 1;
          }
        }
        __syncwarp();
      }
      // drain full batches
      for (;;) {
        unsigned h = 0;
        int got = 0;
        if (lane == 0) {
          h = *reinterpret_cast<volatile unsigned *>(&ctl->qHead);
          const unsigned tl = *reinterpret_cast<volatile unsigned *>(&ctl->qTail);
          if (tl - h >= 32u) got = atomicCAS(&ctl->qHead, h, h + 32u) == h ? 1 : 2;
        }
        got = __shfl_sync(0xffffffffu, got, 0);
        if (got == 0) break;
        if (got == 2) continue; // lost the race, look again
        h = __shfl_sync(0xffffffffu, h, 0);
        processEventBatch<EXACT, RNG_MODE>(C, P, ctl, Q, h, 32);
      }
    }
    // all pushes are complete once every consumer warp is here
    asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumerWarps * 32) : "memory");
    if (warp == 0) {
      unsigned h = *reinterpret_cast<volatile unsigned *>(&ctl->qHead);
      const unsigned tl = *reinterpret_cast<volatile unsigned *>(&ctl->qTail);
      while (h != tl) {
        const int take = min(32u, tl - h);
        processEventBatch<EXACT, RNG_MODE>(C, P, ctl, Q, h, take);
        h += take;
      }
    }
    // the n % kTile particles behind the last whole tile: plain global path
    if (blockIdx.x == 0) {
      const int64_t first = nTiles * kTile;
      const int rounds = (int)((P.n - first + kTmaConsumerWarps * 32 - 1) / (kTmaConsumerWarps * 32));
      for (int r = 0; r < rounds; r++) {
        const int64_t i = first + (int64_t)r * kTmaConsumerWarps * 32 + tid;
        const bool live = i < P.n;
        Particle p;
        Rng rng;
        double e = 0.0, vd = 0.0;
        p.valley = 0;
        if (live) {
          loadParticle(P, i, p, rng);
          vd = bulkParticleStepNI<EXACT, RNG_MODE>(C, P, p, rng, i);
          e = p.energy;
          storeParticleState(P, i, p);
        }
        __syncwarp();
        accumulateObsWarp(C.obs, nV, live, p.valley, e, vd);
      }
    }
    if (single) {
      const double se = warpSum(accE), sv = warpSum(accV);
      const unsigned cnt = __reduce_add_sync(0xffffffffu, accN);
      if (lane == 0 && cnt) {
        atomicAdd(C.obs + 0, se);
        atomicAdd(C.obs + 1, sv);
        atomicAdd(C.obs + 2, (double)cnt);
      }
    }
  }
  __syncthreads();
  for (int j = tid; j < nV * 3; j += blockDim.x) {
    const double v = C.obs[j];
    if (v != 0.0) atomicAdd(P.obs + j, v);
  }
}

// ---------------------------------------------------------------------------
// K1b: nSteps consecutive time steps per launch, particle state held in
// registers between them (136 B of HBM traffic per particle per LAUNCH).
template <bool EXACT, int RNG_MODE>
__global__ void __launch_bounds__(kBulkThreads, 2) bulkStepKernel(const __grid_constant__ BulkParams P) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint64_t tableBar;
  const int nV = P.model->nValleys;
  const int obsPerStep = nV * 3;
  const CtaState C = stageCta(P, smemRaw, &tableBar, P.nSteps * obsPerStep, 0);
  double *sObs = C.obs;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // round the trip count up so that whole warps stay converged for the shuffles
  const int64_t nRounded = (P.n + 31) & ~int64_t(31);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < nRounded; i += stride) {
    const bool live = i < P.n;
    Particle p;
    Rng rng;
    double grain = 0.0;
    if (live) {
      loadParticle(P, i, p, rng);
      attachReplay<RNG_MODE>(P, i, rng);
      if (P.grainTau) grain = P.grainTau[i];
    } else {
      p.valley = 0;
      p.sub = 0;
      p.region = 0;
    }
    for (int s = 0; s < P.nSteps; s++) {
      double e = 0.0, vd = 0.0;
      if (live) {
        rng.step = (uint32_t)(P.step0 + s);
        rng.n = 0;
        vd = bulkParticleStep<EXACT, RNG_MODE>(C, P, p, rng, P.idBase + i, P.step0 + s);
        if (P.grainTau) { // basicBulkParticleHandler.hpp:216-220
          grain = __dsub_rn(grain, P.dt);
          if (grain <= 0.0) {
            grain = grainEvent<RNG_MODE>(P, p, rng);
            vd = driftVelocity<EXACT>(C.model->valleys[p.valley], p.sub, p.k, p.energy, P.dir);
          }
        }
        e = p.energy;
        if (P.velOut) {
          if (P.velComponents == 1) {
            P.velOut[(size_t)s * P.n + i] = vd;
          } else {
            const Vec3 vel = velocityVector(C.model->valleys[p.valley], p.sub, p.k, p.energy);
            double *o = P.velOut + ((size_t)s * P.n + i) * 3;
            o[0] = vel.x;
            o[1] = vel.y;
            o[2] = vel.z;
          }
        }
      }
      // per-valley block partial sums (basicBulkParticleHandler.hpp:289-347)
      double *o = sObs + s * obsPerStep;
      if (nV == 1) {
        const double se = warpSum(e), sv = warpSum(vd);
        const unsigned cnt = __popc(__ballot_sync(0xffffffffu, live));
        if (lane == 0) {
          atomicAdd(o + 0, se);
          atomicAdd(o + 1, sv);
          atomicAdd(o + 2, (double)cnt);
        }
      } else {
        accumulateObsWarp(o, nV, live, p.valley, e, vd);
      }
    }
    if (live) {
      storeParticle<RNG_MODE>(P, i, p, rng);
      if (P.grainTau) P.grainTau[i] = grain;
    }
  }
  __syncthreads();
  for (int j = tid; j < P.nSteps * obsPerStep; j += blockDim.x) {
    const double v = sObs[j];
    if (v != 0.0) atomicAdd(P.obs + j, v);
  }
}

// Observables of the current ensemble without moving it
// (basicBulkParticleHandler.hpp:289-347).
template <bool EXACT>
__global__ void __launch_bounds__(kBulkThreads) bulkObservablesKernel(const BulkParams P) {
  __shared__ double sObs[EMCGPU_MAX_VALLEYS * 3];
  const DevModel &model = *P.model;
  const int nV = model.nValleys;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < nV * 3) sObs[tid] = 0.0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nRounded = (P.n + 31) & ~int64_t(31);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < nRounded; i += stride) {
    const bool live = i < P.n;
    double e = 0.0, vd = 0.0;
    int valley = 0;
    if (live) {
      const uint32_t w = P.packed[i];
      valley = w & 0xffu;
      const int sub = (w >> 8) & 0xffu;
      const Vec3 k{P.stream[EMCGPU_KX][i], P.stream[EMCGPU_KY][i], P.stream[EMCGPU_KZ][i]};
      e = P.stream[EMCGPU_ENERGY][i];
      vd = driftVelocity<EXACT>(model.valleys[valley], sub, k, e, P.dir);
    }
    for (int v = 0; v < nV; v++) {
      const bool mine = live && valley == v;
      const unsigned cnt = __popc(__ballot_sync(0xffffffffu, mine));
      if (cnt == 0) continue;
      const double se = warpSum(mine ? e : 0.0), sv = warpSum(mine ? vd : 0.0);
      if (lane == 0) {
        atomicAdd(sObs + 3 * v + 0, se);
        atomicAdd(sObs + 3 * v + 1, sv);
        atomicAdd(sObs + 3 * v + 2, (double)cnt);
      }
    }
  }
  __syncthreads();
  if (tid < nV * 3 && sObs[tid] != 0.0) atomicAdd(P.obs + tid, sObs[tid]);
}

// Device-side thermal ensemble (emcElectron::generateInitialParticle,
// include/ParticleType/emcElectron.hpp:75-90; emcParticleInitialization.hpp:36-51;
// positions uniform in the periodic box).  Draw order per particle:
// x, y, z, valley, sub-valley, energy, cos(theta), phi, tau.  Philox step word = 0.
struct GenParams {
  double *stream[EMCGPU_N_STREAMS];
  uint32_t *packed;
  int64_t n, idBase;
  const DevModel *model;
  Vec3 box;
  double thermalVoltage;
  int32_t region;
  uint64_t seed;
};

static __global__ void __launch_bounds__(256) bulkGenerateKernel(const GenParams G) {
  const DevModel &model = *G.model;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < G.n; i += stride) {
    const uint64_t id = (uint64_t)(G.idBase + i);
    Rng rng;
    rng.k0 = (uint32_t)G.seed;
    rng.k1 = (uint32_t)(G.seed >> 32);
    rng.idLo = (uint32_t)id;
    rng.idHi = (uint32_t)(id >> 32);
    rng.step = 0;
    rng.n = 0;
    const double x = uniform01(rng.raw<RNG_PHILOX>()) * G.box.x;
    const double y = uniform01(rng.raw<RNG_PHILOX>()) * G.box.y;
    const double z = uniform01(rng.raw<RNG_PHILOX>()) * G.box.z;
    const int valley = (int)floor(model.nValleys * uniformLog(rng.raw<RNG_PHILOX>()));
    const DevValley &v = model.valleys[valley];
    const int sub = (int)floor(v.deg * uniformLog(rng.raw<RNG_PHILOX>()));
    const double e = -1.5 * G.thermalVoltage * log(uniformLog(rng.raw<RNG_PHILOX>()));
    const double r2 = uniform01(rng.raw<RNG_PHILOX>());
    const double r1 = uniform01(rng.raw<RNG_PHILOX>());
    const Vec3 k = randomDirection<true>(normWaveVec<true>(v, e), r1, r2);
    const int set = (G.region < kMaxRegions) ? model.setOf[valley][G.region] : -1;
    const double tau0 = set >= 0 ? model.sets[set].tau : model.defaultTau;
    const double tau = -log(uniformLog(rng.raw<RNG_PHILOX>())) * tau0;
    G.stream[EMCGPU_KX][i] = k.x;
    G.stream[EMCGPU_KY][i] = k.y;
    G.stream[EMCGPU_KZ][i] = k.z;
    G.stream[EMCGPU_ENERGY][i] = e;
    G.stream[EMCGPU_TAU][i] = tau;
    G.stream[EMCGPU_X][i] = x;
    G.stream[EMCGPU_Y][i] = y;
    G.stream[EMCGPU_Z][i] = z;
    G.packed[i] = (uint32_t)valley | ((uint32_t)sub << 8) | ((uint32_t)G.region << 16);
  }
}

} // namespace emc
