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
};

constexpr int kBulkThreads = 256;
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
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra DONE;\n"
               "bra WAIT_LOOP;\n"
               "DONE:\n"
               "}" ::"r"(smemAddr(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void tmaBulkLoad(void *dstSmem, const void *srcGlobal, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemAddr(dstSmem)),
               "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
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
    if (s == 0) {
      FastValley fv;
      fv.fE = v.nonParabolic ? v.fE : 2.0 * v.fE;
      fv.c2a = 2.0 * v.alpha * fv.fE;
      fv.diag = v.rotKind != ROT_GENERAL;
      fv.pad = 0;
      sFastV[i / EMCGPU_MAX_SUBVALLEYS] = fv;
    }
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

// The scattering part of a time step (A.5 of SURVEY.md): entered with the
// particle already drifted to its first scattering time; tRem = dt - tau_old.
template <bool EXACT, int RNG_MODE>
__device__ __forceinline__ void scatterLoop(const CtaState &C, const BulkParams &P, Particle &p, Rng &rng,
                                            int64_t particleId, int64_t step, double tRem) {
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
        sampleFinalState<EXACT, RNG_MODE>(model, mech, p, rng);
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
    const DevValley &v = model.valleys[p.valley];
    drift<EXACT, 3>(v, p, fmin(tRem, newTau), P.force);
    p.pos.x = wrap1<EXACT>(p.pos.x, P.box.x);
    p.pos.y = wrap1<EXACT>(p.pos.y, P.box.y);
    p.pos.z = wrap1<EXACT>(p.pos.z, P.box.z);
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
  {
    const DevValley &v = C.model->valleys[p.valley];
    drift<EXACT, 3>(v, p, fmin(p.tau, dt), P.force);
    p.pos.x = wrap1<EXACT>(p.pos.x, P.box.x);
    p.pos.y = wrap1<EXACT>(p.pos.y, P.box.y);
    p.pos.z = wrap1<EXACT>(p.pos.z, P.box.z);
  }
  scatterLoop<EXACT, RNG_MODE>(C, P, p, rng, particleId, step, A::sub(dt, p.tau));
  p.tau = A::sub(p.tau, dt);
  return driftVelocity<EXACT>(C.model->valleys[p.valley], p.sub, p.k, p.energy, P.dir);
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
  p.energy = P.stream[EMCGPU_ENERGY][i];
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
__global__ void __launch_bounds__(kBulkThreads, 2) bulkStreamKernel(const __grid_constant__ BulkParams P) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint64_t tableBar;
  constexpr int QW = 32 + 32 * VEC;
  const int nV = P.model->nValleys;
  const CtaState C = stageCta(P, smemRaw, &tableBar, nV * 3, (kBulkThreads / 32) * QW);
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
      for (int c = 0; c < EMCGPU_N_STREAMS; c++) VecIO<VEC>::ld(P.stream[c] + i0, s[c]);
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
    if (live) {
      loadParticle(P, i, p, rng);
      attachReplay<RNG_MODE>(P, i, rng);
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
        e = p.energy;
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
    if (live) storeParticle<RNG_MODE>(P, i, p, rng);
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

__global__ void __launch_bounds__(256) bulkGenerateKernel(const GenParams G) {
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
