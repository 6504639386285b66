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

// One full time step of one particle (A.5 of SURVEY.md).
template <bool EXACT, int RNG_MODE>
__device__ __forceinline__ void bulkParticleStep(const DevModel &model, const double *tables,
                                                 const DevMech *mechs, const BulkParams &P, Particle &p,
                                                 Rng &rng, int64_t particleId, int64_t step) {
  using A = Arith<EXACT>;
  const double dt = P.dt;
  {
    const DevValley &v = model.valleys[p.valley];
    drift<EXACT, 3>(v, p, fmin(p.tau, dt), P.force);
    p.pos.x = wrap1<EXACT>(p.pos.x, P.box.x);
    p.pos.y = wrap1<EXACT>(p.pos.y, P.box.y);
    p.pos.z = wrap1<EXACT>(p.pos.z, P.box.z);
  }
  double tRem = A::sub(dt, p.tau);
  while (tRem > 0.0) {
    const int set = (p.region < kMaxRegions) ? model.setOf[p.valley][p.region] : -1;
    double tauTab = model.defaultTau;
    if (set >= 0) {
      const DevTableSet &ts = model.sets[set];
      const int lvl = energyLevel(p.energy, model.dE, model.nLevels);
      const double r = uniform01(rng.raw<RNG_MODE>());
      const double *row = tables + ts.tabOffset + (int64_t)lvl * ts.stride;
      const int m = selectMechanism(row, ts.nMech, r);
      int mechId = -1;
      if (m >= 0) {
        const DevMech &mech = mechs[ts.mechOffset + m];
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
  p.tau = A::sub(p.tau, dt);
}

template <bool EXACT, int RNG_MODE>
__global__ void __launch_bounds__(kBulkThreads, 2) bulkStepKernel(const BulkParams P) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint64_t tableBar;
  // layout: DevModel | obs accumulators | mechs | tables
  DevModel *sModel = reinterpret_cast<DevModel *>(smemRaw);
  size_t off = (sizeof(DevModel) + 15) & ~size_t(15);
  double *sObs = reinterpret_cast<double *>(smemRaw + off);
  const int nV = P.model->nValleys;
  const int obsPerStep = nV * 3;
  off += (size_t)P.nSteps * obsPerStep * sizeof(double);
  off = (off + 15) & ~size_t(15);
  DevMech *sMechs = reinterpret_cast<DevMech *>(smemRaw + off);
  off += (size_t)P.nMechTotal * sizeof(DevMech);
  off = (off + 15) & ~size_t(15);
  double *sTables = reinterpret_cast<double *>(smemRaw + off);

  const int tid = threadIdx.x;
  const uint32_t tableBytes = (uint32_t)(P.model->tableDoubles * sizeof(double));
  if (P.tablesInSmem) {
    if (tid == 0) {
      mbarInit(&tableBar, 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
      mbarExpectTx(&tableBar, tableBytes);
      // chunks of <= 32 KB, each a multiple of 16 B
      const uint32_t chunk = 32768u;
      for (uint32_t o = 0; o < tableBytes; o += chunk) {
        const uint32_t b = min(chunk, tableBytes - o);
        tmaBulkLoad(reinterpret_cast<unsigned char *>(sTables) + o,
                    reinterpret_cast<const unsigned char *>(P.tables) + o, b, &tableBar);
      }
    }
  }
  // model + mechanism descriptors: plain cooperative copy (a few KB)
  {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(P.model);
    uint32_t *dst = reinterpret_cast<uint32_t *>(sModel);
    for (int i = tid; i < (int)(sizeof(DevModel) / 4); i += blockDim.x) dst[i] = src[i];
    const uint32_t *msrc = reinterpret_cast<const uint32_t *>(P.mechs);
    uint32_t *mdst = reinterpret_cast<uint32_t *>(sMechs);
    for (int i = tid; i < (int)(P.nMechTotal * sizeof(DevMech) / 4); i += blockDim.x) mdst[i] = msrc[i];
    for (int i = tid; i < P.nSteps * obsPerStep; i += blockDim.x) sObs[i] = 0.0;
  }
  __syncthreads();
  if (P.tablesInSmem) mbarWait(&tableBar, 0);
  const DevModel &model = *sModel;
  const double *tables = P.tablesInSmem ? sTables : P.tables;

  const int lane = tid & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // round the trip count up so that whole warps stay converged for the shuffles
  const int64_t nRounded = (P.n + 31) & ~int64_t(31);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + tid; i < nRounded; i += stride) {
    const bool live = i < P.n;
    Particle p;
    Rng rng;
    if (live) {
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
      if constexpr (RNG_MODE == RNG_REPLAY) {
        rng.stream = P.draws + P.offsets[i] + P.cursor[i];
        rng.streamEnd = P.draws + P.offsets[i + 1];
      }
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
        bulkParticleStep<EXACT, RNG_MODE>(model, tables, sMechs, P, p, rng, P.idBase + i, P.step0 + s);
        e = p.energy;
        vd = driftVelocity<EXACT>(model.valleys[p.valley], p.sub, p.k, p.energy, P.dir);
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
        for (int v = 0; v < nV; v++) {
          const bool mine = live && p.valley == v;
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
    }
    if (live) {
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
