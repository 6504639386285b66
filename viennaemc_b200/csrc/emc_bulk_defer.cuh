// K1c: several consecutive time steps per launch with DEFERRED scattering events.
//
// Same per-particle arithmetic as K1a / K1b (emc_bulk_kernel.cuh; reference:
// examples/bulkSimulation/basicBulkParticleHandler.hpp:181-225 moveParticles,
// :289-347 observables) -- the trajectories are bit-identical -- but organised so that
// the particle state crosses HBM once per LAUNCH (136 B per particle per nSteps time
// steps) and no warp ever walks the long scattering path with a few lanes:
//
//   main pass     every lane keeps 2 particles in registers and advances them by
//                 full-dt free flights (>= 98 % of all particle-steps) until the step in
//                 which the particle's flight ends (tau < dt).  There it FREEZES: state
//                 and step index stay as they are.  After nSteps the finished particles
//                 go back to global memory with vector stores, the frozen ones -- state,
//                 index and step -- into a CTA-wide shared-memory queue.
//   event batches whenever 32 particles are queued a warp claims them: every lane runs
//                 the complete scattering step of ITS particle at ITS step (drift to the
//                 event, table selection, final state, new flight time), continues with
//                 full-dt flights and either finishes the launch's last step (store) or
//                 meets the next event and queues the particle again.  The random
//                 numbers are Philox(key, id, step, draw), so the order in which events
//                 are served changes nothing.
//   observables   main pass: per-thread accumulators in shared memory, one slot per step
//                 (no atomics, one shuffle per step); event batches: shared-memory
//                 atomics on the per-step sums (lanes are at different steps).
//
// The kernel is bound by the FP64 pipe, not by HBM: ~50 FP64 instructions per
// particle-step at 64 lanes / clk / SM.
#pragma once
#include "emc_bulk_kernel.cuh"

namespace emc {

#ifndef EMC_DEFER_WARPS
#define EMC_DEFER_WARPS 16
#endif
constexpr int kDeferWarps = EMC_DEFER_WARPS;
constexpr int kDeferThreads = kDeferWarps * 32;
constexpr int kDeferChunk = 64;    // particles per warp and pass (2 per lane)
constexpr int kDeferMaxSteps = 8;  // time steps per launch (measured optimum; more steps cost L1 through the observable slots)
constexpr int kDeferDense = 16;    // frozen particles per chunk from which the chunk is finished in place
constexpr int kDeferQueueCap = 32 + kDeferWarps * 32 + 64; // > 31 + kDeferWarps * 32 (see the capacity argument at pushFrozen)
constexpr int kDeferStreams = 7;   // kx ky kz tau x y z (the energy is recomputed by the first drift)

struct DeferControl {
  uint64_t tableBar;
  unsigned qTail, qHead;
  unsigned nextChunk; // chunks of this CTA handed out so far (dynamic claiming)
};
struct DeferQueue {
  double f[kDeferStreams][kDeferQueueCap];
  uint32_t w[kDeferQueueCap];
  uint32_t idx[kDeferQueueCap];
  uint32_t step[kDeferQueueCap];
  uint32_t flag[kDeferQueueCap];
};
constexpr int kDeferQueueWords = (int)((sizeof(DeferControl) + sizeof(DeferQueue) + 3) / 4);

__host__ __device__ inline size_t deferObsOffset(const BulkSmem &L) { return (L.total + 15) & ~size_t(15); }
// per-thread observable slots: [nSteps][2 (sum E, sum v.E)][kDeferThreads]
__host__ __device__ inline size_t deferSmemBytes(const BulkSmem &L, int nSteps, int /*nValleys*/) {
  return deferObsOffset(L) + (size_t)nSteps * 2 * kDeferThreads * sizeof(double);
}

// A whole time step of a particle whose flight outlasts it (tau >= dt): drift(dt), periodic wrap, tau -= dt
// (basicBulkParticleHandler.hpp:195-213); returns v.Ê (:326-347).
template <bool EXACT>
__device__ __forceinline__ double fullDtStep(const CtaState &C, const BulkParams &P, Particle &p) {
  if constexpr (EXACT) {
    const DevValley &v = C.model->valleys[p.valley];
    drift<true, 3>(v, p, P.dt, P.force);
    p.tau = __dsub_rn(p.tau, P.dt);
    p.pos.x = wrap1<true>(p.pos.x, P.box.x);
    p.pos.y = wrap1<true>(p.pos.y, P.box.y);
    p.pos.z = wrap1<true>(p.pos.z, P.box.z);
    return driftVelocity<true>(v, p.sub, p.k, p.energy, P.dir);
  } else {
    return fastStep(C.fast[p.valley * EMCGPU_MAX_SUBVALLEYS + p.sub], C.fastV[p.valley], P.dt, P.box, p.k.x, p.k.y,
                    p.k.z, p.energy, p.tau, p.pos.x, p.pos.y, p.pos.z);
  }
}

// Queue the lanes with `has` (warp-collective).  Capacity: a warp pushes at most 32 entries between two visits of
// the drain loop, and leaves the drain loop only when fewer than 32 entries are queued, so the queue never holds more
// than 31 + kDeferWarps * 32 entries: a pusher never waits for a slot that only itself could free.
__device__ __forceinline__ void pushFrozen(DeferControl *ctl, DeferQueue *Q, bool has, const Particle &p, uint32_t idx,
                                           int step) {
  const unsigned mask = __ballot_sync(0xffffffffu, has);
  if (!mask) return;
  const int lane = threadIdx.x & 31;
  unsigned base = 0;
  if (lane == 0) base = atomicAdd(&ctl->qTail, (unsigned)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (has) {
    const unsigned pos = base + __popc(mask & ((1u << lane) - 1u));
    const unsigned slot = pos % kDeferQueueCap;
    const uint32_t epoch = pos / kDeferQueueCap + 1;
    while (*reinterpret_cast<volatile uint32_t *>(&Q->flag[slot]) != 2 * (epoch - 1)) { // previous tenant read
    }
    Q->f[0][slot] = p.k.x;
    Q->f[1][slot] = p.k.y;
    Q->f[2][slot] = p.k.z;
    Q->f[3][slot] = p.tau;
    Q->f[4][slot] = p.pos.x;
    Q->f[5][slot] = p.pos.y;
    Q->f[6][slot] = p.pos.z;
    Q->w[slot] = (uint32_t)p.valley | ((uint32_t)p.sub << 8) | ((uint32_t)p.region << 16);
    Q->idx[slot] = idx;
    Q->step[slot] = (uint32_t)step;
    __threadfence_block();
    *reinterpret_cast<volatile uint32_t *>(&Q->flag[slot]) = 2 * epoch - 1;
  }
  __syncwarp();
}

// One particle-step of an event lane into the per-step sums: one valley -> the thread's own slots (no atomics);
// several valleys -> atomics on the warp's copy of the per-valley sums.
__device__ __forceinline__ void addObsLane(double *wObs, double *myObs, int obsPerStep, bool single, int s, int valley,
                                           double e, double vd) {
  if (single) {
    myObs[(2 * s) * kDeferThreads] += e;
    myObs[(2 * s + 1) * kDeferThreads] += vd;
  } else {
    double *o = wObs + s * obsPerStep + 3 * valley;
    atomicAdd(o + 0, e);
    atomicAdd(o + 1, vd);
    atomicAdd(o + 2, 1.0);
  }
}

// Advance the lane's particle from step s (relative to P.step0) to the end of the launch or, with `repush`, to its
// next event, where it is queued again.  Warp-collective; lanes may be at different steps.
template <bool EXACT, int RNG_MODE>
__device__ __forceinline__ void runEvents(const CtaState &C, const BulkParams &P, DeferControl *ctl, DeferQueue *Q,
                                          double *obsT, Particle &p, uint32_t idx, int s, bool active, bool repush) {
  const int nV = C.model->nValleys;
  const int obsPerStep = nV * 3;
  const bool single = nV == 1;
  const int nSteps = P.nSteps;
  const double dt = P.dt;
  double *const wObs = C.obs + (threadIdx.x >> 5) * nSteps * obsPerStep; // the warp's own copy of the per-step sums
  double *const myObs = obsT + threadIdx.x;
  for (;;) {
    if (active) {
      Rng rng;
      const uint64_t id = (uint64_t)(P.idBase + idx);
      rng.k0 = (uint32_t)P.seed;
      rng.k1 = (uint32_t)(P.seed >> 32);
      rng.idLo = (uint32_t)id;
      rng.idHi = (uint32_t)(id >> 32);
      rng.status = P.status;
      rng.n = 0;
      rng.step = (uint32_t)(P.step0 + s);
      attachReplay<RNG_MODE>(P, idx, rng);
      double vd = bulkParticleStep<EXACT, RNG_MODE>(C, P, p, rng, P.idBase + idx, P.step0 + s);
      if constexpr (RNG_MODE == RNG_REPLAY) storeCursor(P, idx, rng);
      addObsLane(wObs, myObs, obsPerStep, single, s, p.valley, p.energy, vd);
      s++;
      while (s < nSteps && p.tau >= dt) {
        vd = fullDtStep<EXACT>(C, P, p);
        addObsLane(wObs, myObs, obsPerStep, single, s, p.valley, p.energy, vd);
        s++;
      }
      if (s == nSteps) {
        storeParticleState(P, idx, p);
        active = false;
      }
    }
    __syncwarp();
    if (!__any_sync(0xffffffffu, active)) break;
    if (repush) {
      pushFrozen(ctl, Q, active, p, idx, s);
      break;
    }
  }
}

// The CTA's staged model, rebuilt from the shared-memory symbol itself so that the compiler knows the address space
// (shared-memory loads instead of generic ones in the out-of-line event routines).
__device__ __forceinline__ CtaState localCtaState(const BulkParams &P) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int nV = P.model->nValleys;
  const BulkSmem L(kDeferWarps * P.nSteps * nV * 3, nV, P.nMechTotal, P.model->tableDoubles, P.tablesInSmem != 0,
                   kDeferQueueWords);
  CtaState c;
  c.model = reinterpret_cast<const DevModel *>(smemRaw + L.model);
  c.obs = reinterpret_cast<double *>(smemRaw + L.obs);
  c.mechs = reinterpret_cast<const DevMech *>(smemRaw + L.mechs);
  c.fastV = reinterpret_cast<const FastValley *>(smemRaw + L.fastV);
  c.fast = reinterpret_cast<const FastSub *>(smemRaw + L.fast);
  c.queue = reinterpret_cast<uint32_t *>(smemRaw + L.queue);
  c.tables = P.tablesInSmem ? reinterpret_cast<const double *>(smemRaw + L.tables) : P.tables;
  return c;
}

// Up to 32 queued particles, one per lane.
template <bool EXACT, int RNG_MODE>
__device__ __noinline__ void eventBatch(const CtaState &, const BulkParams &P, DeferControl *ctl, DeferQueue *Q,
                                        double *obsT, unsigned head, int count, bool repush) {
  const CtaState C = localCtaState(P);
  const int lane = threadIdx.x & 31;
  const bool active = lane < count;
  Particle p;
  uint32_t idx = 0;
  int s = 0;
  p.k = Vec3{0.0, 0.0, 0.0};
  p.pos = Vec3{0.0, 0.0, 0.0};
  p.energy = 0.0;
  p.tau = 0.0;
  p.valley = p.sub = p.region = 0;
  if (active) {
    const unsigned pos = head + lane, slot = pos % kDeferQueueCap;
    const uint32_t epoch = pos / kDeferQueueCap + 1;
    while (*reinterpret_cast<const volatile uint32_t *>(&Q->flag[slot]) != 2 * epoch - 1) { // full, this epoch
    }
    __threadfence_block();
    p.k = Vec3{Q->f[0][slot], Q->f[1][slot], Q->f[2][slot]};
    p.tau = Q->f[3][slot];
    p.pos = Vec3{Q->f[4][slot], Q->f[5][slot], Q->f[6][slot]};
    const uint32_t w = Q->w[slot];
    idx = Q->idx[slot];
    s = (int)Q->step[slot];
    __threadfence_block();
    *const_cast<volatile uint32_t *>(&Q->flag[slot]) = 2 * epoch; // consumed: the slot may be refilled
    p.valley = w & 0xffu;
    p.sub = (w >> 8) & 0xffu;
    p.region = w >> 16;
  }
  __syncwarp();
  runEvents<EXACT, RNG_MODE>(C, P, ctl, Q, obsT, p, idx, s, active, repush);
}

// The lane's particle idx, whose state is in global memory, from step s to the end of the launch, in place (rare paths:
// chunks in which most particles scatter, the particles behind the last whole chunk).  Only scalars cross the call.
template <bool EXACT, int RNG_MODE>
__device__ __noinline__ void runFromGlobal(const CtaState &, const BulkParams &P, DeferControl *ctl, DeferQueue *Q,
                                           double *obsT, uint32_t idx, int s, bool active) {
  const CtaState C = localCtaState(P);
  Particle p;
  Rng rng;
  p.k = Vec3{0.0, 0.0, 0.0};
  p.pos = Vec3{0.0, 0.0, 0.0};
  p.energy = p.tau = 0.0;
  p.valley = p.sub = p.region = 0;
  if (active) loadParticle(P, idx, p, rng);
  runEvents<EXACT, RNG_MODE>(C, P, ctl, Q, obsT, p, idx, s, active, false);
}

// serve full batches of the event queue (warp-collective)
template <bool EXACT, int RNG_MODE>
__device__ __forceinline__ void serveBatches(const CtaState &C, const BulkParams &P, DeferControl *ctl, DeferQueue *Q,
                                             double *obsT) {
  const int lane = threadIdx.x & 31;
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
    eventBatch<EXACT, RNG_MODE>(C, P, ctl, Q, obsT, h, 32, true);
  }
}

template <bool EXACT, int RNG_MODE>
__global__ void __launch_bounds__(kDeferThreads, 1) bulkDeferKernel(const __grid_constant__ BulkParams P) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int nV = P.model->nValleys;
  const int nSteps = P.nSteps;
  const int obsPerStep = nV * 3;
  const BulkSmem L(kDeferWarps * nSteps * obsPerStep, nV, P.nMechTotal, P.model->tableDoubles, P.tablesInSmem != 0,
                   kDeferQueueWords);
  DeferControl *ctl = reinterpret_cast<DeferControl *>(smemRaw + L.queue);
  DeferQueue *Q = reinterpret_cast<DeferQueue *>(smemRaw + L.queue + sizeof(DeferControl));
  double *obsT = reinterpret_cast<double *>(smemRaw + deferObsOffset(L)); // [nSteps][2][kDeferThreads]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    ctl->qTail = 0;
    ctl->qHead = 0;
    ctl->nextChunk = 0;
  }
  for (int i = tid; i < kDeferQueueCap; i += kDeferThreads) Q->flag[i] = 0;
  for (int s = 0; s < 2 * nSteps; s++) obsT[s * kDeferThreads + tid] = 0.0;
  const CtaState C = stageCta(P, smemRaw, &ctl->tableBar, kDeferWarps * nSteps * obsPerStep, kDeferQueueWords); // __syncthreads inside
  const bool single = nV == 1;
  const double dt = P.dt;
  const int64_t nChunks = P.n / kDeferChunk;
  // chunks blockIdx.x + k * gridDim.x belong to this CTA; its warps claim them one by one and prefetch the next one
  auto claimChunk = [&]() -> int64_t {
    unsigned k = 0;
    if (lane == 0) k = atomicAdd(&ctl->nextChunk, 1u);
    k = __shfl_sync(0xffffffffu, k, 0);
    return (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
  };
  auto prefetchChunk = [&](int64_t ch) {
    if (ch >= nChunks) return;
    const int64_t i0 = ch * kDeferChunk + 2 * lane;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_KX] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_KY] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_KZ] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_TAU] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_X] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_Y] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_Z] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.packed + i0));
  };

  {
    // ================= general main pass (EXACT arithmetic, several valleys, general rotations) =================
    for (int64_t ch = claimChunk(), chNext; ch < nChunks; ch = chNext) {
      chNext = claimChunk();
      prefetchChunk(chNext);
      const int64_t i0 = ch * kDeferChunk + 2 * lane;
      Particle p[2];
      int fz[2] = {-1, -1};
      {
        const double2 kx = __ldcs(reinterpret_cast<const double2 *>(P.stream[EMCGPU_KX] + i0));
        const double2 ky = __ldcs(reinterpret_cast<const double2 *>(P.stream[EMCGPU_KY] + i0));
        const double2 kz = __ldcs(reinterpret_cast<const double2 *>(P.stream[EMCGPU_KZ] + i0));
        const double2 ta = __ldcs(reinterpret_cast<const double2 *>(P.stream[EMCGPU_TAU] + i0));
        const double2 px = __ldcs(reinterpret_cast<const double2 *>(P.stream[EMCGPU_X] + i0));
        const double2 py = __ldcs(reinterpret_cast<const double2 *>(P.stream[EMCGPU_Y] + i0));
        const double2 pz = __ldcs(reinterpret_cast<const double2 *>(P.stream[EMCGPU_Z] + i0));
        const uint2 w = __ldcs(reinterpret_cast<const uint2 *>(P.packed + i0));
        p[0].k = Vec3{kx.x, ky.x, kz.x};
        p[1].k = Vec3{kx.y, ky.y, kz.y};
        p[0].tau = ta.x;
        p[1].tau = ta.y;
        p[0].pos = Vec3{px.x, py.x, pz.x};
        p[1].pos = Vec3{px.y, py.y, pz.y};
        p[0].energy = p[1].energy = 0.0; // recomputed by the first drift (emcParticleDrift.hpp:25)
        p[0].valley = w.x & 0xffu;
        p[0].sub = (w.x >> 8) & 0xffu;
        p[0].region = w.x >> 16;
        p[1].valley = w.y & 0xffu;
        p[1].sub = (w.y >> 8) & 0xffu;
        p[1].region = w.y >> 16;
      }
      // full-dt flights until the particle's first event
      for (int s = 0; s < nSteps; s++) {
        double e[2] = {0.0, 0.0}, vd[2] = {0.0, 0.0};
        bool moved[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
          moved[j] = false;
          if (fz[j] < 0) {
            if (p[j].tau >= dt) {
              vd[j] = fullDtStep<EXACT>(C, P, p[j]);
              e[j] = p[j].energy;
              moved[j] = true;
            } else {
              fz[j] = s;
            }
          }
        }
        if (single) {
          obsT[(2 * s) * kDeferThreads + tid] += e[0] + e[1];
          obsT[(2 * s + 1) * kDeferThreads + tid] += vd[0] + vd[1];
        } else {
#pragma unroll
          for (int j = 0; j < 2; j++)
            accumulateObsWarp(C.obs + (warp * nSteps + s) * obsPerStep, nV, moved[j], p[j].valley, e[j], vd[j]);
        }
        if (__all_sync(0xffffffffu, fz[0] >= 0 && fz[1] >= 0)) break;
      }
      // finished particles go home, frozen ones into the queue
      if (fz[0] < 0 && fz[1] < 0) {
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_KX] + i0), make_double2(p[0].k.x, p[1].k.x));
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_KY] + i0), make_double2(p[0].k.y, p[1].k.y));
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_KZ] + i0), make_double2(p[0].k.z, p[1].k.z));
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_ENERGY] + i0), make_double2(p[0].energy, p[1].energy));
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_TAU] + i0), make_double2(p[0].tau, p[1].tau));
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_X] + i0), make_double2(p[0].pos.x, p[1].pos.x));
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_Y] + i0), make_double2(p[0].pos.y, p[1].pos.y));
        __stcs(reinterpret_cast<double2 *>(P.stream[EMCGPU_Z] + i0), make_double2(p[0].pos.z, p[1].pos.z));
      } else {
#pragma unroll
        for (int j = 0; j < 2; j++)
          if (fz[j] < 0) storeParticleState(P, i0 + j, p[j]);
      }
      const unsigned f0 = __ballot_sync(0xffffffffu, fz[0] >= 0), f1 = __ballot_sync(0xffffffffu, fz[1] >= 0);
      const int nFz = __popc(f0) + __popc(f1);
      if (nFz >= kDeferDense) {
        // dt >~ tau regime: most particles scatter in every step, finish the chunk right here
        if (fz[0] >= 0) storeParticleState(P, i0, p[0]);
        if (fz[1] >= 0) storeParticleState(P, i0 + 1, p[1]);
        __syncwarp();
        runFromGlobal<EXACT, RNG_MODE>(C, P, ctl, Q, obsT, (uint32_t)i0, fz[0] >= 0 ? fz[0] : 0, fz[0] >= 0);
        runFromGlobal<EXACT, RNG_MODE>(C, P, ctl, Q, obsT, (uint32_t)(i0 + 1), fz[1] >= 0 ? fz[1] : 0, fz[1] >= 0);
      } else if (nFz > 0) {
        pushFrozen(ctl, Q, fz[0] >= 0, p[0], (uint32_t)i0, fz[0]);
        pushFrozen(ctl, Q, fz[1] >= 0, p[1], (uint32_t)(i0 + 1), fz[1]);
      }
      serveBatches<EXACT, RNG_MODE>(C, P, ctl, Q, obsT);
    }
  }
  // every push of the main phase is complete once all warps are here; what is left is finished in place
  __syncthreads();
  for (;;) {
    unsigned h = 0;
    int take = 0;
    if (lane == 0) {
      for (;;) {
        h = *reinterpret_cast<volatile unsigned *>(&ctl->qHead);
        const unsigned tl = *reinterpret_cast<volatile unsigned *>(&ctl->qTail);
        take = (int)min(32u, tl - h);
        if (take == 0 || atomicCAS(&ctl->qHead, h, h + (unsigned)take) == h) break;
      }
    }
    take = __shfl_sync(0xffffffffu, take, 0);
    if (take == 0) break;
    h = __shfl_sync(0xffffffffu, h, 0);
    eventBatch<EXACT, RNG_MODE>(C, P, ctl, Q, obsT, h, take, false);
  }
  // the n % kDeferChunk particles behind the last whole chunk
  if (blockIdx.x == 0 && warp < kDeferChunk / 32) {
    const int64_t i = nChunks * kDeferChunk + tid;
    runFromGlobal<EXACT, RNG_MODE>(C, P, ctl, Q, obsT, (uint32_t)i, 0, i < P.n);
  }
  __syncthreads();
  // ---- per-step sums of the CTA -> global ----
  if (single) {
    for (int r = warp; r < 2 * nSteps; r += kDeferWarps) { // row r = 2 * step + (0: sum E, 1: sum v.E)
      double a = 0.0;
      for (int k = 0; k < kDeferWarps; k++) a += obsT[r * kDeferThreads + 32 * k + lane];
      a = warpSum(a);
      if (lane == 0) atomicAdd(C.obs + (r >> 1) * 3 + (r & 1), a);
    }
    __syncthreads();
  }
  for (int j = tid; j < nSteps * obsPerStep; j += kDeferThreads) {
    double v = 0.0;
    for (int k = 0; k < kDeferWarps; k++) v += C.obs[k * nSteps * obsPerStep + j];
    // one valley: every particle contributes to every step
    if (single && j % 3 == 2) v = blockIdx.x == 0 ? (double)P.n : 0.0;
    if (v != 0.0) atomicAdd(P.obs + j, v);
  }
}

} // namespace emc
