// K1d: several consecutive time steps per launch PAIR -- a flight kernel and an event kernel.
//
// Same per-particle arithmetic as K1a / K1b / K1c (the trajectories are bit-identical; reference:
// examples/bulkSimulation/basicBulkParticleHandler.hpp:181-225 moveParticles, :289-347 observables,
// include/emcParticleDrift.hpp:12-36 drift), organised so that the >= 98 % of the particle-steps that are plain free
// flights run in a kernel that contains nothing else:
//
//   bulkFlightKernel  every lane keeps PPL particles in registers and advances them by full-dt flights, branch-free,
//                     28 FP64 instructions per particle-step (24 for a field along a coordinate axis), all launch constants in the constant bank, the three
//                     Herring-Vogt factors of the particle's sub-valley in registers.  A particle whose flight ends
//                     inside the coming step (tau < dt) FREEZES IN PLACE: its three factors become 0, after which the
//                     same instructions leave k and the position exactly unchanged and contribute exact zeros to the
//                     observables.  After nSteps the whole chunk goes back with vector stores, together with one byte
//                     per particle: the step at which it froze (0xFF = finished).  No queues, no shared-memory staging,
//                     no atomics; the instruction footprint is a few KB.
//   bulkEventKernel   warps scan the byte array, collect the frozen particles in a warp-private list and serve them 32
//                     at a time, one per lane: the complete scattering step of ITS particle at ITS step (flight to the
//                     event, table selection, final state, new flight time, rest of the step: bulkParticleStep), then
//                     full-dt flights to the end of the launch or to the particle's next event.  Random numbers are
//                     Philox(key, particle id, step, draw): the order in which events are served changes nothing.
//
// GRAIN (a grain mechanism is set, emcGrainScatterMechanism): the second free-flight clock of the particle is a ninth fp64
// stream of both kernels (144 B + 1 B of HBM traffic per particle per launch pair instead of 136 + 1).  The reference
// decrements it AFTER the step and scatters when it is <= 0 (basicBulkParticleHandler.hpp:216-220); the flight kernel forms
// g - dt first (one more DADD per particle-step) and freezes the particle at a step whose end would see the clock run out
// (sign test on the high word -- conservative by the subnormals, the event kernel repeats the exact test), so that the
// event kernel serves the step, the grain event at its end and the draws of both in the order of the general kernel.
//
// Observables: per-thread shared-memory slots per step in both kernels (no atomics, no shuffles in the loops); the
// flight kernel sums S - 1 = 2 alpha E instead of E (one FMA from values the flight needs anyway).
#pragma once
#include "emc_bulk_kernel.cuh"

namespace emc {

constexpr int kSplitMaxSteps = 24;     // time steps per launch pair (shared memory of the per-thread observable slots)
constexpr int kFlightThreadsAlone = 512; // flight kernel: 16 warps x PPL particles per lane
constexpr int kEventThreadsAlone = 512;  // event kernel
constexpr int kEventScan = 256;        // bytes of the frozen array a warp scans per refill (8 per lane)
constexpr int kEventListCap = 32 + kEventScan; // a refill starts with fewer than 32 entries
constexpr int kEventDense = 12;        // lanes still busy after an event step from which the batch goes on in place
// particles per claim of a warp of the event kernel: option event_claim (default 1024, emcgpu_internal.cuh)

struct SplitFlightSmem {
  // [nSteps][2][threads] doubles, then the Herring-Vogt factors [EMCGPU_MAX_SUBVALLEYS][4], then the control word
  static __host__ __device__ size_t obsBytes(int nSteps, int threads) { return (size_t)nSteps * 2 * threads * sizeof(double); }
  static __host__ __device__ size_t bytes(int nSteps, int threads) {
    return obsBytes(nSteps, threads) + EMCGPU_MAX_SUBVALLEYS * 4 * sizeof(double) + 16;
  }
};

// signed compare of the bit patterns: for finite doubles with b >= 0 exactly (a < b), including a < 0 and a = -0
__device__ __forceinline__ bool lessThanBits(double a, long long bBits) { return __double_as_longlong(a) < bBits; }

// a += b under a predicate, as one predicated instruction
__device__ __forceinline__ void addIf(double &a, double b, bool p) {
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q add.rn.f64 %0, %0, %1;\n\t}" : "+d"(a) : "d"(b), "r"((uint32_t)p));
}

template <int PPL> struct VecLd;
template <> struct VecLd<2> {
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
  static __device__ __forceinline__ void stFrozen(uint8_t *p, uint32_t f) { *reinterpret_cast<uint16_t *>(p) = (uint16_t)f; }
};
template <> struct VecLd<4> {
  static __device__ __forceinline__ void ld(const double *p, double (&v)[4]) { VecIO<4>::ld(p, v); }
  static __device__ __forceinline__ void st(double *p, const double (&v)[4]) { VecIO<4>::st(p, v); }
  static __device__ __forceinline__ void ldw(const uint32_t *p, uint32_t (&v)[4]) { VecIO<4>::ldw(p, v); }
  static __device__ __forceinline__ void stw(uint32_t *p, const uint32_t (&v)[4]) { VecIO<4>::stw(p, v); }
  static __device__ __forceinline__ void stFrozen(uint8_t *p, uint32_t f) { *reinterpret_cast<uint32_t *>(p) = f; }
};

// ---------------------------------------------------------------------------------------------------------------
// Flight kernel.  Applies to FAST arithmetic, one non-parabolic valley whose sub-valley rotations are signed
// permutations (the Si / Ga2O3 bulk models); everything else runs K1c.
// AXIS: the device axis of a field along a coordinate axis (v.Ê has one term), -1 = any direction.
template <int PPL, int AXIS, int kFlightThreads, bool GRAIN>
__device__ __forceinline__ void flightRole(const BulkParams &P, const int cta, const int nCta) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int nSteps = P.nSteps;
  double *obsT = reinterpret_cast<double *>(smemRaw);
  double *aTab = reinterpret_cast<double *>(smemRaw + SplitFlightSmem::obsBytes(nSteps, kFlightThreads));
  unsigned *nextChunk = reinterpret_cast<unsigned *>(aTab + EMCGPU_MAX_SUBVALLEYS * 4);
  const int tid = threadIdx.x, lane = tid & 31;
  for (int s = 0; s < 2 * nSteps; s++) obsT[s * kFlightThreads + tid] = 0.0;
  if (tid < EMCGPU_MAX_SUBVALLEYS) {
    const DevValley &v = P.model->valleys[0];
    FastSub fs;
    buildFastSub(v, tid < v.deg ? tid : 0, P.force, P.dir, P.dt, fs);
    // AXIS >= 0 (flightCoreAxis): the factors of the two transverse axes doubled
    aTab[4 * tid + 0] = AXIS == 1 || AXIS == 2 ? 2.0 * fs.a[0] : fs.a[0];
    aTab[4 * tid + 1] = AXIS == 0 || AXIS == 2 ? 2.0 * fs.a[1] : fs.a[1];
    aTab[4 * tid + 2] = AXIS == 0 || AXIS == 1 ? 2.0 * fs.a[2] : fs.a[2];
    aTab[4 * tid + 3] = 0.0;
  }
  if (tid == 0) {
    *nextChunk = 0;
    if (cta == 0) *P.claim = 0; // the event kernel of this launch pair starts claiming at 0
  }
  __syncthreads();
  // launch constants: operands from the constant bank
  const FlightConst &f = P.fc[0];
  double *const *const out = P.packedOut ? P.streamOut : P.stream; // out of place: the input ensemble is left as it was
  double *const grainOut = P.packedOut ? P.grainOut : P.grainTau;
  const double dt = P.dt;
  const long long dtBits = __double_as_longlong(dt);
  const uint32_t hiBx = (uint32_t)__double2hiint(P.box.x), hiBy = (uint32_t)__double2hiint(P.box.y),
                 hiBz = (uint32_t)__double2hiint(P.box.z);
  constexpr int kChunk = 32 * PPL;
  const int64_t nChunks = P.n / kChunk;
  const uint32_t obsAddr = smemAddr(obsT) + tid * 8;
  auto claimChunk = [&]() -> int64_t {
    unsigned k = 0;
    if (lane == 0) k = atomicAdd(nextChunk, 1u);
    k = __shfl_sync(0xffffffffu, k, 0);
    return (int64_t)cta + (int64_t)k * nCta;
  };
  auto prefetchChunk = [&](int64_t ch) {
    if (ch >= nChunks) return;
    const int64_t i0 = ch * kChunk + PPL * lane;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_KX] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_KY] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_KZ] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_TAU] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_X] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_Y] + i0));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.stream[EMCGPU_Z] + i0));
    if constexpr (GRAIN) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.grainTau + i0));
    if ((lane & 1) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.packed + i0));
  };

  for (int64_t ch = claimChunk(), chNext; ch < nChunks; ch = chNext) {
    chNext = claimChunk();
    prefetchChunk(chNext);
    const int64_t i0 = ch * kChunk + PPL * lane;
    double kx[PPL], ky[PPL], kz[PPL], tau[PPL], px[PPL], py[PPL], pz[PPL], a0[PPL], a1[PPL], a2[PPL];
    double g[GRAIN ? PPL : 1];
    bool live[PPL];
    uint32_t frz = 0xffffffffu; // byte j: step at which particle j froze
    {
      uint32_t w[PPL];
      VecLd<PPL>::ld(P.stream[EMCGPU_KX] + i0, kx);
      VecLd<PPL>::ld(P.stream[EMCGPU_KY] + i0, ky);
      VecLd<PPL>::ld(P.stream[EMCGPU_KZ] + i0, kz);
      VecLd<PPL>::ld(P.stream[EMCGPU_TAU] + i0, tau);
      VecLd<PPL>::ld(P.stream[EMCGPU_X] + i0, px);
      VecLd<PPL>::ld(P.stream[EMCGPU_Y] + i0, py);
      VecLd<PPL>::ld(P.stream[EMCGPU_Z] + i0, pz);
      if constexpr (GRAIN) VecLd<PPL>::ld(P.grainTau + i0, g);
      VecLd<PPL>::ldw(P.packed + i0, w);
      if (P.packedOut) VecLd<PPL>::stw(P.packedOut + i0, w);
#pragma unroll
      for (int j = 0; j < PPL; j++) {
        const double *a = aTab + 4 * ((w[j] >> 8) & 0xffu);
        const double2 a01 = *reinterpret_cast<const double2 *>(a);
        a0[j] = a01.x;
        a1[j] = a01.y;
        a2[j] = a[2];
        live[j] = true;
      }
    }
#pragma unroll 2
    for (int s = 0; s < nSteps; s++) {
      double sumT, sumV;
      uint32_t wrap = 0;
#pragma unroll
      for (int j = 0; j < PPL; j++) {
        // a particle whose flight ends inside this step freezes: zero factors leave k and the position exactly as
        // they are and make its contributions to the sums exact zeros
        bool ev = live[j] && lessThanBits(tau[j], dtBits);
        double gNext = 0.0;
        if constexpr (GRAIN) { // the clock after this step; <= 0 (or a subnormal): the step ends with a grain event
          gNext = __dsub_rn(g[j], dt);
          ev = ev || (live[j] && __double2hiint(gNext) <= 0);
        }
        if (ev) {
          a0[j] = a1[j] = a2[j] = 0.0;
          live[j] = false;
          frz = (frz & ~(0xffu << (8 * j))) | ((uint32_t)s << (8 * j));
        }
        FlightAux o;
        if constexpr (AXIS < 0)
          flightCore(a0[j], a1[j], a2[j], f.G[0], f.G[1], f.G[2], f.K2, f.c2a, kx[j], ky[j], kz[j], px[j], py[j], pz[j], o);
        else
          flightCoreAxis<AXIS>(a0[j], a1[j], a2[j], f.G[AXIS], f.K2, f.c2a, kx[j], ky[j], kz[j], px[j], py[j], pz[j], o);
        wrap |= (uint32_t)mayNeedWrap(px[j], hiBx) | (uint32_t)mayNeedWrap(py[j], hiBy) | (uint32_t)mayNeedWrap(pz[j], hiBz);
        // live: tau -= dt, sumT += S - 1 (predicated adds); a frozen particle contributes nothing
        const double t = flightSm1(o);
        if (j == 0) sumT = live[j] ? t : 0.0;
        else addIf(sumT, t, live[j]);
        addIf(tau[j], -dt, live[j]);
        if constexpr (GRAIN) g[j] = live[j] ? gNext : g[j];
        const double v = AXIS == 0   ? (f.K4[0] * kx[j]) * o.w0
                         : AXIS == 1 ? (f.K4[1] * ky[j]) * o.w1
                         : AXIS == 2 ? (f.K4[2] * kz[j]) * o.w2
                                     : flightVelocityDt(f.K4[0], f.K4[1], f.K4[2], kx[j], ky[j], kz[j], o);
        if (j == 0) sumV = v;
        else sumV += v;
      }
      if (__any_sync(0xffffffffu, wrap)) { // rare: a particle of the warp is at (or past) a face of the box
#pragma unroll
        for (int j = 0; j < PPL; j++) {
          px[j] = wrapExact(px[j], P.box.x);
          py[j] = wrapExact(py[j], P.box.y);
          pz[j] = wrapExact(pz[j], P.box.z);
        }
      }
      const uint32_t oa = obsAddr + (uint32_t)s * (2 * kFlightThreads * 8);
      double accE, accV;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(accE) : "r"(oa));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(accV) : "r"(oa + kFlightThreads * 8));
      accE += sumT;
      accV += sumV;
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(oa), "d"(accE) : "memory");
      asm volatile("st.shared.f64 [%0], %1;" ::"r"(oa + kFlightThreads * 8), "d"(accV) : "memory");
    }
    // the chunk goes home; the energy of the final state as every other kernel computes it (getEnergy of the final k)
    double en[PPL];
#pragma unroll
    for (int j = 0; j < PPL; j++) {
      FlightAux o;
      o.sq = fma(kz[j], kz[j], fma(ky[j], ky[j], kx[j] * kx[j]));
      o.x = fma(f.c2a, o.sq, 1.0);
      o.r = rsqrtNormal(o.x);
      en[j] = flightEnergy(f.fE, o);
    }
    VecLd<PPL>::st(out[EMCGPU_KX] + i0, kx);
    VecLd<PPL>::st(out[EMCGPU_KY] + i0, ky);
    VecLd<PPL>::st(out[EMCGPU_KZ] + i0, kz);
    VecLd<PPL>::st(out[EMCGPU_ENERGY] + i0, en);
    VecLd<PPL>::st(out[EMCGPU_TAU] + i0, tau);
    VecLd<PPL>::st(out[EMCGPU_X] + i0, px);
    VecLd<PPL>::st(out[EMCGPU_Y] + i0, py);
    VecLd<PPL>::st(out[EMCGPU_Z] + i0, pz);
    if constexpr (GRAIN) VecLd<PPL>::st(grainOut + i0, g);
    VecLd<PPL>::stFrozen(P.frozen + i0, frz);
  }
  // the particles behind the last whole chunk are left to the event kernel, from step 0
  if (cta == 0) {
    const int64_t i = nChunks * kChunk + tid;
    if (tid < kChunk && i < P.n) {
      P.frozen[i] = 0;
      if (P.packedOut) {
        for (int c = 0; c < EMCGPU_N_STREAMS; c++) P.streamOut[c][i] = P.stream[c][i];
        if constexpr (GRAIN) P.grainOut[i] = P.grainTau[i];
        P.packedOut[i] = P.packed[i];
      }
    }
  }
  __syncthreads();
  // ---- per-step sums of the CTA -> global: sum E = sum (S - 1) / (2 alpha) ----
  constexpr int kWarps = kFlightThreads / 32;
  const int warp = tid >> 5;
  for (int r = warp; r < 2 * nSteps; r += kWarps) { // row r = 2 * step + (0: sum S - 1, 1: sum v.E)
    double a = 0.0;
    for (int k = 0; k < kWarps; k++) a += obsT[r * kFlightThreads + 32 * k + lane];
    a = warpSum(a);
    if (lane == 0) {
      if ((r & 1) == 0) a *= f.inv2a;
      if (a != 0.0) atomicAdd(P.obs + (r >> 1) * 3 + (r & 1), a);
    }
  }
  // one valley: every particle contributes to every step
  if (cta == 0 && tid < nSteps) atomicAdd(P.obs + tid * 3 + 2, (double)P.n);
}
template <int PPL, int AXIS, bool GRAIN>
__global__ void __launch_bounds__(kFlightThreadsAlone, 1) bulkFlightKernel(const __grid_constant__ BulkParams P) {
  flightRole<PPL, AXIS, kFlightThreadsAlone, GRAIN>(P, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------------------------------------------
// Event kernel.
struct EventWarpList {
  uint32_t idx[kEventListCap];
  uint8_t step[kEventListCap];
};

template <int RNG_MODE, int kEventThreads, bool GRAIN>
__device__ __forceinline__ void eventRole(const BulkParams &P) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint64_t tableBar;
  const int nSteps = P.nSteps;
  constexpr int kWarps = kEventThreads / 32;
  const BulkSmem L(0, 1, P.nMechTotal, P.model->tableDoubles, P.tablesInSmem != 0, 0);
  double *obsT = reinterpret_cast<double *>(smemRaw + ((L.total + 15) & ~size_t(15))); // [nSteps][2][kEventThreads]
  EventWarpList *lists = reinterpret_cast<EventWarpList *>(obsT + (size_t)nSteps * 2 * kEventThreads);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int s = 0; s < 2 * nSteps; s++) obsT[s * kEventThreads + tid] = 0.0;
  const CtaState C = stageCta(P, smemRaw, &tableBar, 0, 0); // __syncthreads inside
  EventWarpList &list = lists[warp];
  double *const myObs = obsT + tid;
  const double dt = P.dt;
  const unsigned ltMask = (1u << lane) - 1u;
  int count = 0;              // entries in the warp's list (warp-uniform)
  int64_t scanAt = 0, scanEnd = 0; // the warp's current claim of the frozen array
  bool exhausted = false;

  for (;;) {
    // ---- refill: scan the frozen array until 32 particles are listed or nothing is left ----
    while (count < 32 && !exhausted) {
      if (scanAt >= scanEnd) {
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(P.claim, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        scanAt = (int64_t)c * P.eventClaim;
        scanEnd = min(scanAt + (int64_t)P.eventClaim, P.n);
        if (scanAt >= P.n) {
          exhausted = true;
          break;
        }
      }
      // 8 bytes per lane (the claims start on multiples of 256, the array is padded to a multiple of 256)
      const int64_t at = scanAt + 8 * lane;
      uint2 fl = make_uint2(0xffffffffu, 0xffffffffu);
      if (at < scanEnd) fl = *reinterpret_cast<const uint2 *>(P.frozen + at);
      int mine = 0;
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const uint32_t byte = ((b < 4 ? fl.x : fl.y) >> (8 * (b & 3))) & 0xffu;
        mine += (byte != 0xffu && at + b < scanEnd) ? 1 : 0;
      }
      int incl = mine; // inclusive prefix sum over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      int pos = count + incl - mine;
#pragma unroll
      for (int b = 0; b < 8; b++) {
        const uint32_t byte = ((b < 4 ? fl.x : fl.y) >> (8 * (b & 3))) & 0xffu;
        if (byte != 0xffu && at + b < scanEnd) {
          list.idx[pos] = (uint32_t)(at + b);
          list.step[pos] = (uint8_t)byte;
          pos++;
        }
      }
      count += __shfl_sync(0xffffffffu, incl, 31);
      scanAt += kEventScan;
      __syncwarp();
    }
    if (count == 0) break;
    // ---- one batch: the last min(32, count) entries, one per lane ----
    const int take = min(32, count);
    count -= take;
    bool active = lane < take;
    uint32_t idx = 0;
    int s = 0;
    Particle p;
    Rng rng;
    p.k = Vec3{0.0, 0.0, 0.0};
    p.pos = Vec3{0.0, 0.0, 0.0};
    p.energy = p.tau = 0.0;
    p.valley = p.sub = p.region = 0;
    double g = 0.0; // grain clock
    if (active) {
      idx = list.idx[count + lane];
      s = list.step[count + lane];
      loadParticle(P, idx, p, rng);
      if constexpr (GRAIN) g = P.grainTau[idx];
    }
    __syncwarp();
    for (bool first = true;; first = false) {
      if (active) {
        const FastSub &fs = C.fast[p.valley * EMCGPU_MAX_SUBVALLEYS + p.sub];
        const FlightConst &f = C.fastV[p.valley].f; // a signed-permutation valley (splitEligible): the a-form of the flight
        // the full-dt flights between events: fastStep without the accurate energy of every step -- the energy
        // observable takes (S - 1) / (2 alpha) like the flight kernel, the stored energy is formed once after the last flight
        FlightAux o;
        bool flown = false;
        while (s < nSteps && p.tau >= dt) {
          if constexpr (GRAIN) { // a step that ends with a grain event is served below
            const double gNext = __dsub_rn(g, dt);
            if (gNext <= 0.0) break;
            g = gNext;
          }
          flightCore(fs.a[0], fs.a[1], fs.a[2], f.G[0], f.G[1], f.G[2], f.K2, f.c2a, p.k.x, p.k.y, p.k.z, p.pos.x, p.pos.y,
                     p.pos.z, o);
          if (mayNeedWrap(p.pos.x, (uint32_t)__double2hiint(P.box.x)) || mayNeedWrap(p.pos.y, (uint32_t)__double2hiint(P.box.y)) ||
              mayNeedWrap(p.pos.z, (uint32_t)__double2hiint(P.box.z))) {
            p.pos.x = wrapExact(p.pos.x, P.box.x);
            p.pos.y = wrapExact(p.pos.y, P.box.y);
            p.pos.z = wrapExact(p.pos.z, P.box.z);
          }
          p.tau -= dt;
          myObs[(2 * s) * kEventThreads] += flightSm1(o) * f.inv2a;
          myObs[(2 * s + 1) * kEventThreads] += flightVelocityDt(f.K4[0], f.K4[1], f.K4[2], p.k.x, p.k.y, p.k.z, o);
          s++;
          flown = true;
        }
        if (flown) p.energy = flightEnergy(f.fE, o); // as fastStep leaves it (getEnergy of the final k)
        if (s == nSteps) {
          storeParticleState(P, idx, p);
          if constexpr (GRAIN) P.grainTau[idx] = g;
          active = false;
        }
      }
      const unsigned busy = __ballot_sync(0xffffffffu, active);
      if (!busy) break;
      if (!first && __popc(busy) < kEventDense) {
        // few lanes left: their particles wait in the list for a full batch (state parked in global memory)
        if (active) {
          storeParticleState(P, idx, p);
          if constexpr (GRAIN) P.grainTau[idx] = g;
          const int pos = count + __popc(busy & ltMask);
          list.idx[pos] = idx;
          list.step[pos] = (uint8_t)s;
        }
        count += __popc(busy);
        break;
      }
      if (active) { // every busy lane is at the step in which its flight ends (or, GRAIN, its grain clock runs out)
        rng.n = 0;
        rng.step = (uint32_t)(P.step0 + s);
        attachReplay<RNG_MODE>(P, idx, rng);
        double vd = bulkParticleStep<false, RNG_MODE>(C, P, p, rng, P.idBase + idx, P.step0 + s);
        if constexpr (GRAIN) { // basicBulkParticleHandler.hpp:216-220, as the general kernel does it
          g = __dsub_rn(g, dt);
          if (g <= 0.0) {
            g = grainEvent<RNG_MODE>(P, p, rng);
            vd = driftVelocity<false>(C.model->valleys[p.valley], p.sub, p.k, p.energy, P.dir);
          }
        }
        if constexpr (RNG_MODE == RNG_REPLAY) storeCursor(P, idx, rng);
        myObs[(2 * s) * kEventThreads] += p.energy;
        myObs[(2 * s + 1) * kEventThreads] += vd;
        s++;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int r = warp; r < 2 * nSteps; r += kWarps) {
    double a = 0.0;
    for (int k = 0; k < kWarps; k++) a += obsT[r * kEventThreads + 32 * k + lane];
    a = warpSum(a);
    if (lane == 0 && a != 0.0) atomicAdd(P.obs + (r >> 1) * 3 + (r & 1), a);
  }
}
template <int RNG_MODE, bool GRAIN>
__global__ void __launch_bounds__(kEventThreadsAlone, 1) bulkEventKernel(const __grid_constant__ BulkParams P) {
  eventRole<RNG_MODE, kEventThreadsAlone, GRAIN>(P);
}

__host__ __device__ inline size_t splitEventSmemBytes(const BulkSmem &L, int nSteps, int threads) {
  return ((L.total + 15) & ~size_t(15)) + (size_t)nSteps * 2 * threads * sizeof(double) +
         (threads / 32) * sizeof(EventWarpList);
}

} // namespace emc
