// Internals shared by the translation units of libemcgpu.so: the context behind the
// opaque emcgpu_ctx handle of include/emcgpu.h, device buffers and error helpers.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "emc_bulk_kernel.cuh"

namespace emc {

struct DeviceBuffer {
  void *ptr = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&ptr, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <class T> T *as() const { return static_cast<T *>(ptr); }
};

struct DeviceRunState; // emcgpu_device.cu

} // namespace emc

struct emcgpu_ctx {
  int device = 0;
  int smCount = 0;
  int maxSmemOptin = 0;
  int maxSmemPerSm = 0;
  int optVec = 2;          // particles per lane and iteration of the streaming step kernel (1, 2, 4)
  int optKernel = 0;       // one-step kernel: 0 = TMA pipeline when it fits, 1 = plain streaming kernel
  int optStages = 0;       // cap on the TMA ring depth (0 = as many as fit)
  int optDeferTablesSmem = 0; // K1c: 1 = stage the rate tables in shared memory (default: read them through L1/L2)
  int optMultiKernel = 0;  // several steps per launch: 0 = auto (K1d / K1c when the ensemble fills the machine), 1 = in place (K1b), 2 = K1c, 3 = K1d
  int optEventClaim = 1024;     // K1d event kernel: particles per claim of a warp
  int sorClusterWide = -1; // red-black cluster solver on 16 CTAs of 512 threads: -1 undecided, 0 no, 1 yes
  long long sorClusterWideKey = -1; // the grid that decision was made for
  int sorFast = 0;          // red-black cluster solver, fast 2-D form: 0 no, 1 on 8 x 1024 threads, 2 on 16 x 512
  long long sorFastKey = -1; // the grid that decision was made for
  int optSplitPpl = 4;     // K1d flight kernel: particles per lane (2 or 4)
  int optTablesGlobal = 0; // 1 = leave the rate tables in global memory / L2 (more ring stages)
  int optSorKernel = 0;    // 0 = row-per-thread wavefront when it fits, 1 = hyperplane loop
  int optEarlyStep = 1;    // device runs: the step kernel stages its tables while the Poisson solver ahead of it still runs
  int optAssignFp64 = 0;   // 1 = NEC / NEC-VWD deposits as fp64 atomics per corner instead of integer hits per mesh cell
  int optSorOrder = 0;     // 0 = the reference's lexicographic order (bit-identical iterates), 1 = red-black
  int optPoissonInterval = 1; // device run: solve Poisson every n-th step (emcSimulation::setPoissonInterval)
  cudaStream_t stream = nullptr;
  std::string error;
  int64_t launches = 0;

  // model
  bool haveValleys = false, haveTables = false;
  emc::DevModel hModel{};
  std::vector<emc::DevMech> hMechs;
  emc::DeviceBuffer dModel, dMechs, dTables;

  // ensemble
  int64_t n = 0, capacity = 0, idBase = 0;
  emc::DeviceBuffer dEnsemble;
  double *dStream[EMCGPU_N_STREAMS] = {};
  uint32_t *dPacked = nullptr;
  // emcgpu_bulk_step_ahead / emcgpu_bulk_rewind: the ensemble as it was before the last look-ahead call
  emc::DeviceBuffer dEnsembleAlt;
  double *dStreamAlt[EMCGPU_N_STREAMS] = {};
  uint32_t *dPackedAlt = nullptr;
  bool rewindValid = false;
  int64_t rewindStep = 0;
  // emcgpu_bulk_record_velocities: per-particle velocities of the steps of the next step calls, streamed to the host
  int velComponents = 0;
  double *velHost = nullptr;
  int64_t velHostSteps = 0;
  emc::DeviceBuffer dVel[2];
  cudaEvent_t velDone[2] = {}, velCopied[2] = {};
  // "kernel_timing": cudaEvents around the launches of the flight (0) and event (1) kernels, other kernels (2)
  bool optTiming = false;
  struct TimedLaunch {
    int tag;
    cudaEvent_t a, b;
  };
  std::vector<TimedLaunch> timed;
  double timedMs[3] = {0, 0, 0};
  int64_t timedLaunches[3] = {0, 0, 0};

  // slice ring of emcgpu_bulk_run_host (three slices: upload / advance / download)
  emc::DeviceBuffer dSlices;
  cudaStream_t copyIn = nullptr, copyOut = nullptr;
  cudaEvent_t sliceEvents[9] = {};
  bool obsAccumulate = false; // emcgpu_bulk_step_device adds to obsDevice instead of zeroing it first

  // rng
  int rngMode = emc::RNG_PHILOX;
  uint64_t seed = 0;
  emc::DeviceBuffer dDraws, dOffsets, dCursor;

  // bulk configuration
  bool bulkConfigured = false;
  emc::Vec3 box{}, force{}, dir{};
  int mathMode = EMCGPU_MATH_EXACT;
  int64_t nextStep = 1;

  // phonon baths of polar-optical mechanisms (emcgpu_set_phonon_baths)
  int nBaths = 0, nBathBins = 0;
  double bathDq = 0;
  bool bathHasCum = false;
  emc::DeviceBuffer dBathCounts, dBathCum;

  // grain boundaries (emcgpu_set_grain): clock per particle
  bool grainOn = false, grainClockSet = false;
  double grainProb = 0.5, grainTau0 = 1.0;
  emc::DeviceBuffer dGrain, dGrainAlt; // Alt: the clocks of the look-ahead copy (emcgpu_bulk_step_ahead)

  // outputs
  emc::DeviceBuffer dObs, dStatus, dEvents, dEvCount;
  emc::DeviceBuffer dFrozen, dClaim; // split step (K1d): byte per particle, claim counter of the event kernel
  int64_t evCap = 0;

  // device-run path (emcgpu_device.cu), created by emcgpu_device_configure
  emc::DeviceRunState *run = nullptr;
};

namespace emc {

int failWith(emcgpu_ctx *ctx, int code, const char *fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return emc::failWith(ctx, EMCGPU_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

inline void fillGrain(const emcgpu_ctx *ctx, BulkParams &P) {
  P.grainTau = ctx->grainOn ? ctx->dGrain.as<double>() : nullptr;
  P.grainOut = nullptr;
  P.grainProb = ctx->grainProb;
  P.grainTau0 = ctx->grainTau0;
}

inline void fillBathView(const emcgpu_ctx *ctx, BathView &B) {
  B.counts = ctx->nBaths ? ctx->dBathCounts.as<unsigned long long>() : nullptr;
  B.cumW = ctx->bathHasCum ? ctx->dBathCum.as<const double>() : nullptr;
  B.cumWN = ctx->bathHasCum ? ctx->dBathCum.as<const double>() + (size_t)ctx->nBaths * (ctx->nBathBins + 1) : nullptr;
  B.nBaths = ctx->nBaths;
  B.nBins = ctx->nBathBins;
  B.dq = ctx->bathDq;
}

inline int bindDevice(emcgpu_ctx *ctx) {
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  return EMCGPU_OK;
}
// compacted-ensemble bookkeeping shared with the device-run code
int allocEnsembleStreams(emcgpu_ctx *ctx, int64_t n);
void releaseDeviceRun(emcgpu_ctx *ctx);

} // namespace emc
