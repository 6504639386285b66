// C-ABI implementation (include/emcgpu.h): context, model upload, ensemble
// management and kernel launches.  No CPU fallback: every compute entry point
// needs a CUDA device and fails with EMCGPU_E_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "emcgpu_internal.cuh"
#include "emc_bulk_defer.cuh"
#include "emc_bulk_split.cuh"

using namespace emc;

namespace {

thread_local std::string g_createError;

} // namespace

namespace emc {
int failWith(emcgpu_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->error = buf; else g_createError = buf;
  return code;
}
} // namespace emc

namespace {

#define fail emc::failWith

int bind(emcgpu_ctx *ctx) { return emc::bindDevice(ctx); }

// classify R (row-major, rows = ellipse axes): identity, signed permutation, general
int classifyRotation(const double *r, uint16_t *toE, uint16_t *toD) {
  bool identity = true, perm = true;
  int src[3] = {-1, -1, -1}, sign[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) {
    int nz = 0;
    for (int j = 0; j < 3; j++) {
      const double x = r[i * 3 + j];
      if (x != (i == j ? 1.0 : 0.0)) identity = false;
      if (x != 0.0) {
        nz++;
        if (x == 1.0 || x == -1.0) {
          src[i] = j;
          sign[i] = x < 0 ? 1 : 0;
        } else {
          perm = false;
        }
      }
    }
    if (nz != 1) perm = false;
  }
  if (perm && (src[0] == src[1] || src[0] == src[2] || src[1] == src[2])) perm = false;
  *toE = *toD = 0;
  if (!perm) return ROT_GENERAL;
  for (int i = 0; i < 3; i++) {
    // toEllipse: out[i] = sign_i * in[src_i]; toDevice: out[src_i] = sign_i * in[i]
    *toE |= (uint16_t)((src[i] | (sign[i] << 2)) << (4 * i));
    *toD |= (uint16_t)((i | (sign[i] << 2)) << (4 * src[i]));
  }
  return identity ? ROT_IDENTITY : ROT_SIGNED_PERMUTATION;
}

cudaError_t uploadModel(emcgpu_ctx *ctx) {
  cudaError_t e = ctx->dModel.ensure(sizeof(DevModel));
  if (e != cudaSuccess) return e;
  return cudaMemcpyAsync(ctx->dModel.ptr, &ctx->hModel, sizeof(DevModel), cudaMemcpyHostToDevice, ctx->stream);
}

size_t bulkSmemBytes(const emcgpu_ctx *ctx, int nObsDoubles, bool tablesInSmem, int queueWords) {
  return BulkSmem(nObsDoubles, ctx->hModel.nValleys, (int)ctx->hMechs.size(), ctx->hModel.tableDoubles, tablesInSmem,
                  queueWords)
      .total;
}

void fillBulkParams(emcgpu_ctx *ctx, BulkParams &P) {
  memset(&P, 0, sizeof P);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) P.stream[s] = ctx->dStream[s];
  P.packed = ctx->dPacked;
  P.n = ctx->n;
  P.idBase = ctx->idBase;
  P.model = static_cast<const DevModel *>(ctx->dModel.ptr);
  P.tables = static_cast<const double *>(ctx->dTables.ptr);
  P.mechs = static_cast<const DevMech *>(ctx->dMechs.ptr);
  P.nMechTotal = (int32_t)ctx->hMechs.size();
  P.box = ctx->box;
  P.force = ctx->force;
  P.dir = ctx->dir;
  P.seed = ctx->seed;
  P.draws = static_cast<const uint64_t *>(ctx->dDraws.ptr);
  P.offsets = static_cast<const int64_t *>(ctx->dOffsets.ptr);
  P.cursor = static_cast<uint32_t *>(ctx->dCursor.ptr);
  P.events = static_cast<long long *>(ctx->dEvents.ptr);
  P.evCap = ctx->evCap;
  P.evCount = static_cast<unsigned long long *>(ctx->dEvCount.ptr);
  P.status = static_cast<int *>(ctx->dStatus.ptr);
  emc::fillBathView(ctx, P.baths);
  emc::fillGrain(ctx, P);
}

// Launch-uniform flight constants of every valley (FlightConst, emc_device.cuh).  Plain IEEE products and quotients on
// the host (this file is compiled with -ffp-contract=off): every kernel of a launch reads the same numbers.
void buildFlightConsts(const emcgpu_ctx *ctx, BulkParams &P) {
  for (int v = 0; v < ctx->hModel.nValleys && v < EMCGPU_MAX_VALLEYS; v++) {
    const DevValley &dv = ctx->hModel.valleys[v];
    FlightConst &f = P.fc[v];
    f.fE = dv.nonParabolic ? dv.fE : 2.0 * dv.fE;
    f.c2a = 2.0 * dv.alpha * f.fE;
    f.inv2a = dv.nonParabolic && dv.alpha != 0.0 ? 1.0 / (2.0 * dv.alpha) : 0.0;
    const double force[3] = {P.force.x, P.force.y, P.force.z}, dir[3] = {P.dir.x, P.dir.y, P.dir.z};
    f.KP = kHbar / (2.0 * dv.mCond);
    f.K2 = f.KP * P.dt;
    for (int i = 0; i < 3; i++) {
      f.Fh[i] = force[i] / kHbar;
      f.G[i] = f.Fh[i] * P.dt;
      f.KV[i] = dir[i] * (2.0 * f.KP);
      f.K4[i] = f.KV[i] / f.K2;
    }
    // (the diagonal form has one mass; the anisotropic single-layer class moves with m_c and reports velocities with m_DOS)
    f.diag = dv.rotKind != ROT_GENERAL && dv.mBand == dv.mCond;
    f.nonParabolic = dv.nonParabolic;
  }
}

// "kernel_timing": a pair of events around a launch, folded into per-kind totals by emcgpu_kernel_times
cudaError_t timedBegin(emcgpu_ctx *ctx, int tag) {
  if (!ctx->optTiming) return cudaSuccess;
  emcgpu_ctx::TimedLaunch t{tag, nullptr, nullptr};
  cudaError_t e = cudaEventCreate(&t.a);
  if (e == cudaSuccess) e = cudaEventCreate(&t.b);
  if (e == cudaSuccess) e = cudaEventRecord(t.a, ctx->stream);
  if (e == cudaSuccess) ctx->timed.push_back(t);
  return e;
}
void timedEnd(emcgpu_ctx *ctx) {
  if (ctx->optTiming && !ctx->timed.empty()) cudaEventRecord(ctx->timed.back().b, ctx->stream);
}
template <typename K>
cudaError_t launchKernel(emcgpu_ctx *ctx, K kernel, const BulkParams &P, size_t smem, int grid,
                         int threads = kBulkThreads, int tag = 2) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if ((e = timedBegin(ctx, tag)) != cudaSuccess) return e;
  kernel<<<grid, threads, smem, ctx->stream>>>(P);
  timedEnd(ctx);
  ctx->launches++;
  return cudaGetLastError();
}

// K1b, several steps per launch
cudaError_t launchFused(emcgpu_ctx *ctx, const BulkParams &P, size_t smem, int grid) {
  const bool exact = ctx->mathMode == EMCGPU_MATH_EXACT;
  if (ctx->rngMode == RNG_PHILOX)
    return exact ? launchKernel(ctx, bulkStepKernel<true, RNG_PHILOX>, P, smem, grid)
                 : launchKernel(ctx, bulkStepKernel<false, RNG_PHILOX>, P, smem, grid);
  return exact ? launchKernel(ctx, bulkStepKernel<true, RNG_REPLAY>, P, smem, grid)
               : launchKernel(ctx, bulkStepKernel<false, RNG_REPLAY>, P, smem, grid);
}

// K1c, several steps per launch, scattering events deferred into a CTA-wide queue
cudaError_t launchDefer(emcgpu_ctx *ctx, const BulkParams &P, size_t smem, int grid) {
  const bool exact = ctx->mathMode == EMCGPU_MATH_EXACT;
  if (ctx->rngMode == RNG_PHILOX)
    return exact ? launchKernel(ctx, bulkDeferKernel<true, RNG_PHILOX>, P, smem, grid, kDeferThreads)
                 : launchKernel(ctx, bulkDeferKernel<false, RNG_PHILOX>, P, smem, grid, kDeferThreads);
  return exact ? launchKernel(ctx, bulkDeferKernel<true, RNG_REPLAY>, P, smem, grid, kDeferThreads)
               : launchKernel(ctx, bulkDeferKernel<false, RNG_REPLAY>, P, smem, grid, kDeferThreads);
}
// shared memory of K1c for `steps` steps per launch; 0 = does not fit
size_t deferSmem(const emcgpu_ctx *ctx, int steps, bool tablesInSmem) {
  const BulkSmem L(kDeferWarps * steps * ctx->hModel.nValleys * 3, ctx->hModel.nValleys, (int)ctx->hMechs.size(),
                   ctx->hModel.tableDoubles, tablesInSmem, kDeferQueueWords);
  const size_t b = deferSmemBytes(L, steps, ctx->hModel.nValleys);
  return b <= (size_t)ctx->maxSmemOptin ? b : 0;
}

// K1d, several steps per launch pair: flight kernel + event kernel (emc_bulk_split.cuh)
bool splitEligible(const emcgpu_ctx *ctx) {
  const DevValley &v = ctx->hModel.valleys[0];
  return ctx->mathMode == EMCGPU_MATH_FAST && ctx->hModel.nValleys == 1 && v.rotKind != ROT_GENERAL && v.nonParabolic &&
         v.alpha > 0.0 && v.mBand == v.mCond;
}
size_t splitEventSmem(const emcgpu_ctx *ctx, int steps, int threads) {
  const BulkSmem L(0, 1, (int)ctx->hMechs.size(), ctx->hModel.tableDoubles, false, 0);
  return splitEventSmemBytes(L, steps, threads);
}
// a field along a coordinate axis: the drift-velocity observable of the flight kernel has one term
int fieldAxis(const BulkParams &P) {
  const double d[3] = {P.dir.x, P.dir.y, P.dir.z};
  int axis = -1;
  for (int i = 0; i < 3; i++)
    if (d[i] != 0.0 && d[(i + 1) % 3] == 0.0 && d[(i + 2) % 3] == 0.0) axis = i;
  return axis;
}
template <typename F> cudaError_t dispatchFlight(int ppl, int axis, F &&go) {
  if (ppl == 2)
    return axis == 0 ? go(std::integral_constant<int, 2>{}, std::integral_constant<int, 0>{})
           : axis == 1 ? go(std::integral_constant<int, 2>{}, std::integral_constant<int, 1>{})
           : axis == 2 ? go(std::integral_constant<int, 2>{}, std::integral_constant<int, 2>{})
                       : go(std::integral_constant<int, 2>{}, std::integral_constant<int, -1>{});
  return axis == 0 ? go(std::integral_constant<int, 4>{}, std::integral_constant<int, 0>{})
         : axis == 1 ? go(std::integral_constant<int, 4>{}, std::integral_constant<int, 1>{})
         : axis == 2 ? go(std::integral_constant<int, 4>{}, std::integral_constant<int, 2>{})
                     : go(std::integral_constant<int, 4>{}, std::integral_constant<int, -1>{});
}
// the streams of an ensemble of n particles inside one allocation: every stream starts on a 256-byte boundary
void layoutStreams(void *basePtr, int64_t n, double **streams, uint32_t **packed) {
  const size_t strideD = ((size_t)n * sizeof(double) + 255) & ~size_t(255);
  unsigned char *base = static_cast<unsigned char *>(basePtr);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) streams[s] = reinterpret_cast<double *>(base + strideD * s);
  *packed = reinterpret_cast<uint32_t *>(base + strideD * EMCGPU_N_STREAMS);
}
size_t ensembleBytes(int64_t n) {
  const size_t strideD = ((size_t)n * sizeof(double) + 255) & ~size_t(255);
  const size_t strideP = ((size_t)n * sizeof(uint32_t) + 255) & ~size_t(255);
  return strideD * EMCGPU_N_STREAMS + strideP;
}
// the resident ensemble and the look-ahead copy change places
void swapEnsembles(emcgpu_ctx *ctx) {
  std::swap(ctx->dEnsemble, ctx->dEnsembleAlt);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) std::swap(ctx->dStream[s], ctx->dStreamAlt[s]);
  std::swap(ctx->dPacked, ctx->dPackedAlt);
  if (ctx->grainOn) std::swap(ctx->dGrain, ctx->dGrainAlt);
}
// one launch pair on the whole shard: flight kernel, then event kernel.  outOfPlace: the flight kernel writes the
// look-ahead copy (which becomes the resident ensemble), the input stays as it was.
cudaError_t launchSplit(emcgpu_ctx *ctx, BulkParams &P, bool outOfPlace) {
  const int ppl = ctx->optSplitPpl == 2 ? 2 : 4;
  const size_t smemFlight = SplitFlightSmem::bytes(P.nSteps, kFlightThreadsAlone);
  const size_t smemEvent = splitEventSmem(ctx, P.nSteps, kEventThreadsAlone);
  const int64_t nChunks = P.n / (32 * ppl);
  const int warps = kFlightThreadsAlone / 32;
  const int gridFlight = (int)std::max<int64_t>(1, std::min<int64_t>((nChunks + warps - 1) / warps, ctx->smCount));
  P.eventClaim = ctx->optEventClaim;
  const int64_t claims = (P.n + P.eventClaim - 1) / P.eventClaim;
  const int gridEvent = (int)std::max<int64_t>(1, std::min<int64_t>((claims + warps - 1) / warps, ctx->smCount));
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) P.streamOut[s] = outOfPlace ? ctx->dStreamAlt[s] : nullptr;
  P.packedOut = outOfPlace ? ctx->dPackedAlt : nullptr;
  const bool grain = ctx->grainOn;
  P.grainOut = outOfPlace && grain ? ctx->dGrainAlt.as<double>() : nullptr;
  cudaError_t e = dispatchFlight(ppl, fieldAxis(P), [&](auto p, auto a) {
    return grain ? launchKernel(ctx, bulkFlightKernel<decltype(p)::value, decltype(a)::value, true>, P, smemFlight, gridFlight,
                                kFlightThreadsAlone, 0)
                 : launchKernel(ctx, bulkFlightKernel<decltype(p)::value, decltype(a)::value, false>, P, smemFlight, gridFlight,
                                kFlightThreadsAlone, 0);
  });
  if (e != cudaSuccess) return e;
  if (outOfPlace) {
    swapEnsembles(ctx);
    for (int s = 0; s < EMCGPU_N_STREAMS; s++) P.stream[s] = ctx->dStream[s];
    P.packed = ctx->dPacked;
    P.packedOut = nullptr;
    if (grain) P.grainTau = ctx->dGrain.as<double>();
    P.grainOut = nullptr;
  }
  auto event = [&](auto kernel) { return launchKernel(ctx, kernel, P, smemEvent, gridEvent, kEventThreadsAlone, 1); };
  if (ctx->rngMode == RNG_PHILOX) return grain ? event(bulkEventKernel<RNG_PHILOX, true>) : event(bulkEventKernel<RNG_PHILOX, false>);
  return grain ? event(bulkEventKernel<RNG_REPLAY, true>) : event(bulkEventKernel<RNG_REPLAY, false>);
}
// K steps of the whole shard with the flight / event kernels, `window` steps per launch pair
int runSplit(emcgpu_ctx *ctx, BulkParams &P, int nSteps, int window, double *obsDevice, bool keep) {
  const size_t flagBytes = ((size_t)ctx->n + 255) & ~size_t(255);
  CUDA_TRY(ctx, ctx->dFrozen.ensure(flagBytes));
  CUDA_TRY(ctx, ctx->dClaim.ensure(256));
  P.frozen = ctx->dFrozen.as<uint8_t>();
  P.claim = ctx->dClaim.as<unsigned>();
  P.tablesInSmem = 0;
  for (int done = 0; done < nSteps;) {
    const int w = std::min(window, nSteps - done);
    P.nSteps = w;
    P.step0 = ctx->nextStep + done;
    P.obs = obsDevice + (size_t)done * 3; // one valley
    cudaError_t e = launchSplit(ctx, P, keep && done == 0);
    if (e != cudaSuccess) return fail(ctx, EMCGPU_E_CUDA, "bulk step launch failed: %s", cudaGetErrorString(e));
    done += w;
  }
  return EMCGPU_OK;
}

// K1a, one step per launch
template <int VEC> cudaError_t launchStreamVec(emcgpu_ctx *ctx, const BulkParams &P, size_t smem, int grid) {
  const bool exact = ctx->mathMode == EMCGPU_MATH_EXACT;
  if (ctx->rngMode == RNG_PHILOX)
    return exact ? launchKernel(ctx, bulkStreamKernel<true, RNG_PHILOX, VEC>, P, smem, grid, kStreamThreads)
                 : launchKernel(ctx, bulkStreamKernel<false, RNG_PHILOX, VEC>, P, smem, grid, kStreamThreads);
  return exact ? launchKernel(ctx, bulkStreamKernel<true, RNG_REPLAY, VEC>, P, smem, grid, kStreamThreads)
               : launchKernel(ctx, bulkStreamKernel<false, RNG_REPLAY, VEC>, P, smem, grid, kStreamThreads);
}
// K1a/TMA, one step per launch, warp-specialised TMA pipeline
cudaError_t launchTma(emcgpu_ctx *ctx, const BulkParams &P, size_t smem, int grid, int stages) {
  const bool exact = ctx->mathMode == EMCGPU_MATH_EXACT;
  auto go = [&](auto kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kTmaThreads, smem, ctx->stream>>>(P, stages);
    ctx->launches++;
    return cudaGetLastError();
  };
  if (ctx->rngMode == RNG_PHILOX)
    return exact ? go(bulkTmaKernel<true, RNG_PHILOX>) : go(bulkTmaKernel<false, RNG_PHILOX>);
  return exact ? go(bulkTmaKernel<true, RNG_REPLAY>) : go(bulkTmaKernel<false, RNG_REPLAY>);
}
// ring stages that fit beside the model (and the tables) in shared memory; 0 = does not fit
int tmaStages(const emcgpu_ctx *ctx, bool tablesInSmem, size_t *smemOut) {
  const BulkSmem L(ctx->hModel.nValleys * 3, ctx->hModel.nValleys, (int)ctx->hMechs.size(), ctx->hModel.tableDoubles,
                   tablesInSmem, kTmaQueueWords);
  const size_t ring = tmaRingOffset(L);
  if (ring + (size_t)kMinStages * kTileBytes > (size_t)ctx->maxSmemOptin) return 0;
  const int stages = (int)std::min<size_t>(kMaxStages, ((size_t)ctx->maxSmemOptin - ring) / kTileBytes);
  *smemOut = ring + (size_t)stages * kTileBytes;
  return stages;
}

int streamQueueWords(int vec) { return (kStreamThreads / 32) * (32 + 32 * vec); }

int checkReady(emcgpu_ctx *ctx, bool needEnsemble) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!ctx->haveValleys) return fail(ctx, EMCGPU_E_INVALID, "emcgpu_set_valleys has not been called");
  if (!ctx->haveTables) return fail(ctx, EMCGPU_E_INVALID, "emcgpu_set_tables has not been called");
  if (!ctx->bulkConfigured) return fail(ctx, EMCGPU_E_INVALID, "emcgpu_bulk_configure has not been called");
  if (needEnsemble && ctx->n <= 0) return fail(ctx, EMCGPU_E_INVALID, "the ensemble is empty");
  return EMCGPU_OK;
}

int allocEnsemble(emcgpu_ctx *ctx, int64_t n) { return emc::allocEnsembleStreams(ctx, n); }
} // namespace
int emc::allocEnsembleStreams(emcgpu_ctx *ctx, int64_t n) {
  // every stream starts on a 256-byte boundary
  const size_t strideD = ((size_t)n * sizeof(double) + 255) & ~size_t(255);
  const size_t strideP = ((size_t)n * sizeof(uint32_t) + 255) & ~size_t(255);
  CUDA_TRY(ctx, ctx->dEnsemble.ensure(strideD * EMCGPU_N_STREAMS + strideP));
  unsigned char *base = static_cast<unsigned char *>(ctx->dEnsemble.ptr);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) ctx->dStream[s] = reinterpret_cast<double *>(base + strideD * s);
  ctx->dPacked = reinterpret_cast<uint32_t *>(base + strideD * EMCGPU_N_STREAMS);
  ctx->n = n;
  ctx->capacity = n;
  ctx->rewindValid = false;
  return EMCGPU_OK;
}
namespace {

int checkStatusWord(emcgpu_ctx *ctx) {
  int status = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&status, ctx->dStatus.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (status != 0) {
    cudaMemsetAsync(ctx->dStatus.ptr, 0, sizeof(int), ctx->stream);
    if (status == EMCGPU_E_REPLAY_EXHAUSTED)
      return fail(ctx, status, "replay stream exhausted: a particle needed more draws than were recorded");
    return fail(ctx, status, "device reported status %d", status);
  }
  return EMCGPU_OK;
}

} // namespace

extern "C" {

int emcgpu_abi_version(void) { return EMCGPU_ABI_VERSION; }

int emcgpu_create(int cudaDevice, emcgpu_ctx **out) {
  if (!out) return fail(nullptr, EMCGPU_E_INVALID, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, EMCGPU_E_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (cudaDevice < 0 || cudaDevice >= count)
    return fail(nullptr, EMCGPU_E_INVALID, "cudaDevice %d out of range [0,%d)", cudaDevice, count);
  emcgpu_ctx *ctx = new emcgpu_ctx();
  ctx->device = cudaDevice;
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(cudaDevice)) != cudaSuccess ||
      (e = cudaGetDeviceProperties(&prop, cudaDevice)) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, EMCGPU_E_CUDA, "cannot open device %d: %s", cudaDevice, cudaGetErrorString(e));
  }
  if (prop.major != 10) {
    delete ctx;
    return fail(nullptr, EMCGPU_E_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only",
                cudaDevice, prop.major, prop.minor);
  }
  ctx->smCount = prop.multiProcessorCount;
  if (const char *e = std::getenv("EMCGPU_SOR_KERNEL")) // developer switch for unmodified drivers: option sor_kernel
    ctx->optSorKernel = std::max(0, std::min(3, std::atoi(e)));
  if (const char *e = std::getenv("EMCGPU_EARLY_STEP")) // developer switch: option early_step
    ctx->optEarlyStep = std::atoi(e) != 0;
  ctx->maxSmemOptin = (int)prop.sharedMemPerBlockOptin;
  ctx->maxSmemPerSm = (int)prop.sharedMemPerMultiprocessor;
  if ((e = ctx->dStatus.ensure(sizeof(int))) != cudaSuccess ||
      (e = ctx->dEvCount.ensure(sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMemset(ctx->dStatus.ptr, 0, sizeof(int))) != cudaSuccess ||
      (e = cudaMemset(ctx->dEvCount.ptr, 0, sizeof(unsigned long long))) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, EMCGPU_E_CUDA, "device allocation failed: %s", cudaGetErrorString(e));
  }
  *out = ctx;
  return EMCGPU_OK;
}

void emcgpu_destroy(emcgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (DeviceBuffer *b : {&ctx->dModel, &ctx->dMechs, &ctx->dTables, &ctx->dEnsemble, &ctx->dDraws,
                          &ctx->dOffsets, &ctx->dCursor, &ctx->dObs, &ctx->dStatus, &ctx->dEvents,
                          &ctx->dEvCount, &ctx->dSlices, &ctx->dEnsembleAlt, &ctx->dFrozen, &ctx->dClaim, &ctx->dBathCounts, &ctx->dBathCum, &ctx->dGrain, &ctx->dGrainAlt})
    b->release();
  for (cudaEvent_t &e : ctx->sliceEvents)
    if (e) cudaEventDestroy(e);
  for (int b = 0; b < 2; b++) {
    ctx->dVel[b].release();
    if (ctx->velDone[b]) cudaEventDestroy(ctx->velDone[b]);
    if (ctx->velCopied[b]) cudaEventDestroy(ctx->velCopied[b]);
  }
  for (auto &t : ctx->timed) {
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  if (ctx->copyIn) cudaStreamDestroy(ctx->copyIn);
  if (ctx->copyOut) cudaStreamDestroy(ctx->copyOut);
  emc::releaseDeviceRun(ctx);
  delete ctx;
}

const char *emcgpu_last_error(const emcgpu_ctx *ctx) {
  return ctx ? ctx->error.c_str() : g_createError.c_str();
}

int64_t emcgpu_launch_count(const emcgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int emcgpu_set_stream(emcgpu_ctx *ctx, void *cudaStream) {
  if (!ctx) return EMCGPU_E_INVALID;
  ctx->stream = static_cast<cudaStream_t>(cudaStream);
  return EMCGPU_OK;
}

int emcgpu_set_option(emcgpu_ctx *ctx, const char *name, int64_t value) {
  if (!ctx || !name) return EMCGPU_E_INVALID;
  if (!strcmp(name, "vec")) {
    if (value != 1 && value != 2 && value != 4) return fail(ctx, EMCGPU_E_INVALID, "vec must be 1, 2 or 4");
    ctx->optVec = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "kernel")) {
    if (value != 0 && value != 1) return fail(ctx, EMCGPU_E_INVALID, "kernel must be 0 (TMA pipeline) or 1 (streaming)");
    ctx->optKernel = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "tables_global")) {
    ctx->optTablesGlobal = value != 0;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "sor_kernel")) {
    if (value < 0 || value > 3)
      return fail(ctx, EMCGPU_E_INVALID, "sor_kernel must be 0 (default: rows / fastest cluster form), 1 (hyperplanes / one CTA), 2 (red-black: general cluster kernel) or 3 (red-black: fast form on the portable cluster of 8 CTAs)");
    ctx->optSorKernel = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "early_step")) {
    ctx->optEarlyStep = value != 0;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "assign_fp64")) {
    ctx->optAssignFp64 = value != 0;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "sor_order")) {
    if (value != 0 && value != 1) return fail(ctx, EMCGPU_E_INVALID, "sor_order must be 0 (lexicographic) or 1 (red-black)");
    ctx->optSorOrder = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "poisson_interval")) {
    if (value < 1) return fail(ctx, EMCGPU_E_INVALID, "poisson_interval must be at least 1");
    ctx->optPoissonInterval = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "stages")) {
    ctx->optStages = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "defer_tables_smem")) {
    ctx->optDeferTablesSmem = value != 0;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "multi_kernel")) {
    if (value < 0 || value > 3)
      return fail(ctx, EMCGPU_E_INVALID, "multi_kernel must be 0 (flight + event kernels / deferred events when the ensemble is large), 1 (in place), 2 (deferred events always) or 3 (flight + event kernels always, where the model allows)");
    ctx->optMultiKernel = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "kernel_timing")) {
    ctx->optTiming = value != 0;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "event_claim")) {
    if (value < 256 || value > (1 << 20) || value % 256) return fail(ctx, EMCGPU_E_INVALID, "event_claim must be a multiple of 256 in [256, 2^20] (particles per claim of a warp of the event kernel)");
    ctx->optEventClaim = (int)value;
    return EMCGPU_OK;
  }
  if (!strcmp(name, "split_ppl")) {
    if (value != 2 && value != 4) return fail(ctx, EMCGPU_E_INVALID, "split_ppl must be 2 or 4 (particles per lane of the flight kernel)");
    ctx->optSplitPpl = (int)value;
    return EMCGPU_OK;
  }
  return fail(ctx, EMCGPU_E_INVALID, "unknown option '%s'", name);
}

int emcgpu_set_grain(emcgpu_ctx *ctx, double transmissionProbability, double scatterRate) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (transmissionProbability < 0 || transmissionProbability > 1) return fail(ctx, EMCGPU_E_INVALID, "transmission probability outside [0, 1]");
  ctx->grainOn = scatterRate > 0;
  ctx->grainProb = transmissionProbability;
  ctx->grainTau0 = ctx->grainOn ? 1.0 / scatterRate : 1.0;
  return EMCGPU_OK;
}

int emcgpu_set_grain_clock(emcgpu_ctx *ctx, const double *grainTau) {
  if (!ctx || !grainTau) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  CUDA_TRY(ctx, ctx->dGrain.ensure((size_t)std::max<int64_t>(1, ctx->capacity) * sizeof(double)));
  if (ctx->n)
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dGrain.ptr, grainTau, (size_t)ctx->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->grainClockSet = true;
  return EMCGPU_OK;
}

int emcgpu_get_grain_clock(emcgpu_ctx *ctx, double *grainTau) {
  if (!ctx || !grainTau) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  if (!ctx->grainClockSet) return fail(ctx, EMCGPU_E_INVALID, "no grain clocks on the device");
  if (ctx->n)
    CUDA_TRY(ctx, cudaMemcpyAsync(grainTau, ctx->dGrain.ptr, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_set_phonon_baths(emcgpu_ctx *ctx, int nBaths, int nBins, double dq, const double *cumW, const double *cumWN) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  if (nBaths < 0 || nBaths > EMCGPU_MAX_BATHS) return fail(ctx, EMCGPU_E_CAPACITY, "%d phonon baths exceed EMCGPU_MAX_BATHS=%d", nBaths, EMCGPU_MAX_BATHS);
  if (nBaths > 0 && (nBins < 1 || !(dq > 0))) return fail(ctx, EMCGPU_E_INVALID, "bad phonon bath binning");
  if ((cumW == nullptr) != (cumWN == nullptr)) return fail(ctx, EMCGPU_E_INVALID, "cumW and cumWN go together");
  const bool sameShape = nBaths == ctx->nBaths && nBins == ctx->nBathBins;
  const size_t countBytes = (size_t)std::max(1, nBaths) * 2 * std::max(1, nBins) * sizeof(unsigned long long);
  CUDA_TRY(ctx, ctx->dBathCounts.ensure(countBytes));
  if (!sameShape) CUDA_TRY(ctx, cudaMemsetAsync(ctx->dBathCounts.ptr, 0, countBytes, ctx->stream)); // counters survive an update of the sums
  ctx->nBaths = nBaths;
  ctx->nBathBins = nBins;
  ctx->bathDq = dq;
  ctx->bathHasCum = cumW != nullptr && nBaths > 0;
  if (ctx->bathHasCum) {
    const size_t n = (size_t)nBaths * (nBins + 1);
    CUDA_TRY(ctx, ctx->dBathCum.ensure(2 * n * sizeof(double)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dBathCum.ptr, cumW, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dBathCum.as<double>() + n, cumWN, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return EMCGPU_OK;
}

int emcgpu_get_phonon_counts(emcgpu_ctx *ctx, int64_t *emission, int64_t *absorption, int reset) {
  if (!ctx || !emission || !absorption) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  if (ctx->nBaths == 0) return EMCGPU_OK;
  const size_t per = (size_t)ctx->nBathBins;
  std::vector<unsigned long long> h((size_t)ctx->nBaths * 2 * per);
  CUDA_TRY(ctx, cudaMemcpyAsync(h.data(), ctx->dBathCounts.ptr, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  if (reset) CUDA_TRY(ctx, cudaMemsetAsync(ctx->dBathCounts.ptr, 0, h.size() * sizeof(unsigned long long), ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (int b = 0; b < ctx->nBaths; b++)
    for (size_t i = 0; i < per; i++) {
      emission[b * per + i] = (int64_t)h[((size_t)b * 2 + 0) * per + i];
      absorption[b * per + i] = (int64_t)h[((size_t)b * 2 + 1) * per + i];
    }
  return EMCGPU_OK;
}

int emcgpu_synchronize(emcgpu_ctx *ctx) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_set_valleys(emcgpu_ctx *ctx, const emcgpu_valley_t *valleys, int nValleys) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!valleys || nValleys < 1) return fail(ctx, EMCGPU_E_INVALID, "need at least one valley");
  if (nValleys > EMCGPU_MAX_VALLEYS)
    return fail(ctx, EMCGPU_E_CAPACITY, "%d valleys exceed EMCGPU_MAX_VALLEYS=%d", nValleys, EMCGPU_MAX_VALLEYS);
  if (int r = bind(ctx)) return r;
  DevModel &M = ctx->hModel;
  M.nValleys = nValleys;
  for (int i = 0; i < nValleys; i++) {
    const emcgpu_valley_t &in = valleys[i];
    if (in.kind < 0 || in.kind > EMCGPU_VALLEY_NONPARABOLIC_ANISOTROP_SINGLE_LAYER || in.kind == 6)
      return fail(ctx, EMCGPU_E_UNSUPPORTED_VALLEY, "valley %d: unknown valley class %d", i, in.kind);
    if (in.degeneracy < 1 || in.degeneracy > EMCGPU_MAX_SUBVALLEYS)
      return fail(ctx, EMCGPU_E_CAPACITY, "valley %d: degeneracy %d outside [1,%d]", i, in.degeneracy,
                  EMCGPU_MAX_SUBVALLEYS);
    if (!(in.effMassCond > 0) || in.alpha < 0)
      return fail(ctx, EMCGPU_E_INVALID, "valley %d: effective mass must be > 0 and alpha >= 0", i);
    DevValley &v = M.valleys[i];
    memset(&v, 0, sizeof v);
    v.kind = in.kind;
    v.deg = in.degeneracy;
    v.nonParabolic = in.kind & 1;
    v.mCond = in.effMassCond;
    v.alpha = v.nonParabolic ? in.alpha : 0.0;
    v.eBottom = in.bottomEnergy;
    const bool aniso = (in.kind & 2) != 0, singleLayer = (in.kind & 4) != 0;
    for (int d = 0; d < 3; d++) v.vogt[d] = aniso ? in.vogt[d] : 1.0;
    if (singleLayer) v.vogt[2] = 0.0; // no motion out of the plane
    v.mBand = v.mCond;
    if (in.kind == EMCGPU_VALLEY_NONPARABOLIC_ANISOTROP_SINGLE_LAYER) {
      if (!(in.effMassDOS > 0)) return fail(ctx, EMCGPU_E_INVALID, "valley %d: the single-layer class needs effMassDOS > 0", i);
      v.mBand = in.effMassDOS;
    }
    v.xMq = v.mBand * kQ;
    v.xTwoMq = 2 * v.mBand * kQ;
    v.fE = v.nonParabolic ? kHbar * kHbar / (v.mBand * kQ) : kHbar * kHbar / (2 * v.mBand * kQ);
    for (int d = 0; d < 3; d++) {
      v.fPos[d] = kHbar * v.vogt[d] / (2 * v.mCond);
      v.fVel[d] = kHbar * v.vogt[d] / v.mBand;
      v.fDk[d] = v.vogt[d] / kHbar;
    }
    int worst = ROT_IDENTITY;
    for (int s = 0; s < EMCGPU_MAX_SUBVALLEYS; s++) {
      double ident[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
      const double *r = (aniso && s < in.degeneracy) ? in.rot[s] : ident;
      memcpy(v.rot[s], r, sizeof ident);
      const int kind = classifyRotation(r, &v.permToE[s], &v.permToD[s]);
      if (kind > worst) worst = kind;
    }
    v.rotKind = worst;
  }
  ctx->haveValleys = true;
  if (ctx->haveTables) CUDA_TRY(ctx, uploadModel(ctx));
  return EMCGPU_OK;
}

int emcgpu_set_tables(emcgpu_ctx *ctx, const emcgpu_tableset_t *sets, int nSets, int nLevels,
                      double maxEnergy) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!ctx->haveValleys) return fail(ctx, EMCGPU_E_INVALID, "call emcgpu_set_valleys before emcgpu_set_tables");
  if (nSets < 0 || (nSets > 0 && !sets) || nLevels < 1 || !(maxEnergy > 0))
    return fail(ctx, EMCGPU_E_INVALID, "bad table arguments");
  if (nSets > EMCGPU_MAX_TABLESETS)
    return fail(ctx, EMCGPU_E_CAPACITY, "%d table sets exceed EMCGPU_MAX_TABLESETS=%d", nSets,
                EMCGPU_MAX_TABLESETS);
  if (int r = bind(ctx)) return r;
  DevModel &M = ctx->hModel;
  M.nSets = nSets;
  M.nLevels = nLevels;
  M.dE = maxEnergy / nLevels;
  M.defaultTau = 2e-15;
  M.nRegions = 0;
  memset(M.setOf, -1, sizeof M.setOf);
  std::vector<DevMech> mechs;
  std::vector<double> tables;
  for (int i = 0; i < nSets; i++) {
    const emcgpu_tableset_t &in = sets[i];
    if (in.valley < 0 || in.valley >= M.nValleys)
      return fail(ctx, EMCGPU_E_INVALID, "table set %d: valley %d does not exist", i, in.valley);
    if (in.region < 0 || in.region >= kMaxRegions)
      return fail(ctx, EMCGPU_E_CAPACITY, "table set %d: region %d outside [0,%d)", i, in.region, kMaxRegions);
    if (in.nMech < 1 || in.nMech > EMCGPU_MAX_MECH_PER_SET)
      return fail(ctx, EMCGPU_E_CAPACITY, "table set %d: %d mechanisms outside [1,%d]", i, in.nMech,
                  EMCGPU_MAX_MECH_PER_SET);
    if (!in.cum || !in.mech) return fail(ctx, EMCGPU_E_INVALID, "table set %d: NULL tables", i);
    if (M.setOf[in.valley][in.region] >= 0)
      return fail(ctx, EMCGPU_E_INVALID, "duplicate table set for valley %d region %d", in.valley, in.region);
    M.setOf[in.valley][in.region] = (int8_t)i;
    if (in.region + 1 > M.nRegions) M.nRegions = in.region + 1;
    DevTableSet &ts = M.sets[i];
    ts.nMech = in.nMech;
    ts.stride = (in.nMech + 1) & ~1;
    ts.tabOffset = (int32_t)tables.size();
    ts.mechOffset = (int32_t)mechs.size();
    ts.tau = in.tau;
    // transpose to level-major rows so that one selection touches one short row
    tables.resize(tables.size() + (size_t)nLevels * ts.stride, 2.0 /* padding: never selected */);
    for (int m = 0; m < in.nMech; m++)
      for (int l = 0; l < nLevels; l++)
        tables[ts.tabOffset + (size_t)l * ts.stride + m] = in.cum[(size_t)m * nLevels + l];
    for (int m = 0; m < in.nMech; m++) {
      const emcgpu_mech_t &mi = in.mech[m];
      char name[EMCGPU_NAME_LEN + 1];
      memcpy(name, mi.name, EMCGPU_NAME_LEN);
      name[EMCGPU_NAME_LEN] = 0;
      if (mi.sampler <= EMCGPU_SAMPLER_NONE || mi.sampler > EMCGPU_SAMPLER_SINGLE_LAYER_SCREENED_OPTICAL)
        return fail(ctx, EMCGPU_E_UNSUPPORTED_MECHANISM,
                    "scatter mechanism '%s' (valley %d, region %d) has no device sampler; it cannot run on "
                    "the GPU path and there is no CPU fallback",
                    name, in.valley, in.region);
      DevMech d;
      memset(&d, 0, sizeof d);
      d.sampler = mi.sampler;
      d.finalValley = mi.finalValley;
      d.nFinal = mi.nFinal;
      d.mechId = mi.mechId;
      d.param[0] = mi.param[0];
      d.param[1] = mi.param[1];
      d.param[2] = mi.param[2];
      d.bath = -1;
      if (mi.sampler == EMCGPU_SAMPLER_FROEHLICH || mi.sampler == EMCGPU_SAMPLER_SCREENED_FROEHLICH) {
        d.bath = mi.param[2] >= 0 ? (int32_t)mi.param[2] : -1;
        d.flags = mi.param[3] != 0 ? 1 : 0;
        if (d.bath >= ctx->nBaths)
          return fail(ctx, EMCGPU_E_INVALID, "mechanism '%s' refers to phonon bath %d: call emcgpu_set_phonon_baths first", name,
                      d.bath);
        if ((d.flags & 1) && (d.bath < 0 || !ctx->bathHasCum))
          return fail(ctx, EMCGPU_E_INVALID, "mechanism '%s' samples |q| from a phonon bath whose prefix sums were not given", name);
      }
      if (mi.sampler == EMCGPU_SAMPLER_INTERVALLEY || mi.sampler == EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY) {
        if (mi.finalValley < 0 || mi.finalValley >= M.nValleys)
          return fail(ctx, EMCGPU_E_INVALID, "mechanism '%s': final valley %d does not exist", name, mi.finalValley);
        // the single-layer classes have a constructor without a sub-valley map: the sub-valley index is kept
        const int minFinal = mi.sampler == EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY ? 0 : 1;
        if (mi.nFinal < minFinal || mi.nFinal > EMCGPU_MAX_FINAL)
          return fail(ctx, EMCGPU_E_CAPACITY, "mechanism '%s': %d final sub-valleys outside [%d,%d]", name,
                      mi.nFinal, minFinal, EMCGPU_MAX_FINAL);
        if (mi.nFinal == 0 && M.valleys[in.valley].deg > M.valleys[mi.finalValley].deg)
          return fail(ctx, EMCGPU_E_INVALID, "mechanism '%s' keeps the sub-valley index but the final valley has fewer", name);
        const int degF = M.valleys[mi.finalValley].deg;
        for (int s = 0; s < M.valleys[in.valley].deg; s++)
          for (int f = 0; f < mi.nFinal; f++)
            if (mi.finalSub[s][f] >= degF)
              return fail(ctx, EMCGPU_E_INVALID, "mechanism '%s': final sub-valley %d >= degeneracy %d", name,
                          mi.finalSub[s][f], degF);
        memcpy(d.finalSub, mi.finalSub, sizeof d.finalSub);
      }
      mechs.push_back(d);
    }
  }
  if (tables.empty()) tables.resize(2, 2.0);
  M.tableDoubles = (int64_t)tables.size();
  ctx->hMechs = mechs;
  CUDA_TRY(ctx, ctx->dTables.ensure(tables.size() * sizeof(double)));
  CUDA_TRY(ctx, ctx->dMechs.ensure(std::max<size_t>(1, mechs.size()) * sizeof(DevMech)));
  // synchronous copies: the staging vectors die at the end of this call
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaMemcpy(ctx->dTables.ptr, tables.data(), tables.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (!mechs.empty())
    CUDA_TRY(ctx, cudaMemcpy(ctx->dMechs.ptr, mechs.data(), mechs.size() * sizeof(DevMech), cudaMemcpyHostToDevice));
  ctx->haveTables = true;
  CUDA_TRY(ctx, uploadModel(ctx));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_set_ensemble(emcgpu_ctx *ctx, int64_t n, const double *const *soa, const uint32_t *packed,
                        int64_t particleIdBase) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (n < 0 || (n > 0 && (!soa || !packed))) return fail(ctx, EMCGPU_E_INVALID, "bad ensemble arguments");
  if (int r = bind(ctx)) return r;
  ctx->idBase = particleIdBase;
  if (n == 0) {
    ctx->n = 0;
    return EMCGPU_OK;
  }
  ctx->grainClockSet = false; // clocks belong to an ensemble
  if (int r = allocEnsemble(ctx, n)) return r;
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) {
    if (!soa[s]) return fail(ctx, EMCGPU_E_INVALID, "stream %d is NULL", s);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dStream[s], soa[s], n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dPacked, packed, n * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->rngMode == RNG_REPLAY) ctx->rngMode = RNG_PHILOX; // replay streams belong to the old ensemble
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_get_ensemble(emcgpu_ctx *ctx, double *const *soa, uint32_t *packed) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  if (ctx->n == 0) return EMCGPU_OK;
  if (!soa) return fail(ctx, EMCGPU_E_INVALID, "soa is NULL");
  for (int s = 0; s < EMCGPU_N_STREAMS; s++)
    if (soa[s])
      CUDA_TRY(ctx, cudaMemcpyAsync(soa[s], ctx->dStream[s], ctx->n * sizeof(double), cudaMemcpyDeviceToHost,
                                    ctx->stream));
  if (packed)
    CUDA_TRY(ctx, cudaMemcpyAsync(packed, ctx->dPacked, ctx->n * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                  ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int64_t emcgpu_ensemble_size(const emcgpu_ctx *ctx) { return ctx ? ctx->n : 0; }

int emcgpu_ensemble_device_ptrs(emcgpu_ctx *ctx, double **soaOut, uint32_t **packedOut) {
  if (!ctx || !soaOut) return EMCGPU_E_INVALID;
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) soaOut[s] = ctx->dStream[s];
  if (packedOut) *packedOut = ctx->dPacked;
  return EMCGPU_OK;
}

int emcgpu_generate_bulk_ensemble(emcgpu_ctx *ctx, int64_t n, const double box[3], double temperature,
                                  int32_t region, uint64_t seed, int64_t particleIdBase) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!ctx->haveValleys || !ctx->haveTables)
    return fail(ctx, EMCGPU_E_INVALID, "set valleys and tables before generating an ensemble");
  if (n < 1 || !box || !(temperature > 0)) return fail(ctx, EMCGPU_E_INVALID, "bad arguments");
  if (int r = bind(ctx)) return r;
  if (int r = allocEnsemble(ctx, n)) return r;
  ctx->idBase = particleIdBase;
  GenParams G;
  memset(&G, 0, sizeof G);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) G.stream[s] = ctx->dStream[s];
  G.packed = ctx->dPacked;
  G.n = n;
  G.idBase = particleIdBase;
  G.model = static_cast<const DevModel *>(ctx->dModel.ptr);
  G.box = Vec3{box[0], box[1], box[2]};
  G.thermalVoltage = kKB / kQ * temperature; // emcDevice.hpp:87
  G.region = region;
  G.seed = seed;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->smCount * 16);
  bulkGenerateKernel<<<grid, 256, 0, ctx->stream>>>(G);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  if (ctx->rngMode == RNG_REPLAY) ctx->rngMode = RNG_PHILOX;
  return EMCGPU_OK;
}

int emcgpu_rng_philox(emcgpu_ctx *ctx, uint64_t seed) {
  if (!ctx) return EMCGPU_E_INVALID;
  ctx->rngMode = RNG_PHILOX;
  ctx->seed = seed;
  return EMCGPU_OK;
}

int emcgpu_rng_replay(emcgpu_ctx *ctx, const uint64_t *draws, const int64_t *offsets, int64_t n) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!draws || !offsets || n != ctx->n)
    return fail(ctx, EMCGPU_E_INVALID, "replay streams must cover exactly the %lld uploaded particles",
                (long long)ctx->n);
  if (int r = bind(ctx)) return r;
  const int64_t total = offsets[n];
  CUDA_TRY(ctx, ctx->dDraws.ensure(std::max<int64_t>(1, total) * sizeof(uint64_t)));
  CUDA_TRY(ctx, ctx->dOffsets.ensure((n + 1) * sizeof(int64_t)));
  CUDA_TRY(ctx, ctx->dCursor.ensure(n * sizeof(uint32_t)));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (total > 0) CUDA_TRY(ctx, cudaMemcpy(ctx->dDraws.ptr, draws, total * sizeof(uint64_t), cudaMemcpyHostToDevice));
  CUDA_TRY(ctx, cudaMemcpy(ctx->dOffsets.ptr, offsets, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
  CUDA_TRY(ctx, cudaMemset(ctx->dCursor.ptr, 0, n * sizeof(uint32_t)));
  ctx->rngMode = RNG_REPLAY;
  return EMCGPU_OK;
}

int emcgpu_bulk_configure(emcgpu_ctx *ctx, const double box[3], const double fieldDirection[3],
                          double fieldStrength, double charge, int mathMode) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!box || !fieldDirection) return fail(ctx, EMCGPU_E_INVALID, "NULL argument");
  if (mathMode != EMCGPU_MATH_EXACT && mathMode != EMCGPU_MATH_FAST)
    return fail(ctx, EMCGPU_E_INVALID, "unknown math mode %d", mathMode);
  for (int d = 0; d < 3; d++)
    if (!(box[d] > 0)) return fail(ctx, EMCGPU_E_INVALID, "box extent must be positive");
  // basicBulkParticleHandler ctor (:100-102): normalize(dir); field = dir*strength
  double dir[3] = {fieldDirection[0], fieldDirection[1], fieldDirection[2]};
  double sq = 0;
  for (int d = 0; d < 3; d++) sq += dir[d] * dir[d];
  const double nrm = std::sqrt(sq);
  if (nrm != 0)
    for (int d = 0; d < 3; d++) dir[d] /= nrm;
  ctx->box = Vec3{box[0], box[1], box[2]};
  ctx->dir = Vec3{dir[0], dir[1], dir[2]};
  // moveParticles (:186): force = scale(appliedField, charge)
  ctx->force = Vec3{dir[0] * fieldStrength * charge, dir[1] * fieldStrength * charge,
                    dir[2] * fieldStrength * charge};
  ctx->mathMode = mathMode;
  ctx->bulkConfigured = true;
  return EMCGPU_OK;
}

int emcgpu_set_step_index(emcgpu_ctx *ctx, int64_t nextStep) {
  if (!ctx) return EMCGPU_E_INVALID;
  ctx->nextStep = nextStep;
  return EMCGPU_OK;
}
int64_t emcgpu_get_step_index(const emcgpu_ctx *ctx) { return ctx ? ctx->nextStep : 0; }

} // extern "C"
namespace {
// keep: the ensemble as it is now survives the call (emcgpu_bulk_step_ahead), see emcgpu_bulk_rewind
int bulkStepDevice(emcgpu_ctx *ctx, double dt, int nSteps, int stepsPerLaunch, double *obsDevice, bool keep) {
  if (int r = checkReady(ctx, true)) return r;
  if (!(dt > 0) || nSteps < 1 || !obsDevice) return fail(ctx, EMCGPU_E_INVALID, "bad step arguments");
  if (int r = bind(ctx)) return r;
  if (stepsPerLaunch < 1) stepsPerLaunch = 1;
  if (stepsPerLaunch > kMaxStepsPerLaunch) stepsPerLaunch = kMaxStepsPerLaunch;
  const int nV = ctx->hModel.nValleys;
  if (!ctx->obsAccumulate)
    CUDA_TRY(ctx, cudaMemsetAsync(obsDevice, 0, (size_t)nSteps * nV * 3 * sizeof(double), ctx->stream));
  BulkParams P;
  fillBulkParams(ctx, P);
  P.dt = dt;
  buildFlightConsts(ctx, P);
  if (ctx->n >= (int64_t)1 << 32) return fail(ctx, EMCGPU_E_CAPACITY, "at most 2^32-1 particles per context");
  // per-particle velocities of every step go to the host: the general step kernel writes them, chunk by chunk through two
  // device buffers whose download (copy-out stream) overlaps the next chunk's steps
  const bool record = ctx->velComponents != 0;
  int recordChunk = 0;
  if (record) {
    if (nSteps > ctx->velHostSteps) return fail(ctx, EMCGPU_E_CAPACITY, "the velocity record holds %lld steps, %d requested", (long long)ctx->velHostSteps, nSteps);
    const size_t perStep = (size_t)ctx->n * ctx->velComponents * sizeof(double);
    recordChunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::min(nSteps, kMaxStepsPerLaunch), ((size_t)256 << 20) / std::max<size_t>(1, perStep)));
    if (!ctx->copyOut) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copyOut, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
      CUDA_TRY(ctx, ctx->dVel[b].ensure(perStep * recordChunk));
      if (!ctx->velDone[b]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->velDone[b], cudaEventDisableTiming));
      if (!ctx->velCopied[b]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->velCopied[b], cudaEventDisableTiming));
    }
  }
  if (ctx->grainOn && !ctx->grainClockSet)
    return fail(ctx, EMCGPU_E_INVALID, "a grain mechanism is set but the grain clocks were not uploaded (emcgpu_set_grain_clock)");
  if (keep && ctx->grainOn) CUDA_TRY(ctx, ctx->dGrainAlt.ensure(ctx->dGrain.bytes));
  // several steps per launch, plain model: flight kernel + event kernel (K1d) for ensembles that fill the machine
  if (!record) {
    const int ppl = ctx->optSplitPpl == 2 ? 2 : 4;
    const bool split = stepsPerLaunch > 1 && nSteps > 1 && splitEligible(ctx) &&
                       (ctx->optMultiKernel == 3 ||
                        (ctx->optMultiKernel == 0 && ctx->n / (32 * ppl) >= (int64_t)ctx->smCount * (kFlightThreadsAlone / 32)));
    if (split) {
      int window = std::min(stepsPerLaunch, kSplitMaxSteps);
      while (window > 1 && std::max(SplitFlightSmem::bytes(window, kFlightThreadsAlone),
                                    splitEventSmem(ctx, window, kEventThreadsAlone)) > (size_t)ctx->maxSmemOptin)
        window--;
      if (window > 1) {
        if (int r = runSplit(ctx, P, nSteps, window, obsDevice, keep)) return r;
        ctx->nextStep += nSteps;
        return EMCGPU_OK;
      }
    }
  }
  if (keep) { // the other step kernels work in place: copy first
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dEnsembleAlt.ptr, ctx->dEnsemble.ptr, ensembleBytes(ctx->n), cudaMemcpyDeviceToDevice,
                                  ctx->stream));
    if (ctx->grainOn)
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->dGrainAlt.ptr, ctx->dGrain.ptr, (size_t)ctx->n * sizeof(double), cudaMemcpyDeviceToDevice,
                                    ctx->stream));
  }
  int recordIdx = 0;
  for (int done = 0; done < nSteps;) {
    int chunk = std::min(stepsPerLaunch, nSteps - done);
    if (record) {
      chunk = std::min(recordChunk, nSteps - done);
      const int b = recordIdx & 1;
      if (recordIdx >= 2) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->velCopied[b], 0)); // its last download is done
      P.velOut = ctx->dVel[b].as<double>();
      P.velComponents = ctx->velComponents;
    }
    // several steps per launch: deferred-event kernel (K1c) for ensembles that fill the machine
    const int64_t nChunks = ctx->n / kDeferChunk;
    const bool defer = !record && chunk > 1 && !ctx->grainOn &&
                       (ctx->optMultiKernel >= 2 || (ctx->optMultiKernel == 0 && nChunks >= (int64_t)ctx->smCount * kDeferWarps));
    if (defer) chunk = std::min(chunk, kDeferMaxSteps);
    P.nSteps = chunk;
    P.step0 = ctx->nextStep + done;
    P.obs = obsDevice + (size_t)done * nV * 3;
    if (ctx->grainOn && !ctx->grainClockSet)
      return fail(ctx, EMCGPU_E_INVALID, "a grain mechanism is set but the grain clocks were not uploaded (emcgpu_set_grain_clock)");
    if (defer) {
      // K1c reads a table row once per event (1.1 % of the particle-steps): the 80 KB set stays L1/L2-resident and the
      // shared memory it would take is worth more as L1 (measured: +7.6 %); "defer_tables_smem" = 1 stages it anyway
      bool inSmem = ctx->optDeferTablesSmem && !ctx->optTablesGlobal;
      size_t smem = inSmem ? deferSmem(ctx, chunk, true) : 0;
      if (!smem) {
        inSmem = false;
        smem = deferSmem(ctx, chunk, false);
      }
      if (smem) {
        P.tablesInSmem = inSmem ? 1 : 0;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nChunks + kDeferWarps - 1) / kDeferWarps, ctx->smCount));
        cudaError_t e = launchDefer(ctx, P, smem, grid);
        if (e != cudaSuccess) return fail(ctx, EMCGPU_E_CUDA, "bulk step launch failed: %s", cudaGetErrorString(e));
        done += chunk;
        continue;
      }
    }
    // the streaming one-step kernels do not carry the grain clock: with a grain mechanism the general kernel runs
    const bool stream = chunk == 1 && !ctx->grainOn && !record;
    if (stream && ctx->optKernel != 1) {
      // preferred: the TMA pipeline (tables in shared memory if they fit beside >= kMinStages ring stages)
      size_t smem = 0;
      bool inSmem = true;
      int stages = ctx->optTablesGlobal ? 0 : tmaStages(ctx, true, &smem);
      if (!stages) {
        inSmem = false;
        stages = tmaStages(ctx, false, &smem);
      }
      if (stages) {
        if (ctx->optStages >= kMinStages && ctx->optStages < stages) {
          smem -= (size_t)(stages - ctx->optStages) * kTileBytes;
          stages = ctx->optStages;
        }
        P.tablesInSmem = inSmem ? 1 : 0;
        const int64_t nTiles = ctx->n / kTile;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nTiles, ctx->smCount));
        cudaError_t e = launchTma(ctx, P, smem, grid, stages);
        if (e != cudaSuccess) return fail(ctx, EMCGPU_E_CUDA, "bulk step launch failed: %s", cudaGetErrorString(e));
        done += chunk;
        continue;
      }
    }
    const int vec = ctx->optVec;
    const int queueWords = stream ? streamQueueWords(vec) : 0;
    bool inSmem = true;
    size_t smem = bulkSmemBytes(ctx, chunk * nV * 3, true, queueWords);
    if (smem > (size_t)ctx->maxSmemOptin) {
      inSmem = false;
      smem = bulkSmemBytes(ctx, chunk * nV * 3, false, queueWords);
      if (smem > (size_t)ctx->maxSmemOptin)
        return fail(ctx, EMCGPU_E_CAPACITY, "model does not fit in shared memory (%zu bytes)", smem);
    }
    P.tablesInSmem = inSmem ? 1 : 0;
    // persistent grid: as many CTAs as can be resident (2 per SM by launch
    // bounds, fewer if the tables are large), never more than the work needs
    int perSm = (int)std::min<size_t>(2, (size_t)(ctx->maxSmemPerSm) / (smem + 1024));
    if (perSm < 1) perSm = 1;
    const int64_t perCta = stream ? (int64_t)kStreamThreads * vec : (int64_t)kBulkThreads;
    const int blocksNeeded = (int)std::min<int64_t>((ctx->n + perCta - 1) / perCta, 1 << 30);
    const int grid = std::max(1, std::min(blocksNeeded, ctx->smCount * perSm));
    cudaError_t e;
    if (!stream)
      e = launchFused(ctx, P, smem, grid);
    else if (vec == 1)
      e = launchStreamVec<1>(ctx, P, smem, grid);
    else if (vec == 2)
      e = launchStreamVec<2>(ctx, P, smem, grid);
    else
      e = launchStreamVec<4>(ctx, P, smem, grid);
    if (e != cudaSuccess) return fail(ctx, EMCGPU_E_CUDA, "bulk step launch failed: %s", cudaGetErrorString(e));
    if (record) {
      const int b = recordIdx & 1;
      const size_t perStep = (size_t)ctx->n * ctx->velComponents;
      CUDA_TRY(ctx, cudaEventRecord(ctx->velDone[b], ctx->stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copyOut, ctx->velDone[b], 0));
      CUDA_TRY(ctx, cudaMemcpyAsync(ctx->velHost + (size_t)done * perStep, ctx->dVel[b].ptr, perStep * chunk * sizeof(double),
                                    cudaMemcpyDeviceToHost, ctx->copyOut));
      CUDA_TRY(ctx, cudaEventRecord(ctx->velCopied[b], ctx->copyOut));
      recordIdx++;
    }
    done += chunk;
  }
  if (record) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->copyOut));
  ctx->nextStep += nSteps;
  return EMCGPU_OK;
}
} // namespace
extern "C" {

int emcgpu_bulk_step_device(emcgpu_ctx *ctx, double dt, int nSteps, int stepsPerLaunch, double *obsDevice) {
  if (ctx) ctx->rewindValid = false;
  return bulkStepDevice(ctx, dt, nSteps, stepsPerLaunch, obsDevice, false);
}

int emcgpu_bulk_step_ahead(emcgpu_ctx *ctx, double dt, int nSteps, int stepsPerLaunch, double *obs) {
  if (int r = checkReady(ctx, true)) return r;
  if (nSteps < 1) return fail(ctx, EMCGPU_E_INVALID, "nSteps must be >= 1");
  if (ctx->rngMode != RNG_PHILOX) return fail(ctx, EMCGPU_E_INVALID, "emcgpu_bulk_step_ahead needs the Philox streams");
  if (int r = bind(ctx)) return r;
  const size_t bytes = (size_t)nSteps * ctx->hModel.nValleys * 3 * sizeof(double);
  CUDA_TRY(ctx, ctx->dObs.ensure(bytes));
  CUDA_TRY(ctx, ctx->dEnsembleAlt.ensure(ctx->dEnsemble.bytes));
  layoutStreams(ctx->dEnsembleAlt.ptr, ctx->n, ctx->dStreamAlt, &ctx->dPackedAlt);
  const int64_t stepBefore = ctx->nextStep;
  ctx->rewindValid = false;
  if (int r = bulkStepDevice(ctx, dt, nSteps, stepsPerLaunch, static_cast<double *>(ctx->dObs.ptr), true)) return r;
  ctx->rewindValid = true;
  ctx->rewindStep = stepBefore;
  if (obs) CUDA_TRY(ctx, cudaMemcpyAsync(obs, ctx->dObs.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return checkStatusWord(ctx); // synchronises
}

int emcgpu_bulk_rewind(emcgpu_ctx *ctx) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!ctx->rewindValid) return fail(ctx, EMCGPU_E_INVALID, "nothing to rewind: the last step call was not emcgpu_bulk_step_ahead");
  if (int r = bind(ctx)) return r;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  swapEnsembles(ctx);
  ctx->nextStep = ctx->rewindStep;
  ctx->rewindValid = false;
  return EMCGPU_OK;
}

int emcgpu_bulk_record_velocities(emcgpu_ctx *ctx, int components, double *host, int64_t capacitySteps) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (components != 0 && components != 1 && components != 3)
    return fail(ctx, EMCGPU_E_INVALID, "components must be 0 (off), 1 (v.E) or 3 (v)");
  if (components && (!host || capacitySteps < 1)) return fail(ctx, EMCGPU_E_INVALID, "the velocity record needs a host buffer");
  ctx->velComponents = components;
  ctx->velHost = components ? host : nullptr;
  ctx->velHostSteps = components ? capacitySteps : 0;
  return EMCGPU_OK;
}

int emcgpu_kernel_times(emcgpu_ctx *ctx, double *ms, int64_t *launches, int reset) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  for (auto &t : ctx->timed) {
    float f = 0;
    if (cudaEventElapsedTime(&f, t.a, t.b) == cudaSuccess && t.tag >= 0 && t.tag < 3) {
      ctx->timedMs[t.tag] += f;
      ctx->timedLaunches[t.tag]++;
    }
    cudaEventDestroy(t.a);
    cudaEventDestroy(t.b);
  }
  ctx->timed.clear();
  for (int k = 0; k < 3; k++) {
    if (ms) ms[k] = ctx->timedMs[k];
    if (launches) launches[k] = ctx->timedLaunches[k];
    if (reset) {
      ctx->timedMs[k] = 0;
      ctx->timedLaunches[k] = 0;
    }
  }
  return EMCGPU_OK;
}

int emcgpu_bulk_step(emcgpu_ctx *ctx, double dt, int nSteps, int stepsPerLaunch, double *obs) {
  if (int r = checkReady(ctx, true)) return r;
  if (nSteps < 1) return fail(ctx, EMCGPU_E_INVALID, "nSteps must be >= 1");
  if (int r = bind(ctx)) return r;
  const size_t bytes = (size_t)nSteps * ctx->hModel.nValleys * 3 * sizeof(double);
  CUDA_TRY(ctx, ctx->dObs.ensure(bytes));
  if (int r = emcgpu_bulk_step_device(ctx, dt, nSteps, stepsPerLaunch, static_cast<double *>(ctx->dObs.ptr)))
    return r;
  if (obs) CUDA_TRY(ctx, cudaMemcpyAsync(obs, ctx->dObs.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return checkStatusWord(ctx); // synchronises
}

// Host-resident ensemble advanced slice by slice: while slice i runs its nSteps time steps, slice i+1 is on its way
// to the device and slice i-1 on its way back (particles of a bulk run are independent: the Philox stream of a
// particle is keyed by its global id, the observables are sums). The resident ensemble of ctx is left untouched.
int emcgpu_bulk_run_host(emcgpu_ctx *ctx, int64_t n, double *const *soa, uint32_t *packed, int64_t particleIdBase,
                         double dt, int nSteps, int stepsPerLaunch, int64_t sliceParticles, double *obs) {
  if (int r = checkReady(ctx, false)) return r;
  if (n < 0 || (n > 0 && (!soa || !packed)) || !(dt > 0) || nSteps < 1)
    return fail(ctx, EMCGPU_E_INVALID, "bad arguments of emcgpu_bulk_run_host");
  if (ctx->rngMode != RNG_PHILOX) return fail(ctx, EMCGPU_E_INVALID, "emcgpu_bulk_run_host needs the Philox streams (replay streams belong to a resident ensemble)");
  if (ctx->grainOn) return fail(ctx, EMCGPU_E_INVALID, "emcgpu_bulk_run_host does not carry grain clocks");
  if (int r = bind(ctx)) return r;
  for (int s = 0; s < EMCGPU_N_STREAMS && n > 0; s++)
    if (!soa[s]) return fail(ctx, EMCGPU_E_INVALID, "stream %d is NULL", s);
  const int nV = ctx->hModel.nValleys;
  const size_t obsBytes = (size_t)nSteps * nV * 3 * sizeof(double);
  CUDA_TRY(ctx, ctx->dObs.ensure(obsBytes));
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->dObs.ptr, 0, obsBytes, ctx->stream));
  if (n > 0) {
    // slice: a multiple of what one pass of the deferred-event kernel takes (all SMs x warps x chunk), about n/8 -- n/16 for
    // short runs, which are bound by the copies and gain from a shorter fill and drain of the pipeline
    // (measured at n = 1e8 with the flight / event pair, profiles/bench/r2_y_host_run_sweep.json: 1000 steps -- 4 slices 452 ms,
    // 6: 436, 8: 440, 12: 447, 16: 448, 24: 463, 32: 527; 20 steps -- 4: 172, 6: 166, 8: 158, 12: 155, 16: 153, 24: 154, 32: 154)
    const int64_t quantum = (int64_t)ctx->smCount * kDeferWarps * kDeferChunk;
    const int64_t parts = nSteps < 128 ? 16 : 8;
    int64_t slice = sliceParticles > 0 ? sliceParticles : std::max<int64_t>((n / parts + quantum - 1) / quantum * quantum, 16 * quantum);
    slice = std::min(slice, n);
    if (slice >= (int64_t)1 << 32) return fail(ctx, EMCGPU_E_CAPACITY, "at most 2^32-1 particles per slice");
    const int nBuf = 3;
    const size_t strideD = ((size_t)slice * sizeof(double) + 255) & ~size_t(255);
    const size_t strideP = ((size_t)slice * sizeof(uint32_t) + 255) & ~size_t(255);
    const size_t bufBytes = strideD * EMCGPU_N_STREAMS + strideP;
    CUDA_TRY(ctx, ctx->dSlices.ensure(bufBytes * nBuf));
    if (!ctx->copyIn) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copyIn, cudaStreamNonBlocking));
    if (!ctx->copyOut) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copyOut, cudaStreamNonBlocking));
    // the nine events of the slice ring live with the context (created once, destroyed by emcgpu_destroy)
    for (int b = 0; b < 3 * nBuf; b++)
      if (!ctx->sliceEvents[b]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->sliceEvents[b], cudaEventDisableTiming));
    cudaEvent_t *const evIn = ctx->sliceEvents, *const evRun = ctx->sliceEvents + nBuf, *const evOut = ctx->sliceEvents + 2 * nBuf;
    // the step routine works on whatever ensemble the context points at
    double *savedStream[EMCGPU_N_STREAMS];
    for (int s = 0; s < EMCGPU_N_STREAMS; s++) savedStream[s] = ctx->dStream[s];
    uint32_t *savedPacked = ctx->dPacked;
    const int64_t savedN = ctx->n, savedBase = ctx->idBase, step0 = ctx->nextStep;
    const bool savedAcc = ctx->obsAccumulate;
    ctx->obsAccumulate = true;
    int rc = EMCGPU_OK;
    cudaError_t ce = cudaSuccess;
    const int64_t nSlices = (n + slice - 1) / slice;
    for (int64_t i = 0; i < nSlices && rc == EMCGPU_OK && ce == cudaSuccess; i++) {
      const int b = (int)(i % nBuf);
      const int64_t first = i * slice, m = std::min(slice, n - first);
      unsigned char *base = static_cast<unsigned char *>(ctx->dSlices.ptr) + bufBytes * b;
      if (i >= nBuf) ce = cudaStreamWaitEvent(ctx->copyIn, evOut[b], 0); // the slice that used this buffer is back on the host
      for (int s = 0; s < EMCGPU_N_STREAMS && ce == cudaSuccess; s++)
        ce = cudaMemcpyAsync(base + strideD * s, soa[s] + first, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, ctx->copyIn);
      if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(base + strideD * EMCGPU_N_STREAMS, packed + first, (size_t)m * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->copyIn);
      if (ce == cudaSuccess) ce = cudaEventRecord(evIn[b], ctx->copyIn);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->stream, evIn[b], 0);
      if (ce != cudaSuccess) break;
      for (int s = 0; s < EMCGPU_N_STREAMS; s++) ctx->dStream[s] = reinterpret_cast<double *>(base + strideD * s);
      ctx->dPacked = reinterpret_cast<uint32_t *>(base + strideD * EMCGPU_N_STREAMS);
      ctx->n = m;
      ctx->idBase = particleIdBase + first;
      ctx->nextStep = step0;
      rc = emcgpu_bulk_step_device(ctx, dt, nSteps, stepsPerLaunch, static_cast<double *>(ctx->dObs.ptr));
      if (rc != EMCGPU_OK) break;
      ce = cudaEventRecord(evRun[b], ctx->stream);
      if (ce == cudaSuccess) ce = cudaStreamWaitEvent(ctx->copyOut, evRun[b], 0);
      for (int s = 0; s < EMCGPU_N_STREAMS && ce == cudaSuccess; s++)
        ce = cudaMemcpyAsync(soa[s] + first, base + strideD * s, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, ctx->copyOut);
      if (ce == cudaSuccess)
        ce = cudaMemcpyAsync(packed + first, base + strideD * EMCGPU_N_STREAMS, (size_t)m * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->copyOut);
      if (ce == cudaSuccess) ce = cudaEventRecord(evOut[b], ctx->copyOut);
    }
    cudaStreamSynchronize(ctx->copyIn);
    cudaStreamSynchronize(ctx->stream);
    cudaError_t ce2 = cudaStreamSynchronize(ctx->copyOut);
    for (int s = 0; s < EMCGPU_N_STREAMS; s++) ctx->dStream[s] = savedStream[s];
    ctx->dPacked = savedPacked;
    ctx->n = savedN;
    ctx->idBase = savedBase;
    ctx->obsAccumulate = savedAcc;
    ctx->nextStep = step0 + (rc == EMCGPU_OK ? nSteps : 0);
    if (rc != EMCGPU_OK) return rc;
    if (ce == cudaSuccess) ce = ce2;
    if (ce != cudaSuccess) return fail(ctx, EMCGPU_E_CUDA, "emcgpu_bulk_run_host: %s", cudaGetErrorString(ce));
  } else {
    ctx->nextStep += nSteps;
  }
  if (obs) CUDA_TRY(ctx, cudaMemcpyAsync(obs, ctx->dObs.ptr, obsBytes, cudaMemcpyDeviceToHost, ctx->stream));
  return checkStatusWord(ctx); // synchronises
}

int emcgpu_bulk_observables(emcgpu_ctx *ctx, double *obs) {
  if (int r = checkReady(ctx, true)) return r;
  if (!obs) return fail(ctx, EMCGPU_E_INVALID, "obs is NULL");
  if (int r = bind(ctx)) return r;
  const size_t bytes = (size_t)ctx->hModel.nValleys * 3 * sizeof(double);
  CUDA_TRY(ctx, ctx->dObs.ensure(bytes));
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->dObs.ptr, 0, bytes, ctx->stream));
  BulkParams P;
  fillBulkParams(ctx, P);
  P.obs = static_cast<double *>(ctx->dObs.ptr);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n + kBulkThreads - 1) / kBulkThreads,
                                                                (int64_t)ctx->smCount * 8));
  if (ctx->mathMode == EMCGPU_MATH_EXACT)
    bulkObservablesKernel<true><<<grid, kBulkThreads, 0, ctx->stream>>>(P);
  else
    bulkObservablesKernel<false><<<grid, kBulkThreads, 0, ctx->stream>>>(P);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  CUDA_TRY(ctx, cudaMemcpyAsync(obs, ctx->dObs.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_event_log_enable(emcgpu_ctx *ctx, int64_t capacity) {
  if (!ctx || capacity < 0) return EMCGPU_E_INVALID;
  if (int r = bind(ctx)) return r;
  if (capacity > 0) CUDA_TRY(ctx, ctx->dEvents.ensure((size_t)capacity * 4 * sizeof(long long)));
  ctx->evCap = capacity;
  CUDA_TRY(ctx, cudaMemsetAsync(ctx->dEvCount.ptr, 0, sizeof(unsigned long long), ctx->stream));
  return EMCGPU_OK;
}

int64_t emcgpu_event_log_read(emcgpu_ctx *ctx, int64_t *out, int64_t capacity) {
  if (!ctx) return -1;
  if (bind(ctx)) return -1;
  unsigned long long count = 0;
  if (cudaMemcpyAsync(&count, ctx->dEvCount.ptr, sizeof count, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    fail(ctx, EMCGPU_E_CUDA, "cannot read the event counter");
    return -1;
  }
  const int64_t have = std::min<int64_t>((int64_t)count, std::min(capacity, ctx->evCap));
  if (have > 0 && out) {
    if (cudaMemcpy(out, ctx->dEvents.ptr, (size_t)have * 4 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) {
      fail(ctx, EMCGPU_E_CUDA, "cannot read the event log");
      return -1;
    }
  }
  cudaMemsetAsync(ctx->dEvCount.ptr, 0, sizeof(unsigned long long), ctx->stream);
  return (int64_t)count;
}

} // extern "C"
