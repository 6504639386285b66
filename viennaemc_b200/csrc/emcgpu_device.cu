// C-ABI implementation of the device-run path (include/emcgpu.h, emcgpu_device_*):
// grids, Poisson, field, charge assignment, particle step with boundaries, contacts.
#include <algorithm>
#include <cmath>

#include "emc_device_run.cuh"
#include "emcgpu_internal.cuh"

using namespace emc;

namespace emc {

constexpr int kRunChunk = 128; // steps of the device-run loop between two host synchronisations

struct DeviceRunState {
  DevGeometry geo{};
  int dim = 2;
  double charge = 0, nrCarriers = 1;
  DeviceBuffer dRegion, dFace, dDoping, dDopingNorm, dCellKind, dSorHistory;
  DeviceBuffer grid[EMCGPU_N_GRIDS];
  DeviceBuffer dCtl, dFlag, dChunkCount, dListParticle, dListCell, dCellCount, dInjectCount, dSweeps, dReplay, dHits;
  bool chained = false; // inside emcgpu_device_run* on one GPU: kernels of a step are launched as programmatic dependents (chainedLaunch)
  DeviceBuffer dCounters, dSweepsPerStep; // per-step outputs of a chunk of the step loop
  DeviceBuffer altEnsemble, altCursor, altGrain; // second ensemble buffer for the order-preserving compaction
  int64_t reserve = 0;
  int rank = 0, world = 1; // emcgpu_device_set_sharding
  emcgpu_allreduce_fn allreduce = nullptr;
  void *allreduceUser = nullptr;
  DeviceBuffer dShare; // [world][cells] reservoir particles per rank and cell
  int64_t maxInject = -1; // upper bound of the particles the contacts inject in one step
  int64_t runSteps = 0; // steps done by emcgpu_device_run* since configure (frozen-field sub-cycling)
};

void releaseDeviceRun(emcgpu_ctx *ctx) {
  if (!ctx->run) return;
  DeviceRunState *r = ctx->run;
  for (DeviceBuffer *b : {&r->dRegion, &r->dFace, &r->dDoping, &r->dDopingNorm, &r->dCellKind, &r->dSorHistory, &r->dCtl, &r->dFlag, &r->dHits, &r->dChunkCount, &r->dListParticle,
                          &r->dListCell, &r->dCellCount, &r->dInjectCount, &r->dSweeps, &r->dReplay, &r->dCounters,
                          &r->dSweepsPerStep, &r->altEnsemble, &r->altCursor, &r->altGrain, &r->dShare})
    b->release();
  for (auto &g : r->grid) g.release();
  delete r;
  ctx->run = nullptr;
}

} // namespace emc

namespace {

#define fail emc::failWith

int needRun(emcgpu_ctx *ctx) {
  if (!ctx) return EMCGPU_E_INVALID;
  if (!ctx->run) return fail(ctx, EMCGPU_E_INVALID, "emcgpu_device_configure has not been called");
  return emc::bindDevice(ctx);
}

int needModel(emcgpu_ctx *ctx) {
  if (!ctx->haveValleys || !ctx->haveTables)
    return fail(ctx, EMCGPU_E_INVALID, "set valleys and tables before running the device path");
  if (ctx->grainOn && !ctx->grainClockSet)
    return fail(ctx, EMCGPU_E_INVALID, "a grain mechanism is set but the grain clocks were not uploaded (emcgpu_set_grain_clock)");
  return EMCGPU_OK;
}

size_t streamStrideD(int64_t cap) { return ((size_t)cap * sizeof(double) + 255) & ~size_t(255); }
size_t streamStrideP(int64_t cap) { return ((size_t)cap * sizeof(uint32_t) + 255) & ~size_t(255); }

EnsemblePtrs ptrsOf(void *base, int64_t cap, uint32_t *cursor, double *grain = nullptr) {
  EnsemblePtrs p;
  p.grain = grain;
  unsigned char *b = static_cast<unsigned char *>(base);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) p.stream[s] = reinterpret_cast<double *>(b + streamStrideD(cap) * s);
  p.packed = reinterpret_cast<uint32_t *>(b + streamStrideD(cap) * EMCGPU_N_STREAMS);
  p.cursor = cursor;
  return p;
}

// make sure the ensemble allocation (and its twin) can hold `cap` particles, keeping the current content
int growEnsemble(emcgpu_ctx *ctx, int64_t cap) {
  DeviceRunState *r = ctx->run;
  const bool grainReady = !ctx->grainOn || (ctx->dGrain.bytes >= (size_t)ctx->capacity * sizeof(double) &&
                                            r->altGrain.bytes >= (size_t)ctx->capacity * sizeof(double));
  if (cap <= ctx->capacity && r->altEnsemble.bytes >= ctx->dEnsemble.bytes && grainReady) return EMCGPU_OK;
  cap = std::max<int64_t>(cap, ctx->capacity);
  const size_t bytes = streamStrideD(cap) * EMCGPU_N_STREAMS + streamStrideP(cap);
  if (cap > ctx->capacity) {
    DeviceBuffer bigger;
    CUDA_TRY(ctx, bigger.ensure(bytes));
    EnsemblePtrs dst = ptrsOf(bigger.ptr, cap, nullptr);
    for (int s = 0; s < EMCGPU_N_STREAMS; s++)
      if (ctx->n)
        CUDA_TRY(ctx, cudaMemcpyAsync(dst.stream[s], ctx->dStream[s], ctx->n * sizeof(double), cudaMemcpyDeviceToDevice,
                                      ctx->stream));
    if (ctx->n)
      CUDA_TRY(ctx, cudaMemcpyAsync(dst.packed, ctx->dPacked, ctx->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->dEnsemble.release();
    ctx->dEnsemble = bigger;
    for (int s = 0; s < EMCGPU_N_STREAMS; s++) ctx->dStream[s] = dst.stream[s];
    ctx->dPacked = dst.packed;
    ctx->capacity = cap;
  }
  CUDA_TRY(ctx, r->altEnsemble.ensure(ctx->dEnsemble.bytes));
  if (ctx->grainOn) { // grain clocks: one per particle slot, content kept when the allocation grows
    const size_t need = (size_t)ctx->capacity * sizeof(double);
    if (ctx->dGrain.bytes < need) {
      DeviceBuffer bigger;
      CUDA_TRY(ctx, bigger.ensure(need));
      if (ctx->dGrain.ptr && ctx->n)
        CUDA_TRY(ctx, cudaMemcpyAsync(bigger.ptr, ctx->dGrain.ptr, std::min(ctx->dGrain.bytes, (size_t)ctx->n * sizeof(double)),
                                      cudaMemcpyDeviceToDevice, ctx->stream));
      CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
      ctx->dGrain.release();
      ctx->dGrain = bigger;
    }
    CUDA_TRY(ctx, r->altGrain.ensure(need));
  }
  return EMCGPU_OK;
}

void fillParams(emcgpu_ctx *ctx, BulkParams &P) {
  memset(&P, 0, sizeof P);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) P.stream[s] = ctx->dStream[s];
  P.packed = ctx->dPacked;
  P.n = ctx->n;
  P.idBase = ctx->idBase;
  P.model = ctx->dModel.as<const DevModel>();
  P.tables = ctx->dTables.as<const double>();
  P.mechs = ctx->dMechs.as<const DevMech>();
  P.nMechTotal = (int32_t)ctx->hMechs.size();
  P.seed = ctx->seed;
  P.draws = ctx->dDraws.as<const uint64_t>();
  P.offsets = ctx->dOffsets.as<const int64_t>();
  P.cursor = ctx->dCursor.as<uint32_t>();
  P.events = ctx->dEvents.as<long long>();
  P.evCap = ctx->evCap;
  P.evCount = ctx->dEvCount.as<unsigned long long>();
  P.status = ctx->dStatus.as<int>();
  emc::fillBathView(ctx, P.baths);
  emc::fillGrain(ctx, P);
}

int readStatus(emcgpu_ctx *ctx) {
  int status = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&status, ctx->dStatus.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (status != 0) {
    cudaMemsetAsync(ctx->dStatus.ptr, 0, sizeof(int), ctx->stream);
    if (status == EMCGPU_E_REPLAY_EXHAUSTED)
      return fail(ctx, status, "replay stream exhausted: more draws were needed than were recorded");
    if (status == EMCGPU_E_CAPACITY)
      return fail(ctx, status, "the contacts injected more particles than the ensemble allocation holds: "
                               "call emcgpu_device_reserve with a larger capacity");
    return fail(ctx, status, "device reported status %d", status);
  }
  return EMCGPU_OK;
}

int gridBlocks(int cells) { return (cells + 255) / 256; }

// ---- control block: uploaded at the start of every API call, read back at its end ------------------------
int pushCtl(emcgpu_ctx *ctx, int avgFromSlot) {
  DeviceRunState *r = ctx->run;
  if (int rc = growEnsemble(ctx, std::max<int64_t>(ctx->capacity, 1))) return rc;
  const size_t cap = (size_t)ctx->capacity;
  CUDA_TRY(ctx, r->dFlag.ensure(cap * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dListParticle.ensure(cap * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dListCell.ensure(cap * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dChunkCount.ensure((cap / kChunk + 2) * sizeof(int32_t)));
  RunCtl h;
  memset(&h, 0, sizeof h);
  h.n = (int32_t)ctx->n;
  h.nKept = (int32_t)ctx->n;
  h.capacity = (int32_t)ctx->capacity;
  h.avgFromSlot = avgFromSlot;
  h.poissonInterval = ctx->optPoissonInterval;
  h.step = ctx->nextStep;
  h.runSteps = r->runSteps;
  CUDA_TRY(ctx, cudaMemcpyAsync(r->dCtl.ptr, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
  return EMCGPU_OK;
}

// waits for the stream; ensemble size, step index and the per-contact counters of a single step come back
int pullCtl(emcgpu_ctx *ctx, RunCtl *out = nullptr) {
  DeviceRunState *r = ctx->run;
  RunCtl h;
  CUDA_TRY(ctx, cudaMemcpyAsync(&h, r->dCtl.ptr, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->n = h.n;
  ctx->nextStep = h.step;
  r->runSteps = h.runSteps;
  if (out) *out = h;
  return EMCGPU_OK;
}

// Launch on the context's stream; inside the step loop of a run (DeviceRunState::chained) as a PROGRAMMATIC DEPENDENT of the kernel
// ahead of it in the stream: every kernel of the chain releases its dependents first thing (griddepcontrol.launch_dependents) and
// waits for the end of its predecessor before it touches global memory (griddepcontrol.wait), so the CTAs of the next kernel are
// resident -- the step kernel even with its tables staged -- when the predecessor ends, instead of paying a launch after it.
template <typename... P, typename... A>
cudaError_t chainedLaunch(emcgpu_ctx *ctx, void (*kernel)(P...), int grid, int block, size_t smem, A &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  const bool chained = ctx->run && ctx->run->chained;
  cfg.attrs = chained ? attr : nullptr;
  cfg.numAttrs = chained ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

int particleGrid(emcgpu_ctx *ctx, int threads, int perSm) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((ctx->capacity + threads - 1) / threads, (int64_t)perSm * ctx->smCount));
}

// ---- the pieces of one EMC step: all asynchronous on the context's stream, sizes read from the control block ----
int doPoisson(emcgpu_ctx *ctx, bool equilibrium, double accuracyVolt, double omega, bool resetBC, bool inLoop, bool withField) {
  DeviceRunState *r = ctx->run;
  const DevGeometry &G = r->geo;
  double *pot = r->grid[EMCGPU_GRID_POTENTIAL].as<double>();
  if (resetBC) {
    sorResetBcKernel<<<gridBlocks(G.cells), 256, 0, ctx->stream>>>(G, pot, equilibrium ? 0 : 1);
    ctx->launches++;
  }
  SorParams S;
  S.pot = pot;
  S.conc = equilibrium ? nullptr : r->grid[EMCGPU_GRID_CONCENTRATION].as<const double>();
  S.accuracy = accuracyVolt / G.thermalVoltage;
  S.omega = omega;
  S.maxSweeps = 1000000;
  const size_t smem = (size_t)G.cells * sizeof(double);
  S.potInSmem = smem <= (size_t)ctx->maxSmemOptin - 2048 ? 1 : 0;
  S.sweepsOut = r->dSweeps.as<int32_t>();
  CUDA_TRY(ctx, r->dSorHistory.ensure((size_t)kSorRing * G.cells * sizeof(double)));
  S.history = r->dSorHistory.as<double>();
  S.efield = withField ? r->grid[EMCGPU_GRID_EFIELD_X].as<double>() : nullptr;
  S.ctl = inLoop ? r->dCtl.as<RunCtl>() : nullptr;
  S.sweepsPerStep = inLoop ? r->dSweepsPerStep.as<int32_t>() : nullptr;
  // one thread pair per grid row when a wave of >= 12 sweeps fits the CTA that way (and the hand-over records fit
  // next to the potential), hyperplane loop otherwise
  const int nYZ = G.extent[1] * (G.dim > 2 ? G.extent[2] : 1);
  auto launch = [&](auto kernel, size_t handover) {
    size_t bytes = (S.potInSmem ? smem : 0) + handover;
    if (bytes > (size_t)ctx->maxSmemOptin - 2048) { // potential stays in global memory / L1
      S.potInSmem = 0;
      bytes = handover;
    }
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    kernel<<<1, kSorThreads, bytes, ctx->stream>>>(G, S);
    return cudaGetLastError();
  };
  const int ipt = 16 * nYZ <= kSorPairs ? 1 : 16 * nYZ <= 2 * kSorPairs ? 2 : 4;
  if (ctx->optSorOrder == 1) {
    // red-black: a cluster of 8 CTAs with the potential banded over their shared memories when the device offers it;
    // the E field of the result is a separate (tiny) kernel then
    const int nRowsRb = G.extent[1] * (G.dim > 2 ? G.extent[2] : 1);
    const int rowsPerCta = (nRowsRb + kSorClusterSize - 1) / kSorClusterSize;
    const size_t bandBytes = (size_t)rowsPerCta * G.extent[0] * (3 * sizeof(double) + 1) + 16;
    // two dimensions, every thread at most one cell per colour: the fast form of the cluster kernel, on 16 CTAs of 512
    // threads where the device places such a cluster (decided once per grid), else on 8 CTAs of 1024; option
    // sor_kernel = 2 keeps the general cluster kernel
    const int halfX = (G.extent[0] + 1) / 2;
    auto fastLaunch = [&](int clusterSize, int threads, bool probeOnly) -> cudaError_t {
      const int rows = (nRowsRb + clusterSize - 1) / clusterSize;
      const size_t bytes = sorClusterFastSmemBytes(rows, G.extent[0], threads);
      if (G.dim != 2 || G.extent[0] < 2 || nRowsRb < clusterSize || halfX > threads || rows > threads / halfX ||
          bytes > (size_t)ctx->maxSmemOptin - 4096)
        return cudaErrorInvalidConfiguration;
      const void *kernel = clusterSize == kSorClusterSizeWide ? (const void *)sorRedBlackClusterFastKernel<kSorClusterThreadsWide>
                                                              : (const void *)sorRedBlackClusterFastKernel<kSorClusterThreads>;
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e != cudaSuccess) return e;
      if (clusterSize > 8 && (e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) return e;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(clusterSize);
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = bytes;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = clusterSize;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; // behind the charge assignment of the last step
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = !probeOnly && inLoop && r->chained ? 2 : 1;
      if (probeOnly) {
        int nClusters = 0;
        e = cudaOccupancyMaxActiveClusters(&nClusters, kernel, &cfg);
        return e != cudaSuccess ? e : nClusters >= 1 ? cudaSuccess : cudaErrorInvalidConfiguration;
      }
      void *args[] = {(void *)&G, (void *)&S}; // the field follows the potential inside the launch
      return cudaLaunchKernelExC(&cfg, kernel, args);
    };
    const long long fastKey = (long long)G.extent[0] * 1000003LL + nRowsRb * 4LL + G.dim;
    if (ctx->sorFastKey != fastKey) { // 0: general cluster kernel, 1: fast on 8 x 1024, 2: fast on 16 x 512
      ctx->sorFastKey = fastKey;
      ctx->sorFast = fastLaunch(kSorClusterSizeWide, kSorClusterThreadsWide, true) == cudaSuccess ? 2
                     : fastLaunch(kSorClusterSize, kSorClusterThreads, true) == cudaSuccess    ? 1
                                                                                               : 0;
      cudaGetLastError();
    }
    if ((ctx->optSorKernel == 0 || ctx->optSorKernel == 3) && ctx->sorFast > 0) {
      CUDA_TRY(ctx, ctx->sorFast == 2 && ctx->optSorKernel == 0 ? fastLaunch(kSorClusterSizeWide, kSorClusterThreadsWide, false)
                                      : fastLaunch(kSorClusterSize, kSorClusterThreads, false));
    } else if (ctx->optSorKernel != 1 && nRowsRb >= kSorClusterSize && bandBytes <= (size_t)ctx->maxSmemOptin - 4096 &&
        (G.extent[0] + 1) / 2 <= kSorClusterThreads) {
      auto kernel = G.dim == 2 ? sorRedBlackClusterKernel<2> : sorRedBlackClusterKernel<3>;
      CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bandBytes));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(kSorClusterSize);
      cfg.blockDim = dim3(kSorClusterThreads);
      cfg.dynamicSmemBytes = bandBytes;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = kSorClusterSize;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      SorParams S2 = S;
      S2.efield = nullptr;
      CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kernel, G, S2));
      if (S.efield) {
        // the field follows the potential; inside the step loop it must skip with the solve (frozen-field sub-cycling)
        efieldAfterSolveKernel<<<gridBlocks(G.cells), 256, 0, ctx->stream>>>(G, S.pot, S.efield, S.ctl);
        ctx->launches++;
      }
    } else {
      CUDA_TRY(ctx, G.dim == 2 ? launch(sorRedBlackKernel<2>, 0) : launch(sorRedBlackKernel<3>, 0));
    }
  }
  else if (ctx->optSorKernel == 1 || 12 * nYZ > 4 * kSorPairs)
    CUDA_TRY(ctx, launch(sorPlanesKernel, 0));
  else if (ipt == 1)
    CUDA_TRY(ctx, G.dim == 2 ? launch(sorRowsKernel<1, 2>, sorRowsHandoverBytes<1>()) : launch(sorRowsKernel<1, 3>, sorRowsHandoverBytes<1>()));
  else if (ipt == 2)
    CUDA_TRY(ctx, G.dim == 2 ? launch(sorRowsKernel<2, 2>, sorRowsHandoverBytes<2>()) : launch(sorRowsKernel<2, 3>, sorRowsHandoverBytes<2>()));
  else
    CUDA_TRY(ctx, G.dim == 2 ? launch(sorRowsKernel<4, 2>, sorRowsHandoverBytes<4>()) : launch(sorRowsKernel<4, 3>, sorRowsHandoverBytes<4>()));
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return EMCGPU_OK;
}

int doEfield(emcgpu_ctx *ctx) {
  DeviceRunState *r = ctx->run;
  efieldKernel<<<gridBlocks(r->geo.cells), 256, 0, ctx->stream>>>(r->geo, r->grid[EMCGPU_GRID_POTENTIAL].as<const double>(),
                                                                  r->grid[EMCGPU_GRID_EFIELD_X].as<double>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return EMCGPU_OK;
}

// charge assignment; withConc: the last block also forms the concentration; closeStep: ... and ends the step
int doAssign(emcgpu_ctx *ctx, bool withConc, bool closeStep) {
  DeviceRunState *r = ctx->run;
  const DevGeometry &G = r->geo;
  double *count = r->grid[EMCGPU_GRID_COUNT].as<double>();
  AssignParams A;
  A.x = ctx->dStream[EMCGPU_X];
  A.y = ctx->dStream[EMCGPU_Y];
  A.z = ctx->dStream[EMCGPU_Z];
  A.nrCarriers = r->nrCarriers;
  A.count = count;
  A.conc = withConc ? r->grid[EMCGPU_GRID_CONCENTRATION].as<double>() : nullptr;
  A.pot = r->grid[EMCGPU_GRID_POTENTIAL].as<const double>();
  A.sumPot = r->grid[EMCGPU_GRID_SUM_POTENTIAL].as<double>();
  A.sumConc = r->grid[EMCGPU_GRID_SUM_CONCENTRATION].as<double>();
  A.ctl = r->dCtl.as<RunCtl>();
  A.counters = r->dCounters.as<int32_t>();
  A.closeStep = closeStep ? 1 : 0;
  const bool sharded = r->world > 1 && withConc;
  if (sharded) { // deposit only; the counts of all ranks are summed before the concentration is formed
    A.conc = nullptr;
    A.closeStep = 0;
  }
  size_t smem = (size_t)G.cells * sizeof(double);
  A.useSmem = smem <= 96 * 1024 ? 1 : 0;
  // NEC / NEC-VWD deposits of an integer-valued nrCarriers: integer hits per mesh cell (share x hits is exact)
  const bool necHits = (G.pmScheme == PM_NEC || G.pmScheme == PM_NEC_VWD) && r->nrCarriers == std::floor(r->nrCarriers) &&
                       r->nrCarriers > 0 && r->nrCarriers < 16777216.0 && ctx->capacity < (int64_t(1) << 26) &&
                       !ctx->optAssignFp64;
  A.hits = nullptr;
  if (necHits) {
    A.useSmem = 2;
    smem = 0;
    CUDA_TRY(ctx, r->dHits.ensure((size_t)G.cells * sizeof(int32_t)));
    A.hits = r->dHits.as<int32_t>();
    CUDA_TRY(ctx, cudaMemsetAsync(A.hits, 0, (size_t)G.cells * sizeof(int32_t), ctx->stream));
  } else {
    CUDA_TRY(ctx, cudaMemsetAsync(count, 0, (size_t)G.cells * sizeof(double), ctx->stream));
  }
  if (A.useSmem == 1) CUDA_TRY(ctx, cudaFuncSetAttribute(ngpAssignKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (sharded && closeStep) { // the ensemble of this step is the survivors plus the injected particles
    A.closeStep = 2;
  }
  ngpAssignKernel<<<particleGrid(ctx, kAssignThreads, 1), kAssignThreads, A.useSmem == 1 ? smem : 0, ctx->stream>>>(G, A);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  if (sharded) {
    r->allreduce(r->allreduceUser, count, G.cells, ctx->stream);
    A.conc = r->grid[EMCGPU_GRID_CONCENTRATION].as<double>();
    if (closeStep) {
      concentrationCloseKernel<<<1, 1024, 0, ctx->stream>>>(G, A);
    } else {
      concentrationKernel<<<gridBlocks(G.cells), 256, 0, ctx->stream>>>(G, count, A.conc);
    }
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
  } else if (r->world > 1) {
    // deposit alone (emcgpu_device_assign, e.g. the equilibrium charge before the first step): the grid every rank sees
    // is the sum over the ranks all the same -- a replicated Poisson solve must never start from a local charge
    r->allreduce(r->allreduceUser, count, G.cells, ctx->stream);
  }
  return EMCGPU_OK;
}

int doConcentration(emcgpu_ctx *ctx) {
  DeviceRunState *r = ctx->run;
  concentrationKernel<<<gridBlocks(r->geo.cells), 256, 0, ctx->stream>>>(r->geo, r->grid[EMCGPU_GRID_COUNT].as<const double>(),
                                                                         r->grid[EMCGPU_GRID_CONCENTRATION].as<double>());
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  return EMCGPU_OK;
}

// early: programmatic dependent launch -- the CTAs may become resident and stage model and tables while the kernel ahead in the
// stream (the cluster solver, which releases its dependents at its start) still runs; the kernel waits for that kernel's end
// (griddepcontrol.wait) before it touches anything a step produces
template <bool EXACT, int MODE>
cudaError_t launchStepDim(emcgpu_ctx *ctx, const DeviceStepParams &D, size_t smem, int grid, bool early) {
  const DevGeometry &G = ctx->run->geo;
  auto go = [&](auto kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (early) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(kBulkThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = ctx->stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, kernel, G, D);
      if (e != cudaSuccess) return e;
    } else {
      kernel<<<grid, kBulkThreads, smem, ctx->stream>>>(G, D);
    }
    ctx->launches++;
    return cudaGetLastError();
  };
  return ctx->run->dim == 2 ? go(deviceStepKernel<EXACT, MODE, 2>) : go(deviceStepKernel<EXACT, MODE, 3>);
}

// particle step: ensemble in place, flags for the contact handling / compaction, removedPerContact in the control block
// countReservoir: the step kernel also counts the reservoir particles per chunk (the contact handling that follows in the
// same step then starts at its list kernel)
int doStep(emcgpu_ctx *ctx, double dt, bool countReservoir = false, bool early = false) {
  DeviceRunState *r = ctx->run;
  DeviceStepParams D;
  fillParams(ctx, D.P);
  D.P.dt = dt;
  D.P.nSteps = 1;
  D.e = r->grid[EMCGPU_GRID_EFIELD_X].as<const double>();
  D.charge = r->charge;
  D.flag = r->dFlag.as<int32_t>();
  D.ctl = r->dCtl.as<RunCtl>();
  D.chunkCount = countReservoir ? r->dChunkCount.as<int32_t>() : nullptr;
  bool inSmem = true;
  size_t smem = BulkSmem(0, ctx->hModel.nValleys, (int)ctx->hMechs.size(), ctx->hModel.tableDoubles, true, 0).total;
  if (smem > (size_t)ctx->maxSmemOptin / 2) { // two CTAs per SM
    inSmem = false;
    smem = BulkSmem(0, ctx->hModel.nValleys, (int)ctx->hMechs.size(), ctx->hModel.tableDoubles, false, 0).total;
  }
  D.P.tablesInSmem = inSmem ? 1 : 0;
  const int grid = particleGrid(ctx, kBulkThreads, 2);
  const bool exact = ctx->mathMode == EMCGPU_MATH_EXACT;
  cudaError_t e;
  if (ctx->rngMode == RNG_PHILOX)
    e = exact ? launchStepDim<true, RNG_PHILOX>(ctx, D, smem, grid, early) : launchStepDim<false, RNG_PHILOX>(ctx, D, smem, grid, early);
  else
    e = exact ? launchStepDim<true, RNG_REPLAY>(ctx, D, smem, grid, early) : launchStepDim<false, RNG_REPLAY>(ctx, D, smem, grid, early);
  if (e != cudaSuccess) return fail(ctx, EMCGPU_E_CUDA, "device step launch failed: %s", cudaGetErrorString(e));
  return EMCGPU_OK;
}

// drop the particles flagged kGone, keeping the order of the others (into the twin buffer, which becomes the ensemble)
// inject != nullptr: the injection behind the survivors rides in the same launch (inject->ens is set here)
int doCompaction(emcgpu_ctx *ctx, InjectParams *inject = nullptr) {
  DeviceRunState *r = ctx->run;
  RunCtl *ctl = r->dCtl.as<RunCtl>();
  const int32_t *flag = r->dFlag.as<const int32_t>();
  int32_t *chunkCount = r->dChunkCount.as<int32_t>();
  const int grid = particleGrid(ctx, kChunk, 8);
  CUDA_TRY(ctx, chainedLaunch(ctx, selectCountKernel<SELECT_KEPT>, grid, kChunk, 0, flag, ctl, chunkCount));
  const bool replay = ctx->rngMode == RNG_REPLAY;
  if (replay) CUDA_TRY(ctx, r->altCursor.ensure((size_t)ctx->capacity * sizeof(uint32_t)));
  const bool grain = ctx->grainOn;
  EnsemblePtrs src = ptrsOf(ctx->dEnsemble.ptr, ctx->capacity, replay ? ctx->dCursor.as<uint32_t>() : nullptr,
                            grain ? ctx->dGrain.as<double>() : nullptr);
  EnsemblePtrs dst = ptrsOf(r->altEnsemble.ptr, ctx->capacity, replay ? r->altCursor.as<uint32_t>() : nullptr,
                            grain ? r->altGrain.as<double>() : nullptr);
  if (inject) {
    const int injectGrid = 4; // a few hundred particles per step at most; grid-stride
    inject->ens = dst;
    if (r->dim == 2)
      CUDA_TRY(ctx, chainedLaunch(ctx, compactInjectKernel<2>, grid + injectGrid, kChunk, 0, r->geo, *inject, flag, (const int32_t *)chunkCount, src, grid));
    else
      CUDA_TRY(ctx, chainedLaunch(ctx, compactInjectKernel<3>, grid + injectGrid, kChunk, 0, r->geo, *inject, flag, (const int32_t *)chunkCount, src, grid));
  } else {
    compactScatterKernel<<<grid, kChunk, 0, ctx->stream>>>(flag, ctl, chunkCount, src, dst);
  }
  ctx->launches += 2;
  CUDA_TRY(ctx, cudaGetLastError());
  std::swap(ctx->dEnsemble, r->altEnsemble);
  if (replay) std::swap(ctx->dCursor, r->altCursor);
  if (grain) std::swap(ctx->dGrain, r->altGrain);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) ctx->dStream[s] = dst.stream[s];
  ctx->dPacked = dst.packed;
  return EMCGPU_OK;
}

// ohmic contacts on the flags of the step (fromStep) or of the resting ensemble: excess particles -> kGone,
// compaction, injection behind the survivors
// countedByStep: the step kernel has counted the reservoir particles already (doStep(..., true))
int doContacts(emcgpu_ctx *ctx, bool fromStep, const uint64_t *replayDraws, int64_t nReplay, bool countedByStep = false) {
  DeviceRunState *r = ctx->run;
  const DevGeometry &G = r->geo;
  RunCtl *ctl = r->dCtl.as<RunCtl>();
  int32_t *flag = r->dFlag.as<int32_t>();
  int32_t *chunkCount = r->dChunkCount.as<int32_t>();
  const int grid = particleGrid(ctx, kChunk, 8);
  if (!fromStep) {
    reservoirFlagKernel<<<particleGrid(ctx, 256, 4), 256, 0, ctx->stream>>>(G, ctx->dStream[EMCGPU_X], ctx->dStream[EMCGPU_Y],
                                                                            ctx->dStream[EMCGPU_Z], ctl, flag);
    ctx->launches++;
  }
  if (!countedByStep) selectCountKernel<SELECT_RESERVOIR><<<grid, kChunk, 0, ctx->stream>>>(flag, ctl, chunkCount);
  CUDA_TRY(ctx, chainedLaunch(ctx, reservoirListKernel, grid, kChunk, 0, (const int32_t *)flag, (const RunCtl *)ctl, (const int32_t *)chunkCount,
                              r->dListParticle.as<int32_t>(), r->dListCell.as<int32_t>()));
  ContactParams K;
  K.nrCarriers = r->nrCarriers;
  K.expected = r->grid[EMCGPU_GRID_EXPECTED].as<const double>();
  K.listParticle = r->dListParticle.as<const int32_t>();
  K.listCell = r->dListCell.as<const int32_t>();
  K.flag = flag;
  K.cellCount = r->dCellCount.as<int32_t>();
  K.injectCount = r->dInjectCount.as<int32_t>();
  K.ctl = ctl;
  K.share = nullptr;
  K.rank = r->rank;
  K.world = r->world;
  if (r->world > 1) {
    const size_t n = (size_t)r->world * G.cells;
    CUDA_TRY(ctx, r->dShare.ensure(n * sizeof(double)));
    CUDA_TRY(ctx, cudaMemsetAsync(r->dShare.ptr, 0, n * sizeof(double), ctx->stream));
    contactShareKernel<<<std::min(grid, 2 * ctx->smCount), 256, 0, ctx->stream>>>(K, r->dShare.as<double>() + (size_t)r->rank * G.cells);
    ctx->launches++;
    r->allreduce(r->allreduceUser, r->dShare.as<double>(), (int64_t)n, ctx->stream);
    K.share = r->dShare.as<const double>();
  }
  {
    const size_t counterBytes = (size_t)G.cells * sizeof(int);
    const int inSmem = counterBytes <= (size_t)ctx->maxSmemOptin - 20480 ? 1 : 0;
    if (inSmem) CUDA_TRY(ctx, cudaFuncSetAttribute(contactRankKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)counterBytes));
    CUDA_TRY(ctx, chainedLaunch(ctx, contactRankKernel, 1, kRankThreads, inSmem ? counterBytes : 0, G, K, inSmem));
  }
  ctx->launches += 3;
  CUDA_TRY(ctx, cudaGetLastError());
  InjectParams J;
  J.grainTau0 = ctx->grainTau0;
  J.injectCount = K.injectCount;
  J.model = ctx->dModel.as<const DevModel>();
  J.seed = ctx->seed;
  J.ctl = ctl;
  J.replay = nullptr;
  J.replayCount = 0;
  J.status = ctx->dStatus.as<int>();
  J.rank = (uint32_t)r->rank;
  if (replayDraws) {
    CUDA_TRY(ctx, r->dReplay.ensure((size_t)std::max<int64_t>(1, nReplay) * sizeof(uint64_t)));
    CUDA_TRY(ctx, cudaMemcpyAsync(r->dReplay.ptr, replayDraws, nReplay * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    J.replay = r->dReplay.as<const uint64_t>();
    J.replayCount = nReplay;
  }
  return doCompaction(ctx, &J); // compaction and injection in one launch
}

} // namespace

extern "C" {

int emcgpu_device_configure(emcgpu_ctx *ctx, const emcgpu_device_t *dev, double charge, double nrCarriers,
                            const double *expected, int mathMode) {
  if (!ctx || !dev) return EMCGPU_E_INVALID;
  if (dev->dim != 2 && dev->dim != 3) return fail(ctx, EMCGPU_E_INVALID, "device dimension must be 2 or 3");
  if (dev->nContacts < 0 || dev->nContacts > kMaxContacts)
    return fail(ctx, EMCGPU_E_CAPACITY, "%d contacts exceed EMCGPU_MAX_CONTACTS=%d", dev->nContacts, kMaxContacts);
  if (!dev->region || !dev->faceContact || !dev->doping || (dev->nContacts && (!dev->contactType || !dev->contactVoltage)))
    return fail(ctx, EMCGPU_E_INVALID, "NULL device array");
  if (mathMode != EMCGPU_MATH_EXACT && mathMode != EMCGPU_MATH_FAST) return fail(ctx, EMCGPU_E_INVALID, "unknown math mode");
  if (dev->pmScheme < EMCGPU_PM_NGP || dev->pmScheme > EMCGPU_PM_NEC_VWD)
    return fail(ctx, EMCGPU_E_INVALID, "unknown particle-mesh scheme %d", dev->pmScheme);
  if (dev->dim == 3 && (dev->pmScheme == EMCGPU_PM_NEC || dev->pmScheme == EMCGPU_PM_NEC_VWD))
    return fail(ctx, EMCGPU_E_INVALID, "the NEC schemes interpolate forces in 2-D only (emcNECScheme.hpp:99-113)");
  if (!(nrCarriers > 0)) return fail(ctx, EMCGPU_E_INVALID, "nrCarriersPerParticle must be positive");
  if (int r = emc::bindDevice(ctx)) return r;
  int64_t cells = 1;
  for (int i = 0; i < dev->dim; i++) {
    if (dev->extent[i] < 3) return fail(ctx, EMCGPU_E_INVALID, "grid extent must be at least 3 per dimension");
    if (!(dev->spacing[i] > 0) || !(dev->maxPos[i] > 0)) return fail(ctx, EMCGPU_E_INVALID, "bad spacing / maxPos");
    cells *= dev->extent[i];
  }
  if (cells > (1 << 28)) return fail(ctx, EMCGPU_E_CAPACITY, "grid too large");
  // every argument is checked BEFORE the previous configuration is given up
  for (int c = 0; c < dev->nContacts; c++)
    if (dev->contactType[c] == EMCGPU_CONTACT_GATE && !(dev->gateThickness && dev->gateThickness[c] > 0))
      return fail(ctx, EMCGPU_E_INVALID, "gate contact %d needs an oxide thickness", c);
  for (int64_t i = 0; i < cells * 2 * dev->dim; i++)
    if (dev->faceContact[i] < -2 || dev->faceContact[i] >= dev->nContacts)
      return fail(ctx, EMCGPU_E_INVALID, "faceContact entry %lld out of range", (long long)i);
  // a failed allocation below leaves NO configuration (needRun() then rejects the run calls) rather than half of one
  struct Guard {
    emcgpu_ctx *ctx;
    bool ok = false;
    ~Guard() {
      if (!ok) emc::releaseDeviceRun(ctx);
    }
  } guard{ctx};
  emc::releaseDeviceRun(ctx);
  DeviceRunState *r = ctx->run = new DeviceRunState();
  DevGeometry &G = r->geo;
  memset(&G, 0, sizeof G);
  G.dim = r->dim = dev->dim;
  G.nContacts = dev->nContacts;
  G.cells = (int32_t)cells;
  G.pmScheme = dev->pmScheme;
  for (int i = 0; i < 3; i++) {
    G.extent[i] = i < dev->dim ? dev->extent[i] : 1;
    G.spacing[i] = i < dev->dim ? dev->spacing[i] : 1.0;
    G.maxPos[i] = i < dev->dim ? dev->maxPos[i] : 0.0;
  }
  G.thermalVoltage = dev->thermalVoltage;
  G.debyeLength = dev->debyeLength;
  G.ni = dev->ni;
  G.cellVolume = dev->cellVolume;
  G.epsR = dev->epsR;
  for (int c = 0; c < dev->nContacts; c++) {
    G.contactType[c] = dev->contactType[c];
    G.contactVoltage[c] = dev->contactVoltage[c];
    G.gateEpsOx[c] = dev->gateEpsOx ? dev->gateEpsOx[c] : 0.0;
    G.gateThickness[c] = dev->gateThickness ? dev->gateThickness[c] : 0.0;
    G.gateBarrier[c] = dev->gateBarrier ? dev->gateBarrier[c] : 0.0;
  }
  r->charge = charge;
  r->nrCarriers = nrCarriers;
  ctx->mathMode = mathMode;
  CUDA_TRY(ctx, r->dRegion.ensure(cells * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dFace.ensure(cells * 2 * dev->dim));
  CUDA_TRY(ctx, r->dDoping.ensure(cells * sizeof(double)));
  CUDA_TRY(ctx, cudaMemcpy(r->dRegion.ptr, dev->region, cells * sizeof(int32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(ctx, cudaMemcpy(r->dFace.ptr, dev->faceContact, cells * 2 * dev->dim, cudaMemcpyHostToDevice));
  CUDA_TRY(ctx, cudaMemcpy(r->dDoping.ptr, dev->doping, cells * sizeof(double), cudaMemcpyHostToDevice));
  G.region = r->dRegion.as<const int32_t>();
  G.faceContact = r->dFace.as<const int8_t>();
  G.doping = r->dDoping.as<const double>();
  {
    // per-cell constants of the Poisson sweeps
    std::vector<double> norm(cells);
    std::vector<uint8_t> kind(cells, 0);
    for (int64_t i = 0; i < cells; i++) {
      norm[i] = dev->doping[i] / dev->ni;
      const int8_t *fc = dev->faceContact + i * 2 * dev->dim;
      int first = -1;
      bool seen = false;
      for (int f = 0; f < 2 * dev->dim; f++) {
        if (fc[f] == -2) continue;
        if (!seen) first = fc[f], seen = true;
        if (fc[f] >= 0 && dev->contactType[fc[f]] == EMCGPU_CONTACT_GATE) kind[i] |= 2;
      }
      if (first >= 0 && dev->contactType[first] != EMCGPU_CONTACT_GATE) kind[i] |= 1;
    }
    CUDA_TRY(ctx, r->dDopingNorm.ensure(cells * sizeof(double)));
    CUDA_TRY(ctx, r->dCellKind.ensure(cells));
    CUDA_TRY(ctx, cudaMemcpy(r->dDopingNorm.ptr, norm.data(), cells * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(r->dCellKind.ptr, kind.data(), cells, cudaMemcpyHostToDevice));
    G.dopingNorm = r->dDopingNorm.as<const double>();
    G.cellKind = r->dCellKind.as<const uint8_t>();
  }
  // E field components are one allocation [dim][cells] starting at EFIELD_X
  for (int g = 0; g < EMCGPU_N_GRIDS; g++) {
    if (g == EMCGPU_GRID_EFIELD_Y || g == EMCGPU_GRID_EFIELD_Z) continue;
    const size_t bytes = (g == EMCGPU_GRID_EFIELD_X ? 3 : 1) * cells * sizeof(double);
    CUDA_TRY(ctx, r->grid[g].ensure(bytes));
    CUDA_TRY(ctx, cudaMemset(r->grid[g].ptr, 0, bytes));
  }
  CUDA_TRY(ctx, r->dCtl.ensure(sizeof(RunCtl)));
  CUDA_TRY(ctx, cudaMemset(r->dCtl.ptr, 0, sizeof(RunCtl)));
  CUDA_TRY(ctx, r->dCellCount.ensure(cells * sizeof(int32_t)));
  CUDA_TRY(ctx, cudaMemset(r->dCellCount.ptr, 0, cells * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dCounters.ensure((size_t)kRunChunk * 2 * kMaxContacts * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dSweepsPerStep.ensure((size_t)kRunChunk * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dInjectCount.ensure((cells + 1) * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dSweeps.ensure(sizeof(int32_t)));
  CUDA_TRY(ctx, cudaMemset(r->dSweeps.ptr, 0, sizeof(int32_t)));
  // initial guess of the potential and the contact populations
  std::vector<double> pot(cells), exp(cells, 0.0);
  for (int64_t i = 0; i < cells; i++) pot[i] = std::asinh(0.5 * (dev->doping[i] / dev->ni));
  if (expected) {
    std::copy(expected, expected + cells, exp.begin());
  } else {
    for (int64_t i = 0; i < cells; i++) {
      const int8_t *fc = dev->faceContact + i * 2 * dev->dim;
      int contact = -1;
      for (int f = 0; f < 2 * dev->dim; f++)
        if (fc[f] != -2) {
          contact = fc[f];
          break;
        }
      if (contact < 0 || dev->contactType[contact] == EMCGPU_CONTACT_GATE) continue;
      double v = dev->cellVolume * dev->doping[i];
      int64_t rem = i;
      for (int d = 0; d < dev->dim; d++) {
        const int64_t c = rem % dev->extent[d];
        rem /= dev->extent[d];
        if (c == 0 || c == dev->extent[d] - 1) v *= 0.5;
      }
      exp[i] = v;
    }
  }
  CUDA_TRY(ctx, cudaMemcpy(r->grid[EMCGPU_GRID_POTENTIAL].ptr, pot.data(), cells * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_TRY(ctx, cudaMemcpy(r->grid[EMCGPU_GRID_EXPECTED].ptr, exp.data(), cells * sizeof(double), cudaMemcpyHostToDevice));
  guard.ok = true;
  return EMCGPU_OK;
}

static double *gridPtr(emcgpu_ctx *ctx, int grid) {
  DeviceRunState *r = ctx->run;
  if (grid == EMCGPU_GRID_EFIELD_Y) return r->grid[EMCGPU_GRID_EFIELD_X].as<double>() + r->geo.cells;
  if (grid == EMCGPU_GRID_EFIELD_Z) return r->grid[EMCGPU_GRID_EFIELD_X].as<double>() + 2 * (size_t)r->geo.cells;
  return r->grid[grid].as<double>();
}

int emcgpu_device_set_surface(emcgpu_ctx *ctx, int face, int kind, double parameter) {
  if (int r = needRun(ctx)) return r;
  if (face < 0 || face >= 2 * ctx->run->dim) return fail(ctx, EMCGPU_E_INVALID, "face %d does not exist in a %d-D device", face, ctx->run->dim);
  if (kind < EMCGPU_SURFACE_SPECULAR || kind > EMCGPU_SURFACE_MOMENTUM_DEPENDENT)
    return fail(ctx, EMCGPU_E_UNSUPPORTED_MECHANISM, "surface scatter mechanism %d has no device implementation", kind);
  ctx->run->geo.surfaceKind[face] = kind;
  ctx->run->geo.surfaceParam[face] = parameter;
  return EMCGPU_OK;
}

int emcgpu_device_set_sharding(emcgpu_ctx *ctx, int rank, int world, emcgpu_allreduce_fn allreduceSum, void *user) {
  if (int r = needRun(ctx)) return r;
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, EMCGPU_E_INVALID, "rank %d of %d", rank, world);
  if (world > 1 && !allreduceSum) return fail(ctx, EMCGPU_E_INVALID, "a sharded run needs the all-reduce callback");
  ctx->run->rank = rank;
  ctx->run->world = world;
  ctx->run->allreduce = allreduceSum;
  ctx->run->allreduceUser = user;
  return EMCGPU_OK;
}

int emcgpu_device_set_particle_kind(emcgpu_ctx *ctx, int kind) {
  if (int r = needRun(ctx)) return r;
  if (kind != EMCGPU_PARTICLE_ELECTRON && kind != EMCGPU_PARTICLE_ELECTRON_VWD)
    return fail(ctx, EMCGPU_E_INVALID, "unknown particle kind %d", kind);
  ctx->run->geo.particleKind = kind;
  return EMCGPU_OK;
}

int emcgpu_device_set_grid(emcgpu_ctx *ctx, int grid, const double *host) {
  if (int r = needRun(ctx)) return r;
  if (grid < 0 || grid >= EMCGPU_N_GRIDS || !host) return fail(ctx, EMCGPU_E_INVALID, "bad grid argument");
  if (grid == EMCGPU_GRID_EFIELD_Z && ctx->run->dim < 3) return fail(ctx, EMCGPU_E_INVALID, "no z field in a 2-D device");
  if (grid == EMCGPU_GRID_EXPECTED) ctx->run->maxInject = -1;
  CUDA_TRY(ctx, cudaMemcpyAsync(gridPtr(ctx, grid), host, (size_t)ctx->run->geo.cells * sizeof(double), cudaMemcpyHostToDevice,
                                ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_device_get_grid(emcgpu_ctx *ctx, int grid, double *host) {
  if (int r = needRun(ctx)) return r;
  if (grid < 0 || grid >= EMCGPU_N_GRIDS || !host) return fail(ctx, EMCGPU_E_INVALID, "bad grid argument");
  if (grid == EMCGPU_GRID_EFIELD_Z && ctx->run->dim < 3) return fail(ctx, EMCGPU_E_INVALID, "no z field in a 2-D device");
  CUDA_TRY(ctx, cudaMemcpyAsync(host, gridPtr(ctx, grid), (size_t)ctx->run->geo.cells * sizeof(double), cudaMemcpyDeviceToHost,
                                ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_device_reserve(emcgpu_ctx *ctx, int64_t capacity) {
  if (int r = needRun(ctx)) return r;
  ctx->run->reserve = capacity;
  return growEnsemble(ctx, capacity);
}

int emcgpu_device_poisson(emcgpu_ctx *ctx, int equilibrium, double accuracyVolt, double omega, int resetBC, int32_t *sweeps) {
  if (int r = needRun(ctx)) return r;
  if (!(accuracyVolt > 0) || !(omega > 0 && omega < 2)) return fail(ctx, EMCGPU_E_INVALID, "bad SOR parameters");
  if (int r = doPoisson(ctx, equilibrium != 0, accuracyVolt, omega, resetBC != 0, false, false)) return r;
  if (sweeps)
    CUDA_TRY(ctx, cudaMemcpyAsync(sweeps, ctx->run->dSweeps.ptr, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_device_efield(emcgpu_ctx *ctx) {
  if (int r = needRun(ctx)) return r;
  return doEfield(ctx);
}

int emcgpu_device_assign(emcgpu_ctx *ctx) {
  if (int r = needRun(ctx)) return r;
  if (int r = pushCtl(ctx, 0)) return r;
  return doAssign(ctx, false, false);
}

int emcgpu_device_concentration(emcgpu_ctx *ctx) {
  if (int r = needRun(ctx)) return r;
  return doConcentration(ctx);
}

int emcgpu_device_step(emcgpu_ctx *ctx, double dt, int32_t *removedPerContact) {
  if (int r = needRun(ctx)) return r;
  if (int r = needModel(ctx)) return r;
  if (!(dt > 0)) return fail(ctx, EMCGPU_E_INVALID, "dt must be positive");
  if (int r = pushCtl(ctx, 0)) return r;
  if (int r = doStep(ctx, dt)) return r;
  if (int r = doCompaction(ctx)) return r;
  RunCtl h;
  if (int r = pullCtl(ctx, &h)) return r;
  ctx->n = h.nKept;
  ctx->nextStep = h.step + 1;
  for (int c = 0; c < ctx->run->geo.nContacts && removedPerContact; c++) removedPerContact[c] = h.removedPerContact[c];
  return readStatus(ctx);
}

} // extern "C"
namespace {
// Upper bound of the particles the contacts inject in ONE step: a reservoir cell never misses more than its expected
// population (handleOhmicContacts, emcBasicParticleHandler.hpp:158-192).  The ensemble therefore grows by at most this
// many particles per step, whatever the bias or the initial population.
int updateMaxInject(emcgpu_ctx *ctx) {
  DeviceRunState *r = ctx->run;
  if (r->maxInject >= 0) return EMCGPU_OK;
  std::vector<double> exp(r->geo.cells);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaMemcpy(exp.data(), r->grid[EMCGPU_GRID_EXPECTED].ptr, exp.size() * sizeof(double), cudaMemcpyDeviceToHost));
  double total = 0;
  for (double v : exp) total += std::ceil(std::max(0.0, v) / r->nrCarriers);
  r->maxInject = (int64_t)total;
  return EMCGPU_OK;
}
} // namespace
extern "C" {

int emcgpu_device_contacts(emcgpu_ctx *ctx, int32_t *netPerContact, const uint64_t *replayDraws, int64_t nReplayDraws) {
  if (int r = needRun(ctx)) return r;
  if (int r = needModel(ctx)) return r;
  // one reservoir cell never misses more than its expected population: room for the worst case
  {
    DeviceRunState *r = ctx->run;
    if (int rc = updateMaxInject(ctx)) return rc;
    if (int rc = growEnsemble(ctx, std::max<int64_t>(ctx->n + r->maxInject, r->reserve))) return rc;
  }
  ctx->nextStep--; // the contacts belong to the step that was just done (Philox counter of the injected particles)
  if (int r = pushCtl(ctx, 0)) return r;
  ctx->nextStep++;
  if (int r = doContacts(ctx, false, replayDraws, nReplayDraws)) return r;
  RunCtl h;
  if (int r = pullCtl(ctx, &h)) return r;
  ctx->n = h.nKept + h.toInject;
  ctx->nextStep = h.step + 1;
  for (int c = 0; c < ctx->run->geo.nContacts && netPerContact; c++) netPerContact[c] = h.net[c];
  if (ctx->rngMode == RNG_REPLAY && h.toInject > 0 && ctx->dCursor.bytes < (size_t)ctx->capacity * sizeof(uint32_t))
    return fail(ctx, EMCGPU_E_INVALID, "replay streams do not cover injected particles: upload the ensemble again");
  return readStatus(ctx);
}

int emcgpu_device_run(emcgpu_ctx *ctx, double dt, int nSteps, double accuracyVolt, double omega, int resetBCFirst,
                      int32_t *counters, int32_t *sweeps) {
  return emcgpu_device_run_averaging(ctx, dt, nSteps, 0, accuracyVolt, omega, resetBCFirst, counters, sweeps);
}

int emcgpu_device_run_averaging(emcgpu_ctx *ctx, double dt, int nSteps, int nAverage, double accuracyVolt, double omega,
                                int resetBCFirst, int32_t *counters, int32_t *sweeps) {
  if (int r = needRun(ctx)) return r;
  if (int r = needModel(ctx)) return r;
  if (!(dt > 0) || nSteps < 1 || !(accuracyVolt > 0) || nAverage < 0 || nAverage > nSteps)
    return fail(ctx, EMCGPU_E_INVALID, "bad run arguments");
  DeviceRunState *r = ctx->run;
  const int nC = r->geo.nContacts;
  std::vector<int32_t> hCounters((size_t)kRunChunk * 2 * std::max(1, nC)), hSweeps(kRunChunk);
  for (int done = 0; done < nSteps;) {
    // Head room for the particles the contacts may inject during the chunk.  The ensemble grows by at most maxInject per
    // step (see updateMaxInject), so capacity >= n + chunk * maxInject cannot overflow inside the chunk: grow to a whole
    // chunk's worth when that is cheap (a few times the ensemble), otherwise shorten the chunk to what the room allows.
    if (int rc = updateMaxInject(ctx)) return rc;
    const int64_t perStep = std::max<int64_t>(1, r->maxInject);
    const int64_t wholeChunk = ctx->n + (int64_t)kRunChunk * perStep;
    const int64_t want = std::max<int64_t>(std::min<int64_t>(wholeChunk, 4 * ctx->n + ((int64_t)1 << 20)), ctx->n + perStep);
    if (ctx->capacity < want || ctx->capacity < r->reserve)
      if (int rc = growEnsemble(ctx, std::max<int64_t>(want + want / 8, r->reserve))) return rc;
    const int chunk = (int)std::min<int64_t>(std::min(kRunChunk, nSteps - done), std::max<int64_t>(1, (ctx->capacity - ctx->n) / perStep));
    if (int rc = pushCtl(ctx, std::max(0, (nSteps - nAverage) - done))) return rc;
    r->chained = ctx->optEarlyStep != 0 && r->world == 1;
    for (int s = 0; s < chunk; s++) {
      // performEMCStep (emcSimulation.hpp:177-192)
      if (int rc = doPoisson(ctx, false, accuracyVolt, omega, resetBCFirst && done + s == 0, true, true)) return rc;
      if (int rc = doStep(ctx, dt, true, r->chained)) return rc;
      if (int rc = doContacts(ctx, true, nullptr, 0, true)) return rc;
      if (int rc = doAssign(ctx, true, true)) return rc;
    }
    r->chained = false;
    if (counters)
      CUDA_TRY(ctx, cudaMemcpyAsync(hCounters.data(), r->dCounters.ptr, (size_t)chunk * 2 * nC * sizeof(int32_t),
                                    cudaMemcpyDeviceToHost, ctx->stream));
    if (sweeps)
      CUDA_TRY(ctx, cudaMemcpyAsync(hSweeps.data(), r->dSweepsPerStep.ptr, (size_t)chunk * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                    ctx->stream));
    if (int rc = pullCtl(ctx)) return rc;
    if (int rc = readStatus(ctx)) return rc;
    if (counters) std::copy(hCounters.begin(), hCounters.begin() + (size_t)chunk * 2 * nC, counters + (size_t)done * 2 * nC);
    if (sweeps) std::copy(hSweeps.begin(), hSweeps.begin() + chunk, sweeps + done);
    done += chunk;
  }
  return EMCGPU_OK;
}

} // extern "C"
