// C-ABI implementation of the device-run path (include/emcgpu.h, emcgpu_device_*):
// grids, Poisson, field, charge assignment, particle step with boundaries, contacts.
#include <algorithm>
#include <cmath>

#include "emc_device_run.cuh"
#include "emcgpu_internal.cuh"

using namespace emc;

namespace emc {

struct DeviceRunState {
  DevGeometry geo{};
  int dim = 2;
  double charge = 0, nrCarriers = 1;
  DeviceBuffer dRegion, dFace, dDoping;
  DeviceBuffer grid[EMCGPU_N_GRIDS];
  DeviceBuffer dRemoved, dRemovedPerContact, dBlockCount, dHave, dNet, dInjectCount, dSweeps, dReplay;
  DeviceBuffer altEnsemble, altCursor; // second ensemble buffer for the order-preserving compaction
  int64_t reserve = 0;
};

void releaseDeviceRun(emcgpu_ctx *ctx) {
  if (!ctx->run) return;
  DeviceRunState *r = ctx->run;
  for (DeviceBuffer *b : {&r->dRegion, &r->dFace, &r->dDoping, &r->dRemoved, &r->dRemovedPerContact, &r->dBlockCount,
                          &r->dHave, &r->dNet, &r->dInjectCount, &r->dSweeps, &r->dReplay, &r->altEnsemble,
                          &r->altCursor})
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
  return EMCGPU_OK;
}

size_t streamStrideD(int64_t cap) { return ((size_t)cap * sizeof(double) + 255) & ~size_t(255); }
size_t streamStrideP(int64_t cap) { return ((size_t)cap * sizeof(uint32_t) + 255) & ~size_t(255); }

EnsemblePtrs ptrsOf(void *base, int64_t cap, uint32_t *cursor) {
  EnsemblePtrs p;
  unsigned char *b = static_cast<unsigned char *>(base);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) p.stream[s] = reinterpret_cast<double *>(b + streamStrideD(cap) * s);
  p.packed = reinterpret_cast<uint32_t *>(b + streamStrideD(cap) * EMCGPU_N_STREAMS);
  p.cursor = cursor;
  return p;
}

// make sure the ensemble allocation (and its twin) can hold `cap` particles, keeping the current content
int growEnsemble(emcgpu_ctx *ctx, int64_t cap) {
  DeviceRunState *r = ctx->run;
  if (cap <= ctx->capacity && r->altEnsemble.bytes >= ctx->dEnsemble.bytes) return EMCGPU_OK;
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
  return EMCGPU_OK;
}

// drop the particles flagged in `drop`, keeping the order of the others; returns the new count through ctx->n
int compactEnsemble(emcgpu_ctx *ctx, const int8_t *drop) {
  DeviceRunState *r = ctx->run;
  const int64_t n = ctx->n;
  if (n == 0) return EMCGPU_OK;
  if (int rc = growEnsemble(ctx, ctx->capacity)) return rc;
  const int nBlocks = (int)((n + kCompactThreads - 1) / kCompactThreads);
  CUDA_TRY(ctx, r->dBlockCount.ensure((size_t)(nBlocks + 1) * sizeof(int32_t)));
  int32_t *blockCount = r->dBlockCount.as<int32_t>();
  compactCountKernel<<<nBlocks, kCompactThreads, 0, ctx->stream>>>(drop, n, blockCount);
  compactScanKernel<<<1, 1024, 0, ctx->stream>>>(blockCount, nBlocks);
  const bool replay = ctx->rngMode == RNG_REPLAY;
  if (replay) CUDA_TRY(ctx, r->altCursor.ensure((size_t)ctx->capacity * sizeof(uint32_t)));
  EnsemblePtrs src = ptrsOf(ctx->dEnsemble.ptr, ctx->capacity, replay ? ctx->dCursor.as<uint32_t>() : nullptr);
  EnsemblePtrs dst = ptrsOf(r->altEnsemble.ptr, ctx->capacity, replay ? r->altCursor.as<uint32_t>() : nullptr);
  compactScatterKernel<<<nBlocks, kCompactThreads, 0, ctx->stream>>>(drop, n, blockCount, src, dst);
  ctx->launches += 3;
  CUDA_TRY(ctx, cudaGetLastError());
  int32_t kept = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&kept, blockCount + nBlocks, sizeof kept, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  std::swap(ctx->dEnsemble, r->altEnsemble);
  if (replay) std::swap(ctx->dCursor, r->altCursor);
  for (int s = 0; s < EMCGPU_N_STREAMS; s++) ctx->dStream[s] = dst.stream[s];
  ctx->dPacked = dst.packed;
  ctx->n = kept;
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
}

template <bool EXACT, int MODE> cudaError_t launchStepDim(emcgpu_ctx *ctx, const DeviceStepParams &D, size_t smem, int grid) {
  const DevGeometry &G = ctx->run->geo;
  auto go = [&](auto kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kBulkThreads, smem, ctx->stream>>>(G, D);
    ctx->launches++;
    return cudaGetLastError();
  };
  return ctx->run->dim == 2 ? go(deviceStepKernel<EXACT, MODE, 2>) : go(deviceStepKernel<EXACT, MODE, 3>);
}

int readStatus(emcgpu_ctx *ctx) {
  int status = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&status, ctx->dStatus.ptr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (status != 0) {
    cudaMemsetAsync(ctx->dStatus.ptr, 0, sizeof(int), ctx->stream);
    if (status == EMCGPU_E_REPLAY_EXHAUSTED)
      return fail(ctx, status, "replay stream exhausted: more draws were needed than were recorded");
    return fail(ctx, status, "device reported status %d", status);
  }
  return EMCGPU_OK;
}

int gridBlocks(int cells) { return (cells + 255) / 256; }

// ---- the pieces of one EMC step, asynchronous on the context's stream unless they return counters -------
int doPoisson(emcgpu_ctx *ctx, bool equilibrium, double accuracyVolt, double omega, bool resetBC, int32_t *sweepsHost) {
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
  S.potInSmem = smem <= (size_t)ctx->maxSmemOptin - 1024 ? 1 : 0;
  S.sweepsOut = r->dSweeps.as<int32_t>();
  if (S.potInSmem)
    CUDA_TRY(ctx, cudaFuncSetAttribute(sorKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sorKernel<<<1, kSorThreads, S.potInSmem ? smem : 0, ctx->stream>>>(G, S);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  if (sweepsHost) {
    CUDA_TRY(ctx, cudaMemcpyAsync(sweepsHost, r->dSweeps.ptr, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
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

int doAssign(emcgpu_ctx *ctx) {
  DeviceRunState *r = ctx->run;
  const DevGeometry &G = r->geo;
  double *count = r->grid[EMCGPU_GRID_COUNT].as<double>();
  CUDA_TRY(ctx, cudaMemsetAsync(count, 0, (size_t)G.cells * sizeof(double), ctx->stream));
  if (ctx->n == 0) return EMCGPU_OK;
  AssignParams A;
  A.x = ctx->dStream[EMCGPU_X];
  A.y = ctx->dStream[EMCGPU_Y];
  A.z = ctx->dStream[EMCGPU_Z];
  A.n = ctx->n;
  A.nrCarriers = r->nrCarriers;
  A.count = count;
  const size_t smem = (size_t)G.cells * sizeof(double);
  A.useSmem = smem <= 96 * 1024 ? 1 : 0;
  if (A.useSmem) CUDA_TRY(ctx, cudaFuncSetAttribute(ngpAssignKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n + 255) / 256, ctx->smCount));
  ngpAssignKernel<<<grid, 256, A.useSmem ? smem : 0, ctx->stream>>>(G, A);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
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

int doStep(emcgpu_ctx *ctx, double dt, int32_t *removedHost) {
  DeviceRunState *r = ctx->run;
  const DevGeometry &G = r->geo;
  for (int c = 0; c < G.nContacts && removedHost; c++) removedHost[c] = 0;
  if (ctx->n == 0) {
    ctx->nextStep++;
    return EMCGPU_OK;
  }
  CUDA_TRY(ctx, r->dRemoved.ensure((size_t)ctx->capacity));
  CUDA_TRY(ctx, cudaMemsetAsync(r->dRemovedPerContact.ptr, 0, kMaxContacts * sizeof(int32_t), ctx->stream));
  DeviceStepParams D;
  fillParams(ctx, D.P);
  D.P.dt = dt;
  D.P.nSteps = 1;
  D.P.step0 = ctx->nextStep;
  D.e = r->grid[EMCGPU_GRID_EFIELD_X].as<const double>();
  D.charge = r->charge;
  D.removed = r->dRemoved.as<int8_t>();
  D.removedPerContact = r->dRemovedPerContact.as<int32_t>();
  bool inSmem = true;
  size_t smem = BulkSmem(0, ctx->hModel.nValleys, (int)ctx->hMechs.size(), ctx->hModel.tableDoubles, true, 0).total;
  if (smem > (size_t)ctx->maxSmemOptin / 2) { // two CTAs per SM
    inSmem = false;
    smem = BulkSmem(0, ctx->hModel.nValleys, (int)ctx->hMechs.size(), ctx->hModel.tableDoubles, false, 0).total;
  }
  D.P.tablesInSmem = inSmem ? 1 : 0;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((ctx->n + kBulkThreads - 1) / kBulkThreads, 2 * ctx->smCount));
  const bool exact = ctx->mathMode == EMCGPU_MATH_EXACT;
  cudaError_t e;
  if (ctx->rngMode == RNG_PHILOX)
    e = exact ? launchStepDim<true, RNG_PHILOX>(ctx, D, smem, grid) : launchStepDim<false, RNG_PHILOX>(ctx, D, smem, grid);
  else
    e = exact ? launchStepDim<true, RNG_REPLAY>(ctx, D, smem, grid) : launchStepDim<false, RNG_REPLAY>(ctx, D, smem, grid);
  if (e != cudaSuccess) return fail(ctx, EMCGPU_E_CUDA, "device step launch failed: %s", cudaGetErrorString(e));
  ctx->nextStep++;
  if (removedHost)
    CUDA_TRY(ctx, cudaMemcpyAsync(removedHost, r->dRemovedPerContact.ptr, G.nContacts * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, ctx->stream));
  if (int rc = compactEnsemble(ctx, r->dRemoved.as<const int8_t>())) return rc; // synchronises
  return readStatus(ctx);
}

int doContacts(emcgpu_ctx *ctx, int32_t *netHost, const uint64_t *replayDraws, int64_t nReplay) {
  DeviceRunState *r = ctx->run;
  const DevGeometry &G = r->geo;
  CUDA_TRY(ctx, r->dRemoved.ensure((size_t)std::max<int64_t>(1, ctx->capacity)));
  ContactParams K;
  K.x = ctx->dStream[EMCGPU_X];
  K.y = ctx->dStream[EMCGPU_Y];
  K.z = ctx->dStream[EMCGPU_Z];
  K.n = ctx->n;
  K.nrCarriers = r->nrCarriers;
  K.expected = r->grid[EMCGPU_GRID_EXPECTED].as<const double>();
  K.have = r->dHave.as<double>();
  K.drop = r->dRemoved.as<int8_t>();
  K.net = r->dNet.as<int32_t>();
  K.injectCount = r->dInjectCount.as<int32_t>();
  contactMarkKernel<<<1, 32, 0, ctx->stream>>>(G, K);
  ctx->launches++;
  CUDA_TRY(ctx, cudaGetLastError());
  int32_t toInject = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&toInject, K.injectCount + G.cells, sizeof toInject, cudaMemcpyDeviceToHost, ctx->stream));
  if (netHost)
    CUDA_TRY(ctx, cudaMemcpyAsync(netHost, r->dNet.ptr, G.nContacts * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (int rc = compactEnsemble(ctx, K.drop)) return rc; // synchronises: toInject / netHost are valid now
  if (ctx->n == 0) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (toInject > 0) {
    if (int rc = growEnsemble(ctx, std::max<int64_t>(ctx->n + toInject, r->reserve))) return rc;
    InjectParams J;
    J.ens = ptrsOf(ctx->dEnsemble.ptr, ctx->capacity, ctx->rngMode == RNG_REPLAY ? ctx->dCursor.as<uint32_t>() : nullptr);
    J.first = ctx->n;
    J.injectCount = K.injectCount;
    J.model = ctx->dModel.as<const DevModel>();
    J.seed = ctx->seed;
    J.step = ctx->nextStep - 1; // the step whose contacts are handled
    J.replay = nullptr;
    J.replayCount = 0;
    J.status = ctx->dStatus.as<int>();
    if (replayDraws) {
      CUDA_TRY(ctx, r->dReplay.ensure((size_t)std::max<int64_t>(1, nReplay) * sizeof(uint64_t)));
      CUDA_TRY(ctx, cudaMemcpyAsync(r->dReplay.ptr, replayDraws, nReplay * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
      J.replay = r->dReplay.as<const uint64_t>();
      J.replayCount = nReplay;
    }
    const int grid = (toInject + 127) / 128;
    if (r->dim == 2)
      contactInjectKernel<2><<<grid, 128, 0, ctx->stream>>>(G, J);
    else
      contactInjectKernel<3><<<grid, 128, 0, ctx->stream>>>(G, J);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    if (ctx->rngMode == RNG_REPLAY && ctx->dCursor.bytes < (size_t)ctx->capacity * sizeof(uint32_t))
      return fail(ctx, EMCGPU_E_INVALID, "replay streams do not cover injected particles: upload the ensemble again");
    ctx->n += toInject;
    return readStatus(ctx);
  }
  return EMCGPU_OK;
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
  if (!(nrCarriers > 0)) return fail(ctx, EMCGPU_E_INVALID, "nrCarriersPerParticle must be positive");
  if (int r = emc::bindDevice(ctx)) return r;
  int64_t cells = 1;
  for (int i = 0; i < dev->dim; i++) {
    if (dev->extent[i] < 3) return fail(ctx, EMCGPU_E_INVALID, "grid extent must be at least 3 per dimension");
    if (!(dev->spacing[i] > 0) || !(dev->maxPos[i] > 0)) return fail(ctx, EMCGPU_E_INVALID, "bad spacing / maxPos");
    cells *= dev->extent[i];
  }
  if (cells > (1 << 28)) return fail(ctx, EMCGPU_E_CAPACITY, "grid too large");
  emc::releaseDeviceRun(ctx);
  DeviceRunState *r = ctx->run = new DeviceRunState();
  DevGeometry &G = r->geo;
  memset(&G, 0, sizeof G);
  G.dim = r->dim = dev->dim;
  G.nContacts = dev->nContacts;
  G.cells = (int32_t)cells;
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
    if (G.contactType[c] == EMCGPU_CONTACT_GATE && !(G.gateThickness[c] > 0))
      return fail(ctx, EMCGPU_E_INVALID, "gate contact %d needs an oxide thickness", c);
  }
  for (int64_t i = 0; i < cells * 2 * dev->dim; i++)
    if (dev->faceContact[i] < -2 || dev->faceContact[i] >= dev->nContacts)
      return fail(ctx, EMCGPU_E_INVALID, "faceContact entry %lld out of range", (long long)i);
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
  // E field components are one allocation [dim][cells] starting at EFIELD_X
  for (int g = 0; g < EMCGPU_N_GRIDS; g++) {
    if (g == EMCGPU_GRID_EFIELD_Y || g == EMCGPU_GRID_EFIELD_Z) continue;
    const size_t bytes = (g == EMCGPU_GRID_EFIELD_X ? 3 : 1) * cells * sizeof(double);
    CUDA_TRY(ctx, r->grid[g].ensure(bytes));
    CUDA_TRY(ctx, cudaMemset(r->grid[g].ptr, 0, bytes));
  }
  CUDA_TRY(ctx, r->dRemovedPerContact.ensure(kMaxContacts * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dNet.ensure(kMaxContacts * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dHave.ensure(cells * sizeof(double)));
  CUDA_TRY(ctx, r->dInjectCount.ensure((cells + 1) * sizeof(int32_t)));
  CUDA_TRY(ctx, r->dSweeps.ensure(sizeof(int32_t)));
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
  return EMCGPU_OK;
}

static double *gridPtr(emcgpu_ctx *ctx, int grid) {
  DeviceRunState *r = ctx->run;
  if (grid == EMCGPU_GRID_EFIELD_Y) return r->grid[EMCGPU_GRID_EFIELD_X].as<double>() + r->geo.cells;
  if (grid == EMCGPU_GRID_EFIELD_Z) return r->grid[EMCGPU_GRID_EFIELD_X].as<double>() + 2 * (size_t)r->geo.cells;
  return r->grid[grid].as<double>();
}

int emcgpu_device_set_grid(emcgpu_ctx *ctx, int grid, const double *host) {
  if (int r = needRun(ctx)) return r;
  if (grid < 0 || grid >= EMCGPU_N_GRIDS || !host) return fail(ctx, EMCGPU_E_INVALID, "bad grid argument");
  if (grid == EMCGPU_GRID_EFIELD_Z && ctx->run->dim < 3) return fail(ctx, EMCGPU_E_INVALID, "no z field in a 2-D device");
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
  if (int r = doPoisson(ctx, equilibrium != 0, accuracyVolt, omega, resetBC != 0, sweeps)) return r;
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

int emcgpu_device_efield(emcgpu_ctx *ctx) {
  if (int r = needRun(ctx)) return r;
  return doEfield(ctx);
}

int emcgpu_device_assign(emcgpu_ctx *ctx) {
  if (int r = needRun(ctx)) return r;
  return doAssign(ctx);
}

int emcgpu_device_concentration(emcgpu_ctx *ctx) {
  if (int r = needRun(ctx)) return r;
  return doConcentration(ctx);
}

int emcgpu_device_step(emcgpu_ctx *ctx, double dt, int32_t *removedPerContact) {
  if (int r = needRun(ctx)) return r;
  if (int r = needModel(ctx)) return r;
  if (!(dt > 0)) return fail(ctx, EMCGPU_E_INVALID, "dt must be positive");
  return doStep(ctx, dt, removedPerContact);
}

int emcgpu_device_contacts(emcgpu_ctx *ctx, int32_t *netPerContact, const uint64_t *replayDraws, int64_t nReplayDraws) {
  if (int r = needRun(ctx)) return r;
  if (int r = needModel(ctx)) return r;
  return doContacts(ctx, netPerContact, replayDraws, nReplayDraws);
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
  const int nC = ctx->run->geo.nContacts;
  for (int s = 0; s < nSteps; s++) {
    if (int r = doPoisson(ctx, false, accuracyVolt, omega, resetBCFirst && s == 0, sweeps ? sweeps + s : nullptr)) return r;
    if (int r = doEfield(ctx)) return r;
    if (int r = doStep(ctx, dt, counters ? counters + (size_t)s * 2 * nC : nullptr)) return r;
    if (int r = doContacts(ctx, counters ? counters + (size_t)s * 2 * nC + nC : nullptr, nullptr, 0)) return r;
    if (int r = doAssign(ctx)) return r;
    if (int r = doConcentration(ctx)) return r;
    if (s >= nSteps - nAverage) {
      DeviceRunState *r = ctx->run;
      accumulateKernel<<<gridBlocks(r->geo.cells), 256, 0, ctx->stream>>>(
          r->geo.cells, r->grid[EMCGPU_GRID_POTENTIAL].as<const double>(), r->grid[EMCGPU_GRID_CONCENTRATION].as<const double>(),
          r->grid[EMCGPU_GRID_SUM_POTENTIAL].as<double>(), r->grid[EMCGPU_GRID_SUM_CONCENTRATION].as<double>());
      ctx->launches++;
      CUDA_TRY(ctx, cudaGetLastError());
    }
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return EMCGPU_OK;
}

} // extern "C"
