// Kernels of the device-run path: self-consistent
// particle-mesh loop on a box device with contacts.
//
//   K2 deviceStepKernel        emcBasicParticleHandler::driftScatterParticles (:76-145) incl.
//                              emcAbstractParticleHandler::driftParticle (:238-249),
//                              handleParticleAtBoundary (emcParticleDrift.hpp:42-66), default
//                              specular reflection (emcScatterHandler.hpp:172-191), NGP force
//                              gather (emcNGPScheme.hpp:51-66)
//   K3 ngpAssignKernel         emcNGPScheme::assignToMesh (:36-47)
//   K4 concentrationKernel     emcSimulationResults::updateCurrentParticleConcentrations (:98-116)
//      efieldKernel            calcEFieldAtGridPts + setEFieldBoundaryValues (emcEFieldCalculation.hpp:13-30, :58-82)
//   K5 sorKernel               emcSORSolver::calc{Equilibrium,NonEquilibrium}Potential (:49-197)
//   K6 contact kernels         emcBasicParticleHandler::handleOhmicContacts (:158-192),
//                              generateInjectedParticles (emcAbstractParticleHandler.hpp:200-216)
//
// Grids are x-fastest flat arrays (emcGrid.hpp:242-254).  The grids of the configs are
// tiny (2 121 / 12 726 points), the ensembles small (1e4 / 1.5e5 particles): these kernels
// are latency- not bandwidth-bound, so they are written for few launches and few
// synchronisations rather than for streaming throughput.
#pragma once
#include <cooperative_groups.h>

#include "emc_bulk_kernel.cuh"

namespace emc {

constexpr int kMaxContacts = 16;

// flattened emcDevice + emcSurface + emcDopingProfile (include/emcgpu.h, emcgpu_device_t)
struct DevGeometry {
  int32_t dim, nContacts;
  int32_t extent[3];
  int32_t cells;
  double spacing[3], maxPos[3];
  double thermalVoltage, debyeLength, ni, cellVolume, epsR;
  int32_t contactType[kMaxContacts];
  double contactVoltage[kMaxContacts], gateEpsOx[kMaxContacts], gateThickness[kMaxContacts], gateBarrier[kMaxContacts];
  const int32_t *region;     // [cells]
  const int8_t *faceContact; // [cells][2*dim]
  const double *doping;      // [cells], 1/m^3
  const double *dopingNorm;  // [cells], doping / Ni (emcDevice::normalizeDoping)
  const uint8_t *cellKind;   // [cells]: bit 0 reservoir contact cell (ohmic / Schottky), bit 1 has a gate face
  // plug-in variants (include/emcgpu.h): particle-mesh scheme, particle creation rules, wall mechanisms per face
  int32_t pmScheme;          // EMCGPU_PM_*
  int32_t particleKind;      // EMCGPU_PARTICLE_*
  int32_t surfaceKind[6];    // EMCGPU_SURFACE_* per face XMIN .. ZMAX
  double surfaceParam[6];
};
enum { PM_NGP = 0, PM_CIC = 1, PM_NEC = 2, PM_NEC_VWD = 3 };
enum { SURFACE_SPECULAR = 0, SURFACE_CONSTANT = 1, SURFACE_MOMENTUM = 2 };

__device__ __forceinline__ void cellCoord(const DevGeometry &g, int cell, int c[3]) {
  c[0] = cell % g.extent[0];
  c[1] = (cell / g.extent[0]) % g.extent[1];
  c[2] = g.dim > 2 ? cell / (g.extent[0] * g.extent[1]) : 0;
}
// emcSurface::getBoundaryPos (:340-349): first face, in the order XMIN XMAX YMIN YMAX ZMIN ZMAX
__device__ __forceinline__ int cellContact(const DevGeometry &g, int cell) {
  const int8_t *fc = g.faceContact + cell * 2 * g.dim;
  for (int f = 0; f < 2 * g.dim; f++)
    if (fc[f] != -2) return fc[f];
  return -1;
}
__device__ __forceinline__ int cellFirstFace(const DevGeometry &g, int cell) {
  const int8_t *fc = g.faceContact + cell * 2 * g.dim;
  for (int f = 0; f < 2 * g.dim; f++)
    if (fc[f] != -2) return f;
  return -1;
}
__device__ __forceinline__ bool cellIsOhmic(const DevGeometry &g, int cell) {
  const int c = cellContact(g, cell);
  return c >= 0 && g.contactType[c] == 0;
}
__device__ __forceinline__ bool cellIsReservoir(const DevGeometry &g, int cell) {
  const int c = cellContact(g, cell);
  return c >= 0 && g.contactType[c] != 2; // ohmic or Schottky (emcSurface.hpp:260-262)
}
// emcDevice::posToCoord (:274-281): std::round(pos / spacing), individually rounded division
__device__ __forceinline__ int posToCell(const DevGeometry &g, double x, double y, double z) {
  int cell = (int)round(__ddiv_rn(x, g.spacing[0])) + g.extent[0] * (int)round(__ddiv_rn(y, g.spacing[1]));
  if (g.dim > 2) cell += g.extent[0] * g.extent[1] * (int)round(__ddiv_rn(z, g.spacing[2]));
  return cell;
}
// lower grid point of the mesh cell a position lies in (floor(pos / spacing)) and the distance from it in cell units:
// CIC / NEC schemes.  The index is kept inside the grid so that the upper neighbours exist (the reference aborts for a
// particle exactly on the max face).
__device__ __forceinline__ int posToLowerCell(const DevGeometry &g, const double pos[3], double w[3], int c[3]) {
  int cell = 0, stride = 1;
  c[0] = c[1] = c[2] = 0;
  w[0] = w[1] = w[2] = 0.0;
  for (int i = 0; i < g.dim; i++) {
    const double q = __ddiv_rn(pos[i], g.spacing[i]);
    const double f = floor(q);
    w[i] = __dsub_rn(q, f);
    c[i] = max(0, min((int)f, g.extent[i] - 2));
    cell += stride * c[i];
    stride *= g.extent[i];
  }
  return cell;
}

// ---------------------------------------------------------------------------
// Control block of the step loop, resident in device memory.  Everything the kernels of one EMC step hand to
// each other (ensemble size, list sizes, step index, counters) lives here, so that the host never has to
// wait for a kernel between the steps of a chunk and the per-step launch sequence is the same for every step.
struct RunCtl {
  int32_t n;          // live particles
  int32_t nKept;      // survivors of this step's compaction
  int32_t nReservoir; // entries of this step's reservoir list
  int32_t toInject;   // particles the contacts inject in this step
  int32_t slot;       // index of the step inside the running chunk (row of the counter tables)
  int32_t avgFromSlot; // slots >= this one add potential / concentration to the running sums
  int32_t poissonInterval;
  int32_t capacity;   // particles the ensemble allocation can hold
  long long step;     // Philox step index of the next particle step
  long long runSteps; // steps done since configure (frozen-field sub-cycling)
  unsigned int ticket[4]; // "last block done" counters of the multi-block kernels
  int32_t removedPerContact[kMaxContacts];
  int32_t net[kMaxContacts];
};

// true in exactly one block of the grid: the one that finishes last.  Its threads see everything the other
// blocks wrote before calling this.  The ticket resets itself for the next launch.
__device__ __forceinline__ bool lastBlockDone(unsigned int *ticket) {
  __shared__ bool sLast;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) sLast = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  if (sLast) __threadfence();
  return sLast;
}

// in-place exclusive scan of a[0..n) by one block (any size that is a multiple of 32, <= 1024); returns the total.
// Every warp scans a contiguous segment of the array on its own (coalesced rows of 32, a running carry -- no block
// barrier per row), the warp totals are scanned once, a second pass adds the warp's offset: two block barriers in all.
__device__ int blockExclusiveScan(int32_t *a, int n) {
  __shared__ int sWarp[32];
  __shared__ int sTotal;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
  const int rows = (n + 32 * nWarps - 1) / (32 * nWarps); // rows of 32 elements per warp
  const int begin = min(n, warp * rows * 32), end = min(n, begin + rows * 32);
  int carry = 0;
  for (int base = begin; base < end; base += 32) {
    const int i = base + lane;
    const int v = i < end ? __ldcg(a + i) : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (i < end) a[i] = carry + incl - v;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) sWarp[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < nWarps ? sWarp[lane] : 0;
    int incl = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    sWarp[lane] = incl - w; // exclusive: elements ahead of the warp's segment
    if (lane == 31) sTotal = incl;
  }
  __syncthreads();
  const int offset = sWarp[warp];
  if (offset)
    for (int i = begin + lane; i < end; i += 32) a[i] += offset;
  __syncthreads(); // sWarp / sTotal may be reused by the caller's next scan
  return sTotal;
}

// ---------------------------------------------------------------------------
// K4: concentration from counts (emcSimulationResults::updateCurrentParticleConcentrations :98-116)
__device__ __forceinline__ double cellConcentration(const DevGeometry &G, int cell, double count) {
  int c[3];
  cellCoord(G, cell, c);
  double v = __dmul_rn(__ddiv_rn(count, G.ni), __ddiv_rn(1.0, G.cellVolume));
  for (int i = 0; i < G.dim; i++)
    if (c[i] == 0 || c[i] == G.extent[i] - 1) v = __dmul_rn(v, 2.0);
  return v;
}
__global__ void concentrationKernel(const __grid_constant__ DevGeometry G, const double *count, double *conc) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < G.cells) conc[cell] = cellConcentration(G, cell, count[cell]);
}

// E = -grad(phi), calcEField of the particle-mesh scheme.
//   NGP, CIC: calcEFieldAtGridPts (emcEFieldCalculation.hpp:13-30): Vt (phi[prev] - phi[next]) / (2 h) in the interior;
//   NEC:      calcEFieldAtEdgeMidPts (:35-53): Vt (phi[here] - phi[next]) / h in the interior;
//   both with setEFieldBoundaryValues (:58-82): on a face 0 (artificial boundary) or the inner neighbour's value (contact);
//   NEC-VWD:  examples/mosfet2D/NECSchemeVWD.hpp:82-99: forward difference everywhere, 0 on the max face.
// CG: the potential was just written by other SMs of the same launch (cluster solver): read it past L1
template <bool CG = false>
__device__ __forceinline__ void cellEField(const DevGeometry &G, int cell, const double *potIn, double *e) {
  struct Pot {
    const double *p;
    __device__ __forceinline__ double operator[](int i) const { return CG ? __ldcg(p + i) : p[i]; }
  } pot{potIn};
  int c[3];
  cellCoord(G, cell, c);
  const int stride[3] = {1, G.extent[0], G.extent[0] * G.extent[1]};
  for (int i = 0; i < G.dim; i++) {
    double v = 0.0;
    if (G.pmScheme == PM_NEC_VWD) {
      if (c[i] != G.extent[i] - 1)
        v = __ddiv_rn(__dmul_rn(__dsub_rn(pot[cell], pot[cell + stride[i]]), G.thermalVoltage), G.spacing[i]);
    } else {
      int at = cell; // the cell whose difference quotient this cell reports
      bool zero = false;
      if (c[i] == 0) {
        zero = G.faceContact[cell * 2 * G.dim + 2 * i] == -1;
        at = cell + stride[i];
      } else if (c[i] == G.extent[i] - 1) {
        zero = G.faceContact[cell * 2 * G.dim + 2 * i + 1] == -1;
        at = cell - stride[i];
      }
      if (!zero) {
        if (G.pmScheme == PM_NEC)
          v = __ddiv_rn(__dmul_rn(__dsub_rn(pot[at], pot[at + stride[i]]), G.thermalVoltage), G.spacing[i]);
        else
          v = __ddiv_rn(__dmul_rn(__dsub_rn(pot[at - stride[i]], pot[at + stride[i]]), G.thermalVoltage),
                        __dmul_rn(2.0, G.spacing[i]));
      }
    }
    e[(size_t)i * G.cells + cell] = v;
  }
}
__global__ void efieldKernel(const __grid_constant__ DevGeometry G, const double *pot, double *e) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < G.cells) cellEField(G, cell, pot, e);
}
// the same after a Poisson solve inside the step loop: skipped together with the solve (frozen-field sub-cycling)
__global__ void efieldAfterSolveKernel(const __grid_constant__ DevGeometry G, const double *pot, double *e, const RunCtl *ctl) {
  if (ctl && ctl->runSteps % ctl->poissonInterval != 0) return;
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell < G.cells) cellEField(G, cell, pot, e);
}

// ---------------------------------------------------------------------------
// K3: nearest-grid-point charge assignment (emcNGPScheme::assignToMesh :36-47).  Every add is the same
// integer-valued nrCarriers, so the fp64 sums are exact and independent of the order of the atomics.
// Warp-aggregated (__match_any_sync on the cell index) into a shared-memory copy of the grid when it fits,
// flushed with one global atomic per touched cell and CTA.  The block that finishes last turns the counts
// into concentrations, adds to the running sums when asked (updateAverageCharacteristics :87-93) and, inside
// the step loop, closes the step in the control block.
struct AssignParams {
  const double *x, *y, *z;
  double nrCarriers;
  double *count; // [cells], zeroed by the caller
  double *conc;  // [cells] or nullptr: counts only
  const double *pot;
  double *sumPot, *sumConc;
  RunCtl *ctl;
  int32_t *counters; // [slots][2][nContacts] or nullptr
  int32_t closeStep; // 1: the ensemble is ctl->nKept + ctl->toInject particles and the step ends here; 2: only the former
  int32_t useSmem;   // 1: fp64 copy of the grid in shared memory, 2: integer hits per mesh cell (NEC / NEC-VWD)
  int32_t *hits;     // [cells], zeroed by the caller (useSmem == 2)
};

constexpr int kAssignThreads = 1024; // the block that finishes last forms the concentration of the whole grid alone
__global__ void __launch_bounds__(kAssignThreads) ngpAssignKernel(const __grid_constant__ DevGeometry G, const AssignParams A) {
  asm volatile("griddepcontrol.launch_dependents;"); // the solver of the next step (chainedLaunch, emcgpu_device.cu)
  extern __shared__ double sCount[];
  const int64_t n = A.closeStep ? A.ctl->nKept + A.ctl->toInject : A.ctl->n;
  if (A.useSmem == 1) {
    for (int i = threadIdx.x; i < G.cells; i += blockDim.x) sCount[i] = 0.0;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nRounded = (n + 31) & ~int64_t(31);
  double *target = A.useSmem ? sCount : A.count;
  if (A.useSmem == 2) {
    // NEC / NEC-VWD with an integer-valued nrCarriers (the host checks it): every deposit is the same share, so a node's
    // sum is share x (particles in the up to 2^dim mesh cells it is a corner of), exact whatever the order.  Particles are
    // counted per mesh cell (one warp-aggregated integer atomic per particle instead of 2^dim fp64 atomics); the block that
    // finishes last lets every node collect its cells.
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nRounded; i += stride) {
      const bool live = i < n;
      int base = -1;
      if (live) {
        const double pos[3] = {A.x[i], A.y[i], G.dim > 2 ? A.z[i] : 0.0};
        double w[3];
        int c[3];
        base = posToLowerCell(G, pos, w, c);
      }
      const unsigned peers = __match_any_sync(0xffffffffu, base);
      if (live && lane == __ffs(peers) - 1) atomicAdd(&A.hits[base], __popc(peers));
    }
  } else if (G.pmScheme == PM_NGP) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nRounded; i += stride) {
      const bool live = i < n;
      const int cell = live ? posToCell(G, A.x[i], A.y[i], G.dim > 2 ? A.z[i] : 0.0) : -1;
      const unsigned peers = __match_any_sync(0xffffffffu, cell);
      if (live && lane == __ffs(peers) - 1) atomicAdd(&target[cell], A.nrCarriers * __popc(peers));
    }
  } else {
    // CIC (emcCICScheme.hpp:71-118: the distance from the LOWER point weights the lower point) and NEC / NEC-VWD
    // (emcNECScheme.hpp:62-96, NECSchemeVWD.hpp:39-52: equal shares) deposit on the 2^dim corners of the mesh cell.
    // NEC shares are multiples of 1/8, so those sums are exact in any order; CIC sums depend on the order of the
    // atomics in the last bits.
    const int sy = G.extent[0], sz = G.extent[0] * G.extent[1];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const double pos[3] = {A.x[i], A.y[i], G.dim > 2 ? A.z[i] : 0.0};
      double w[3];
      int c[3];
      const int base = posToLowerCell(G, pos, w, c);
      if (G.pmScheme == PM_CIC) {
        const double wX = w[0], wY = w[1], wZ = w[2], uX = __dsub_rn(1.0, wX), uY = __dsub_rn(1.0, wY), uZ = __dsub_rn(1.0, wZ);
        if (G.dim == 2) {
          atomicAdd(&target[base], __dmul_rn(__dmul_rn(wX, wY), A.nrCarriers));
          atomicAdd(&target[base + 1], __dmul_rn(__dmul_rn(uX, wY), A.nrCarriers));
          atomicAdd(&target[base + sy], __dmul_rn(__dmul_rn(wX, uY), A.nrCarriers));
          atomicAdd(&target[base + 1 + sy], __dmul_rn(__dmul_rn(uX, uY), A.nrCarriers));
        } else {
          atomicAdd(&target[base], __dmul_rn(__dmul_rn(__dmul_rn(wX, wY), wZ), A.nrCarriers));
          atomicAdd(&target[base + 1], __dmul_rn(__dmul_rn(__dmul_rn(uX, wY), wZ), A.nrCarriers));
          atomicAdd(&target[base + sy], __dmul_rn(__dmul_rn(__dmul_rn(wX, uY), wZ), A.nrCarriers));
          atomicAdd(&target[base + 1 + sy], __dmul_rn(__dmul_rn(__dmul_rn(uX, uY), wZ), A.nrCarriers));
          atomicAdd(&target[base + sz], __dmul_rn(__dmul_rn(__dmul_rn(wX, wY), uZ), A.nrCarriers));
          atomicAdd(&target[base + 1 + sz], __dmul_rn(__dmul_rn(__dmul_rn(uX, wY), uZ), A.nrCarriers));
          atomicAdd(&target[base + sy + sz], __dmul_rn(__dmul_rn(__dmul_rn(wX, uY), uZ), A.nrCarriers));
          atomicAdd(&target[base + 1 + sy + sz], __dmul_rn(__dmul_rn(__dmul_rn(uX, uY), uZ), A.nrCarriers));
        }
      } else {
        const double share = __dmul_rn(G.dim == 2 ? 0.25 : 0.125, A.nrCarriers);
        for (int dz = 0; dz < (G.dim > 2 ? 2 : 1); dz++)
          for (int dy = 0; dy < 2; dy++)
            for (int dx = 0; dx < 2; dx++) atomicAdd(&target[base + dx + dy * sy + dz * sz], share);
      }
    }
  }
  if (A.useSmem == 1) {
    __syncthreads();
    for (int i = threadIdx.x; i < G.cells; i += blockDim.x)
      if (sCount[i] != 0.0) atomicAdd(&A.count[i], sCount[i]);
  }
  const bool fromHits = A.useSmem == 2;
  if (!fromHits && !A.conc && A.closeStep != 1) return;
  if (!lastBlockDone(&A.ctl->ticket[3])) return;
  const bool average = A.closeStep == 1 && A.ctl->slot >= A.ctl->avgFromSlot;
  if (A.conc || fromHits) {
    // one block, the whole grid: kBatch cells per thread at a time, all their loads in flight before the first division
    constexpr int kBatch = 4;
    const int sy = G.extent[0], sz = G.extent[0] * G.extent[1];
    const double share = __dmul_rn(G.dim == 2 ? 0.25 : 0.125, A.nrCarriers);
    for (int cell0 = threadIdx.x; cell0 < G.cells; cell0 += kBatch * blockDim.x) {
      double count[kBatch], pot[kBatch], sumPot[kBatch], sumConc[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; k++) {
        const int cell = cell0 + k * blockDim.x;
        if (cell >= G.cells) break;
        if (fromHits) {
          // a mesh cell never starts in the last column / row / plane, so the cells reached across a row or plane end hold 0
          int h = 0;
          for (int dz = 0; dz < (G.dim > 2 ? 2 : 1); dz++)
            for (int dy = 0; dy < 2; dy++)
              for (int dx = 0; dx < 2; dx++) {
                const int b = cell - dx - dy * sy - dz * sz;
                if (b >= 0) h += __ldcg(A.hits + b);
              }
          count[k] = __dmul_rn(share, (double)h);
        } else {
          count[k] = __ldcg(A.count + cell);
        }
        if (average) {
          pot[k] = A.pot[cell];
          sumPot[k] = A.sumPot[cell];
          sumConc[k] = A.sumConc[cell];
        }
      }
#pragma unroll
      for (int k = 0; k < kBatch; k++) {
        const int cell = cell0 + k * blockDim.x;
        if (cell >= G.cells) break;
        if (fromHits) A.count[cell] = count[k];
        if (!A.conc) continue;
        const double v = cellConcentration(G, cell, count[k]);
        A.conc[cell] = v;
        if (average) {
          A.sumPot[cell] = __dadd_rn(sumPot[k], pot[k]);
          A.sumConc[cell] = __dadd_rn(sumConc[k], v);
        }
      }
    }
  }
  if (A.closeStep == 1 && threadIdx.x == 0) {
    RunCtl &c = *A.ctl;
    if (A.counters)
      for (int k = 0; k < G.nContacts; k++) {
        A.counters[(c.slot * 2 + 0) * G.nContacts + k] = c.removedPerContact[k];
        A.counters[(c.slot * 2 + 1) * G.nContacts + k] = c.net[k];
      }
    for (int k = 0; k < kMaxContacts; k++) c.removedPerContact[k] = c.net[k] = 0;
    c.n = c.nKept + c.toInject;
    c.slot++;
    c.step++;
    c.runSteps++;
  }
}

// Sharded ensembles: the counts of all ranks are summed between the deposit and this kernel (one block), which then does
// what the last block of ngpAssignKernel does on a single GPU.
__global__ void __launch_bounds__(1024) concentrationCloseKernel(const __grid_constant__ DevGeometry G, const AssignParams A) {
  const bool average = A.ctl->slot >= A.ctl->avgFromSlot;
  for (int cell = threadIdx.x; cell < G.cells; cell += blockDim.x) {
    const double v = cellConcentration(G, cell, A.count[cell]);
    A.conc[cell] = v;
    if (average) {
      A.sumPot[cell] = __dadd_rn(A.sumPot[cell], A.pot[cell]);
      A.sumConc[cell] = __dadd_rn(A.sumConc[cell], v);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    RunCtl &c = *A.ctl;
    if (A.counters)
      for (int k = 0; k < G.nContacts; k++) {
        A.counters[(c.slot * 2 + 0) * G.nContacts + k] = c.removedPerContact[k];
        A.counters[(c.slot * 2 + 1) * G.nContacts + k] = c.net[k];
      }
    for (int k = 0; k < kMaxContacts; k++) c.removedPerContact[k] = c.net[k] = 0;
    c.n = c.nKept + c.toInject;
    c.slot++;
    c.step++;
    c.runSteps++;
  }
}

// ---------------------------------------------------------------------------
// K5: nonlinear SOR in the reference's own (lexicographic Gauss-Seidel) update order, as a PIPELINED
// WAVEFRONT.  The update of a cell needs the NEW values of its lower neighbours and the OLD values of its
// upper neighbours, so all cells on a hyperplane x+y(+z) = const are independent, and sweeping the
// hyperplanes in order reproduces the sequential sweep of emcSORSolver.hpp:157-196 exactly.  A sweep alone
// is a chain of nPlanes dependent stages with a handful of cells each (121 stages of <= 21 cells on the
// resistor grid) -- latency, not work.  But sweep s+1 may update plane p as soon as sweep s has finished
// plane p+1: consecutive sweeps run TWO planes apart, in place, on the one copy of the potential (in a
// stage the planes written all have one parity and the planes read the other).  S sweeps then take
// nPlanes + 2(S-1) stages instead of S * nPlanes.
//
// How many sweeps are needed is only known when a sweep ends (max |delta| <= accuracy), by which time later
// sweeps have already touched the potential.  Every update is therefore also logged into a ring of per-sweep
// snapshots in global memory (write-only, off the dependency chain); when sweep S turns out to be the
// converged one the result is read from snapshot S.  Sweeps are started in waves of W = (sweeps of the
// previous solve) + 2, so next to nothing is computed speculatively; a wave that ends unconverged is
// followed by another.  Iterates, result and sweep count are those of the sequential sweep.
// One CTA, potential resident in shared memory.
constexpr int kSorThreads = 1024;
constexpr int kSorRing = 64; // snapshots / sweeps in flight per wave

struct SorParams {
  double *pot;        // [cells] in/out
  const double *conc; // [cells] or nullptr: equilibrium solve (n = exp(phi), p = 1/n)
  double *history;    // [kSorRing][cells] snapshots
  double accuracy;    // normalised (volts / Vt)
  double omega;
  int32_t maxSweeps;
  int32_t potInSmem;
  int32_t *sweepsOut; // in: sweep count of the previous solve (wave width hint); out: of this one
  double *efield;     // [dim][cells] or nullptr: E = -grad(phi) of the result (pmScheme.calcEField), fused
  RunCtl *ctl;        // inside the step loop (or nullptr): frozen-field sub-cycling, per-step sweep counts
  int32_t *sweepsPerStep; // [slots] or nullptr
};

__global__ void __launch_bounds__(kSorThreads) sorPlanesKernel(const __grid_constant__ DevGeometry G, const SorParams S) {
  extern __shared__ double sPot[];
  __shared__ unsigned long long sErr[kSorRing]; // max |delta| per sweep in flight, as the bits of a non-negative double
  const int tid = threadIdx.x;
  if (S.ctl && S.ctl->runSteps % S.ctl->poissonInterval != 0) { // emcSimulation.hpp:116, :180-184: reuse the field
    if (tid == 0 && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = 0;
    return;
  }
  double *pot = S.potInSmem ? sPot : S.pot;
  if (S.potInSmem)
    for (int i = tid; i < G.cells; i += blockDim.x) sPot[i] = S.pot[i];
  if (tid < kSorRing) sErr[tid] = 0ull;
  // emcSORSolver.hpp:330-350
  double h[3] = {1, 1, 1}, hF[3] = {0, 0, 0};
  for (int i = 0; i < G.dim; i++) h[i] = __ddiv_rn(G.spacing[i], G.debyeLength);
  if (G.dim == 2) {
    hF[0] = __ddiv_rn(h[1], h[0]);
    hF[1] = __ddiv_rn(h[0], h[1]);
  } else {
    hF[0] = __ddiv_rn(__dmul_rn(h[1], h[2]), h[0]);
    hF[1] = __ddiv_rn(__dmul_rn(h[0], h[2]), h[1]);
    hF[2] = __ddiv_rn(__dmul_rn(h[0], h[1]), h[2]);
  }
  double hFSum = 0.0, hProd = 1.0;
  for (int i = 0; i < G.dim; i++) {
    hFSum = __dadd_rn(hFSum, hF[i]);
    hProd = __dmul_rn(hProd, h[i]);
  }
  const double twoHFSum = __dmul_rn(2.0, hFSum);
  const int stride[3] = {1, G.extent[0], G.extent[0] * G.extent[1]};
  const int ex = G.extent[0], ey = G.extent[1], ez = G.dim > 2 ? G.extent[2] : 1;
  const int nPlanes = ex + ey + ez - 2;
  const int nYZ = ey * ez;
  const int hint = *S.sweepsOut;
  int width = hint > 0 ? hint + 2 : 24;
  width = min(width, kSorRing - 1);
  __syncthreads();

  int firstSweep = 0; // of the current wave
  int converged = -1;
  while (converged < 0) {
    if (S.maxSweeps > 0) width = min(width, S.maxSweeps - firstSweep);
    const int nStages = nPlanes + 2 * (width - 1);
    for (int t = 0; t < nStages && converged < 0; t++) {
      // sweeps of the wave that are inside the grid at this stage: local index j, plane t - 2j
      const int jLo = t >= nPlanes ? (t - nPlanes + 2) / 2 : 0;
      const int jHi = min(width - 1, t / 2);
      const int items = (jHi - jLo + 1) * nYZ;
      for (int idx = tid; idx < items; idx += blockDim.x) {
        const int j = jLo + idx / nYZ, yz = idx % nYZ;
        const int y = yz % ey, z = yz / ey;
        const int x = t - 2 * j - y - z;
        if (x < 0 || x >= ex) continue;
        const int cell = x + ex * (y + ey * z);
        const unsigned kind = __ldg(G.cellKind + cell);
        if (kind & 1u) continue;
        const int sweep = firstSweep + j;
        const int c[3] = {x, y, z};
        const double cur = pot[cell];
        double p, n;
        if (S.conc) {
          p = exp(-cur);
          n = __ldg(S.conc + cell);
        } else {
          n = exp(cur);
          p = __ddiv_rn(1.0, n);
        }
        const double dop = __ldg(G.dopingNorm + cell);
        double num = __dmul_rn(hProd, __dadd_rn(__dadd_rn(__dsub_rn(p, n), dop), __dmul_rn(cur, __dadd_rn(p, n))));
        double den = __dadd_rn(twoHFSum, __dmul_rn(hProd, __dadd_rn(n, p)));
        for (int i = 0; i < G.dim; i++) {
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const bool atFace = side == 0 ? c[i] == 0 : c[i] == G.extent[i] - 1;
            if (!atFace) {
              num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? -stride[i] : stride[i])], hF[i]));
            } else {
              num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? stride[i] : -stride[i])], hF[i]));
              const int ct = (kind & 2u) ? G.faceContact[cell * 2 * G.dim + 2 * i + side] : -1;
              if (ct >= 0 && G.contactType[ct] == 2) { // gate: Robin term (:399-412)
                const double gammaOx = __ddiv_rn(G.gateEpsOx[ct], G.epsR);
                const double tOx = __ddiv_rn(G.gateThickness[ct], G.debyeLength);
                const double gF = __ddiv_rn(__dmul_rn(2.0, gammaOx), tOx);
                double Vg = __ddiv_rn(G.gateBarrier[ct], G.thermalVoltage);
                if (S.conc) Vg = __dadd_rn(Vg, __ddiv_rn(G.contactVoltage[ct], G.thermalVoltage));
                num = __dadd_rn(num, __dmul_rn(__dmul_rn(__dmul_rn(gF, Vg), hF[i]), h[i]));
                den = __dadd_rn(den, __dmul_rn(__dmul_rn(gF, hF[i]), h[i]));
              }
            }
          }
        }
        const double delta = __dmul_rn(S.omega, __dsub_rn(__ddiv_rn(num, den), cur));
        const double next = __dadd_rn(cur, delta);
        pot[cell] = next;
        S.history[(size_t)(sweep % kSorRing) * G.cells + cell] = next;
        atomicMax(&sErr[sweep % kSorRing], (unsigned long long)__double_as_longlong(fabs(delta)));
      }
      __syncthreads();
      // the sweep whose last plane was this stage's
      const int jFin = t - (nPlanes - 1);
      if (jFin >= 0 && (jFin & 1) == 0 && (jFin >> 1) < width) {
        const int sweep = firstSweep + (jFin >> 1);
        const double err = __longlong_as_double((long long)sErr[sweep % kSorRing]);
        if (!(err > S.accuracy) || (S.maxSweeps > 0 && sweep + 1 >= S.maxSweeps)) converged = sweep;
      }
    }
    if (converged < 0) {
      __syncthreads();
      if (tid < kSorRing) sErr[tid] = 0ull;
      __syncthreads();
      firstSweep += width;
      width = min(2 * width, kSorRing - 1);
    }
  }
  // result = snapshot of the converged sweep (cells that are never updated keep their value)
  for (int i = tid; i < G.cells; i += blockDim.x) {
    const bool fixed = G.cellKind[i] & 1u;
    if (!fixed)
      S.pot[i] = S.history[(size_t)(converged % kSorRing) * G.cells + i];
    else if (S.potInSmem)
      S.pot[i] = sPot[i];
  }
  if (tid == 0) {
    *S.sweepsOut = converged + 1;
    if (S.ctl && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = converged + 1;
  }
  if (S.efield) {
    __syncthreads();
    for (int i = tid; i < G.cells; i += blockDim.x) cellEField(G, i, S.pot, S.efield);
  }
}

// The same pipelined wavefront with a FIXED assignment of work to threads and the update of a cell split
// over two threads of different warps.  One grid row (y, z) of one sweep of the wave is walked along x, one
// cell per stage, by a PREPARER and a FINISHER:
//   preparer, one stage ahead: old value of the cell (final by then), exp(phi), charge term, denominator;
//   finisher: neighbour sums in the reference's order, division, relaxation, store, error maximum (kept in
//             a register until the row ends).
// Warps issue in order, so within one thread the two halves would simply add up; on different warps they
// overlap and a stage costs the longer of the two.  Hand-over through a double-buffered shared-memory
// record per row.  Used when a wave of rows fits the CTA (IPT rows per thread pair); sorPlanesKernel is the
// general form.
struct SorGate { // Robin term of a gate face (emcSORSolver.hpp:399-412)
  double numTerm, denTerm;
};
__device__ __forceinline__ SorGate sorGateTerm(const DevGeometry &G, int ct, bool nonEquilibrium, double hFi, double hi) {
  const double gammaOx = __ddiv_rn(G.gateEpsOx[ct], G.epsR);
  const double tOx = __ddiv_rn(G.gateThickness[ct], G.debyeLength);
  const double gF = __ddiv_rn(__dmul_rn(2.0, gammaOx), tOx);
  double Vg = __ddiv_rn(G.gateBarrier[ct], G.thermalVoltage);
  if (nonEquilibrium) Vg = __dadd_rn(Vg, __ddiv_rn(G.contactVoltage[ct], G.thermalVoltage));
  SorGate g;
  g.numTerm = __dmul_rn(__dmul_rn(__dmul_rn(gF, Vg), hFi), hi);
  g.denTerm = __dmul_rn(__dmul_rn(gF, hFi), hi);
  return g;
}

constexpr int kSorPairs = kSorThreads / 2; // finisher threads [0, kSorPairs), preparers behind them

// shared memory: potential (cells doubles, when it fits) followed by the hand-over records
template <int IPT> constexpr size_t sorRowsHandoverBytes() { return (size_t)2 * IPT * kSorPairs * (3 * sizeof(double) + sizeof(uint32_t)); }

template <int IPT, int DIM>
__global__ void __launch_bounds__(kSorThreads) sorRowsKernel(const __grid_constant__ DevGeometry G, const SorParams S) {
  extern __shared__ double sPot[];
  __shared__ unsigned long long sErr[kSorRing];
  const int tid = threadIdx.x;
  if (S.ctl && S.ctl->runSteps % S.ctl->poissonInterval != 0) {
    if (tid == 0 && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = 0;
    return;
  }
  constexpr int kItems = IPT * kSorPairs;
  double *pot = S.potInSmem ? sPot : S.pot;
  double *hand = sPot + (S.potInSmem ? G.cells : 0); // [2][3][kItems] cur, a, den; then [2][kItems] kind
  uint32_t *handKind = reinterpret_cast<uint32_t *>(hand + 2 * 3 * kItems);
  if (S.potInSmem)
    for (int i = tid; i < G.cells; i += blockDim.x) sPot[i] = S.pot[i];
  if (tid < kSorRing) sErr[tid] = 0ull;
  double h[3] = {1, 1, 1}, hF[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < DIM; i++) h[i] = __ddiv_rn(G.spacing[i], G.debyeLength);
  if (DIM == 2) {
    hF[0] = __ddiv_rn(h[1], h[0]);
    hF[1] = __ddiv_rn(h[0], h[1]);
  } else {
    hF[0] = __ddiv_rn(__dmul_rn(h[1], h[2]), h[0]);
    hF[1] = __ddiv_rn(__dmul_rn(h[0], h[2]), h[1]);
    hF[2] = __ddiv_rn(__dmul_rn(h[0], h[1]), h[2]);
  }
  double hFSum = 0.0, hProd = 1.0;
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    hFSum = __dadd_rn(hFSum, hF[i]);
    hProd = __dmul_rn(hProd, h[i]);
  }
  const double twoHFSum = __dmul_rn(2.0, hFSum);
  const int ex = G.extent[0], ey = G.extent[1], ez = DIM > 2 ? G.extent[2] : 1;
  const int strideY = ex, strideZ = ex * ey;
  const int nPlanes = ex + ey + ez - 2;
  const int nYZ = ey * ez;
  const bool nonEq = S.conc != nullptr;
  const bool preparer = tid >= kSorPairs;
  const int lane0 = preparer ? tid - kSorPairs : tid;
  const int hint = *S.sweepsOut;
  int width = hint > 0 ? hint + 2 : 24;
  const int widthCap = min(kSorRing - 1, max(1, kItems / nYZ));
  width = min(width, widthCap);
  __syncthreads();

  int firstSweep = 0, converged = -1;
  while (converged < 0) {
    if (S.maxSweeps > 0) width = min(width, S.maxSweeps - firstSweep);
    int planeOff[IPT], rowBase[IPT], rowY[IPT], rowZ[IPT], slot[IPT];
    double err[IPT];
    bool live[IPT];
#pragma unroll
    for (int q = 0; q < IPT; q++) {
      const int it = lane0 + q * kSorPairs;
      live[q] = it < width * nYZ;
      const int j = it / nYZ, yz = it % nYZ;
      rowY[q] = yz % ey;
      rowZ[q] = yz / ey;
      planeOff[q] = 2 * j + rowY[q] + rowZ[q];
      rowBase[q] = strideY * rowY[q] + strideZ * rowZ[q];
      slot[q] = (firstSweep + j) % kSorRing;
      err[q] = 0.0;
    }
    const int nStages = nPlanes + 2 * (width - 1);
    for (int t = -1; t < nStages && converged < 0; t++) {
      if (preparer) {
        // cell x of stage t + 1: its old value is final now
        double *out = hand + (size_t)((t + 1) & 1) * 3 * kItems;
        uint32_t *outKind = handKind + (size_t)((t + 1) & 1) * kItems;
#pragma unroll
        for (int q = 0; q < IPT; q++) {
          const int x = t + 1 - planeOff[q];
          if (!live[q] || x < 0 || x >= ex) continue;
          const int it = lane0 + q * kSorPairs;
          const int cell = rowBase[q] + x;
          const unsigned kind = __ldg(G.cellKind + cell);
          outKind[it] = kind;
          if (kind & 1u) continue;
          const double cur = pot[cell];
          double p, n;
          if (nonEq) {
            p = exp(-cur);
            n = __ldg(S.conc + cell);
          } else {
            n = exp(cur);
            p = __ddiv_rn(1.0, n);
          }
          const double dop = __ldg(G.dopingNorm + cell);
          const double a = __dmul_rn(hProd, __dadd_rn(__dadd_rn(__dsub_rn(p, n), dop), __dmul_rn(cur, __dadd_rn(p, n))));
          double den = __dadd_rn(twoHFSum, __dmul_rn(hProd, __dadd_rn(n, p)));
          if (kind & 2u) { // gate faces add to the denominator in face order
            const int c[3] = {x, rowY[q], rowZ[q]};
            const int last[3] = {ex - 1, ey - 1, ez - 1};
#pragma unroll
            for (int i = 0; i < DIM; i++)
#pragma unroll
              for (int side = 0; side < 2; side++) {
                if (!(side == 0 ? c[i] == 0 : c[i] == last[i])) continue;
                const int ct = G.faceContact[cell * 2 * DIM + 2 * i + side];
                if (ct >= 0 && G.contactType[ct] == 2) den = __dadd_rn(den, sorGateTerm(G, ct, nonEq, hF[i], h[i]).denTerm);
              }
          }
          out[it] = cur;
          out[kItems + it] = a;
          out[2 * kItems + it] = den;
        }
      } else if (t >= 0) {
        const double *in = hand + (size_t)(t & 1) * 3 * kItems;
        const uint32_t *inKind = handKind + (size_t)(t & 1) * kItems;
#pragma unroll
        for (int q = 0; q < IPT; q++) {
          const int x = t - planeOff[q];
          if (!live[q] || x < 0 || x >= ex) continue;
          const int it = lane0 + q * kSorPairs;
          const int cell = rowBase[q] + x;
          const unsigned kind = inKind[it];
          if (!(kind & 1u)) {
            const double cur = in[it], den = in[2 * kItems + it];
            double num = in[kItems + it];
            const int c[3] = {x, rowY[q], rowZ[q]};
            const int last[3] = {ex - 1, ey - 1, ez - 1};
            const int stride[3] = {1, strideY, strideZ};
#pragma unroll
            for (int i = 0; i < DIM; i++) {
#pragma unroll
              for (int side = 0; side < 2; side++) {
                const bool atFace = side == 0 ? c[i] == 0 : c[i] == last[i];
                if (!atFace) {
                  num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? -stride[i] : stride[i])], hF[i]));
                } else {
                  num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? stride[i] : -stride[i])], hF[i]));
                  if (kind & 2u) {
                    const int ct = G.faceContact[cell * 2 * DIM + 2 * i + side];
                    if (ct >= 0 && G.contactType[ct] == 2) num = __dadd_rn(num, sorGateTerm(G, ct, nonEq, hF[i], h[i]).numTerm);
                  }
                }
              }
            }
            const double delta = __dmul_rn(S.omega, __dsub_rn(__ddiv_rn(num, den), cur));
            const double next = __dadd_rn(cur, delta);
            err[q] = fmax(err[q], fabs(delta));
            pot[cell] = next;
            S.history[(size_t)slot[q] * G.cells + cell] = next;
          }
          if (x == ex - 1) atomicMax(&sErr[slot[q]], (unsigned long long)__double_as_longlong(err[q]));
        }
      }
      __syncthreads();
      const int jFin = t - (nPlanes - 1);
      if (jFin >= 0 && (jFin & 1) == 0 && (jFin >> 1) < width) {
        const int sweep = firstSweep + (jFin >> 1);
        const double e = __longlong_as_double((long long)sErr[sweep % kSorRing]);
        if (!(e > S.accuracy) || (S.maxSweeps > 0 && sweep + 1 >= S.maxSweeps)) converged = sweep;
      }
    }
    if (converged < 0) {
      __syncthreads();
      if (tid < kSorRing) sErr[tid] = 0ull;
      __syncthreads();
      firstSweep += width;
      width = min(2 * width, widthCap);
    }
  }
  for (int i = tid; i < G.cells; i += blockDim.x) {
    const bool fixed = G.cellKind[i] & 1u;
    if (!fixed)
      S.pot[i] = S.history[(size_t)(converged % kSorRing) * G.cells + i];
    else if (S.potInSmem)
      S.pot[i] = sPot[i];
  }
  if (tid == 0) {
    *S.sweepsOut = converged + 1;
    if (S.ctl && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = converged + 1;
  }
  if (S.efield) {
    __syncthreads();
    for (int i = tid; i < G.cells; i += blockDim.x) cellEField(G, i, S.pot, S.efield);
  }
}

// Red-black ordering of the same relaxation (BASELINE.json north_star: "device-resident red-black SOR"): the cells
// with even x+y+z are updated from the old values of their (odd) neighbours, then the odd ones from the new even
// values.  Same equation, same relaxation factor, same stopping rule (max |delta| of a full sweep <= accuracy),
// but a different -- order-independent, fully parallel -- sequence of iterates than the reference's lexicographic
// sweep: the converged potential agrees with it to about the accuracy of the solver (tests), not bit for bit.
// Opt-in (emcgpu_set_option "sor_order" = 1); two barriers per sweep instead of a chain of nPlanes stages.
template <int DIM>
__global__ void __launch_bounds__(kSorThreads) sorRedBlackKernel(const __grid_constant__ DevGeometry G, const SorParams S) {
  extern __shared__ double sPot[];
  __shared__ double sWarpErr[kSorThreads / 32];
  __shared__ double sMax;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (S.ctl && S.ctl->runSteps % S.ctl->poissonInterval != 0) {
    if (tid == 0 && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = 0;
    return;
  }
  double *pot = S.potInSmem ? sPot : S.pot;
  if (S.potInSmem)
    for (int i = tid; i < G.cells; i += blockDim.x) sPot[i] = S.pot[i];
  double h[3] = {1, 1, 1}, hF[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < DIM; i++) h[i] = __ddiv_rn(G.spacing[i], G.debyeLength);
  if (DIM == 2) {
    hF[0] = __ddiv_rn(h[1], h[0]);
    hF[1] = __ddiv_rn(h[0], h[1]);
  } else {
    hF[0] = __ddiv_rn(__dmul_rn(h[1], h[2]), h[0]);
    hF[1] = __ddiv_rn(__dmul_rn(h[0], h[2]), h[1]);
    hF[2] = __ddiv_rn(__dmul_rn(h[0], h[1]), h[2]);
  }
  double hFSum = 0.0, hProd = 1.0;
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    hFSum = __dadd_rn(hFSum, hF[i]);
    hProd = __dmul_rn(hProd, h[i]);
  }
  const double twoHFSum = __dmul_rn(2.0, hFSum);
  const int ex = G.extent[0], ey = G.extent[1], ez = DIM > 2 ? G.extent[2] : 1;
  const int halfX = (ex + 1) / 2; // cells of one colour per row, at most
  const int nHalf = halfX * ey * ez;
  const bool nonEq = S.conc != nullptr;
  __syncthreads();
  int sweeps = 0;
  for (;;) {
    double myErr = 0.0;
#pragma unroll 1
    for (int colour = 0; colour < 2; colour++) {
      for (int idx = tid; idx < nHalf; idx += blockDim.x) {
        const int row = idx / halfX, y = row % ey, z = row / ey;
        const int x = 2 * (idx - row * halfX) + ((y + z + colour) & 1);
        if (x >= ex) continue;
        const int cell = x + ex * row;
        const unsigned kind = __ldg(G.cellKind + cell);
        if (kind & 1u) continue;
        const double cur = pot[cell];
        double p, n;
        if (nonEq) {
          p = exp(-cur);
          n = __ldg(S.conc + cell);
        } else {
          n = exp(cur);
          p = __ddiv_rn(1.0, n);
        }
        const double dop = __ldg(G.dopingNorm + cell);
        double num = __dmul_rn(hProd, __dadd_rn(__dadd_rn(__dsub_rn(p, n), dop), __dmul_rn(cur, __dadd_rn(p, n))));
        double den = __dadd_rn(twoHFSum, __dmul_rn(hProd, __dadd_rn(n, p)));
        const int c[3] = {x, y, z};
        const int last[3] = {ex - 1, ey - 1, ez - 1};
        const int stride[3] = {1, ex, ex * ey};
#pragma unroll
        for (int i = 0; i < DIM; i++) {
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const bool atFace = side == 0 ? c[i] == 0 : c[i] == last[i];
            if (!atFace) {
              num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? -stride[i] : stride[i])], hF[i]));
            } else {
              num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? stride[i] : -stride[i])], hF[i]));
              if (kind & 2u) {
                const int ct = G.faceContact[cell * 2 * DIM + 2 * i + side];
                if (ct >= 0 && G.contactType[ct] == 2) {
                  const SorGate g = sorGateTerm(G, ct, nonEq, hF[i], h[i]);
                  num = __dadd_rn(num, g.numTerm);
                  den = __dadd_rn(den, g.denTerm);
                }
              }
            }
          }
        }
        const double delta = __dmul_rn(S.omega, __dsub_rn(__ddiv_rn(num, den), cur));
        pot[cell] = __dadd_rn(cur, delta);
        myErr = fmax(myErr, fabs(delta));
      }
      __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) myErr = fmax(myErr, __shfl_xor_sync(0xffffffffu, myErr, o));
    if (lane == 0) sWarpErr[warp] = myErr;
    __syncthreads();
    if (warp == 0) {
      double m = lane < (int)(blockDim.x >> 5) ? sWarpErr[lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) sMax = m;
    }
    __syncthreads();
    sweeps++;
    if (!(sMax > S.accuracy) || (S.maxSweeps > 0 && sweeps >= S.maxSweeps)) break;
  }
  if (S.potInSmem)
    for (int i = tid; i < G.cells; i += blockDim.x) S.pot[i] = sPot[i];
  if (tid == 0) {
    *S.sweepsOut = sweeps;
    if (S.ctl && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = sweeps;
  }
  if (S.efield) {
    __syncthreads();
    for (int i = tid; i < G.cells; i += blockDim.x) cellEField(G, i, S.pot, S.efield);
  }
}

// Red-black relaxation spread over a thread-block CLUSTER: the grid rows (y, z) are dealt out in contiguous bands to the
// CTAs of one cluster (8 SMs), each CTA keeps its band of the potential in its own shared memory and reads the
// neighbour rows of the adjacent bands through distributed shared memory; a cluster barrier separates the two colours
// and the per-sweep error maxima are exchanged through DSMEM as well.  Same iterates as sorRedBlackKernel (a colour only
// reads the other colour, so the order inside a half sweep does not matter), but a half sweep costs one pass of
// 1/8 of the cells plus a ~0.2 us cluster barrier instead of a single-SM pass over all of them.
constexpr int kSorClusterSize = 8;
constexpr int kSorClusterThreads = 1024;

template <int DIM>
__global__ void __launch_bounds__(kSorClusterThreads) sorRedBlackClusterKernel(const __grid_constant__ DevGeometry G, const SorParams S) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ double sBand[]; // [rowsPerCta][ex] potential, then electron density, normalised doping, cell kinds
  __shared__ double sWarpErr[kSorClusterThreads / 32];
  __shared__ double sCtaErr; // this CTA's max |delta| of the sweep, read by the whole cluster
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cta = (int)cluster.block_rank(), nCta = (int)cluster.num_blocks();
  if (S.ctl && S.ctl->runSteps % S.ctl->poissonInterval != 0) { // uniform over the cluster
    if (cta == 0 && tid == 0 && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = 0;
    return;
  }
  double h[3] = {1, 1, 1}, hF[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < DIM; i++) h[i] = __ddiv_rn(G.spacing[i], G.debyeLength);
  if (DIM == 2) {
    hF[0] = __ddiv_rn(h[1], h[0]);
    hF[1] = __ddiv_rn(h[0], h[1]);
  } else {
    hF[0] = __ddiv_rn(__dmul_rn(h[1], h[2]), h[0]);
    hF[1] = __ddiv_rn(__dmul_rn(h[0], h[2]), h[1]);
    hF[2] = __ddiv_rn(__dmul_rn(h[0], h[1]), h[2]);
  }
  double hFSum = 0.0, hProd = 1.0;
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    hFSum = __dadd_rn(hFSum, hF[i]);
    hProd = __dmul_rn(hProd, h[i]);
  }
  const double twoHFSum = __dmul_rn(2.0, hFSum);
  const int ex = G.extent[0], ey = G.extent[1], ez = DIM > 2 ? G.extent[2] : 1;
  const int nRows = ey * ez;
  const int rowsPerCta = (nRows + nCta - 1) / nCta;
  const int row0 = cta * rowsPerCta, row1 = min(nRows, row0 + rowsPerCta);
  const int myRows = max(0, row1 - row0);
  const bool nonEq = S.conc != nullptr;
  // everything a cell update reads besides the neighbours stays in shared memory as well: cluster barriers invalidate
  // the L1 cache, so per-sweep global loads would go to L2 every time
  double *sConc = sBand + (size_t)rowsPerCta * ex, *sDop = sConc + (size_t)rowsPerCta * ex;
  unsigned char *sKind = reinterpret_cast<unsigned char *>(sDop + (size_t)rowsPerCta * ex);
  for (int i = tid; i < myRows * ex; i += blockDim.x) {
    sBand[i] = S.pot[row0 * ex + i];
    sConc[i] = nonEq ? S.conc[row0 * ex + i] : 0.0;
    sDop[i] = G.dopingNorm[row0 * ex + i];
    sKind[i] = G.cellKind[row0 * ex + i];
  }
  // the row above / below a band edge lives in the neighbouring CTA (its last / first row)
  const double *bandBelow = cta > 0 ? cluster.map_shared_rank(sBand, cta - 1) + (size_t)(rowsPerCta - 1) * ex : sBand;
  const double *bandAbove = cta + 1 < nCta ? cluster.map_shared_rank(sBand, cta + 1) : sBand;
  // value of any cell (3-D: the z neighbours are ey rows away, possibly several bands)
  auto potAt = [&](int row, int x) -> double {
    const int owner = row / rowsPerCta;
    const double *band = owner == cta ? sBand : cluster.map_shared_rank(sBand, owner);
    return band[(row - owner * rowsPerCta) * ex + x];
  };
  // fixed thread -> (column pair, row) assignment: no index arithmetic inside the sweeps
  const int halfX = (ex + 1) / 2;
  const int tx = tid % halfX, ty = tid / halfX, rowStep = max(1, (int)blockDim.x / halfX);
  const bool worker = ty < rowStep;
  const double *errOfPeer = cluster.map_shared_rank(&sCtaErr, lane % nCta);
  cluster.sync();
  int sweeps = 0;
  for (;;) {
    double myErr = 0.0;
#pragma unroll 1
    for (int colour = 0; colour < 2; colour++) {
      for (int local = ty; worker && local < myRows; local += rowStep) {
        const int row = row0 + local;
        const int y = DIM == 2 ? row : row % ey, z = DIM == 2 ? 0 : row / ey;
        const int x = 2 * tx + ((y + z + colour) & 1);
        if (x >= ex) continue;
        const int at = local * ex + x;
        const unsigned kind = sKind[at];
        if (kind & 1u) continue;
        const double cur = sBand[at];
        double p, n;
        if (nonEq) {
          p = exp(-cur);
          n = sConc[at];
        } else {
          n = exp(cur);
          p = __ddiv_rn(1.0, n);
        }
        double num = __dmul_rn(hProd, __dadd_rn(__dadd_rn(__dsub_rn(p, n), sDop[at]), __dmul_rn(cur, __dadd_rn(p, n))));
        double den = __dadd_rn(twoHFSum, __dmul_rn(hProd, __dadd_rn(n, p)));
        const int c[3] = {x, y, z};
        const int last[3] = {ex - 1, ey - 1, ez - 1};
#pragma unroll
        for (int i = 0; i < DIM; i++) {
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const bool atFace = side == 0 ? c[i] == 0 : c[i] == last[i];
            const int dir = (side == 0) != atFace ? -1 : 1; // towards the neighbour, mirrored at a face
            double nb;
            if (i == 0) {
              nb = sBand[at + dir];
            } else if (i == 1) {
              const int nl = local + dir;
              nb = nl < 0 ? bandBelow[x] : nl >= myRows ? bandAbove[x] : sBand[nl * ex + x];
            } else {
              nb = potAt(row + dir * ey, x);
            }
            num = __dadd_rn(num, __dmul_rn(nb, hF[i]));
            if (atFace && (kind & 2u)) {
              const int cell = x + ex * row;
              const int ct = G.faceContact[cell * 2 * DIM + 2 * i + side];
              if (ct >= 0 && G.contactType[ct] == 2) {
                const SorGate g = sorGateTerm(G, ct, nonEq, hF[i], h[i]);
                num = __dadd_rn(num, g.numTerm);
                den = __dadd_rn(den, g.denTerm);
              }
            }
          }
        }
        const double delta = __dmul_rn(S.omega, __dsub_rn(__ddiv_rn(num, den), cur));
        sBand[at] = __dadd_rn(cur, delta);
        myErr = fmax(myErr, fabs(delta));
      }
      if (colour == 1) { // publish this CTA's maximum before the barrier that ends the sweep
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) myErr = fmax(myErr, __shfl_xor_sync(0xffffffffu, myErr, o));
        if (lane == 0) sWarpErr[warp] = myErr;
        __syncthreads();
        if (warp == 0) {
          double m = sWarpErr[lane];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
          if (lane == 0) sCtaErr = m;
        }
      }
      cluster.sync();
    }
    // every warp gathers the maxima of all CTAs: lane l reads CTA l % nCta
    double err = *errOfPeer;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) err = fmax(err, __shfl_xor_sync(0xffffffffu, err, o));
    sweeps++;
    if (!(err > S.accuracy) || (S.maxSweeps > 0 && sweeps >= S.maxSweeps)) break;
  }
  for (int i = tid; i < myRows * ex; i += blockDim.x) S.pot[row0 * ex + i] = sBand[i];
  if (cta == 0 && tid == 0) {
    *S.sweepsOut = sweeps;
    if (S.ctl && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = sweeps;
  }
  // keep the bands alive until every peer has read the last error value
  cluster.sync();
}

// The cluster relaxation for the two-dimensional device-run grids, where every thread owns at most ONE cell per colour
// (rows per CTA <= THREADS / ceil(ex / 2)).  Same bands, same iterates as sorRedBlackClusterKernel (the operations of a
// cell update in the same order: bit-identical potentials and sweep counts), but everything an update needs besides
// the potentials is fixed before the sweeps start: the shared-memory address of the cell, the addresses of the row
// neighbours (mirrored at the faces; rows of the adjacent bands as shared::cluster addresses of the peer CTA), the Robin
// term of a gate face (per thread and colour in shared memory).  A sweep is then 7 shared-memory loads, exp, ~20 FP64
// operations, one division and a store per cell; the stopping test is a __syncthreads_or whose result every CTA stores
// into every peer's shared memory before the barrier that ends the sweep.
// Launched as 16 CTAs of 512 threads (non-portable cluster size) where the device places such a cluster, else 8 of 1024.
constexpr int kSorClusterSizeWide = 16;
constexpr int kSorClusterThreadsWide = 512;
__host__ __device__ inline size_t sorClusterFastSmemBytes(int rowsPerCta, int ex, int threads) {
  const size_t band = ((size_t)rowsPerCta * ex * (3 * sizeof(double) + 1) + 15) & ~size_t(15);
  return band + (size_t)8 * threads * sizeof(double) + 16; // + [colour][x face: num, den | y face: num, den][thread] gate terms
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) sorRedBlackClusterFastKernel(const __grid_constant__ DevGeometry G, const SorParams S) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ double sBand[]; // [rowsPerCta][ex] potential, electron density, normalised doping, cell kinds; gate terms
  __shared__ int sOver[kSorClusterSizeWide];
  const int tid = threadIdx.x;
  const int cta = (int)cluster.block_rank(), nCta = (int)cluster.num_blocks();
  // the solver keeps 16 (8) SMs busy: a kernel behind it that was launched as a programmatic dependent (the particle step of
  // emcgpu_device_run*) may take the other SMs now and do what does not depend on this solve; it waits for the end of this grid
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory"); // itself a programmatic dependent of the charge assignment of the last step
  if (S.ctl && S.ctl->runSteps % S.ctl->poissonInterval != 0) { // uniform over the cluster
    if (cta == 0 && tid == 0 && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = 0;
    return;
  }
  const double h0 = __ddiv_rn(G.spacing[0], G.debyeLength), h1 = __ddiv_rn(G.spacing[1], G.debyeLength);
  const double hF0 = __ddiv_rn(h1, h0), hF1 = __ddiv_rn(h0, h1);
  const double hProd = __dmul_rn(__dmul_rn(1.0, h0), h1);
  const double twoHFSum = __dmul_rn(2.0, __dadd_rn(__dadd_rn(0.0, hF0), hF1));
  const int ex = G.extent[0], ey = G.extent[1];
  const int rowsPerCta = (ey + nCta - 1) / nCta;
  const int row0 = cta * rowsPerCta, row1 = min(ey, row0 + rowsPerCta);
  const int myRows = max(0, row1 - row0);
  const bool nonEq = S.conc != nullptr;
  const int cellsPerBand = rowsPerCta * ex;
  double *sConc = sBand + cellsPerBand, *sDop = sConc + cellsPerBand;
  unsigned char *sKind = reinterpret_cast<unsigned char *>(sDop + cellsPerBand);
  double *sGate = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(sBand) + (((size_t)cellsPerBand * 25 + 15) & ~size_t(15)));
  for (int i = tid; i < myRows * ex; i += THREADS) {
    sBand[i] = S.pot[row0 * ex + i];
    sConc[i] = nonEq ? S.conc[row0 * ex + i] : 0.0;
    sDop[i] = G.dopingNorm[row0 * ex + i];
    sKind[i] = G.cellKind[row0 * ex + i];
  }
  __syncthreads();
  const int halfX = (ex + 1) / 2;
  const int tx = tid % halfX, local = tid / halfX; // the thread's row of the band
  const uint32_t band = (uint32_t)__cvta_generic_to_shared(sBand);
  const uint32_t concOff = 8u * (uint32_t)cellsPerBand, dopOff = 2u * concOff;
  const uint32_t gate = (uint32_t)__cvta_generic_to_shared(sGate) + 8u * (uint32_t)tid;
  auto mapa = [](uint32_t addr, int rank) -> uint32_t {
    uint32_t r;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
  };
  // per colour c, bits 8c...: 0 the cell is relaxed, 1 left neighbour mirrored (x = 0), 2 right neighbour mirrored
  // (x = ex - 1), 3 the cell has a gate face, 4-7 a gate term follows the neighbour term of face x-, x+, y-, y+ (ex, ey > 1:
  // at most one gate face per axis)
  uint32_t flags = 0;
  uint32_t at0 = 0, at1 = 0, down0 = 0, down1 = 0, up0 = 0, up1 = 0;
  auto setUp = [&](int colour, uint32_t &at, uint32_t &down, uint32_t &up) {
    const int y = row0 + local;
    const int x = 2 * tx + ((y + colour) & 1);
    if (local >= myRows || x >= ex) return;
    const int cellLocal = local * ex + x;
    const unsigned kind = sKind[cellLocal];
    if (kind & 1u) return;
    uint32_t f = 1u;
    if (x == 0) f |= 2u;
    if (x == ex - 1) f |= 4u;
    at = band + 8u * (uint32_t)cellLocal;
    // the row below / above: mirrored at the faces of the device; outside the band it is the last row of the CTA below or
    // the first row of the CTA above
    auto rowAddr = [&](int nl) -> uint32_t {
      if (nl < 0) return mapa(band + 8u * (uint32_t)((rowsPerCta - 1) * ex + x), cta - 1);
      if (nl >= myRows) return mapa(band + 8u * (uint32_t)x, cta + 1);
      return mapa(band + 8u * (uint32_t)(nl * ex + x), cta);
    };
    down = rowAddr(local + (y == 0 ? 1 : -1));
    up = rowAddr(local + (y == ey - 1 ? -1 : 1));
    if (kind & 2u) { // a cell on a contact face: Robin term of a gate (emcSORSolver.hpp:399-412)
      const int cell = x + ex * y;
      const bool atFace[4] = {x == 0, x == ex - 1, y == 0, y == ey - 1};
      for (int face = 0; face < 4; face++) {
        if (!atFace[face]) continue;
        const int ct = G.faceContact[cell * 4 + face];
        if (ct >= 0 && G.contactType[ct] == 2) {
          const SorGate g = sorGateTerm(G, ct, nonEq, face < 2 ? hF0 : hF1, face < 2 ? h0 : h1);
          sGate[(4 * colour + (face < 2 ? 0 : 2)) * THREADS + tid] = g.numTerm;
          sGate[(4 * colour + (face < 2 ? 1 : 3)) * THREADS + tid] = g.denTerm;
          f |= 8u | (16u << face);
        }
      }
    }
    flags |= f << (8 * colour);
  };
  setUp(0, at0, down0, up0);
  setUp(1, at1, down1, up1);
  uint32_t overOfPeer = 0; // thread t < nCta: the address of sOver[cta] in CTA t
  if (tid < nCta) overOfPeer = mapa((uint32_t)__cvta_generic_to_shared(&sOver[cta]), tid);
  auto ldShared = [](uint32_t addr) -> double {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
  };
  auto ldCluster = [](uint32_t addr) -> double {
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
  };
  const double omega = S.omega, accuracy = S.accuracy;
  // one cell update; the order of the operations is that of sorRedBlackKernel / sorRedBlackClusterKernel
  auto relax = [&](uint32_t fl, uint32_t at, uint32_t down, uint32_t up, uint32_t gateAt) -> bool {
    const double cur = ldShared(at);
    const double nbL = ldShared((fl & 2u) ? at + 8u : at - 8u);
    const double nbR = ldShared((fl & 4u) ? at - 8u : at + 8u);
    const double nbD = ldCluster(down), nbU = ldCluster(up);
    const double dop = ldShared(at + dopOff);
    double p, n;
    if (nonEq) {
      p = exp(-cur);
      n = ldShared(at + concOff);
    } else {
      n = exp(cur);
      p = __ddiv_rn(1.0, n);
    }
    double num = __dmul_rn(hProd, __dadd_rn(__dadd_rn(__dsub_rn(p, n), dop), __dmul_rn(cur, __dadd_rn(p, n))));
    double den = __dadd_rn(twoHFSum, __dmul_rn(hProd, __dadd_rn(n, p)));
    if (!(fl & 8u)) {
      num = __dadd_rn(num, __dmul_rn(nbL, hF0));
      num = __dadd_rn(num, __dmul_rn(nbR, hF0));
      num = __dadd_rn(num, __dmul_rn(nbD, hF1));
      num = __dadd_rn(num, __dmul_rn(nbU, hF1));
    } else {
      const double gxNum = ldShared(gateAt), gxDen = ldShared(gateAt + 8u * THREADS);
      const double gyNum = ldShared(gateAt + 16u * THREADS), gyDen = ldShared(gateAt + 24u * THREADS);
      const double nb[4] = {nbL, nbR, nbD, nbU};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        num = __dadd_rn(num, __dmul_rn(nb[k], k < 2 ? hF0 : hF1));
        if (fl & (16u << k)) {
          num = __dadd_rn(num, k < 2 ? gxNum : gyNum);
          den = __dadd_rn(den, k < 2 ? gxDen : gyDen);
        }
      }
    }
    const double delta = __dmul_rn(omega, __dsub_rn(__ddiv_rn(num, den), cur));
    const double next = __dadd_rn(cur, delta);
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(at), "d"(next) : "memory");
    return fabs(delta) > accuracy;
  };
  cluster.sync();
  int sweeps = 0;
  for (;;) {
    bool over = false;
    if (flags & 1u) over = relax(flags & 0xffu, at0, down0, up0, gate);
    cluster.sync();
    if (flags & 0x100u) over |= relax((flags >> 8) & 0xffu, at1, down1, up1, gate + 32u * THREADS);
    const int any = __syncthreads_or(over ? 1 : 0);
    if (tid < nCta) asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(overOfPeer), "r"(any) : "memory");
    cluster.sync();
    int anyOfAll = 0;
    for (int k = 0; k < nCta; k++) anyOfAll |= sOver[k];
    sweeps++;
    if (!anyOfAll || (S.maxSweeps > 0 && sweeps >= S.maxSweeps)) break;
  }
  for (int i = tid; i < myRows * ex; i += THREADS) S.pot[row0 * ex + i] = sBand[i];
  if (cta == 0 && tid == 0) {
    *S.sweepsOut = sweeps;
    if (S.ctl && S.sweepsPerStep) S.sweepsPerStep[S.ctl->slot] = sweeps;
  }
  // no CTA leaves while a peer may still read its band (the neighbour rows of the last sweep); with the field asked for,
  // the barrier also publishes the potential every CTA has just written (device-scope fence before it)
  if (S.efield) __threadfence();
  cluster.sync();
  // the field of this CTA's rows follows the potential inside the same launch (difference quotients reach up to two rows
  // into the neighbouring bands: read from global memory, past L1)
  if (S.efield)
    for (int i = tid; i < myRows * ex; i += THREADS) cellEField<true>(G, row0 * ex + i, S.pot, S.efield);
}

// Dirichlet values at ohmic contacts (emcSORSolver.hpp:57-73, :139-155); faces in the reference's order
__global__ void sorResetBcKernel(const __grid_constant__ DevGeometry G, double *pot, int nonEquilibrium) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.cells) return;
  for (int f = 0; f < 2 * G.dim; f++) {
    const int c = G.faceContact[cell * 2 * G.dim + f];
    if (c >= 0 && G.contactType[c] == 0) {
      const double builtIn = asinh(__dmul_rn(0.5, __ddiv_rn(G.doping[cell], G.ni)));
      pot[cell] = nonEquilibrium ? __dadd_rn(__ddiv_rn(G.contactVoltage[c], G.thermalVoltage), builtIn) : builtIn;
    }
  }
}

// ---------------------------------------------------------------------------
// K2: one time step of every particle of a device run.
// Per particle it leaves a flag for the contact handling and the compaction that follow:
//   kGone (-2)  left through an ohmic contact (removed),  kFree (-1) alive,  >= 0 alive in that reservoir cell.
constexpr int32_t kGone = -2, kFree = -1;

struct DeviceStepParams {
  BulkParams P;      // ensemble, model, tables, rng (box / force / dir / n / step0 unused)
  const double *e;   // [dim][cells]
  double charge;
  int32_t *flag;     // [capacity] out
  RunCtl *ctl;       // n, step index, removedPerContact
  // not nullptr: the kernel also does the work of selectCountKernel<SELECT_RESERVOIR> -- reservoir particles per chunk of
  // kChunk consecutive particles, turned into offsets by the block that finishes last (ctl->nReservoir)
  int32_t *chunkCount;
};

// emcSurfaceScatterMechanism (SurfaceScatterMechanisms/emcSurfaceScatterMechanism.hpp): with probability pDiff the
// particle leaves the wall in a new direction (polar angle theta from the wall normal, azimuth phi), otherwise -- and
// after that in any case -- it is reflected; unlike the default reflection of the scatter handler this one only turns
// k where it points out of the device (:81-93).
//   constant (emcConstantSurfaceScatterMechanism.hpp):  pDiff = 1 - specularity, theta = asin(sqrt(r))
//   momentum dependent (emcMomentumDependentSurfaceScatterMechanism.hpp): pDiff = 1 - exp(-(2 h k_perp)^2),
//     theta from a Newton iteration on the cumulative distribution (:39-68)
__device__ __forceinline__ double surfaceSolveTheta(double r, double height, double speed) {
  const double t = 2.0 * height * speed, c = t * t;
  const double ee = exp(-c);
  double x = sqrt(r * (1.0 / (1.0 - ee) - 1.0 / c));
  double error = 1.0;
  int it = 0;
  while (error > 1e-10 && it < 10000) {
    double si, co;
    sincos(x, &si, &co);
    const double ecos = exp(-c * (co * co));
    const double esin = exp(-c * (si * si));
    const double x1 = x - (((ee - ecos) / c + si * si - r * (1.0 - (1.0 - ee) / c)) * (2.0 * esin * c * si * co * (ecos - 1.0)) /
                           (ee * (c - 1.0) - 1.0));
    error = fabs(x1 - x);
    x = x1;
    it++;
  }
  return x;
}

template <int RNG_MODE, int DIM>
__device__ __forceinline__ void surfaceScatter(const DevGeometry &G, int face, double pos[3], double k[3], Rng &rng) {
  const int kind = G.surfaceKind[face];
  const int perp = face >> 1;
  double pDiff;
  if (kind == SURFACE_CONSTANT) {
    pDiff = 1.0 - G.surfaceParam[face];
  } else {
    const double t = 2.0 * G.surfaceParam[face] * k[perp];
    pDiff = 1.0 - exp(-(t * t));
  }
  if (uniform01(rng.raw<RNG_MODE>()) < pDiff) {
    const double speed = sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);
    double theta;
    if (kind == SURFACE_CONSTANT)
      theta = asin(sqrt(uniform01(rng.raw<RNG_MODE>())));
    else
      theta = surfaceSolveTheta(uniform01(rng.raw<RNG_MODE>()), G.surfaceParam[face], speed);
    const double phi = 2.0 * 3.14159265358979323846 * uniform01(rng.raw<RNG_MODE>());
    double st, ct, sp, cp;
    sincos(theta, &st, &ct);
    sincos(phi, &sp, &cp);
    k[perp] = speed * ct * ((face & 1) ? -1.0 : 1.0);
    k[(perp + 1) % 3] = speed * st * cp;
    k[(perp + 2) % 3] = speed * st * sp;
  }
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    if (pos[i] < 0.0) {
      pos[i] = -pos[i];
      if (k[i] < 0.0) k[i] = -k[i];
    } else if (pos[i] > G.maxPos[i]) {
      pos[i] = 2.0 * G.maxPos[i] - pos[i];
      if (k[i] > 0.0) k[i] = -k[i];
    }
  }
}

template <bool EXACT, int RNG_MODE, int DIM>
__device__ __forceinline__ bool deviceDriftParticle(const DevGeometry &G, const DevModel &model, Particle &p, double dt,
                                                    const Vec3 &force, Rng &rng) {
  drift<EXACT, DIM>(model.valleys[p.valley], p, dt, force);
  double pos[3] = {p.pos.x, p.pos.y, p.pos.z};
  double k[3] = {p.k.x, p.k.y, p.k.z};
  bool out = false;
#pragma unroll
  for (int i = 0; i < DIM; i++) out = out || pos[i] < 0.0 || pos[i] > G.maxPos[i];
  bool removed = false;
  if (out) {
    double cl[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < DIM; i++) cl[i] = fmax(0.0, fmin(pos[i], G.maxPos[i]));
    const int wallCell = posToCell(G, cl[0], cl[1], cl[2]);
    const int face = cellFirstFace(G, wallCell);
    if (cellIsOhmic(G, wallCell)) {
      removed = true;
#pragma unroll
      for (int i = 0; i < DIM; i++) pos[i] = cl[i];
    } else if (face >= 0 && G.surfaceKind[face] != SURFACE_SPECULAR) {
      surfaceScatter<RNG_MODE, DIM>(G, face, pos, k, rng);
    } else {
#pragma unroll
      for (int i = 0; i < DIM; i++) {
        if (pos[i] < 0.0) {
          pos[i] = -pos[i];
          k[i] = -k[i];
        } else if (pos[i] > G.maxPos[i]) {
          pos[i] = Arith<true>::sub(Arith<true>::mul(2.0, G.maxPos[i]), pos[i]);
          k[i] = -k[i];
        }
      }
    }
  }
  p.pos = Vec3{pos[0], pos[1], pos[2]};
  p.k = Vec3{k[0], k[1], k[2]};
  if (!removed) p.region = G.region[posToCell(G, pos[0], pos[1], pos[2])];
  return removed;
}

// interpolateForce of the particle-mesh scheme: emcNGPScheme.hpp:51-66 (field of the nearest grid point),
// emcCICScheme.hpp:122-173 (corner fields weighted like the deposit -- including the reference's (1 - wY)(1 - wY)
// weight of the upper-right corner), emcNECScheme.hpp:99-113 (mean of the two edge mid-point fields of the mesh cell,
// 2-D), NECSchemeVWD.hpp:57-76 (the same with the x index rounded instead of floored).
template <int DIM>
__device__ __forceinline__ Vec3 pmForce(const DevGeometry &G, const double *e, const Particle &p, double charge) {
  Vec3 f;
  f.z = 0.0;
  const size_t n = (size_t)G.cells;
  if (G.pmScheme == PM_NGP) {
    const int cell = posToCell(G, p.pos.x, p.pos.y, p.pos.z);
    f.x = __dmul_rn(charge, e[cell]);
    f.y = __dmul_rn(charge, e[n + cell]);
    if (DIM > 2) f.z = __dmul_rn(charge, e[2 * n + cell]);
    return f;
  }
  const double pos[3] = {p.pos.x, p.pos.y, DIM > 2 ? p.pos.z : 0.0};
  double w[3];
  int c[3];
  int base = posToLowerCell(G, pos, w, c);
  const int sy = G.extent[0], sz = G.extent[0] * G.extent[1];
  if (G.pmScheme == PM_CIC) {
    const double wX = w[0], wY = w[1], wZ = w[2], uX = 1.0 - wX, uY = 1.0 - wY, uZ = 1.0 - wZ;
    double out[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      const double *E = e + (size_t)i * n;
      if (DIM == 2)
        out[i] = charge * (E[base] * wX * wY + E[base + 1] * uX * wY + E[base + sy] * wX * uY + E[base + 1 + sy] * uY * uY);
      else
        out[i] = (E[base] * wX * wY * wZ + E[base + 1] * uX * wY * wZ + E[base + sy] * wX * uY * wZ +
                  E[base + 1 + sy] * uY * uY * wZ + E[base + sz] * wX * wY * uZ + E[base + 1 + sz] * uX * wY * uZ +
                  E[base + sy + sz] * wX * uY * uZ + E[base + 1 + sy + sz] * uY * uY * uZ) *
                 charge;
    }
    f.x = out[0];
    f.y = out[1];
    f.z = out[2];
    return f;
  }
  // NEC / NEC-VWD (2-D)
  bool lastColumn = false;
  if (G.pmScheme == PM_NEC_VWD) {
    const int cx = (int)round(__ddiv_rn(pos[0], G.spacing[0]));
    base += cx - c[0];
    lastColumn = cx == G.extent[0] - 1;
  }
  f.x = charge * (e[base] + e[base + sy]) / 2.0;
  f.y = lastColumn ? charge * e[n + base] : charge * (e[n + base] + e[n + base + 1]) / 2.0;
  return f;
}

template <bool EXACT, int RNG_MODE, int DIM>
__global__ void __launch_bounds__(kBulkThreads, 2)
    deviceStepKernel(const __grid_constant__ DevGeometry G, const __grid_constant__ DeviceStepParams D) {
  asm volatile("griddepcontrol.launch_dependents;"); // the reservoir list kernel behind it (chainedLaunch, emcgpu_device.cu)
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint64_t tableBar;
  const BulkParams &P = D.P;
  const CtaState C = stageCta(P, smemRaw, &tableBar, 0, 0);
  // launched as a programmatic dependent of the kernel ahead in the stream (emcgpu_device_run*): model and tables are staged,
  // everything below reads what the earlier kernels of the step wrote -- wait for them (returns at once in a plain launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const DevModel &model = *C.model;
  using A = Arith<EXACT>;
  const int64_t n = D.ctl->n;
  const long long step = D.ctl->step;
  // a block takes whole chunks of kBulkThreads (= kChunk of the ordered selections below) consecutive particles
  const int nChunks = (int)((n + kBulkThreads - 1) / kBulkThreads);
  for (int chunk = blockIdx.x; chunk < nChunks; chunk += gridDim.x) {
    const int64_t i = (int64_t)chunk * kBulkThreads + threadIdx.x;
    bool inReservoir = false;
    if (i < n) {
    Particle p;
    Rng rng;
    loadParticle(P, i, p, rng);
    p.energy = P.stream[EMCGPU_ENERGY][i]; // Coulomb and the table look-up read it before the next drift
    if (DIM < 3) p.pos.z = 0.0;
    attachReplay<RNG_MODE>(P, i, rng);
    rng.step = (uint32_t)step;
    const double dt = P.dt;
    Vec3 force = pmForce<DIM>(G, D.e, p, D.charge);
    bool removed;
    if (p.tau >= dt) {
      removed = deviceDriftParticle<EXACT, RNG_MODE, DIM>(G, model, p, dt, force, rng);
    } else {
      removed = deviceDriftParticle<EXACT, RNG_MODE, DIM>(G, model, p, p.tau, force, rng);
      double tRem = A::sub(dt, p.tau);
      while (tRem > 0.0 && !removed) {
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
            const unsigned long long ev = atomicAdd(P.evCount, 1ull);
            if ((long long)ev < P.evCap) {
              long long *dst = P.events + 4 * ev;
              dst[0] = step;
              dst[1] = P.idBase + i;
              dst[2] = m;
              dst[3] = mechId;
            }
          }
        }
        {
          const int set2 = (p.region < kMaxRegions) ? model.setOf[p.valley][p.region] : -1;
          if (set2 >= 0) tauTab = model.sets[set2].tau;
        }
        const double newTau = A::mul(-log(uniformLog(rng.raw<RNG_MODE>())), tauTab);
        p.tau = A::add(p.tau, newTau);
        force = pmForce<DIM>(G, D.e, p, D.charge); // re-interpolated after every scattering (:112)
        removed = deviceDriftParticle<EXACT, RNG_MODE, DIM>(G, model, p, fmin(tRem, newTau), force, rng);
        tRem = A::sub(tRem, newTau);
      }
    }
    p.tau = A::sub(p.tau, dt);
    if (P.grainTau) { // emcBasicParticleHandler.hpp:134-138
      double grain = __dsub_rn(P.grainTau[i], dt);
      if (grain <= 0.0 && !removed) grain = grainEvent<RNG_MODE>(P, p, rng);
      P.grainTau[i] = grain;
    }
    storeParticle<RNG_MODE>(P, i, p, rng);
    const int cell = posToCell(G, p.pos.x, p.pos.y, p.pos.z);
    if (removed) {
      D.flag[i] = kGone;
      atomicAdd(&D.ctl->removedPerContact[cellContact(G, cell)], 1);
    } else {
      inReservoir = (G.cellKind[cell] & 1u) != 0;
      D.flag[i] = inReservoir ? cell : kFree;
    }
    }
    if (D.chunkCount) { // block-uniform
      const int count = __syncthreads_count(inReservoir);
      if (threadIdx.x == 0) D.chunkCount[chunk] = count;
    }
  }
  if (D.chunkCount) {
    if (!lastBlockDone(&D.ctl->ticket[0])) return; // the ticket of selectCountKernel<SELECT_RESERVOIR>
    const int total = blockExclusiveScan(D.chunkCount, nChunks);
    if (threadIdx.x == 0) D.ctl->nReservoir = total;
  }
}

// flags of a resting ensemble (emcgpu_device_contacts on its own): kFree or the reservoir cell
__global__ void __launch_bounds__(256)
    reservoirFlagKernel(const __grid_constant__ DevGeometry G, const double *x, const double *y, const double *z,
                        const RunCtl *ctl, int32_t *flag) {
  const int64_t n = ctl->n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cell = posToCell(G, x[i], y[i], G.dim > 2 ? z[i] : 0.0);
    flag[i] = (G.cellKind[cell] & 1u) ? cell : kFree;
  }
}

// ---------------------------------------------------------------------------
// Ordered selection over the ensemble: particles are handled in chunks of kChunk consecutive indices; a
// "count" kernel leaves the number of selected particles per chunk and its last block turns them into
// offsets, a second kernel then writes the selected particles in index order.  Used twice per step:
//   SELECT_RESERVOIR  flag >= 0     -> the list of (particle, cell) the contact handling ranks
//   SELECT_KEPT       flag != kGone -> order-preserving compaction (removeParticles, emcBasicParticleHandler.hpp:267-277)
constexpr int kChunk = 256;
enum { SELECT_RESERVOIR = 0, SELECT_KEPT = 1 };

template <int WHAT> __device__ __forceinline__ bool selected(int32_t flag) {
  return WHAT == SELECT_RESERVOIR ? flag >= 0 : flag != kGone;
}

template <int WHAT>
__global__ void __launch_bounds__(kChunk) selectCountKernel(const int32_t *flag, RunCtl *ctl, int32_t *chunkCount) {
  asm volatile("griddepcontrol.launch_dependents;"); // see chainedLaunch (emcgpu_device.cu): no-ops in a plain launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ int sWarp[kChunk / 32];
  const int64_t n = ctl->n;
  const int nChunks = (int)((n + kChunk - 1) / kChunk);
  for (int chunk = blockIdx.x; chunk < nChunks; chunk += gridDim.x) {
    const int64_t i = (int64_t)chunk * kChunk + threadIdx.x;
    const bool sel = i < n && selected<WHAT>(flag[i]);
    const unsigned b = __ballot_sync(0xffffffffu, sel);
    if ((threadIdx.x & 31) == 0) sWarp[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x == 0) {
      int sum = 0;
      for (int w = 0; w < kChunk / 32; w++) sum += sWarp[w];
      chunkCount[chunk] = sum;
    }
    __syncthreads();
  }
  if (!lastBlockDone(&ctl->ticket[WHAT])) return;
  const int total = blockExclusiveScan(chunkCount, nChunks);
  if (threadIdx.x == 0) {
    if (WHAT == SELECT_RESERVOIR)
      ctl->nReservoir = total;
    else
      ctl->nKept = total;
  }
}

__global__ void __launch_bounds__(kChunk)
    reservoirListKernel(const int32_t *flag, const RunCtl *ctl, const int32_t *chunkOffset, int32_t *listParticle,
                        int32_t *listCell) {
  asm volatile("griddepcontrol.launch_dependents;"); // see chainedLaunch (emcgpu_device.cu): no-ops in a plain launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ int sWarp[kChunk / 32];
  const int64_t n = ctl->n;
  const int nChunks = (int)((n + kChunk - 1) / kChunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int chunk = blockIdx.x; chunk < nChunks; chunk += gridDim.x) {
    const int64_t i = (int64_t)chunk * kChunk + threadIdx.x;
    const int32_t f = i < n ? flag[i] : kFree;
    const unsigned b = __ballot_sync(0xffffffffu, f >= 0);
    if (lane == 0) sWarp[warp] = __popc(b);
    __syncthreads();
    if (f >= 0) {
      int off = chunkOffset[chunk];
      for (int w = 0; w < warp; w++) off += sWarp[w];
      off += __popc(b & ((1u << lane) - 1u));
      listParticle[off] = (int32_t)i;
      listCell[off] = f;
    }
    __syncthreads();
  }
}

struct EnsemblePtrs {
  double *stream[EMCGPU_N_STREAMS];
  uint32_t *packed;
  uint32_t *cursor; // replay cursors travel with their particle (may be null)
  double *grain;    // grain clocks travel with their particle (may be null)
};

// (block `blk` of `nBlk` blocks of kChunk threads: the kernel below, or the compaction blocks of compactInjectKernel)
__device__ __forceinline__ void compactScatterRole(const int32_t *flag, const RunCtl *ctl, const int32_t *chunkOffset,
                                                   const EnsemblePtrs &src, const EnsemblePtrs &dst, int blk, int nBlk) {
  __shared__ int sWarp[kChunk / 32];
  const int64_t n = ctl->n;
  const int nChunks = (int)((n + kChunk - 1) / kChunk);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int chunk = blk; chunk < nChunks; chunk += nBlk) {
    const int64_t i = (int64_t)chunk * kChunk + threadIdx.x;
    const bool keep = i < n && flag[i] != kGone;
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) sWarp[warp] = __popc(b);
    __syncthreads();
    if (keep) {
      int off = chunkOffset[chunk];
      for (int w = 0; w < warp; w++) off += sWarp[w];
      const int64_t j = off + __popc(b & ((1u << lane) - 1u));
#pragma unroll
      for (int c = 0; c < EMCGPU_N_STREAMS; c++) dst.stream[c][j] = src.stream[c][i];
      dst.packed[j] = src.packed[i];
      if (src.cursor) dst.cursor[j] = src.cursor[i];
      if (src.grain) dst.grain[j] = src.grain[i];
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(kChunk)
    compactScatterKernel(const int32_t *flag, const RunCtl *ctl, const int32_t *chunkOffset, EnsemblePtrs src, EnsemblePtrs dst) {
  compactScatterRole(flag, ctl, chunkOffset, src, dst, (int)blockIdx.x, (int)gridDim.x);
}

// ---------------------------------------------------------------------------
// K6: ohmic contacts (emcBasicParticleHandler::handleOhmicContacts :158-192).  The reference scans the ensemble in
// index order and keeps a particle of a reservoir cell while the cell is below its expected population, so the FIRST
// particles of a cell (by index) survive: a particle is deleted iff its rank among the particles of its cell is
// >= slots = ceil(expected / carriers per particle).  The ranks are counted over the (short) ordered reservoir list
// by all blocks; the block that finishes last derives, cell by cell in storage order, how many particles each
// reservoir cell is missing (while (diff > 0) { inject; diff -= carriers }) and the per-contact bookkeeping.
struct ContactParams {
  double nrCarriers;
  const double *expected;  // [cells]
  const int32_t *listParticle, *listCell;
  int32_t *flag;           // [capacity]: excess particles become kGone
  int32_t *cellCount;      // [cells] scratch, all zero between launches
  int32_t *injectCount;    // [cells + 1] out: particles to inject per cell as an exclusive scan, total at [cells]
  RunCtl *ctl;
  // ensemble sharded over several GPUs (emcgpu_device_set_sharding): share[r][cell] = reservoir particles of rank r in
  // that cell, summed over the ranks before this kernel runs; nullptr on a single GPU
  const double *share;
  int32_t rank, world;
};

// How many of the `missing` particles of a reservoir cell this rank injects: an even split whose remainder rotates with
// the cell and the time step, so that no rank keeps collecting the injected particles (viennaemc_b200/sharding.py holds
// the same function for the host-side tests).
__host__ __device__ __forceinline__ int injectShareOfRank(int missing, int rank, int world, int cell, long long step) {
  const int rotated = (int)((rank + world - (int)((cell + step) % world)) % world);
  return missing / world + (rotated < missing % world ? 1 : 0);
}

// reservoir particles of this rank per cell -> its row of the share table
__global__ void __launch_bounds__(256) contactShareKernel(const ContactParams K, double *shareRow) {
  const int m = K.ctl->nReservoir;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) atomicAdd(&shareRow[K.listCell[j]], 1.0);
}

__device__ __forceinline__ int reservoirSlots(double expected, double nrCarriers) {
  return expected > 0.0 ? (int)ceil(expected / nrCarriers) : 0;
}

// One CTA.  The rank of a particle among the particles of its cell -- in index order -- is (particles of the cell in
// earlier tiles and earlier warps of this tile) + (earlier lanes of its warp with the same cell): per-cell counters (shared
// memory when the grid fits, else the zeroed global scratch) are advanced warp by warp in index order, lanes of a warp
// that share a cell are ranked with __match_any_sync.  O(M) instead of comparing every pair of list entries.
constexpr int kRankThreads = 1024;

__global__ void __launch_bounds__(kRankThreads) contactRankKernel(const __grid_constant__ DevGeometry G, const ContactParams K,
                                                                   const int countersInSmem) {
  asm volatile("griddepcontrol.launch_dependents;"); // see chainedLaunch (emcgpu_device.cu): no-ops in a plain launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ int sCellCount[];
  __shared__ int sGroupCell[kRankThreads], sGroupSize[kRankThreads], sGroupBefore[kRankThreads];
  int *cnt = countersInSmem ? sCellCount : K.cellCount;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (countersInSmem)
    for (int c = tid; c < G.cells; c += blockDim.x) sCellCount[c] = 0;
  __syncthreads();
  const int m = K.ctl->nReservoir;
  for (int base = 0; base < m; base += kRankThreads) {
    const int j = base + tid;
    const int cell = j < m ? K.listCell[j] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    const int leaderLane = __ffs(peers) - 1;
    // the groups (warp, cell) of this chunk take their places in the order of the warps: every group leader posts its cell
    // and size, warp 0 walks through the 32 warps' groups -- the groups of one warp have different cells, so its lanes
    // never collide -- and posts the number of particles of the cell ahead of each group
    sGroupCell[tid] = (cell >= 0 && lane == leaderLane) ? cell : -1;
    sGroupSize[tid] = __popc(peers);
    __syncthreads();
    if (warp == 0) {
      for (int w = 0; w < kRankThreads / 32; w++) {
        const int g = 32 * w + lane, c = sGroupCell[g];
        if (c >= 0) {
          const int b = cnt[c];
          cnt[c] = b + sGroupSize[g];
          sGroupBefore[g] = b;
        }
        __syncwarp();
      }
    }
    __syncthreads();
    if (cell >= 0) {
      const int before = sGroupBefore[32 * warp + leaderLane]; // particles of the cell ahead of this warp
      int rank = before + __popc(peers & ((1u << lane) - 1u));
      if (K.share) // particles of the lower ranks come first in the global index order
        for (int r = 0; r < K.rank; r++) rank += (int)K.share[(size_t)r * G.cells + cell];
      if (rank >= reservoirSlots(K.expected[cell], K.nrCarriers)) {
        K.flag[K.listParticle[j]] = kGone;
        atomicAdd(&K.ctl->net[cellContact(G, cell)], -1);
      }
    }
  }
  __syncthreads();
  for (int c = tid; c < G.cells; c += blockDim.x) {
    int n = 0;
    if (G.cellKind[c] & 1u) {
      int group = cnt[c];
      if (K.share) {
        group = 0;
        for (int r = 0; r < K.world; r++) group += (int)K.share[(size_t)r * G.cells + c];
      }
      if (!countersInSmem) cnt[c] = 0; // the global scratch stays zero between launches
      const int kept = min(group, reservoirSlots(K.expected[c], K.nrCarriers));
      const double diff = K.expected[c] - kept * K.nrCarriers;
      n = diff > 0.0 ? (int)ceil(diff / K.nrCarriers) : 0;
      if (K.share) n = injectShareOfRank(n, K.rank, K.world, c, K.ctl->step);
      if (n) atomicAdd(&K.ctl->net[cellContact(G, c)], n);
    }
    K.injectCount[c] = n;
  }
  __syncthreads();
  const int total = blockExclusiveScan(K.injectCount, G.cells);
  if (tid == 0) {
    K.injectCount[G.cells] = total;
    K.ctl->toInject = total;
  }
}

// emcBasicParticleHandler::addParticle(isInitial = false): initParticlePos (emcParticleInitialization.hpp:14-29) +
// emcElectron::generateInjectedParticle (emcElectron.hpp:92-104).  Draw order: position (dim draws), valley,
// sub-valley, energy, cos(theta), phi, tau, grainTau = dim + 7 draws per particle; injected particle j of this
// step uses Philox(seed, counter = (j, 0xC0117AC7, step, .)) or, in replay mode, draws[j * (dim + 7) ...].
// The particles are appended behind the ctl->nKept survivors of the compaction.
struct InjectParams {
  EnsemblePtrs ens;
  const int32_t *injectCount; // exclusive scan per cell, total at [cells]
  const DevModel *model;
  uint64_t seed;
  RunCtl *ctl;
  const uint64_t *replay; // flat draw stream of the contact phase or nullptr
  int64_t replayCount;
  int *status;
  uint32_t rank; // of a sharded ensemble: keeps the streams of the ranks apart
  double grainTau0; // mean time between grain events (1 s without a grain mechanism)
};

template <int DIM>
__device__ __forceinline__ void contactInjectRole(const DevGeometry &G, const InjectParams &J, int blk, int nBlk) {
  const DevModel &model = *J.model;
  const int total = J.injectCount[G.cells];
  const int64_t first = J.ctl->nKept;
  const long long step = J.ctl->step;
  if (first + total > J.ctl->capacity) { // the host reserves head room between chunks; running out of it is an error
    if (blk == 0 && threadIdx.x == 0) {
      atomicExch(J.status, (int)EMCGPU_E_CAPACITY);
      J.ctl->toInject = 0; // keeps the later kernels of the chunk inside the allocation
    }
    return;
  }
  for (int j = blk * blockDim.x + threadIdx.x; j < total; j += nBlk * blockDim.x) {
    // cell of injected particle j: the last cell whose offset is <= j (empty cells share the offset of their successor)
    int lo = 0, hi = G.cells - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (J.injectCount[mid] <= j) lo = mid; else hi = mid - 1;
    }
    const int cell = lo;
    int c[3];
    cellCoord(G, cell, c);
    constexpr int kDraws = DIM + 7;
    uint64_t raw[kDraws];
    if (J.replay) {
      for (int d = 0; d < kDraws; d++) {
        const int64_t at = (int64_t)j * kDraws + d;
        if (at >= J.replayCount) {
          atomicExch(J.status, (int)EMCGPU_E_REPLAY_EXHAUSTED);
          raw[d] = 0x8000000000000000ull;
        } else {
          raw[d] = J.replay[at];
        }
      }
    } else {
      for (int d = 0; d < kDraws; d += 2) {
        uint32_t o[4];
        philox4x32_10((uint32_t)j, 0xC0117AC7u + J.rank, (uint32_t)step, (uint32_t)(d >> 1), (uint32_t)J.seed,
                      (uint32_t)(J.seed >> 32), o);
        raw[d] = (uint64_t)o[1] << 32 | o[0];
        if (d + 1 < kDraws) raw[d + 1] = (uint64_t)o[3] << 32 | o[2];
      }
    }
    int d = 0;
    double pos[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < DIM; i++) {
      const double u = uniform01(raw[d++]);
      if (c[i] == G.extent[i] - 1)
        pos[i] = __dmul_rn(__dsub_rn((double)c[i], __dmul_rn(u, 0.5)), G.spacing[i]);
      else if (c[i] == 0)
        pos[i] = __dmul_rn(__dmul_rn(u, 0.5), G.spacing[i]);
      else
        pos[i] = __dmul_rn(__dsub_rn(__dadd_rn((double)c[i], u), 0.5), G.spacing[i]);
    }
    const int region = G.region[cell];
    // emcElectron draws valley / sub-valley from U[1e-6, 1), mosfet2D's electronVWD from U[0, 1) (electronVWD.hpp:25, :82-85)
    const bool vwd = G.particleKind == 1;
    const uint64_t rawValley = raw[d++], rawSub = raw[d++];
    const int valley = (int)floor(__dmul_rn((double)model.nValleys, vwd ? uniform01(rawValley) : uniformLog(rawValley)));
    const DevValley &v = model.valleys[valley];
    const int sub = (int)floor(__dmul_rn((double)v.deg, vwd ? uniform01(rawSub) : uniformLog(rawSub)));
    const double energy = __dmul_rn(__dmul_rn(-1.5, G.thermalVoltage), log(uniformLog(raw[d++])));
    const double r2 = uniform01(raw[d++]);
    const double r1 = uniform01(raw[d++]);
    Vec3 k = randomDirection<true>(normWaveVec<true>(v, energy), r1, r2);
    double kk[3] = {k.x, k.y, k.z};
    for (int i = 0; i < DIM; i++)
      if ((c[i] == 0 && kk[i] < 0.0) || (c[i] == G.extent[i] - 1 && kk[i] > 0.0)) kk[i] = -kk[i];
    // electronVWD passes the valley index where the region index belongs (electronVWD.hpp:87)
    const int tauRegion = vwd ? valley : region;
    const int set = (tauRegion >= 0 && tauRegion < kMaxRegions) ? model.setOf[valley][tauRegion] : -1;
    const double tau0 = set >= 0 ? model.sets[set].tau : model.defaultTau;
    const double tau = __dmul_rn(-log(uniformLog(raw[d++])), tau0);
    const int64_t at = first + j;
    J.ens.stream[EMCGPU_KX][at] = kk[0];
    J.ens.stream[EMCGPU_KY][at] = kk[1];
    J.ens.stream[EMCGPU_KZ][at] = kk[2];
    J.ens.stream[EMCGPU_ENERGY][at] = energy;
    J.ens.stream[EMCGPU_TAU][at] = tau;
    J.ens.stream[EMCGPU_X][at] = pos[0];
    J.ens.stream[EMCGPU_Y][at] = pos[1];
    J.ens.stream[EMCGPU_Z][at] = pos[2];
    J.ens.packed[at] = (uint32_t)valley | ((uint32_t)sub << 8) | ((uint32_t)region << 16);
    if (J.ens.cursor) J.ens.cursor[at] = 0;
    if (J.ens.grain) J.ens.grain[at] = __dmul_rn(-log(uniformLog(raw[d++])), J.grainTau0);
  }
}
template <int DIM>
__global__ void __launch_bounds__(128) contactInjectKernel(const __grid_constant__ DevGeometry G, const InjectParams J) {
  contactInjectRole<DIM>(G, J, (int)blockIdx.x, (int)gridDim.x);
}
// The order-preserving compaction and the injection behind the survivors in ONE launch: the two write disjoint parts of the
// new ensemble ([0, nKept) and [nKept, nKept + toInject)) and both only need the counts of the kernels before them.  The
// first compactBlocks blocks compact, the others inject (J.ens = the ensemble the compaction writes).
template <int DIM>
__global__ void __launch_bounds__(kChunk)
    compactInjectKernel(const __grid_constant__ DevGeometry G, const __grid_constant__ InjectParams J, const int32_t *flag,
                        const int32_t *chunkOffset, const __grid_constant__ EnsemblePtrs src, int compactBlocks) {
  asm volatile("griddepcontrol.launch_dependents;"); // see chainedLaunch (emcgpu_device.cu): no-ops in a plain launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if ((int)blockIdx.x < compactBlocks)
    compactScatterRole(flag, J.ctl, chunkOffset, src, J.ens, (int)blockIdx.x, compactBlocks);
  else
    contactInjectRole<DIM>(G, J, (int)blockIdx.x - compactBlocks, (int)gridDim.x - compactBlocks);
}

} // namespace emc
