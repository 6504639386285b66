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
};

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

// ---------------------------------------------------------------------------
// K3: nearest-grid-point charge assignment.  Every add is the same integer-valued
// nrCarriers, so the fp64 sums are exact and independent of the order of the atomics.
// Warp-aggregated (__match_any_sync on the cell index) into a shared-memory copy of
// the grid when it fits, flushed with one global atomic per touched cell and CTA.
struct AssignParams {
  const double *x, *y, *z;
  int64_t n;
  double nrCarriers;
  double *count; // [cells], zeroed by the caller
  int32_t useSmem;
};

__global__ void __launch_bounds__(256) ngpAssignKernel(const __grid_constant__ DevGeometry G, const AssignParams A) {
  extern __shared__ double sCount[];
  if (A.useSmem) {
    for (int i = threadIdx.x; i < G.cells; i += blockDim.x) sCount[i] = 0.0;
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t nRounded = (A.n + 31) & ~int64_t(31);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nRounded; i += stride) {
    const bool live = i < A.n;
    const int cell = live ? posToCell(G, A.x[i], A.y[i], G.dim > 2 ? A.z[i] : 0.0) : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    if (live && lane == __ffs(peers) - 1) {
      const double add = A.nrCarriers * __popc(peers);
      if (A.useSmem)
        atomicAdd(&sCount[cell], add);
      else
        atomicAdd(&A.count[cell], add);
    }
  }
  if (A.useSmem) {
    __syncthreads();
    for (int i = threadIdx.x; i < G.cells; i += blockDim.x)
      if (sCount[i] != 0.0) atomicAdd(&A.count[i], sCount[i]);
  }
}

// ---------------------------------------------------------------------------
// K4: concentration from counts; E = -grad(phi) with the reference's boundary rules
__global__ void concentrationKernel(const __grid_constant__ DevGeometry G, const double *count, double *conc) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.cells) return;
  int c[3];
  cellCoord(G, cell, c);
  double v = __dmul_rn(__ddiv_rn(count[cell], G.ni), __ddiv_rn(1.0, G.cellVolume));
  for (int i = 0; i < G.dim; i++)
    if (c[i] == 0 || c[i] == G.extent[i] - 1) v = __dmul_rn(v, 2.0);
  conc[cell] = v;
}

// emcSimulationResults::updateAverageCharacteristics (:87-93): running sums of potential and concentration
__global__ void accumulateKernel(int cells, const double *pot, const double *conc, double *sumPot, double *sumConc) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= cells) return;
  sumPot[cell] = __dadd_rn(sumPot[cell], pot[cell]);
  sumConc[cell] = __dadd_rn(sumConc[cell], conc[cell]);
}

// interior: Vt (phi[prev] - phi[next]) / (2 h); on a face: 0 (artificial boundary) or the inner neighbour's value (contact)
__global__ void efieldKernel(const __grid_constant__ DevGeometry G, const double *pot, double *e) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= G.cells) return;
  int c[3];
  cellCoord(G, cell, c);
  const int stride[3] = {1, G.extent[0], G.extent[0] * G.extent[1]};
  for (int i = 0; i < G.dim; i++) {
    int at = cell; // the cell whose central difference this cell reports
    bool zero = false;
    if (c[i] == 0) {
      zero = G.faceContact[cell * 2 * G.dim + 2 * i] == -1;
      at = cell + stride[i];
    } else if (c[i] == G.extent[i] - 1) {
      zero = G.faceContact[cell * 2 * G.dim + 2 * i + 1] == -1;
      at = cell - stride[i];
    }
    double v = 0.0;
    if (!zero)
      v = __ddiv_rn(__dmul_rn(__dsub_rn(pot[at - stride[i]], pot[at + stride[i]]), G.thermalVoltage),
                    __dmul_rn(2.0, G.spacing[i]));
    e[(size_t)i * G.cells + cell] = v;
  }
}

// ---------------------------------------------------------------------------
// K5: nonlinear SOR in the reference's own (lexicographic Gauss-Seidel) update order.
// The update of a cell needs the NEW values of its lower neighbours and the OLD values of
// its upper neighbours; all cells on a hyperplane x+y(+z) = const are therefore independent,
// and sweeping the hyperplanes in order reproduces the sequential sweep of
// emcSORSolver.hpp:157-196 exactly (same iterates, same sweep count up to the last bits of exp).
// One CTA, potential resident in shared memory, convergence loop inside the kernel.
struct SorParams {
  double *pot;        // [cells] in/out
  const double *conc; // [cells] or nullptr: equilibrium solve (n = exp(phi), p = 1/n)
  double accuracy;    // normalised (volts / Vt)
  double omega;
  int32_t maxSweeps;
  int32_t potInSmem;
  int32_t *sweepsOut;
};

constexpr int kSorThreads = 512;

__global__ void __launch_bounds__(kSorThreads) sorKernel(const __grid_constant__ DevGeometry G, const SorParams S) {
  extern __shared__ double sPot[];
  __shared__ double sErr[kSorThreads / 32];
  __shared__ double sMax;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *pot = S.potInSmem ? sPot : S.pot;
  if (S.potInSmem) {
    for (int i = tid; i < G.cells; i += blockDim.x) sPot[i] = S.pot[i];
  }
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
  const int stride[3] = {1, G.extent[0], G.extent[0] * G.extent[1]};
  const int ex = G.extent[0], ey = G.extent[1], ez = G.dim > 2 ? G.extent[2] : 1;
  const int nPlanes = ex + ey + ez - 2;
  __syncthreads();
  int sweeps = 0;
  for (;;) {
    double myErr = 0.0;
    for (int plane = 0; plane < nPlanes; plane++) {
      // cells with x + y + z == plane: enumerate (y, z) pairs, x follows
      const int nYZ = ey * ez;
      for (int yz = tid; yz < nYZ; yz += blockDim.x) {
        const int y = yz % ey, z = yz / ey;
        const int x = plane - y - z;
        if (x < 0 || x >= ex) continue;
        const int cell = x + ex * (y + ey * z);
        if (cellIsReservoir(G, cell)) continue;
        const int c[3] = {x, y, z};
        const double cur = pot[cell];
        double p, n;
        if (S.conc) {
          p = exp(-cur);
          n = S.conc[cell];
        } else {
          n = exp(cur);
          p = __ddiv_rn(1.0, n);
        }
        const double dop = __ddiv_rn(G.doping[cell], G.ni);
        double num = __dmul_rn(hProd, __dadd_rn(__dadd_rn(__dsub_rn(p, n), dop), __dmul_rn(cur, __dadd_rn(p, n))));
        double den = __dadd_rn(__dmul_rn(2.0, hFSum), __dmul_rn(hProd, __dadd_rn(n, p)));
        for (int i = 0; i < G.dim; i++) {
#pragma unroll
          for (int side = 0; side < 2; side++) {
            const bool atFace = side == 0 ? c[i] == 0 : c[i] == G.extent[i] - 1;
            if (!atFace) {
              num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? -stride[i] : stride[i])], hF[i]));
            } else {
              num = __dadd_rn(num, __dmul_rn(pot[cell + (side == 0 ? stride[i] : -stride[i])], hF[i]));
              const int ct = G.faceContact[cell * 2 * G.dim + 2 * i + side];
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
        pot[cell] = __dadd_rn(cur, delta);
        myErr = fmax(myErr, fabs(delta));
      }
      __syncthreads();
    }
    // max |delta| of the sweep
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) myErr = fmax(myErr, __shfl_xor_sync(0xffffffffu, myErr, o));
    if (lane == 0) sErr[warp] = myErr;
    __syncthreads();
    if (tid == 0) {
      double m = 0.0;
      for (int w = 0; w < (int)blockDim.x / 32; w++) m = fmax(m, sErr[w]);
      sMax = m;
    }
    __syncthreads();
    sweeps++;
    if (!(sMax > S.accuracy) || (S.maxSweeps > 0 && sweeps >= S.maxSweeps)) break;
  }
  if (S.potInSmem)
    for (int i = tid; i < G.cells; i += blockDim.x) S.pot[i] = sPot[i];
  if (tid == 0 && S.sweepsOut) *S.sweepsOut = sweeps;
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
struct DeviceStepParams {
  BulkParams P;      // ensemble, model, tables, rng (box / force / dir unused)
  const double *e;   // [dim][cells]
  double charge;
  int8_t *removed;   // [n] out
  int32_t *removedPerContact; // [nContacts], zeroed by the caller
};

template <bool EXACT, int DIM>
__device__ __forceinline__ bool deviceDriftParticle(const DevGeometry &G, const DevModel &model, Particle &p, double dt,
                                                    const Vec3 &force) {
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
    if (cellIsOhmic(G, posToCell(G, cl[0], cl[1], cl[2]))) {
      removed = true;
#pragma unroll
      for (int i = 0; i < DIM; i++) pos[i] = cl[i];
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

template <int DIM>
__device__ __forceinline__ Vec3 ngpForce(const DevGeometry &G, const double *e, const Particle &p, double charge) {
  const int cell = posToCell(G, p.pos.x, p.pos.y, p.pos.z);
  Vec3 f;
  f.x = __dmul_rn(charge, e[cell]);
  f.y = __dmul_rn(charge, e[(size_t)G.cells + cell]);
  f.z = DIM > 2 ? __dmul_rn(charge, e[2 * (size_t)G.cells + cell]) : 0.0;
  return f;
}

template <bool EXACT, int RNG_MODE, int DIM>
__global__ void __launch_bounds__(kBulkThreads, 2)
    deviceStepKernel(const __grid_constant__ DevGeometry G, const __grid_constant__ DeviceStepParams D) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  __shared__ uint64_t tableBar;
  const BulkParams &P = D.P;
  const CtaState C = stageCta(P, smemRaw, &tableBar, 0, 0);
  const DevModel &model = *C.model;
  using A = Arith<EXACT>;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
    Particle p;
    Rng rng;
    loadParticle(P, i, p, rng);
    p.energy = P.stream[EMCGPU_ENERGY][i]; // Coulomb and the table look-up read it before the next drift
    if (DIM < 3) p.pos.z = 0.0;
    attachReplay<RNG_MODE>(P, i, rng);
    rng.step = (uint32_t)P.step0;
    const double dt = P.dt;
    Vec3 force = ngpForce<DIM>(G, D.e, p, D.charge);
    bool removed;
    if (p.tau >= dt) {
      removed = deviceDriftParticle<EXACT, DIM>(G, model, p, dt, force);
    } else {
      removed = deviceDriftParticle<EXACT, DIM>(G, model, p, p.tau, force);
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
            sampleFinalState<EXACT, RNG_MODE>(model, mech, p, rng);
          }
          if (P.evCap > 0) {
            const unsigned long long ev = atomicAdd(P.evCount, 1ull);
            if ((long long)ev < P.evCap) {
              long long *dst = P.events + 4 * ev;
              dst[0] = P.step0;
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
        force = ngpForce<DIM>(G, D.e, p, D.charge); // re-interpolated after every scattering (:112)
        removed = deviceDriftParticle<EXACT, DIM>(G, model, p, fmin(tRem, newTau), force);
        tRem = A::sub(tRem, newTau);
      }
    }
    p.tau = A::sub(p.tau, dt);
    storeParticle<RNG_MODE>(P, i, p, rng);
    D.removed[i] = removed ? 1 : 0;
    if (removed) atomicAdd(&D.removedPerContact[cellContact(G, posToCell(G, p.pos.x, p.pos.y, p.pos.z))], 1);
  }
}

// ---------------------------------------------------------------------------
// Order-preserving compaction (removeParticles, emcBasicParticleHandler.hpp:267-277): keep[i] != 0
// survives.  Three small kernels: per-block counts, exclusive scan of the counts (one CTA), scatter.
constexpr int kCompactThreads = 256;

__global__ void __launch_bounds__(kCompactThreads) compactCountKernel(const int8_t *drop, int64_t n, int32_t *blockCount) {
  __shared__ int sWarp[kCompactThreads / 32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool keep = i < n && !drop[i];
  const unsigned b = __ballot_sync(0xffffffffu, keep);
  if ((threadIdx.x & 31) == 0) sWarp[threadIdx.x >> 5] = __popc(b);
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kCompactThreads / 32; w++) s += sWarp[w];
    blockCount[blockIdx.x] = s;
  }
}

// exclusive scan of up to a few thousand block counts by one CTA; total -> out[nBlocks]
__global__ void __launch_bounds__(1024) compactScanKernel(int32_t *blockCount, int nBlocks) {
  __shared__ int sCarry;
  __shared__ int sWarp[32];
  if (threadIdx.x == 0) sCarry = 0;
  __syncthreads();
  for (int base = 0; base < nBlocks; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < nBlocks ? blockCount[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) sWarp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = sWarp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += t;
      }
      sWarp[threadIdx.x] = w;
    }
    __syncthreads();
    const int warpOff = (threadIdx.x >> 5) ? sWarp[(threadIdx.x >> 5) - 1] : 0;
    const int carry = sCarry;
    if (i < nBlocks) blockCount[i] = carry + warpOff + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) sCarry = carry + warpOff + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) blockCount[nBlocks] = sCarry;
}

struct EnsemblePtrs {
  double *stream[EMCGPU_N_STREAMS];
  uint32_t *packed;
  uint32_t *cursor; // replay cursors travel with their particle (may be null)
};

__global__ void __launch_bounds__(kCompactThreads)
    compactScatterKernel(const int8_t *drop, int64_t n, const int32_t *blockOffset, EnsemblePtrs src, EnsemblePtrs dst) {
  __shared__ int sWarp[kCompactThreads / 32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool keep = i < n && !drop[i];
  const unsigned b = __ballot_sync(0xffffffffu, keep);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sWarp[warp] = __popc(b);
  __syncthreads();
  int off = blockOffset[blockIdx.x];
  for (int w = 0; w < warp; w++) off += sWarp[w];
  if (keep) {
    const int64_t j = off + __popc(b & ((1u << lane) - 1u));
#pragma unroll
    for (int c = 0; c < EMCGPU_N_STREAMS; c++) dst.stream[c][j] = src.stream[c][i];
    dst.packed[j] = src.packed[i];
    if (src.cursor) dst.cursor[j] = src.cursor[i];
  }
}

// ---------------------------------------------------------------------------
// K6: ohmic contacts.  (1) mark the excess particles of every reservoir cell -- the reference scans the
// ensemble in index order and keeps a particle while the cell is below its expected population, so the
// FIRST particles of a cell (by index) survive; (2) inject the missing ones, cell by cell in storage order.
// Step (1) is done by one warp walking the ensemble in index order (the reservoir population is a few
// hundred particles; ranks within a cell come from __match_any_sync); step (2) is embarrassingly parallel.
struct ContactParams {
  const double *x, *y, *z;
  int64_t n;
  double nrCarriers;
  const double *expected; // [cells]
  double *have;           // [cells] out: population kept per reservoir cell
  int8_t *drop;           // [n] out
  int32_t *net;           // [nContacts] out: injected - deleted
  int32_t *injectCount;   // [cells + 1] out: particles to inject per cell, exclusive scan, total at [cells]
};

__global__ void __launch_bounds__(32) contactMarkKernel(const __grid_constant__ DevGeometry G, const ContactParams K) {
  const int lane = threadIdx.x;
  for (int c = lane; c < G.cells; c += 32) K.have[c] = 0.0;
  for (int c = lane; c < G.nContacts; c += 32) K.net[c] = 0;
  __syncwarp();
  for (int64_t base = 0; base < K.n; base += 32) {
    const int64_t i = base + lane;
    int cell = -1;
    if (i < K.n) {
      const int c = posToCell(G, K.x[i], K.y[i], G.dim > 2 ? K.z[i] : 0.0);
      if (cellIsReservoir(G, c)) cell = c;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    bool dropIt = false;
    if (cell >= 0) {
      // population of the cell seen by this particle = what earlier chunks left + earlier lanes of this chunk that were kept.
      // Within a cell every kept particle adds nrCarriers and, once the cell is full, all later ones are dropped, so the
      // number kept before lane l is min(rank, slots) with slots = ceil((expected - have) / nrCarriers) clipped at 0.
      const int rank = __popc(peers & ((1u << lane) - 1u));
      const double have = K.have[cell];
      const double room = K.expected[cell] - have;
      const int slots = room > 0.0 ? (int)ceil(room / K.nrCarriers) : 0;
      dropIt = rank >= slots;
      const int group = __popc(peers);
      if (rank == 0) { // the first lane of the group updates the cell and the contact counter
        const int kept = min(group, slots);
        K.have[cell] = have + kept * K.nrCarriers;
        if (group > kept) atomicAdd(&K.net[cellContact(G, cell)], -(group - kept));
      }
    }
    if (i < K.n) K.drop[i] = dropIt ? 1 : 0;
    __syncwarp();
  }
  // particles to inject per reservoir cell: while (diff > 0) { inject; diff -= nrCarriers }
  int carry = 0;
  for (int base = 0; base < G.cells; base += 32) {
    const int c = base + lane;
    int cnt = 0;
    if (c < G.cells && cellIsReservoir(G, c)) {
      const double diff = K.expected[c] - K.have[c];
      cnt = diff > 0.0 ? (int)ceil(diff / K.nrCarriers) : 0;
      if (cnt) atomicAdd(&K.net[cellContact(G, c)], cnt);
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (c < G.cells) K.injectCount[c] = carry + incl - cnt;
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) K.injectCount[G.cells] = carry;
}

// emcBasicParticleHandler::addParticle(isInitial = false): initParticlePos (emcParticleInitialization.hpp:14-29) +
// emcElectron::generateInjectedParticle (emcElectron.hpp:92-104).  Draw order: position (dim draws), valley,
// sub-valley, energy, cos(theta), phi, tau, grainTau = dim + 7 draws per particle; injected particle j of this
// step uses Philox(seed, counter = (j, 0xC0117AC7, step, .)) or, in replay mode, draws[j * (dim + 7) ...].
struct InjectParams {
  EnsemblePtrs ens;
  int64_t first;              // index of the first injected particle in the (already compacted) ensemble
  const int32_t *injectCount; // exclusive scan per cell, total at [cells]
  const DevModel *model;
  uint64_t seed;
  int64_t step;
  const uint64_t *replay; // flat draw stream of the contact phase or nullptr
  int64_t replayCount;
  int *status;
};

template <int DIM>
__global__ void __launch_bounds__(128) contactInjectKernel(const __grid_constant__ DevGeometry G, const InjectParams J) {
  const DevModel &model = *J.model;
  const int total = J.injectCount[G.cells];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
    // cell of injected particle j: last cell with injectCount[cell] <= j (binary search over the scan)
    int lo = 0, hi = G.cells - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (J.injectCount[mid] <= j) lo = mid; else hi = mid - 1;
    }
    // skip empty cells that share the same offset: move to the last cell whose range contains j
    while (lo + 1 < G.cells && J.injectCount[lo + 1] <= j) lo++;
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
        philox4x32_10((uint32_t)j, 0xC0117AC7u, (uint32_t)J.step, (uint32_t)(d >> 1), (uint32_t)J.seed,
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
    const int valley = (int)floor(__dmul_rn((double)model.nValleys, uniformLog(raw[d++])));
    const DevValley &v = model.valleys[valley];
    const int sub = (int)floor(__dmul_rn((double)v.deg, uniformLog(raw[d++])));
    const double energy = __dmul_rn(__dmul_rn(-1.5, G.thermalVoltage), log(uniformLog(raw[d++])));
    const double r2 = uniform01(raw[d++]);
    const double r1 = uniform01(raw[d++]);
    Vec3 k = randomDirection<true>(normWaveVec<true>(v, energy), r1, r2);
    double kk[3] = {k.x, k.y, k.z};
    for (int i = 0; i < DIM; i++)
      if ((c[i] == 0 && kk[i] < 0.0) || (c[i] == G.extent[i] - 1 && kk[i] > 0.0)) kk[i] = -kk[i];
    const int set = (region >= 0 && region < kMaxRegions) ? model.setOf[valley][region] : -1;
    const double tau0 = set >= 0 ? model.sets[set].tau : model.defaultTau;
    const double tau = __dmul_rn(-log(uniformLog(raw[d++])), tau0);
    const int64_t at = J.first + j;
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
  }
}

} // namespace emc
