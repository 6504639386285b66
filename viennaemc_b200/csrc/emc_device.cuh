// Device functions of the particle loop: arithmetic policy (EXACT / FAST),
// counter-based and replay RNG, valley math, free flight, selection and the
// final-state samplers.  Every function cites the reference code it replaces
// (paths relative to the reference tree).
#pragma once
#include <cstdint>

#include "emc_model.cuh"

namespace emc {

// ---------------------------------------------------------------------------
// Arithmetic policy.  EXACT: every operation individually rounded in the
// reference's order (the __d*_rn intrinsics are never contracted into FMAs),
// so that replay runs track the reference to the last bits of libm.  FAST:
// plain operators, nvcc contracts a*b+c into DFMA.
template <bool EXACT> struct Arith {
  static __device__ __forceinline__ double mul(double a, double b) {
    if constexpr (EXACT) return __dmul_rn(a, b); else return a * b;
  }
  static __device__ __forceinline__ double add(double a, double b) {
    if constexpr (EXACT) return __dadd_rn(a, b); else return a + b;
  }
  static __device__ __forceinline__ double sub(double a, double b) {
    if constexpr (EXACT) return __dsub_rn(a, b); else return a - b;
  }
  static __device__ __forceinline__ double div(double a, double b) {
    if constexpr (EXACT) return __ddiv_rn(a, b); else return a / b;
  }
  static __device__ __forceinline__ double sqrt(double a) {
    if constexpr (EXACT) return __dsqrt_rn(a); else return ::sqrt(a);
  }
};

// ---------------------------------------------------------------------------
// Random numbers.
enum RngMode : int { RNG_PHILOX = 0, RNG_REPLAY = 1 };

// Philox4x32-10 (Salmon et al. 2011).  One call yields two 64-bit draws.
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Per (particle, step) stream of raw 64-bit draws.  Draw i is word pair (i&1)
// of Philox(key = seed, counter = (id_lo, id_hi, step, i>>1)), or the next
// entry of the particle's recorded replay stream.
struct Rng {
  // philox
  uint32_t k0, k1, idLo, idHi, step, n;
  uint64_t cached;
  // replay
  const uint64_t *stream; // first unread draw of this particle
  const uint64_t *streamEnd;
  int *status;

  template <int MODE> __device__ __forceinline__ uint64_t raw() {
    if constexpr (MODE == RNG_PHILOX) {
      uint64_t r;
      if ((n & 1u) == 0u) {
        uint32_t o[4];
        philox4x32_10(idLo, idHi, step, n >> 1, k0, k1, o);
        r = (uint64_t)o[1] << 32 | o[0];
        cached = (uint64_t)o[3] << 32 | o[2];
      } else {
        r = cached;
      }
      n++;
      return r;
    } else {
      if (stream >= streamEnd) {
        atomicExch(status, (int)EMCGPU_E_REPLAY_EXHAUSTED);
        return 0x8000000000000000ull;
      }
      return *stream++;
    }
  }
};

// libstdc++ generate_canonical<double,53> on a 64-bit engine followed by
// uniform_real_distribution's u*(b-a)+a: always individually
// rounded, in both math modes, so that indices never depend on the mode.
__device__ __forceinline__ double canonical(uint64_t raw) {
  double u = __dmul_rn(__ull2double_rn(raw), 5.42101086242752217e-20 /* 2^-64 */);
  return u >= 1.0 ? 0.99999999999999988898 /* nextafter(1,0) */ : u;
}
__device__ __forceinline__ double uniform01(uint64_t raw) {
  // (u * (1 - 0)) + 0
  return canonical(raw);
}
__device__ __forceinline__ double uniformLog(uint64_t raw) {
  // U[1e-6, 1): u * (1. - 1e-6) + 1e-6   (emcParticleType.hpp:37)
  return __dadd_rn(__dmul_rn(canonical(raw), 1.0 - 1e-6), 1e-6);
}

// ---------------------------------------------------------------------------
// Valley math (include/ValleyTypes/*.hpp)
struct Vec3 {
  double x, y, z;
};

__device__ __forceinline__ double pick(const Vec3 &v, unsigned code) {
  const unsigned i = code & 3u;
  const double a = i == 0 ? v.x : (i == 1 ? v.y : v.z);
  return (code & 4u) ? -a : a;
}
__device__ __forceinline__ Vec3 permute(const Vec3 &v, unsigned p) {
  return Vec3{pick(v, p), pick(v, p >> 4), pick(v, p >> 8)};
}

// transformToEllipseCoord (emcNonParabolicAnistropValley.hpp:140-153): R_s * v
template <bool EXACT>
__device__ __forceinline__ Vec3 toEllipse(const DevValley &v, int s, const Vec3 &a) {
  using A = Arith<EXACT>;
  if (v.rotKind == ROT_IDENTITY) return a;
  if (v.rotKind == ROT_SIGNED_PERMUTATION) return permute(a, v.permToE[s]);
  const double *r = v.rot[s];
  Vec3 o;
  o.x = A::add(A::add(A::mul(a.x, r[0]), A::mul(a.y, r[1])), A::mul(a.z, r[2]));
  o.y = A::add(A::add(A::mul(a.x, r[3]), A::mul(a.y, r[4])), A::mul(a.z, r[5]));
  o.z = A::add(A::add(A::mul(a.x, r[6]), A::mul(a.y, r[7])), A::mul(a.z, r[8]));
  return o;
}
// transformToDeviceCoord (:157-170): R_s^T * v
template <bool EXACT>
__device__ __forceinline__ Vec3 toDevice(const DevValley &v, int s, const Vec3 &a) {
  using A = Arith<EXACT>;
  if (v.rotKind == ROT_IDENTITY) return a;
  if (v.rotKind == ROT_SIGNED_PERMUTATION) return permute(a, v.permToD[s]);
  const double *r = v.rot[s];
  Vec3 o;
  o.x = A::add(A::add(A::mul(a.x, r[0]), A::mul(a.y, r[3])), A::mul(a.z, r[6]));
  o.y = A::add(A::add(A::mul(a.x, r[1]), A::mul(a.y, r[4])), A::mul(a.z, r[7]));
  o.z = A::add(A::add(A::mul(a.x, r[2]), A::mul(a.y, r[5])), A::mul(a.z, r[8]));
  return o;
}

// getGamma (emcNonParabolicAnistropValley.hpp:122; parabolic: E)
template <bool EXACT> __device__ __forceinline__ double gammaOf(const DevValley &v, double e) {
  using A = Arith<EXACT>;
  if (!v.nonParabolic) return e;
  return A::mul(e, A::add(1.0, A::mul(v.alpha, e)));
}

// getEnergy(k): non-parabolic g = hbar*hbar*|k|^2/(m q), E = g/(1+sqrt(1+2 a g))
// (emcNonParabolicAnistropValley.hpp:109-113, emcNonParabolicIsotropValley.hpp:78-82);
// parabolic hbar*hbar*|k|^2/(2 m q) (emcParabolicIsotropValley.hpp:62-65)
template <bool EXACT> __device__ __forceinline__ double energyOfSq(const DevValley &v, double sq) {
  using A = Arith<EXACT>;
  if constexpr (EXACT) {
    const double h2 = kHbar * kHbar; // folded exactly like the reference's constexpr product
    if (v.nonParabolic) {
      const double g = A::div(A::mul(h2, sq), v.xMq);
      return A::div(g, A::add(1.0, A::sqrt(A::add(1.0, A::mul(A::mul(2.0, v.alpha), g)))));
    }
    return A::div(A::mul(h2, sq), v.xTwoMq);
  } else {
    const double g = v.fE * sq;
    if (v.nonParabolic) return g / (1.0 + ::sqrt(fma(2.0 * v.alpha, g, 1.0)));
    return g;
  }
}
template <bool EXACT> __device__ __forceinline__ double sqNorm(const Vec3 &k) {
  using A = Arith<EXACT>;
  // emcUtil.hpp:26-31: res = 0; res += e*e, in order
  return A::add(A::add(A::mul(k.x, k.x), A::mul(k.y, k.y)), A::mul(k.z, k.z));
}

// getEffMassCond(E) (:90-92): m (1 + 2 E alpha); parabolic: m
template <bool EXACT> __device__ __forceinline__ double massFactor(const DevValley &v, double e) {
  using A = Arith<EXACT>;
  if (!v.nonParabolic) return 1.0;
  return A::add(1.0, A::mul(A::mul(2.0, e), v.alpha));
}

// getNormWaveVec(E): sqrt(2 m gamma q)/hbar (aniso non-parabolic, :103-106) or
// sqrt(2 m q gamma)/hbar (the three other classes)
template <bool EXACT> __device__ __forceinline__ double normWaveVec(const DevValley &v, double e) {
  using A = Arith<EXACT>;
  const double g = gammaOf<EXACT>(v, e);
  const double twoM = A::mul(2.0, v.mBand);
  double arg;
  if (v.kind == EMCGPU_VALLEY_NONPARABOLIC_ANISOTROP)
    arg = A::mul(A::mul(twoM, g), kQ);
  else
    arg = A::mul(A::mul(twoM, kQ), g);
  return A::div(A::sqrt(arg), kHbar);
}

// getVelocity(k, E, s) . dirE, where dirE is the field direction already in
// the sub-valley's ellipse frame (emcNonParabolicAnistropValley.hpp:126-136 and
// the three sibling classes; basicBulkParticleHandler.hpp:326-347).
template <bool EXACT>
__device__ __forceinline__ double driftVelocity(const DevValley &v, int s, const Vec3 &k, double e,
                                                const Vec3 &dir) {
  using A = Arith<EXACT>;
  if constexpr (EXACT) {
    double npf = 1.0;
    if (v.nonParabolic)
      npf = A::sqrt(A::add(1.0, A::mul(A::mul(4.0, v.alpha), gammaOf<EXACT>(v, e))));
    Vec3 vel;
    if (v.kind & 2) { // anisotropic classes
      const Vec3 ke = toEllipse<EXACT>(v, s, k);
      const double den = v.nonParabolic ? A::mul(v.mBand, npf) : v.mBand;
      Vec3 ve;
      ve.x = A::div(A::mul(A::mul(kHbar, v.vogt[0]), ke.x), den);
      ve.y = A::div(A::mul(A::mul(kHbar, v.vogt[1]), ke.y), den);
      ve.z = A::div(A::mul(A::mul(kHbar, v.vogt[2]), ke.z), den);
      vel = toDevice<EXACT>(v, s, ve);
    } else {
      const double f = v.nonParabolic ? A::div(kHbar, A::mul(v.mBand, npf)) : A::div(kHbar, v.mBand);
      vel.x = A::mul(k.x, f);
      vel.y = A::mul(k.y, f);
      vel.z = A::mul(k.z, f);
    }
    // innerProduct (emcUtil.hpp:51-54)
    return A::add(A::add(A::mul(vel.x, dir.x), A::mul(vel.y, dir.y)), A::mul(vel.z, dir.z));
  } else {
    // sqrt(1 + 4 a gamma(E)) == 1 + 2 a E for the Kane dispersion
    const Vec3 ke = toEllipse<false>(v, s, k);
    const Vec3 de = toEllipse<false>(v, s, dir);
    const double num = v.fVel[0] * ke.x * de.x + v.fVel[1] * ke.y * de.y + v.fVel[2] * ke.z * de.z;
    return v.nonParabolic ? num / fma(2.0 * v.alpha, e, 1.0) : num;
  }
}

// getVelocity(k, E, s) in the device frame (emcNonParabolicAnistropValley.hpp:126-136 and the three sibling classes), in the
// reference's operation order (printVelocities, basicBulkParticleHandler.hpp:251-285)
__device__ __forceinline__ Vec3 velocityVector(const DevValley &v, int s, const Vec3 &k, double e) {
  using A = Arith<true>;
  double npf = 1.0;
  if (v.nonParabolic) npf = A::sqrt(A::add(1.0, A::mul(A::mul(4.0, v.alpha), gammaOf<true>(v, e))));
  if (v.kind & 2) {
    const Vec3 ke = toEllipse<true>(v, s, k);
    const double den = v.nonParabolic ? A::mul(v.mBand, npf) : v.mBand;
    Vec3 ve;
    ve.x = A::div(A::mul(A::mul(kHbar, v.vogt[0]), ke.x), den);
    ve.y = A::div(A::mul(A::mul(kHbar, v.vogt[1]), ke.y), den);
    ve.z = A::div(A::mul(A::mul(kHbar, v.vogt[2]), ke.z), den);
    return toDevice<true>(v, s, ve);
  }
  const double f = v.nonParabolic ? A::div(kHbar, A::mul(v.mBand, npf)) : A::div(kHbar, v.mBand);
  return Vec3{A::mul(k.x, f), A::mul(k.y, f), A::mul(k.z, f)};
}

// ---------------------------------------------------------------------------
// Particle state held in registers during a step.
struct Particle {
  Vec3 k;
  double energy, tau;
  Vec3 pos;
  int valley, sub, region;
};

// drift() (include/emcParticleDrift.hpp:12-36): Herring-Vogt free flight of
// duration dt under `force` (device frame), leapfrog position update with the
// conduction mass at the NEW energy.  DIM = number of position components that
// are advanced (2-D devices drop z, :33-35).
template <bool EXACT, int DIM>
__device__ __forceinline__ void drift(const DevValley &v, Particle &p, double dt, const Vec3 &force) {
  using A = Arith<EXACT>;
  const Vec3 kOld = toEllipse<EXACT>(v, p.sub, p.k);
  const Vec3 fE = toEllipse<EXACT>(v, p.sub, force);
  Vec3 kNew, dP;
  if constexpr (EXACT) {
    kNew.x = A::add(kOld.x, A::div(A::mul(A::mul(fE.x, dt), v.vogt[0]), kHbar));
    kNew.y = A::add(kOld.y, A::div(A::mul(A::mul(fE.y, dt), v.vogt[1]), kHbar));
    kNew.z = A::add(kOld.z, A::div(A::mul(A::mul(fE.z, dt), v.vogt[2]), kHbar));
    p.k = toDevice<EXACT>(v, p.sub, kNew);
    p.energy = energyOfSq<EXACT>(v, sqNorm<EXACT>(p.k));
    const double mass = v.nonParabolic ? A::mul(v.mCond, massFactor<EXACT>(v, p.energy)) : v.mCond;
    dP.x = A::div(A::mul(A::mul(A::mul(kHbar, v.vogt[0]), A::div(A::add(kNew.x, kOld.x), 2.0)), dt), mass);
    dP.y = A::div(A::mul(A::mul(A::mul(kHbar, v.vogt[1]), A::div(A::add(kNew.y, kOld.y), 2.0)), dt), mass);
    dP.z = A::div(A::mul(A::mul(A::mul(kHbar, v.vogt[2]), A::div(A::add(kNew.z, kOld.z), 2.0)), dt), mass);
  } else {
    kNew.x = fma(fE.x * v.fDk[0], dt, kOld.x);
    kNew.y = fma(fE.y * v.fDk[1], dt, kOld.y);
    kNew.z = fma(fE.z * v.fDk[2], dt, kOld.z);
    p.k = toDevice<EXACT>(v, p.sub, kNew);
    p.energy = energyOfSq<EXACT>(v, kNew.x * kNew.x + kNew.y * kNew.y + kNew.z * kNew.z);
    const double w = v.nonParabolic ? dt / fma(2.0 * v.alpha, p.energy, 1.0) : dt;
    dP.x = v.fPos[0] * (kNew.x + kOld.x) * w;
    dP.y = v.fPos[1] * (kNew.y + kOld.y) * w;
    dP.z = v.fPos[2] * (kNew.z + kOld.z) * w;
  }
  const Vec3 d = toDevice<EXACT>(v, p.sub, dP);
  p.pos.x = A::add(p.pos.x, d.x);
  if constexpr (DIM > 1) p.pos.y = A::add(p.pos.y, d.y);
  if constexpr (DIM > 2) p.pos.z = A::add(p.pos.z, d.z);
}

// ---------------------------------------------------------------------------
// FAST-mode full-dt free flight (the >= 98 % case: tau >= dt, no scattering).
// Because the field is uniform and the rotations are linear, the Herring-Vogt
// update of emcParticleDrift.hpp:12-36 collapses, per (valley, sub-valley), to
//   k'   = k + dk                       dk = R^T diag(vogt/hbar) R F dt
//   E'   = g / (1 + S)                  g = hbar^2 |k'|^2/(m q), S = sqrt(1 + 2 a g)
//   pos' = pos + M (k' + k) dt / S      M = R^T diag(hbar vogt/(2 m)) R ; 1 + 2 a E' == S
//   v.Ê  = (c . k') / S                 c = R^T diag(hbar vogt/m) R Ê
// (parabolic valleys: a = 0, S = 1).  The constants are rebuilt by every CTA at
// kernel start from the valley description; one reciprocal square root and one
// reciprocal replace the reference's sqrt + 5 divisions.
struct FastSub {
  double dk[3];
  double m[9];
  double c[3];
  double a[3]; // diagonal of R^T diag(vogt) R: the Herring-Vogt factor seen along each DEVICE axis (signed permutations)
};
// Launch-uniform constants of the flights of one valley.  Built on the host (emcgpu.cu: buildFlightConst) and passed as kernel
// parameters, so that every kernel -- and the flight kernel, which reads them straight from the constant bank -- uses the
// same numbers.  With a signed-permutation rotation the collapsed step above becomes, per device axis i,
//   k'_i = k_i + a_i G_i      pos'_i = pos_i + (k'_i + k_i) a_i K2 / S      v.Ê = sum_i K4_i k'_i a_i K2 / S
struct FlightConst {
  double fE;    // hbar^2/(m q)  (parabolic: hbar^2/(2 m q) * 2, see stageCta)
  double c2a;   // 2 alpha fE
  double inv2a; // 1/(2 alpha) (0 for parabolic valleys)
  double Fh[3]; // force / hbar
  double G[3];  // Fh dt
  double KP;    // hbar/(2 m)
  double K2;    // KP dt
  double KV[3]; // Ê 2 KP   (v.Ê = r sum_i KV_i a_i k'_i)
  double K4[3]; // KV / K2
  int32_t diag, nonParabolic;
};
struct FastValley {
  FlightConst f;
};

__device__ __forceinline__ void buildFastSub(const DevValley &v, int s, const Vec3 &force, const Vec3 &dir,
                                             double dt, FastSub &o) {
  const double *r = v.rot[s];
  const double f[3] = {force.x, force.y, force.z}, d[3] = {dir.x, dir.y, dir.z};
  double fe[3], de[3];
  for (int i = 0; i < 3; i++) {
    fe[i] = (r[3 * i] * f[0] + r[3 * i + 1] * f[1] + r[3 * i + 2] * f[2]) * v.fDk[i] * dt;
    de[i] = (r[3 * i] * d[0] + r[3 * i + 1] * d[1] + r[3 * i + 2] * d[2]) * v.fVel[i];
  }
  for (int a = 0; a < 3; a++) {
    o.dk[a] = r[a] * fe[0] + r[3 + a] * fe[1] + r[6 + a] * fe[2];
    o.c[a] = r[a] * de[0] + r[3 + a] * de[1] + r[6 + a] * de[2];
    for (int b = 0; b < 3; b++)
      o.m[3 * a + b] = r[a] * v.fPos[0] * r[b] + r[3 + a] * v.fPos[1] * r[3 + b] + r[6 + a] * v.fPos[2] * r[6 + b];
    // exact for signed permutations: one term, (+-1)^2 vogt
    o.a[a] = r[a] * r[a] * v.vogt[0] + r[3 + a] * r[3 + a] * v.vogt[1] + r[6 + a] * r[6 + a] * v.vogt[2];
  }
}

// 1/sqrt(x) and 1/x for x in the normal range, without the special-case
// branches of the library versions: MUFU seed (reads the high word only,
// ~2^-20) + one cubic Newton step (-> ~2^-60 before rounding).
__device__ __forceinline__ double rsqrtNormal(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0); // 1 - x y^2
  const double t = fma(0.375, e, 0.5) * e;
  return fma(y, t, y);
}
__device__ __forceinline__ double rcpNormal(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x, y, 1.0);
  const double t = fma(e, e, e);
  return fma(y, t, y);
}

// ---- flights through a valley whose rotations are signed permutations (FlightConst::diag) ----------------------
// The core of drift() (emcParticleDrift.hpp:12-36) for one flight: k' = k + a g (g = force/hbar * duration), the
// position advance with the conduction mass of the NEW energy (kp = hbar/(2m) * duration).  Leaves |k'|^2, x = 1 + 2 a g
// = S^2 and r = 1/S for the energy / velocity forms below.  One MUFU + 25 FP64 instructions; no divisions, no branches.
struct FlightAux {
  double sq, x, r, w0, w1, w2;
};
__device__ __forceinline__ void flightCore(double a0, double a1, double a2, double g0, double g1, double g2, double kp,
                                           double c2a, double &kx, double &ky, double &kz, double &px, double &py,
                                           double &pz, FlightAux &o) {
  const double nx = fma(a0, g0, kx), ny = fma(a1, g1, ky), nz = fma(a2, g2, kz);
  o.sq = fma(nz, nz, fma(ny, ny, nx * nx));
  o.x = fma(c2a, o.sq, 1.0);
  o.r = rsqrtNormal(o.x);
  const double w = kp * o.r;
  o.w0 = a0 * w;
  o.w1 = a1 * w;
  o.w2 = a2 * w;
  px = fma(nx + kx, o.w0, px);
  py = fma(ny + ky, o.w1, py);
  pz = fma(nz + kz, o.w2, pz);
  kx = nx;
  ky = ny;
  kz = nz;
}
// The same flight when the force has ONE component, along device axis AXIS (g = that component): the two transverse
// components of k do not change, so their two FMAs and the two additions k' + k of the position advance drop out.  The
// caller passes the transverse Herring-Vogt factors DOUBLED (aT2 = 2 a): (k' + k) (a w) = (2 k)(a w) = k ((2 a) w) with
// exact scalings by two, i.e. the results equal flightCore's bit for bit (21 FP64 instructions + MUFU instead of 25).
// o.w0 / w1 / w2: only the component of AXIS is the w of flightCore, the transverse ones are doubled.
template <int AXIS>
__device__ __forceinline__ void flightCoreAxis(double a0, double a1, double a2, double g, double kp, double c2a, double &kx,
                                               double &ky, double &kz, double &px, double &py, double &pz, FlightAux &o) {
  const double nx = AXIS == 0 ? fma(a0, g, kx) : kx, ny = AXIS == 1 ? fma(a1, g, ky) : ky, nz = AXIS == 2 ? fma(a2, g, kz) : kz;
  o.sq = fma(nz, nz, fma(ny, ny, nx * nx));
  o.x = fma(c2a, o.sq, 1.0);
  o.r = rsqrtNormal(o.x);
  const double w = kp * o.r;
  o.w0 = a0 * w;
  o.w1 = a1 * w;
  o.w2 = a2 * w;
  px = fma(AXIS == 0 ? nx + kx : kx, o.w0, px);
  py = fma(AXIS == 1 ? ny + ky : ky, o.w1, py);
  pz = fma(AXIS == 2 ? nz + kz : kz, o.w2, pz);
  kx = nx;
  ky = ny;
  kz = nz;
}
// E = g/(1 + S), g = fE |k|^2 (getEnergy, emcNonParabolicAnistropValley.hpp:109-113), to the last bit or two
__device__ __forceinline__ double flightEnergy(double fE, const FlightAux &o) {
  const double g = fE * o.sq;
  const double d = fma(o.x, o.r, 1.0); // 1 + S
  const double y = rcpNormal(d);
  const double e = g * y;
  return fma(fma(-d, e, g), y, e);
}
// S - 1 = 2 alpha E with one rounding (the per-step energy observable of the flight kernel sums this)
__device__ __forceinline__ double flightSm1(const FlightAux &o) { return fma(o.x, o.r, -1.0); }
// v.Ê of the state after a full-dt flight: sum_i K4_i k'_i a_i K2 / S
__device__ __forceinline__ double flightVelocityDt(double k40, double k41, double k42, double kx, double ky, double kz,
                                                   const FlightAux &o) {
  return fma(k42 * kz, o.w2, fma(k41 * ky, o.w1, (k40 * kx) * o.w0));
}
// v.Ê of a state with known r = 1/S: r sum_i KV_i a_i k_i
__device__ __forceinline__ double flightVelocity(const FlightConst &f, const FastSub &fs, double kx, double ky, double kz,
                                                 double r) {
  return fma(f.KV[2] * fs.a[2], kz, fma(f.KV[1] * fs.a[1], ky, (f.KV[0] * fs.a[0]) * kx)) * r;
}
// exact periodic wrap (basicBulkParticleHandler.hpp:600-613) behind one unsigned compare of the high words: a coordinate in
// [0, box) with a high word below the box's needs nothing; negative values (sign bit) and values near or beyond the box go on
// to the exact test.
__device__ __forceinline__ bool mayNeedWrap(double x, uint32_t boxHi) { return (uint32_t)__double2hiint(x) >= boxHi; }
__device__ __forceinline__ double wrapExact(double x, double b) { return x < 0.0 ? x + b : (x > b ? x - b : x); }

// One whole time step of a particle that does not scatter (tau >= dt):
// drift(dt), periodic wrap, tau -= dt; returns v.Ê for the drift-velocity
// observable (basicBulkParticleHandler.hpp:195-213, :326-347).
__device__ __forceinline__ double fastStep(const FastSub &fs, const FastValley &fv, double dt, const Vec3 &box,
                                           double &kx, double &ky, double &kz, double &energy, double &tau,
                                           double &px, double &py, double &pz) {
  const FlightConst &f = fv.f;
  if (f.diag) {
    FlightAux o;
    flightCore(fs.a[0], fs.a[1], fs.a[2], f.G[0], f.G[1], f.G[2], f.K2, f.c2a, kx, ky, kz, px, py, pz, o);
    if (mayNeedWrap(px, (uint32_t)__double2hiint(box.x)) || mayNeedWrap(py, (uint32_t)__double2hiint(box.y)) ||
        mayNeedWrap(pz, (uint32_t)__double2hiint(box.z))) {
      px = wrapExact(px, box.x);
      py = wrapExact(py, box.y);
      pz = wrapExact(pz, box.z);
    }
    energy = flightEnergy(f.fE, o);
    tau -= dt;
    return flightVelocityDt(f.K4[0], f.K4[1], f.K4[2], kx, ky, kz, o);
  }
  const double nx = kx + fs.dk[0], ny = ky + fs.dk[1], nz = kz + fs.dk[2];
  const double sq = fma(nz, nz, fma(ny, ny, nx * nx));
  const double g = f.fE * sq;
  const double x = fma(f.c2a, sq, 1.0);
  const double r = rsqrtNormal(x); // 1/S
  const double d = fma(x, r, 1.0); // 1 + S
  const double y = rcpNormal(d);
  double e = g * y;
  e = fma(fma(-d, e, g), y, e);
  const double w = dt * r;
  const double sx = (nx + kx) * w, sy = (ny + ky) * w, sz = (nz + kz) * w;
  const double dx = fma(fs.m[2], sz, fma(fs.m[1], sy, fs.m[0] * sx));
  const double dy = fma(fs.m[5], sz, fma(fs.m[4], sy, fs.m[3] * sx));
  const double dz = fma(fs.m[8], sz, fma(fs.m[7], sy, fs.m[6] * sx));
  px += dx;
  py += dy;
  pz += dz;
  px += px < 0.0 ? box.x : (px > box.x ? -box.x : 0.0);
  py += py < 0.0 ? box.y : (py > box.y ? -box.y : 0.0);
  pz += pz < 0.0 ? box.z : (pz > box.z ? -box.z : 0.0);
  kx = nx;
  ky = ny;
  kz = nz;
  energy = e;
  tau -= dt;
  return fma(fs.c[2], nz, fma(fs.c[1], ny, fs.c[0] * nx)) * r;
}

// periodic wrap of the bulk handler (basicBulkParticleHandler.hpp:600-613)
template <bool EXACT> __device__ __forceinline__ double wrap1(double x, double maxPos) {
  using A = Arith<EXACT>;
  if (x < 0.0) return A::add(x, maxPos);
  if (x > maxPos) return A::sub(x, maxPos);
  return x;
}

// initRandomDirection (include/emcUtil.hpp:131-139)
template <bool EXACT> __device__ __forceinline__ Vec3 randomDirection(double norm, double rand1, double rand2) {
  using A = Arith<EXACT>;
  const double phi = A::mul(2.0 * kPi, rand1);
  const double c = A::sub(1.0, A::mul(2.0, rand2));
  double sp, cp;
  sincos(phi, &sp, &cp);
  const double ns = A::mul(norm, A::sqrt(A::sub(1.0, A::mul(c, c))));
  return Vec3{A::mul(ns, cp), A::mul(ns, sp), A::mul(norm, c)};
}

// initRandomDirectionWithRespectToCurrentK (include/emcUtil.hpp:143-175)
template <bool EXACT>
__device__ __forceinline__ Vec3 randomDirectionWrtK(const Vec3 &k, double cosTheta, double rnd) {
  using A = Arith<EXACT>;
  const double kxy = A::sqrt(A::add(A::mul(k.x, k.x), A::mul(k.y, k.y)));
  const double normK = A::sqrt(A::add(A::mul(kxy, kxy), A::mul(k.z, k.z)));
  if (normK == 0.0) return Vec3{0.0, 0.0, 0.0};
  const double ct0 = A::div(k.z, normK), st0 = A::div(kxy, normK);
  const double cfi0 = kxy > 0.0 ? A::div(k.x, kxy) : 1.0;
  const double sfi0 = kxy > 0.0 ? A::div(k.y, kxy) : 0.0;
  const double st = A::sqrt(A::sub(1.0, A::mul(cosTheta, cosTheta)));
  const double phi = A::mul(2.0 * kPi, rnd);
  double sp, cp;
  sincos(phi, &sp, &cp);
  const double kxp = A::mul(A::mul(normK, st), cp);
  const double kyp = A::mul(A::mul(normK, st), sp);
  const double kzp = A::mul(normK, cosTheta);
  Vec3 o;
  o.x = A::add(A::sub(A::mul(A::mul(kxp, cfi0), ct0), A::mul(kyp, sfi0)), A::mul(A::mul(kzp, cfi0), st0));
  o.y = A::add(A::add(A::mul(A::mul(kxp, sfi0), ct0), A::mul(kyp, cfi0)), A::mul(A::mul(kzp, sfi0), st0));
  o.z = A::add(A::mul(-kxp, st0), A::mul(kzp, ct0));
  return o;
}

// getEnergyLevel (include/emcScatterHandler.hpp:237-244), with the x86-64 g++
// behaviour of the size_t conversion
__device__ __forceinline__ int energyLevel(double energy, double dE, int nLevels) {
  const double f = floor(__ddiv_rn(energy, dE)) - 1.0;
  if (f == -1.0) return 0;
  if (!(f >= 0.0) || f > (double)(nLevels - 1)) return nLevels - 1;
  return (int)f;
}

// Null-scatter selection (include/emcScatterHandler.hpp:148-170) as a binary
// search: first m with r < cum[m]; -1 = self-scattering.  row points to the
// nMech cumulative entries of the particle's energy level.
__device__ __forceinline__ int selectMechanism(const double *row, int nMech, double r) {
  if (r > row[nMech - 1]) return -1;
  int lo = 0, hi = nMech; // answer in [lo, hi]; hi == nMech means none
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (r < row[mid]) hi = mid; else lo = mid + 1;
  }
  return lo < nMech ? lo : -1;
}

// Final-state samplers, selected by mechanism ID.
// emcPhononBath::recordEmission / recordAbsorption (:237-253): one event in the |q| bin of the bath
__device__ __forceinline__ void bathRecord(const BathView &B, int bath, double q, bool emission) {
  if (!B.counts || bath < 0 || bath >= B.nBaths) return;
  const unsigned long long idx = (unsigned long long)(q / B.dq);
  const int bin = idx >= (unsigned long long)B.nBins ? B.nBins - 1 : (int)idx;
  atomicAdd(&B.counts[((size_t)bath * 2 + (emission ? 0 : 1)) * B.nBins + bin], 1ull);
}
// emcPhononBath::sampleQ (:423-458)
__device__ __forceinline__ double bathSampleQ(const BathView &B, int bath, double qMin, double qMax, bool emission, double r) {
  if (B.nBins < 2 || qMax <= qMin) return qMin;
  long lo = (long)floor(qMin / B.dq), hi = (long)ceil(qMax / B.dq);
  if (lo < 0) lo = 0;
  if (hi > (long)B.nBins) hi = (long)B.nBins;
  if (hi - lo < 1) return qMin;
  const double *cw = B.cumW + (size_t)bath * (B.nBins + 1), *cwn = B.cumWN + (size_t)bath * (B.nBins + 1);
  auto S = [&](long i) { return emission ? cwn[i] + cw[i] : cwn[i]; };
  const double sLo = S(lo), sHi = S(hi), span = sHi - sLo;
  if (!(span > 0.0)) return 0.5 * (qMin + qMax);
  const double target = sLo + r * span;
  long a = lo, b = hi;
  while (b - a > 1) {
    const long mid = (a + b) / 2;
    if (S(mid) <= target) a = mid; else b = mid;
  }
  const double sA = S(a), sB = S(a + 1);
  const double frac = (sB > sA) ? (target - sA) / (sB - sA) : 0.5;
  return fmax(qMin, fmin(qMax, ((double)a + frac) * B.dq));
}

// ---- long-range single-layer mechanisms (emcFroehlichInteractionSingleLayer.hpp, emcPiezoelectricSingleLayerScatterMechanism.hpp)
// 1 / eps(q)^2, eps(q) = 1 + q_s / q  (emc2DScreening.hpp:65-76); plain IEEE operations in the reference's order
__device__ __forceinline__ double slScreeningFactor(double q, double qs) {
  const double eps = (q <= 0.0 || qs <= 0.0) ? 1.0 : 1.0 + qs / q;
  return 1.0 / (eps * eps);
}
// weight of a deflection angle: Froehlich erfc(w q/2)^2 / (eps^2 q) with q^2 = k^2 + k'^2 - 2 k k' cos(psi) (:48-56),
// piezoelectric erfc(w q/2)^2 / eps^2 with q = 2 k sin(theta/2) (:62-66)
template <bool FROEHLICH>
__device__ __forceinline__ double slAngularWeight(double angle, double k, double kPrime, double width, double qs) {
  if constexpr (FROEHLICH) {
    const double q2 = k * k + kPrime * kPrime - 2.0 * k * kPrime * cos(angle);
    const double q = sqrt(q2 > 0.0 ? q2 : 0.0);
    if (q <= 0.0) return 0.0;
    const double ff = erfc(width * q / 2.0);
    return ff * ff * slScreeningFactor(q, qs) / q;
  } else {
    const double q = 2.0 * k * sin(angle / 2.0);
    const double ff = erfc(width * q / 2.0);
    return ff * ff * slScreeningFactor(q, qs);
  }
}
// magnitude of the deflection in [0, pi] by inversion of the 128-point cumulative sum of the weight (midpoint rule).  The
// reference fills an array; the same running sum is formed twice here (total, then up to the target) -- identical partial sums,
// no array in local memory.  `total` comes back so that the caller can take the degenerate branch.
template <bool FROEHLICH>
__device__ __forceinline__ double slRunningTotal(double k, double kPrime, double width, double qs) {
  const double dAngle = kPi / 128;
  double c = 0.0;
  for (int i = 1; i <= 128; ++i) c = c + slAngularWeight<FROEHLICH>(((double)i - 0.5) * dAngle, k, kPrime, width, qs);
  return c;
}
template <bool FROEHLICH>
__device__ __forceinline__ double slInvertAngle(double target, double k, double kPrime, double width, double qs) {
  const double dAngle = kPi / 128;
  double below = 0.0, at = 0.0;
  int lo = 0;
  do { // lo: the first index with cdf[lo] >= target, or 128
    ++lo;
    below = at;
    at = below + slAngularWeight<FROEHLICH>(((double)lo - 0.5) * dAngle, k, kPrime, width, qs);
  } while (lo < 128 && at < target);
  return ((double)lo - 1.0 + (target - below) / (at - below)) * dAngle;
}

// ---- the other angle-resolved single-layer mechanisms: 2-D charged impurities, interface roughness, remote surface-optical
// phonons, screened intravalley optical phonons.  One shape (reference emc2DChargedImpurityScatterMechanism.hpp:107-139,
// emcSurfaceRoughnessScatterMechanism.hpp:94-126, emcRemoteSurfaceOpticalPhononMechanism.hpp:112-149,
// emcScreenedIntravalleyOpticalMechanism.hpp:104-141): E += dE (0: elastic), deflection magnitude by inversion of an N-point
// cumulative sum of the mechanism's weight (isotropic magnitude pi u when the sum vanishes), the side by a second draw,
// k = |k'| (cos(phi + angle), sin(phi + angle), 0).  KIND = sampler id.
template <int KIND> struct SlAngular;
template <> struct SlAngular<EMCGPU_SAMPLER_SINGLE_LAYER_CHARGED_IMPURITY> { // :65-72: (exp(-q d) / (q_s + q + r0 q^2))^2, q = 2 k sin(theta/2)
  static constexpr int steps = 512;
  static constexpr bool elastic = true;
  static __device__ __forceinline__ double weight(double theta, double k, double, const double *par) {
    const double q = 2.0 * k * sin(theta / 2.0);
    const double denom = par[2] + q + par[1] * q * q;
    if (denom <= 0.0) return 0.0;
    const double v = exp(-q * par[0]) / denom;
    return v * v;
  }
};
template <> struct SlAngular<EMCGPU_SAMPLER_SINGLE_LAYER_SURFACE_ROUGHNESS> { // :59-63: exp(-q^2 Lambda^2 / 4) / eps(q)^2
  static constexpr int steps = 256;
  static constexpr bool elastic = true;
  static __device__ __forceinline__ double weight(double theta, double k, double, const double *par) {
    const double q = 2.0 * k * sin(theta / 2.0);
    const double formFactor = exp(-q * q * par[1] / 4.0);
    return formFactor * slScreeningFactor(q, par[2]);
  }
};
template <> struct SlAngular<EMCGPU_SAMPLER_SINGLE_LAYER_REMOTE_SO> { // :63-70: exp(-2 q d) / (q eps(q)^2), q^2 = k^2 + k'^2 - 2 k k' cos
  static constexpr int steps = 128;
  static constexpr bool elastic = false;
  static __device__ __forceinline__ double weight(double theta, double k, double kPrime, const double *par) {
    const double q2 = k * k + kPrime * kPrime - 2.0 * k * kPrime * cos(theta);
    const double q = sqrt(fmax(0.0, q2));
    if (q <= 0.0) return 0.0;
    return exp(-2.0 * q * par[1]) * slScreeningFactor(q, par[2]) / q;
  }
};
template <> struct SlAngular<EMCGPU_SAMPLER_SINGLE_LAYER_SCREENED_OPTICAL> { // :58-62: 1 / eps(q)^2
  static constexpr int steps = 128;
  static constexpr bool elastic = false;
  static __device__ __forceinline__ double weight(double theta, double k, double kPrime, const double *par) {
    const double q2 = k * k + kPrime * kPrime - 2.0 * k * kPrime * cos(theta);
    return slScreeningFactor(sqrt(fmax(0.0, q2)), par[2]);
  }
};
// the running sum is formed twice (total, then up to the target) like slRunningTotal / slInvertAngle above: the same partial
// sums as the reference's array, nothing in local memory.  Out of line and by value (the loops stay out of the step kernels'
// hot code, no particle state behind a pointer): the signed deflection angle from the two uniform draws.
template <int KIND>
__device__ __noinline__ double slSignedAngle(double k, double kPrime, double p0, double p1, double p2, double u1, double u2) {
  using W = SlAngular<KIND>;
  const double par[3] = {p0, p1, p2};
  const double dAngle = kPi / W::steps;
  double total = 0.0;
  for (int i = 1; i <= W::steps; ++i) total = total + W::weight(((double)i - 0.5) * dAngle, k, kPrime, par);
  double angle;
  if (!(total > 0.0)) {
    angle = kPi * u1;
  } else {
    const double target = u1 * total;
    double below = 0.0, at = 0.0;
    int lo = 0;
    do { // lo: the first index whose cumulative sum reaches the target, or the last one
      ++lo;
      below = at;
      at = below + W::weight(((double)lo - 0.5) * dAngle, k, kPrime, par);
    } while (lo < W::steps && at < target);
    angle = ((double)lo - 1.0 + (target - below) / (at - below)) * dAngle;
  }
  return u2 < 0.5 ? -angle : angle; // left / right
}
template <int KIND, int RNG_MODE>
__device__ __forceinline__ void slAngularScatter(const DevValley &v, const DevMech &mech, Particle &p, Rng &rng) {
  const double k = normWaveVec<true>(v, p.energy);
  double kPrime = k;
  if constexpr (!SlAngular<KIND>::elastic) {
    p.energy = p.energy + mech.param[0];
    kPrime = normWaveVec<true>(v, p.energy);
  }
  const double u1 = uniform01(rng.raw<RNG_MODE>());
  const double u2 = uniform01(rng.raw<RNG_MODE>());
  const double angle = slSignedAngle<KIND>(k, kPrime, mech.param[0], mech.param[1], mech.param[2], u1, u2);
  double sa, ca;
  sincos(atan2(p.k.y, p.k.x) + angle, &sa, &ca);
  p.k = Vec3{kPrime * ca, kPrime * sa, 0.0};
}

template <bool EXACT, int RNG_MODE>
__device__ __forceinline__ void sampleFinalState(const DevModel &model, const DevMech &mech, Particle &p,
                                                 Rng &rng, const BathView &baths) {
  using A = Arith<EXACT>;
  switch (mech.sampler) {
  case EMCGPU_SAMPLER_ISOTROPIC_ELASTIC:
  case EMCGPU_SAMPLER_INTERVALLEY: {
    // one body for both so that a warp with lanes of either kind walks the common tail (two draws, sincos,
    // sqrt) once.  Elastic (emcAcousticScatterMechanism.hpp:70-72): |k| kept.  Intervalley
    // (emcZeroOrderInterValleyScatterMechanism.hpp:119-129 / :262-272, emcFirstOrderInterValleyScatterMechanism.hpp
    // :121-131 / :269-279): final sub-valley drawn first, energy shifted, |k| of the new energy.
    double nrm;
    if (mech.sampler == EMCGPU_SAMPLER_INTERVALLEY) {
      const uint64_t raw = rng.raw<RNG_MODE>();
      p.sub = mech.finalSub[p.sub][raw % (uint64_t)mech.nFinal];
      p.valley = mech.finalValley;
      p.energy = A::add(p.energy, mech.param[0]);
      nrm = normWaveVec<EXACT>(model.valleys[p.valley], p.energy);
    } else {
      nrm = A::sqrt(sqNorm<EXACT>(p.k));
    }
    // g++ evaluates the two dist(rng) arguments right to left: first draw = cos(theta) variate
    const double r2 = uniform01(rng.raw<RNG_MODE>());
    const double r1 = uniform01(rng.raw<RNG_MODE>());
    p.k = randomDirection<EXACT>(nrm, r1, r2);
    break;
  }
  case EMCGPU_SAMPLER_SINGLE_LAYER_ELASTIC:
  case EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY: {
    // emcAcousticSingleLayerScatterMechanism.hpp:63-81; emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp:116-147 /
    // :293-324: the final sub-valley (if the mechanism lists any) is finalSub[sub][floor(u nFinal)], then the in-plane
    // direction weighted by the Herring-Vogt factors of the (final) valley
    if (mech.sampler == EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY) {
      if (mech.nFinal > 0) {
        const double u = uniform01(rng.raw<RNG_MODE>());
        p.sub = mech.finalSub[p.sub][(int)floor(__dmul_rn(u, (double)mech.nFinal))];
      }
      p.valley = mech.finalValley;
      p.energy = A::add(p.energy, mech.param[0]);
    }
    const DevValley &v = model.valleys[p.valley];
    const double angle = A::mul(2.0 * kPi, uniform01(rng.raw<RNG_MODE>()));
    double sa, ca;
    sincos(angle, &sa, &ca);
    if (mech.sampler == EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY && mech.param[1] != 0.0) {
      // first-order classes (emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp:121-125, :270-274): |k| (cos, sin) without
      // the Herring-Vogt weighting, k_z left as it is
      const double nrm = normWaveVec<EXACT>(v, p.energy);
      p.k.x = A::mul(nrm, ca);
      p.k.y = A::mul(nrm, sa);
      break;
    }
    double kx = A::div(ca, v.vogt[0]), ky = A::div(sa, v.vogt[1]);
    const double factor = A::div(1.0, A::sqrt(A::add(A::mul(kx, kx), A::mul(ky, ky))));
    const double nrm = normWaveVec<EXACT>(v, p.energy);
    kx = A::mul(A::mul(kx, factor), nrm);
    ky = A::mul(A::mul(ky, factor), nrm);
    p.k = Vec3{kx, ky, 0.0};
    break;
  }
  case EMCGPU_SAMPLER_SINGLE_LAYER_FROEHLICH: {
    // emcFroehlichInteractionSingleLayer.hpp:149-168 (absorption), :273-292 (emission)
    const DevValley &v = model.valleys[p.valley];
    const double width = mech.param[1], qs = mech.param[2];
    const double kI = normWaveVec<true>(v, p.energy);
    p.energy = p.energy + mech.param[0];
    const double kF = normWaveVec<true>(v, p.energy);
    const double phi = atan2(p.k.y, p.k.x);
    const double total = slRunningTotal<true>(kI, kF, width, qs);
    double psi;
    if (!(total > 0.0)) {
      psi = 2.0 * kPi * uniform01(rng.raw<RNG_MODE>()); // degenerate: isotropic
    } else {
      const double target = uniform01(rng.raw<RNG_MODE>()) * total;
      const double psiMag = slInvertAngle<true>(target, kI, kF, width, qs);
      psi = (uniform01(rng.raw<RNG_MODE>()) < 0.5) ? psiMag : (2.0 * kPi - psiMag);
    }
    double sa, ca;
    sincos(phi + psi, &sa, &ca);
    p.k = Vec3{kF * ca, kF * sa, 0.0};
    break;
  }
  case EMCGPU_SAMPLER_SINGLE_LAYER_PIEZOELECTRIC: {
    // emcPiezoelectricSingleLayerScatterMechanism.hpp:110-139
    const DevValley &v = model.valleys[p.valley];
    const double width = mech.param[1], qs = mech.param[2];
    const double k = normWaveVec<true>(v, p.energy);
    const double total = slRunningTotal<false>(k, k, width, qs);
    const double phi = atan2(p.k.y, p.k.x);
    double theta;
    if (!(total > 0.0)) {
      theta = kPi * uniform01(rng.raw<RNG_MODE>());
    } else {
      const double target = uniform01(rng.raw<RNG_MODE>()) * total;
      theta = slInvertAngle<false>(target, k, k, width, qs);
    }
    if (uniform01(rng.raw<RNG_MODE>()) < 0.5) theta = -theta; // left / right
    double sa, ca;
    sincos(phi + theta, &sa, &ca);
    p.k = Vec3{k * ca, k * sa, 0.0};
    break;
  }
  case EMCGPU_SAMPLER_SINGLE_LAYER_CHARGED_IMPURITY:
    slAngularScatter<EMCGPU_SAMPLER_SINGLE_LAYER_CHARGED_IMPURITY, RNG_MODE>(model.valleys[p.valley], mech, p, rng);
    break;
  case EMCGPU_SAMPLER_SINGLE_LAYER_SURFACE_ROUGHNESS:
    slAngularScatter<EMCGPU_SAMPLER_SINGLE_LAYER_SURFACE_ROUGHNESS, RNG_MODE>(model.valleys[p.valley], mech, p, rng);
    break;
  case EMCGPU_SAMPLER_SINGLE_LAYER_REMOTE_SO:
    slAngularScatter<EMCGPU_SAMPLER_SINGLE_LAYER_REMOTE_SO, RNG_MODE>(model.valleys[p.valley], mech, p, rng);
    break;
  case EMCGPU_SAMPLER_SINGLE_LAYER_SCREENED_OPTICAL:
    slAngularScatter<EMCGPU_SAMPLER_SINGLE_LAYER_SCREENED_OPTICAL, RNG_MODE>(model.valleys[p.valley], mech, p, rng);
    break;
  case EMCGPU_SAMPLER_COULOMB: {
    // emcCoulombScatterMechanism.hpp:48-59
    const double g = gammaOf<EXACT>(model.valleys[p.valley], p.energy);
    const double rnd = uniform01(rng.raw<RNG_MODE>());
    const double den = A::add(A::div(A::mul(A::sub(1.0, rnd), g), mech.param[0]), 1.0);
    const double c = A::sub(1.0, A::div(A::mul(rnd, 2.0), den));
    const double r = uniform01(rng.raw<RNG_MODE>());
    p.k = randomDirectionWrtK<EXACT>(p.k, c, r);
    break;
  }
  case EMCGPU_SAMPLER_FROEHLICH:
  case EMCGPU_SAMPLER_SCREENED_FROEHLICH: {
    // polar-optical family (emcFroehlichInteraction.hpp, emcHotPhononFroehlichMechanism.hpp,
    // emcScreenedFroehlichInteraction.hpp); plain IEEE operations in the reference's order (-fmad=false)
    const DevValley &v = model.valleys[p.valley];
    const Vec3 kOld = p.k;
    const bool emission = mech.param[0] < 0.0;
    double cosTheta, kNew;
    if (mech.sampler == EMCGPU_SAMPLER_FROEHLICH) {
      const double initEnergy = p.energy;
      p.energy = p.energy + mech.param[0];
      const double finalEnergy = p.energy;
      const double d = sqrt(initEnergy) - sqrt(finalEnergy);
      const double f = 2.0 * sqrt(initEnergy * finalEnergy) / d / d;
      cosTheta = (1.0 + f - pow(1.0 + 2.0 * f, uniform01(rng.raw<RNG_MODE>()))) / f;
      cosTheta = fmax(-1.0, fmin(1.0, cosTheta));
      kNew = normWaveVec<true>(v, p.energy);
    } else {
      const double kI = normWaveVec<true>(v, p.energy);
      p.energy = p.energy + mech.param[0];
      const double kF = normWaveVec<true>(v, p.energy);
      const double r = uniform01(rng.raw<RNG_MODE>());
      const double B = 2.0 * kI * kF;
      if (B <= 0.0) {
        cosTheta = 1.0 - 2.0 * r;
      } else if (mech.flags & 1) {
        const double q = bathSampleQ(baths, mech.bath, fabs(kI - kF), kI + kF, emission, r);
        cosTheta = fmax(-1.0, fmin(1.0, (kI * kI + kF * kF - q * q) / B));
      } else {
        const double Ap = kI * kI + kF * kF + mech.param[1];
        const double num = Ap - B, den = Ap + B;
        if (num <= 0.0 || den <= 0.0)
          cosTheta = 1.0 - 2.0 * r;
        else
          cosTheta = fmax(-1.0, fmin(1.0, (Ap - den * pow(num / den, r)) / B));
      }
      kNew = kF;
    }
    Vec3 k = randomDirectionWrtK<true>(p.k, cosTheta, uniform01(rng.raw<RNG_MODE>()));
    const double kCurr = sqrt(sqNorm<true>(k));
    if (kCurr > 0.0) {
      const double s = kNew / kCurr;
      k.x = k.x * s;
      k.y = k.y * s;
      k.z = k.z * s;
    }
    p.k = k;
    if (mech.bath >= 0) {
      const Vec3 q = Vec3{k.x - kOld.x, k.y - kOld.y, k.z - kOld.z};
      bathRecord(baths, mech.bath, sqrt(sqNorm<true>(q)), emission);
    }
    break;
  }
  default:
    break;
  }
}

} // namespace emc
