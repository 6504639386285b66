/* TEST INFRASTRUCTURE (oracle/) -- see emc_oracle.h for the rules.
 *
 * Plain-C restatement of the reference algorithm for the per-time-step
 * particle loop.  Operation ORDER follows the reference expression by
 * expression (left-to-right as g++ evaluates them, no FMA contraction:
 * compile with -ffp-contract=off) so that results are bit-identical to the
 * reference built with the same compiler flags.
 */
#include "emc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* reference: include/emcConstants.hpp:9-27 (literal, non-CODATA values) */
static const double C_PI = 3.14159265358979323846;
static const double C_Q = 1.60219e-19;
static const double C_KB = 1.38066e-23;
static const double C_HBAR = 1.05459e-34;
static const double C_EPS0 = 8.85419e-12;
static const double C_ME = 9.11e-31;

/* ------------------------------------------------------------------ RNG */
/* std::mt19937_64 (reference: include/emcUtil.hpp:15).  Published MT19937-64
 * recurrence (Matsumoto & Nishimura); pinned by the 10000th output of the
 * default seed 5489 == 9981545732273789042 (C++ standard [rand.predef]). */
#define MT_N 312
#define MT_M 156
void orc_mt_seed(uint64_t *st, uint64_t seed) {
  st[0] = seed;
  for (int i = 1; i < MT_N; i++)
    st[i] = 6364136223846793005ULL * (st[i - 1] ^ (st[i - 1] >> 62)) + (uint64_t)i;
  st[MT_N] = MT_N;
}
uint64_t orc_mt_next(uint64_t *st) {
  if (st[MT_N] >= MT_N) {
    const uint64_t UM = 0xFFFFFFFF80000000ULL, LM = 0x7FFFFFFFULL;
    for (int i = 0; i < MT_N; i++) {
      uint64_t x = (st[i] & UM) | (st[(i + 1) % MT_N] & LM);
      uint64_t xa = x >> 1;
      if (x & 1ULL)
        xa ^= 0xB5026F5AA96619E9ULL;
      st[i] = st[(i + MT_M) % MT_N] ^ xa;
    }
    st[MT_N] = 0;
  }
  uint64_t y = st[st[MT_N]++];
  y ^= (y >> 29) & 0x5555555555555555ULL;
  y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
  y ^= (y << 37) & 0xFFF7EEE000000000ULL;
  y ^= (y >> 43);
  return y;
}
void orc_mt_fill(uint64_t seed, uint64_t *out, int64_t n) {
  uint64_t st[MT_N + 1];
  orc_mt_seed(st, seed);
  for (int64_t i = 0; i < n; i++)
    out[i] = orc_mt_next(st);
}

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the
 * B200 path.  Not part of the reference; restated here so that CPU and GPU
 * consume identical streams in PHILOX mode. */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* draw `idx` of (particle, step): counter = (particle lo, particle hi, step,
 * idx/2); word pair idx%2 of the 128-bit block, low word first. */
uint64_t orc_philox_draw(uint64_t seed, uint64_t particle, uint64_t step, uint64_t idx) {
  uint32_t ctr[4] = {(uint32_t)particle, (uint32_t)(particle >> 32),
                     (uint32_t)step, (uint32_t)(idx >> 1)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  orc_philox4x32(ctr, key, o);
  return (idx & 1) ? ((uint64_t)o[3] << 32 | o[2]) : ((uint64_t)o[1] << 32 | o[0]);
}

/* libstdc++ std::uniform_real_distribution<double>(a,b) on a 64-bit engine:
 * one raw draw, generate_canonical = double(x) / 2^64, clamped below 1
 * (SURVEY App. A.2 [probe]); then u*(b-a)+a. */
double orc_uniform(uint64_t raw, double a, double b) {
  double u = (double)raw / 18446744073709551616.0;
  if (u >= 1.0)
    u = nextafter(1.0, 0.0);
  return u * (b - a) + a;
}

typedef struct {
  const orc_rng_cfg_t *cfg;
  int64_t particle; /* local index */
  uint64_t step;
  uint64_t k; /* draws consumed by this particle in this step (PHILOX) */
  int32_t *recPid;
  int64_t recCap, recCount;
} rng_t;

static uint64_t rng_raw(rng_t *r) {
  uint64_t x;
  switch (r->cfg->mode) {
  case ORC_RNG_MT_GLOBAL:
    x = orc_mt_next(r->cfg->mtState);
    break;
  case ORC_RNG_STREAMS: {
    int64_t p = r->particle;
    x = r->cfg->draws[r->cfg->offsets[p] + r->cfg->cursor[p]++];
    break;
  }
  default:
    x = orc_philox_draw(r->cfg->philoxSeed,
                        (uint64_t)(r->cfg->particleIdBase + r->particle), r->step, r->k++);
  }
  if (r->recPid && r->recCount < r->recCap)
    r->recPid[r->recCount] = (int32_t)r->particle;
  r->recCount++;
  return x;
}
static double rng_u01(rng_t *r) { return orc_uniform(rng_raw(r), 0., 1.); }
static double rng_ulog(rng_t *r) { return orc_uniform(rng_raw(r), 1e-6, 1.); }

/* ---------------------------------------------------------- valley math */
static double sq3(const double k[3]) {
  /* emcUtil.hpp:26-31: res = 0; res += e*e in order */
  double res = 0;
  res += k[0] * k[0];
  res += k[1] * k[1];
  res += k[2] * k[2];
  return res;
}
static int is_aniso(const orc_valley_t *v) { return (v->kind & 2) != 0; }
/* the mass of the dispersion (getEnergy, getNormWaveVec, getVelocity): the conduction mass in the 3-D classes, the
 * density-of-states mass sqrt(ml mt) in emcNonParabolicAnisotropSingleLayerValley.hpp:123-148 */
static double band_mass(const orc_valley_t *v) { return v->kind == ORC_VALLEY_NONPARABOLIC_ANISO_SL ? v->mDos : v->mCond; }
static int is_nonparabolic(const orc_valley_t *v) { return v->kind & 1; }

/* emcNonParabolicAnistropValley.hpp:122, emcNonParabolicIsotropValley.hpp:99,
 * emcParabolic*Valley.hpp (gamma = E) */
double orc_gamma(const orc_valley_t *v, double e) {
  return is_nonparabolic(v) ? e * (1 + v->alpha * e) : e;
}
/* :90-92 / :60-62 ; parabolic: constant */
double orc_eff_mass_cond(const orc_valley_t *v, double e) {
  return is_nonparabolic(v) ? v->mCond * (1 + 2 * e * v->alpha) : v->mCond;
}
/* emcNonParabolicAnistropValley.hpp:109-113, emcNonParabolicIsotropValley.hpp:78-82,
 * emcParabolicIsotropValley.hpp:62-65, emcParabolicAnisotropValley.hpp:100-103 */
double orc_energy(const orc_valley_t *v, const double k[3]) {
  /* single-layer classes (...SingleLayerValley.hpp getEnergy): k[0]*k[0] + k[1]*k[1], which is sq3(k) bit for bit with
   * k[2] = 0 */
  if (is_nonparabolic(v)) {
    double gamma = C_HBAR * C_HBAR * sq3(k) / (band_mass(v) * C_Q);
    return gamma / (1 + sqrt(1 + 2 * v->alpha * gamma));
  }
  return C_HBAR * C_HBAR * sq3(k) / (2 * v->mCond * C_Q);
}
/* aniso non-parabolic :103-106 multiplies (2 m gamma q); the three others
 * multiply (2 m q gamma|E) */
double orc_norm_wave_vec(const orc_valley_t *v, double e) {
  if (v->kind == ORC_VALLEY_NONPARABOLIC_ANISO)
    return sqrt(2 * v->mCond * orc_gamma(v, e) * C_Q) / C_HBAR;
  return sqrt(2 * band_mass(v) * C_Q * orc_gamma(v, e)) / C_HBAR;
}
/* :140-153 */
void orc_to_ellipse(const orc_valley_t *v, int s, const double in[3], double out[3]) {
  if (!is_aniso(v)) {
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
    return;
  }
  const double *r = v->rot[s];
  double a = in[0], b = in[1], c = in[2];
  out[0] = a * r[0] + b * r[1] + c * r[2];
  out[1] = a * r[3] + b * r[4] + c * r[5];
  out[2] = a * r[6] + b * r[7] + c * r[8];
}
/* :157-170 */
void orc_to_device(const orc_valley_t *v, int s, const double in[3], double out[3]) {
  if (!is_aniso(v)) {
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
    return;
  }
  const double *r = v->rot[s];
  double a = in[0], b = in[1], c = in[2];
  out[0] = a * r[0] + b * r[3] + c * r[6];
  out[1] = a * r[1] + b * r[4] + c * r[7];
  out[2] = a * r[2] + b * r[5] + c * r[8];
}
/* getVelocity: aniso NP :126-136, iso NP :85-90, iso P :74-77, aniso P :106-114 */
void orc_velocity(const orc_valley_t *v, const double k[3], double e, int s, double out[3]) {
  switch (v->kind) {
  case ORC_VALLEY_NONPARABOLIC_ANISO_SL: /* emcNonParabolicAnisotropSingleLayerValley.hpp:137-148: m_DOS; z stays 0 */
  case ORC_VALLEY_NONPARABOLIC_ANISO: {
    double ke[3], ve[3];
    orc_to_ellipse(v, s, k, ke);
    double npf = sqrt(1 + 4 * v->alpha * orc_gamma(v, e));
    for (int i = 0; i < 3; i++)
      ve[i] = C_HBAR * v->vogt[i] * ke[i] / (band_mass(v) * npf);
    if (v->kind == ORC_VALLEY_NONPARABOLIC_ANISO_SL)
      ve[2] = 0;
    orc_to_device(v, s, ve, out);
    break;
  }
  case ORC_VALLEY_PARABOLIC_ANISO: {
    double ke[3], ve[3];
    orc_to_ellipse(v, s, k, ke);
    for (int i = 0; i < 3; i++)
      ve[i] = C_HBAR * v->vogt[i] * ke[i] / v->mCond;
    orc_to_device(v, s, ve, out);
    break;
  }
  case ORC_VALLEY_NONPARABOLIC_ISO_SL:
  case ORC_VALLEY_NONPARABOLIC_ISO: {
    double f = C_HBAR / (v->mCond * sqrt(1 + 4 * v->alpha * orc_gamma(v, e)));
    for (int i = 0; i < 3; i++)
      out[i] = k[i] * f;
    break;
  }
  default: {
    double f = C_HBAR / v->mCond;
    for (int i = 0; i < 3; i++)
      out[i] = k[i] * f;
  }
  }
}

/* include/emcParticleDrift.hpp:12-36 */
void orc_drift(const orc_valley_t *v, double dt, double k[3], double *energy, int s,
               double pos[3], int dim, const double force[3]) {
  double kOld[3], kNew[3], fE[3], dP[3], dPd[3];
  orc_to_ellipse(v, s, k, kOld);
  memcpy(kNew, kOld, sizeof kNew);
  orc_to_ellipse(v, s, force, fE);
  for (int i = 0; i < 3; i++)
    kNew[i] += fE[i] * dt * v->vogt[i] / C_HBAR;
  orc_to_device(v, s, kNew, k);
  *energy = orc_energy(v, k);
  for (int i = 0; i < 3; i++) {
    double avgK = (kNew[i] + kOld[i]) / 2;
    dP[i] = C_HBAR * v->vogt[i] * avgK * dt / orc_eff_mass_cond(v, *energy);
  }
  orc_to_device(v, s, dP, dPd);
  for (int i = 0; i < dim; i++)
    pos[i] = pos[i] + dPd[i];
}

/* include/emcUtil.hpp:131-139 */
void orc_random_direction(double norm, double rand1, double rand2, double out[3]) {
  double phi = 2 * C_PI * rand1;
  double cosTheta = 1 - 2 * rand2;
  out[0] = norm * sqrt(1 - cosTheta * cosTheta) * cos(phi);
  out[1] = norm * sqrt(1 - cosTheta * cosTheta) * sin(phi);
  out[2] = norm * cosTheta;
}
/* include/emcUtil.hpp:143-175 */
void orc_random_direction_wrt_k(const double k[3], double cosTheta, double rnd, double out[3]) {
  double kxy = sqrt(k[0] * k[0] + k[1] * k[1]);
  double normK = sqrt(kxy * kxy + k[2] * k[2]);
  if (normK == 0.) {
    out[0] = out[1] = out[2] = 0.;
    return;
  }
  double ct0 = k[2] / normK;
  double st0 = kxy / normK;
  double cfi0 = (kxy > 0.) ? k[0] / kxy : 1.;
  double sfi0 = (kxy > 0.) ? k[1] / kxy : 0.;
  double st = sqrt(1.0 - cosTheta * cosTheta);
  double phi = 2.0 * C_PI * rnd;
  double kxp = normK * st * cos(phi);
  double kyp = normK * st * sin(phi);
  double kzp = normK * cosTheta;
  out[0] = kxp * cfi0 * ct0 - kyp * sfi0 + kzp * cfi0 * st0;
  out[1] = kxp * sfi0 * ct0 + kyp * cfi0 + kzp * sfi0 * st0;
  out[2] = -kxp * st0 + kzp * ct0;
}

/* ------------------------------------------------------------------ model */
enum { MK_ACOUSTIC = 1, MK_ZERO = 2, MK_FIRST = 3, MK_COULOMB = 4, MK_FROEHLICH = 5, MK_ACOUSTIC_SL = 6, MK_ZERO_SL = 7, MK_FIRST_SL = 8, MK_FROEHLICH_SL = 9, MK_PIEZO_SL = 10,
       MK_IMPURITY_SL = 11, MK_ROUGHNESS_SL = 12, MK_REMOTE_SO_SL = 13, MK_SCREENED_OPTICAL_SL = 14 };

typedef struct {
  int kind, valley, finalValley, region, emission, nFinal, nInitSub;
  double scatterConst, scatterConst2, phononEnergy, regionDoping;
  int32_t finalSub[ORC_MAX_SUB][ORC_MAX_FINAL];
  /* Froehlich family */
  int variant, bath, qResolved, qResolvedAngle;
  double effMass, nBose;
  /* long-range single-layer mechanisms: form-factor width [m], 2-D screening wave vector [1/m] */
  double slWidth, slQs;
  /* the other angle-resolved single-layer mechanisms: sl[0], sl[1] as documented at ORC_SAMPLER_SL_CHARGED_IMPURITY ... */
  double sl[2];
} mech_t;

/* emcPhononBath.hpp */
struct orc_bath {
  int nBins;
  double dq, tauLO, N0, Vsim;
  double *Nq, *nEm, *nAbs, *cumW, *cumWN;
  double qs2;
  double loE, latT;
  int hasAc, hasRidley;
  double acE, tauAc, Nac, NacEq, wRidley, toE, tauTO, Nto, NtoEq;
};

typedef struct {
  int valley, region, nMech;
  int mechIdx[64];
  double *cum; /* [nMech][nLevels] */
  double tau;
} tableset_t;

struct orc_model {
  int nLevels;
  double maxEnergy, dE, temperature, rho, vSound;
  int nValleys;
  orc_valley_t valleys[16];
  int nMech;
  mech_t mech[256];
  int nSets;
  tableset_t sets[64];
  int nBaths;
  orc_bath_t *baths[16];
  double qs2; /* emcPlasmonScreening::getQs2() */
  int hasGrain;
  double grainProb, grainTau; /* grainTau = 1 / rate (emcScatterHandler.hpp:234), 1 s without a mechanism */
  double initEnergy; /* > 0: mono-energetic initial ensemble (emcElectron / emcHole initEnergyEV), 0: Maxwellian */
  int electron2D;    /* > 0: examples/singleLayerMoS2/electron2D.hpp -- that many particles per grid point, in-plane k */
};

orc_model_t *orc_model_create(int nLevels, double maxEnergy, double temperature,
                              double rho, double vSound) {
  orc_model_t *m = (orc_model_t *)calloc(1, sizeof *m);
  m->nLevels = nLevels;
  m->maxEnergy = maxEnergy;
  m->dE = maxEnergy / nLevels; /* emcScatterHandler.hpp:76 */
  m->temperature = temperature;
  m->rho = rho;
  m->vSound = vSound;
  m->grainTau = 1.;
  return m;
}
/* emcElectron.hpp:85-88 / emcHole.hpp:95-98: initParticleKSpaceFixed instead of the Maxwellian (no energy draw) */
void orc_model_set_init_energy(orc_model_t *m, double energyEV) { m->initEnergy = energyEV; }
/* examples/singleLayerMoS2/electron2D.hpp:38-67 as the particle type of the initial ensemble */
void orc_model_set_electron2d(orc_model_t *m, int perGridPoint) { m->electron2D = perGridPoint; }
void orc_model_destroy(orc_model_t *m) {
  if (!m)
    return;
  for (int i = 0; i < m->nSets; i++)
    free(m->sets[i].cum);
  free(m);
}

/* valley constructors: emcNonParabolicAnistropValley.hpp:54-80,185-204 etc. */
int orc_add_valley(orc_model_t *m, int kind, const double relMass[3], double particleMass,
                   int deg, double alpha, double eBottom, const double *dirs) {
  if (m->nValleys >= 16 || deg > ORC_MAX_SUB)
    return -1;
  orc_valley_t *v = &m->valleys[m->nValleys];
  memset(v, 0, sizeof *v);
  v->kind = kind;
  v->deg = deg;
  v->alpha = (kind & 1) ? alpha : 0.;
  v->eBottom = eBottom;
  for (int s = 0; s < ORC_MAX_SUB; s++)
    v->rot[s][0] = v->rot[s][4] = v->rot[s][8] = 1.;
  if (kind == ORC_VALLEY_NONPARABOLIC_ANISO_SL) {
    /* emcNonParabolicAnisotropSingleLayerValley.hpp:79-113: relMass = {longitudinal, transversal, -}; dirs = one
     * rotation angle per sub-valley (first entry of each 9-block); frame rows (cos, -sin, 0), (sin, cos, 0), (0, 0, 0) */
    const double mL = relMass[0] * particleMass, mT = relMass[1] * particleMass;
    v->mCond = 2. / (1. / mL + 1. / mT);
    v->mDos = sqrt(mL * mT);
    v->vogt[0] = sqrt(v->mDos / mL);
    v->vogt[1] = sqrt(v->mDos / mT);
    v->vogt[2] = 0;
    for (int s = 0; s < deg; s++) {
      const double angle = dirs ? dirs[s * 9] : 0.;
      const double sn = sin(angle), cs = cos(angle);
      double *r = v->rot[s];
      memset(r, 0, 9 * sizeof(double));
      r[0] = cs; r[1] = -sn;
      r[3] = sn; r[4] = cs;
    }
  } else if (kind == ORC_VALLEY_PARABOLIC_ISO_SL || kind == ORC_VALLEY_NONPARABOLIC_ISO_SL) {
    v->mCond = v->mDos = relMass[0] * particleMass;
    v->vogt[0] = v->vogt[1] = 1.;
    v->vogt[2] = 0.;
  } else if (kind >= 2) {
    double prod = 1.;
    for (int i = 0; i < 3; i++)
      prod = prod * relMass[i];
    v->mDos = pow(prod, 1. / 3.) * particleMass;
    double acc = 0.;
    for (int i = 0; i < 3; i++)
      acc += 1. / relMass[i];
    v->mCond = 3. * particleMass / acc;
    for (int i = 0; i < 3; i++)
      v->vogt[i] = sqrt(orc_eff_mass_cond(v, 0.) / (relMass[i] * particleMass));
    if (dirs) {
      for (int s = 0; s < deg; s++)
        for (int r = 0; r < 3; r++) {
          const double *d = dirs + (s * 3 + r) * 3;
          double nrm = sqrt(sq3(d)); /* emcUtil.hpp:34-47 */
          for (int c = 0; c < 3; c++)
            v->rot[s][r * 3 + c] = (nrm == 0.) ? d[c] : d[c] / nrm;
        }
    }
  } else {
    v->mCond = v->mDos = relMass[0] * particleMass;
    v->vogt[0] = v->vogt[1] = v->vogt[2] = 1.;
  }
  return m->nValleys++;
}

static double dos_mass_at_zero(const orc_valley_t *v) {
  /* getEffMassDOS(0): non-parabolic = m*pow(1+2*alpha*0, 3.) */
  return is_nonparabolic(v) ? v->mDos * pow(1 + 2 * v->alpha * 0., 3.) : v->mDos;
}

/* emcAcousticScatterMechanism.hpp:47-56 */
int orc_add_acoustic(orc_model_t *m, int valley, int region, double sigma) {
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = MK_ACOUSTIC;
  x->valley = x->finalValley = valley;
  x->region = region;
  double cL = m->rho * pow(m->vSound, 2);
  x->scatterConst = sqrt(2.0 * C_Q) * pow(sigma * C_Q, 2) * C_KB * m->temperature /
                    (C_PI * cL * pow(C_HBAR, 4));
  return m->nMech++;
}

/* emcZeroOrderInterValleyScatterMechanism.hpp:12-21, emcFirstOrder...:12-20 */
int orc_add_intervalley(orc_model_t *m, int order, int emission, int valley, int finalValley,
                        int region, double defPot, double phE, int nInitSub, int nFinal,
                        const int32_t *finalSub) {
  if (nFinal > ORC_MAX_FINAL || nInitSub > ORC_MAX_SUB)
    return -1;
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = order == 0 ? MK_ZERO : MK_FIRST;
  x->valley = valley;
  x->finalValley = finalValley;
  x->region = region;
  x->emission = emission;
  x->nFinal = nFinal;
  x->nInitSub = nInitSub;
  x->phononEnergy = phE;
  for (int i = 0; i < nInitSub; i++)
    for (int j = 0; j < nFinal; j++)
      x->finalSub[i][j] = finalSub[i * nFinal + j];
  double result;
  if (order == 0)
    result = nFinal * sqrt(C_Q) * pow(defPot / C_HBAR, 2) * C_Q /
             (C_PI * m->rho * phE * sqrt(2));
  else
    result = nFinal * sqrt(2) * pow(C_Q, 5. / 2.) * pow(defPot, 2) /
             (C_PI * m->rho * pow(C_HBAR, 4) * phE);
  double nrPh = 1. / (exp(C_Q * phE / (C_KB * m->temperature)) - 1.);
  x->scatterConst = emission ? result * (nrPh + 1) : result * nrPh;
  return m->nMech++;
}

/* emcAcousticSingleLayerScatterMechanism.hpp:42-50 */
int orc_add_acoustic_sl(orc_model_t *m, int valley, int region, double sigma, double density2D, double vSound) {
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = MK_ACOUSTIC_SL;
  x->valley = x->finalValley = valley;
  x->region = region;
  double cL = density2D * pow(vSound, 2);
  x->scatterConst = pow(sigma * C_Q, 2) * C_KB * m->temperature / (cL * pow(C_HBAR, 3));
  return m->nMech++;
}

/* emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp:66-88 (absorption), :243-266 (emission); nFinal = 0: the
 * one-valley constructor (:42-49): nrFinalValleys = 1 and no sub-valley draw */
int orc_add_intervalley_sl(orc_model_t *m, int order, int emission, int valley, int finalValley, int region, double sigma,
                           double density2D, double phE, int nInitSub, int nFinal, const int32_t *finalSub) {
  if (nFinal > ORC_MAX_FINAL || nInitSub > ORC_MAX_SUB)
    return -1;
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = order == 0 ? MK_ZERO_SL : MK_FIRST_SL;
  x->valley = valley;
  x->finalValley = finalValley;
  x->region = region;
  x->emission = emission;
  x->nFinal = nFinal;
  x->nInitSub = nInitSub;
  x->phononEnergy = phE;
  for (int i = 0; i < nInitSub; i++)
    for (int j = 0; j < nFinal; j++)
      x->finalSub[i][j] = finalSub[i * nFinal + j];
  const double nrFinal = nFinal > 0 ? (double)nFinal : 1.;
  double exponent = phE * C_Q / (C_KB * m->temperature);
  double omega = phE * C_Q / C_HBAR;
  if (order != 0) {
    /* emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp:82-88 (absorption), :229-236 (emission); sigma in eV */
    if (emission)
      x->scatterConst = nrFinal * pow(sigma * C_Q, 2) * C_Q * exp(exponent) /
                        (density2D * omega * (exp(exponent) - 1) * pow(C_HBAR, 4));
    else
      x->scatterConst = nrFinal * pow(sigma * C_Q, 2) * C_Q / (density2D * omega * (exp(exponent) - 1) * pow(C_HBAR, 4));
  } else if (emission)
    x->scatterConst = nrFinal * pow(sigma * C_Q / C_HBAR, 2) * exp(exponent) / (2 * density2D * omega * (exp(exponent) - 1));
  else
    x->scatterConst = nrFinal * pow(sigma * C_Q / C_HBAR, 2) / (2 * density2D * omega * (exp(exponent) - 1));
  return m->nMech++;
}

/* ---- long-range single-layer mechanisms (Froehlich, piezoelectric): emc2DScreening.hpp:65-76 and the angular weights */
static double sl_screening_factor(double q, double qs) {
  double eps = (q <= 0 || qs <= 0) ? 1. : 1. + qs / q;
  return 1. / (eps * eps);
}
/* emcFroehlichInteractionSingleLayer.hpp:48-56: weight of the deflection angle psi */
static double sl_froehlich_weight(double psi, double k, double kPrime, double width, double qs) {
  double q2 = k * k + kPrime * kPrime - 2 * k * kPrime * cos(psi);
  double q = sqrt(q2 > 0 ? q2 : 0);
  if (q <= 0)
    return 0;
  double ff = erfc(width * q / 2);
  return ff * ff * sl_screening_factor(q, qs) / q;
}
/* emcPiezoelectricSingleLayerScatterMechanism.hpp:62-66 */
static double sl_piezo_weight(double theta, double k, double width, double qs) {
  double q = 2 * k * sin(theta / 2);
  double ff = erfc(width * q / 2);
  return ff * ff * sl_screening_factor(q, qs);
}
/* rate integrands of the Froehlich classes (:190-197 absorption, :303-318 emission) and their midpoint rule (:201-208) */
static double sl_froehlich_integrand(int emission, double theta, double k, double eFactor, double width, double qs) {
  double cosTheta = cos(theta);
  if (!emission) {
    double root = sqrt(cosTheta * cosTheta + eFactor);
    double q = k * (-cosTheta + root);
    return (-cosTheta + root) / root * pow(erfc(width * q / 2.), 2) * sl_screening_factor(q, qs);
  }
  double root = sqrt(cosTheta * cosTheta - eFactor);
  double qPlus = k * (cosTheta + root);
  double partPlus = (cosTheta + root);
  partPlus *= pow(erfc(width * qPlus / 2.), 2) * sl_screening_factor(qPlus, qs);
  double qMinus = k * (cosTheta - root);
  double partMinus = (cosTheta - root);
  partMinus *= pow(erfc(width * qMinus / 2.), 2) * sl_screening_factor(qMinus, qs);
  return (partPlus + partMinus) / root;
}
static double sl_froehlich_integral(int emission, double a, double b, int n, double k, double eFactor, double width, double qs) {
  double dx = (b - a) / (double)n;
  double result = 0;
  for (double x = a + dx / 2.; x <= b - dx / 2.; x += dx)
    result += sl_froehlich_integrand(emission, x, k, eFactor, width, qs);
  return result * dx;
}
/* emcFroehlichInteractionSingleLayer.hpp:120-132 (absorption), :239-251 (emission) */
int orc_add_froehlich_sl(orc_model_t *m, int emission, int valley, int region, double phE, double couplingConst, double width,
                         double qs) {
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = MK_FROEHLICH_SL;
  x->valley = x->finalValley = valley;
  x->region = region;
  x->emission = emission;
  x->phononEnergy = phE;
  x->slWidth = width;
  x->slQs = qs;
  double exponent = C_Q * phE / (C_KB * m->temperature);
  double nrPhonons = emission ? exp(exponent) / (exp(exponent) - 1.) : 1. / (exp(exponent) - 1.);
  x->scatterConst = pow(couplingConst * C_Q, 2) * nrPhonons / (2 * C_PI * pow(C_HBAR, 3));
  return m->nMech++;
}
/* emcPiezoelectricSingleLayerScatterMechanism.hpp:71-86 */
int orc_add_piezo_sl(orc_model_t *m, int valley, int region, double piezoConst, double width, double density2D, double vSound,
                     double qs) {
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = MK_PIEZO_SL;
  x->valley = x->finalValley = valley;
  x->region = region;
  x->slWidth = width;
  x->slQs = qs;
  double couplingEnergy = piezoConst * C_Q / C_EPS0;
  x->scatterConst = 0.5 * couplingEnergy * couplingEnergy * C_KB * m->temperature / (density2D * vSound * vSound * pow(C_HBAR, 3));
  return m->nMech++;
}

/* ---- the other angle-resolved single-layer mechanisms.  Weights: emc2DChargedImpurityScatterMechanism.hpp:65-72,
 * emcSurfaceRoughnessScatterMechanism.hpp:59-63, emcRemoteSurfaceOpticalPhononMechanism.hpp:63-70,
 * emcScreenedIntravalleyOpticalMechanism.hpp:58-62.  par = {p0, p1, q_s} as in the sampler description (emc_oracle.h) */
static int sl_angular_steps(int sampler) {
  return sampler == ORC_SAMPLER_SL_CHARGED_IMPURITY ? 512 : sampler == ORC_SAMPLER_SL_SURFACE_ROUGHNESS ? 256 : 128;
}
static double sl_angular_weight(int sampler, double theta, double k, double kPrime, const double *par) {
  switch (sampler) {
  case ORC_SAMPLER_SL_CHARGED_IMPURITY: {
    double q = 2 * k * sin(theta / 2);
    double denom = par[2] + q + par[1] * q * q;
    if (denom <= 0)
      return 0;
    double v = exp(-q * par[0]) / denom;
    return v * v;
  }
  case ORC_SAMPLER_SL_SURFACE_ROUGHNESS: {
    double q = 2 * k * sin(theta / 2);
    double formFactor = exp(-q * q * par[1] / 4);
    return formFactor * sl_screening_factor(q, par[2]);
  }
  case ORC_SAMPLER_SL_REMOTE_SO: {
    double q2 = k * k + kPrime * kPrime - 2 * k * kPrime * cos(theta);
    double q = sqrt(q2 > 0 ? q2 : 0);
    if (q <= 0)
      return 0;
    return exp(-2 * q * par[1]) * sl_screening_factor(q, par[2]) / q;
  }
  default: { /* ORC_SAMPLER_SL_SCREENED_OPTICAL */
    double q2 = k * k + kPrime * kPrime - 2 * k * kPrime * cos(theta);
    double q = sqrt(q2 > 0 ? q2 : 0);
    return sl_screening_factor(q, par[2]);
  }
  }
}
static int sl_angular_sampler_of(int kind) {
  return kind == MK_IMPURITY_SL ? ORC_SAMPLER_SL_CHARGED_IMPURITY : kind == MK_ROUGHNESS_SL ? ORC_SAMPLER_SL_SURFACE_ROUGHNESS
         : kind == MK_REMOTE_SO_SL ? ORC_SAMPLER_SL_REMOTE_SO : ORC_SAMPLER_SL_SCREENED_OPTICAL;
}
static mech_t *sl_new_mech(orc_model_t *m, int kind, int valley, int region) {
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = kind;
  x->valley = x->finalValley = valley;
  x->region = region;
  return x;
}
/* emc2DChargedImpurityScatterMechanism.hpp:77-92 */
int orc_add_charged_impurity_sl(orc_model_t *m, int valley, int region, double impurityDensity, double epsAvg, double qs,
                                double rytovaKeldyshLength, double remoteDistance, double chargeNumber) {
  mech_t *x = sl_new_mech(m, MK_IMPURITY_SL, valley, region);
  x->sl[0] = remoteDistance;
  x->sl[1] = rytovaKeldyshLength;
  x->slQs = qs;
  double A = chargeNumber * C_Q * C_Q / (2 * C_EPS0 * epsAvg);
  x->scatterConst = impurityDensity * A * A / (C_PI * pow(C_HBAR, 3));
  return m->nMech++;
}
/* emcSurfaceRoughnessScatterMechanism.hpp:68-79 */
int orc_add_surface_roughness_sl(orc_model_t *m, int valley, int region, double effectiveField, double roughnessAmplitude,
                                 double correlationLength, double qs) {
  mech_t *x = sl_new_mech(m, MK_ROUGHNESS_SL, valley, region);
  x->sl[1] = correlationLength * correlationLength;
  x->slQs = qs;
  double eF = C_Q * effectiveField;
  x->scatterConst = eF * eF * roughnessAmplitude * roughnessAmplitude * x->sl[1] / pow(C_HBAR, 3);
  return m->nMech++;
}
/* emcRemoteSurfaceOpticalPhononMechanism.hpp:75-91 */
int orc_add_remote_so_sl(orc_model_t *m, int emission, int valley, int region, double phE, double couplingD,
                         double remoteDistance, double qs) {
  mech_t *x = sl_new_mech(m, MK_REMOTE_SO_SL, valley, region);
  x->emission = emission;
  x->phononEnergy = phE;
  x->sl[1] = remoteDistance;
  x->slQs = qs;
  double omega = phE * C_Q / C_HBAR;
  double C = C_Q * C_Q * omega * couplingD / (4 * C_PI * C_EPS0 * C_HBAR * C_HBAR);
  double xx = phE * C_Q / (C_KB * m->temperature);
  double nBose = 1. / (exp(xx) - 1.);
  x->scatterConst = C * (emission ? nBose + 1 : nBose);
  return m->nMech++;
}
/* emcScreenedIntravalleyOpticalMechanism.hpp:67-81 */
int orc_add_screened_optical_sl(orc_model_t *m, int emission, int valley, int region, double sigma, double density2D, double phE,
                                double qs) {
  mech_t *x = sl_new_mech(m, MK_SCREENED_OPTICAL_SL, valley, region);
  x->emission = emission;
  x->phononEnergy = phE;
  x->slQs = qs;
  double xx = phE * C_Q / (C_KB * m->temperature);
  double omega = phE * C_Q / C_HBAR;
  double nBose = 1. / (exp(xx) - 1.);
  double nFactor = emission ? nBose + 1 : nBose;
  x->scatterConst = pow(sigma * C_Q / C_HBAR, 2) * nFactor / (2 * density2D * omega);
  return m->nMech++;
}

/* emcCoulombScatterMechanism.hpp:23-32 */
int orc_add_coulomb(orc_model_t *m, int valley, int region, double epsR, double regionDoping) {
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = MK_COULOMB;
  x->valley = x->finalValley = valley;
  x->region = region;
  double epsMat = C_EPS0 * epsR;
  double Vt = C_KB / C_Q * m->temperature; /* emcDevice.hpp:87 */
  x->scatterConst = sqrt(2 * C_Q) * pow(C_KB * m->temperature, 2) / (C_PI * pow(C_HBAR, 4));
  x->scatterConst2 = 8 * epsMat * Vt / (C_HBAR * C_HBAR);
  x->regionDoping = fabs(regionDoping);
  return m->nMech++;
}

static double delta_valley(const orc_model_t *m, const mech_t *x) {
  return m->valleys[x->finalValley].eBottom - m->valleys[x->valley].eBottom;
}

/* emcScreenedFroehlichInteraction.hpp:54-63 */
static double screened_log_factor(double kI, double kF, double qs2) {
  const double qPlus2 = (kI + kF) * (kI + kF);
  const double qMinus2 = (kI - kF) * (kI - kF);
  if (qs2 <= 0) {
    if (qMinus2 <= 0)
      return 0;
    return 0.5 * log(qPlus2 / qMinus2);
  }
  return 0.5 * log((qPlus2 + qs2) / (qMinus2 + qs2));
}

/* emcFroehlichInteraction.hpp:33-50 */
int orc_add_froehlich(orc_model_t *m, int variant, int emission, int valley, int region, double phononEnergy,
                      double relEffMass, double epsHi, double epsLo, double temperature, int bath, int qResolved,
                      int qResolvedAngle) {
  if ((variant == 1 || variant == 3) && (bath < 0 || bath >= m->nBaths))
    return -1;
  mech_t *x = &m->mech[m->nMech];
  memset(x, 0, sizeof *x);
  x->kind = MK_FROEHLICH;
  x->variant = variant;
  x->valley = x->finalValley = valley;
  x->region = region;
  x->emission = emission;
  x->phononEnergy = phononEnergy;
  x->effMass = relEffMass * C_ME;
  x->bath = (variant == 1 || variant == 3) ? bath : -1;
  x->qResolved = qResolved;
  x->qResolvedAngle = qResolvedAngle;
  double omega0 = phononEnergy * C_Q / C_HBAR;
  x->scatterConst = C_Q * C_Q * omega0 * x->effMass / (4. * C_PI * C_HBAR * C_HBAR) * (1. / epsHi - 1. / epsLo) / C_EPS0;
  double xx = C_Q * phononEnergy / (C_KB * temperature);
  x->nBose = 1. / (exp(xx) - 1.);
  return m->nMech++;
}
void orc_model_set_grain(orc_model_t *m, double transmissionProb, double scatterRate) {
  m->hasGrain = scatterRate > 0;
  m->grainProb = transmissionProb;
  m->grainTau = m->hasGrain ? 1. / scatterRate : 1.;
}
int orc_model_add_bath(orc_model_t *m, orc_bath_t *b) {
  if (m->nBaths >= 16)
    return -1;
  m->baths[m->nBaths] = b;
  return m->nBaths++;
}
void orc_model_set_qs2(orc_model_t *m, double qs2) { m->qs2 = qs2; }
double orc_plasmon_qs2(double density, double carrierTemp, double epsStatic) {
  if (density <= 0 || carrierTemp <= 0)
    return 0;
  return density * C_Q * C_Q / (epsStatic * C_EPS0 * C_KB * carrierTemp);
}

/* getScatterRate of each built-in mechanism */
double orc_raw_rate(const orc_model_t *m, int g, double energy) {
  const mech_t *x = &m->mech[g];
  const orc_valley_t *vi = &m->valleys[x->valley];
  const orc_valley_t *vf = &m->valleys[x->finalValley];
  switch (x->kind) {
  case MK_ACOUSTIC: { /* emcAcousticScatterMechanism.hpp:60-67 */
    double md = dos_mass_at_zero(vi);
    double alpha = vi->alpha;
    double gamma = orc_gamma(vi, energy);
    return x->scatterConst * pow(md, 3. / 2.) * sqrt(gamma) * (2 * alpha * energy + 1.0);
  }
  case MK_ZERO: { /* emcZeroOrder...:103-117, :246-260 */
    double dV = delta_valley(m, x);
    double ef = x->emission ? energy - x->phononEnergy - dV : energy + x->phononEnergy - dV;
    if (ef > 0) {
      double md = dos_mass_at_zero(vf);
      double alpha = vf->alpha;
      double gamma = orc_gamma(vf, ef);
      return x->scatterConst * pow(md, 3. / 2.) * sqrt(gamma) * (2 * alpha * ef + 1.0);
    }
    return 0;
  }
  case MK_ACOUSTIC_SL: { /* emcAcousticSingleLayerScatterMechanism.hpp:55-60 */
    double md = dos_mass_at_zero(vi);
    double alpha = vi->alpha;
    return md * x->scatterConst * (1 + 2 * alpha * energy);
  }
  case MK_ZERO_SL: { /* emcZeroOrderSingleLayer...:96-108, :274-286 */
    double dV = delta_valley(m, x);
    double ef = x->emission ? energy - x->phononEnergy - dV : energy + x->phononEnergy - dV;
    if (ef > 0) {
      double md = dos_mass_at_zero(vf);
      double alpha = vf->alpha;
      return md * x->scatterConst * (1 + 2 * alpha * ef);
    }
    return 0;
  }
  case MK_FROEHLICH_SL: { /* emcFroehlichInteractionSingleLayer.hpp:138-145, :257-269 */
    if (x->emission && !(energy > x->phononEnergy))
      return 0;
    double md = dos_mass_at_zero(vi);
    double k = orc_norm_wave_vec(vi, energy);
    double eFactor = x->phononEnergy / energy;
    double integral;
    if (x->emission) {
      double thetaMax = acos(sqrt(eFactor));
      integral = sl_froehlich_integral(1, -thetaMax, thetaMax, 10000, k, eFactor, x->slWidth, x->slQs);
    } else {
      integral = sl_froehlich_integral(0, 0., 2 * C_PI, 10000, k, eFactor, x->slWidth, x->slQs);
    }
    return integral * x->scatterConst * md;
  }
  case MK_IMPURITY_SL:   /* emc2DChargedImpurityScatterMechanism.hpp:96-105 */
  case MK_ROUGHNESS_SL: { /* emcSurfaceRoughnessScatterMechanism.hpp:83-92 */
    const int sampler = sl_angular_sampler_of(x->kind), n = sl_angular_steps(sampler);
    const double par[3] = {x->sl[0], x->sl[1], x->slQs};
    double mc = orc_eff_mass_cond(vi, energy);
    double k = orc_norm_wave_vec(vi, energy);
    double dtheta = C_PI / n;
    double integral = 0;
    for (int i = 0; i < n; ++i)
      integral += sl_angular_weight(sampler, (i + 0.5) * dtheta, k, k, par);
    integral *= dtheta;
    return x->scatterConst * mc * integral;
  }
  case MK_REMOTE_SO_SL: { /* emcRemoteSurfaceOpticalPhononMechanism.hpp:97-110 */
    if (x->emission && energy <= x->phononEnergy)
      return 0;
    const double par[3] = {0, x->sl[1], x->slQs};
    double finalEnergy = x->emission ? energy - x->phononEnergy : energy + x->phononEnergy;
    double k = orc_norm_wave_vec(vi, energy);
    double kPrime = orc_norm_wave_vec(vi, finalEnergy);
    double mc = orc_eff_mass_cond(vi, finalEnergy);
    double dtheta = C_PI / 128;
    double integral = 0;
    for (int i = 0; i < 128; ++i)
      integral += sl_angular_weight(ORC_SAMPLER_SL_REMOTE_SO, (i + 0.5) * dtheta, k, kPrime, par);
    integral *= dtheta;
    return 2 * x->scatterConst * mc * integral;
  }
  case MK_SCREENED_OPTICAL_SL: { /* emcScreenedIntravalleyOpticalMechanism.hpp:88-102 */
    if (x->emission && energy <= x->phononEnergy)
      return 0;
    const double par[3] = {0, 0, x->slQs};
    double finalEnergy = x->emission ? energy - x->phononEnergy : energy + x->phononEnergy;
    double md = dos_mass_at_zero(vi);
    double alpha = vi->alpha;
    double k = orc_norm_wave_vec(vi, energy);
    double kPrime = orc_norm_wave_vec(vi, finalEnergy);
    double dtheta = C_PI / 128;
    double screenAvg = 0;
    for (int i = 0; i < 128; ++i)
      screenAvg += sl_angular_weight(ORC_SAMPLER_SL_SCREENED_OPTICAL, (i + 0.5) * dtheta, k, kPrime, par);
    screenAvg /= 128;
    return md * x->scatterConst * (1 + 2 * alpha * finalEnergy) * screenAvg;
  }
  case MK_PIEZO_SL: { /* emcPiezoelectricSingleLayerScatterMechanism.hpp:92-105 */
    double md = dos_mass_at_zero(vi);
    double alpha = vi->alpha;
    double k = orc_norm_wave_vec(vi, energy);
    double dtheta = C_PI / 128;
    double integral = 0;
    for (int i = 0; i < 128; ++i)
      integral += sl_piezo_weight((i + 0.5) * dtheta, k, x->slWidth, x->slQs);
    integral *= dtheta / C_PI;
    return md * x->scatterConst * (1 + 2 * alpha * energy) * integral;
  }
  case MK_FIRST_SL: { /* emcFirstOrderSingleLayer...:96-101, :244-252: mass and non-parabolicity of the INITIAL valley */
    double md = dos_mass_at_zero(vi);
    double alpha = vi->alpha;
    if (x->emission) {
      if (energy > x->phononEnergy) {
        double rate = md * md * x->scatterConst * (2 * energy - x->phononEnergy);
        return rate * (1 + 2 * alpha * energy);
      }
      return 0;
    }
    double rate = md * md * x->scatterConst * (2 * energy + x->phononEnergy);
    return rate * (1 + 2 * alpha * energy);
  }
  case MK_FIRST: { /* emcFirstOrder...:102-119, :250-267 */
    double dV = delta_valley(m, x);
    double ef = x->emission ? energy - x->phononEnergy - dV : energy + x->phononEnergy - dV;
    if (ef > 0) {
      double md = dos_mass_at_zero(vf);
      double alpha = vf->alpha;
      double gamma = orc_gamma(vi, energy);
      double gammaF = orc_gamma(vf, ef);
      return x->scatterConst * pow(md, 5. / 2.) * sqrt(gammaF) * (2 * alpha * ef + 1.0) *
             (gamma + gammaF);
    }
    return 0;
  }
  case MK_FROEHLICH: {
    /* emcFroehlichInteraction.hpp:96-107 / :168-181, emcHotPhononFroehlichMechanism.hpp:79-92 / :158-173,
     * emcScreenedFroehlichInteraction.hpp:127-138 ... :339-352 */
    if (x->emission && energy <= x->phononEnergy)
      return 0;
    double gammaI = orc_gamma(vi, energy);
    double gammaF = orc_gamma(vi, x->emission ? energy - x->phononEnergy : energy + x->phononEnergy);
    if (gammaI <= 0 || gammaF <= 0)
      return 0;
    double kI = sqrt(2 * x->effMass * gammaI * C_Q) / C_HBAR;
    double kF = sqrt(2 * x->effMass * gammaF * C_Q) / C_HBAR;
    double occ;
    if (x->variant == 0 || x->variant == 2)
      occ = x->nBose;
    else if (x->variant == 3 && x->qResolved)
      occ = orc_bath_nq_window(m->baths[x->bath], fabs(kI - kF), kI + kF);
    else
      occ = orc_bath_mean_nq(m->baths[x->bath]);
    if (x->emission)
      occ = occ + 1;
    if (x->variant < 2) {
      double lnFactor = x->emission ? log((kI + kF) / (kI - kF)) : log((kI + kF) / (kF - kI));
      return x->scatterConst * occ / kI * lnFactor;
    }
    return x->scatterConst * occ / kI * screened_log_factor(kI, kF, m->qs2);
  }
  case MK_COULOMB: { /* emcCoulombScatterMechanism.hpp:36-46 */
    double md = dos_mass_at_zero(vi);
    double mc = orc_eff_mass_cond(vi, 0.);
    double alpha = vi->alpha;
    double gamma = orc_gamma(vi, energy);
    return x->scatterConst * pow(md, 3. / 2.) / x->regionDoping * sqrt(gamma) *
           (2 * alpha * energy + 1.0) / (1 + (x->scatterConst2 * mc * gamma / x->regionDoping));
  }
  }
  return 0;
}

/* emcScatterHandler.hpp:220-273: std::map keyed (valley, region) => sets sorted
 * lexicographically; mechanisms in insertion order inside a set. */
int orc_build_tables(orc_model_t *m) {
  for (int i = 0; i < m->nSets; i++)
    free(m->sets[i].cum);
  m->nSets = 0;
  for (int g = 0; g < m->nMech; g++) {
    int found = -1;
    for (int i = 0; i < m->nSets; i++)
      if (m->sets[i].valley == m->mech[g].valley && m->sets[i].region == m->mech[g].region)
        found = i;
    if (found < 0) {
      if (m->nSets >= 64)
        return -1;
      found = m->nSets++;
      m->sets[found].valley = m->mech[g].valley;
      m->sets[found].region = m->mech[g].region;
      m->sets[found].nMech = 0;
      m->sets[found].cum = NULL;
    }
    if (m->sets[found].nMech >= 64)
      return -1;
    m->sets[found].mechIdx[m->sets[found].nMech++] = g;
  }
  /* sort sets by (valley, region) like std::map<tuple> */
  for (int i = 0; i < m->nSets; i++)
    for (int j = i + 1; j < m->nSets; j++) {
      tableset_t *a = &m->sets[i], *b = &m->sets[j];
      if (b->valley < a->valley || (b->valley == a->valley && b->region < a->region)) {
        tableset_t t = *a;
        *a = *b;
        *b = t;
      }
    }
  const int L = m->nLevels;
  for (int i = 0; i < m->nSets; i++) {
    tableset_t *s = &m->sets[i];
    s->cum = (double *)malloc(sizeof(double) * s->nMech * L);
    for (int t = 0; t < s->nMech; t++)
      for (int l = 0; l < L; l++)
        s->cum[t * L + l] = orc_raw_rate(m, s->mechIdx[t], (l + 1) * m->dE);
    for (int t = 1; t < s->nMech; t++)
      for (int l = 0; l < L; l++)
        s->cum[t * L + l] = s->cum[t * L + l] + s->cum[(t - 1) * L + l];
    double maxRate = s->cum[(s->nMech - 1) * L];
    for (int l = 1; l < L; l++)
      if (s->cum[(s->nMech - 1) * L + l] > maxRate)
        maxRate = s->cum[(s->nMech - 1) * L + l];
    for (int t = 0; t < s->nMech; t++)
      for (int l = 0; l < L; l++)
        s->cum[t * L + l] /= maxRate;
    s->tau = 1. / maxRate;
  }
  return 0;
}

int orc_n_valleys(const orc_model_t *m) { return m->nValleys; }
int orc_get_valley(const orc_model_t *m, int v, orc_valley_t *out) {
  if (v < 0 || v >= m->nValleys)
    return -1;
  *out = m->valleys[v];
  return 0;
}
int orc_n_mechanisms(const orc_model_t *m) { return m->nMech; }
int orc_n_tablesets(const orc_model_t *m) { return m->nSets; }
int orc_tableset_info(const orc_model_t *m, int i, int32_t *valley, int32_t *region,
                      int32_t *nMech, double *tau) {
  if (i < 0 || i >= m->nSets)
    return -1;
  *valley = m->sets[i].valley;
  *region = m->sets[i].region;
  *nMech = m->sets[i].nMech;
  *tau = m->sets[i].tau;
  return 0;
}
static void fill_mech_desc(const orc_model_t *m, int g, orc_mech_t *d) {
  const mech_t *x = &m->mech[g];
  memset(d, 0, sizeof *d);
  d->globalId = g;
  d->finalValley = x->finalValley;
  d->nFinal = x->nFinal;
  memcpy(d->finalSub, x->finalSub, sizeof d->finalSub);
  switch (x->kind) {
  case MK_ACOUSTIC:
    d->sampler = ORC_SAMPLER_ISOTROPIC_ELASTIC;
    break;
  case MK_ZERO:
  case MK_FIRST: {
    d->sampler = ORC_SAMPLER_INTERVALLEY;
    /* abs: E += (hw - dV); em: E -= (hw + dV)  (emcZeroOrder...:125, :268).
     * a - b == a + (-b) exactly, so a signed shift reproduces both. */
    double dV = delta_valley(m, x);
    d->p[0] = x->emission ? -(x->phononEnergy + dV) : (x->phononEnergy - dV);
    break;
  }
  case MK_ACOUSTIC_SL:
    d->sampler = ORC_SAMPLER_SL_ELASTIC;
    break;
  case MK_ZERO_SL:
  case MK_FIRST_SL: { /* abs: E += hw - dV (:127); em: E -= (dV + hw) (:304); the first-order classes alike */
    d->sampler = ORC_SAMPLER_SL_INTERVALLEY;
    double dV = delta_valley(m, x);
    d->p[0] = x->emission ? -(dV + x->phononEnergy) : (x->phononEnergy - dV);
    d->p[1] = x->kind == MK_FIRST_SL ? 1. : 0.; /* plain in-plane direction, k_z left as it is */
    break;
  }
  case MK_FROEHLICH_SL:
    d->sampler = ORC_SAMPLER_SL_FROEHLICH;
    d->p[0] = x->emission ? -x->phononEnergy : x->phononEnergy;
    d->p[1] = x->slWidth;
    d->p[2] = x->slQs;
    break;
  case MK_IMPURITY_SL:
  case MK_ROUGHNESS_SL:
  case MK_REMOTE_SO_SL:
  case MK_SCREENED_OPTICAL_SL:
    d->sampler = sl_angular_sampler_of(x->kind);
    d->p[0] = (x->kind == MK_REMOTE_SO_SL || x->kind == MK_SCREENED_OPTICAL_SL) ? (x->emission ? -x->phononEnergy : x->phononEnergy)
                                                                               : x->sl[0];
    d->p[1] = x->sl[1];
    d->p[2] = x->slQs;
    break;
  case MK_PIEZO_SL:
    d->sampler = ORC_SAMPLER_SL_PIEZO;
    d->p[1] = x->slWidth;
    d->p[2] = x->slQs;
    break;
  case MK_FROEHLICH:
    d->sampler = x->variant < 2 ? ORC_SAMPLER_FROEHLICH : ORC_SAMPLER_SCREENED_FROEHLICH;
    d->p[0] = x->emission ? -x->phononEnergy : x->phononEnergy;
    d->p[1] = m->qs2;
    d->p[2] = (double)x->bath;
    d->p[3] = (x->variant == 3 && x->qResolved && x->qResolvedAngle) ? 1. : 0.;
    break;
  case MK_COULOMB: {
    d->sampler = ORC_SAMPLER_COULOMB;
    double mc = orc_eff_mass_cond(&m->valleys[x->valley], 0.);
    d->p[0] = x->regionDoping / (x->scatterConst2 * mc); /* debyeEnergy :55 */
    break;
  }
  }
}
int orc_tableset_copy(const orc_model_t *m, int i, double *cum, orc_mech_t *mech) {
  if (i < 0 || i >= m->nSets)
    return -1;
  const tableset_t *s = &m->sets[i];
  if (cum)
    memcpy(cum, s->cum, sizeof(double) * s->nMech * m->nLevels);
  if (mech)
    for (int t = 0; t < s->nMech; t++)
      fill_mech_desc(m, s->mechIdx[t], &mech[t]);
  return 0;
}
static int find_set(const orc_model_t *m, int valley, int region) {
  for (int i = 0; i < m->nSets; i++)
    if (m->sets[i].valley == valley && m->sets[i].region == region)
      return i;
  return -1;
}
/* emcScatterHandler.hpp:79-84 */
double orc_tau(const orc_model_t *m, int valley, int region) {
  int i = find_set(m, valley, region);
  return i >= 0 ? m->sets[i].tau : 2e-15;
}
double orc_dE(const orc_model_t *m) { return m->dE; }

/* emcScatterHandler.hpp:237-244 (x86-64 g++ behaviour of the size_t cast,
 * SURVEY App. B.4): -1 -> 0, above range -> n-1 */
int orc_energy_level(const orc_model_t *m, double energy) {
  double f = floor(energy / m->dE) - 1;
  int64_t lvl = (int64_t)f;
  if (lvl == -1)
    return 0;
  if (lvl < 0 || lvl > m->nLevels - 1)
    return m->nLevels - 1;
  return (int)lvl;
}
/* emcScatterHandler.hpp:148-170; returns table index or -1 (self-scatter) */
int orc_select(const orc_model_t *m, int si, double energy, double r) {
  const tableset_t *s = &m->sets[si];
  const int L = m->nLevels;
  int lvl = orc_energy_level(m, energy);
  if (r > s->cum[(s->nMech - 1) * L + lvl])
    return -1;
  double lower = 0;
  for (int t = 0; t < s->nMech; t++) {
    double upper = s->cum[t * L + lvl];
    if (r >= lower && r < upper)
      return t;
    lower = upper;
  }
  return -1;
}

/* ------------------------------------------------------- final states */
/* test aid: marks the particles whose q-resolved |q| sample came back ON a kinematic limit (forward / backward scattering).
 * There cos(theta) = (kI^2 + kF^2 - q^2) / (2 kI kF) is 1 - O(eps), so sin(theta) = O(sqrt(eps)) is made of the rounding of
 * kI and kF: the reference's own result moves by ~1e-8 when its input moves by one ulp (tests/test_oracle_mhp.py shows it) */
static unsigned char *g_limitFlags = 0;
void orc_set_limit_flags(unsigned char *flags) { g_limitFlags = flags; }
static void scatter_with(const orc_model_t *m, const orc_mech_t *d, orc_ensemble_t *e,
                         int64_t p, rng_t *rng) {
  double k[3] = {e->kx[p], e->ky[p], e->kz[p]}, out[3];
  switch (d->sampler) {
  case ORC_SAMPLER_ISOTROPIC_ELASTIC: {
    /* emcAcousticScatterMechanism.hpp:70-72; g++ evaluates the two dist(rng)
     * arguments right-to-left: first draw -> rand2 (cos theta) */
    double nrm = sqrt(sq3(k));
    double r2 = rng_u01(rng);
    double r1 = rng_u01(rng);
    orc_random_direction(nrm, r1, r2, out);
    break;
  }
  case ORC_SAMPLER_INTERVALLEY: {
    /* emcZeroOrder...:119-129 / :262-272, emcFirstOrder...:121-131 / :269-279 */
    int subOld = e->sub[p];
    e->valley[p] = d->finalValley;
    uint64_t raw = rng_raw(rng);
    e->sub[p] = d->finalSub[subOld][raw % (uint64_t)d->nFinal];
    e->energy[p] += d->p[0];
    double knew = orc_norm_wave_vec(&m->valleys[d->finalValley], e->energy[p]);
    double r2 = rng_u01(rng);
    double r1 = rng_u01(rng);
    orc_random_direction(knew, r1, r2, out);
    break;
  }
  case ORC_SAMPLER_SL_ELASTIC:
  case ORC_SAMPLER_SL_INTERVALLEY: {
    /* emcAcousticSingleLayerScatterMechanism.hpp:63-81; emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp:116-147,
     * :293-324 */
    if (d->sampler == ORC_SAMPLER_SL_INTERVALLEY) {
      int subOld = e->sub[p];
      e->valley[p] = d->finalValley;
      if (d->nFinal > 0)
        e->sub[p] = d->finalSub[subOld][(int)floor(rng_u01(rng) * d->nFinal)];
      e->energy[p] += d->p[0];
    }
    const orc_valley_t *v = &m->valleys[e->valley[p]];
    double angle = 2 * C_PI * rng_u01(rng);
    if (d->sampler == ORC_SAMPLER_SL_INTERVALLEY && d->p[1] != 0) {
      /* emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp:121-125, :270-274: no Herring-Vogt weighting */
      double normK = orc_norm_wave_vec(v, e->energy[p]);
      out[0] = normK * cos(angle);
      out[1] = normK * sin(angle);
      out[2] = k[2];
      break;
    }
    /* random direction weighted by the Herring-Vogt factors */
    out[0] = cos(angle) / v->vogt[0];
    out[1] = sin(angle) / v->vogt[1];
    out[2] = 0;
    double factor = 1. / (sqrt(out[0] * out[0] + out[1] * out[1]));
    out[0] *= factor;
    out[1] *= factor;
    double normK = orc_norm_wave_vec(v, e->energy[p]);
    out[0] *= normK;
    out[1] *= normK;
    break;
  }
  case ORC_SAMPLER_SL_FROEHLICH:
  case ORC_SAMPLER_SL_PIEZO: {
    /* emcFroehlichInteractionSingleLayer.hpp:45-80 + :149-168 / :273-292; emcPiezoelectricSingleLayerScatterMechanism.hpp
     * :110-139: deflection angle by inversion of a 128-point cumulative sum of the angular weight, random sign */
    const orc_valley_t *v = &m->valleys[e->valley[p]];
    const int fro = d->sampler == ORC_SAMPLER_SL_FROEHLICH;
    const double width = d->p[1], qs = d->p[2];
    double kI, kF;
    if (fro) {
      double initEnergy = e->energy[p];
      e->energy[p] = initEnergy + d->p[0];
      kI = orc_norm_wave_vec(v, initEnergy);
      kF = orc_norm_wave_vec(v, e->energy[p]);
    } else {
      kI = kF = orc_norm_wave_vec(v, e->energy[p]);
    }
    double cdf[129];
    const double dpsi = C_PI / 128;
    cdf[0] = 0;
    for (int i = 1; i <= 128; ++i)
      cdf[i] = cdf[i - 1] + (fro ? sl_froehlich_weight((i - 0.5) * dpsi, kI, kF, width, qs)
                                 : sl_piezo_weight((i - 0.5) * dpsi, kI, width, qs));
    const double total = cdf[128];
    double angle;
    if (fro) {
      double phi = atan2(k[1], k[0]);
      double psi;
      if (!(total > 0)) {
        psi = 2 * C_PI * rng_u01(rng);
      } else {
        double target = rng_u01(rng) * total;
        int lo = 1;
        while (lo < 128 && cdf[lo] < target)
          ++lo;
        double frac = (target - cdf[lo - 1]) / (cdf[lo] - cdf[lo - 1]);
        double psiMag = ((double)lo - 1 + frac) * dpsi;
        psi = (rng_u01(rng) < 0.5) ? psiMag : (2 * C_PI - psiMag);
      }
      angle = phi + psi;
    } else {
      double phi = atan2(k[1], k[0]);
      double theta;
      if (!(total > 0)) {
        theta = C_PI * rng_u01(rng);
      } else {
        double target = rng_u01(rng) * total;
        int lo = 1;
        while (lo < 128 && cdf[lo] < target)
          ++lo;
        theta = ((double)lo - 1 + (target - cdf[lo - 1]) / (cdf[lo] - cdf[lo - 1])) * dpsi;
      }
      if (rng_u01(rng) < 0.5)
        theta = -theta;
      angle = phi + theta;
    }
    out[0] = kF * cos(angle);
    out[1] = kF * sin(angle);
    out[2] = 0;
    break;
  }
  case ORC_SAMPLER_SL_CHARGED_IMPURITY:
  case ORC_SAMPLER_SL_SURFACE_ROUGHNESS:
  case ORC_SAMPLER_SL_REMOTE_SO:
  case ORC_SAMPLER_SL_SCREENED_OPTICAL: {
    /* emc2DChargedImpurityScatterMechanism.hpp:107-139, emcSurfaceRoughnessScatterMechanism.hpp:94-126,
     * emcRemoteSurfaceOpticalPhononMechanism.hpp:112-149, emcScreenedIntravalleyOpticalMechanism.hpp:104-141 */
    const orc_valley_t *v = &m->valleys[e->valley[p]];
    const int inelastic = d->sampler == ORC_SAMPLER_SL_REMOTE_SO || d->sampler == ORC_SAMPLER_SL_SCREENED_OPTICAL;
    const int n = sl_angular_steps(d->sampler);
    double kI = orc_norm_wave_vec(v, e->energy[p]), kF = kI;
    if (inelastic) {
      e->energy[p] = e->energy[p] + d->p[0];
      kF = orc_norm_wave_vec(v, e->energy[p]);
    }
    double cdf[513];
    const double dtheta = C_PI / n;
    cdf[0] = 0;
    for (int i = 1; i <= n; ++i)
      cdf[i] = cdf[i - 1] + sl_angular_weight(d->sampler, (i - 0.5) * dtheta, kI, kF, d->p);
    const double total = cdf[n];
    double phi = atan2(k[1], k[0]);
    double theta;
    if (!(total > 0)) {
      theta = C_PI * rng_u01(rng);
    } else {
      double target = rng_u01(rng) * total;
      int lo = 1;
      while (lo < n && cdf[lo] < target)
        ++lo;
      theta = ((double)lo - 1 + (target - cdf[lo - 1]) / (cdf[lo] - cdf[lo - 1])) * dtheta;
    }
    if (rng_u01(rng) < 0.5)
      theta = -theta;
    out[0] = kF * cos(phi + theta);
    out[1] = kF * sin(phi + theta);
    out[2] = 0;
    break;
  }
  case ORC_SAMPLER_COULOMB: {
    /* emcCoulombScatterMechanism.hpp:48-59 */
    const orc_valley_t *v = &m->valleys[e->valley[p]];
    double gamma = orc_gamma(v, e->energy[p]);
    double rnd = rng_u01(rng);
    double debyeEnergy = d->p[0];
    double cosTheta = 1.0 - rnd * 2.0 / ((1 - rnd) * gamma / debyeEnergy + 1.0);
    double r = rng_u01(rng);
    orc_random_direction_wrt_k(k, cosTheta, r, out);
    break;
  }
  case ORC_SAMPLER_FROEHLICH: {
    /* emcFroehlichInteraction.hpp:109-128 / :183-201 (+ the bath record of the hot-phonon classes) */
    const orc_valley_t *v = &m->valleys[e->valley[p]];
    double initEnergy = e->energy[p];
    e->energy[p] += d->p[0];
    double finalEnergy = e->energy[p];
    double f = 2 * sqrt(initEnergy * finalEnergy) / (sqrt(initEnergy) - sqrt(finalEnergy)) /
               (sqrt(initEnergy) - sqrt(finalEnergy));
    double cosTheta = (1 + f - pow(1 + 2 * f, rng_u01(rng))) / f;
    cosTheta = fmax(-1., fmin(1., cosTheta));
    double kNew = orc_norm_wave_vec(v, e->energy[p]);
    orc_random_direction_wrt_k(k, cosTheta, rng_u01(rng), out);
    double kCurr = sqrt(sq3(out));
    if (kCurr > 0) {
      double s = kNew / kCurr;
      for (int i = 0; i < 3; i++)
        out[i] = out[i] * s;
    }
    if (d->p[2] >= 0) {
      double q[3] = {out[0] - k[0], out[1] - k[1], out[2] - k[2]};
      orc_bath_record(m->baths[(int)d->p[2]], sqrt(sq3(q)), d->p[0] < 0);
    }
    break;
  }
  case ORC_SAMPLER_SCREENED_FROEHLICH: {
    /* emcScreenedFroehlichInteraction.hpp:140-153, :199-213, :272-294, :354-374; helpers :66-97 */
    const orc_valley_t *v = &m->valleys[e->valley[p]];
    const int emission = d->p[0] < 0;
    double kI = orc_norm_wave_vec(v, e->energy[p]);
    e->energy[p] += d->p[0];
    double kF = orc_norm_wave_vec(v, e->energy[p]);
    double r = rng_u01(rng);
    double cosTheta;
    const double B = 2 * kI * kF;
    if (B <= 0) {
      cosTheta = 1 - 2 * r;
    } else if (d->p[3] != 0) {
      double q = orc_bath_sample_q(m->baths[(int)d->p[2]], fabs(kI - kF), kI + kF, emission, r);
      if (g_limitFlags && (q == fabs(kI - kF) || q == kI + kF))
        g_limitFlags[p] = 1;
      cosTheta = fmax(-1., fmin(1., (kI * kI + kF * kF - q * q) / B));
    } else {
      const double Ap = kI * kI + kF * kF + d->p[1];
      const double num = Ap - B, den = Ap + B;
      if (num <= 0 || den <= 0)
        cosTheta = 1 - 2 * r;
      else
        cosTheta = fmax(-1., fmin(1., (Ap - den * pow(num / den, r)) / B));
    }
    orc_random_direction_wrt_k(k, cosTheta, rng_u01(rng), out);
    double kCurr = sqrt(sq3(out));
    if (kCurr > 0) {
      double s = kF / kCurr;
      for (int i = 0; i < 3; i++)
        out[i] = out[i] * s;
    }
    if (d->p[2] >= 0) {
      double q[3] = {out[0] - k[0], out[1] - k[1], out[2] - k[2]};
      orc_bath_record(m->baths[(int)d->p[2]], sqrt(sq3(q)), emission);
    }
    break;
  }
  default:
    return;
  }
  e->kx[p] = out[0];
  e->ky[p] = out[1];
  e->kz[p] = out[2];
}

/* ------------------------------------------------------- phonon bath (emcPhononBath.hpp) */
static double bath_planck(double energyEV, double tempK) {
  double x = C_Q * energyEV / (C_KB * tempK);
  return 1. / (exp(x) - 1.);
}
static double bath_planck_temp(double energyEV, double N) {
  if (N <= 0)
    return 0;
  return C_Q * energyEV / (C_KB * log(1. + 1. / N));
}
static void bath_rebuild_sums(orc_bath_t *b) { /* :122-135 */
  b->cumW[0] = b->cumWN[0] = 0;
  for (int i = 0; i < b->nBins; i++) {
    const double q = ((double)i + 0.5) * b->dq;
    const double denom = q * q + b->qs2;
    const double w = (denom > 0) ? q / denom : 0;
    b->cumW[i + 1] = b->cumW[i] + w;
    b->cumWN[i + 1] = b->cumWN[i] + w * b->Nq[i];
  }
}
orc_bath_t *orc_bath_create(int nBins, double dq, double tauLO, double phononEnergy, double latticeTemp, double Vsim,
                            int enableAcoustic, double acPhononEnergy, double tauAcoustic, double wRidley,
                            double toPhononEnergy, double tauTO) {
  orc_bath_t *b = (orc_bath_t *)calloc(1, sizeof *b);
  b->nBins = nBins;
  b->dq = dq;
  b->tauLO = tauLO;
  b->Vsim = Vsim;
  b->loE = phononEnergy;
  b->latT = latticeTemp;
  b->Nq = (double *)calloc((size_t)nBins, sizeof(double));
  b->nEm = (double *)calloc((size_t)nBins, sizeof(double));
  b->nAbs = (double *)calloc((size_t)nBins, sizeof(double));
  b->cumW = (double *)calloc((size_t)nBins + 1, sizeof(double));
  b->cumWN = (double *)calloc((size_t)nBins + 1, sizeof(double));
  b->N0 = bath_planck(phononEnergy, latticeTemp);
  for (int i = 0; i < nBins; i++)
    b->Nq[i] = b->N0;
  if (enableAcoustic) {
    b->hasAc = 1;
    b->acE = acPhononEnergy;
    b->tauAc = tauAcoustic;
    b->NacEq = bath_planck(acPhononEnergy, latticeTemp);
    b->Nac = b->NacEq;
  }
  if (enableAcoustic && wRidley > 0 && toPhononEnergy > 0) {
    b->hasRidley = 1;
    b->wRidley = wRidley > 1 ? 1 : wRidley;
    b->toE = toPhononEnergy;
    b->tauTO = tauTO;
    b->NtoEq = bath_planck(toPhononEnergy, latticeTemp);
    b->Nto = b->NtoEq;
  }
  bath_rebuild_sums(b);
  return b;
}
void orc_bath_destroy(orc_bath_t *b) {
  if (!b)
    return;
  free(b->Nq); free(b->nEm); free(b->nAbs); free(b->cumW); free(b->cumWN);
  free(b);
}
void orc_bath_set_qs2(orc_bath_t *b, double qs2) {
  if (qs2 == b->qs2)
    return;
  b->qs2 = qs2;
  bath_rebuild_sums(b);
}
static int bath_bin(const orc_bath_t *b, double q) { /* binOf :228-231 */
  uint64_t idx = (uint64_t)(q / b->dq);
  return idx >= (uint64_t)b->nBins ? b->nBins - 1 : (int)idx;
}
void orc_bath_record(orc_bath_t *b, double q, int emission) {
  if (emission)
    b->nEm[bath_bin(b, q)] += 1.;
  else
    b->nAbs[bath_bin(b, q)] += 1.;
}
void orc_bath_add_counts(orc_bath_t *b, const int64_t *emission, const int64_t *absorption) {
  for (int i = 0; i < b->nBins; i++) {
    b->nEm[i] += (double)emission[i];
    b->nAbs[i] += (double)absorption[i];
  }
}
double orc_bath_mean_nq(const orc_bath_t *b) {
  double sumW = 0, sumWN = 0;
  for (int i = 0; i < b->nBins; i++) {
    double q = ((double)i + 0.5) * b->dq;
    double w = q * q;
    sumW += w;
    sumWN += w * b->Nq[i];
  }
  return sumW > 0 ? sumWN / sumW : b->N0;
}
void orc_bath_update(orc_bath_t *b, double dt) {
  double N_target = b->N0;
  if (b->hasAc) {
    double N_klemens = b->N0;
    if (b->Nac > b->NacEq)
      N_klemens = bath_planck(b->loE, bath_planck_temp(b->acE, b->Nac));
    if (b->hasRidley) {
      double N_ridley = b->N0;
      if (b->Nto > b->NtoEq)
        N_ridley = bath_planck(b->loE, bath_planck_temp(b->toE, b->Nto));
      N_target = (1. - b->wRidley) * N_klemens + b->wRidley * N_ridley;
    } else {
      N_target = N_klemens;
    }
  }
  double sumW = 0, sumWdN = 0;
  for (int i = 0; i < b->nBins; i++) {
    double q = ((double)i + 0.5) * b->dq;
    double Dph = q * q * b->dq * b->Vsim / (2. * C_PI * C_PI);
    double g = (Dph > 0) ? (b->nEm[i] - b->nAbs[i]) / (Dph * dt) : 0;
    const double tau_i = b->tauLO; /* no tauLOProfile in the configs */
    if (b->hasAc) {
      double w = q * q;
      sumW += w;
      sumWdN += w * (b->Nq[i] - N_target);
    }
    b->Nq[i] += g * dt - (dt / tau_i) * (b->Nq[i] - N_target);
    if (b->Nq[i] < 0)
      b->Nq[i] = 0;
    b->nEm[i] = 0;
    b->nAbs[i] = 0;
  }
  if (b->hasAc) {
    double fKlemens = b->hasRidley ? (1. - b->wRidley) : 1.;
    double dN_LO_mean = (sumW > 0) ? sumWdN / sumW : 0;
    b->Nac += dt * fKlemens * dN_LO_mean / b->tauLO - dt * (b->Nac - b->NacEq) / b->tauAc;
    if (b->Nac < b->NacEq)
      b->Nac = b->NacEq;
    if (b->hasRidley) {
      b->Nto += dt * b->wRidley * dN_LO_mean / b->tauLO - dt * (b->Nto - b->NtoEq) / b->tauTO;
      if (b->Nto < b->NtoEq)
        b->Nto = b->NtoEq;
    }
  }
  bath_rebuild_sums(b);
}
double orc_bath_nq_window(const orc_bath_t *b, double qMin, double qMax) {
  if (b->nBins < 2 || qMax <= qMin)
    return orc_bath_mean_nq(b);
  long lo = (long)floor(qMin / b->dq);
  long hi = (long)ceil(qMax / b->dq);
  if (lo < 0)
    lo = 0;
  if (hi > (long)b->nBins)
    hi = (long)b->nBins;
  if (hi - lo < 1)
    return orc_bath_mean_nq(b);
  const double wSum = b->cumW[hi] - b->cumW[lo];
  if (wSum <= 0)
    return orc_bath_mean_nq(b);
  return (b->cumWN[hi] - b->cumWN[lo]) / wSum;
}
double orc_bath_sample_q(const orc_bath_t *b, double qMin, double qMax, int emission, double r) {
  if (b->nBins < 2 || qMax <= qMin)
    return qMin;
  long lo = (long)floor(qMin / b->dq);
  long hi = (long)ceil(qMax / b->dq);
  if (lo < 0)
    lo = 0;
  if (hi > (long)b->nBins)
    hi = (long)b->nBins;
  if (hi - lo < 1)
    return qMin;
#define BATH_S(i) (emission ? (b->cumWN[i] + b->cumW[i]) : b->cumWN[i])
  const double sLo = BATH_S(lo), sHi = BATH_S(hi);
  const double span = sHi - sLo;
  if (!(span > 0))
    return 0.5 * (qMin + qMax);
  const double target = sLo + r * span;
  long a = lo, bb = hi;
  while (bb - a > 1) {
    const long mid = (a + bb) / 2;
    if (BATH_S(mid) <= target)
      a = mid;
    else
      bb = mid;
  }
  const double sA = BATH_S(a), sB = BATH_S(a + 1);
#undef BATH_S
  const double frac = (sB > sA) ? (target - sA) / (sB - sA) : 0.5;
  const double q = ((double)a + frac) * b->dq;
  return fmax(qMin, fmin(qMax, q));
}
double orc_bath_acoustic_temp(const orc_bath_t *b) {
  if (!b->hasAc || b->Nac <= b->NacEq)
    return b->latT;
  return bath_planck_temp(b->acE, b->Nac);
}
double orc_bath_n0(const orc_bath_t *b) { return b->N0; }
int orc_bath_copy(const orc_bath_t *b, int which, double *out) {
  const double *src[5] = {b->Nq, b->nEm, b->nAbs, b->cumW, b->cumWN};
  if (which < 0 || which > 4)
    return -1;
  memcpy(out, src[which], sizeof(double) * (size_t)(b->nBins + (which >= 3 ? 1 : 0)));
  return 0;
}

/* emcGrainScatterMechanism::scatterParticle (:40-77): reflected into the opposite or transmitted into the same hemisphere
 * about the current k; then the new clock (emcParticleType.hpp:191-193) */
static void grain_event(const orc_model_t *m, orc_ensemble_t *e, int64_t p, rng_t *rng) {
  if (m->hasGrain) {
    double k[3] = {e->kx[p], e->ky[p], e->kz[p]}, out[3];
    double rand;
    if (rng_u01(rng) > m->grainProb) {
      rand = rng_u01(rng);
      if (rand < 0.5)
        rand += 0.5;
    } else {
      rand = rng_u01(rng);
      if (rand > 0.5)
        rand -= 0.5;
    }
    orc_random_direction_wrt_k(k, 1 - 2 * rand, rng_u01(rng), out);
    e->kx[p] = out[0];
    e->ky[p] = out[1];
    e->kz[p] = out[2];
  }
  e->grainTau[p] = -log(rng_ulog(rng)) * m->grainTau;
}

static void drift_wrap(const orc_model_t *m, orc_ensemble_t *e, int64_t p, double dt,
                       const double force[3], const double box[3]) {
  /* basicBulkParticleHandler.hpp:600-613 */
  double k[3] = {e->kx[p], e->ky[p], e->kz[p]};
  double pos[3] = {e->x[p], e->y[p], e->z[p]};
  orc_drift(&m->valleys[e->valley[p]], dt, k, &e->energy[p], e->sub[p], pos, 3, force);
  for (int d = 0; d < 3; d++) {
    if (pos[d] < 0)
      pos[d] = pos[d] + box[d];
    else if (pos[d] > box[d])
      pos[d] = pos[d] - box[d];
  }
  e->kx[p] = k[0]; e->ky[p] = k[1]; e->kz[p] = k[2];
  e->x[p] = pos[0]; e->y[p] = pos[1]; e->z[p] = pos[2];
}

void orc_bulk_observables(const orc_model_t *m, const orc_ensemble_t *e,
                          const double fieldDir[3], double *obs) {
  /* :289-347; fieldDir already normalised (ctor :101) */
  for (int v = 0; v < m->nValleys * 3; v++)
    obs[v] = 0.;
  for (int64_t p = 0; p < e->n; p++) {
    int v = e->valley[p];
    double k[3] = {e->kx[p], e->ky[p], e->kz[p]}, vel[3];
    orc_velocity(&m->valleys[v], k, e->energy[p], e->sub[p], vel);
    obs[v * 3 + 0] += e->energy[p];
    obs[v * 3 + 1] += vel[0] * fieldDir[0] + vel[1] * fieldDir[1] + vel[2] * fieldDir[2];
    obs[v * 3 + 2] += 1.;
  }
}

int orc_bulk_steps(const orc_model_t *m, orc_ensemble_t *e, const double box[3],
                   const double fieldDirIn[3], double fieldStrength, double charge,
                   double dt, int nSteps,
                   int64_t firstStep, const orc_rng_cfg_t *cfg, double *obs,
                   int32_t *recPid, int64_t recCap, int64_t *recCount, int64_t *events,
                   int64_t evCap, int64_t *evCount) {
  rng_t rng;
  memset(&rng, 0, sizeof rng);
  rng.cfg = cfg;
  rng.recPid = recPid;
  rng.recCap = recCap;
  int64_t nEv = 0;
  /* ctor :100-102: normalize(dir); appliedField = scale(dir, strength) */
  double dir[3] = {fieldDirIn[0], fieldDirIn[1], fieldDirIn[2]};
  {
    double nrm = sqrt(sq3(dir));
    if (nrm != 0.)
      for (int d = 0; d < 3; d++)
        dir[d] /= nrm;
  }
  double field[3] = {dir[0] * fieldStrength, dir[1] * fieldStrength, dir[2] * fieldStrength};
  /* :186 force = scale(appliedField, charge) */
  double force[3] = {field[0] * charge, field[1] * charge, field[2] * charge};
  orc_mech_t desc;
  for (int s = 0; s < nSteps; s++) {
    for (int64_t p = 0; p < e->n; p++) {
      rng.particle = p;
      rng.step = (uint64_t)(firstStep + s);
      rng.k = 0;
      /* :195-213 */
      double tau = e->tau[p];
      drift_wrap(m, e, p, tau < dt ? tau : dt, force, box);
      double tRem = dt - tau;
      while (tRem > 0) {
        int si = find_set(m, e->valley[p], e->region[p]);
        int t = -2;
        if (si >= 0 && m->sets[si].nMech > 0) {
          double r = rng_u01(&rng);
          t = orc_select(m, si, e->energy[p], r);
          if (t >= 0) {
            fill_mech_desc(m, m->sets[si].mechIdx[t], &desc);
            scatter_with(m, &desc, e, p, &rng);
          }
          if (events && nEv < evCap) {
            events[nEv * 4 + 0] = firstStep + s;
            events[nEv * 4 + 1] = p;
            events[nEv * 4 + 2] = t;
            events[nEv * 4 + 3] = t >= 0 ? m->sets[si].mechIdx[t] : -1;
          }
          nEv++;
        }
        /* emcParticleType.hpp:187-189: uses the NEW valley, unchanged region */
        double newTau = -log(rng_ulog(&rng)) * orc_tau(m, e->valley[p], e->region[p]);
        tau += newTau;
        drift_wrap(m, e, p, tRem < newTau ? tRem : newTau, force, box);
        tRem -= newTau;
      }
      tau -= dt;
      e->tau[p] = tau;
      /* :216-220 grain clock (no grain mechanism: only the clock runs) */
      if (e->grainTau) {
        e->grainTau[p] -= dt;
        if (e->grainTau[p] <= 0)
          grain_event(m, e, p, &rng);
      }
    }
    if (obs)
      orc_bulk_observables(m, e, dir, obs + (size_t)s * m->nValleys * 3);
  }
  if (recCount)
    *recCount = rng.recCount;
  if (evCount)
    *evCount = nEv;
  return 0;
}

/* ---------------------------------------------------------- initialisation */
int64_t orc_generate_initial(const orc_model_t *m, const double box[3], const int32_t cells[3],
                             double doping, uint64_t *mtState, orc_ensemble_t *out,
                             int64_t capacity, int64_t *drawsConsumed) {
  orc_rng_cfg_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.mode = ORC_RNG_MT_GLOBAL;
  cfg.mtState = mtState;
  rng_t rng;
  memset(&rng, 0, sizeof rng);
  rng.cfg = &cfg;
  double h[3];
  int64_t ext[3];
  for (int d = 0; d < 3; d++) {
    h[d] = box[d] / cells[d];
    ext[d] = (int64_t)round(box[d] / h[d]) + 1; /* emcUtil.hpp:109-119 */
  }
  /* emcDevice.hpp:333-343 */
  double cellVolume = 1.;
  for (int d = 0; d < 3; d++)
    cellVolume = cellVolume * h[d];
  const double Vt = C_KB / C_Q * m->temperature;
  int64_t n = 0;
  for (int64_t cz = 0; cz < ext[2]; cz++)
    for (int64_t cy = 0; cy < ext[1]; cy++)
      for (int64_t cx = 0; cx < ext[0]; cx++) {
        int64_t c[3] = {cx, cy, cz};
        /* emcElectron.hpp:48-61 */
        double dens = doping;
        for (int d = 0; d < 3; d++)
          if (c[d] == 0 || c[d] == ext[d] - 1)
            dens *= 0.5;
        double nr = dens * cellVolume;
        if (m->electron2D > 0)
          nr = m->electron2D; /* electron2D.hpp:38-41 */
        /* basicBulkParticleHandler.hpp:150-156 */
        for (;;) {
          int create = 0;
          if (nr >= 1) {
            create = 1;
          } else {
            if (rng_u01(&rng) < nr)
              create = 2;
            else
              break;
          }
          if (n < capacity) {
            /* emcParticleInitialization.hpp:14-29 */
            double pos[3];
            for (int d = 0; d < 3; d++) {
              if (c[d] == ext[d] - 1)
                pos[d] = ((double)c[d] - rng_u01(&rng) * 0.5) * h[d];
              else if (c[d] == 0)
                pos[d] = rng_u01(&rng) * 0.5 * h[d];
              else
                pos[d] = ((double)c[d] + rng_u01(&rng) - 0.5) * h[d];
            }
            int valley, sub;
            double energy, k[3];
            if (m->electron2D > 0) {
              /* electron2D.hpp:43-67: first valley, random sub-valley, 2-D thermal energy, in-plane direction */
              valley = 0;
              const orc_valley_t *v = &m->valleys[0];
              sub = (int)floor(v->deg * rng_u01(&rng));
              energy = -Vt * log(rng_ulog(&rng));
              double normK = orc_norm_wave_vec(v, energy);
              double angle = 2 * C_PI * rng_u01(&rng);
              k[0] = normK * cos(angle);
              k[1] = normK * sin(angle);
              k[2] = 0;
            } else {
              /* emcElectron.hpp:75-90 */
              valley = (int)floor(m->nValleys * rng_ulog(&rng));
              const orc_valley_t *v = &m->valleys[valley];
              sub = (int)floor(v->deg * rng_ulog(&rng));
              /* emcParticleInitialization.hpp:36-51; :60-74 for a fixed start energy (no draw) */
              energy = m->initEnergy > 0. ? m->initEnergy : -1.5 * Vt * log(rng_ulog(&rng));
              double r2 = rng_u01(&rng); /* right-to-left: first draw is rand2 */
              double r1 = rng_u01(&rng);
              orc_random_direction(orc_norm_wave_vec(v, energy), r1, r2, k);
              for (int d = 0; d < 3; d++)
                if ((c[d] == 0 && k[d] < 0) || (c[d] == ext[d] - 1 && k[d] > 0))
                  k[d] *= -1;
            }
            double tau = -log(rng_ulog(&rng)) * orc_tau(m, valley, 0);
            double gtau = -log(rng_ulog(&rng)) * m->grainTau;
            out->kx[n] = k[0]; out->ky[n] = k[1]; out->kz[n] = k[2];
            out->energy[n] = energy;
            out->tau[n] = tau;
            if (out->grainTau)
              out->grainTau[n] = gtau;
            out->x[n] = pos[0]; out->y[n] = pos[1]; out->z[n] = pos[2];
            out->valley[n] = valley;
            out->sub[n] = sub;
            out->region[n] = 0;
          } else {
            /* still consume the draws so the count stays meaningful */
            for (int i = 0; i < 10; i++)
              rng_raw(&rng);
          }
          n++;
          if (create == 2)
            break;
          nr--;
        }
      }
  out->n = n <= capacity ? n : capacity;
  if (drawsConsumed)
    *drawsConsumed = rng.recCount;
  return n <= capacity ? n : -n;
}

/* ======================================================================================
 * Device-run path.  All file:line citations relative to the reference tree.
 * ====================================================================================== */
int64_t orc_dev_cells(const orc_device_t *d) {
  int64_t n = 1;
  for (int i = 0; i < d->dim; i++)
    n *= d->extent[i];
  return n;
}
static void dev_coord(const orc_device_t *d, int64_t cell, int64_t c[3]) {
  c[0] = cell % d->extent[0];
  c[1] = (cell / d->extent[0]) % d->extent[1];
  c[2] = d->dim > 2 ? cell / ((int64_t)d->extent[0] * d->extent[1]) : 0;
}
static int64_t dev_cell(const orc_device_t *d, const int64_t c[3]) {
  int64_t idx = c[0] + (int64_t)d->extent[0] * c[1];
  if (d->dim > 2)
    idx += (int64_t)d->extent[0] * d->extent[1] * c[2];
  return idx;
}
/* first face (XMIN, XMAX, YMIN, YMAX, ZMIN, ZMAX) the cell lies on, or -1 (emcSurface.hpp:340-349) */
static int dev_first_face(const orc_device_t *d, int64_t cell) {
  const int8_t *fc = d->faceContact + cell * 2 * d->dim;
  for (int f = 0; f < 2 * d->dim; f++)
    if (fc[f] != -2)
      return f;
  return -1;
}
int orc_dev_contact_idx(const orc_device_t *d, int64_t cell) {
  int f = dev_first_face(d, cell);
  return f < 0 ? -1 : d->faceContact[cell * 2 * d->dim + f];
}
int orc_dev_is_ohmic(const orc_device_t *d, int64_t cell) {
  int c = orc_dev_contact_idx(d, cell);
  return c >= 0 && d->contactType[c] == ORC_CONTACT_OHMIC;
}
int orc_dev_is_reservoir(const orc_device_t *d, int64_t cell) {
  int c = orc_dev_contact_idx(d, cell);
  return c >= 0 && (d->contactType[c] == ORC_CONTACT_OHMIC || d->contactType[c] == ORC_CONTACT_SCHOTTKY);
}
/* emcDevice.hpp:274-281 posToCoord: round(pos / spacing) per dimension */
static int64_t dev_pos_to_cell(const orc_device_t *d, const double pos[3]) {
  int64_t c[3] = {0, 0, 0};
  for (int i = 0; i < d->dim; i++)
    c[i] = (int64_t)round(pos[i] / d->spacing[i]);
  return dev_cell(d, c);
}

void orc_initial_potential(const orc_device_t *d, double *pot) {
  const int64_t n = orc_dev_cells(d);
  for (int64_t i = 0; i < n; i++)
    pot[i] = asinh(0.5 * (d->doping[i] / d->ni));
}

int orc_sor(const orc_device_t *d, double *pot, const double *conc, double accuracyVolt, double omega, int resetBC,
            int maxSweeps) {
  const int dim = d->dim;
  const int64_t n = orc_dev_cells(d);
  /* emcSORSolver.hpp:27, :330-368 */
  const double accuracy = accuracyVolt / d->thermalVoltage;
  double h[3], hF[3];
  for (int i = 0; i < dim; i++)
    h[i] = d->spacing[i] / d->debyeLength;
  if (dim == 2) {
    hF[0] = h[1] / h[0];
    hF[1] = h[0] / h[1];
  } else {
    hF[0] = h[1] * h[2] / h[0];
    hF[1] = h[0] * h[2] / h[1];
    hF[2] = h[0] * h[1] / h[2];
  }
  double hFSum = 0., hProd = 1.;
  for (int i = 0; i < dim; i++) {
    hFSum += hF[i];
    hProd = hProd * h[i];
  }
  int64_t stride[3] = {1, d->extent[0], (int64_t)d->extent[0] * d->extent[1]};
  if (resetBC) {
    /* :57-73 / :139-155: faces in the order XMIN, XMAX, YMIN, ... ; ohmic cells become Dirichlet values */
    for (int f = 0; f < 2 * dim; f++)
      for (int64_t cell = 0; cell < n; cell++) {
        int c = d->faceContact[cell * 2 * dim + f];
        if (c >= 0 && d->contactType[c] == ORC_CONTACT_OHMIC) {
          double builtIn = asinh(0.5 * (d->doping[cell] / d->ni));
          pot[cell] = conc ? d->contactVoltage[c] / d->thermalVoltage + builtIn : builtIn;
        }
      }
  }
  int sweeps = 0;
  double error;
  do {
    error = 0;
    for (int64_t cell = 0; cell < n; cell++) {
      if (orc_dev_is_reservoir(d, cell))
        continue;
      int64_t c[3];
      dev_coord(d, cell, c);
      const double cur = pot[cell];
      double p, nn;
      if (conc) {
        p = exp(-cur);
        nn = conc[cell];
      } else {
        nn = exp(cur);
        p = 1. / nn;
      }
      const double dop = d->doping[cell] / d->ni;
      double num = hProd * (p - nn + dop + cur * (p + nn));
      double den = 2 * hFSum + hProd * (nn + p);
      for (int i = 0; i < dim; i++) {
        for (int side = 0; side < 2; side++) {
          const int atFace = side == 0 ? c[i] == 0 : c[i] == d->extent[i] - 1;
          if (!atFace) {
            num += pot[cell + (side == 0 ? -stride[i] : stride[i])] * hF[i];
          } else {
            num += pot[cell + (side == 0 ? stride[i] : -stride[i])] * hF[i]; /* mirror */
            /* checkForGateContact :399-412 */
            int ct = d->faceContact[cell * 2 * dim + 2 * i + side];
            if (ct >= 0 && d->contactType[ct] == ORC_CONTACT_GATE) {
              double gammaOx = d->gateEpsOx[ct] / d->epsR;
              double tOx = d->gateThickness[ct] / d->debyeLength;
              double gF = 2 * gammaOx / tOx;
              double Vg = d->gateBarrier[ct] / d->thermalVoltage;
              if (conc)
                Vg += d->contactVoltage[ct] / d->thermalVoltage;
              num += gF * Vg * hF[i] * h[i];
              den += gF * hF[i] * h[i];
            }
          }
        }
      }
      const double delta = omega * (num / den - cur);
      pot[cell] = cur + delta;
      if (fabs(delta) > error)
        error = fabs(delta);
    }
    sweeps++;
  } while (error > accuracy && (maxSweeps <= 0 || sweeps < maxSweeps));
  return sweeps;
}

void orc_efield(const orc_device_t *d, const double *pot, double *e) {
  const int dim = d->dim;
  const int64_t n = orc_dev_cells(d);
  int64_t stride[3] = {1, d->extent[0], (int64_t)d->extent[0] * d->extent[1]};
  if (d->pmScheme == ORC_PM_NEC_VWD) {
    /* NECSchemeVWD.hpp:82-99: forward difference everywhere but on the max face, where E = 0; no contact rule */
    for (int i = 0; i < dim; i++) {
      double *ed = e + (int64_t)i * n;
      for (int64_t cell = 0; cell < n; cell++) {
        int64_t c[3];
        dev_coord(d, cell, c);
        if (c[i] != d->extent[i] - 1)
          ed[cell] = ((pot[cell] - pot[cell + stride[i]]) * d->thermalVoltage) / d->spacing[i];
        else
          ed[cell] = 0;
      }
    }
    return;
  }
  const int midPts = d->pmScheme == ORC_PM_NEC; /* calcEFieldAtEdgeMidPts :35-53 instead of calcEFieldAtGridPts :13-30 */
  for (int i = 0; i < dim; i++) {
    double *ed = e + (int64_t)i * n;
    for (int64_t cell = 0; cell < n; cell++) {
      int64_t c[3];
      dev_coord(d, cell, c);
      if (c[i] != 0 && c[i] != d->extent[i] - 1) {
        if (midPts)
          ed[cell] = ((pot[cell] - pot[cell + stride[i]]) * d->thermalVoltage) / d->spacing[i];
        else
          ed[cell] = ((pot[cell - stride[i]] - pot[cell + stride[i]]) * d->thermalVoltage) / (2 * d->spacing[i]);
      }
    }
    /* setEFieldBoundaryValues :58-82: normal component 0 on artificial boundaries, copied from the
     * inner neighbour at contacts */
    for (int64_t cell = 0; cell < n; cell++) {
      int64_t c[3];
      dev_coord(d, cell, c);
      if (c[i] == 0) {
        int ct = d->faceContact[cell * 2 * dim + 2 * i];
        ed[cell] = ct == -1 ? 0. : ed[cell + stride[i]];
      } else if (c[i] == d->extent[i] - 1) {
        int ct = d->faceContact[cell * 2 * dim + 2 * i + 1];
        ed[cell] = ct == -1 ? 0. : ed[cell - stride[i]];
      }
    }
  }
}

/* lower-left grid point of the cell a position lies in: floor(pos / spacing) */
static void dev_floor_coord(const orc_device_t *d, const double pos[3], int64_t c[3], double w[3]) {
  c[0] = c[1] = c[2] = 0;
  w[0] = w[1] = w[2] = 0;
  for (int i = 0; i < d->dim; i++) {
    c[i] = (int64_t)floor(pos[i] / d->spacing[i]);
    w[i] = pos[i] / d->spacing[i] - (double)c[i];
  }
}

int orc_assign(const orc_device_t *d, int64_t n, const double *x, const double *y, const double *z, double nrCarriers,
               double *count) {
  const int64_t sx = 1, sy = d->extent[0], sz = (int64_t)d->extent[0] * d->extent[1];
  for (int64_t p = 0; p < n; p++) {
    const double pos[3] = {x[p], y[p], d->dim > 2 ? z[p] : 0.};
    if (d->pmScheme == ORC_PM_NGP) {
      count[dev_pos_to_cell(d, pos)] += nrCarriers;
      continue;
    }
    int64_t c[3];
    double w[3];
    dev_floor_coord(d, pos, c, w);
    const int64_t base = dev_cell(d, c);
    if (d->pmScheme == ORC_PM_CIC) {
      /* emcCICScheme.hpp:43-118: the weight w (distance from the LOWER point) goes to the lower point */
      const double wX = w[0], wY = w[1], wZ = w[2];
      if (d->dim == 2) {
        count[base] += wX * wY * nrCarriers;
        count[base + sx] += (1 - wX) * wY * nrCarriers;
        count[base + sy] += wX * (1 - wY) * nrCarriers;
        count[base + sx + sy] += (1 - wX) * (1 - wY) * nrCarriers;
      } else {
        count[base] += wX * wY * wZ * nrCarriers;
        count[base + sx] += (1 - wX) * wY * wZ * nrCarriers;
        count[base + sy] += wX * (1 - wY) * wZ * nrCarriers;
        count[base + sx + sy] += (1 - wX) * (1 - wY) * wZ * nrCarriers;
        count[base + sz] += wX * wY * (1 - wZ) * nrCarriers;
        count[base + sx + sz] += (1 - wX) * wY * (1 - wZ) * nrCarriers;
        count[base + sy + sz] += wX * (1 - wY) * (1 - wZ) * nrCarriers;
        count[base + sx + sy + sz] += (1 - wX) * (1 - wY) * (1 - wZ) * nrCarriers;
      }
    } else { /* NEC, NEC-VWD: equal shares (emcNECScheme.hpp:33-96, NECSchemeVWD.hpp:25-52) */
      const double weight = (d->dim == 2 ? 0.25 : 0.125) * nrCarriers;
      for (int dz = 0; dz < (d->dim > 2 ? 2 : 1); dz++)
        for (int dy = 0; dy < 2; dy++)
          for (int dx = 0; dx < 2; dx++)
            count[base + dx * sx + dy * sy + dz * sz] += weight;
    }
  }
  return 0;
}

int orc_ngp_assign(const orc_device_t *d, int64_t n, const double *x, const double *y, const double *z,
                   double nrCarriers, double *count) {
  for (int64_t p = 0; p < n; p++) {
    const double pos[3] = {x[p], y[p], d->dim > 2 ? z[p] : 0.};
    count[dev_pos_to_cell(d, pos)] += nrCarriers;
  }
  return 0;
}

void orc_concentration(const orc_device_t *d, const double *count, double *conc) {
  const int64_t n = orc_dev_cells(d);
  const double factor = 1. / d->cellVolume;
  for (int64_t cell = 0; cell < n; cell++) {
    int64_t c[3];
    dev_coord(d, cell, c);
    double v = (count[cell] / d->ni) * factor;
    for (int i = 0; i < d->dim; i++)
      if (c[i] == 0 || c[i] == d->extent[i] - 1)
        v *= 2;
    conc[cell] = v;
  }
}

void orc_expected_at_contact(const orc_device_t *d, double *expected) {
  const int64_t n = orc_dev_cells(d);
  for (int64_t cell = 0; cell < n; cell++) {
    expected[cell] = 0;
    if (!orc_dev_is_reservoir(d, cell))
      continue;
    int64_t c[3];
    dev_coord(d, cell, c);
    double v = d->cellVolume * d->doping[cell];
    for (int i = 0; i < d->dim; i++)
      if (c[i] == 0 || c[i] == d->extent[i] - 1)
        v *= 0.5;
    expected[cell] = d->electronKind == 1 ? round(v) : v; /* electronVWD.hpp:53-62 */
  }
}

double orc_initial_nr_particles(const orc_device_t *d, int64_t cell, const double *pot) {
  int64_t c[3];
  dev_coord(d, cell, c);
  /* emcElectron.hpp:48-61 / electronVWD.hpp:40-49 */
  double dens = pot ? exp(pot[cell]) * d->ni : d->doping[cell]; /* electronVWD: pot must be given */
  for (int i = 0; i < d->dim; i++)
    if (c[i] == 0 || c[i] == d->extent[i] - 1)
      dens *= 0.5;
  double nr = dens * d->cellVolume;
  return d->electronKind == 1 ? round(nr) : nr;
}

/* emcBasicParticleHandler.hpp:239-253 addParticle: position (emcParticleInitialization.hpp:14-29), then
 * emcElectron::generate{Initial,Injected}Particle (emcElectron.hpp:75-104) */
static void dev_create_particle(const orc_model_t *m, const orc_device_t *d, const int64_t c[3], rng_t *rng,
                                orc_ensemble_t *out, int64_t slot) {
  double pos[3] = {0, 0, 0};
  for (int i = 0; i < d->dim; i++) {
    if (c[i] == d->extent[i] - 1)
      pos[i] = ((double)c[i] - rng_u01(rng) * 0.5) * d->spacing[i];
    else if (c[i] == 0)
      pos[i] = rng_u01(rng) * 0.5 * d->spacing[i];
    else
      pos[i] = ((double)c[i] + rng_u01(rng) - 0.5) * d->spacing[i];
  }
  const int region = d->region[dev_cell(d, c)];
  /* emcElectron draws valley / sub-valley from U[1e-6, 1), electronVWD from U[0, 1) (electronVWD.hpp:25, :69-72) */
  const int vwd = d->electronKind == 1;
  const int valley = (int)floor(m->nValleys * (vwd ? rng_u01(rng) : rng_ulog(rng)));
  const orc_valley_t *v = &m->valleys[valley];
  const int sub = (int)floor(v->deg * (vwd ? rng_u01(rng) : rng_ulog(rng)));
  const double energy = -1.5 * d->thermalVoltage * log(rng_ulog(rng));
  const double r2 = rng_u01(rng);
  const double r1 = rng_u01(rng);
  double k[3];
  orc_random_direction(orc_norm_wave_vec(v, energy), r1, r2, k);
  for (int i = 0; i < d->dim; i++)
    if ((c[i] == 0 && k[i] < 0) || (c[i] == d->extent[i] - 1 && k[i] > 0))
      k[i] *= -1;
  /* electronVWD passes the valley index where the region index belongs (electronVWD.hpp:74, :87) */
  const double tau = -log(rng_ulog(rng)) * orc_tau(m, valley, vwd ? valley : region);
  const double gtau = -log(rng_ulog(rng)) * m->grainTau;
  if (slot >= 0) {
    out->kx[slot] = k[0]; out->ky[slot] = k[1]; out->kz[slot] = k[2];
    out->energy[slot] = energy;
    out->tau[slot] = tau;
    if (out->grainTau)
      out->grainTau[slot] = gtau;
    out->x[slot] = pos[0]; out->y[slot] = pos[1]; out->z[slot] = pos[2];
    out->valley[slot] = valley;
    out->sub[slot] = sub;
    out->region[slot] = region;
  }
}

int64_t orc_device_generate_initial_pot(const orc_model_t *m, const orc_device_t *d, double nrCarriers, uint64_t *mtState,
                                        const double *pot, orc_ensemble_t *out, int64_t capacity) {
  orc_rng_cfg_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.mode = ORC_RNG_MT_GLOBAL;
  cfg.mtState = mtState;
  rng_t rng;
  memset(&rng, 0, sizeof rng);
  rng.cfg = &cfg;
  const int64_t cells = orc_dev_cells(d);
  int64_t n = 0;
  for (int64_t cell = 0; cell < cells; cell++) {
    int64_t c[3];
    dev_coord(d, cell, c);
    double nr = orc_initial_nr_particles(d, cell, pot);
    while (nr >= 1) { /* emcAbstractParticleHandler.hpp:139-146 */
      dev_create_particle(m, d, c, &rng, out, n < capacity ? n : -1);
      n++;
      nr -= nrCarriers;
    }
    if (rng_u01(&rng) < nr) {
      dev_create_particle(m, d, c, &rng, out, n < capacity ? n : -1);
      n++;
    }
  }
  out->n = n <= capacity ? n : capacity;
  return n <= capacity ? n : -n;
}
int64_t orc_device_generate_initial(const orc_model_t *m, const orc_device_t *d, double nrCarriers, uint64_t *mtState,
                                    orc_ensemble_t *out, int64_t capacity) {
  return orc_device_generate_initial_pot(m, d, nrCarriers, mtState, NULL, out, capacity);
}

/* emcAbstractParticleHandler.hpp:238-249 driftParticle = drift + handleParticleAtBoundary
 * (emcParticleDrift.hpp:42-66, default specular reflection emcScatterHandler.hpp:172-191) + region update */
/* emcSurfaceScatterMechanism::scatterParticleSpecularly (:81-93): unlike the default reflection of the scatter handler,
 * k is only turned when it points out of the device */
static void surface_specular(const orc_device_t *d, double pos[3], double k[3]) {
  for (int i = 0; i < d->dim; i++) {
    if (pos[i] < 0) {
      pos[i] *= -1;
      if (k[i] < 0)
        k[i] *= -1;
    } else if (pos[i] > d->maxPos[i]) {
      pos[i] = 2 * d->maxPos[i] - pos[i];
      if (k[i] > 0)
        k[i] *= -1;
    }
  }
}
/* emcMomentumDependentSurfaceScatterMechanism::solveForTheta (:39-68) */
static double surface_solve_theta(double r, double height, double speed) {
  double c = pow(2 * height * speed, 2);
  double ee = exp(-c);
  double x = sqrt(r * (1 / (1 - ee) - 1 / c));
  double error = 1, tol = 1e-10;
  int it = 0;
  while (error > tol && it < 10000) {
    double co = cos(x), si = sin(x);
    double ecos = exp(-c * pow(co, 2));
    double esin = exp(-c * pow(si, 2));
    double x1 = x - (((ee - ecos) / c + pow(si, 2) - r * (1 - (1 - ee) / c)) * (2 * esin * c * si * co * (ecos - 1)) /
                     (ee * (c - 1) - 1));
    error = fabs(x1 - x);
    x = x1;
    it++;
  }
  return x;
}
/* emcSurfaceScatterMechanism::scatterParticle (:40-46) for the mechanism set on `face` */
static void surface_scatter(const orc_device_t *d, int face, double pos[3], double k[3], rng_t *rng) {
  const int kind = d->surfaceKind[face];
  const int perp = face / 2;
  double pDiff;
  if (kind == ORC_SURFACE_CONSTANT)
    pDiff = 1 - d->surfaceParam[face];
  else
    pDiff = 1 - (exp(-pow(2 * d->surfaceParam[face] * k[perp], 2)));
  if (rng_u01(rng) < pDiff) {
    double theta, phi;
    if (kind == ORC_SURFACE_CONSTANT) {
      theta = asin(sqrt(rng_u01(rng)));
      phi = 2 * C_PI * rng_u01(rng);
    } else {
      double speed = sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);
      theta = surface_solve_theta(rng_u01(rng), d->surfaceParam[face], speed);
      phi = 2 * C_PI * rng_u01(rng);
    }
    /* calculateAndAssignKAndPos (:127-147) */
    double speed = sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);
    double sign = face % 2 == 1 ? -1 : 1;
    k[perp] = speed * cos(theta) * sign;
    k[(perp + 1) % 3] = speed * sin(theta) * cos(phi);
    k[(perp + 2) % 3] = speed * sin(theta) * sin(phi);
  }
  surface_specular(d, pos, k);
}

static int dev_drift_particle(const orc_model_t *m, const orc_device_t *d, orc_ensemble_t *e, int64_t p, double dt,
                              const double force[3], rng_t *rng) {
  double k[3] = {e->kx[p], e->ky[p], e->kz[p]};
  double pos[3] = {e->x[p], e->y[p], d->dim > 2 ? e->z[p] : 0.};
  orc_drift(&m->valleys[e->valley[p]], dt, k, &e->energy[p], e->sub[p], pos, d->dim, force);
  int removed = 0, out = 0;
  for (int i = 0; i < d->dim; i++)
    if (pos[i] < 0 || pos[i] > d->maxPos[i])
      out = 1;
  if (out) {
    double clamped[3] = {0, 0, 0};
    for (int i = 0; i < d->dim; i++) {
      double lo = pos[i] < d->maxPos[i] ? pos[i] : d->maxPos[i]; /* std::max(0., std::min(pos, max)) */
      clamped[i] = 0. > lo ? 0. : lo;
    }
    const int64_t wallCell = dev_pos_to_cell(d, clamped);
    const int face = dev_first_face(d, wallCell);
    if (orc_dev_is_ohmic(d, wallCell)) {
      removed = 1;
      for (int i = 0; i < d->dim; i++)
        pos[i] = clamped[i];
    } else if (face >= 0 && d->surfaceKind[face] != ORC_SURFACE_SPECULAR) {
      surface_scatter(d, face, pos, k, rng);
    } else {
      for (int i = 0; i < d->dim; i++) {
        if (pos[i] < 0) {
          pos[i] = -pos[i];
          k[i] = -k[i];
        } else if (pos[i] > d->maxPos[i]) {
          pos[i] = 2 * d->maxPos[i] - pos[i];
          k[i] = -k[i];
        }
      }
    }
  }
  e->kx[p] = k[0]; e->ky[p] = k[1]; e->kz[p] = k[2];
  e->x[p] = pos[0]; e->y[p] = pos[1];
  if (d->dim > 2)
    e->z[p] = pos[2];
  if (!removed)
    e->region[p] = d->region[dev_pos_to_cell(d, pos)];
  return removed;
}

/* interpolateForce of the device's PM scheme: emcNGPScheme.hpp:51-66, emcCICScheme.hpp:122-173 (with its
 * (1 - wY)(1 - wY) weight of the upper-right point), emcNECScheme.hpp:99-113, NECSchemeVWD.hpp:57-76 */
static void dev_force(const orc_device_t *d, const orc_ensemble_t *e, int64_t p, const double *ef, double charge,
                      double force[3]) {
  const double pos[3] = {e->x[p], e->y[p], d->dim > 2 ? e->z[p] : 0.};
  const int64_t n = orc_dev_cells(d);
  const int64_t sx = 1, sy = d->extent[0], sz = (int64_t)d->extent[0] * d->extent[1];
  force[2] = 0;
  if (d->pmScheme == ORC_PM_NGP) {
    const int64_t cell = dev_pos_to_cell(d, pos);
    for (int i = 0; i < d->dim; i++)
      force[i] = charge * ef[(int64_t)i * n + cell];
    return;
  }
  int64_t c[3];
  double w[3];
  dev_floor_coord(d, pos, c, w);
  if (d->pmScheme == ORC_PM_CIC) {
    const int64_t base = dev_cell(d, c);
    const double wX = w[0], wY = w[1], wZ = w[2];
    for (int i = 0; i < d->dim; i++) {
      const double *E = ef + (int64_t)i * n;
      if (d->dim == 2)
        force[i] = charge * (E[base] * wX * wY + E[base + sx] * (1 - wX) * wY + E[base + sy] * wX * (1 - wY) +
                             E[base + sx + sy] * (1 - wY) * (1 - wY));
      else
        force[i] = (E[base] * wX * wY * wZ + E[base + sx] * (1 - wX) * wY * wZ + E[base + sy] * wX * (1 - wY) * wZ +
                    E[base + sx + sy] * (1 - wY) * (1 - wY) * wZ + E[base + sz] * wX * wY * (1 - wZ) +
                    E[base + sx + sz] * (1 - wX) * wY * (1 - wZ) + E[base + sy + sz] * wX * (1 - wY) * (1 - wZ) +
                    E[base + sx + sy + sz] * (1 - wY) * (1 - wY) * (1 - wZ)) *
                   charge;
    }
    return;
  }
  /* NEC (2-D only in the reference) */
  if (d->pmScheme == ORC_PM_NEC_VWD)
    c[0] = (int64_t)round(pos[0] / d->spacing[0]); /* "round x-position, instead of floor" */
  const int64_t base = dev_cell(d, c);
  const double *Ex = ef, *Ey = ef + n;
  force[0] = charge * (Ex[base] + Ex[base + sy]) / 2;
  if (d->pmScheme == ORC_PM_NEC_VWD && c[0] == d->extent[0] - 1)
    force[1] = charge * Ey[base];
  else
    force[1] = charge * (Ey[base] + Ey[base + sx]) / 2;
}

int orc_device_step(const orc_model_t *m, const orc_device_t *d, orc_ensemble_t *e, const double *ef, double charge,
                    double dt, int64_t stepIndex, const orc_rng_cfg_t *cfg, int8_t *removedOut,
                    int32_t *removedPerContact, int32_t *recPid, int64_t recCap, int64_t *recCount, int64_t *events,
                    int64_t evCap, int64_t *evCount) {
  rng_t rng;
  memset(&rng, 0, sizeof rng);
  rng.cfg = cfg;
  rng.recPid = recPid;
  rng.recCap = recCap;
  int64_t nEv = 0;
  orc_mech_t desc;
  for (int c = 0; c < d->nContacts; c++)
    removedPerContact[c] = 0;
  for (int64_t p = 0; p < e->n; p++) {
    rng.particle = p;
    rng.step = (uint64_t)stepIndex;
    rng.k = 0;
    double force[3];
    dev_force(d, e, p, ef, charge, force);
    int removed = 0;
    double tau = e->tau[p];
    if (tau >= dt) {
      removed = dev_drift_particle(m, d, e, p, dt, force, &rng);
    } else {
      removed = dev_drift_particle(m, d, e, p, tau, force, &rng);
      double tRem = dt - tau;
      while (tRem > 0 && !removed) {
        int si = find_set(m, e->valley[p], e->region[p]);
        int t = -2;
        if (si >= 0 && m->sets[si].nMech > 0) {
          double r = rng_u01(&rng);
          t = orc_select(m, si, e->energy[p], r);
          if (t >= 0) {
            fill_mech_desc(m, m->sets[si].mechIdx[t], &desc);
            scatter_with(m, &desc, e, p, &rng);
          }
          if (events && nEv < evCap) {
            events[nEv * 4 + 0] = stepIndex;
            events[nEv * 4 + 1] = p;
            events[nEv * 4 + 2] = t;
            events[nEv * 4 + 3] = t >= 0 ? m->sets[si].mechIdx[t] : -1;
          }
          nEv++;
        }
        double newTau = -log(rng_ulog(&rng)) * orc_tau(m, e->valley[p], e->region[p]);
        tau += newTau;
        dev_force(d, e, p, ef, charge, force);
        removed = dev_drift_particle(m, d, e, p, tRem < newTau ? tRem : newTau, force, &rng);
        tRem -= newTau;
      }
    }
    tau -= dt;
    e->tau[p] = tau;
    removedOut[p] = (int8_t)removed;
    if (removed) {
      const double pos[3] = {e->x[p], e->y[p], d->dim > 2 ? e->z[p] : 0.};
      removedPerContact[orc_dev_contact_idx(d, dev_pos_to_cell(d, pos))]++;
    }
    if (e->grainTau) { /* emcBasicParticleHandler.hpp:134-138 */
      e->grainTau[p] -= dt;
      if (e->grainTau[p] <= 0 && !removed)
        grain_event(m, e, p, &rng);
    }
  }
  if (recCount)
    *recCount = rng.recCount;
  if (evCount)
    *evCount = nEv;
  return 0;
}

static void ens_move(orc_ensemble_t *e, int64_t dst, int64_t src) {
  e->kx[dst] = e->kx[src]; e->ky[dst] = e->ky[src]; e->kz[dst] = e->kz[src];
  e->energy[dst] = e->energy[src];
  e->tau[dst] = e->tau[src];
  if (e->grainTau)
    e->grainTau[dst] = e->grainTau[src];
  e->x[dst] = e->x[src]; e->y[dst] = e->y[src]; e->z[dst] = e->z[src];
  e->valley[dst] = e->valley[src]; e->sub[dst] = e->sub[src]; e->region[dst] = e->region[src];
}

int64_t orc_compact(orc_ensemble_t *e, const int8_t *removed) {
  int64_t w = 0;
  for (int64_t p = 0; p < e->n; p++)
    if (!removed[p]) {
      if (w != p)
        ens_move(e, w, p);
      w++;
    }
  e->n = w;
  return w;
}

int64_t orc_contacts(const orc_model_t *m, const orc_device_t *d, orc_ensemble_t *e, int64_t capacity,
                     const double *expected, double nrCarriers, uint64_t *mtState, int32_t *net) {
  const int64_t cells = orc_dev_cells(d);
  double *have = (double *)calloc((size_t)cells, sizeof(double));
  for (int c = 0; c < d->nContacts; c++)
    net[c] = 0;
  /* :165-181: scan in index order, keep a particle while the cell is below its expected population */
  int64_t w = 0;
  for (int64_t p = 0; p < e->n; p++) {
    const double pos[3] = {e->x[p], e->y[p], d->dim > 2 ? e->z[p] : 0.};
    const int64_t cell = dev_pos_to_cell(d, pos);
    int keep = 1;
    if (orc_dev_is_reservoir(d, cell)) {
      if (have[cell] < expected[cell])
        have[cell] += nrCarriers;
      else {
        keep = 0;
        net[orc_dev_contact_idx(d, cell)]--;
      }
    }
    if (keep) {
      if (w != p)
        ens_move(e, w, p);
      w++;
    }
  }
  e->n = w;
  /* emcAbstractParticleHandler.hpp:200-216 generateInjectedParticles */
  orc_rng_cfg_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.mode = ORC_RNG_MT_GLOBAL;
  cfg.mtState = mtState;
  rng_t rng;
  memset(&rng, 0, sizeof rng);
  rng.cfg = &cfg;
  int64_t total = w;
  for (int64_t cell = 0; cell < cells; cell++) {
    if (!orc_dev_is_reservoir(d, cell))
      continue;
    int64_t c[3];
    dev_coord(d, cell, c);
    double diff = expected[cell] - have[cell];
    while (diff > 0) {
      dev_create_particle(m, d, c, &rng, e, total < capacity ? total : -1);
      total++;
      net[orc_dev_contact_idx(d, cell)]++;
      diff -= nrCarriers;
    }
  }
  free(have);
  e->n = total <= capacity ? total : capacity;
  return total <= capacity ? total : -total;
}
