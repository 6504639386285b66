/*
 * emcgpu.h -- C ABI of the B200-native ensemble Monte Carlo particle loop.
 *
 * This is the drop-in boundary for ONE path of ViennaEMC: the per-time-step
 * particle loop (free flight, null-scatter selection, final-state sampling,
 * per-step observables; for device runs additionally charge assignment and the
 * Poisson update).  The reference is a header-only C++17 library with no FFI of
 * its own; the entry points below are what its particle handlers bind to once
 * their bodies are replaced.  Each entry point cites the
 * reference interface it replaces (paths relative to the reference tree).
 *
 * Conventions
 *   - plain C, plain pointers and sizes, no C++/torch types;
 *   - every call returns an emcgpu_status (0 = ok); a human readable message
 *     for the last failure is available from emcgpu_last_error();
 *   - host buffers are caller-owned and copied during the call; device memory is
 *     owned by the context.  Pointers documented as DEVICE pointers must point
 *     to memory of the context's CUDA device;
 *   - all calls on one context must come from one host thread at a time (the
 *     reference drives its handlers from a single thread as well);
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     EMCGPU_E_CUDA, and a scatter mechanism without a device sampler is
 *     rejected with EMCGPU_E_UNSUPPORTED_MECHANISM (message carries getName()).
 */
#ifndef EMCGPU_H
#define EMCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMCGPU_ABI_VERSION 1
#define EMCGPU_MAX_VALLEYS 8
#define EMCGPU_MAX_SUBVALLEYS 8
#define EMCGPU_MAX_FINAL 8
#define EMCGPU_MAX_MECH_PER_SET 32
#define EMCGPU_MAX_TABLESETS 32
#define EMCGPU_NAME_LEN 48

typedef enum {
  EMCGPU_OK = 0,
  EMCGPU_E_INVALID = 1,               /* bad argument / call order            */
  EMCGPU_E_CUDA = 2,                  /* CUDA runtime failure, no device      */
  EMCGPU_E_UNSUPPORTED_MECHANISM = 3, /* mechanism has no device sampler      */
  EMCGPU_E_UNSUPPORTED_VALLEY = 4,
  EMCGPU_E_CAPACITY = 5,              /* fixed-size limit above exceeded      */
  EMCGPU_E_REPLAY_EXHAUSTED = 6       /* replay stream ran out of draws       */
} emcgpu_status;

/* Valley classes of include/ValleyTypes/ (3-D):
 *   emcParabolicIsotropValley.hpp, emcNonParabolicIsotropValley.hpp,
 *   emcParabolicAnisotropValley.hpp, emcNonParabolicAnistropValley.hpp */
typedef enum {
  EMCGPU_VALLEY_PARABOLIC_ISOTROP = 0,
  EMCGPU_VALLEY_NONPARABOLIC_ISOTROP = 1,
  EMCGPU_VALLEY_PARABOLIC_ANISOTROP = 2,
  EMCGPU_VALLEY_NONPARABOLIC_ANISOTROP = 3,
  /* single-layer (2-D material in the x-y plane) classes: bit 0 non-parabolic, bit 1 anisotropic, bit 2 single layer.
   *   emcParabolicIsotropSingleLayerValley.hpp, emcNonParabolicIsotropSingleLayerValley.hpp: Herring-Vogt factors
   *   (1, 1, 0) -- k_z and z never change;
   *   emcNonParabolicAnisotropSingleLayerValley.hpp: vogt = (sqrt(m_DOS/m_l), sqrt(m_DOS/m_t), 0), m_DOS = sqrt(m_l m_t)
   *   is the mass of the dispersion, of |k|(E) and of the velocity (:123-148), m_c = 2/(1/m_l + 1/m_t) that of the
   *   position update (getEffMassCond, :117-119); rot[s] = rows (cos a, -sin a, 0), (sin a, cos a, 0), (0, 0, 0) of the
   *   sub-valley's in-plane angle (:151-169).  (The reference has no parabolic anisotropic single-layer class.) */
  EMCGPU_VALLEY_PARABOLIC_ISOTROP_SINGLE_LAYER = 4,
  EMCGPU_VALLEY_NONPARABOLIC_ISOTROP_SINGLE_LAYER = 5,
  EMCGPU_VALLEY_NONPARABOLIC_ANISOTROP_SINGLE_LAYER = 7
} emcgpu_valley_kind;

/* One emcAbstractValley (include/ValleyTypes/emcAbstractValley.hpp:20-91) as the
 * numbers its virtuals return: getEffMassCond(0), getEffMassDOS(0),
 * getNonParabolicity(), getBottomEnergy(), getDegeneracyFactor(),
 * getVogtTransformationFactor() and, per sub-valley, the row-major rotation
 * used by transformToEllipseCoord (emcNonParabolicAnistropValley.hpp:140-153). */
typedef struct {
  int32_t kind; /* emcgpu_valley_kind */
  int32_t degeneracy;
  double effMassCond; /* [kg], band edge */
  double effMassDOS;  /* [kg], band edge */
  double alpha;       /* [1/eV], 0 for parabolic */
  double bottomEnergy; /* [eV] */
  double vogt[3];
  double rot[EMCGPU_MAX_SUBVALLEYS][9];
} emcgpu_valley_t;

/* Device final-state samplers, selected by mechanism ID.  Each replaces the
 * scatterParticle() of the named reference classes. */
typedef enum {
  EMCGPU_SAMPLER_NONE = 0, /* rejected: "no device sampler" */
  /* emcAcousticScatterMechanism.hpp:70-72 (elastic, isotropic) */
  EMCGPU_SAMPLER_ISOTROPIC_ELASTIC = 1,
  /* emcZeroOrderInterValleyScatterMechanism.hpp:119-129, :262-272 and
   * emcFirstOrderInterValleyScatterMechanism.hpp:121-131, :269-279:
   * valley <- finalValley, subValley <- finalSub[sub][raw % nFinal],
   * E += param[0] (signed: +(hw - dE_valley) absorption, -(hw + dE_valley)
   * emission), |k| <- k_norm(E) in the final valley, isotropic direction */
  EMCGPU_SAMPLER_INTERVALLEY = 2,
  /* emcCoulombScatterMechanism.hpp:48-59 (Brooks-Herring); param[0] = Debye
   * energy N_D/(c2 m_c) of the table set's region */
  EMCGPU_SAMPLER_COULOMB = 3,
  /* emcFroehlichInteraction.hpp:109-128, :183-201 and emcHotPhononFroehlichMechanism.hpp:94-117, :175-199 (polar optical,
   * unscreened): E += param[0] (signed phonon energy), cos(theta) = (1 + f - (1 + 2f)^r) / f with
   * f = 2 sqrt(E E') / (sqrt(E) - sqrt(E'))^2, new direction about the current k, |k| <- k_norm(E').
   * param[2] >= 0: phonon bath that counts the event in its |k' - k| bin (emission if param[0] < 0) */
  EMCGPU_SAMPLER_FROEHLICH = 4,
  /* emcScreenedFroehlichInteraction.hpp:140-153, :199-213, :272-294, :354-374: the same with the screened polar angle
   * (helpers :66-97): param[1] = qs^2 [1/m^2]; param[3] != 0: |q| drawn from the occupation-weighted window of bath
   * param[2] (emcPhononBath::sampleQ, q-resolved) instead of the closed form */
  EMCGPU_SAMPLER_SCREENED_FROEHLICH = 5,
  /* emcAcousticSingleLayerScatterMechanism.hpp:63-81 (elastic): in-plane angle 2 pi u, direction (cos/vogt_x, sin/vogt_y, 0)
   * normalised, |k| <- k_norm(E) of the particle's valley */
  EMCGPU_SAMPLER_SINGLE_LAYER_ELASTIC = 6,
  /* emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp:116-147, :293-324: valley <- finalValley; nFinal > 0:
   * subValley <- finalSub[sub][floor(u nFinal)] (nFinal = 0: the one-valley constructor, no draw); E += param[0]
   * (+(hw - dE_valley) absorption, -(dE_valley + hw) emission); then the direction of SINGLE_LAYER_ELASTIC in the final
   * valley.  param[1] != 0: the first-order classes (emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp:104-126, :255-275):
   * k_x = |k| cos, k_y = |k| sin without the Herring-Vogt weighting, k_z kept */
  EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY = 7,
  /* emcFroehlichInteractionSingleLayer.hpp (:45-80 sampleSingleLayerFroehlichDeflectionAngle, :149-168 / :273-292 the two
   * classes): E += param[0] (signed phonon energy); the in-plane direction of k is turned by psi, |k| <- k_norm(E'), k_z = 0.
   * psi by inversion of the 128-point cumulative sum of erfc(w q/2)^2 / (eps(q)^2 q), q^2 = k^2 + k'^2 - 2 k k' cos(psi),
   * eps(q) = 1 + q_s/q (emc2DScreening.hpp); one draw for the magnitude, one for the side.  param[1] = form-factor width w
   * [m], param[2] = 2-D screening wave vector q_s [1/m] (0: unscreened) */
  EMCGPU_SAMPLER_SINGLE_LAYER_FROEHLICH = 8,
  /* emcPiezoelectricSingleLayerScatterMechanism.hpp:110-139: elastic; deflection theta by the same inversion with the weight
   * erfc(w q/2)^2 / eps(q)^2, q = 2 k sin(theta/2); param[1], param[2] as above */
  EMCGPU_SAMPLER_SINGLE_LAYER_PIEZOELECTRIC = 9,
  /* The four other angle-resolved single-layer mechanisms share one final state: E += dE (elastic: unchanged), the in-plane
   * direction of k is turned by +-angle, |k| <- k_norm(E'), k_z = 0; the magnitude of the angle by inversion of an N-point
   * cumulative sum (midpoint rule on [0, pi]) of the mechanism's weight, pi u if the sum vanishes; one more draw for the side.
   * q = 2 k sin(angle/2) (elastic) or q^2 = k^2 + k'^2 - 2 k k' cos(angle); eps(q) = 1 + q_s/q; param[2] = q_s [1/m] in all.
   * emc2DChargedImpurityScatterMechanism.hpp:65-72, :107-139: elastic, N = 512, weight (exp(-q d) / (q_s + q + r0 q^2))^2;
   * param[0] = d [m] (impurity-to-sheet distance), param[1] = r0 [m] (Rytova-Keldysh length) */
  EMCGPU_SAMPLER_SINGLE_LAYER_CHARGED_IMPURITY = 10,
  /* emcSurfaceRoughnessScatterMechanism.hpp:59-63, :94-126: elastic, N = 256, weight exp(-q^2 Lambda^2/4) / eps(q)^2;
   * param[1] = Lambda^2 [m^2] */
  EMCGPU_SAMPLER_SINGLE_LAYER_SURFACE_ROUGHNESS = 11,
  /* emcRemoteSurfaceOpticalPhononMechanism.hpp:63-70, :112-149: param[0] = signed phonon energy, N = 128, weight
   * exp(-2 q d) / (q eps(q)^2); param[1] = d [m] (carrier-to-surface distance) */
  EMCGPU_SAMPLER_SINGLE_LAYER_REMOTE_SO = 12,
  /* emcScreenedIntravalleyOpticalMechanism.hpp:58-62, :104-141: param[0] = signed phonon energy, N = 128, weight 1/eps(q)^2 */
  EMCGPU_SAMPLER_SINGLE_LAYER_SCREENED_OPTICAL = 13
} emcgpu_sampler_id;
#define EMCGPU_MAX_BATHS 8

/* One emcScatterMechanism (include/ScatterMechanisms/emcScatterMechanism.hpp:17-53)
 * as seen by the device: which sampler, its parameters, its name for errors. */
typedef struct {
  int32_t sampler; /* emcgpu_sampler_id */
  int32_t finalValley;
  int32_t nFinal;
  int32_t mechId; /* caller's global mechanism index, echoed in event logs */
  double param[4];
  uint8_t finalSub[EMCGPU_MAX_SUBVALLEYS][EMCGPU_MAX_FINAL];
  char name[EMCGPU_NAME_LEN];
} emcgpu_mech_t;

/* The normalised cumulative scatter tables of one (valley, region) key exactly
 * as emcScatterHandler::renormalizeTables leaves them
 * (include/emcScatterHandler.hpp:248-273): cum[m][l], m in insertion order,
 * l = energy level ((l+1)*dE), all divided by the maximal cumulative rate;
 * tau = 1/that maximum. */
typedef struct {
  int32_t valley;
  int32_t region;
  int32_t nMech;
  int32_t reserved;
  double tau;
  const double *cum;         /* HOST, [nMech][nLevels] */
  const emcgpu_mech_t *mech; /* HOST, [nMech] */
} emcgpu_tableset_t;

/* SoA ensemble streams (replaces the AoS emcParticle<T> of include/emcParticle.hpp:10-18
 * plus the separate position vector, basicBulkParticleHandler.hpp:58-59). */
enum { EMCGPU_KX = 0, EMCGPU_KY, EMCGPU_KZ, EMCGPU_ENERGY, EMCGPU_TAU,
       EMCGPU_X, EMCGPU_Y, EMCGPU_Z, EMCGPU_N_STREAMS };
/* packed index word: valley | subValley << 8 | region << 16 */
#define EMCGPU_PACK(valley, sub, region) \
  ((uint32_t)(valley) | ((uint32_t)(sub) << 8) | ((uint32_t)(region) << 16))

/* math mode of the step kernels */
typedef enum {
  /* reference operation order, every op individually rounded (no FMA
   * contraction): the mode used for replay parity */
  EMCGPU_MATH_EXACT = 0,
  /* hoisted constants + FMA; agrees with EXACT to ~1e-15 per step */
  EMCGPU_MATH_FAST = 1
} emcgpu_math_mode;

typedef struct emcgpu_ctx emcgpu_ctx;

/* ---- life cycle ------------------------------------------------------- */
int emcgpu_abi_version(void);
/* cudaDevice: ordinal of the GPU this context lives on. */
int emcgpu_create(int cudaDevice, emcgpu_ctx **out);
void emcgpu_destroy(emcgpu_ctx *ctx);
/* message of the last failed call on ctx (ctx == NULL: last emcgpu_create failure) */
const char *emcgpu_last_error(const emcgpu_ctx *ctx);
/* number of kernels launched by this context so far (bench bookkeeping) */
int64_t emcgpu_launch_count(const emcgpu_ctx *ctx);
/* run all later work of ctx on this cudaStream_t (NULL = default stream) */
int emcgpu_set_stream(emcgpu_ctx *ctx, void *cudaStream);
int emcgpu_synchronize(emcgpu_ctx *ctx);
/* options: "vec" = particles per lane and loop iteration of the one-step kernel
 * (1, 2 or 4; default 2; no effect on results); "poisson_interval" = n: emcgpu_device_run*
 * solves Poisson only every n-th step (emcSimulation::setPoissonInterval, emcSimulation.hpp:80);
 * "sor_order" = 0: the reference's lexicographic Gauss-Seidel order (default; iterates and sweep counts
 * are the reference's), 1: red-black ordering (same equation and stopping rule, parallel, converges to
 * the same potential within the solver's accuracy); "sor_kernel" = 1 forces the general hyperplane
 * form of the lexicographic solver / the one-CTA form of the red-black solver, 2 the general cluster kernel of the
 * red-black solver, 3 its fast 2-D form on the portable cluster of 8 CTAs (default 0: the fastest form the grid and the
 * device allow -- red-black: 16 CTAs of 512 threads; no effect on results); "early_step" = 0: the kernels of a step of
 * emcgpu_device_run* as plain launches (default 1: a chain of programmatic dependent launches -- the next kernel's CTAs are
 * resident, the particle step's tables staged, while the predecessor still runs; same results); "assign_fp64" = 1: NEC / NEC-VWD charge
 * assignment with one fp64 atomic per corner instead of integer hits per mesh cell (same sums, slower); "multi_kernel": kernel of emcgpu_bulk_step*
 * with stepsPerLaunch > 1 -- 0 (default): for ensembles that fill the GPU the flight + event kernel pair (up to 24 steps
 * per launch pair; FAST arithmetic, one non-parabolic valley with signed-permutation rotations) or else the deferred-event
 * kernel (up to 8 steps per launch), events in place otherwise; 1: always in place; 2: always deferred; 3: always the
 * flight + event pair where the model allows (no effect on results: the trajectories are bit-identical);
 * "split_ppl" = 2 | 4: particles per lane of the flight kernel; "event_claim" = particles per claim of a warp of the event kernel
 * (a multiple of 256; work-distribution grain, no effect on results); "kernel_timing" = 1: see emcgpu_kernel_times; "defer_tables_smem" = 1: the deferred-event kernel
 * stages the rate tables in shared memory instead of reading them through L1/L2 (slower, no effect on results) */
int emcgpu_set_option(emcgpu_ctx *ctx, const char *name, int64_t value);

/* ---- physics model (built on the host by the reference-compatible API) - */
/* replaces the per-particle virtual calls into emcParticleType::valleys
 * (include/ParticleType/emcParticleType.hpp:34, :121-124) */
int emcgpu_set_valleys(emcgpu_ctx *ctx, const emcgpu_valley_t *valleys, int nValleys);
/* replaces the std::map look-ups of emcScatterHandler::scatterParticle
 * (include/emcScatterHandler.hpp:148-170) and getTau (:79-84).  Can be called
 * again at any time (emcScatterHandler::reinitScatterTables, :100-108). */
int emcgpu_set_tables(emcgpu_ctx *ctx, const emcgpu_tableset_t *sets, int nSets,
                      int nLevels, double maxEnergy);

/* emcGrainScatterMechanism (include/emcGrainScatterMechanism.hpp) + the grain clock of the particle handlers
 * (basicBulkParticleHandler.hpp:216-220, emcBasicParticleHandler.hpp:134-138): every particle carries a second exponential
 * clock; when it runs out the particle is reflected into the opposite (probability 1 - transmissionProbability) or
 * transmitted into the same hemisphere about its k (:40-77) and draws a new clock with mean 1 / scatterRate.
 * scatterRate <= 0 removes the mechanism.  The clocks (emcParticle::grainTau, HOST [n]) are uploaded after the ensemble;
 * with a grain mechanism bulk steps run on the general step kernel (the streaming one-step kernels do not carry the clock). */
int emcgpu_set_grain(emcgpu_ctx *ctx, double transmissionProbability, double scatterRate);
int emcgpu_set_grain_clock(emcgpu_ctx *ctx, const double *grainTau);
int emcgpu_get_grain_clock(emcgpu_ctx *ctx, double *grainTau);

/* emcPhononBath (include/emcPhononBath.hpp): |q|-binned occupation of a polar phonon mode coupled to the ensemble.  The
 * bath itself (update :264-358, occupations, relaxation) stays a host object of the drop-in API; the device side is
 *   - the event counters of recordEmission / recordAbsorption (:237-253), one pair per bin and bath, and
 *   - for the q-resolved polar angle, the prefix sums cumW / cumWN (:122-135) that sampleQ (:423-458) searches.
 * cumW / cumWN: HOST [nBaths][nBins + 1], may be NULL when no mechanism samples |q| from the bath.  Call again whenever
 * the bath was updated (like emcgpu_set_tables after reinitScatterTables). */
int emcgpu_set_phonon_baths(emcgpu_ctx *ctx, int nBaths, int nBins, double dq, const double *cumW, const double *cumWN);
/* event counts since the last call with reset != 0: HOST [nBaths][nBins] each */
int emcgpu_get_phonon_counts(emcgpu_ctx *ctx, int64_t *emission, int64_t *absorption, int reset);

/* ---- ensemble --------------------------------------------------------- */
/* upload n particles; soa[EMCGPU_N_STREAMS] are HOST arrays of length n.
 * particleIdBase: global id of particle 0 (keys the Philox streams so that
 * trajectories do not depend on how the ensemble is sharded across GPUs). */
int emcgpu_set_ensemble(emcgpu_ctx *ctx, int64_t n, const double *const *soa,
                        const uint32_t *packed, int64_t particleIdBase);
int emcgpu_get_ensemble(emcgpu_ctx *ctx, double *const *soa, uint32_t *packed);
int64_t emcgpu_ensemble_size(const emcgpu_ctx *ctx);
/* Device-side creation of a thermal bulk ensemble (same distributions as
 * emcElectron::generateInitialParticle, include/ParticleType/emcElectron.hpp:75-90,
 * emcParticleInitialization.hpp:14-51, positions uniform in the box), keyed by
 * Philox(seed, particleIdBase + i).  For ensembles too large to build on the host. */
int emcgpu_generate_bulk_ensemble(emcgpu_ctx *ctx, int64_t n, const double box[3],
                                  double temperature, int32_t region, uint64_t seed,
                                  int64_t particleIdBase);
/* DEVICE pointers to the SoA streams (for zero-copy consumers, e.g. dlpack) */
int emcgpu_ensemble_device_ptrs(emcgpu_ctx *ctx, double **soaOut /*[8]*/, uint32_t **packedOut);

/* ---- random numbers --------------------------------------------------- */
/* counter-based Philox4x32-10: draw i of particle p in step s is word pair
 * (i & 1) of Philox(key = seed, counter = (p_lo, p_hi, s, i >> 1)).  Replaces
 * the per-thread std::mt19937_64 of basicBulkParticleHandler.hpp:110-119. */
int emcgpu_rng_philox(emcgpu_ctx *ctx, uint64_t seed);
/* replay: particle p consumes draws[offsets[p]], draws[offsets[p]+1], ... (raw
 * 64-bit engine outputs recorded from the reference run).  HOST arrays;
 * offsets has n+1 entries. */
int emcgpu_rng_replay(emcgpu_ctx *ctx, const uint64_t *draws, const int64_t *offsets, int64_t n);

/* ---- bulk run: basicBulkParticleHandler (examples/bulkSimulation/
 *      basicBulkParticleHandler.hpp) -------------------------------------- */
/* ctor / resetAppliedFieldStrength (:93-137): periodic box, field = strength *
 * normalised(direction), force = charge * field (:186). */
int emcgpu_bulk_configure(emcgpu_ctx *ctx, const double box[3], const double fieldDirection[3],
                          double fieldStrength, double charge, int mathMode);
/* nSteps x { moveParticles(dt) (:181-225) ; getAvgEnergy, getAvgDriftVelocity,
 * getValleyOccupationProbability (:289-347) }.  obs (HOST, may be NULL) receives
 * [nSteps][nValleys][3] = { sum of energies, sum of v.E_dir, particle count }
 * per valley after each step (sums over THIS context's particles; divide after
 * the cross-GPU reduction). stepsPerLaunch > 1 keeps the state in registers for
 * that many consecutive steps. */
int emcgpu_bulk_step(emcgpu_ctx *ctx, double dt, int nSteps, int stepsPerLaunch, double *obs);
/* same, observables left on the device: obsDevice is a DEVICE buffer of
 * nSteps*nValleys*3 doubles that is zeroed and accumulated into; asynchronous
 * on the context's stream. */
int emcgpu_bulk_step_device(emcgpu_ctx *ctx, double dt, int nSteps, int stepsPerLaunch,
                            double *obsDevice);
/* Look-ahead for drivers that call one step at a time (basicBulkParticleHandler::moveParticles(dt),
 * examples/bulkSimulation/bulkSimulation.cpp:150-157): advances nSteps like emcgpu_bulk_step AND keeps the ensemble as it
 * was before the call, so that a host which turns out to need the state of an earlier step can emcgpu_bulk_rewind and
 * step again (the Philox streams are keyed by particle id and step: the same steps give the same trajectories).  With
 * the flight / event kernels the first flight launch simply writes a second set of streams (no copy); other kernels copy
 * the ensemble device-to-device first.  Philox streams only, no grain clocks.  Doubles the device memory of the ensemble. */
int emcgpu_bulk_step_ahead(emcgpu_ctx *ctx, double dt, int nSteps, int stepsPerLaunch, double *obs);
/* back to the ensemble and step index before the last emcgpu_bulk_step_ahead (valid once, and only directly after it) */
int emcgpu_bulk_rewind(emcgpu_ctx *ctx);
/* Per-particle velocities of every time step, streamed to the host (printDriftVelocities / printVelocities,
 * examples/bulkSimulation/basicBulkParticleHandler.hpp:251-285; consumed by examples/singleLayerMoS2/
 * calcMobilityFromVACF.py).  components = 1: v.Ê (projection on the field direction), 3: the velocity vector, 0: off.
 * While on, emcgpu_bulk_step / emcgpu_bulk_step_ahead write host[step][particle][component] for the steps of the call
 * (at most capacitySteps per call): the general step kernel stores them into one of two device buffers whose download
 * overlaps the next steps (asynchronous when `host` is pinned). */
int emcgpu_bulk_record_velocities(emcgpu_ctx *ctx, int components, double *host, int64_t capacitySteps);
/* with emcgpu_set_option("kernel_timing", 1): device time (cudaEvents on the launching stream) and launch counts of the
 * flight kernel [0], the event kernel [1] and all other bulk kernels [2] since the last reset; synchronises */
int emcgpu_kernel_times(emcgpu_ctx *ctx, double *ms, int64_t *launches, int reset);
/* the same nSteps x { moveParticles ; observables } for an ensemble that lives in HOST memory (soa / packed as
 * emcgpu_set_ensemble, updated IN PLACE).  The arrays should be PINNED (cudaHostAlloc / cudaHostRegister): only then
 * are the copies asynchronous and overlap the kernels; with pageable arrays the call is still correct, but every
 * cudaMemcpyAsync blocks the host thread and the copies and kernels run one after the other.  The ensemble is cut into
 * slices of sliceParticles (<= 0: about n/8, n/16 for runs of fewer than 128 steps) and slice i runs its nSteps steps while slice i+1 is copied to the
 * device and slice i-1 back, so the PCIe transfers hide behind the step kernels and n is not limited by the HBM
 * size. Results equal emcgpu_set_ensemble + emcgpu_bulk_step + emcgpu_get_ensemble (particle states bit for bit,
 * the Philox stream of a particle is keyed by particleIdBase + index; obs = the same sums in another order).
 * Philox streams only; the context's resident ensemble is not touched; the step index advances by nSteps. */
int emcgpu_bulk_run_host(emcgpu_ctx *ctx, int64_t n, double *const *soa, uint32_t *packed, int64_t particleIdBase,
                         double dt, int nSteps, int stepsPerLaunch, int64_t sliceParticles, double *obs);
/* observables of the current state without moving (:289-347), HOST out [nValleys][3] */
int emcgpu_bulk_observables(emcgpu_ctx *ctx, double *obs);
/* index of the next time step (Philox counter word); starts at 1 like
 * bulkSimulation.cpp:150 and advances by nSteps per emcgpu_bulk_step call */
int emcgpu_set_step_index(emcgpu_ctx *ctx, int64_t nextStep);
int64_t emcgpu_get_step_index(const emcgpu_ctx *ctx);

/* ---- device run: emcSimulation + emcBasicParticleHandler + emcNGPScheme + emcSORSolver --------------
 * (include/emcSimulation.hpp:139-193, include/ParticleHandler/emcBasicParticleHandler.hpp,
 *  include/PMSchemes/emcNGPScheme.hpp, include/PMSchemes/emcEFieldCalculation.hpp,
 *  include/PoissonSolver/emcSORSolver.hpp, include/emcSimulationResults.hpp:98-116)
 * The box device with its doping regions and contacts crosses the boundary as flat arrays: */
#define EMCGPU_MAX_CONTACTS 16
typedef enum { EMCGPU_CONTACT_OHMIC = 0, EMCGPU_CONTACT_SCHOTTKY = 1, EMCGPU_CONTACT_GATE = 2 } emcgpu_contact_type;
/* particle-mesh schemes: include/PMSchemes/emcNGPScheme.hpp, emcCICScheme.hpp, emcNECScheme.hpp (2-D) and
 * examples/mosfet2D/NECSchemeVWD.hpp (2-D).  The scheme selects the variants of the charge-assignment kernel, of the
 * force gather in the particle step and of E = -grad(phi). */
typedef enum { EMCGPU_PM_NGP = 0, EMCGPU_PM_CIC = 1, EMCGPU_PM_NEC = 2, EMCGPU_PM_NEC_VWD = 3 } emcgpu_pm_scheme;
/* wall mechanisms per face: none = the specular reflection of emcScatterHandler.hpp:172-191;
 * SurfaceScatterMechanisms/emcConstantSurfaceScatterMechanism.hpp (parameter: specularity),
 * emcMomentumDependentSurfaceScatterMechanism.hpp (parameter: rms roughness height [m]) */
typedef enum { EMCGPU_SURFACE_SPECULAR = 0, EMCGPU_SURFACE_CONSTANT = 1, EMCGPU_SURFACE_MOMENTUM_DEPENDENT = 2 } emcgpu_surface_kind;
/* creation rules of injected particles: ParticleType/emcElectron.hpp:92-104 or examples/mosfet2D/electronVWD.hpp:78-90
 * (valley draws from U[0,1), tau looked up with the valley index as region) */
typedef enum { EMCGPU_PARTICLE_ELECTRON = 0, EMCGPU_PARTICLE_ELECTRON_VWD = 1 } emcgpu_particle_kind;

typedef struct {
  int32_t dim; /* 2 or 3 */
  int32_t nContacts;
  int32_t extent[3];   /* grid points per dimension (emcDevice::getGridExtent) */
  int32_t pmScheme;    /* emcgpu_pm_scheme */
  double spacing[3];   /* [m] */
  double maxPos[3];    /* [m] */
  double thermalVoltage, debyeLength, ni, cellVolume, epsR; /* emcDevice.hpp:87-90, :333-343 */
  const int32_t *contactType;                               /* HOST [nContacts], emcgpu_contact_type */
  const double *contactVoltage;                             /* HOST [nContacts], volts */
  const double *gateEpsOx, *gateThickness, *gateBarrier;    /* HOST [nContacts] (gate contacts) */
  const int32_t *region;     /* HOST [cells], x fastest: emcDopingProfile::getDopingRegionIdx */
  const int8_t *faceContact; /* HOST [cells][2*dim] in face order XMIN XMAX YMIN YMAX ZMIN ZMAX:
                                -2 cell not on that face, -1 artificial boundary, >= 0 contact index
                                (emcSurface::idxContactGrid incl. the corner sharing of updateAllOccurences) */
  const double *doping;      /* HOST [cells], 1/m^3 */
} emcgpu_device_t;

typedef enum {
  EMCGPU_GRID_POTENTIAL = 0, /* normalised by Vt */
  EMCGPU_GRID_CONCENTRATION, /* normalised by Ni (particle type of this context) */
  EMCGPU_GRID_COUNT,         /* carriers per grid point (emcSimulationResults::nrPart) */
  EMCGPU_GRID_EFIELD_X,
  EMCGPU_GRID_EFIELD_Y,
  EMCGPU_GRID_EFIELD_Z,
  EMCGPU_GRID_EXPECTED,      /* expected reservoir population per contact cell (expNrPart) */
  EMCGPU_GRID_SUM_POTENTIAL,     /* running sums of emcSimulationResults::updateAverageCharacteristics (:87-93) */
  EMCGPU_GRID_SUM_CONCENTRATION,
  EMCGPU_N_GRIDS
} emcgpu_grid_id;

/* Device + particle charge + carriers per simulated particle.  Allocates the device-resident grids;
 * the potential starts as asinh(doping / 2 Ni) (emcSimulationResults.hpp:208-213).  `expected` (HOST,
 * [cells], may be NULL = cellVolume * doping * 1/2 per boundary dimension at reservoir cells,
 * emcElectron.hpp:63-73) is the population the contact cells are kept at. */
int emcgpu_device_configure(emcgpu_ctx *ctx, const emcgpu_device_t *device, double charge, double nrCarriersPerParticle,
                            const double *expected, int mathMode);
/* emcParticleType::setSurfaceScatterMechanism (emcParticleType.hpp:159-166): wall mechanism of one face
 * (0 XMIN, 1 XMAX, 2 YMIN, 3 YMAX, 4 ZMIN, 5 ZMAX) */
int emcgpu_device_set_surface(emcgpu_ctx *ctx, int face, int kind, double parameter);
/* which generateInjectedParticle the contacts use (emcgpu_particle_kind) */
int emcgpu_device_set_particle_kind(emcgpu_ctx *ctx, int kind);
int emcgpu_device_set_grid(emcgpu_ctx *ctx, int grid, const double *host);
int emcgpu_device_get_grid(emcgpu_ctx *ctx, int grid, double *host);
/* room for particles injected at contacts: the ensemble is re-allocated for at least this many */
int emcgpu_device_reserve(emcgpu_ctx *ctx, int64_t capacity);
/* emcSORSolver::calcEquilibriumPotential (:49-128, equilibrium != 0) / calcNonEquilibriumPotential (:131-197)
 * on the POTENTIAL grid, with the CONCENTRATION grid as electron density; accuracy in volts.  Same update
 * order as the reference (hyperplane sweep).  sweeps (HOST, may be NULL) receives the sweep count. */
int emcgpu_device_poisson(emcgpu_ctx *ctx, int equilibrium, double accuracyVolt, double omega, int resetBC,
                          int32_t *sweeps);
/* pmScheme.calcEField (emcNGPScheme.hpp:69-73): POTENTIAL -> EFIELD_* */
int emcgpu_device_efield(emcgpu_ctx *ctx);
/* handler.assignParticlesToMesh (emcNGPScheme.hpp:36-47): ensemble -> COUNT (zeroed first) */
int emcgpu_device_assign(emcgpu_ctx *ctx);
/* results.updateCurrentParticleConcentrations (:98-116): COUNT -> CONCENTRATION */
int emcgpu_device_concentration(emcgpu_ctx *ctx);
/* handler.driftScatterParticles(dt, eField) (:76-145): one time step of every particle in EFIELD_*, particles
 * that leave through an ohmic contact are removed (order of the others kept).  removedPerContact: HOST
 * [nContacts] out.  Uses the context's rng (Philox: step index advances by one per call). */
int emcgpu_device_step(emcgpu_ctx *ctx, double dt, int32_t *removedPerContact);
/* handler.handleOhmicContacts() (:158-192): delete the excess particles of every reservoir cell (first come,
 * first kept, in index order), inject the missing ones (appended, cells in storage order).
 * netPerContact: HOST [nContacts] out = injected - deleted.  replayDraws (HOST, may be NULL): the raw draws the
 * reference consumed in this call, (dim + 7) per injected particle. */
int emcgpu_device_contacts(emcgpu_ctx *ctx, int32_t *netPerContact, const uint64_t *replayDraws, int64_t nReplayDraws);
/* nSteps x performEMCStep (emcSimulation.hpp:177-192): poisson(resetBC only in the first step when asked) ->
 * efield -> step -> contacts -> assign -> concentration, everything device resident.  counters (HOST, may be
 * NULL): [nSteps][2][nContacts] = {left through contact, injected - deleted} per step; sweeps (HOST, may be
 * NULL): [nSteps] SOR sweeps per step. */
int emcgpu_device_run(emcgpu_ctx *ctx, double dt, int nSteps, double accuracyVolt, double omega, int resetBCFirst,
                      int32_t *counters, int32_t *sweeps);
/* the same; the last nAverage steps also add POTENTIAL / CONCENTRATION to the SUM_* grids
 * (results.updateAverageCharacteristics, emcSimulation.hpp:122-123) */
int emcgpu_device_run_averaging(emcgpu_ctx *ctx, double dt, int nSteps, int nAverage, double accuracyVolt, double omega,
                                int resetBCFirst, int32_t *counters, int32_t *sweeps);

/* ---- device run on an ensemble sharded over several GPUs ------------------------------------------
 * One context per GPU holds a block of the particles; potential, field and concentration are replicated.  Per step the
 * ranks exchange (a) how many reservoir particles each holds per cell -- handleOhmicContacts keeps the FIRST particles of
 * a cell in global index order, rank r's particles counting before rank r+1's, and the particles a cell is missing are
 * injected by the ranks in even shares -- and (b) the carriers per grid point after the charge assignment.  Both are
 * in-place sums over the ranks of a DEVICE buffer of doubles, done by the callback on the given stream (ncclAllReduce
 * over NVLink in production).  The Poisson solve that follows is replicated and bitwise identical on
 * every rank.  Counters returned by emcgpu_device_run* are this rank's; sum them over the ranks. */
typedef void (*emcgpu_allreduce_fn)(void *user, double *deviceBuffer, int64_t count, void *cudaStream);
int emcgpu_device_set_sharding(emcgpu_ctx *ctx, int rank, int world, emcgpu_allreduce_fn allreduceSum, void *user);

/* ---- diagnostics used by the parity tests ------------------------------ */
/* log up to capacity scatter selections as (step, particleId, tableIndex or -1
 * for self-scattering, mechId or -1); 0 disables. */
int emcgpu_event_log_enable(emcgpu_ctx *ctx, int64_t capacity);
/* copies the log (HOST, [count][4] int64) and returns the number of events
 * seen (may exceed capacity); resets the log. */
int64_t emcgpu_event_log_read(emcgpu_ctx *ctx, int64_t *out, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* EMCGPU_H */
