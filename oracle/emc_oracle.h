/* TEST INFRASTRUCTURE (oracle/).  CPU restatement of ViennaEMC's per-time-step
 * particle loop, used ONLY as the checker by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.  Never linked into, imported by or executed
 * from the product path (viennaemc_b200/, include/).
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against
 * outputs of the unmodified reference (oracle/_ref/ref_bulk_driver, built from
 * /root/reference by oracle/Makefile; vectors committed under tests/golden/ by
 * oracle/make_golden.py) and against the reference's own known-answer tests
 * (tests/testParticleMovement, tests/testValleyCoordinateTransformation).
 *
 * All file:line citations are relative to the reference tree.
 */
#ifndef EMC_ORACLE_H
#define EMC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_SUB 8
#define ORC_MAX_FINAL 8

/* bit 0: non-parabolic, bit 1: anisotropic (sub-valley frames), bit 2: single layer (2-D material in the x-y plane:
 * ValleyTypes/emcParabolicIsotropSingleLayerValley.hpp, emcNonParabolicIsotropSingleLayerValley.hpp,
 * emcNonParabolicAnisotropSingleLayerValley.hpp; the reference has no parabolic anisotropic single-layer class) */
enum { ORC_VALLEY_PARABOLIC_ISO = 0, ORC_VALLEY_NONPARABOLIC_ISO = 1,
       ORC_VALLEY_PARABOLIC_ANISO = 2, ORC_VALLEY_NONPARABOLIC_ANISO = 3,
       ORC_VALLEY_PARABOLIC_ISO_SL = 4, ORC_VALLEY_NONPARABOLIC_ISO_SL = 5, ORC_VALLEY_NONPARABOLIC_ANISO_SL = 7 };

enum { ORC_SAMPLER_NONE = 0, ORC_SAMPLER_ISOTROPIC_ELASTIC = 1,
       ORC_SAMPLER_INTERVALLEY = 2, ORC_SAMPLER_COULOMB = 3,
       /* emcFroehlichInteraction.hpp:109-128 / :183-201, emcHotPhononFroehlichMechanism.hpp:94-117 / :175-199:
        * p[0] = signed phonon energy, p[2] = phonon bath index or -1 */
       ORC_SAMPLER_FROEHLICH = 4,
       /* emcScreenedFroehlichInteraction.hpp:140-153 ... :354-374: p[0] = signed phonon energy, p[1] = qs^2,
        * p[2] = phonon bath index or -1, p[3] = 1: polar angle through bath.sampleQ (q-resolved) */
       ORC_SAMPLER_SCREENED_FROEHLICH = 5,
       /* emcAcousticSingleLayerScatterMechanism.hpp:63-81: elastic, in-plane angle 2 pi u weighted by the Herring-Vogt
        * factors of the particle's valley, k_z = 0 */
       ORC_SAMPLER_SL_ELASTIC = 6,
       /* emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp:116-147, :293-324: valley <- finalValley; nFinal > 0:
        * sub <- finalSub[sub][floor(u nFinal)] (one draw; none for the one-valley constructor); E += p[0]; then the
        * direction of SL_ELASTIC in the FINAL valley.  p[1] != 0 (emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp
        * :104-126, :255-275): k_x = |k| cos, k_y = |k| sin without the Herring-Vogt weighting, k_z kept */
       ORC_SAMPLER_SL_INTERVALLEY = 7,
       /* emcFroehlichInteractionSingleLayer.hpp (:45-80 the angle, :149-168 / :273-292 the classes): E += p[0] (signed phonon
        * energy); deflection psi about the in-plane direction of k by inversion of a 128-point cumulative sum of
        * erfc(w q/2)^2 / (eps(q)^2 q), q^2 = k^2 + k'^2 - 2 k k' cos(psi); p[1] = form-factor width w, p[2] = screening q_s */
       ORC_SAMPLER_SL_FROEHLICH = 8,
       /* emcPiezoelectricSingleLayerScatterMechanism.hpp:110-139: elastic; weight erfc(w q/2)^2 / eps(q)^2, q = 2 k sin(theta/2) */
       ORC_SAMPLER_SL_PIEZO = 9,
       /* one final state for four mechanisms: E += dE (elastic: none), in-plane direction turned by +-angle, the magnitude by
        * inversion of an N-point cumulative sum of the weight (pi u if the sum vanishes), one more draw for the side; p[2] = q_s.
        * emc2DChargedImpurityScatterMechanism.hpp:107-139: elastic, N = 512, (exp(-q d) / (q_s + q + r0 q^2))^2; p[0] = d, p[1] = r0 */
       ORC_SAMPLER_SL_CHARGED_IMPURITY = 10,
       /* emcSurfaceRoughnessScatterMechanism.hpp:94-126: elastic, N = 256, exp(-q^2 Lambda^2/4) / eps(q)^2; p[1] = Lambda^2 */
       ORC_SAMPLER_SL_SURFACE_ROUGHNESS = 11,
       /* emcRemoteSurfaceOpticalPhononMechanism.hpp:112-149: p[0] = signed phonon energy, N = 128, exp(-2 q d) / (q eps^2); p[1] = d */
       ORC_SAMPLER_SL_REMOTE_SO = 12,
       /* emcScreenedIntravalleyOpticalMechanism.hpp:104-141: p[0] = signed phonon energy, N = 128, 1 / eps(q)^2 */
       ORC_SAMPLER_SL_SCREENED_OPTICAL = 13 };

enum { ORC_RNG_MT_GLOBAL = 0, ORC_RNG_STREAMS = 1, ORC_RNG_PHILOX = 2 };

typedef struct {
  int32_t kind, deg;
  double mCond, mDos, alpha, eBottom;
  double vogt[3];
  double rot[ORC_MAX_SUB][9];
} orc_valley_t;

typedef struct {
  int32_t sampler, finalValley, nFinal, globalId;
  double p[4]; /* INTERVALLEY: p[0] = signed energy shift; COULOMB: p[0] = Debye energy */
  int32_t finalSub[ORC_MAX_SUB][ORC_MAX_FINAL];
} orc_mech_t;

typedef struct {
  int64_t n;
  double *kx, *ky, *kz, *energy, *tau, *grainTau, *x, *y, *z;
  int32_t *valley, *sub, *region;
} orc_ensemble_t;

typedef struct {
  int32_t mode;
  /* MT_GLOBAL: one std::mt19937_64-equivalent stream shared by all particles,
   * consumed in particle order like the reference's single-thread rngs[0] */
  uint64_t *mtState; /* [313], see orc_mt_seed */
  /* STREAMS (replay): particle p consumes draws[offsets[p] + cursor[p]++] */
  const uint64_t *draws;
  const int64_t *offsets; /* [n+1] */
  int64_t *cursor;        /* [n], in/out */
  /* PHILOX: key = seed, counter = (particle id, step, draw index / 2) */
  uint64_t philoxSeed;
  int64_t particleIdBase; /* global id of local particle 0 (multi-GPU shards) */
} orc_rng_cfg_t;

typedef struct orc_model orc_model_t;

/* ---- model construction (host-side math of the reference) ---- */
orc_model_t *orc_model_create(int nLevels, double maxEnergy, double temperature,
                              double rho, double vSound);
void orc_model_destroy(orc_model_t *m);
void orc_model_set_electron2d(orc_model_t *m, int perGridPoint); /* examples/singleLayerMoS2/electron2D.hpp */
void orc_model_set_init_energy(orc_model_t *m, double energyEV); /* emcElectron.hpp:85-88, emcHole.hpp:95-98 */
/* single-layer mechanisms; density2D [kg/m^2], the model's temperature */
int orc_add_acoustic_sl(orc_model_t *m, int valley, int region, double sigma, double density2D, double vSound);
int orc_add_intervalley_sl(orc_model_t *m, int order, int emission, int valley, int finalValley, int region, double sigma,
                           double density2D, double phE, int nInitSub, int nFinal, const int32_t *finalSub);
int orc_add_froehlich_sl(orc_model_t *m, int emission, int valley, int region, double phE, double couplingConst, double width,
                         double qs);
int orc_add_piezo_sl(orc_model_t *m, int valley, int region, double piezoConst, double width, double density2D, double vSound,
                     double qs);
int orc_add_charged_impurity_sl(orc_model_t *m, int valley, int region, double impurityDensity, double epsAvg, double qs,
                                double rytovaKeldyshLength, double remoteDistance, double chargeNumber);
int orc_add_surface_roughness_sl(orc_model_t *m, int valley, int region, double effectiveField, double roughnessAmplitude,
                                 double correlationLength, double qs);
int orc_add_remote_so_sl(orc_model_t *m, int emission, int valley, int region, double phE, double couplingD,
                         double remoteDistance, double qs);
int orc_add_screened_optical_sl(orc_model_t *m, int emission, int valley, int region, double sigma, double density2D, double phE,
                                double qs);
int orc_add_valley(orc_model_t *m, int kind, const double relMass[3],
                   double particleMass, int deg, double alpha, double eBottom,
                   const double *dirs /* [deg][3][3] un-normalised or NULL */);
int orc_add_acoustic(orc_model_t *m, int valley, int region, double sigma);
int orc_add_intervalley(orc_model_t *m, int order, int emission, int valley,
                        int finalValley, int region, double defPot,
                        double phononEnergy, int nInitSub, int nFinal,
                        const int32_t *finalSub /* [nInitSub][nFinal] */);
int orc_add_coulomb(orc_model_t *m, int valley, int region, double epsR,
                    double regionDoping);
int orc_build_tables(orc_model_t *m);
/* emcGrainScatterMechanism (include/emcGrainScatterMechanism.hpp): transmission probability, scatter rate [1/s];
 * rate <= 0: no grain mechanism (the clock still runs with grainTau = 1 s, emcScatterHandler.hpp:62) */
void orc_model_set_grain(orc_model_t *m, double transmissionProb, double scatterRate);

/* ---- polar-optical (Froehlich) family and the phonon bath (config 5, rows a11 / a21) ----------------------------
 * emcPhononBath (include/emcPhononBath.hpp): q-binned LO occupation, event counters, relaxation. */
typedef struct orc_bath orc_bath_t;
orc_bath_t *orc_bath_create(int nBins, double dq, double tauLO, double phononEnergy, double latticeTemp, double Vsim,
                            int enableAcoustic, double acPhononEnergy, double tauAcoustic, double wRidley,
                            double toPhononEnergy, double tauTO);             /* ctor :161-196 */
void orc_bath_destroy(orc_bath_t *b);
void orc_bath_set_qs2(orc_bath_t *b, double qs2);                               /* setScreeningQ2 :200-205 */
void orc_bath_record(orc_bath_t *b, double q, int emission);                     /* :237-253 */
void orc_bath_update(orc_bath_t *b, double dt);                                 /* :264-358 */
double orc_bath_mean_nq(const orc_bath_t *b);                                    /* :461-470 */
double orc_bath_nq_window(const orc_bath_t *b, double qMin, double qMax);        /* :378-394 */
void orc_set_limit_flags(unsigned char *flags); /* test aid: see emc_oracle.c */
double orc_bath_sample_q(const orc_bath_t *b, double qMin, double qMax, int emission, double r); /* :423-458 */
double orc_bath_acoustic_temp(const orc_bath_t *b);                             /* :477-481 */
double orc_bath_n0(const orc_bath_t *b);
/* which: 0 Nq, 1 nEm, 2 nAbs (nBins doubles); 3 cumW, 4 cumWN (nBins + 1 doubles) */
int orc_bath_copy(const orc_bath_t *b, int which, double *out);
/* add event counts gathered elsewhere (the device) to nEm / nAbs */
void orc_bath_add_counts(orc_bath_t *b, const int64_t *emission, const int64_t *absorption);
/* the model refers to baths by index; the caller keeps ownership */
int orc_model_add_bath(orc_model_t *m, orc_bath_t *b);
/* emcPlasmonScreening::getQs2 as seen by the screened mechanisms */
void orc_model_set_qs2(orc_model_t *m, double qs2);
double orc_plasmon_qs2(double density, double carrierTemp, double epsStatic); /* emcPlasmonScreening.hpp:69-76 */
/* variant: 0 emcFroehlich{Absorption,Emission}3D, 1 emcHotPhononFroehlich*, 2 emcScreenedFroehlich*,
 * 3 emcScreenedHotPhononFroehlich* (qResolved / qResolvedAngle as its ctor); bath: index or -1 */
int orc_add_froehlich(orc_model_t *m, int variant, int emission, int valley, int region, double phononEnergy,
                      double relEffMass, double epsHi, double epsLo, double temperature, int bath, int qResolved,
                      int qResolvedAngle);

int orc_n_valleys(const orc_model_t *m);
int orc_get_valley(const orc_model_t *m, int v, orc_valley_t *out);
int orc_n_mechanisms(const orc_model_t *m);
double orc_raw_rate(const orc_model_t *m, int globalMech, double energy);
int orc_n_tablesets(const orc_model_t *m);
int orc_tableset_info(const orc_model_t *m, int i, int32_t *valley,
                      int32_t *region, int32_t *nMech, double *tau);
int orc_tableset_copy(const orc_model_t *m, int i, double *cum /*[nMech][nLevels]*/,
                      orc_mech_t *mech /*[nMech]*/);
double orc_tau(const orc_model_t *m, int valley, int region);
double orc_dE(const orc_model_t *m);

/* ---- valley math / kernels of the path ---- */
double orc_energy(const orc_valley_t *v, const double k[3]);
double orc_norm_wave_vec(const orc_valley_t *v, double energy);
double orc_gamma(const orc_valley_t *v, double energy);
double orc_eff_mass_cond(const orc_valley_t *v, double energy);
void orc_velocity(const orc_valley_t *v, const double k[3], double energy,
                  int sub, double out[3]);
void orc_to_ellipse(const orc_valley_t *v, int sub, const double in[3], double out[3]);
void orc_to_device(const orc_valley_t *v, int sub, const double in[3], double out[3]);
void orc_drift(const orc_valley_t *v, double dt, double k[3], double *energy,
               int sub, double pos[3], int dim, const double force[3]);
void orc_random_direction(double norm, double rand1, double rand2, double out[3]);
void orc_random_direction_wrt_k(const double k[3], double cosTheta, double rand,
                                double out[3]);
double orc_uniform(uint64_t raw, double a, double b);
int orc_energy_level(const orc_model_t *m, double energy);
int orc_select(const orc_model_t *m, int tableset, double energy, double r);

/* ---- RNG primitives ---- */
void orc_mt_seed(uint64_t *state /*[313]*/, uint64_t seed);
uint64_t orc_mt_next(uint64_t *state);
void orc_mt_fill(uint64_t seed, uint64_t *out, int64_t n);
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
uint64_t orc_philox_draw(uint64_t seed, uint64_t particle, uint64_t step, uint64_t idx);

/* ---- ensemble ---- */
/* reference: basicBulkParticleHandler::generateInitialParticles (:143-159) with a
 * single constant doping region; draws come from the given mt19937_64 state.
 * Returns number of particles created, or -needed if capacity is too small.
 * drawsConsumed (optional) receives the number of raw draws used. */
int64_t orc_generate_initial(const orc_model_t *m, const double box[3],
                             const int32_t cells[3], double doping,
                             uint64_t *mtState, orc_ensemble_t *out,
                             int64_t capacity, int64_t *drawsConsumed);

/* One or more time steps of basicBulkParticleHandler::moveParticles (:181-225)
 * followed by the three observable passes (:289-347).
 *  obs: [nSteps][nValleys][3] = {sum E, sum v.Edir, count} (may be NULL)
 *  recPid/recCap/recCount: optional log of which particle consumed each draw
 *  events/evCap/evCount: optional log (step, particle, mechIndexInTableset or -1, globalMech or -1)
 */
int orc_bulk_steps(const orc_model_t *m, orc_ensemble_t *ens, const double box[3],
                   const double fieldDir[3] /* un-normalised ok */, double fieldStrength,
                   double charge, double dt, int nSteps,
                   int64_t firstStepIndex, const orc_rng_cfg_t *rng,
                   double *obs, int32_t *recPid,
                   int64_t recCap, int64_t *recCount, int64_t *events,
                   int64_t evCap, int64_t *evCount);

/* observables only (reference :289-347), sums not yet divided */
void orc_bulk_observables(const orc_model_t *m, const orc_ensemble_t *ens,
                          const double fieldDir[3], double *obs /*[nValleys][3]*/);

#ifdef __cplusplus
}
#endif

/* ======================================================================================
 * Device-run path.  Parity status: PINNED against
 * oracle/_ref/ref_device_driver (unmodified reference headers; fixtures tests/golden/device_*.npz).
 * ====================================================================================== */
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_CONTACT_OHMIC = 0, ORC_CONTACT_SCHOTTKY = 1, ORC_CONTACT_GATE = 2 }; /* emcContact.hpp */

typedef struct {
  int32_t dim; /* 2 or 3 */
  int32_t extent[3];
  double spacing[3], maxPos[3];
  double thermalVoltage, debyeLength, ni, cellVolume, epsR;
  int32_t nContacts;
  const int32_t *contactType;
  const double *contactVoltage, *gateEpsOx, *gateThickness, *gateBarrier; /* [nContacts] */
  const int32_t *region;     /* [cells], x fastest */
  const int8_t *faceContact; /* [cells][2*dim]: -2 cell not on that face, -1 artificial boundary, >= 0 contact */
  const double *doping;      /* [cells], 1/m^3 */
  /* --- plug-in variants; all zero = emcNGPScheme + emcElectron + default specular walls --- */
  int32_t pmScheme;       /* ORC_PM_* */
  int32_t electronKind;   /* 0 emcElectron (ParticleType/emcElectron.hpp), 1 electronVWD (examples/mosfet2D/electronVWD.hpp) */
  int32_t surfaceKind[6]; /* per face XMIN..ZMAX: ORC_SURFACE_* (emcScatterHandler.hpp:172-191) */
  double surfaceParam[6]; /* specularity parameter (constant) / rms roughness height [m] (momentum dependent) */
} orc_device_t;

/* PMSchemes/emcNGPScheme.hpp, emcCICScheme.hpp, emcNECScheme.hpp, examples/mosfet2D/NECSchemeVWD.hpp */
enum { ORC_PM_NGP = 0, ORC_PM_CIC = 1, ORC_PM_NEC = 2, ORC_PM_NEC_VWD = 3 };
/* SurfaceScatterMechanisms/emcConstantSurfaceScatterMechanism.hpp, emcMomentumDependentSurfaceScatterMechanism.hpp */
enum { ORC_SURFACE_SPECULAR = 0, ORC_SURFACE_CONSTANT = 1, ORC_SURFACE_MOMENTUM = 2 };

int64_t orc_dev_cells(const orc_device_t *d);
int orc_dev_is_ohmic(const orc_device_t *d, int64_t cell);     /* emcSurface.hpp isOhmicContact */
int orc_dev_is_reservoir(const orc_device_t *d, int64_t cell); /* isReservoirContact */
int orc_dev_contact_idx(const orc_device_t *d, int64_t cell);  /* getOhmicContactIdx */

/* emcSimulationResults.hpp:208-213 */
void orc_initial_potential(const orc_device_t *d, double *pot);
/* emcSORSolver.hpp:49-128 (conc == NULL) / :131-197; returns the number of sweeps */
int orc_sor(const orc_device_t *d, double *pot, const double *conc, double accuracyVolt, double omega, int resetBC,
            int maxSweeps);
/* calcEField of the device's PM scheme; e[dim][cells].  NGP / CIC: calcEFieldAtGridPts, NEC: calcEFieldAtEdgeMidPts
 * (emcEFieldCalculation.hpp:13-82); NEC-VWD: mosfet2D/NECSchemeVWD.hpp:82-99 */
void orc_efield(const orc_device_t *d, const double *pot, double *e);
/* assignToMesh of the device's PM scheme (adds to count): emcNGPScheme.hpp:36-47, emcCICScheme.hpp:71-118,
 * emcNECScheme.hpp:62-96, mosfet2D/NECSchemeVWD.hpp:39-52 */
int orc_assign(const orc_device_t *d, int64_t n, const double *x, const double *y, const double *z, double nrCarriers,
               double *count);
/* the same for ORC_PM_NGP only (kept for older callers) */
int orc_ngp_assign(const orc_device_t *d, int64_t n, const double *x, const double *y, const double *z,
                   double nrCarriers, double *count);
/* emcSimulationResults.hpp:98-116 */
void orc_concentration(const orc_device_t *d, const double *count, double *conc);
/* emcAbstractParticleHandler.hpp:263-277 + emcElectron.hpp:63-73 */
void orc_expected_at_contact(const orc_device_t *d, double *expected);
/* particles per cell at start / expected per contact cell: emcElectron.hpp:48-73 (pot == NULL: density from the
 * doping) or electronVWD.hpp:40-62 (rounded, always from the potential) */
double orc_initial_nr_particles(const orc_device_t *d, int64_t cell, const double *pot);
/* emcAbstractParticleHandler.hpp:133-148; pot == NULL: density from the doping (usePotentialForInit = false) */
int64_t orc_device_generate_initial_pot(const orc_model_t *m, const orc_device_t *d, double nrCarriers, uint64_t *mtState,
                                        const double *pot, orc_ensemble_t *out, int64_t capacity);
int64_t orc_device_generate_initial(const orc_model_t *m, const orc_device_t *d, double nrCarriers, uint64_t *mtState,
                                    orc_ensemble_t *out, int64_t capacity);
/* emcBasicParticleHandler.hpp:76-145 for one step: per-particle removed flags and per-contact counts; the
 * ensemble is NOT compacted (orc_compact does that, keeping the order like removeParticles :267-277). */
int orc_device_step(const orc_model_t *m, const orc_device_t *d, orc_ensemble_t *ens, const double *e, double charge,
                    double dt, int64_t stepIndex, const orc_rng_cfg_t *rng, int8_t *removed, int32_t *removedPerContact,
                    int32_t *recPid, int64_t recCap, int64_t *recCount, int64_t *events, int64_t evCap, int64_t *evCount);
int64_t orc_compact(orc_ensemble_t *ens, const int8_t *removed);
/* emcBasicParticleHandler.hpp:158-192: delete excess reservoir particles in index order, inject the missing
 * ones (cells in storage order, draws from the global mt19937_64); net[contact] = injected - deleted */
int64_t orc_contacts(const orc_model_t *m, const orc_device_t *d, orc_ensemble_t *ens, int64_t capacity,
                     const double *expected, double nrCarriers, uint64_t *mtState, int32_t *net);

#ifdef __cplusplus
}
#endif
#endif
