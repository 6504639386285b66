/*
 * emchost.h -- C entry points of the host-side model builder (libemchost.so).
 *
 * The library is the drop-in C++17 host API (viennaemc_b200/host/include, the
 * reference-compatible emcDevice / emcElectron / emcScatterMechanism classes)
 * instantiated for the silicon model of the reference examples, for callers
 * that are not C++ (bench.py, the Python tests).  It builds the rate tables on
 * the host exactly as the reference does (emcScatterHandler::initScatterTables,
 * include/emcScatterHandler.hpp:91-95) and hands them to the CUDA library through
 * the C ABI of emcgpu.h.  Nothing here moves particles.
 */
#ifndef EMCHOST_H
#define EMCHOST_H

#include <stdint.h>

#include "emcgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { EMCHOST_ACOUSTIC = 1, EMCHOST_ZERO_ORDER = 2, EMCHOST_FIRST_ORDER = 4, EMCHOST_COULOMB = 8 };

/* Silicon electrons in a box with one constant doping region (region 0), the set-up of
 * examples/bulkSimulation/bulkSimulation.cpp:87-103 with its constants as parameters. */
typedef struct {
  int32_t nLevels;       /* energy levels of the rate tables (1000) */
  int32_t coulombSecond; /* order of the Coulomb mechanism: 1 = right after Acoustic (mosfet2D), 0 = last (resistor2D) */
  uint32_t mechanisms;   /* EMCHOST_* mask; shipped bulk example: ACOUSTIC | ZERO_ORDER | FIRST_ORDER */
  uint32_t reserved;
  double maxEnergy;   /* [eV] (1.0) */
  double temperature; /* [K] (300) */
  double doping;      /* [1/m^3] (1e23) */
  double box[3];      /* [m] */
  double spacing[3];  /* [m]; only used to create the initial ensemble cell by cell */
} emchost_si_spec;

/* emcgpu_set_valleys + emcgpu_set_tables for that model.  Returns an emcgpu_status. */
int emchost_si_upload(emcgpu_ctx *ctx, const emchost_si_spec *spec);

/* CPU only: the normalised cumulative tables [nMech][nLevels] and tau of (valley 0, region 0).
 * cum may be NULL to query nMech. */
int emchost_si_tables(const emchost_si_spec *spec, double *cum, int64_t cumCapacity, double *tau, int32_t *nMech);

/* CPU only: the initial ensemble the bulk handler would create for a seed
 * (basicBulkParticleHandler::generateInitialParticles).  soa[EMCGPU_N_STREAMS], packed and
 * grainTau have room for `capacity` particles (any may be NULL to only count).  Returns the
 * number of particles, or -1 if capacity is too small. */
int64_t emchost_si_initial_ensemble(const emchost_si_spec *spec, uint64_t seed, int64_t capacity, double *const *soa,
                                    uint32_t *packed, double *grainTau);

/* ---- beta-Ga2O3 with polar-optical scattering and phonon baths (config 5) -------------------------------------
 * The set-up of examples/hotPhononGa2O3 (Ga2O3Functions.hpp parameters; one region, Gamma valley, acoustic + non-polar
 * optical [+ Brooks-Herring] + polar optical) built with the drop-in host classes emcPhononBath, emcPlasmonScreening and
 * the eight Froehlich mechanism classes. */
typedef struct {
  int32_t polar;      /* 0 emcFroehlich*3D, 1 emcHotPhononFroehlich*3D, 2 emcScreenedFroehlich*3D, 3 emcScreenedHotPhononFroehlich*3D */
  int32_t multimode, screening, qResolved, qResolvedAngle, acousticBath, impurity, nLevels;
  double maxEnergy, temperature, doping, box, tauLO, tauAc;
} emchost_ga2o3_spec;

/* CPU only: drives the HOST side of the hot-phonon loop (hotPhononGa2O3.cpp:270-282) with given inputs -- per step the
 * event counts per bath and |q| bin (counts: [nSteps][nBaths][2][nBins], emission then absorption; may be NULL without
 * baths) and the mean carrier energy (meanEnergy: [nSteps], used by the screening update) -- and returns what the
 * reference would have: cumulative tables before the first step and after the last one ([nMech][nLevels] each), tau
 * after every table rebuild, <N_q> of every bath after every update ([nSteps][nBaths]) and the final occupations
 * ([nBaths][nBins]).  Any output may be NULL.  Returns the number of mechanisms or a negative emcgpu_status. */
int emchost_ga2o3_host_loop(const emchost_ga2o3_spec *spec, int nSteps, double dt, const double *counts, const double *meanEnergy,
                            double *cumInitial, double *cumFinal, double *tauSeries, double *meanNq, double *finalNq);
/* emcgpu_set_valleys + emcgpu_set_phonon_baths + emcgpu_set_tables for that model (initial state) */
int emchost_ga2o3_upload(emcgpu_ctx *ctx, const emchost_ga2o3_spec *spec);

#ifdef __cplusplus
}
#endif
#endif
