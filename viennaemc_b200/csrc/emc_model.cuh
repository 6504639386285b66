// Device-resident physics model of the particle loop: valley constants,
// table-set directory, mechanism descriptors.  Built once on the host from the
// C-ABI structs (include/emcgpu.h) and staged into shared memory by each CTA.
#pragma once
#include <cstdint>

#include "../../include/emcgpu.h"

namespace emc {

// reference: include/emcConstants.hpp:9-27 -- literal values, not CODATA
constexpr double kPi = 3.14159265358979323846;
constexpr double kQ = 1.60219e-19;
constexpr double kKB = 1.38066e-23;
constexpr double kHbar = 1.05459e-34;

enum RotKind : int32_t { ROT_IDENTITY = 0, ROT_SIGNED_PERMUTATION = 1, ROT_GENERAL = 2 };

// One valley, laid out for the kernels.  "x" members serve the EXACT math mode
// (reference operation order), "f" members the FAST mode (hoisted products).
struct DevValley {
  int32_t kind;    // emcgpu_valley_kind
  int32_t deg;
  int32_t rotKind; // RotKind, worst case over the sub-valleys
  int32_t nonParabolic;
  double mCond, alpha, eBottom;
  double mBand;    // mass of the dispersion, of |k|(E) and of the velocity: mCond, except m_DOS in the anisotropic
                   // single-layer class (emcNonParabolicAnisotropSingleLayerValley.hpp:123-148)
  double vogt[3];
  double xMq;      // mBand*q                  (denominator of getEnergy, non-parabolic)
  double xTwoMq;   // (2*mBand)*q              (denominator of getEnergy, parabolic)
  double fE;       // hbar^2/(mBand q) or hbar^2/(2 mBand q)
  double fPos[3];  // hbar*vogt/(2 mCond)
  double fVel[3];  // hbar*vogt/mBand
  double fDk[3];   // vogt/hbar
  // signed permutations, 4 bits per component: source index | sign bit << 2
  uint16_t permToE[EMCGPU_MAX_SUBVALLEYS];
  uint16_t permToD[EMCGPU_MAX_SUBVALLEYS];
  double rot[EMCGPU_MAX_SUBVALLEYS][9];
};

struct DevMech {
  int32_t sampler, finalValley, nFinal, mechId;
  double param[3]; // [0]: energy change / Debye energy; [1], [2]: sampler specific (include/emcgpu.h)
  int32_t bath;  // phonon bath of a polar-optical mechanism or -1
  int32_t flags; // bit 0: polar angle through the bath's |q| distribution
  uint8_t finalSub[EMCGPU_MAX_SUBVALLEYS][EMCGPU_MAX_FINAL];
};

struct DevTableSet {
  int32_t nMech;
  int32_t stride;    // doubles per energy level row (nMech rounded up to even)
  int32_t tabOffset; // offset (in doubles) of this set's [nLevels][stride] block
  int32_t mechOffset;
  double tau;
};

constexpr int kMaxRegions = 16;

// phonon baths as the kernels see them (emcgpu_set_phonon_baths)
struct BathView {
  unsigned long long *counts; // [nBaths][2][nBins]: emission, absorption
  const double *cumW, *cumWN; // [nBaths][nBins + 1] or nullptr
  int32_t nBaths, nBins;
  double dq;
};

// Everything the kernels need besides the tables themselves; one copy in
// global memory, staged to shared memory per CTA (a few KB).
struct DevModel {
  int32_t nValleys, nSets, nLevels, nRegions;
  double dE;          // maxEnergy / nLevels  (emcScatterHandler.hpp:76)
  double defaultTau;  // 2e-15                (emcScatterHandler.hpp:83)
  int64_t tableDoubles;  // total size of the table block
  int8_t setOf[EMCGPU_MAX_VALLEYS][kMaxRegions]; // (valley, region) -> set or -1
  DevTableSet sets[EMCGPU_MAX_TABLESETS];
  DevValley valleys[EMCGPU_MAX_VALLEYS];
};

} // namespace emc
