// Radiative and Auger band-to-band recombination (reference include/emcRecombination.hpp, ctor :107-111, counters
// :206-207): removes electron-hole pairs from the two host ensembles with Poisson-distributed event counts per step.
// It changes the ensemble SIZE from the host side and couples two species; present for source compatibility, rejected by
// the GPU bulk handler (basicBulkParticleHandler::recombine).
#ifndef EMC_RECOMBINATION_HPP
#define EMC_RECOMBINATION_HPP

#include <emcUtil.hpp>

template <class T> class emcRecombination {
  SizeType nRadiative = 0, nAuger = 0;

public:
  T B, C_n, C_p, E_gap, relEffMassE, relEffMassH, Vsim;
  emcRecombination() = delete;
  emcRecombination(T inB, T inCn, T inCp, T inGap, T inRelEffMassE, T inRelEffMassH, T inVsim)
      : B(inB), C_n(inCn), C_p(inCp), E_gap(inGap), relEffMassE(inRelEffMassE), relEffMassH(inRelEffMassH), Vsim(inVsim) {}
  SizeType getNrRadiative() const { return nRadiative; }
  SizeType getNrAuger() const { return nAuger; }
  static const char *name() { return "emcRecombination"; }
};

#endif
