// Debye-screened binary carrier-carrier collisions inside one species (reference include/emcCarrierCarrierScatter.hpp,
// ctor :72-80).  The reference pairs particles at random and rotates their relative momentum SEQUENTIALLY on the host
// ensemble; that pairwise, order-dependent step is outside the data-parallel particle loop this library accelerates.
// The class exists so that drivers of the hot-carrier example compile; basicBulkParticleHandler::carrierCarrierScatter
// rejects the call with an error that names it (no CPU fallback).
#ifndef EMC_CARRIER_CARRIER_SCATTER_HPP
#define EMC_CARRIER_CARRIER_SCATTER_HPP

#include <emcUtil.hpp>

template <class T> class emcCarrierCarrierScatter {
public:
  T epsR, relEffMass, Vsim, latTempK;
  emcCarrierCarrierScatter() = delete;
  emcCarrierCarrierScatter(T inEpsR, T inRelEffMass, T inVsim, T inTempK)
      : epsR(inEpsR), relEffMass(inRelEffMass), Vsim(inVsim), latTempK(inTempK) {}
  static const char *name() { return "emcCarrierCarrierScatter"; }
};

#endif
