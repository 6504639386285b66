// Electron-hole binary collisions (reference include/emcInterCarrierScatter.hpp, ctor :87-100): pairwise and sequential
// like emcCarrierCarrierScatter; present for source compatibility, rejected by the GPU bulk handler.
#ifndef EMC_INTER_CARRIER_SCATTER_HPP
#define EMC_INTER_CARRIER_SCATTER_HPP

#include <emcUtil.hpp>

template <class T> class emcInterCarrierScatter {
public:
  T epsR, relEffMassE, relEffMassH, Vsim, latTempK;
  emcInterCarrierScatter() = delete;
  emcInterCarrierScatter(T inEpsR, T inRelEffMassE, T inRelEffMassH, T inVsim, T inTempK)
      : epsR(inEpsR), relEffMassE(inRelEffMassE), relEffMassH(inRelEffMassH), Vsim(inVsim), latTempK(inTempK) {}
  static const char *name() { return "emcInterCarrierScatter"; }
};

#endif
