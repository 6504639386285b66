// Energy-selective contact: removes carriers whose energy lies in a window around E_ex with probability dt / tau_ex per
// step (reference include/emcEnergySelectiveContact.hpp, ctor :84-88, accessors :114-143).  It edits the host ensemble
// particle by particle; present for source compatibility, rejected by the GPU bulk handler
// (basicBulkParticleHandler::extractCarriers).  The counters stay at zero.
#ifndef EMC_ENERGY_SELECTIVE_CONTACT_HPP
#define EMC_ENERGY_SELECTIVE_CONTACT_HPP

#include <emcConstants.hpp>
#include <emcUtil.hpp>

template <class T> class emcEnergySelectiveContact {
  T E_ex, halfDeltaE, tauEx;
  SizeType nExtracted = 0;
  T sumEnergy = T(0);

public:
  emcEnergySelectiveContact() = delete;
  emcEnergySelectiveContact(T inExtractionEnergy, T inDeltaE, T inTauEx)
      : E_ex(inExtractionEnergy), halfDeltaE(T(0.5) * inDeltaE), tauEx(inTauEx) {}
  SizeType getNrExtracted() const { return nExtracted; }
  T getMeanExtractedEnergy() const { return nExtracted ? sumEnergy / T(nExtracted) : T(0); }
  // q N / (A t)  [A/m^2]
  T getCurrentDensity(T contactArea, T elapsedTime) const {
    return (contactArea > T(0) && elapsedTime > T(0)) ? T(constants::q) * T(nExtracted) / (contactArea * elapsedTime) : T(0);
  }
  void resetCounters() {
    nExtracted = 0;
    sumEnergy = T(0);
  }
  T getExtractionEnergy() const { return E_ex; }
  T getWindowWidth() const { return T(2) * halfDeltaE; }
  T getExtractionTimeCst() const { return tauEx; }
  static const char *name() { return "emcEnergySelectiveContact"; }
};

#endif
