// Conduction-band electrons.  Interface mirrored: reference include/ParticleType/emcElectron.hpp (ctor :28-35); the body is
// detail/emcBandCarrier.hpp with charge -q.
#ifndef EMC_ELECTRON_HPP
#define EMC_ELECTRON_HPP

#include <detail/emcBandCarrier.hpp>

template <class T, class DeviceType> struct emcElectron : public emcdetail::BandCarrier<T, DeviceType, -1> {
  emcElectron(SizeType inHandlerNrEnergyLevels = 1000, T inHandlerMaxEnergy = 4., bool inUsePotentialForInit = true,
              T inInitEnergyEV = T(0))
      : emcdetail::BandCarrier<T, DeviceType, -1>(inHandlerNrEnergyLevels, inHandlerMaxEnergy, inUsePotentialForInit,
                                                  inInitEnergyEV) {}
};

#endif
