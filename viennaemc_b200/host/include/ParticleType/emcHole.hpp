// Valence-band holes.  Interface mirrored: reference include/ParticleType/emcHole.hpp (ctor :38-44); the body is
// detail/emcBandCarrier.hpp with charge +q.
#ifndef EMC_HOLE_HPP
#define EMC_HOLE_HPP

#include <detail/emcBandCarrier.hpp>

template <class T, class DeviceType> struct emcHole : public emcdetail::BandCarrier<T, DeviceType, +1> {
  emcHole(SizeType inHandlerNrEnergyLevels = 1000, T inHandlerMaxEnergy = 4., bool inUsePotentialForInit = false,
          T inInitEnergyEV = T(0))
      : emcdetail::BandCarrier<T, DeviceType, +1>(inHandlerNrEnergyLevels, inHandlerMaxEnergy, inUsePotentialForInit,
                                                  inInitEnergyEV) {}
};

#endif
