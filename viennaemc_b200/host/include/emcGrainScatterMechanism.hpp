// Grain-boundary scattering (second exponential clock).  Interface mirrored: reference
// include/emcGrainScatterMechanism.hpp (ctor: transmission probability, scatter rate [1/s];
// getScatterRate()).  No device sampler exists yet: a particle type that carries one is rejected
// when it is handed to a GPU particle handler (it is off in every example configuration).
#ifndef EMC_GRAIN_SCATTER_MECHANISM_HPP
#define EMC_GRAIN_SCATTER_MECHANISM_HPP

#include <emcUtil.hpp>

template <class T> class emcGrainScatterMechanism {
  T transmissionProbability;
  T scatterRate;

public:
  emcGrainScatterMechanism() = delete;
  emcGrainScatterMechanism(T inTransmissionProbability, T inScatterRate)
      : transmissionProbability(inTransmissionProbability), scatterRate(inScatterRate) {}
  T getScatterRate() const { return scatterRate; }
  T getTransmissionProbability() const { return transmissionProbability; }
};

#endif
