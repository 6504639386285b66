// Grain-boundary scattering (second exponential clock).  Interface mirrored: reference
// include/emcGrainScatterMechanism.hpp (ctor: transmission probability, scatter rate [1/s];
// getScatterRate()).  The event itself -- reflection into the opposite or transmission into the same hemisphere about
// the current k (:40-77) -- and the clock run on the device (grainEvent, viennaemc_b200/csrc/emc_bulk_kernel.cuh); the
// GPU particle handlers hand transmission probability and rate over with emcgpu_set_grain.
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
