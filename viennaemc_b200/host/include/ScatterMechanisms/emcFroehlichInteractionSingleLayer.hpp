// Froehlich (polar-optical) interaction in a single layer.  Names mirrored: reference
// include/ScatterMechanisms/emcFroehlichInteractionSingleLayer.hpp.
// NAME ONLY (detail/emcNoDeviceSampler.hpp): constructible with the reference's arguments, rejected with its name when added
// to a particle type -- no device final-state sampler yet, and nothing is ever scattered on the CPU.
#ifndef EMC_FROEHLICH_INTERACTION_SINGLE_LAYER_HPP
#define EMC_FROEHLICH_INTERACTION_SINGLE_LAYER_HPP

#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <detail/emcNoDeviceSampler.hpp>

template <class T> struct emcFroehlichInteractionAbsorptionSL : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcFroehlichInteractionAbsorptionSL(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("froehlichAbsorptionSL", inValley) {}
};

template <class T> struct emcFroehlichInteractionEmissionSL : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcFroehlichInteractionEmissionSL(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("froehlichEmissionSL", inValley) {}
};

#endif
