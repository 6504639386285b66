// First-order intervalley phonon scattering in a single layer.  Names mirrored: reference
// include/ScatterMechanisms/emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp.
// NAME ONLY (detail/emcNoDeviceSampler.hpp): constructible with the reference's arguments, rejected with its name when added
// to a particle type -- no device final-state sampler yet, and nothing is ever scattered on the CPU.
#ifndef EMC_FIRST_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP
#define EMC_FIRST_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP

#include <string>

#include <detail/emcNoDeviceSampler.hpp>

template <class T> struct emcFirstOrderSingleLayerInterValleyAbsorptionScatterMechanism : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcFirstOrderSingleLayerInterValleyAbsorptionScatterMechanism(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("FirstInterValleyAbsorptionSL", inValley) {}
};

template <class T> struct emcFirstOrderSingleLayerInterValleyEmissionScatterMechanism : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcFirstOrderSingleLayerInterValleyEmissionScatterMechanism(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("FirstInterValleyEmissionSL", inValley) {}
};

#endif
