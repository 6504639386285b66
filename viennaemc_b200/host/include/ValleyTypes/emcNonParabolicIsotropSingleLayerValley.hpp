// Isotropic single-layer valley with Kane non-parabolicity alpha [1/eV].  Interface mirrored: reference
// include/ValleyTypes/emcNonParabolicIsotropSingleLayerValley.hpp.
#ifndef EMC_NONPARABOLIC_ISOTROP_SINGLELAYER_VALLEY_HPP
#define EMC_NONPARABOLIC_ISOTROP_SINGLELAYER_VALLEY_HPP

#include <detail/emcSingleLayerValley.hpp>

template <class T> class emcNonParabolicIsotropSingleLayerValley : public emcdetail::SingleLayerValley<T, false, true> {
public:
  emcNonParabolicIsotropSingleLayerValley() = delete;
  emcNonParabolicIsotropSingleLayerValley(T inRelEffMass, T inParticleMass, SizeType inDegFactor, T inAlpha,
                                          T inBottomValleyEnergy = 0.)
      : emcdetail::SingleLayerValley<T, false, true>(inRelEffMass, inParticleMass, inDegFactor, inAlpha,
                                                     inBottomValleyEnergy) {}
};

#endif
