// Parabolic isotropic valley of a single layer (2-D extension in x and y).  Interface mirrored: reference
// include/ValleyTypes/emcParabolicIsotropSingleLayerValley.hpp.
#ifndef EMC_PARABOLIC_ISOTROP_SINGLELAYER_VALLEY_HPP
#define EMC_PARABOLIC_ISOTROP_SINGLELAYER_VALLEY_HPP

#include <detail/emcSingleLayerValley.hpp>

template <class T> class emcParabolicIsotropSingleLayerValley : public emcdetail::SingleLayerValley<T, false, false> {
public:
  emcParabolicIsotropSingleLayerValley() = delete;
  emcParabolicIsotropSingleLayerValley(T inRelEffMass, T inParticleMass, SizeType inDegFactor, T inBottomValleyEnergy = 0.)
      : emcdetail::SingleLayerValley<T, false, false>(inRelEffMass, inParticleMass, inDegFactor, T(0), inBottomValleyEnergy) {}
};

#endif
