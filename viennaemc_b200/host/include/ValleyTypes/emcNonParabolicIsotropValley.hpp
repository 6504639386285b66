// Spherical valley with Kane non-parabolicity alpha [1/eV].  Interface mirrored:
// reference include/ValleyTypes/emcNonParabolicIsotropValley.hpp.
#ifndef EMC_NONPARABOLIC_ISOTROP_VALLEY_HPP
#define EMC_NONPARABOLIC_ISOTROP_VALLEY_HPP

#include <detail/emcEllipsoidalValley.hpp>

template <class T> class emcNonParabolicIsotropValley : public emcdetail::EllipsoidalValley<T, false, true> {
public:
  emcNonParabolicIsotropValley() = delete;
  emcNonParabolicIsotropValley(T inRelEffMass, T inParticleMass, SizeType inDegFactor, T inAlpha,
                               T inBottomEnergy = 0.)
      : emcdetail::EllipsoidalValley<T, false, true>({inRelEffMass, inRelEffMass, inRelEffMass}, inParticleMass,
                                                     inDegFactor, inAlpha, inBottomEnergy) {}
};

#endif
