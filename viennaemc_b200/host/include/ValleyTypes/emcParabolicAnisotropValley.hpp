// Ellipsoidal parabolic valley (relative masses along the three ellipse axes).
// Interface mirrored: reference include/ValleyTypes/emcParabolicAnisotropValley.hpp.
#ifndef EMC_PARABOLIC_ANISOTROP_VALLEY_HPP
#define EMC_PARABOLIC_ANISOTROP_VALLEY_HPP

#include <detail/emcEllipsoidalValley.hpp>

template <class T> class emcParabolicAnisotropValley : public emcdetail::EllipsoidalValley<T, true, false> {
public:
  emcParabolicAnisotropValley() = delete;
  emcParabolicAnisotropValley(std::array<T, 3> inRelEffMass, T inParticleMass, SizeType inDegFactor,
                              T inBottomEnergy = 0.)
      : emcdetail::EllipsoidalValley<T, true, false>(inRelEffMass, inParticleMass, inDegFactor, T(0), inBottomEnergy) {}
};

#endif
