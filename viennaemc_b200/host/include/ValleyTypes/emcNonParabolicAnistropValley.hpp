// Ellipsoidal valley with Kane non-parabolicity (the silicon X valleys).
// Interface mirrored: reference include/ValleyTypes/emcNonParabolicAnistropValley.hpp
// (the file name carries the reference's spelling so that user includes keep working).
#ifndef EMC_NONPARABOLIC_ANISOTROP_VALLEY_HPP
#define EMC_NONPARABOLIC_ANISOTROP_VALLEY_HPP

#include <detail/emcEllipsoidalValley.hpp>

template <class T> class emcNonParabolicAnisotropValley : public emcdetail::EllipsoidalValley<T, true, true> {
public:
  emcNonParabolicAnisotropValley() = delete;
  emcNonParabolicAnisotropValley(std::array<T, 3> inRelEffMass, T inParticleMass, SizeType inDegFactor, T inAlpha,
                                 T inBottomEnergy = 0.)
      : emcdetail::EllipsoidalValley<T, true, true>(inRelEffMass, inParticleMass, inDegFactor, inAlpha,
                                                    inBottomEnergy) {}
};

#endif
