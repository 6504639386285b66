// Spherical parabolic valley.  Interface mirrored: reference
// include/ValleyTypes/emcParabolicIsotropValley.hpp (ctor: relative effective mass,
// particle rest mass, degeneracy, bottom energy [eV]).
#ifndef EMC_PARABOLIC_ISOTROP_VALLEY_HPP
#define EMC_PARABOLIC_ISOTROP_VALLEY_HPP

#include <detail/emcEllipsoidalValley.hpp>

template <class T> class emcParabolicIsotropValley : public emcdetail::EllipsoidalValley<T, false, false> {
public:
  emcParabolicIsotropValley() = delete;
  emcParabolicIsotropValley(T inRelEffMass, T inParticleMass, SizeType inDegFactor, T inBottomEnergy = 0.)
      : emcdetail::EllipsoidalValley<T, false, false>({inRelEffMass, inRelEffMass, inRelEffMass}, inParticleMass,
                                                      inDegFactor, T(0), inBottomEnergy) {}
};

#endif
