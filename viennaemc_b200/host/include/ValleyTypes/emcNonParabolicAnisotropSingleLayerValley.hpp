// Anisotropic single-layer valley (longitudinal / transversal mass, one in-plane rotation angle per sub-valley) with Kane
// non-parabolicity.  Interface mirrored: reference include/ValleyTypes/emcNonParabolicAnisotropSingleLayerValley.hpp.
#ifndef EMC_NONPARABOLIC_ANISOTROP_SINGLELAYER_VALLEY_HPP
#define EMC_NONPARABOLIC_ANISOTROP_SINGLELAYER_VALLEY_HPP

#include <detail/emcSingleLayerValley.hpp>

template <class T> class emcNonParabolicAnisotropSingleLayerValley : public emcdetail::SingleLayerValley<T, true, true> {
public:
  emcNonParabolicAnisotropSingleLayerValley() = delete;
  // all sub-valleys aligned with the x axis
  emcNonParabolicAnisotropSingleLayerValley(T relEffMassLongitudinal, T relEffMassTransversal, T inParticleMass,
                                            SizeType inDegFactor, T inAlpha, T inBottomEnergy = 0.)
      : emcNonParabolicAnisotropSingleLayerValley(relEffMassLongitudinal, relEffMassTransversal, inParticleMass, inDegFactor,
                                                  inAlpha, std::vector<T>(inDegFactor, 0.), inBottomEnergy) {}
  emcNonParabolicAnisotropSingleLayerValley(T relEffMassLongitudinal, T relEffMassTransversal, T inParticleMass,
                                            SizeType inDegFactor, T inAlpha, std::vector<T> inRotationAngles,
                                            T inBottomEnergy = 0.)
      : emcdetail::SingleLayerValley<T, true, true>(relEffMassLongitudinal, relEffMassTransversal, inParticleMass, inDegFactor,
                                                    inAlpha, std::move(inRotationAngles), inBottomEnergy) {}
};

#endif
