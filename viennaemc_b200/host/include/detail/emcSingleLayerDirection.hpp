// The final direction shared by the elastic and the intervalley single-layer mechanisms: an in-plane angle 2 pi u, weighted
// by the Herring-Vogt factors of the (final) valley, scaled to |k|(E); k_z = 0.
// Arithmetic mirrored: reference include/ScatterMechanisms/emcAcousticSingleLayerScatterMechanism.hpp:63-81,
// emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp:131-147.  Device: EMCGPU_SAMPLER_SINGLE_LAYER_*.
#ifndef EMC_DETAIL_SINGLE_LAYER_DIRECTION_HPP
#define EMC_DETAIL_SINGLE_LAYER_DIRECTION_HPP

#include <cmath>

#include <ValleyTypes/emcAbstractValley.hpp>
#include <emcConstants.hpp>

namespace emcdetail {

template <class T> std::array<T, 3> singleLayerDirection(const emcAbstractValley<T> *valley, T energy, T uniformDraw) {
  const T phi = 2 * constants::pi * uniformDraw;
  const auto &vogt = valley->getVogtTransformationFactor();
  std::array<T, 3> k{std::cos(phi) / vogt[0], std::sin(phi) / vogt[1], 0};
  const T toUnit = 1. / (std::sqrt(k[0] * k[0] + k[1] * k[1]));
  k[0] *= toUnit;
  k[1] *= toUnit;
  const T kNorm = valley->getNormWaveVec(energy);
  k[0] *= kNorm;
  k[1] *= kNorm;
  return k;
}

} // namespace emcdetail

#endif
