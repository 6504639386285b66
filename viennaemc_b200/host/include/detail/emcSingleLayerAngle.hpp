// Deflection angle of the long-range single-layer mechanisms (Froehlich, piezoelectric): drawn by inversion of a 128-point
// cumulative sum (midpoint rule) of an angular weight on [0, pi]; the side is a second draw.
// Arithmetic mirrored: reference include/ScatterMechanisms/emcFroehlichInteractionSingleLayer.hpp:45-80,
// emcPiezoelectricSingleLayerScatterMechanism.hpp:62-66, :117-134.  Device: slAngularWeight / slInvertAngle (emc_device.cuh).
#ifndef EMC_DETAIL_SINGLE_LAYER_ANGLE_HPP
#define EMC_DETAIL_SINGLE_LAYER_ANGLE_HPP

#include <array>
#include <cmath>
#include <random>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <emcConstants.hpp>
#include <emcUtil.hpp>

namespace emcdetail {

constexpr SizeType singleLayerAngleSteps = 128;

// cumulative sums c[0] = 0, c[i] = c[i-1] + weight((i - 1/2) pi / 128)
template <class T, class Weight> std::array<T, singleLayerAngleSteps + 1> cumulativeAngularWeight(Weight &&weight) {
  std::array<T, singleLayerAngleSteps + 1> sums;
  const T step = constants::pi / singleLayerAngleSteps;
  sums[0] = T(0);
  for (SizeType i = 1; i <= singleLayerAngleSteps; ++i)
    sums[i] = sums[i - 1] + weight((i - T(0.5)) * step);
  return sums;
}

// the angle in [0, pi] at which the cumulative sum reaches `target` (linear inside the bin)
template <class T> T invertAngularWeight(const std::array<T, singleLayerAngleSteps + 1> &sums, T target) {
  SizeType bin = 1;
  while (bin < singleLayerAngleSteps && sums[bin] < target)
    ++bin;
  const T step = constants::pi / singleLayerAngleSteps;
  return (T(bin) - 1 + (target - sums[bin - 1]) / (sums[bin] - sums[bin - 1])) * step;
}

// erfc(w q / 2)^2 / eps(q)^2: finite-thickness form factor of the layer times the 2-D free-carrier screening
template <class T> T formFactorScreened(T q, T width, T screeningWavevector) {
  const T ff = std::erfc(width * q / 2);
  return ff * ff * twoDScreeningFactor(q, screeningWavevector);
}

// ---- the same inversion with N bins, and the whole final state of the mechanisms that turn the in-plane wave vector by a
// weighted angle of either sign (reference emc2DChargedImpurityScatterMechanism.hpp:107-139, emcSurfaceRoughnessScatterMechanism.hpp
// :94-126, emcRemoteSurfaceOpticalPhononMechanism.hpp:112-149, emcScreenedIntravalleyOpticalMechanism.hpp:104-141).
// Device: slAngularScatter (emc_device.cuh).
template <SizeType N, class T, class Weight> T midpointAngularSum(Weight &&weight) { // sum_i weight((i + 1/2) pi / N), i < N
  const T step = constants::pi / N;
  T sum = 0;
  for (SizeType i = 0; i < N; ++i)
    sum += weight((i + T(0.5)) * step);
  return sum;
}

template <SizeType N, class T, class Particle, class Rng, class Weight>
void turnByWeightedAngle(Particle &particle, Rng &rng, std::uniform_real_distribution<T> &uniform, T finalNorm, Weight &&weight) {
  std::array<T, N + 1> sums;
  const T step = constants::pi / N;
  sums[0] = T(0);
  for (SizeType i = 1; i <= N; ++i)
    sums[i] = sums[i - 1] + weight((i - T(0.5)) * step);
  const T total = sums[N];
  const T phi = std::atan2(particle.k[1], particle.k[0]);
  T angle;
  if (!(total > T(0))) {
    angle = constants::pi * uniform(rng); // nothing to weight with: any magnitude
  } else {
    const T target = uniform(rng) * total;
    SizeType bin = 1;
    while (bin < N && sums[bin] < target)
      ++bin;
    angle = (T(bin) - 1 + (target - sums[bin - 1]) / (sums[bin] - sums[bin - 1])) * step;
  }
  if (uniform(rng) < T(0.5))
    angle = -angle;
  particle.k = {finalNorm * std::cos(phi + angle), finalNorm * std::sin(phi + angle), 0};
}

} // namespace emcdetail

#endif
