// Initial / injected particle state.
// Interface mirrored: reference include/emcParticleInitialization.hpp -- initParticlePos
// :14-29 (uniform inside the grid cell of a coordinate; half cells at the box faces),
// initParticleKSpaceMaxwellian :36-51 (E = -1.5 Vt ln U[1e-6,1), isotropic direction,
// k turned inward in boundary cells), initParticleKSpaceFixed.
// The draw order is part of the contract: a seeded run creates the same ensemble
// as the reference.
#ifndef EMC_PARTICLE_INITIALIZATION_HPP
#define EMC_PARTICLE_INITIALIZATION_HPP

#include <array>
#include <cmath>
#include <random>

#include <emcConstants.hpp>
#include <emcParticle.hpp>
#include <emcUtil.hpp>

template <class T, SizeType Dim>
std::array<T, Dim> initParticlePos(const std::array<SizeType, Dim> &coord, const std::array<SizeType, Dim> &extent,
                                   const std::array<T, Dim> &spacing, emcRNG &rng) {
  std::uniform_real_distribution<T> uniform(0., 1.);
  std::array<T, Dim> pos;
  for (SizeType d = 0; d < Dim; d++) {
    const T u = uniform(rng);
    if (coord[d] == extent[d] - 1)
      pos[d] = (coord[d] - u * 0.5) * spacing[d];
    else if (coord[d] == 0)
      pos[d] = u * 0.5 * spacing[d];
    else
      pos[d] = (coord[d] + u - 0.5) * spacing[d];
  }
  return pos;
}

namespace emcdetail {
// k must point into the box in cells that touch a face
template <class T, SizeType Dim>
void turnInward(emcParticle<T> &part, const std::array<SizeType, Dim> &coord, const std::array<SizeType, Dim> &extent) {
  for (SizeType d = 0; d < Dim; d++)
    if ((coord[d] == 0 && part.k[d] < 0) || (coord[d] == extent[d] - 1 && part.k[d] > 0))
      part.k[d] *= -1;
}
template <class T, class ValleyType> void isotropicK(emcParticle<T> &part, const ValleyType *valley, emcRNG &rng) {
  std::uniform_real_distribution<T> uniform(0., 1.);
  const T cosDraw = uniform(rng); // drawn first (g++ evaluates call arguments right to left)
  const T phiDraw = uniform(rng);
  part.k = initRandomDirection(valley->getNormWaveVec(part.energy), phiDraw, cosDraw);
}
} // namespace emcdetail

template <class T, SizeType Dim, template <class, SizeType> class DeviceType, class ValleyType>
void initParticleKSpaceMaxwellian(emcParticle<T> &part, const std::array<SizeType, Dim> &coord,
                                  const DeviceType<T, Dim> &device, const ValleyType *valley, emcRNG &rng) {
  std::uniform_real_distribution<T> forLog(1e-6, 1.);
  part.energy = -1.5 * device.getThermalVoltage() * std::log(forLog(rng));
  emcdetail::isotropicK(part, valley, rng);
  emcdetail::turnInward<T, Dim>(part, coord, device.getGridExtent());
}

template <class T, SizeType Dim, template <class, SizeType> class DeviceType, class ValleyType>
void initParticleKSpaceFixed(emcParticle<T> &part, T initEnergyEV, const std::array<SizeType, Dim> &coord,
                             const DeviceType<T, Dim> &device, const ValleyType *valley, emcRNG &rng) {
  part.energy = initEnergyEV;
  emcdetail::isotropicK(part, valley, rng);
  emcdetail::turnInward<T, Dim>(part, coord, device.getGridExtent());
}

#endif
