// Rough wall whose diffusive fraction grows with the momentum normal to it: p = 1 - exp(-(2 h k_perp)^2) with the rms
// roughness height h; the polar angle of a diffusive event follows the matching cumulative distribution (Newton
// iteration on the device).
// Interface mirrored: reference include/SurfaceScatterMechanisms/emcMomentumDependentSurfaceScatterMechanism.hpp.
#ifndef EMC_MOMENTUM_DEPENDENT_SURFACE_SCATTER_MECHANISM_HPP
#define EMC_MOMENTUM_DEPENDENT_SURFACE_SCATTER_MECHANISM_HPP

#include <cmath>

#include <SurfaceScatterMechanisms/emcSurfaceScatterMechanism.hpp>

template <class T, class DeviceType, SizeType Dim = DeviceType::Dimension>
class emcMomentumDependentSurfaceScatterMechanism : public emcSurfaceScatterMechanism<T, DeviceType> {
  T roughnessHeight;

public:
  emcMomentumDependentSurfaceScatterMechanism(T inRoughnessHeight, std::array<T, Dim> inMaxPos)
      : emcSurfaceScatterMechanism<T, DeviceType>(inMaxPos), roughnessHeight(inRoughnessHeight) {}

  T getDiffScatterProb(emcParticle<T> &particle) const override {
    return 1 - (std::exp(-std::pow(2 * roughnessHeight * particle.k[this->getIndexFromBoundaryPos()], 2)));
  }
  int deviceSurfaceKind() const override { return EMCGPU_SURFACE_MOMENTUM_DEPENDENT; }
  T deviceSurfaceParameter() const override { return roughnessHeight; }
};

#endif
