// Wall with a constant specularity: a fraction (1 - specularity) of the hits is diffusive, with the polar angle of the
// new direction theta = asin(sqrt(r)) (cosine law) and a uniform azimuth.
// Interface mirrored: reference include/SurfaceScatterMechanisms/emcConstantSurfaceScatterMechanism.hpp.
#ifndef EMC_CONSTANT_SURFACE_SCATTER_MECHANISM_HPP
#define EMC_CONSTANT_SURFACE_SCATTER_MECHANISM_HPP

#include <SurfaceScatterMechanisms/emcSurfaceScatterMechanism.hpp>

template <class T, class DeviceType, SizeType Dim = DeviceType::Dimension>
class emcConstantSurfaceScatterMechanism : public emcSurfaceScatterMechanism<T, DeviceType> {
  T specularityParam;

public:
  emcConstantSurfaceScatterMechanism() = delete;
  emcConstantSurfaceScatterMechanism(T inSpecularityParam, std::array<T, Dim> inMaxPos)
      : emcSurfaceScatterMechanism<T, DeviceType>(inMaxPos), specularityParam(inSpecularityParam) {}

  T getDiffScatterProb(emcParticle<T> & /*particle*/) const override { return 1 - specularityParam; }
  int deviceSurfaceKind() const override { return EMCGPU_SURFACE_CONSTANT; }
  T deviceSurfaceParameter() const override { return specularityParam; }
};

#endif
