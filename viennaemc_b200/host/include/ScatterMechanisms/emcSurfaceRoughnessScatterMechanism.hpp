// Surface-roughness scattering of a 2-D layer.  Name mirrored: reference
// include/ScatterMechanisms/emcSurfaceRoughnessScatterMechanism.hpp.
// NAME ONLY (detail/emcNoDeviceSampler.hpp): constructible with the reference's arguments, rejected with its name when added
// to a particle type -- no device final-state sampler yet, and nothing is ever scattered on the CPU.
#ifndef EMC_SURFACE_ROUGHNESS_SCATTER_MECHANISM_HPP
#define EMC_SURFACE_ROUGHNESS_SCATTER_MECHANISM_HPP

#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <detail/emcNoDeviceSampler.hpp>

template <class T> struct emcSurfaceRoughnessScatterMechanism : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcSurfaceRoughnessScatterMechanism(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("SurfaceRoughness", inValley) {}
};

#endif
