// Remote surface-optical phonon scattering.  Name mirrored: reference
// include/ScatterMechanisms/emcRemoteSurfaceOpticalPhononMechanism.hpp.
// NAME ONLY (detail/emcNoDeviceSampler.hpp): constructible with the reference's arguments, rejected with its name when added
// to a particle type -- no device final-state sampler yet, and nothing is ever scattered on the CPU.
#ifndef EMC_REMOTE_SURFACE_OPTICAL_PHONON_MECHANISM_HPP
#define EMC_REMOTE_SURFACE_OPTICAL_PHONON_MECHANISM_HPP

#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <detail/emcNoDeviceSampler.hpp>

template <class T> struct emcRemoteSurfaceOpticalPhononMechanism : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcRemoteSurfaceOpticalPhononMechanism(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("RemoteSO", inValley) {}
};

#endif
