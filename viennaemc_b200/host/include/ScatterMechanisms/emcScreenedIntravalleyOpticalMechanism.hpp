// Free-carrier-screened intravalley optical scattering of a 2-D layer.  Name mirrored: reference
// include/ScatterMechanisms/emcScreenedIntravalleyOpticalMechanism.hpp.
// NAME ONLY (detail/emcNoDeviceSampler.hpp): constructible with the reference's arguments, rejected with its name when added
// to a particle type -- no device final-state sampler yet, and nothing is ever scattered on the CPU.
#ifndef EMC_SCREENED_INTRAVALLEY_OPTICAL_MECHANISM_HPP
#define EMC_SCREENED_INTRAVALLEY_OPTICAL_MECHANISM_HPP

#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <detail/emcNoDeviceSampler.hpp>

template <class T> struct emcScreenedIntravalleyOpticalMechanism : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcScreenedIntravalleyOpticalMechanism(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("ScreenedIntraOptical", inValley) {}
};

#endif
