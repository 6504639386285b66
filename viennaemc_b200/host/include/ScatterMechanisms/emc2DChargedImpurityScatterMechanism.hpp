// Charged-impurity scattering in a 2-D semiconductor.  Name mirrored: reference
// include/ScatterMechanisms/emc2DChargedImpurityScatterMechanism.hpp.
// NAME ONLY (detail/emcNoDeviceSampler.hpp): constructible with the reference's arguments, rejected with its name when added
// to a particle type -- no device final-state sampler yet, and nothing is ever scattered on the CPU.
#ifndef EMC_2D_CHARGED_IMPURITY_SCATTER_MECHANISM_HPP
#define EMC_2D_CHARGED_IMPURITY_SCATTER_MECHANISM_HPP

#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <detail/emcNoDeviceSampler.hpp>

template <class T> struct emc2DChargedImpurityScatterMechanism : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emc2DChargedImpurityScatterMechanism(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("ChargedImpurity2D", inValley) {}
};

#endif
