// Piezoelectric acoustic-phonon scattering in a single layer.  Name mirrored: reference
// include/ScatterMechanisms/emcPiezoelectricSingleLayerScatterMechanism.hpp.
// NAME ONLY (detail/emcNoDeviceSampler.hpp): constructible with the reference's arguments, rejected with its name when added
// to a particle type -- no device final-state sampler yet, and nothing is ever scattered on the CPU.
#ifndef EMC_PIEZOELECTRIC_SINGLE_LAYER_SCATTER_MECHANISM_HPP
#define EMC_PIEZOELECTRIC_SINGLE_LAYER_SCATTER_MECHANISM_HPP

#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <detail/emcNoDeviceSampler.hpp>

template <class T> struct emcPiezoelectricSingleLayerMechanism : public emcdetail::NoDeviceSamplerMechanism<T> {
  template <class... Args>
  explicit emcPiezoelectricSingleLayerMechanism(SizeType inValley, Args &&...) : emcdetail::NoDeviceSamplerMechanism<T>("PiezoelectricSL", inValley) {}
};

#endif
