// First-order intervalley phonon scattering in a single layer, absorption and emission.
// Interface mirrored: reference include/ScatterMechanisms/emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp; the body is
// detail/emcSingleLayerInterValley.hpp.  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY with param[1] = 1.
#ifndef EMC_FIRST_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP
#define EMC_FIRST_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP

#include <detail/emcSingleLayerInterValley.hpp>

EMC_SINGLE_LAYER_INTERVALLEY_CLASS(emcFirstOrderSingleLayerInterValleyAbsorptionScatterMechanism, 1, true);
EMC_SINGLE_LAYER_INTERVALLEY_CLASS(emcFirstOrderSingleLayerInterValleyEmissionScatterMechanism, 1, false);

#endif
