// Zero-order intervalley (or optical intravalley) phonon scattering in a single layer, absorption and emission.
// Interface mirrored: reference include/ScatterMechanisms/emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp; the body is
// detail/emcSingleLayerInterValley.hpp.  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY.
#ifndef EMC_ZERO_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP
#define EMC_ZERO_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP

#include <detail/emcSingleLayerInterValley.hpp>

EMC_SINGLE_LAYER_INTERVALLEY_CLASS(emcZeroOrderSingleLayerInterValleyAbsorptionScatterMechanism, 0, true);
EMC_SINGLE_LAYER_INTERVALLEY_CLASS(emcZeroOrderSingleLayerInterValleyEmissionScatterMechanism, 0, false);

#endif
