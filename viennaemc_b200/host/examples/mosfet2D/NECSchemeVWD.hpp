// Nearest-element-centre scheme as done in ViennaWD -- the particle-mesh scheme of the MOSFET example.
// Interface mirrored: reference examples/mosfet2D/NECSchemeVWD.hpp (same class name, 2-D only): equal charge shares
// for the four corners of the mesh cell (:25-52); force from the edge mid-point fields with the x index ROUNDED
// instead of floored and a one-sided y average in the last column (:57-76); E field by forward differences with
// E = 0 on the max faces and no contact rule (:82-99).  All three are variants of the device kernels
// (emcgpu_pm_scheme EMCGPU_PM_NEC_VWD); see emcDevicePMScheme in PMSchemes/emcAbstractPMScheme.hpp.
#ifndef EMC_NEC_SCHEME_VWD_HPP
#define EMC_NEC_SCHEME_VWD_HPP

#include <PMSchemes/emcAbstractPMScheme.hpp>

template <class T, class DeviceType> class emcNECSchemeVWD : public emcDevicePMScheme<T, DeviceType> {
public:
  static_assert(DeviceType::Dimension == 2, "PMScheme NEC-VWD is only implemented for 2D.");
  emcNECSchemeVWD() : emcDevicePMScheme<T, DeviceType>("emcNECSchemeVWD") {}
  int deviceSchemeId() const override { return 4; }
};

#endif
