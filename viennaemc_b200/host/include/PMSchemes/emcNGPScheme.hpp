// Nearest-grid-point scheme.  Interface mirrored: reference include/PMSchemes/emcNGPScheme.hpp
// (assignToMesh :36-47, interpolateForce :51-66, calcEField :69-73 = calcEFieldAtGridPts).
// Device kernels: see emcDevicePMScheme (PMSchemes/emcAbstractPMScheme.hpp).
#ifndef EMC_NGP_SCHEME_HPP
#define EMC_NGP_SCHEME_HPP

#include <PMSchemes/emcAbstractPMScheme.hpp>

template <class T, class DeviceType> class emcNGPScheme : public emcDevicePMScheme<T, DeviceType> {
public:
  emcNGPScheme() : emcDevicePMScheme<T, DeviceType>("emcNGPScheme") {}
  int deviceSchemeId() const override { return 1; }
};

#endif
