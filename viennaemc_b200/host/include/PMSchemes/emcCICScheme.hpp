// Cloud-in-cell scheme: charge and force are shared between the 2^Dim corners of the mesh cell a particle lies in.
// Interface mirrored: reference include/PMSchemes/emcCICScheme.hpp (assignToMesh :43-118, interpolateForce :122-173,
// calcEField :176-180 = calcEFieldAtGridPts).  The device kernels reproduce the reference's weights as they are: the
// distance from the LOWER grid point weights the lower grid point, and the force uses (1 - wY)(1 - wY) for the
// upper-right corner.
#ifndef EMC_CIC_SCHEME_HPP
#define EMC_CIC_SCHEME_HPP

#include <PMSchemes/emcAbstractPMScheme.hpp>

template <class T, class DeviceType> class emcCICScheme : public emcDevicePMScheme<T, DeviceType> {
public:
  emcCICScheme() : emcDevicePMScheme<T, DeviceType>("emcCICScheme") {}
  int deviceSchemeId() const override { return 2; }
};

#endif
