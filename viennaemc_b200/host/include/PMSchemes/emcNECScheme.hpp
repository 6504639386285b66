// Nearest-element-centre scheme: equal charge shares for the corners of the mesh cell, force from the fields at the
// edge mid-points of that cell.
// Interface mirrored: reference include/PMSchemes/emcNECScheme.hpp (assignToMesh :33-96, interpolateForce :99-113 --
// 2-D only in the reference, and so here --, calcEField :116-120 = calcEFieldAtEdgeMidPts).
#ifndef EMC_NEC_SCHEME_HPP
#define EMC_NEC_SCHEME_HPP

#include <PMSchemes/emcAbstractPMScheme.hpp>

template <class T, class DeviceType> class emcNECScheme : public emcDevicePMScheme<T, DeviceType> {
public:
  emcNECScheme() : emcDevicePMScheme<T, DeviceType>("emcNECScheme") {}
  int deviceSchemeId() const override { return 3; }
};

#endif
