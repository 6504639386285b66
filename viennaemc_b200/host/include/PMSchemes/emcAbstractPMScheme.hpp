// Plug-in interface of the particle-mesh scheme.  Interface mirrored: reference
// include/PMSchemes/emcAbstractPMScheme.hpp:20-77 (assignToMesh x2, interpolateForce, calcEField).
//
// On the GPU path the scheme is a kernel variant, selected by the additive deviceSchemeId(); a scheme
// without one cannot be used with the GPU particle handler.
#ifndef EMC_ABSTRACT_PM_SCHEME_HPP
#define EMC_ABSTRACT_PM_SCHEME_HPP

#include <array>
#include <vector>

#include <emcGrid.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> class emcAbstractPMScheme {
public:
  static const SizeType Dim = DeviceType::Dimension;
  virtual ~emcAbstractPMScheme() = default;
  virtual void assignToMesh(const std::array<T, Dim> &pos, SizeType nrCarriers, const std::array<T, Dim> &spacing,
                            emcGrid<T, Dim> &gridNrParticles) const = 0;
  virtual void assignToMesh(const std::vector<std::array<T, Dim>> &position, SizeType nrCarriers,
                            const std::array<T, Dim> &spacing, emcGrid<T, Dim> &gridNrParticles) const = 0;
  virtual std::array<T, 3> interpolateForce(const std::vector<emcGrid<T, Dim>> &eField, const std::array<T, Dim> &pos,
                                            const std::array<T, Dim> &spacing, T charge) const = 0;
  virtual void calcEField(std::vector<emcGrid<T, Dim>> &eField, const emcGrid<T, Dim> &potential,
                          const DeviceType &device) const = 0;
  // 0 = no device implementation; 1 = nearest grid point
  virtual int deviceSchemeId() const { return 0; }
};

#endif
