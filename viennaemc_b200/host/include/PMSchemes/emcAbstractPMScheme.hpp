// Plug-in interface of the particle-mesh scheme.  Interface mirrored: reference
// include/PMSchemes/emcAbstractPMScheme.hpp:20-77 (assignToMesh x2, interpolateForce, calcEField).
//
// On the GPU path the scheme is a kernel variant, selected by the additive deviceSchemeId(); a scheme
// without one cannot be used with the GPU particle handler.
#ifndef EMC_ABSTRACT_PM_SCHEME_HPP
#define EMC_ABSTRACT_PM_SCHEME_HPP

#include <array>
#include <vector>

#include <string>

#include <emcGrid.hpp>
#include <emcMessage.hpp>
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
  // 0 = no device implementation; otherwise emcgpu_pm_scheme + 1: 1 nearest grid point, 2 cloud in cell,
  // 3 nearest element centre, 4 nearest element centre as in ViennaWD (examples/mosfet2D)
  virtual int deviceSchemeId() const { return 0; }
};

// Schemes whose work -- charge assignment, force gather, E = -grad(phi) -- is done by the device kernels
// (ngpAssignKernel, pmForce in deviceStepKernel, cellEField) inside the GPU particle handler / emcSimulation.  The
// per-call host entry points of the interface are not a second implementation: they report that the scheme runs on the
// GPU.
template <class T, class DeviceType> class emcDevicePMScheme : public emcAbstractPMScheme<T, DeviceType> {
  const char *name;
  void gpuOnly(const char *what) const {
    emcMessage::getInstance()
        .addError(std::string(name) + "::" + what +
                  " runs on the GPU inside emcBasicParticleHandler / emcSimulation; there is no host implementation.")
        .print();
  }

protected:
  explicit emcDevicePMScheme(const char *inName) : name(inName) {}

public:
  static const SizeType Dim = DeviceType::Dimension;
  void assignToMesh(const std::array<T, Dim> &, SizeType, const std::array<T, Dim> &, emcGrid<T, Dim> &) const override {
    gpuOnly("assignToMesh");
  }
  void assignToMesh(const std::vector<std::array<T, Dim>> &, SizeType, const std::array<T, Dim> &,
                    emcGrid<T, Dim> &) const override {
    gpuOnly("assignToMesh");
  }
  std::array<T, 3> interpolateForce(const std::vector<emcGrid<T, Dim>> &, const std::array<T, Dim> &,
                                    const std::array<T, Dim> &, T) const override {
    gpuOnly("interpolateForce");
    return {0, 0, 0};
  }
  void calcEField(std::vector<emcGrid<T, Dim>> &, const emcGrid<T, Dim> &, const DeviceType &) const override {
    gpuOnly("calcEField");
  }
};

#endif
