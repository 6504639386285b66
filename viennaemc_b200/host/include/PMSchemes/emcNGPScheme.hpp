// Nearest-grid-point scheme.  Interface mirrored: reference include/PMSchemes/emcNGPScheme.hpp
// (assignToMesh :36-47, interpolateForce :51-66, calcEField :69-73).
// The work itself -- charge assignment, force gather, E = -grad(phi) -- is done by the device kernels
// (ngpAssignKernel, deviceStepKernel, efieldKernel) inside the GPU particle handler / emcSimulation.
// The per-call host entry points of the interface are not a second implementation: they report that
// the scheme runs on the GPU.
#ifndef EMC_NGP_SCHEME_HPP
#define EMC_NGP_SCHEME_HPP

#include <PMSchemes/emcAbstractPMScheme.hpp>
#include <emcMessage.hpp>

template <class T, class DeviceType> class emcNGPScheme : public emcAbstractPMScheme<T, DeviceType> {
  static void gpuOnly(const char *what) {
    emcMessage::getInstance()
        .addError(std::string("emcNGPScheme::") + what +
                  " runs on the GPU inside emcBasicParticleHandler / emcSimulation; there is no host implementation.")
        .print();
  }

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
  int deviceSchemeId() const override { return 1; }
};

#endif
