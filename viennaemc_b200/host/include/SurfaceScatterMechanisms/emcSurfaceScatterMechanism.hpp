// Scattering at a wall of the simulation box (a face that is not an ohmic contact).
// Interface mirrored: reference include/SurfaceScatterMechanisms/emcSurfaceScatterMechanism.hpp (ctor: box extent;
// setBoundaryPosition :49-51; the probability of a diffusive event getDiffScatterProb :59).
//
// The event itself -- decide diffusive / specular, draw the new direction, fold the position back into the box
// (scatterParticle :40-46, calculateAndAssignKAndPos :127-147, scatterParticleSpecularly :81-93) -- is device code
// (surfaceScatter, viennaemc_b200/csrc/emc_device_run.cuh), selected by the additive deviceSurfaceKind().  A subclass
// without a device implementation is rejected when its particle type is handed to a GPU particle handler.
#ifndef EMC_SURFACE_SCATTER_MECHANISM_HPP
#define EMC_SURFACE_SCATTER_MECHANISM_HPP

#include <array>

#include <emcgpu.h>

#include <emcBoundaryPos.hpp>
#include <emcParticle.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> class emcSurfaceScatterMechanism {
protected:
  static const SizeType Dim = DeviceType::Dimension;
  emcBoundaryPos boundaryPos = emcBoundaryPos::INVALID;
  std::array<T, Dim> maxPos;

  // index of the k / position component normal to the wall (+delta, cyclic)
  int getIndexFromBoundaryPos(int delta = 0) const { return static_cast<int>((toUnderlying(boundaryPos) / 2 + delta) % 3); }
  bool isScatteringAtMaxPos() const { return toUnderlying(boundaryPos) % 2 == 1; }

public:
  emcSurfaceScatterMechanism() = delete;
  explicit emcSurfaceScatterMechanism(std::array<T, Dim> inMaxPos) : maxPos(inMaxPos) {}
  virtual ~emcSurfaceScatterMechanism() = default;

  void setBoundaryPosition(emcBoundaryPos inBoundaryPos) { boundaryPos = inBoundaryPos; }
  emcBoundaryPos getBoundaryPosition() const { return boundaryPos; }

  // probability that a particle hitting the wall leaves it in a new random direction
  virtual T getDiffScatterProb(emcParticle<T> &particle) const = 0;

  // --- additive: how the device runs this mechanism (emcgpu_surface_kind, one parameter) ---
  virtual int deviceSurfaceKind() const { return -1; } // -1: no device implementation
  virtual T deviceSurfaceParameter() const { return 0; }
};

#endif
