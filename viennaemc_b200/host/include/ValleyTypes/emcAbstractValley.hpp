// Plug-in interface of one band-structure valley.
// Interface mirrored: reference include/ValleyTypes/emcAbstractValley.hpp:20-91 (the
// twelve pure virtuals and check()).
//
// Additive for the GPU path: deviceValleyKind().  The device kernels evaluate
// the dispersion themselves (include/emcgpu.h, emcgpu_valley_t); a valley
// class tells the binding which of the built-in dispersions it is.  A user
// valley that does not override it is rejected with a clear error when the
// ensemble is moved to the GPU -- it is never evaluated on the CPU instead.
#ifndef EMC_ABSTRACT_VALLEY_HPP
#define EMC_ABSTRACT_VALLEY_HPP

#include <array>

#include <emcMessage.hpp>
#include <emcUtil.hpp>

template <class T> class emcAbstractValley {
public:
  virtual ~emcAbstractValley() = default;

  virtual T getEffMassDOS(T energy = 0) const = 0;
  virtual T getEffMassCond(T energy = 0) const = 0;
  virtual T getNonParabolicity() const = 0;
  virtual T getBottomEnergy() const = 0;
  virtual SizeType getDegeneracyFactor() const = 0;
  virtual T getNormWaveVec(T energy) const = 0;
  virtual T getEnergy(const std::array<T, 3> &k) const = 0;
  virtual T getGamma(T energy) const = 0;
  virtual std::array<T, 3> getVelocity(const std::array<T, 3> &k, T energy, SizeType idxSubValley) const = 0;
  virtual const std::array<T, 3> &getVogtTransformationFactor() const = 0;
  virtual std::array<T, 3> transformToDeviceCoord(SizeType idxSubValley, const std::array<T, 3> &vec) const = 0;
  virtual std::array<T, 3> transformToEllipseCoord(SizeType idxSubValley, const std::array<T, 3> &vec) const = 0;

  // emcgpu_valley_kind of include/emcgpu.h, or -1: no device implementation
  virtual int deviceValleyKind() const { return -1; }

  void check() const {
    if (getDegeneracyFactor() < 1)
      emcMessage::getInstance().addError("Degeneracy Factor of a valley has to be at least 1.").print();
    if (getNonParabolicity() < 0)
      emcMessage::getInstance().addError("NonParabolicity of a valley has to be positive.").print();
  }
};

#endif
