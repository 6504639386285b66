// One implementation behind the three single-layer (2-D material in the x-y plane) valley classes of the API:
// emcParabolicIsotropSingleLayerValley, emcNonParabolicIsotropSingleLayerValley, emcNonParabolicAnisotropSingleLayerValley.
//
// Arithmetic mirrored operation by operation (the rate tables and the initial ensemble built on the host must agree with the
// reference's to the last bit): reference include/ValleyTypes/emcParabolicIsotropSingleLayerValley.hpp:44-80,
// emcNonParabolicIsotropSingleLayerValley.hpp:50-93, emcNonParabolicAnisotropSingleLayerValley.hpp:79-169.
//   * isotropic classes: one mass, Herring-Vogt factors (1, 1, 0), no sub-valley frames;
//   * anisotropic class: m_DOS = sqrt(m_l m_t) is the mass of the dispersion, of |k|(E) and of the velocity,
//     m_c = 2 / (1/m_l + 1/m_t) the one getEffMassCond() reports (the position update of the free flight uses it),
//     vogt = (sqrt(m_DOS/m_l), sqrt(m_DOS/m_t), 0), one in-plane rotation angle per sub-valley;
//   * k_z, v_z and the z component of every transformed vector are 0.
// Device dispersion: EMCGPU_VALLEY_*_SINGLE_LAYER of include/emcgpu.h.
#ifndef EMC_DETAIL_SINGLE_LAYER_VALLEY_HPP
#define EMC_DETAIL_SINGLE_LAYER_VALLEY_HPP

#include <cmath>
#include <vector>

#include <ValleyTypes/emcAbstractValley.hpp>
#include <emcConstants.hpp>

namespace emcdetail {

template <class T, bool Anisotropic, bool NonParabolic> class SingleLayerValley : public emcAbstractValley<T> {
  SizeType degeneracy;
  T alpha;
  T bottomEnergy;
  T massCond, massBand; // massBand: m_DOS of the anisotropic class, the one mass of the isotropic classes
  std::array<T, 3> vogt;
  std::vector<T> angle, sinAngle, cosAngle;

protected:
  // isotropic
  SingleLayerValley(T relMass, T particleMass, SizeType inDegeneracy, T inAlpha, T inBottomEnergy)
      : degeneracy(inDegeneracy), alpha(NonParabolic ? inAlpha : T(0)), bottomEnergy(inBottomEnergy),
        massCond(relMass * particleMass), massBand(relMass * particleMass), vogt({1, 1, 0}) {}
  // anisotropic: longitudinal / transversal relative masses, one rotation angle [0, 2 pi] per sub-valley
  SingleLayerValley(T relMassLong, T relMassTrans, T particleMass, SizeType inDegeneracy, T inAlpha,
                    std::vector<T> inAngles, T inBottomEnergy)
      : degeneracy(inDegeneracy), alpha(inAlpha), bottomEnergy(inBottomEnergy), angle(std::move(inAngles)) {
    const T massLong = relMassLong * particleMass, massTrans = relMassTrans * particleMass;
    massCond = 2. / (1. / massLong + 1. / massTrans);
    massBand = std::sqrt(massLong * massTrans);
    vogt = {std::sqrt(massBand / massLong), std::sqrt(massBand / massTrans), 0};
    if (angle.size() != degeneracy)
      emcMessage::getInstance().addError("Wrong size of rotationAngle vector!").print();
    sinAngle.assign(degeneracy, 0.);
    cosAngle.assign(degeneracy, 0.);
    for (SizeType s = 0; s < degeneracy; s++)
      setSubValleyEllipseCoordSystem(s, angle[s]);
  }

public:
  void setSubValleyEllipseCoordSystem(SizeType idxSubValley, T newRotationAngle) {
    static_assert(Anisotropic, "only the anisotropic single-layer valley has sub-valley frames");
    if (newRotationAngle < 0 || newRotationAngle > 2 * constants::pi)
      emcMessage::getInstance().addError("Rotation angle has to between 0 and 2 * PI.").print();
    angle[idxSubValley] = newRotationAngle;
    sinAngle[idxSubValley] = std::sin(newRotationAngle);
    cosAngle[idxSubValley] = std::cos(newRotationAngle);
  }

  T getBottomEnergy() const override { return bottomEnergy; }
  T getEffMassDOS(T energy = 0) const override {
    return NonParabolic ? massBand * std::pow(1 + 2 * alpha * energy, 3.) : massBand;
  }
  T getEffMassCond(T energy = 0) const override { return NonParabolic ? massCond * (1 + 2 * energy * alpha) : massCond; }
  T getNonParabolicity() const override { return alpha; }
  SizeType getDegeneracyFactor() const override { return degeneracy; }
  T getGamma(T energy) const override { return NonParabolic ? energy * (1 + alpha * energy) : energy; }

  T getNormWaveVec(T energy) const override {
    return std::sqrt(2 * massBand * constants::q * getGamma(energy)) / constants::hbar;
  }
  T getEnergy(const std::array<T, 3> &k) const override {
    if (NonParabolic) {
      const T g = constants::hbar * constants::hbar * (k[0] * k[0] + k[1] * k[1]) / (massBand * constants::q);
      return g / (1 + std::sqrt(1 + 2 * alpha * g));
    }
    return constants::hbar * constants::hbar * (k[0] * k[0] + k[1] * k[1]) / (2 * massBand * constants::q);
  }
  std::array<T, 3> getVelocity(const std::array<T, 3> &k, T energy, SizeType idxSubValley) const override {
    if (!Anisotropic)
      return scale(k, NonParabolic ? constants::hbar / (massBand * std::sqrt(1 + 4 * alpha * getGamma(energy)))
                                   : constants::hbar / massBand);
    std::array<T, 3> v = transformToEllipseCoord(idxSubValley, k);
    const T npf = std::sqrt(1 + 4 * alpha * getGamma(energy));
    for (int i = 0; i < 2; i++)
      v[i] = constants::hbar * vogt[i] * v[i] / (massBand * npf);
    return transformToDeviceCoord(idxSubValley, v);
  }
  const std::array<T, 3> &getVogtTransformationFactor() const override { return vogt; }

  // in-plane rotation by the sub-valley's angle, and back
  std::array<T, 3> transformToEllipseCoord(SizeType s, const std::array<T, 3> &v) const override {
    if (!Anisotropic)
      return v;
    return {v[0] * cosAngle[s] - v[1] * sinAngle[s], v[0] * sinAngle[s] + v[1] * cosAngle[s], 0};
  }
  std::array<T, 3> transformToDeviceCoord(SizeType s, const std::array<T, 3> &v) const override {
    if (!Anisotropic)
      return v;
    return {v[0] * cosAngle[s] + v[1] * sinAngle[s], -v[0] * sinAngle[s] + v[1] * cosAngle[s], 0};
  }

  // EMCGPU_VALLEY_{PARABOLIC_ISOTROP, NONPARABOLIC_ISOTROP, NONPARABOLIC_ANISOTROP}_SINGLE_LAYER = 4, 5, 7
  int deviceValleyKind() const override { return 4 + (Anisotropic ? 2 : 0) + (NonParabolic ? 1 : 0); }
};

} // namespace emcdetail

#endif
