// One implementation behind the four 3-D valley classes of the API
// (emcParabolicIsotropValley, emcNonParabolicIsotropValley,
// emcParabolicAnisotropValley, emcNonParabolicAnisotropValley): an ellipsoidal
// valley with optional Kane non-parabolicity in Herring-Vogt coordinates.
//
// The reference has four independent classes
// (include/ValleyTypes/emc{,Non}Parabolic{Isotrop,Anisotrop}Valley.hpp); their
// arithmetic differs only in details -- which are kept here, operation by
// operation, because the rate tables and the initial ensemble built on the host
// must agree with the reference's to the last bit:
//   * isotropic classes: m_c = m_DOS = rel * m0, vogt = 1, identity frames;
//   * anisotropic classes: m_DOS = cbrt(m0 m1 m2) m0, m_c = 3 m0 / sum(1/m_i),
//     vogt_i = sqrt(m_c / (m_i m0))                (reference ...Anistrop...:185-204);
//   * |k|(E): sqrt(2 m_c gamma q)/hbar in the non-parabolic anisotropic class,
//     sqrt(2 m q gamma)/hbar in the three others   (:103-106 vs. isotropic :72-75).
#ifndef EMC_DETAIL_ELLIPSOIDAL_VALLEY_HPP
#define EMC_DETAIL_ELLIPSOIDAL_VALLEY_HPP

#include <cmath>
#include <vector>

#include <ValleyTypes/emcAbstractValley.hpp>
#include <emcConstants.hpp>

namespace emcdetail {

template <class T, bool Anisotropic, bool NonParabolic> class EllipsoidalValley : public emcAbstractValley<T> {
  T particleMass;
  SizeType degeneracy;
  T alpha;
  T bottomEnergy;
  T massCond, massDOS;
  std::array<T, 3> vogt;
  std::vector<std::array<T, 9>> frames; // row-major, rows = ellipse axes in device coordinates

  static std::array<T, 9> identity() { return {1, 0, 0, 0, 1, 0, 0, 0, 1}; }

protected:
  EllipsoidalValley(std::array<T, 3> relMass, T inParticleMass, SizeType inDegeneracy, T inAlpha, T inBottomEnergy)
      : particleMass(inParticleMass), degeneracy(inDegeneracy), alpha(NonParabolic ? inAlpha : T(0)),
        bottomEnergy(inBottomEnergy), frames(inDegeneracy, identity()) {
    if (Anisotropic) {
      T prod = 1.;
      for (auto m : relMass)
        prod = prod * m;
      massDOS = std::pow(prod, 1. / 3.) * particleMass;
      T inv = 0.;
      for (auto m : relMass)
        inv += 1. / m;
      massCond = 3. * particleMass / inv;
      for (int i = 0; i < 3; i++)
        vogt[i] = std::sqrt(massCond / (relMass[i] * particleMass));
    } else {
      massCond = massDOS = relMass[0] * particleMass;
      vogt = {1, 1, 1};
    }
  }

public:
  // orthogonal (not necessarily normalised) ellipse axes of one sub-valley in device coordinates
  void setSubValleyEllipseCoordSystem(SizeType idxSubValley, std::array<T, 3> dir1, std::array<T, 3> dir2,
                                      std::array<T, 3> dir3) {
    static_assert(Anisotropic, "only anisotropic valleys have sub-valley frames");
    if (idxSubValley >= degeneracy)
      emcMessage::getInstance().addError("Using invalid subvalley index.").print();
    if (innerProduct(dir1, dir2) != 0 || innerProduct(dir1, dir3) != 0 || innerProduct(dir2, dir3) != 0)
      emcMessage::getInstance()
          .addError("The given coordinate system for a subvalley is not orthogonal, adapt that.")
          .print();
    std::array<T, 3> axes[3] = {dir1, dir2, dir3};
    for (int r = 0; r < 3; r++) {
      normalize(axes[r]);
      for (int c = 0; c < 3; c++)
        frames[idxSubValley][3 * r + c] = axes[r][c];
    }
  }

  T getEffMassDOS(T energy = 0) const override {
    return NonParabolic ? massDOS * std::pow(1 + 2 * alpha * energy, 3.) : massDOS;
  }
  T getEffMassCond(T energy = 0) const override { return NonParabolic ? massCond * (1 + 2 * energy * alpha) : massCond; }
  T getNonParabolicity() const override { return alpha; }
  T getBottomEnergy() const override { return bottomEnergy; }
  SizeType getDegeneracyFactor() const override { return degeneracy; }
  T getGamma(T energy) const override { return NonParabolic ? energy * (1 + alpha * energy) : energy; }

  T getNormWaveVec(T energy) const override {
    if (Anisotropic && NonParabolic)
      return std::sqrt(2 * massCond * getGamma(energy) * constants::q) / constants::hbar;
    return std::sqrt(2 * massCond * constants::q * getGamma(energy)) / constants::hbar;
  }
  T getEnergy(const std::array<T, 3> &k) const override {
    if (NonParabolic) {
      const T g = constants::hbar * constants::hbar * square(k) / (massCond * constants::q);
      return g / (1 + std::sqrt(1 + 2 * alpha * g));
    }
    return constants::hbar * constants::hbar * square(k) / (2 * massCond * constants::q);
  }
  std::array<T, 3> getVelocity(const std::array<T, 3> &k, T energy, SizeType idxSubValley) const override {
    const T npf = NonParabolic ? std::sqrt(1 + 4 * alpha * getGamma(energy)) : T(1);
    if (!Anisotropic)
      return scale(k, NonParabolic ? constants::hbar / (massCond * npf) : constants::hbar / massCond);
    std::array<T, 3> v = transformToEllipseCoord(idxSubValley, k);
    for (int i = 0; i < 3; i++)
      v[i] = NonParabolic ? constants::hbar * vogt[i] * v[i] / (massCond * npf)
                          : constants::hbar * vogt[i] * v[i] / massCond;
    return transformToDeviceCoord(idxSubValley, v);
  }
  const std::array<T, 3> &getVogtTransformationFactor() const override { return vogt; }

  std::array<T, 3> transformToEllipseCoord(SizeType s, const std::array<T, 3> &v) const override {
    if (!Anisotropic)
      return v;
    const auto &r = frames[s];
    return {v[0] * r[0] + v[1] * r[1] + v[2] * r[2], v[0] * r[3] + v[1] * r[4] + v[2] * r[5],
            v[0] * r[6] + v[1] * r[7] + v[2] * r[8]};
  }
  std::array<T, 3> transformToDeviceCoord(SizeType s, const std::array<T, 3> &v) const override {
    if (!Anisotropic)
      return v;
    const auto &r = frames[s];
    return {v[0] * r[0] + v[1] * r[3] + v[2] * r[6], v[0] * r[1] + v[1] * r[4] + v[2] * r[7],
            v[0] * r[2] + v[1] * r[5] + v[2] * r[8]};
  }

  // EMCGPU_VALLEY_* of include/emcgpu.h: parabolic iso 0, non-parabolic iso 1, parabolic aniso 2, non-parabolic aniso 3
  int deviceValleyKind() const override { return (Anisotropic ? 2 : 0) + (NonParabolic ? 1 : 0); }
  // row-major frame of a sub-valley (what emcgpu_valley_t::rot takes)
  const std::array<T, 9> &frame(SizeType idxSubValley) const { return frames[idxSubValley]; }
};

} // namespace emcdetail

#endif
