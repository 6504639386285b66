// Common part of the polar-optical (Froehlich) mechanisms: coupling constant, rate, device sampler descriptor.
// Logic mirrored: reference include/ScatterMechanisms/emcFroehlichInteraction.hpp (coupling constant :33-40,
// Bose-Einstein occupation :43-47, rates :96-107 / :168-181), emcHotPhononFroehlichMechanism.hpp (occupation from a
// phonon bath :79-92 / :158-173) and emcScreenedFroehlichInteraction.hpp (screened logarithm :54-63, q-resolved
// occupation :256-270 / :339-352).  The eight public classes differ only in where the phonon occupation comes from,
// whether the coupling is screened and how the polar angle is drawn -- they are thin wrappers around this one.
//
// Rates are host code (they feed emcScatterHandler's tables, as in the reference).  The final state -- polar angle,
// rotation about the current k, new |k|, and for hot-phonon mechanisms the event count in the |q| bin of the bath --
// is device code: EMCGPU_SAMPLER_FROEHLICH / EMCGPU_SAMPLER_SCREENED_FROEHLICH (viennaemc_b200/csrc/emc_device.cuh).
#ifndef EMC_DETAIL_POLAR_OPTICAL_MECHANISM_HPP
#define EMC_DETAIL_POLAR_OPTICAL_MECHANISM_HPP

#include <cmath>
#include <memory>
#include <string>

#include <emcgpu.h>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <emcConstants.hpp>
#include <emcMessage.hpp>
#include <emcPhononBath.hpp>
#include <emcPlasmonScreening.hpp>

// C = e^2 w0 m* / (4 pi hbar^2) (1/eps_hi - 1/eps_lo) / eps0   [1/(m s)]
template <class T> T froehlichScatterConst3D(T phononEnergy, T effMass, T eps_hi, T eps_lo) {
  const T omega0 = phononEnergy * constants::q / constants::hbar;
  return constants::q * constants::q * omega0 * effMass / (4. * constants::pi * constants::hbar * constants::hbar) *
         (T(1) / eps_hi - T(1) / eps_lo) / constants::eps0;
}
template <class T> T boseEinstein(T phononEnergy, T temperature) {
  const T x = constants::q * phononEnergy / (constants::kB * temperature);
  return T(1) / (std::exp(x) - T(1));
}
// (1/2) ln((q+^2 + qs^2) / (q-^2 + qs^2)) with q+- = kI +- kF
template <class T> T screenedFroehlichLogFactor(T kI, T kF, T qs2) {
  const T qPlus2 = (kI + kF) * (kI + kF), qMinus2 = (kI - kF) * (kI - kF);
  if (qs2 <= T(0))
    return qMinus2 <= T(0) ? T(0) : T(0.5) * std::log(qPlus2 / qMinus2);
  return T(0.5) * std::log((qPlus2 + qs2) / (qMinus2 + qs2));
}

namespace emcdetail {

template <class T> class PolarOpticalMechanism : public emcScatterMechanism<T> {
protected:
  const bool emission;
  const bool screened; // screened coupling + closed-form / q-resolved angle instead of the unscreened power law
  T phononEnergy;      // [eV]
  T effMass;           // [kg]
  T scatterConst;
  T nBose = T(0); // used when no bath is attached
  std::shared_ptr<emcPhononBath<T>> phononBath;
  std::shared_ptr<emcPlasmonScreening<T>> screening;
  bool qResolved = false, qResolvedAngle = true;
  std::string baseName, nameSuffix;

  PolarOpticalMechanism(const char *inBaseName, bool inEmission, bool inScreened, SizeType inValley, T inPhononEnergy,
                        T relEffMass, T eps_hi, T eps_lo, std::string inNameSuffix)
      : emcScatterMechanism<T>(inValley), emission(inEmission), screened(inScreened), phononEnergy(inPhononEnergy),
        effMass(relEffMass * constants::me),
        scatterConst(froehlichScatterConst3D(inPhononEnergy, relEffMass * constants::me, eps_hi, eps_lo)),
        baseName(inBaseName), nameSuffix(std::move(inNameSuffix)) {}

  T qs2() const { return screening ? screening->getQs2() : T(0); }

public:
  std::string getName() const override { return baseName + nameSuffix; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    if (emission && energy <= phononEnergy)
      return T(0);
    const auto *valley = this->ptrValley[this->idxValley];
    const T gammaI = valley->getGamma(energy);
    const T gammaF = valley->getGamma(emission ? energy - phononEnergy : energy + phononEnergy);
    if (gammaI <= T(0) || gammaF <= T(0))
      return T(0);
    const T kI = std::sqrt(T(2) * effMass * gammaI * constants::q) / constants::hbar;
    const T kF = std::sqrt(T(2) * effMass * gammaF * constants::q) / constants::hbar;
    T occupation = nBose;
    if (phononBath)
      occupation = (screened && qResolved) ? phononBath->getNqInWindow(std::fabs(kI - kF), kI + kF) : phononBath->getMeanNq();
    if (emission)
      occupation = occupation + T(1);
    if (screened)
      return scatterConst * occupation / kI * screenedFroehlichLogFactor(kI, kF, qs2());
    const T lnFactor = emission ? std::log((kI + kF) / (kI - kF)) : std::log((kI + kF) / (kF - kI));
    return scatterConst * occupation / kI * lnFactor;
  }

  // the final state is sampled on the device; this entry point of the plug-in interface is not a second implementation
  void scatterParticle(emcParticle<T> &, emcRNG &) const override {
    emcMessage::getInstance()
        .addError(getName() + "::scatterParticle runs on the GPU (EMCGPU_SAMPLER_FROEHLICH / _SCREENED_FROEHLICH); there is "
                              "no host implementation.")
        .print();
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = screened ? EMCGPU_SAMPLER_SCREENED_FROEHLICH : EMCGPU_SAMPLER_FROEHLICH;
    d.finalValley = this->idxValley;
    d.param[0] = emission ? -phononEnergy : phononEnergy;
    d.param[1] = qs2();
    d.param[2] = -1; // bath index: filled in by the GPU binding
    d.param[3] = (screened && phononBath && qResolved && qResolvedAngle) ? 1. : 0.;
    return d;
  }
  emcPhononBath<T> *devicePhononBath() const override { return phononBath.get(); }
};

} // namespace emcdetail

#endif
