// Polar-optical (Froehlich) scattering in a single layer, absorption and emission (Kaasbjerg et al., PRB 85, 115317; parabolic
// bands), optionally screened by the 2-D carrier gas.  Interface mirrored: reference
// include/ScatterMechanisms/emcFroehlichInteractionSingleLayer.hpp (free function :45-80, ctors :120-132 / :239-251, rates :138-145
// / :257-269 with the midpoint rule :201-208, samplers :149-168 / :273-292).  One implementation for both classes.
// Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_FROEHLICH (param: signed phonon energy, form-factor width, screening wave vector).
#ifndef EMC_FROEHLICH_INTERACTION_SINGLE_LAYER_HPP
#define EMC_FROEHLICH_INTERACTION_SINGLE_LAYER_HPP

#include <cassert>
#include <cmath>
#include <random>
#include <string>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerAngle.hpp>
#include <emcConstants.hpp>

// deflection between k and k' of a polar-optical event in a single layer: weight erfc(w q/2)^2 / (eps(q)^2 q) with
// q^2 = k^2 + k'^2 - 2 k k' cos(psi); returns psi in [0, 2 pi)
template <class T>
T sampleSingleLayerFroehlichDeflectionAngle(T k, T kPrime, T effWidth, T screeningWavevector, emcRNG &rng,
                                            std::uniform_real_distribution<T> &dist) {
  const auto sums = emcdetail::cumulativeAngularWeight<T>([&](T psi) -> T {
    const T q2 = k * k + kPrime * kPrime - 2 * k * kPrime * std::cos(psi);
    const T q = std::sqrt(std::max(T(0), q2));
    if (q <= T(0))
      return T(0);
    return emcdetail::formFactorScreened(q, effWidth, screeningWavevector) / q;
  });
  const T total = sums[emcdetail::singleLayerAngleSteps];
  if (!(total > T(0)))
    return 2 * constants::pi * dist(rng); // nothing to weight with: isotropic
  const T magnitude = emcdetail::invertAngularWeight(sums, dist(rng) * total);
  return (dist(rng) < T(0.5)) ? magnitude : (2 * constants::pi - magnitude);
}

namespace emcdetail {

template <class T, bool Absorption> class SingleLayerFroehlich : public emcScatterMechanism<T> {
  T phononEnergy;
  T effWidth;
  T prefactor; // (C q)^2 N / (2 pi hbar^3), N = n_B (absorption) or n_B + 1 (emission)
  T screeningWavevector;
  std::string nameSuffix;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  // integrand of the rate over the scattering angle (momentum transfer of the branch(es) open at this angle)
  T integrand(T theta, T k, T eFactor) const {
    const T cosTheta = std::cos(theta);
    if (Absorption) {
      const T root = std::sqrt(cosTheta * cosTheta + eFactor);
      const T q = k * (-cosTheta + root);
      return (-cosTheta + root) / root * std::pow(std::erfc(effWidth * q / 2.), 2) *
             twoDScreeningFactor(q, screeningWavevector);
    }
    const T root = std::sqrt(cosTheta * cosTheta - eFactor);
    const T qPlus = k * (cosTheta + root);
    T partPlus = (cosTheta + root);
    partPlus *= std::pow(std::erfc(effWidth * qPlus / 2.), 2) * twoDScreeningFactor(qPlus, screeningWavevector);
    const T qMinus = k * (cosTheta - root);
    T partMinus = (cosTheta - root);
    partMinus *= std::pow(std::erfc(effWidth * qMinus / 2.), 2) * twoDScreeningFactor(qMinus, screeningWavevector);
    return (partPlus + partMinus) / root;
  }
  T midpointRule(T a, T b, SizeType nrIntervals, T k, T eFactor) const {
    const T dx = (b - a) / (T)nrIntervals;
    T sum = 0;
    for (T at = a + dx / 2.; at <= b - dx / 2.; at += dx)
      sum += integrand(at, k, eFactor);
    return sum * dx;
  }

public:
  SingleLayerFroehlich() = delete;
  // phonon energy [eV], coupling constant [eV/m], effective width of the layer [m]; screening wave vector [1/m], 0: none
  SingleLayerFroehlich(SizeType inValley, T inPhononEnergy, T couplingConstant, T effectiveWidth, T temperature,
                       std::string inNameSuffix = "", T inScreeningWavevector = 0)
      : emcScatterMechanism<T>(inValley), phononEnergy(inPhononEnergy), effWidth(effectiveWidth),
        screeningWavevector(inScreeningWavevector), nameSuffix(inNameSuffix) {
    const T exponent = constants::q * phononEnergy / (constants::kB * temperature);
    const T nrPhonons = Absorption ? 1. / (std::exp(exponent) - 1.) : std::exp(exponent) / (std::exp(exponent) - 1.);
    prefactor = std::pow(couplingConstant * constants::q, 2) * nrPhonons / (2 * constants::pi * std::pow(constants::hbar, 3));
  }

  std::string getName() const override { return std::string("froehlich") + (Absorption ? "Absorption" : "Emission") + "SL" + nameSuffix; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    if (!Absorption && !(energy > phononEnergy))
      return 0;
    const auto *valley = this->ptrValley[this->idxValley];
    const T md = valley->getEffMassDOS();
    const T k = valley->getNormWaveVec(energy);
    const T eFactor = phononEnergy / energy;
    T integral;
    if (Absorption) {
      integral = midpointRule(0., 2 * constants::pi, 10000, k, eFactor);
    } else {
      const T thetaMax = std::acos(std::sqrt(eFactor));
      integral = midpointRule(-thetaMax, thetaMax, 10000, k, eFactor);
    }
    return integral * prefactor * md;
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    const auto *valley = this->ptrValley[this->idxValley];
    const T before = particle.energy;
    particle.energy = Absorption ? before + phononEnergy : before - phononEnergy;
    assert(particle.energy > 0);
    const T k = valley->getNormWaveVec(before);
    const T kPrime = valley->getNormWaveVec(particle.energy);
    const T phi = std::atan2(particle.k[1], particle.k[0]);
    const T psi = sampleSingleLayerFroehlichDeflectionAngle(k, kPrime, effWidth, screeningWavevector, rng, uniform);
    particle.k = {kPrime * std::cos(phi + psi), kPrime * std::sin(phi + psi), 0};
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 8; // EMCGPU_SAMPLER_SINGLE_LAYER_FROEHLICH
    d.finalValley = this->idxValley;
    d.param[0] = Absorption ? phononEnergy : -phononEnergy;
    d.param[1] = effWidth;
    d.param[2] = screeningWavevector;
    return d;
  }
};

} // namespace emcdetail

template <class T> struct emcFroehlichInteractionAbsorptionSL : public emcdetail::SingleLayerFroehlich<T, true> {
  using emcdetail::SingleLayerFroehlich<T, true>::SingleLayerFroehlich;
};
template <class T> struct emcFroehlichInteractionEmissionSL : public emcdetail::SingleLayerFroehlich<T, false> {
  using emcdetail::SingleLayerFroehlich<T, false>::SingleLayerFroehlich;
};

#endif
