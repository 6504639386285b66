// Intravalley optical-phonon scattering (zero-order deformation potential) of a 2-D carrier gas with the coupling divided by
// the static dielectric function of the carriers: the rate of the unscreened zero-order mechanism times the angular average of
// 1/eps(q)^2, final angle drawn from 1/eps(q)^2.  Valley and sub-valley do not change.
// Interface mirrored: reference include/ScatterMechanisms/emcScreenedIntravalleyOpticalMechanism.hpp (weight :58-62, ctor :67-81,
// rate :88-102, sampler :104-141).  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_SCREENED_OPTICAL.
#ifndef EMC_SCREENED_INTRAVALLEY_OPTICAL_MECHANISM_HPP
#define EMC_SCREENED_INTRAVALLEY_OPTICAL_MECHANISM_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <random>
#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerAngle.hpp>
#include <emcConstants.hpp>

template <class T> class emcScreenedIntravalleyOpticalMechanism : public emcScatterMechanism<T> {
  static constexpr SizeType angleSteps = 128;
  T phononEnergy;
  T prefactor; // (sigma e / hbar)^2 (N_q or N_q + 1) / (2 rho omega)
  T screeningWavevector;
  bool emission;
  std::string nameSuffix;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  T angularWeight(T theta, T k, T kFinal) const { // q^2 = k^2 + k'^2 - 2 k k' cos(theta)
    const T q = std::sqrt(std::max(T(0), k * k + kFinal * kFinal - 2 * k * kFinal * std::cos(theta)));
    return twoDScreeningFactor(q, screeningWavevector);
  }
  T finalEnergy(T energy) const { return emission ? energy - phononEnergy : energy + phononEnergy; }

public:
  emcScreenedIntravalleyOpticalMechanism() = delete;
  // optical deformation potential [eV/m]; sheet mass density [kg/m^2]; lattice temperature [K]; phonon energy [eV]
  emcScreenedIntravalleyOpticalMechanism(SizeType inValley, T sigma, T densityMaterial, T temperature, T inPhononEnergy,
                                         bool inEmission, T inScreeningWavevector = 0, std::string inNameSuffix = "")
      : emcScatterMechanism<T>(inValley), phononEnergy(inPhononEnergy), screeningWavevector(inScreeningWavevector),
        emission(inEmission), nameSuffix(inNameSuffix) {
    const T x = phononEnergy * constants::q / (constants::kB * temperature);
    const T omega = phononEnergy * constants::q / constants::hbar;
    const T occupation = 1. / (std::exp(x) - 1.);
    const T phonons = emission ? occupation + 1 : occupation;
    prefactor = std::pow(sigma * constants::q / constants::hbar, 2) * phonons / (2 * densityMaterial * omega);
  }

  std::string getName() const override { return std::string("ScreenedIntraOptical") + (emission ? "Em" : "Ab") + nameSuffix; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    if (emission && energy <= phononEnergy)
      return 0;
    const auto *valley = this->ptrValley[this->idxValley];
    const T after = finalEnergy(energy);
    const T md = valley->getEffMassDOS();
    const T alpha = valley->getNonParabolicity();
    const T k = valley->getNormWaveVec(energy);
    const T kFinal = valley->getNormWaveVec(after);
    T average = emcdetail::midpointAngularSum<angleSteps, T>([&](T theta) { return angularWeight(theta, k, kFinal); });
    average /= angleSteps;
    return md * prefactor * (1 + 2 * alpha * after) * average;
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    const auto *valley = this->ptrValley[this->idxValley];
    const T before = particle.energy;
    particle.energy = finalEnergy(before);
    assert(particle.energy > 0);
    const T k = valley->getNormWaveVec(before);
    const T kFinal = valley->getNormWaveVec(particle.energy);
    emcdetail::turnByWeightedAngle<angleSteps>(particle, rng, uniform, kFinal, [&](T theta) { return angularWeight(theta, k, kFinal); });
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 13; // EMCGPU_SAMPLER_SINGLE_LAYER_SCREENED_OPTICAL
    d.finalValley = this->idxValley;
    d.param[0] = emission ? -phononEnergy : phononEnergy;
    d.param[2] = screeningWavevector;
    return d;
  }
};

#endif
