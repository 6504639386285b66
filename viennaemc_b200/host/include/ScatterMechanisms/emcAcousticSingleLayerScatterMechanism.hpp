// Intravalley acoustic-phonon scattering in a single layer (elastic; Kaasbjerg et al., PRB 85, 115317).
// Interface mirrored: reference include/ScatterMechanisms/emcAcousticSingleLayerScatterMechanism.hpp (ctor :42-50, rate
// :55-60, sampler :63-81).  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_ELASTIC.
#ifndef ACOUSTIC_SINGLE_LAYER_SCATTER_MECHANISM_HPP
#define ACOUSTIC_SINGLE_LAYER_SCATTER_MECHANISM_HPP

#include <cmath>
#include <random>
#include <string>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerDirection.hpp>
#include <emcConstants.hpp>

template <class T> class emcAcousticSingleLayerMechanism : public emcScatterMechanism<T> {
  T prefactor; // (sigma q)^2 kB T / (rho vs^2 hbar^3)
  std::string nameSuffix;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

public:
  emcAcousticSingleLayerMechanism() = delete;

  // sigma: deformation potential [eV]; densityMaterial: sheet mass density [kg/m^2]; velSound [m/s]
  emcAcousticSingleLayerMechanism(SizeType inValley, T sigma, T densityMaterial, T velSound, T temperature,
                                  std::string inNameSuffix = "")
      : emcScatterMechanism<T>(inValley), nameSuffix(inNameSuffix) {
    const T elasticConstant = densityMaterial * std::pow(velSound, 2);
    prefactor = std::pow(sigma * constants::q, 2) * constants::kB * temperature /
                (elasticConstant * std::pow(constants::hbar, 3));
  }

  std::string getName() const override { return "AcousticSL" + nameSuffix; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    const auto *valley = this->ptrValley[this->idxValley];
    const T md = valley->getEffMassDOS();
    const T alpha = valley->getNonParabolicity();
    return md * prefactor * (1 + 2 * alpha * energy);
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    particle.k = emcdetail::singleLayerDirection(this->ptrValley[particle.valley], particle.energy, uniform(rng));
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 6; // EMCGPU_SAMPLER_SINGLE_LAYER_ELASTIC
    d.finalValley = this->idxValley;
    return d;
  }
};

#endif
