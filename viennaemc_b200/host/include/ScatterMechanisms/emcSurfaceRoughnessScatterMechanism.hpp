// Interface-roughness scattering (Prange-Nee / Ando) of a 2-D carrier gas: static potential e F Delta with a Gaussian
// autocorrelation of length Lambda, screened by the carriers; elastic.
// Interface mirrored: reference include/ScatterMechanisms/emcSurfaceRoughnessScatterMechanism.hpp (weight :59-63, ctor :68-79,
// rate :83-92, sampler :94-126).  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_SURFACE_ROUGHNESS.
#ifndef EMC_SURFACE_ROUGHNESS_SCATTER_MECHANISM_HPP
#define EMC_SURFACE_ROUGHNESS_SCATTER_MECHANISM_HPP

#include <cmath>
#include <random>
#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerAngle.hpp>
#include <emcConstants.hpp>

template <class T> class emcSurfaceRoughnessScatterMechanism : public emcScatterMechanism<T> {
  static constexpr SizeType angleSteps = 256;
  T prefactor; // (e F)^2 Delta^2 Lambda^2 / hbar^3
  T lambdaSquared;
  T screeningWavevector;
  std::string nameSuffix;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  T angularWeight(T theta, T k) const { // q = 2 k sin(theta / 2)
    const T q = 2 * k * std::sin(theta / 2);
    const T powerSpectrum = std::exp(-q * q * lambdaSquared / 4);
    return powerSpectrum * twoDScreeningFactor(q, screeningWavevector);
  }

public:
  emcSurfaceRoughnessScatterMechanism() = delete;
  // field normal to the sheet [V/m]; RMS roughness Delta [m]; correlation length Lambda [m]; 2-D screening wave vector [1/m]
  emcSurfaceRoughnessScatterMechanism(SizeType inValley, T effectiveField, T roughnessAmplitude, T correlationLength,
                                      T inScreeningWavevector, std::string inNameSuffix = "")
      : emcScatterMechanism<T>(inValley), lambdaSquared(correlationLength * correlationLength),
        screeningWavevector(inScreeningWavevector), nameSuffix(inNameSuffix) {
    const T force = constants::q * effectiveField;
    prefactor = force * force * roughnessAmplitude * roughnessAmplitude * lambdaSquared / std::pow(constants::hbar, 3);
  }

  std::string getName() const override { return "SurfaceRoughness" + nameSuffix; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    const auto *valley = this->ptrValley[this->idxValley];
    const T mc = valley->getEffMassCond(energy);
    const T k = valley->getNormWaveVec(energy);
    T integral = emcdetail::midpointAngularSum<angleSteps, T>([&](T theta) { return angularWeight(theta, k); });
    integral *= constants::pi / angleSteps;
    return prefactor * mc * integral;
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    const T k = this->ptrValley[this->idxValley]->getNormWaveVec(particle.energy);
    emcdetail::turnByWeightedAngle<angleSteps>(particle, rng, uniform, k, [&](T theta) { return angularWeight(theta, k); });
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 11; // EMCGPU_SAMPLER_SINGLE_LAYER_SURFACE_ROUGHNESS
    d.finalValley = this->idxValley;
    d.param[1] = lambdaSquared;
    d.param[2] = screeningWavevector;
    return d;
  }
};

#endif
