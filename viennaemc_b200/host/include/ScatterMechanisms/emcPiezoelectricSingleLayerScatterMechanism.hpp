// Piezoelectric acoustic-phonon scattering in a single layer of a 2-D semiconductor without inversion centre (Kaasbjerg,
// Thygesen & Jauho, PRB 87, 235312): quasi-elastic, long-range, optionally screened by the 2-D carrier gas.
// Interface mirrored: reference include/ScatterMechanisms/emcPiezoelectricSingleLayerScatterMechanism.hpp (weight :62-66, ctor
// :71-86, rate :92-105, sampler :110-139).  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_PIEZOELECTRIC.
#ifndef PIEZOELECTRIC_SINGLE_LAYER_SCATTER_MECHANISM_HPP
#define PIEZOELECTRIC_SINGLE_LAYER_SCATTER_MECHANISM_HPP

#include <cmath>
#include <random>
#include <string>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerAngle.hpp>
#include <emcConstants.hpp>

template <class T> class emcPiezoelectricSingleLayerMechanism : public emcScatterMechanism<T> {
  T effWidth;
  T prefactor; // (e11 q / eps0)^2 <A^2> kB T / (rho c^2 hbar^3), <A^2> = 1/2
  T screeningWavevector;
  std::string nameSuffix;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  // weight of a deflection theta of an elastic event at |k|: momentum transfer q = 2 k sin(theta / 2)
  T angularWeight(T theta, T k) const {
    return emcdetail::formFactorScreened(2 * k * std::sin(theta / 2), effWidth, screeningWavevector);
  }

public:
  emcPiezoelectricSingleLayerMechanism() = delete;
  // piezoConst e11 [C/m]; width of the wave functions [m]; sheet mass density [kg/m^2]; sound velocity of the branch [m/s]
  emcPiezoelectricSingleLayerMechanism(SizeType inValley, T piezoConst, T inEffWidth, T densityMaterial, T velSound, T temperature,
                                       std::string inNameSuffix = "", T inScreeningWavevector = 0)
      : emcScatterMechanism<T>(inValley), effWidth(inEffWidth), screeningWavevector(inScreeningWavevector),
        nameSuffix(inNameSuffix) {
    const T couplingEnergy = piezoConst * constants::q / constants::eps0;
    prefactor = 0.5 * couplingEnergy * couplingEnergy * constants::kB * temperature /
                (densityMaterial * velSound * velSound * std::pow(constants::hbar, 3));
  }

  std::string getName() const override { return "PiezoelectricSL" + nameSuffix; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    const auto *valley = this->ptrValley[this->idxValley];
    const T md = valley->getEffMassDOS();
    const T alpha = valley->getNonParabolicity();
    const T k = valley->getNormWaveVec(energy);
    // (1/pi) int_0^pi weight dtheta by the midpoint rule; -> 1 for long waves without screening
    const T dtheta = constants::pi / emcdetail::singleLayerAngleSteps;
    T integral = 0;
    for (SizeType i = 0; i < emcdetail::singleLayerAngleSteps; ++i)
      integral += angularWeight((i + T(0.5)) * dtheta, k);
    integral *= dtheta / constants::pi;
    return md * prefactor * (1 + 2 * alpha * energy) * integral;
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    const T k = this->ptrValley[this->idxValley]->getNormWaveVec(particle.energy);
    const auto sums = emcdetail::cumulativeAngularWeight<T>([&](T theta) { return angularWeight(theta, k); });
    const T total = sums[emcdetail::singleLayerAngleSteps];
    const T phi = std::atan2(particle.k[1], particle.k[0]);
    T theta = (total > T(0)) ? emcdetail::invertAngularWeight(sums, uniform(rng) * total) : constants::pi * uniform(rng);
    if (uniform(rng) < T(0.5))
      theta = -theta;
    particle.k = {k * std::cos(phi + theta), k * std::sin(phi + theta), 0};
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 9; // EMCGPU_SAMPLER_SINGLE_LAYER_PIEZOELECTRIC
    d.finalValley = this->idxValley;
    d.param[1] = effWidth;
    d.param[2] = screeningWavevector;
    return d;
  }
};

#endif
