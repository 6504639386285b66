// Elastic scattering off a sheet of charged impurities at distance d of a 2-D semiconductor; statically screened Coulomb
// potential with the Rytova-Keldysh correction: |V(q)|^2 ~ (exp(-q d) / (q_s + q + r0 q^2))^2.
// Interface mirrored: reference include/ScatterMechanisms/emc2DChargedImpurityScatterMechanism.hpp (weight :65-72, ctor :77-92,
// rate :96-105, sampler :107-139).  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_CHARGED_IMPURITY.
#ifndef EMC_2D_CHARGED_IMPURITY_SCATTER_MECHANISM_HPP
#define EMC_2D_CHARGED_IMPURITY_SCATTER_MECHANISM_HPP

#include <cmath>
#include <random>
#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerAngle.hpp>
#include <emcConstants.hpp>

template <class T> class emc2DChargedImpurityScatterMechanism : public emcScatterMechanism<T> {
  static constexpr SizeType angleSteps = 512;
  T prefactor; // N_imp A^2 / (pi hbar^3), A = Z e^2 / (2 eps0 eps_avg)
  T screeningWavevector, rytovaKeldyshLength, remoteDistance;
  std::string nameSuffix;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  T angularWeight(T theta, T k) const { // q = 2 k sin(theta / 2)
    const T q = 2 * k * std::sin(theta / 2);
    const T denom = screeningWavevector + q + rytovaKeldyshLength * q * q;
    if (denom <= T(0))
      return T(0);
    const T v = std::exp(-q * remoteDistance) / denom;
    return v * v;
  }

public:
  emc2DChargedImpurityScatterMechanism() = delete;
  // areal impurity density [1/m^2]; average relative permittivity of the surroundings; 2-D screening wave vector [1/m];
  // Rytova-Keldysh length r0 [m] (0: plain Coulomb); impurity-to-sheet distance d [m]; charge number Z
  emc2DChargedImpurityScatterMechanism(SizeType inValley, T impurityDensity, T epsAvg, T inScreeningWavevector,
                                       T inRytovaKeldyshLength = 0, T inRemoteDistance = 0, T chargeNumber = 1,
                                       std::string inNameSuffix = "")
      : emcScatterMechanism<T>(inValley), screeningWavevector(inScreeningWavevector), rytovaKeldyshLength(inRytovaKeldyshLength),
        remoteDistance(inRemoteDistance), nameSuffix(inNameSuffix) {
    const T A = chargeNumber * constants::q * constants::q / (2 * constants::eps0 * epsAvg);
    prefactor = impurityDensity * A * A / (constants::pi * std::pow(constants::hbar, 3));
  }

  std::string getName() const override { return "ChargedImpurity2D" + nameSuffix; }

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
    d.samplerId = 10; // EMCGPU_SAMPLER_SINGLE_LAYER_CHARGED_IMPURITY
    d.finalValley = this->idxValley;
    d.param[0] = remoteDistance;
    d.param[1] = rytovaKeldyshLength;
    d.param[2] = screeningWavevector;
    return d;
  }
};

#endif
