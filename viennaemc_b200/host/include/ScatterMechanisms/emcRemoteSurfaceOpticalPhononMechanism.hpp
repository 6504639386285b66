// Remote surface-optical phonons of a polar substrate or gate dielectric at distance d (Froehlich-like coupling
// ~ exp(-2 q d) / q), optionally screened by the 2-D carrier gas; one object per branch (absorption or emission).
// Interface mirrored: reference include/ScatterMechanisms/emcRemoteSurfaceOpticalPhononMechanism.hpp (weight :63-70, ctor
// :75-91, rate :97-110, sampler :112-149).  Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_REMOTE_SO.
#ifndef EMC_REMOTE_SURFACE_OPTICAL_PHONON_MECHANISM_HPP
#define EMC_REMOTE_SURFACE_OPTICAL_PHONON_MECHANISM_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <random>
#include <string>

#include <ScatterMechanisms/emc2DScreening.hpp>
#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerAngle.hpp>
#include <emcConstants.hpp>

template <class T> class emcRemoteSurfaceOpticalPhononMechanism : public emcScatterMechanism<T> {
  static constexpr SizeType angleSteps = 128;
  T phononEnergy, remoteDistance;
  T prefactor; // e^2 omega D / (4 pi eps0 hbar^2) x (N_q or N_q + 1)
  T screeningWavevector;
  bool emission;
  std::string nameSuffix;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  T angularWeight(T theta, T k, T kFinal) const { // q^2 = k^2 + k'^2 - 2 k k' cos(theta)
    const T q = std::sqrt(std::max(T(0), k * k + kFinal * kFinal - 2 * k * kFinal * std::cos(theta)));
    if (q <= T(0))
      return T(0);
    return std::exp(-2 * q * remoteDistance) * twoDScreeningFactor(q, screeningWavevector) / q;
  }
  T finalEnergy(T energy) const { return emission ? energy - phononEnergy : energy + phononEnergy; }

public:
  emcRemoteSurfaceOpticalPhononMechanism() = delete;
  // phonon energy [eV]; dielectric step D (dimensionless); carrier-to-surface distance [m]; lattice temperature [K]
  emcRemoteSurfaceOpticalPhononMechanism(SizeType inValley, T inPhononEnergy, T couplingD, T inRemoteDistance, T temperature,
                                         bool inEmission, T inScreeningWavevector = 0, std::string inNameSuffix = "")
      : emcScatterMechanism<T>(inValley), phononEnergy(inPhononEnergy), remoteDistance(inRemoteDistance),
        screeningWavevector(inScreeningWavevector), emission(inEmission), nameSuffix(inNameSuffix) {
    const T omega = phononEnergy * constants::q / constants::hbar;
    const T coupling = constants::q * constants::q * omega * couplingD / (4 * constants::pi * constants::eps0 * constants::hbar * constants::hbar);
    const T x = phononEnergy * constants::q / (constants::kB * temperature);
    const T occupation = 1. / (std::exp(x) - 1.);
    prefactor = coupling * (emission ? occupation + 1 : occupation);
  }

  std::string getName() const override { return std::string("RemoteSO") + (emission ? "Em" : "Ab") + nameSuffix; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    if (emission && energy <= phononEnergy)
      return 0;
    const auto *valley = this->ptrValley[this->idxValley];
    const T after = finalEnergy(energy);
    const T k = valley->getNormWaveVec(energy);
    const T kFinal = valley->getNormWaveVec(after);
    const T mc = valley->getEffMassCond(after);
    T integral = emcdetail::midpointAngularSum<angleSteps, T>([&](T theta) { return angularWeight(theta, k, kFinal); });
    integral *= constants::pi / angleSteps;
    return 2 * prefactor * mc * integral;
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
    d.samplerId = 12; // EMCGPU_SAMPLER_SINGLE_LAYER_REMOTE_SO
    d.finalValley = this->idxValley;
    d.param[0] = emission ? -phononEnergy : phononEnergy;
    d.param[1] = remoteDistance;
    d.param[2] = screeningWavevector;
    return d;
  }
};

#endif
