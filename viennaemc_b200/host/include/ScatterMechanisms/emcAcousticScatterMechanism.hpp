// Acoustic deformation-potential scattering (elastic, isotropic).
// Interface mirrored: reference include/ScatterMechanisms/emcAcousticScatterMechanism.hpp
// (ctors :30-56, rate :60-67, sampler :70-72).
// Device sampler: EMCGPU_SAMPLER_ISOTROPIC_ELASTIC.
#ifndef EMC_ACOUSTIC_SCATTER_MECHANISM_HPP
#define EMC_ACOUSTIC_SCATTER_MECHANISM_HPP

#include <cmath>
#include <random>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <emcConstants.hpp>

template <class T> class emcAcousticScatterMechanism : public emcScatterMechanism<T> {
  T prefactor; // sqrt(2q) (sigma q)^2 kB T / (pi rho vs^2 hbar^4)
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

public:
  emcAcousticScatterMechanism() = delete;

  // sigma: acoustic deformation potential [eV]; density and sound velocity from the device's material
  template <class DeviceType>
  emcAcousticScatterMechanism(SizeType inIdxValley, T sigma, const DeviceType &device)
      : emcAcousticScatterMechanism(inIdxValley, sigma, device.getMaterial().getRho(),
                                    device.getMaterial().getVelSound(), device) {}

  template <class DeviceType>
  emcAcousticScatterMechanism(SizeType inIdxValley, T sigma, T materialDensity, T velSound, const DeviceType &device)
      : emcScatterMechanism<T>(inIdxValley) {
    const T elasticConstant = materialDensity * std::pow(velSound, 2);
    prefactor = std::sqrt(2.0 * constants::q) * std::pow(sigma * constants::q, 2) * constants::kB *
                device.getTemperature() / (constants::pi * elasticConstant * pow(constants::hbar, 4));
  }

  std::string getName() const override { return "Acoustic"; }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    const auto *valley = this->ptrValley[this->idxValley];
    const T md = valley->getEffMassDOS();
    const T alpha = valley->getNonParabolicity();
    const T gamma = valley->getGamma(energy);
    return prefactor * pow(md, 3. / 2.) * std::sqrt(gamma) * (2 * alpha * energy + 1.0);
  }

  // g++ evaluates the two draws right to left: the first one becomes cos(theta)
  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    const T cosDraw = uniform(rng);
    const T phiDraw = uniform(rng);
    particle.k = initRandomDirection(norm(particle.k), phiDraw, cosDraw);
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 1; // EMCGPU_SAMPLER_ISOTROPIC_ELASTIC
    d.finalValley = this->idxValley;
    return d;
  }
};

#endif
