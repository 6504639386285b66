// Ionised-impurity scattering, Brooks-Herring model.
// Interface mirrored: reference include/ScatterMechanisms/emcCoulombScatterMechanism.hpp
// (ctor :23-32, rate :36-46, sampler :48-59).  The rate depends on the doping of the
// region the table is built for; the device sampler (EMCGPU_SAMPLER_COULOMB) gets
// that region's Debye energy N_D/(c2 m_c) as param[0].
#ifndef EMC_COULOMB_SCATTER_MECHANISM_HPP
#define EMC_COULOMB_SCATTER_MECHANISM_HPP

#include <cmath>
#include <random>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <emcConstants.hpp>

template <class T, class DeviceType> class emcCoulombScatterMechanism : public emcScatterMechanism<T> {
  DeviceType &device;
  T c1; // sqrt(2q) (kB T)^2 / (pi hbar^4)
  T c2; // 8 eps Vt / hbar^2
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  T impurityDensity(SizeType idxRegion) const { return std::fabs(device.getDopingProfile().getDoping(idxRegion)); }

public:
  emcCoulombScatterMechanism() = delete;
  emcCoulombScatterMechanism(SizeType inIdxValley, T epsR, DeviceType &inDevice)
      : emcScatterMechanism<T>(inIdxValley), device(inDevice) {
    const T epsMat = constants::eps0 * epsR;
    const T temperature = device.getTemperature();
    const T Vt = device.getThermalVoltage();
    c1 = std::sqrt(2 * constants::q) * std::pow(constants::kB * temperature, 2) /
         (constants::pi * pow(constants::hbar, 4));
    c2 = 8 * epsMat * Vt / (constants::hbar * constants::hbar);
  }

  std::string getName() const override { return "Coulomb"; }

  T getScatterRate(T energy, SizeType idxRegion) const override {
    const T nd = impurityDensity(idxRegion);
    const auto *valley = this->ptrValley[this->idxValley];
    const T md = valley->getEffMassDOS();
    const T mc = valley->getEffMassCond();
    const T alpha = valley->getNonParabolicity();
    const T gamma = valley->getGamma(energy);
    return c1 * pow(md, 3. / 2.) / nd * std::sqrt(gamma) * (2 * alpha * energy + 1.0) / (1 + (c2 * mc * gamma / nd));
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    const auto *valley = this->ptrValley[this->idxValley];
    const T gamma = valley->getGamma(particle.energy);
    const T debyeEnergy = impurityDensity(particle.region) / (c2 * valley->getEffMassCond());
    const T r = uniform(rng);
    const T cosTheta = 1.0 - r * 2.0 / ((1 - r) * gamma / debyeEnergy + 1.0);
    particle.k = initRandomDirectionWithRespectToCurrentK(particle.k, cosTheta, uniform(rng));
  }

  emcDeviceSamplerDesc deviceSampler(SizeType idxRegion) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 3; // EMCGPU_SAMPLER_COULOMB
    d.finalValley = this->idxValley;
    d.param[0] = impurityDensity(idxRegion) / (c2 * this->ptrValley[this->idxValley]->getEffMassCond());
    return d;
  }
};

#endif
