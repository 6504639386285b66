// Shared implementation of the four phonon-assisted intervalley mechanisms
// (zero / first order, absorption / emission).  They differ in the prefactor,
// in the energy exchanged with the lattice and, for first order, in one extra
// factor of the rate; the final state is the same: the carrier lands in one of
// the listed final sub-valleys with E' = E +- hw - (E_bottom,f - E_bottom,i) and
// an isotropic direction.
//
// Arithmetic mirrored (operation order kept: the tables must match bit for bit):
//   reference include/ScatterMechanisms/emcZeroOrderInterValleyScatterMechanism.hpp:12-21, :103-129, :246-272
//   reference include/ScatterMechanisms/emcFirstOrderInterValleyScatterMechanism.hpp:12-20, :102-131, :250-279
// Device sampler: EMCGPU_SAMPLER_INTERVALLEY with param[0] = the signed energy change.
#ifndef EMC_DETAIL_INTERVALLEY_MECHANISM_HPP
#define EMC_DETAIL_INTERVALLEY_MECHANISM_HPP

#include <cmath>
#include <map>
#include <random>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <emcConstants.hpp>
#include <emcMessage.hpp>

namespace emcdetail {

// Bose-Einstein occupation of a phonon of energy phEnergy [eV] at temp [K]
template <class T> T phononOccupation(T phEnergy, T temp) {
  return 1. / (std::exp(constants::q * phEnergy / (constants::kB * temp)) - 1.);
}

template <class T, int Order, bool Absorption> class InterValleyMechanism : public emcScatterMechanism<T> {
  std::string nameSuffix;
  T phononEnergy;
  T prefactor;
  SizeType nrFinal;
  SizeType idxFinalValley;
  std::map<SizeType, std::vector<SizeType>> finalSubValleys;
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  T bottomDifference() const {
    return this->ptrValley[idxFinalValley]->getBottomEnergy() - this->ptrValley[this->idxValley]->getBottomEnergy();
  }
  // energy after the event minus energy before it
  T energyChange() const {
    return Absorption ? (phononEnergy - bottomDifference()) : -(phononEnergy + bottomDifference());
  }

public:
  InterValleyMechanism() = delete;

  template <class DeviceType>
  InterValleyMechanism(std::string inNameSuffix, SizeType inIdxValley, SizeType inIdxFinalValley,
                       std::map<SizeType, std::vector<SizeType>> inFinalSubValleys, T defPotential, T inPhononEnergy,
                       const DeviceType &device)
      : emcScatterMechanism<T>(inIdxValley), nameSuffix(inNameSuffix), phononEnergy(inPhononEnergy),
        nrFinal(inFinalSubValleys.at(0).size()), idxFinalValley(inIdxFinalValley),
        finalSubValleys(inFinalSubValleys) {
    const T rho = device.getMaterial().getRho();
    if (Order == 0)
      prefactor = nrFinal * std::sqrt(constants::q) * std::pow(defPotential / constants::hbar, 2) * constants::q /
                  (constants::pi * rho * phononEnergy * std::sqrt(2));
    else
      prefactor = nrFinal * std::sqrt(2) * pow(constants::q, 5. / 2.) * pow(defPotential, 2) /
                  (constants::pi * rho * pow(constants::hbar, 4) * phononEnergy);
    const T nPhonon = phononOccupation(phononEnergy, T(device.getTemperature()));
    prefactor = Absorption ? prefactor * nPhonon : prefactor * (nPhonon + 1);
  }

  std::string getName() const override {
    return std::string(Order == 0 ? "Zero" : "First") + "InterValley" + (Absorption ? "Absorption" : "Emission") +
           nameSuffix;
  }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    const auto *from = this->ptrValley[this->idxValley];
    const auto *to = this->ptrValley[idxFinalValley];
    const T shift = bottomDifference();
    const T finalEnergy = Absorption ? energy + phononEnergy - shift : energy - phononEnergy - shift;
    if (!(finalEnergy > 0))
      return 0;
    const T md = to->getEffMassDOS();
    const T alpha = to->getNonParabolicity();
    if (Order == 0) {
      const T gammaFinal = to->getGamma(finalEnergy);
      return prefactor * std::pow(md, 3. / 2.) * std::sqrt(gammaFinal) * (2 * alpha * finalEnergy + 1.0);
    }
    const T gammaInitial = from->getGamma(energy);
    const T gammaFinal = to->getGamma(finalEnergy);
    return prefactor * pow(md, 5. / 2.) * std::sqrt(gammaFinal) * (2 * alpha * finalEnergy + 1.0) *
           (gammaInitial + gammaFinal);
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    particle.valley = idxFinalValley;
    particle.subValley = finalSubValleys.at(particle.subValley)[rng() % nrFinal];
    if (Absorption)
      particle.energy += (phononEnergy - bottomDifference());
    else
      particle.energy -= (phononEnergy + bottomDifference());
    const T kNew = this->ptrValley[idxFinalValley]->getNormWaveVec(particle.energy);
    const T cosDraw = uniform(rng); // first draw -> cos(theta), second -> phi (g++ argument order)
    const T phiDraw = uniform(rng);
    particle.k = initRandomDirection(kNew, phiDraw, cosDraw);
  }

  void check() final {
    auto &msg = emcMessage::getInstance();
    if (idxFinalValley >= this->ptrValley.size())
      msg.addError(getName() + ": idxFinalValley " + std::to_string(idxFinalValley) + " is not valid.").print();
    const SizeType degInitial = this->ptrValley[this->idxValley]->getDegeneracyFactor();
    const SizeType degFinal = this->ptrValley[idxFinalValley]->getDegeneracyFactor();
    for (SizeType s = 0; s < degInitial; s++) {
      auto it = finalSubValleys.find(s);
      if (it == finalSubValleys.end())
        msg.addError(getName() + ": no final subvalleys given for initial subvalley " + std::to_string(s) + ".")
            .print();
      if (it->second.size() != nrFinal)
        msg.addWarning(getName() + ": Nr. of final subvalleys not consistent.").print();
      for (auto f : it->second)
        if (f >= degFinal)
          msg.addError(getName() + ": final subvalley " + std::to_string(f) + " does not exist.").print();
    }
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 2; // EMCGPU_SAMPLER_INTERVALLEY
    d.finalValley = idxFinalValley;
    d.finalSubValleys = finalSubValleys;
    d.param[0] = energyChange();
    return d;
  }
};

} // namespace emcdetail

#endif
