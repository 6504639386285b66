// Zero-order intervalley (or optical intravalley) phonon scattering in a single layer, absorption and emission
// (Kaasbjerg et al., PRB 85, 115317).  Interface mirrored: reference
// include/ScatterMechanisms/emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp (ctors :42-88 / :219-266, rates :96-108 /
// :274-286, samplers :111-147 / :289-324, check :150-181).  One implementation for both classes; they differ in the Bose
// factor of the prefactor and in the sign of the phonon energy.
// Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY with param[0] = the signed energy change.
#ifndef EMC_ZERO_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP
#define EMC_ZERO_ORDER_SINGLE_LAYER_INTERVALLEY_SCATTER_MECHANISM_HPP

#include <cassert>
#include <cmath>
#include <map>
#include <random>
#include <string>
#include <vector>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <detail/emcSingleLayerDirection.hpp>
#include <emcConstants.hpp>
#include <emcMessage.hpp>

namespace emcdetail {

template <class T, bool Absorption> class SingleLayerInterValley : public emcScatterMechanism<T> {
  T prefactor;
  T phononEnergy;
  SizeType idxFinalValley;
  SizeType nrFinal;
  std::string nameSuffix;
  std::vector<std::vector<SizeType>> finalSubValleys; // [initial sub-valley] -> candidates; empty: sub-valley kept
  mutable std::uniform_real_distribution<T> uniform{0., 1.};

  T bottomDifference() const {
    return this->ptrValley[idxFinalValley]->getBottomEnergy() - this->ptrValley[this->idxValley]->getBottomEnergy();
  }

public:
  SingleLayerInterValley() = delete;

  // sigma: deformation potential [eV/m]; densityMaterial: sheet mass density [kg/m^2]
  SingleLayerInterValley(SizeType inValley, SizeType inFinalValley, T sigma, T densityMaterial, T temperature,
                         T inPhononEnergy, std::vector<std::vector<SizeType>> inFinalSubValleys, std::string inNameSuffix)
      : emcScatterMechanism<T>(inValley), phononEnergy(inPhononEnergy), idxFinalValley(inFinalValley),
        nameSuffix(inNameSuffix), finalSubValleys(std::move(inFinalSubValleys)) {
    nrFinal = finalSubValleys.empty() ? 1 : finalSubValleys[0].size();
    const T exponent = phononEnergy * constants::q / (constants::kB * temperature);
    const T omega = phononEnergy * constants::q / constants::hbar;
    if (Absorption)
      prefactor = nrFinal * std::pow(sigma * constants::q / constants::hbar, 2) /
                  (2 * densityMaterial * omega * (std::exp(exponent) - 1));
    else
      prefactor = nrFinal * std::pow(sigma * constants::q / constants::hbar, 2) * std::exp(exponent) /
                  (2 * densityMaterial * omega * (std::exp(exponent) - 1));
  }

  std::string getName() const override {
    return std::string("ZeroInterValley") + (Absorption ? "Absorption" : "Emission") + "SL" + nameSuffix;
  }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    const auto *to = this->ptrValley[idxFinalValley];
    const T shift = bottomDifference();
    const T finalEnergy = Absorption ? energy + phononEnergy - shift : energy - phononEnergy - shift;
    if (finalEnergy > 0) {
      const T md = to->getEffMassDOS();
      const T alpha = to->getNonParabolicity();
      return md * prefactor * (1 + 2 * alpha * finalEnergy);
    }
    return 0;
  }

  void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const override {
    particle.valley = idxFinalValley;
    if (!finalSubValleys.empty())
      particle.subValley = finalSubValleys[particle.subValley][std::floor(uniform(rng) * nrFinal)];
    if (Absorption)
      particle.energy += phononEnergy - bottomDifference();
    else
      particle.energy -= (bottomDifference() + phononEnergy);
    particle.k = singleLayerDirection(this->ptrValley[idxFinalValley], particle.energy, uniform(rng));
  }

  void check() final {
    auto &msg = emcMessage::getInstance();
    if (idxFinalValley >= this->ptrValley.size())
      msg.addError(getName() + ": idxFinalValley " + std::to_string(idxFinalValley) + " is not valid.").print();
    const SizeType degInitial = this->ptrValley[this->idxValley]->getDegeneracyFactor();
    const SizeType degFinal = this->ptrValley[idxFinalValley]->getDegeneracyFactor();
    if (finalSubValleys.empty())
      return;
    for (SizeType s = 0; s < degInitial; s++) {
      if (finalSubValleys.at(s).size() != nrFinal)
        msg.addWarning(getName() + ": Nr. of final subvalleys not consistent.").print();
      for (auto f : finalSubValleys.at(s))
        if (f >= degFinal)
          msg.addError(getName() + ": Used idx " + std::to_string(f) + " for valley of degeneracy " +
                       std::to_string(degFinal) + " is not valid.")
              .print();
    }
  }

  emcDeviceSamplerDesc deviceSampler(SizeType) const override {
    emcDeviceSamplerDesc d;
    d.samplerId = 7; // EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY
    d.finalValley = idxFinalValley;
    for (SizeType s = 0; s < finalSubValleys.size(); s++)
      d.finalSubValleys[s] = finalSubValleys[s];
    d.param[0] = Absorption ? (phononEnergy - bottomDifference()) : -(bottomDifference() + phononEnergy);
    return d;
  }
};

} // namespace emcdetail

template <class T>
class emcZeroOrderSingleLayerInterValleyAbsorptionScatterMechanism : public emcdetail::SingleLayerInterValley<T, true> {
  using Base = emcdetail::SingleLayerInterValley<T, true>;

public:
  emcZeroOrderSingleLayerInterValleyAbsorptionScatterMechanism() = delete;
  // one valley with one sub-valley (comparison with Kaasbjerg et al.): the sub-valley index is kept
  emcZeroOrderSingleLayerInterValleyAbsorptionScatterMechanism(SizeType inValley, T sigma, T densityMaterial, T temperature,
                                                               T inPhononEnergy, std::string inNameSuffix = "")
      : Base(inValley, inValley, sigma, densityMaterial, temperature, inPhononEnergy, {}, inNameSuffix) {}
  emcZeroOrderSingleLayerInterValleyAbsorptionScatterMechanism(SizeType inValley, SizeType inFinalValley, T sigma,
                                                               T densityMaterial, T temperature, T inPhononEnergy,
                                                               std::vector<std::vector<SizeType>> inFinalSubValleys,
                                                               std::string inNameSuffix = "")
      : Base(inValley, inFinalValley, sigma, densityMaterial, temperature, inPhononEnergy, std::move(inFinalSubValleys),
             inNameSuffix) {}
};

template <class T>
class emcZeroOrderSingleLayerInterValleyEmissionScatterMechanism : public emcdetail::SingleLayerInterValley<T, false> {
  using Base = emcdetail::SingleLayerInterValley<T, false>;

public:
  emcZeroOrderSingleLayerInterValleyEmissionScatterMechanism() = delete;
  emcZeroOrderSingleLayerInterValleyEmissionScatterMechanism(SizeType inValley, T sigma, T densityMaterial, T temperature,
                                                             T inPhononEnergy, std::string inNameSuffix = "")
      : Base(inValley, inValley, sigma, densityMaterial, temperature, inPhononEnergy, {}, inNameSuffix) {}
  emcZeroOrderSingleLayerInterValleyEmissionScatterMechanism(SizeType inValley, SizeType inFinalValley, T sigma,
                                                             T densityMaterial, T temperature, T inPhononEnergy,
                                                             std::vector<std::vector<SizeType>> inFinalSubValleys,
                                                             std::string inNameSuffix = "")
      : Base(inValley, inFinalValley, sigma, densityMaterial, temperature, inPhononEnergy, std::move(inFinalSubValleys),
             inNameSuffix) {}
};

#endif
