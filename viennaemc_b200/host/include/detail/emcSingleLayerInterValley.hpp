// Shared implementation of the four phonon-assisted intervalley (or optical intravalley) mechanisms of a single layer: zero /
// first order, absorption / emission (Kaasbjerg et al., PRB 85, 115317).  They differ in the prefactor, in the rate formula and
// in the sign of the phonon energy; zero order weights the final direction with the Herring-Vogt factors of the final valley,
// first order does not.
// Arithmetic mirrored (operation order kept: the tables must match bit for bit):
//   reference include/ScatterMechanisms/emcZeroOrderSingleLayerInterValleyScatterMechanism.hpp (ctors :42-88 / :219-266, rates
//   :96-108 / :274-286, samplers :111-147 / :289-324, check :150-181),
//   reference include/ScatterMechanisms/emcFirstOrderSingleLayerIntervalleyScatterMechanism.hpp (ctors :44-89 / :195-237, rates
//   :96-101 / :244-252, samplers :104-126 / :255-275).
// Device sampler: EMCGPU_SAMPLER_SINGLE_LAYER_INTERVALLEY with param[0] = the signed energy change, param[1] = 1 for first order.
#ifndef EMC_DETAIL_SINGLE_LAYER_INTERVALLEY_HPP
#define EMC_DETAIL_SINGLE_LAYER_INTERVALLEY_HPP

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

template <class T, int Order, bool Absorption> class SingleLayerInterValley : public emcScatterMechanism<T> {
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

  // sigma: deformation potential, [eV/m] zero order, [eV] first order; densityMaterial: sheet mass density [kg/m^2]
  SingleLayerInterValley(SizeType inValley, SizeType inFinalValley, T sigma, T densityMaterial, T temperature,
                         T inPhononEnergy, std::vector<std::vector<SizeType>> inFinalSubValleys, std::string inNameSuffix)
      : emcScatterMechanism<T>(inValley), phononEnergy(inPhononEnergy), idxFinalValley(inFinalValley),
        nameSuffix(inNameSuffix), finalSubValleys(std::move(inFinalSubValleys)) {
    nrFinal = finalSubValleys.empty() ? 1 : finalSubValleys[0].size();
    const T exponent = phononEnergy * constants::q / (constants::kB * temperature);
    const T omega = phononEnergy * constants::q / constants::hbar;
    if (Order == 0) {
      if (Absorption)
        prefactor = nrFinal * std::pow(sigma * constants::q / constants::hbar, 2) /
                    (2 * densityMaterial * omega * (std::exp(exponent) - 1));
      else
        prefactor = nrFinal * std::pow(sigma * constants::q / constants::hbar, 2) * std::exp(exponent) /
                    (2 * densityMaterial * omega * (std::exp(exponent) - 1));
    } else {
      if (Absorption)
        prefactor = nrFinal * std::pow(sigma * constants::q, 2) * constants::q /
                    (densityMaterial * omega * (std::exp(exponent) - 1) * std::pow(constants::hbar, 4));
      else
        prefactor = nrFinal * std::pow(sigma * constants::q, 2) * constants::q * std::exp(exponent) /
                    (densityMaterial * omega * (std::exp(exponent) - 1) * std::pow(constants::hbar, 4));
    }
  }

  std::string getName() const override {
    return std::string(Order == 0 ? "Zero" : "First") + "InterValley" + (Absorption ? "Absorption" : "Emission") + "SL" + nameSuffix;
  }

  T getScatterRate(T energy, SizeType /*idxRegion*/) const override {
    if (Order != 0) { // first order: mass and non-parabolicity of the INITIAL valley, threshold at the phonon energy
      if (!Absorption && !(energy > phononEnergy))
        return 0;
      const auto *from = this->ptrValley[this->idxValley];
      const T md = from->getEffMassDOS();
      const T alpha = from->getNonParabolicity();
      const T rate = md * md * prefactor * (Absorption ? 2 * energy + phononEnergy : 2 * energy - phononEnergy);
      return rate * (1 + 2 * alpha * energy);
    }
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
    if (Order == 0) {
      particle.k = singleLayerDirection(this->ptrValley[idxFinalValley], particle.energy, uniform(rng));
    } else { // no Herring-Vogt weighting, k_z left as it is
      const T phi = 2 * constants::pi * uniform(rng);
      const T kNorm = this->ptrValley[idxFinalValley]->getNormWaveVec(particle.energy);
      particle.k[0] = kNorm * std::cos(phi);
      particle.k[1] = kNorm * std::sin(phi);
    }
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
    d.param[1] = Order == 0 ? 0. : 1.; // first order: plain in-plane direction
    return d;
  }
};

} // namespace emcdetail

// the four public classes: the two constructors of the reference each (one valley with one sub-valley -- the sub-valley index
// is kept --, or a final valley with a map initial sub-valley -> candidates)
#define EMC_SINGLE_LAYER_INTERVALLEY_CLASS(NAME, ORDER, ABSORPTION)                                                             \
  template <class T> class NAME : public emcdetail::SingleLayerInterValley<T, ORDER, ABSORPTION> {                             \
    using Base = emcdetail::SingleLayerInterValley<T, ORDER, ABSORPTION>;                                                      \
                                                                                                                               \
  public:                                                                                                                      \
    NAME() = delete;                                                                                                           \
    NAME(SizeType inValley, T sigma, T densityMaterial, T temperature, T inPhononEnergy, std::string inNameSuffix = "")       \
        : Base(inValley, inValley, sigma, densityMaterial, temperature, inPhononEnergy, {}, inNameSuffix) {}                   \
    NAME(SizeType inValley, SizeType inFinalValley, T sigma, T densityMaterial, T temperature, T inPhononEnergy,              \
         std::vector<std::vector<SizeType>> inFinalSubValleys, std::string inNameSuffix = "")                                 \
        : Base(inValley, inFinalValley, sigma, densityMaterial, temperature, inPhononEnergy, std::move(inFinalSubValleys),     \
               inNameSuffix) {}                                                                                                \
  }

#endif
