// Zero-order (deformation potential D [eV/m]) intervalley phonon scattering.
// Interface mirrored: reference
// include/ScatterMechanisms/emcZeroOrderInterValleyScatterMechanism.hpp -- two classes
// (Absorption / Emission), each with a same-valley and a different-valley constructor
// (name suffix, initial valley[, final valley], final sub-valley map, D, phonon
// energy [eV], device).  Implementation: detail/emcInterValleyMechanism.hpp.
#ifndef EMC_ZERO_ORDER_INTERVALLEY_SCATTER_MECHANISM_HPP
#define EMC_ZERO_ORDER_INTERVALLEY_SCATTER_MECHANISM_HPP

#include <detail/emcInterValleyMechanism.hpp>

// prefactor of the rate without the final density of states (reference :12-21)
template <class T>
T getZeroOrderScatterConst(T defPot, T phEnergy, T rho, T temp, SizeType nrFValleys, bool isAbsorption = true) {
  const T c = nrFValleys * std::sqrt(constants::q) * std::pow(defPot / constants::hbar, 2) * constants::q /
              (constants::pi * rho * phEnergy * std::sqrt(2));
  const T n = emcdetail::phononOccupation(phEnergy, temp);
  return isAbsorption ? c * n : c * (n + 1);
}

#define EMC_DECLARE_INTERVALLEY(ClassName, Order, Absorption)                                                          \
  template <class T> class ClassName : public emcdetail::InterValleyMechanism<T, Order, Absorption> {                  \
    typedef emcdetail::InterValleyMechanism<T, Order, Absorption> Base;                                                \
                                                                                                                       \
  public:                                                                                                              \
    ClassName() = delete;                                                                                              \
    /* final valley == initial valley */                                                                               \
    template <class DeviceType>                                                                                        \
    ClassName(std::string inNameSuffix, SizeType inIdxValley,                                                          \
              std::map<SizeType, std::vector<SizeType>> inFinalSubValleys, T defPotential, T inPhononEnergy,          \
              const DeviceType &device)                                                                                \
        : Base(inNameSuffix, inIdxValley, inIdxValley, inFinalSubValleys, defPotential, inPhononEnergy, device) {}     \
    template <class DeviceType>                                                                                        \
    ClassName(std::string inNameSuffix, SizeType inIdxValley, SizeType inIdxFinalValley,                               \
              std::map<SizeType, std::vector<SizeType>> inFinalSubValleys, T defPotential, T inPhononEnergy,          \
              const DeviceType &device)                                                                                \
        : Base(inNameSuffix, inIdxValley, inIdxFinalValley, inFinalSubValleys, defPotential, inPhononEnergy,           \
               device) {}                                                                                              \
  }

EMC_DECLARE_INTERVALLEY(emcZeroOrderInterValleyAbsorptionScatterMechanism, 0, true);
EMC_DECLARE_INTERVALLEY(emcZeroOrderInterValleyEmissionScatterMechanism, 0, false);

#endif
