// First-order intervalley phonon scattering.
// Interface mirrored: reference
// include/ScatterMechanisms/emcFirstOrderInterValleyScatterMechanism.hpp -- classes
// emcFirstOrderInterValley{Absorption,Emission}ScatterMechanism with the same two
// constructors as the zero-order pair.  Implementation: detail/emcInterValleyMechanism.hpp.
#ifndef EMC_FIRST_ORDER_INTERVALLEY_SCATTER_MECHANISM_HPP
#define EMC_FIRST_ORDER_INTERVALLEY_SCATTER_MECHANISM_HPP

#include <ScatterMechanisms/emcZeroOrderInterValleyScatterMechanism.hpp>

// prefactor of the rate (reference :12-20)
template <class T>
T getFirstOrderScatterConst(T defPot, T phEnergy, T rho, T temp, SizeType nrFValleys, bool isAbsorption = true) {
  const T c = nrFValleys * std::sqrt(2) * pow(constants::q, 5. / 2.) * pow(defPot, 2) /
              (constants::pi * rho * pow(constants::hbar, 4) * phEnergy);
  const T n = emcdetail::phononOccupation(phEnergy, temp);
  return isAbsorption ? c * n : c * (n + 1);
}

EMC_DECLARE_INTERVALLEY(emcFirstOrderInterValleyAbsorptionScatterMechanism, 1, true);
EMC_DECLARE_INTERVALLEY(emcFirstOrderInterValleyEmissionScatterMechanism, 1, false);

#endif
