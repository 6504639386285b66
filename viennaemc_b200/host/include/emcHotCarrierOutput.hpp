// Carrier-temperature and device-metric estimators of the hot-carrier example (host-side post-processing of the
// ensemble averages; nothing here touches particles).
// Interface mirrored: reference include/emcHotCarrierOutput.hpp -- fermiIntegral :73-85, getMBTemp :89-92, FDFitResult
// :97-103, getFDFit :115-180, getVOC :196-209, getPCE :221-224.
#ifndef EMC_HOT_CARRIER_OUTPUT_HPP
#define EMC_HOT_CARRIER_OUTPUT_HPP

#include <algorithm>
#include <cmath>

#include <emcConstants.hpp>

namespace emcHotCarrierOutput {

// F_j(eta) = int_0^inf x^j / (1 + exp(x - eta)) dx by the midpoint rule on [0, max(10, eta + 30)], 1200 intervals
template <class T> T fermiIntegral(T j, T eta) {
  const T xMax = std::max(T(10), eta + T(30));
  const int n = 1200;
  const T dx = xMax / T(n);
  T sum = T(0);
  for (int i = 0; i < n; i++) {
    const T x = (T(i) + T(0.5)) * dx;
    sum += std::pow(x, j) / (T(1) + std::exp(std::min(x - eta, T(700))));
  }
  return sum * dx;
}

// equipartition: <E> = 3/2 kB T
template <class T> T getMBTemp(T avgEnergyEV) { return T(2) * avgEnergyEV * T(constants::q) / (T(3) * T(constants::kB)); }

template <class T> struct FDFitResult {
  T T_FD;  // carrier temperature [K]
  T mu_eV; // chemical potential [eV]
  T eta;   // mu / (kB T_FD)
  bool converged;
};

// Fermi-Dirac fit of a 3-D parabolic band from density n and mean energy <E>:
//   n = (2 m kT / hbar^2)^{3/2} F_1/2(eta) / (2 pi^2),  <E> = kT F_3/2(eta) / F_1/2(eta)
// eliminate T:  F_3/2^{3/2} / F_1/2^{5/2} = (2 m <E> / hbar^2)^{3/2} / (2 pi^2 n)  -> eta by bisection on [-60, 60]
// (the left side falls monotonically with eta), then kT = <E> F_1/2 / F_3/2.
template <class T> FDFitResult<T> getFDFit(T avgEnergyEV, T n, T effMassRel, T latTempK = T(300)) {
  const T mass = effMassRel * T(constants::me);
  const T energyJ = avgEnergyEV * T(constants::q);
  const T target = std::pow(T(2) * mass * energyJ / (T(constants::hbar) * T(constants::hbar)), T(1.5)) /
                   (T(2) * T(constants::pi) * T(constants::pi)) / n;
  const auto lhs = [](T eta) {
    return std::pow(fermiIntegral<T>(T(1.5), eta), T(1.5)) / std::pow(fermiIntegral<T>(T(0.5), eta), T(2.5));
  };
  T lo = T(-60), hi = T(60);
  bool converged = false;
  for (int it = 0; it < 80 && !converged; it++) {
    const T mid = T(0.5) * (lo + hi);
    if (lhs(mid) > target)
      lo = mid;
    else
      hi = mid;
    converged = hi - lo < T(1e-5);
  }
  const T eta = T(0.5) * (lo + hi);
  const T kT = energyJ * fermiIntegral<T>(T(0.5), eta) / fermiIntegral<T>(T(1.5), eta);
  T temp = kT / T(constants::kB);
  if (temp < latTempK)
    temp = latTempK;
  return {temp, eta * kT / T(constants::q), eta, converged};
}

// q V_OC = dMu (T_L / T_eh) + dE (1 - T_L / T_eh),  dMu = mu_e + mu_h + E_gap,  dE = <E_e> + <E_h> + E_gap
template <class T> T getVOC(T muE, T muH, T tempE, T tempH, T tempLattice, T gapEV, T avgEnergyE, T avgEnergyH) {
  const T tEH = T(0.5) * (tempE + tempH);
  const T ratio = tEH > T(0) ? tempLattice / tEH : T(1);
  return (muE + muH + gapEV) * ratio + (avgEnergyE + avgEnergyH + gapEV) * (T(1) - ratio);
}

// J [mA/cm^2] x V_OC [V] x FF / P_sun [mW/cm^2]
template <class T> T getPCE(T currentDensity, T voc, T fillFactor = T(0.85), T sunPower = T(100)) {
  return currentDensity * voc * fillFactor / sunPower;
}

} // namespace emcHotCarrierOutput

#endif
