// Static free-carrier screening of a 2-D carrier gas: the screening wave vector the single-layer parameter sets pass to their
// long-range mechanisms.  Interface mirrored: reference include/ScatterMechanisms/emc2DScreening.hpp:49-76
//   q_s = q^2 (dn/dmu) / (2 eps0 eps_r),  dn/dmu = D0 (1 - exp(-n / (D0 kB T))),  D0 = g m / (2 pi hbar^2)
//   eps(q) = 1 + q_s / q.
#ifndef EMC_2D_SCREENING_HPP
#define EMC_2D_SCREENING_HPP

#include <cmath>

#include <emcConstants.hpp>

// carrierDensity [1/m^2], temperature [K], relative permittivity of the environment, density-of-states mass [kg],
// g = spin x valley degeneracy; 0 without carriers
template <class T>
T twoDStaticScreeningWavevector(T carrierDensity, T temperature, T envPermittivity, T dosEffMass, T degeneracy = T(4)) {
  if (carrierDensity <= T(0) || temperature <= T(0) || dosEffMass <= T(0))
    return T(0);
  const T dosZero = degeneracy * dosEffMass / (2 * constants::pi * constants::hbar * constants::hbar);
  const T thermal = constants::kB * temperature;
  const T dnDmu = dosZero * (1 - std::exp(-carrierDensity / (dosZero * thermal)));
  return constants::q * constants::q * dnDmu / (2 * constants::eps0 * envPermittivity);
}

template <class T> T twoDStaticDielectric(T q, T screeningWavevector) {
  if (q <= T(0) || screeningWavevector <= T(0))
    return T(1);
  return T(1) + screeningWavevector / q;
}

// 1 / eps(q)^2: what a bare |g(q)|^2 is multiplied with
template <class T> T twoDScreeningFactor(T q, T screeningWavevector) {
  const T eps = twoDStaticDielectric(q, screeningWavevector);
  return T(1) / (eps * eps);
}

#endif
