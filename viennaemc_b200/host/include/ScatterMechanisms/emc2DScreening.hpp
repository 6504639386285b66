// Static free-carrier screening of a 2-D carrier gas -- the screening wave vector the single-layer parameter sets hand to their
// long-range mechanisms, and the dielectric function that goes with it.  Interface mirrored: reference
// include/ScatterMechanisms/emc2DScreening.hpp:49-76 (three free function templates; operation order kept).
#ifndef EMC_2D_SCREENING_HPP
#define EMC_2D_SCREENING_HPP

#include <cmath>

#include <emcConstants.hpp>

namespace emcdetail {
// thermally averaged density of states dn/dmu = D0 (1 - exp(-n / (D0 kB T))) with D0 = g m / (2 pi hbar^2): D0 in the
// degenerate limit, n / (kB T) in the Debye limit
template <class T> T sheetCompressibility(T sheetDensity, T temperature, T mass, T degeneracy) {
  const T dosAtEdge = degeneracy * mass / (2 * constants::pi * constants::hbar * constants::hbar);
  return dosAtEdge * (1 - std::exp(-sheetDensity / (dosAtEdge * (constants::kB * temperature))));
}
} // namespace emcdetail

// q_s = q^2 (dn/dmu) / (2 eps0 eps_r) [1/m]; sheet density [1/m^2], temperature [K], relative permittivity of the
// surroundings, density-of-states mass [kg], g = spin x valley degeneracy.  No carriers (or no temperature / mass): 0.
template <class T>
T twoDStaticScreeningWavevector(T carrierDensity, T temperature, T envPermittivity, T dosEffMass, T degeneracy = T(4)) {
  const bool defined = carrierDensity > T(0) && temperature > T(0) && dosEffMass > T(0);
  return defined ? constants::q * constants::q * emcdetail::sheetCompressibility(carrierDensity, temperature, dosEffMass, degeneracy) /
                       (2 * constants::eps0 * envPermittivity)
                 : T(0);
}

// eps(q) = 1 + q_s / q; 1 where either argument is not positive
template <class T> T twoDStaticDielectric(T q, T screeningWavevector) {
  return (q > T(0) && screeningWavevector > T(0)) ? T(1) + screeningWavevector / q : T(1);
}

// 1 / eps(q)^2: the factor on a bare |g(q)|^2
template <class T> T twoDScreeningFactor(T q, T screeningWavevector) {
  const T dielectric = twoDStaticDielectric(q, screeningWavevector);
  return T(1) / (dielectric * dielectric);
}

#endif
