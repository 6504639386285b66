// Physical constants of the drop-in host API.
// Interface mirrored: reference include/emcConstants.hpp:9-27.  The values are the
// reference's literals (not CODATA): every rate table and every trajectory
// depends on them bit for bit.
#ifndef EMC_CONSTANTS_HPP
#define EMC_CONSTANTS_HPP

namespace constants {
constexpr double pi = 3.14159265358979323846;
constexpr double q = 1.60219e-19;    // elementary charge [C]
constexpr double kB = 1.38066e-23;   // Boltzmann constant [J/K]
constexpr double hbar = 1.05459e-34; // reduced Planck constant [Js]
constexpr double eps0 = 8.85419e-12; // vacuum permittivity [F/m]
constexpr double me = 9.11e-31;      // electron rest mass [kg]
constexpr double ke = 1. / (4 * pi * eps0);
} // namespace constants

#endif
