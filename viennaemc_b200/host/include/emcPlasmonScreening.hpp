// Free-carrier (Debye) screening of the polar-optical coupling: one shared number, the squared inverse screening
// length qs^2 = n q^2 / (eps_s eps_0 kB T_e), that the screened Froehlich mechanisms read when their rates are tabulated
// and that travels to the device with the mechanism descriptors.
// Interface mirrored: reference include/emcPlasmonScreening.hpp (ctor, update :69-76, setQs2, getQs2, getQs,
// getScreeningLength, isEnabled).
#ifndef EMC_PLASMON_SCREENING_HPP
#define EMC_PLASMON_SCREENING_HPP

#include <cmath>
#include <limits>

#include <emcConstants.hpp>
#include <emcUtil.hpp>

template <class T> class emcPlasmonScreening {
  T qs2 = T(0);
  T epsStatic;
  bool enabled;

public:
  emcPlasmonScreening() = delete;
  explicit emcPlasmonScreening(T inEpsStatic, bool inEnabled = true) : epsStatic(inEpsStatic), enabled(inEnabled) {}

  // carrier density [1/m^3] and carrier temperature [K]; a disabled screening keeps qs^2 = 0
  void update(T density, T carrierTemp) {
    const bool active = enabled && density > T(0) && carrierTemp > T(0);
    qs2 = active ? density * constants::q * constants::q / (epsStatic * constants::eps0 * constants::kB * carrierTemp) : T(0);
  }
  void setQs2(T inQs2) { qs2 = enabled ? inQs2 : T(0); }
  T getQs2() const { return qs2; }
  T getQs() const { return std::sqrt(qs2); }
  T getScreeningLength() const { return qs2 > T(0) ? T(1) / std::sqrt(qs2) : std::numeric_limits<T>::infinity(); }
  bool isEnabled() const { return enabled; }
};

#endif
