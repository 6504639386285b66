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
public:
  // Debye: qs^2 = n q^2 / (eps_s eps_0 kB T) for a carrier density [1/m^3] at a carrier temperature [K]
  static T debyeWaveVectorSquared(T density, T carrierTemp, T inEpsStatic) {
    return density * constants::q * constants::q / (inEpsStatic * constants::eps0 * constants::kB * carrierTemp);
  }

  emcPlasmonScreening() = delete;
  explicit emcPlasmonScreening(T inEpsStatic, bool inEnabled = true) : staticPermittivity(inEpsStatic), switchedOn(inEnabled) {}

  bool isEnabled() const { return switchedOn; }
  // a disabled screening, an empty band or a cold one keep qs^2 = 0
  void update(T density, T carrierTemp) {
    value = (switchedOn && density > T(0) && carrierTemp > T(0)) ? debyeWaveVectorSquared(density, carrierTemp, staticPermittivity)
                                                                 : T(0);
  }
  void setQs2(T inQs2) { value = switchedOn ? inQs2 : T(0); }
  T getQs2() const { return value; }
  T getQs() const { return std::sqrt(value); }
  T getScreeningLength() const { return value > T(0) ? T(1) / std::sqrt(value) : std::numeric_limits<T>::infinity(); }

private:
  T value = T(0); // qs^2 [1/m^2]
  T staticPermittivity;
  bool switchedOn;
};

#endif
