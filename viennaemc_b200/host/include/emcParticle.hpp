// One particle as the user-facing API sees it (AoS).  Interface mirrored:
// reference include/emcParticle.hpp:10-18.  On the GPU the ensemble lives as SoA
// streams (include/emcgpu.h); this struct is only the host-side exchange format
// (initial particle generation, plug-in signatures).
#ifndef EMC_PARTICLE_HPP
#define EMC_PARTICLE_HPP

#include <array>

#include <emcUtil.hpp>

template <class T> struct emcParticle {
  std::array<T, 3> k = {0, 0, 0}; // Herring-Vogt wave vector, device axes [1/m]
  T energy = 0;                   // [eV]
  T tau = 1;                      // remaining free-flight time [s]
  T grainTau = 1;                 // remaining time to the next grain boundary [s]
  SizeType valley = 0;
  SizeType subValley = 0;
  SizeType region = 0;
};

#endif
