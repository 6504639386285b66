// Bulk material parameters.  Interface mirrored: reference include/emcMaterial.hpp.
#ifndef EMC_MATERIAL_HPP
#define EMC_MATERIAL_HPP

#include <emcConstants.hpp>
#include <emcUtil.hpp>

template <class T> class emcMaterial {
  T epsR, rho, Ni, velSound, bandGap;

public:
  emcMaterial() = delete;
  // relative permittivity, density [kg/m3], intrinsic carrier concentration [1/m3],
  // velocity of sound [m/s], band gap [eV]
  emcMaterial(T inEpsR, T inRho, T inNi, T inVelSound, T inBandGap)
      : epsR(inEpsR), rho(inRho), Ni(inNi), velSound(inVelSound), bandGap(inBandGap) {}
  T getEpsR() const { return epsR; }
  T getDielectricConstant() const { return epsR * constants::eps0; }
  T getRho() const { return rho; }
  T getNi() const { return Ni; }
  T getVelSound() const { return velSound; }
  T getBandGap() const { return bandGap; }
};

#endif
