// Polar-optical (Froehlich) scattering with equilibrium phonons, unscreened.
// Interface mirrored: reference include/ScatterMechanisms/emcFroehlichInteraction.hpp (emcFroehlichAbsorption3D,
// emcFroehlichEmission3D: valley, phonon energy [eV], relative effective mass, eps_hi, eps_lo, lattice temperature,
// name suffix).  Rates and device sampler: detail/emcPolarOpticalMechanism.hpp.
#ifndef EMC_FROEHLICH_INTERACTION_HPP
#define EMC_FROEHLICH_INTERACTION_HPP

#include <detail/emcPolarOpticalMechanism.hpp>

template <class T> class emcFroehlichAbsorption3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcFroehlichAbsorption3D() = delete;
  emcFroehlichAbsorption3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo, T temperature,
                           std::string inNameSuffix = "")
      : emcdetail::PolarOpticalMechanism<T>("FroehlichAbsorption3D", false, false, inValley, inPhononEnergy, relEffMass,
                                            eps_hi, eps_lo, inNameSuffix) {
    this->nBose = boseEinstein(inPhononEnergy, temperature);
  }
};

template <class T> class emcFroehlichEmission3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcFroehlichEmission3D() = delete;
  emcFroehlichEmission3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo, T temperature,
                         std::string inNameSuffix = "")
      : emcdetail::PolarOpticalMechanism<T>("FroehlichEmission3D", true, false, inValley, inPhononEnergy, relEffMass, eps_hi,
                                            eps_lo, inNameSuffix) {
    this->nBose = boseEinstein(inPhononEnergy, temperature);
  }
};

#endif
