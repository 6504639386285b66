// Polar-optical (Froehlich) scattering coupled to a non-equilibrium phonon bath, unscreened: the occupation in the rates
// is the bath's mean occupation, every event is counted in the bath's |q| bin.
// Interface mirrored: reference include/ScatterMechanisms/emcHotPhononFroehlichMechanism.hpp
// (emcHotPhononFroehlichAbsorption3D, emcHotPhononFroehlichEmission3D).
#ifndef EMC_HOT_PHONON_FROEHLICH_MECHANISM_HPP
#define EMC_HOT_PHONON_FROEHLICH_MECHANISM_HPP

#include <ScatterMechanisms/emcFroehlichInteraction.hpp>

template <class T> class emcHotPhononFroehlichAbsorption3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcHotPhononFroehlichAbsorption3D() = delete;
  emcHotPhononFroehlichAbsorption3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo,
                                    std::shared_ptr<emcPhononBath<T>> inPhononBath, std::string inNameSuffix = "")
      : emcdetail::PolarOpticalMechanism<T>("HotPhononFroehlichAbsorption3D", false, false, inValley, inPhononEnergy,
                                            relEffMass, eps_hi, eps_lo, inNameSuffix) {
    this->phononBath = std::move(inPhononBath);
  }
};

template <class T> class emcHotPhononFroehlichEmission3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcHotPhononFroehlichEmission3D() = delete;
  emcHotPhononFroehlichEmission3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo,
                                  std::shared_ptr<emcPhononBath<T>> inPhononBath, std::string inNameSuffix = "")
      : emcdetail::PolarOpticalMechanism<T>("HotPhononFroehlichEmission3D", true, false, inValley, inPhononEnergy, relEffMass,
                                            eps_hi, eps_lo, inNameSuffix) {
    this->phononBath = std::move(inPhononBath);
  }
};

#endif
