// Polar-optical (Froehlich) scattering with a free-carrier-screened coupling: equilibrium phonons
// (emcScreenedFroehlich{Absorption,Emission}3D) or a non-equilibrium phonon bath
// (emcScreenedHotPhononFroehlich{Absorption,Emission}3D, optionally with the occupation -- and the polar angle --
// resolved in the transferred |q|).
// Interface mirrored: reference include/ScatterMechanisms/emcScreenedFroehlichInteraction.hpp (constructor signatures
// :112-125, :171-184, :229-246, :311-328).
#ifndef EMC_SCREENED_FROEHLICH_INTERACTION_HPP
#define EMC_SCREENED_FROEHLICH_INTERACTION_HPP

#include <ScatterMechanisms/emcFroehlichInteraction.hpp>

template <class T> class emcScreenedFroehlichAbsorption3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcScreenedFroehlichAbsorption3D() = delete;
  emcScreenedFroehlichAbsorption3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo, T temperature,
                                   std::shared_ptr<emcPlasmonScreening<T>> inScreening, std::string inNameSuffix = "")
      : emcdetail::PolarOpticalMechanism<T>("ScreenedFroehlichAbsorption3D", false, true, inValley, inPhononEnergy, relEffMass,
                                            eps_hi, eps_lo, inNameSuffix) {
    this->nBose = boseEinstein(inPhononEnergy, temperature);
    this->screening = std::move(inScreening);
  }
};

template <class T> class emcScreenedFroehlichEmission3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcScreenedFroehlichEmission3D() = delete;
  emcScreenedFroehlichEmission3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo, T temperature,
                                 std::shared_ptr<emcPlasmonScreening<T>> inScreening, std::string inNameSuffix = "")
      : emcdetail::PolarOpticalMechanism<T>("ScreenedFroehlichEmission3D", true, true, inValley, inPhononEnergy, relEffMass,
                                            eps_hi, eps_lo, inNameSuffix) {
    this->nBose = boseEinstein(inPhononEnergy, temperature);
    this->screening = std::move(inScreening);
  }
};

template <class T> class emcScreenedHotPhononFroehlichAbsorption3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcScreenedHotPhononFroehlichAbsorption3D() = delete;
  emcScreenedHotPhononFroehlichAbsorption3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo,
                                            std::shared_ptr<emcPhononBath<T>> inPhononBath,
                                            std::shared_ptr<emcPlasmonScreening<T>> inScreening, bool inQResolved = false,
                                            std::string inNameSuffix = "", bool inQResolvedAngle = true)
      : emcdetail::PolarOpticalMechanism<T>("ScreenedHotPhononFroehlichAbsorption3D", false, true, inValley, inPhononEnergy,
                                            relEffMass, eps_hi, eps_lo, inNameSuffix) {
    this->phononBath = std::move(inPhononBath);
    this->screening = std::move(inScreening);
    this->qResolved = inQResolved;
    this->qResolvedAngle = inQResolvedAngle;
  }
};

template <class T> class emcScreenedHotPhononFroehlichEmission3D : public emcdetail::PolarOpticalMechanism<T> {
public:
  emcScreenedHotPhononFroehlichEmission3D() = delete;
  emcScreenedHotPhononFroehlichEmission3D(SizeType inValley, T inPhononEnergy, T relEffMass, T eps_hi, T eps_lo,
                                          std::shared_ptr<emcPhononBath<T>> inPhononBath,
                                          std::shared_ptr<emcPlasmonScreening<T>> inScreening, bool inQResolved = false,
                                          std::string inNameSuffix = "", bool inQResolvedAngle = true)
      : emcdetail::PolarOpticalMechanism<T>("ScreenedHotPhononFroehlichEmission3D", true, true, inValley, inPhononEnergy,
                                            relEffMass, eps_hi, eps_lo, inNameSuffix) {
    this->phononBath = std::move(inPhononBath);
    this->screening = std::move(inScreening);
    this->qResolved = inQResolved;
    this->qResolvedAngle = inQResolvedAngle;
  }
};

#endif
