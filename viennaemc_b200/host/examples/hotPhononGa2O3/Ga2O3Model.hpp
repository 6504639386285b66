// beta-Ga2O3 conduction band for the drop-in API: material constants, Gamma valley, and the scattering set of the
// hot-phonon example (acoustic, non-polar optical, polar optical with equilibrium or non-equilibrium phonons).
// Parameter values: reference examples/hotPhononGa2O3/Ga2O3Functions.hpp:45-88 (Ma 2016, Ghosh 2017, Santia 2019).
// The reference's own Ga2O3Functions.hpp also compiles unchanged against these headers; this file exists so that the
// package does not depend on the reference tree at build time.
#ifndef EMC_EXAMPLES_GA2O3_MODEL_HPP
#define EMC_EXAMPLES_GA2O3_MODEL_HPP

#include <map>
#include <memory>
#include <string>
#include <vector>

#include <ScatterMechanisms/emcAcousticScatterMechanism.hpp>
#include <ScatterMechanisms/emcCoulombScatterMechanism.hpp>
#include <ScatterMechanisms/emcHotPhononFroehlichMechanism.hpp>
#include <ScatterMechanisms/emcScreenedFroehlichInteraction.hpp>
#include <ScatterMechanisms/emcZeroOrderInterValleyScatterMechanism.hpp>
#include <ValleyTypes/emcNonParabolicIsotropValley.hpp>
#include <emcMaterial.hpp>
#include <emcPhononBath.hpp>
#include <emcPlasmonScreening.hpp>

namespace Ga2O3Model {

struct Parameters {
  double epsLo = 10.2, epsHi = 3.573, rho = 5880., vSound = 6800., bandGap = 4.85, Ni = 1.;
  double relEffMass = 0.284, alpha = 0.106;
  double hwPOP = 0.044;                                  // single-mode lump [eV]
  std::vector<double> hwModes = {0.0302, 0.0429, 0.0646, 0.0796, 0.0936}; // five polar modes [eV]
  std::vector<double> modeWeight = {0.1382, 0.0326, 0.2448, 0.1510, 0.4334};
  double hwNPO = 0.090, defPotNPO = 8.05e10;             // non-polar optical [eV], [eV/m]
  double sigmaAc = 4.8;                                  // acoustic deformation potential [eV]
  double tauLO = 5e-12, tauAc = 20e-12;                  // phonon decay times [s]
  SizeType nrPhononBins = 300;
  double dqBin = 1e7;                                    // [1/m]
};

// effective static permittivity of each mode so that the mode couplings add up to the full one (sum rule)
inline std::vector<double> modeEpsLo(const Parameters &p) {
  const double invHi = 1. / p.epsHi, total = invHi - 1. / p.epsLo;
  std::vector<double> out;
  for (double w : p.modeWeight)
    out.push_back(1. / (invHi - w * total));
  return out;
}

template <class T> emcMaterial<T> material(const Parameters &p = Parameters()) {
  return emcMaterial<T>(p.epsLo, p.rho, p.Ni, p.vSound, p.bandGap);
}

struct PolarSetup { // what a driver keeps to run the hot-phonon loop
  std::vector<double> modeEnergy, modeEpsLo, modeWeight;
  std::shared_ptr<emcPlasmonScreening<double>> screening;
  std::vector<std::shared_ptr<emcPhononBath<double>>> baths;
};

// Gamma valley + acoustic + non-polar optical [+ Brooks-Herring] + polar optical (screened classes; hot: one bath per mode)
template <class T, class ParticleTypePtr, class DeviceType>
PolarSetup addBandAndScattering(ParticleTypePtr &type, DeviceType &device, bool hotPhonons, bool multimode, bool screeningOn,
                                bool qResolved, bool qResolvedAngle, bool impurities, bool acousticBath, T temperature, T Vsim,
                                const Parameters &p = Parameters()) {
  typedef std::map<SizeType, std::vector<SizeType>> SubValleyMap;
  type->addValley(std::make_unique<emcNonParabolicIsotropValley<T>>(p.relEffMass, type->getMass(), 1, p.alpha));
  const std::vector<int> regions = {0};
  type->addScatterMechanism(regions, std::make_unique<emcAcousticScatterMechanism<T>>(0, p.sigmaAc, device));
  const SubValleyMap same = {{0, {0}}};
  type->addScatterMechanism(regions, std::make_unique<emcZeroOrderInterValleyAbsorptionScatterMechanism<T>>(
                                         "NPO", 0, same, p.defPotNPO, p.hwNPO, device));
  type->addScatterMechanism(regions, std::make_unique<emcZeroOrderInterValleyEmissionScatterMechanism<T>>(
                                         "NPO", 0, same, p.defPotNPO, p.hwNPO, device));
  if (impurities)
    type->addScatterMechanism(regions, std::make_unique<emcCoulombScatterMechanism<T, DeviceType>>(0, p.epsLo, device));
  PolarSetup s;
  s.modeEnergy = multimode ? p.hwModes : std::vector<double>{p.hwPOP};
  s.modeEpsLo = multimode ? modeEpsLo(p) : std::vector<double>{p.epsLo};
  s.modeWeight = multimode ? p.modeWeight : std::vector<double>{1.};
  s.screening = std::make_shared<emcPlasmonScreening<T>>(p.epsLo, screeningOn);
  for (SizeType m = 0; m < s.modeEnergy.size(); m++) {
    const std::string suffix = "Ga2O3-" + std::to_string(m);
    const T hw = s.modeEnergy[m];
    if (hotPhonons) {
      s.baths.push_back(std::make_shared<emcPhononBath<T>>(p.nrPhononBins, p.dqBin, p.tauLO, hw, temperature, Vsim, acousticBath,
                                                          hw / 2., p.tauAc));
      type->addScatterMechanism(regions, std::make_unique<emcScreenedHotPhononFroehlichAbsorption3D<T>>(
                                             0, hw, p.relEffMass, p.epsHi, s.modeEpsLo[m], s.baths[m], s.screening, qResolved,
                                             suffix, qResolvedAngle));
      type->addScatterMechanism(regions, std::make_unique<emcScreenedHotPhononFroehlichEmission3D<T>>(
                                             0, hw, p.relEffMass, p.epsHi, s.modeEpsLo[m], s.baths[m], s.screening, qResolved,
                                             suffix, qResolvedAngle));
    } else {
      type->addScatterMechanism(regions, std::make_unique<emcScreenedFroehlichAbsorption3D<T>>(
                                             0, hw, p.relEffMass, p.epsHi, s.modeEpsLo[m], temperature, s.screening, suffix));
      type->addScatterMechanism(regions, std::make_unique<emcScreenedFroehlichEmission3D<T>>(
                                             0, hw, p.relEffMass, p.epsHi, s.modeEpsLo[m], temperature, s.screening, suffix));
    }
  }
  return s;
}

} // namespace Ga2O3Model

#endif
