// Silicon conduction band for the drop-in API: material constants, X valleys and the
// scattering set of the bulk / resistor / MOSFET examples.
// Parameter values: reference examples/SiliconFunctions.hpp:21-51 (Vasileska et al.);
// mechanism order as assembled by the reference examples (bulkSimulation.cpp:100-103,
// resistor2D.cpp:109-115).  The reference's own SiliconFunctions.hpp also compiles
// unchanged against these headers; this file exists so that the package does not
// depend on the reference tree at build time.
#ifndef EMC_EXAMPLES_SILICON_MODEL_HPP
#define EMC_EXAMPLES_SILICON_MODEL_HPP

#include <map>
#include <memory>
#include <vector>

#include <ScatterMechanisms/emcAcousticScatterMechanism.hpp>
#include <ScatterMechanisms/emcCoulombScatterMechanism.hpp>
#include <ScatterMechanisms/emcFirstOrderInterValleyScatterMechanism.hpp>
#include <ScatterMechanisms/emcZeroOrderInterValleyScatterMechanism.hpp>
#include <ValleyTypes/emcNonParabolicAnistropValley.hpp>
#include <emcMaterial.hpp>

namespace SiliconModel {

struct Parameters {
  double epsR = 11.8, rho = 2329., Ni = 1.45e16, vSound = 9040., bandgap = 1.15;
  double massLong = 0.916, massTrans = 0.196, alpha = 0.5; // X valley
  double sigmaAcoustic = 9.;                               // [eV]
  double D0f = 5.23e10, D0g = 5.23e10, hw0f = 0.06, hw0g = 0.06;   // zero order [eV/m], [eV]
  double D1f = 2.5, D1g = 4., hw1f = 0.023, hw1g = 0.018;          // first order
};

typedef std::map<SizeType, std::vector<SizeType>> SubValleyMap;
// g process: same axis; f process: one of the four valleys on the two other axes
inline SubValleyMap gFinal() { return {{0, {0}}, {1, {1}}, {2, {2}}}; }
inline SubValleyMap fFinal() { return {{0, {1, 1, 2, 2}}, {1, {0, 0, 2, 2}}, {2, {0, 0, 1, 1}}}; }

template <class T> emcMaterial<T> material(const Parameters &p = Parameters()) {
  return emcMaterial<T>(p.epsR, p.rho, p.Ni, p.vSound, p.bandgap);
}

// the six X valleys as one valley with three sub-valley frames (x, y, z longitudinal)
template <class T, class ParticleTypePtr> void addXValley(ParticleTypePtr &type, const Parameters &p = Parameters()) {
  auto valley = std::make_unique<emcNonParabolicAnisotropValley<T>>(
      std::array<T, 3>{p.massLong, p.massTrans, p.massTrans}, type->getMass(), 3, p.alpha);
  valley->setSubValleyEllipseCoordSystem(0, {1, 0, 0}, {0, 1, 0}, {0, 0, 1});
  valley->setSubValleyEllipseCoordSystem(1, {0, 1, 0}, {1, 0, 0}, {0, 0, 1});
  valley->setSubValleyEllipseCoordSystem(2, {0, 0, 1}, {0, 1, 0}, {1, 0, 0});
  type->addValley(std::move(valley));
}

enum Mechanisms : unsigned { ACOUSTIC = 1, ZERO_ORDER = 2, FIRST_ORDER = 4, COULOMB = 8 };

// Acoustic, [Coulomb,] zero-order f/g abs/em, first-order f/g abs/em -- the order of the reference drivers
// (Coulomb last, as in resistor2D, unless coulombSecond: the order of mosfet2D)
template <class T, class ParticleTypePtr, class DeviceType>
void addScattering(ParticleTypePtr &type, DeviceType &device, const std::vector<int> &regions, unsigned which,
                   bool coulombSecond = false, const Parameters &p = Parameters()) {
  typedef emcZeroOrderInterValleyAbsorptionScatterMechanism<T> Z_A;
  typedef emcZeroOrderInterValleyEmissionScatterMechanism<T> Z_E;
  typedef emcFirstOrderInterValleyAbsorptionScatterMechanism<T> F_A;
  typedef emcFirstOrderInterValleyEmissionScatterMechanism<T> F_E;
  auto coulomb = [&] {
    type->addScatterMechanism(regions, std::make_unique<emcCoulombScatterMechanism<T, DeviceType>>(0, p.epsR, device));
  };
  if (which & ACOUSTIC)
    type->addScatterMechanism(regions, std::make_unique<emcAcousticScatterMechanism<T>>(0, p.sigmaAcoustic, device));
  if ((which & COULOMB) && coulombSecond)
    coulomb();
  if (which & ZERO_ORDER) {
    type->addScatterMechanism(regions, std::make_unique<Z_A>("F", 0, fFinal(), p.D0f, p.hw0f, device));
    type->addScatterMechanism(regions, std::make_unique<Z_E>("F", 0, fFinal(), p.D0f, p.hw0f, device));
    type->addScatterMechanism(regions, std::make_unique<Z_A>("G", 0, gFinal(), p.D0g, p.hw0g, device));
    type->addScatterMechanism(regions, std::make_unique<Z_E>("G", 0, gFinal(), p.D0g, p.hw0g, device));
  }
  if (which & FIRST_ORDER) {
    type->addScatterMechanism(regions, std::make_unique<F_A>("F", 0, fFinal(), p.D1f, p.hw1f, device));
    type->addScatterMechanism(regions, std::make_unique<F_E>("F", 0, fFinal(), p.D1f, p.hw1f, device));
    type->addScatterMechanism(regions, std::make_unique<F_A>("G", 0, gFinal(), p.D1g, p.hw1g, device));
    type->addScatterMechanism(regions, std::make_unique<F_E>("G", 0, gFinal(), p.D1g, p.hw1g, device));
  }
  if ((which & COULOMB) && !coulombSecond)
    coulomb();
}

} // namespace SiliconModel

#endif
