// One implementation behind the two mobile carrier types of the API, emcElectron (conduction band, charge -q) and emcHole
// (valence band, charge +q).  The reference has two classes (include/ParticleType/emcElectron.hpp, emcHole.hpp) that differ
// in four places, kept here as properties of the sign:
//   * charge and name;
//   * the initial density of a cell: ni exp(+potential) or the doping for electrons (emcElectron.hpp:48-60), ni exp(-potential)
//     or |doping| for holes (emcHole.hpp:56-71); the expected population of a contact cell likewise (:63-73 / :73-84);
//   * the default of usePotentialForInit (true for electrons, false for holes -- the constructors' defaults);
//   * an INJECTED electron is always thermal (emcElectron.hpp:92-104), an injected hole is created like an initial one
//     (emcHole.hpp:104-108).
// Draw order of a new particle: valley, sub-valley, [energy,] cos(theta), phi, tau, grainTau.
// On the device a carrier of either sign is a particle like any other: valleys, tables and samplers cross the C ABI the same
// way, the sign of the charge only enters the force (emcgpu_bulk_configure / emcgpu_device_configure).
#ifndef EMC_DETAIL_BAND_CARRIER_HPP
#define EMC_DETAIL_BAND_CARRIER_HPP

#include <cmath>
#include <random>
#include <string>

#include <emcgpu.h>

#include <ParticleType/emcParticleType.hpp>
#include <emcConstants.hpp>
#include <emcParticleInitialization.hpp>
#include <emcUtil.hpp>

namespace emcdetail {

template <class T, class DeviceType, int ChargeSign> struct BandCarrier : public emcParticleType<T, DeviceType> {
  static_assert(ChargeSign == 1 || ChargeSign == -1, "charge sign");
  typedef typename DeviceType::ValueVec ValueVec;
  typedef typename DeviceType::SizeVec SizeVec;
  static const SizeType Dim = DeviceType::Dimension;
  static constexpr bool isHole = ChargeSign > 0;

  std::uniform_real_distribution<T> dist{1e-6, 1.};
  bool usePotentialForInit; // initial density from the potential (ni exp(-+ potential)) instead of the doping
  T initEnergyEV;           // > 0: mono-energetic (photo-excited) start; 0: Maxwellian at the lattice temperature

  BandCarrier(SizeType nrEnergyLevels, T maxEnergy, bool inUsePotentialForInit, T inInitEnergyEV)
      : emcParticleType<T, DeviceType>(nrEnergyLevels, maxEnergy), usePotentialForInit(inUsePotentialForInit),
        initEnergyEV(inInitEnergyEV) {}

  std::string getName() const override { return isHole ? "Holes" : "Electrons"; }
  T getMass() const override { return constants::me; }
  T getCharge() const override { return isHole ? +constants::q : -constants::q; }
  bool isMoved() const override { return true; }
  bool isInjected() const override { return true; }
  // creation rule at contacts: thermal, by the initial-particle rule
  int deviceParticleKind() const override { return EMCGPU_PARTICLE_ELECTRON; }

  T getInitialNrParticles(const SizeVec &coord, const DeviceType &device, const emcGrid<T, Dim> &potential) override {
    T density;
    if (usePotentialForInit)
      density = (isHole ? std::exp(-potential[coord]) : std::exp(potential[coord])) * device.getMaterial().getNi();
    else
      density = dopingOf(device, coord);
    return halveAtFaces(density, coord, [&](SizeType d) { return potential.getSize(d); }) * device.getCellVolume();
  }

  T getExpectedNrParticlesAtContact(const SizeVec &coord, const DeviceType &device) override {
    const auto extent = device.getGridExtent();
    return halveAtFaces(device.getCellVolume() * dopingOf(device, coord), coord, [&](SizeType d) { return extent[d]; });
  }

  emcParticle<T> generateInitialParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    return create(coord, device, rng, initEnergyEV > T(0));
  }
  emcParticle<T> generateInjectedParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    return create(coord, device, rng, isHole && initEnergyEV > T(0));
  }

private:
  static T dopingOf(const DeviceType &device, const SizeVec &coord) {
    const T doping = device.getDopingProfile().getDoping(coord);
    return isHole ? std::fabs(doping) : doping;
  }
  // a grid point on a face of the box owns half a cell per face it lies on
  template <class SizeOf> static T halveAtFaces(T value, const SizeVec &coord, SizeOf &&sizeOf) {
    for (SizeType d = 0; d < Dim; d++)
      if (coord[d] == 0 || coord[d] == sizeOf(d) - 1)
        value *= T(0.5);
    return value;
  }
  emcParticle<T> create(const SizeVec &coord, const DeviceType &device, emcRNG &rng, bool monoEnergetic) {
    emcParticle<T> part;
    part.region = device.getDopingProfile().getDopingRegionIdx(coord);
    part.valley = std::floor(this->getNrValleys() * dist(rng));
    auto valley = this->getValley(part.valley);
    part.subValley = std::floor(valley->getDegeneracyFactor() * dist(rng));
    if (monoEnergetic)
      initParticleKSpaceFixed(part, initEnergyEV, coord, device, valley, rng);
    else
      initParticleKSpaceMaxwellian(part, coord, device, valley, rng);
    part.tau = this->getNewTau(part.valley, part.region, rng);
    part.grainTau = this->getNewGrainTau(rng);
    return part;
  }
};

} // namespace emcdetail

#endif
