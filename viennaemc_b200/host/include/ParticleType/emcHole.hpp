// Valence-band holes: the mirror image of emcElectron (charge +q, density from exp(-potential) ni or from |doping|).
// Interface mirrored: reference include/ParticleType/emcHole.hpp (ctor :38-44, getInitialNrParticles :56-71,
// getExpectedNrParticlesAtContact :73-84, generateInitialParticle :86-102, generateInjectedParticle :104-108).
// On the device a hole is a particle like any other: its valleys, tables and samplers cross the C ABI exactly as the
// electrons' do, the sign of the charge only enters the force (emcgpu_bulk_configure / emcgpu_device_configure).
#ifndef EMC_HOLE_HPP
#define EMC_HOLE_HPP

#include <emcgpu.h>

#include <ParticleType/emcParticleType.hpp>
#include <emcConstants.hpp>
#include <emcParticleInitialization.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> struct emcHole : public emcParticleType<T, DeviceType> {
  typedef typename DeviceType::ValueVec ValueVec;
  typedef typename DeviceType::SizeVec SizeVec;
  static const SizeType Dim = DeviceType::Dimension;

  std::uniform_real_distribution<T> dist{1e-6, 1.};
  bool usePotentialForInit; // initial density from exp(-potential) ni instead of |doping|
  T initEnergyEV;           // > 0: mono-energetic (photo-excited) start; 0: Maxwellian at the lattice temperature

  emcHole(SizeType inHandlerNrEnergyLevels = 1000, T inHandlerMaxEnergy = 4., bool inUsePotentialForInit = false,
          T inInitEnergyEV = T(0))
      : emcParticleType<T, DeviceType>(inHandlerNrEnergyLevels, inHandlerMaxEnergy),
        usePotentialForInit(inUsePotentialForInit), initEnergyEV(inInitEnergyEV) {}

  std::string getName() const override { return "Holes"; }
  T getMass() const override { return constants::me; }
  T getCharge() const override { return +constants::q; }
  bool isMoved() const override { return true; }
  bool isInjected() const override { return true; }
  // the creation rule at contacts is the electrons' (thermal, initial-particle rule): same device kind
  int deviceParticleKind() const override { return EMCGPU_PARTICLE_ELECTRON; }

  T getInitialNrParticles(const SizeVec &coord, const DeviceType &device, const emcGrid<T, Dim> &potential) override {
    T density = usePotentialForInit ? std::exp(-potential[coord]) * device.getMaterial().getNi()
                                    : std::fabs(device.getDopingProfile().getDoping(coord));
    for (SizeType d = 0; d < Dim; d++)
      if (coord[d] == 0 || coord[d] == potential.getSize(d) - 1)
        density *= T(0.5); // half cell at a face
    return density * device.getCellVolume();
  }

  T getExpectedNrParticlesAtContact(const SizeVec &coord, const DeviceType &device) override {
    T expected = device.getCellVolume() * std::fabs(device.getDopingProfile().getDoping(coord));
    const auto extent = device.getGridExtent();
    for (SizeType d = 0; d < Dim; d++)
      if (coord[d] == 0 || coord[d] == extent[d] - 1)
        expected *= T(0.5);
    return expected;
  }

  // draw order: valley, sub-valley, [energy,] cos(theta), phi, tau, grainTau
  emcParticle<T> generateInitialParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    emcParticle<T> part;
    part.region = device.getDopingProfile().getDopingRegionIdx(coord);
    part.valley = std::floor(this->getNrValleys() * dist(rng));
    auto valley = this->getValley(part.valley);
    part.subValley = std::floor(valley->getDegeneracyFactor() * dist(rng));
    if (initEnergyEV > T(0))
      initParticleKSpaceFixed(part, initEnergyEV, coord, device, valley, rng);
    else
      initParticleKSpaceMaxwellian(part, coord, device, valley, rng);
    part.tau = this->getNewTau(part.valley, part.region, rng);
    part.grainTau = this->getNewGrainTau(rng);
    return part;
  }
  emcParticle<T> generateInjectedParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    return generateInitialParticle(coord, device, rng);
  }
};

#endif
