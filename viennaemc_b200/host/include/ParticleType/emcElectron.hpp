// Conduction-band electrons.
// Interface mirrored: reference include/ParticleType/emcElectron.hpp (ctor :28-35,
// getInitialNrParticles :48-60, getExpectedNrParticlesAtContact :63-73,
// generateInitialParticle :75-90, generateInjectedParticle :92-104).
#ifndef EMC_ELECTRON_HPP
#define EMC_ELECTRON_HPP

#include <emcgpu.h>

#include <ParticleType/emcParticleType.hpp>
#include <emcConstants.hpp>
#include <emcParticleInitialization.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> struct emcElectron : public emcParticleType<T, DeviceType> {
  typedef typename DeviceType::ValueVec ValueVec;
  typedef typename DeviceType::SizeVec SizeVec;
  static const SizeType Dim = DeviceType::Dimension;

  std::uniform_real_distribution<T> dist{1e-6, 1.};
  bool usePotentialForInit; // initial density from exp(potential) Ni instead of the doping
  T initEnergyEV;           // > 0: mono-energetic start; 0: Maxwellian at the lattice temperature

  emcElectron(SizeType inHandlerNrEnergyLevels = 1000, T inHandlerMaxEnergy = 4., bool inUsePotentialForInit = true,
              T inInitEnergyEV = T(0))
      : emcParticleType<T, DeviceType>(inHandlerNrEnergyLevels, inHandlerMaxEnergy),
        usePotentialForInit(inUsePotentialForInit), initEnergyEV(inInitEnergyEV) {}

  std::string getName() const override { return "Electrons"; }
  T getMass() const override { return constants::me; }
  T getCharge() const override { return -constants::q; }
  bool isMoved() const override { return true; }
  bool isInjected() const override { return true; }
  int deviceParticleKind() const override { return EMCGPU_PARTICLE_ELECTRON; }

  T getInitialNrParticles(const SizeVec &coord, const DeviceType &device, const emcGrid<T, Dim> &potential) override {
    T density = usePotentialForInit ? std::exp(potential[coord]) * device.getMaterial().getNi()
                                    : device.getDopingProfile().getDoping(coord);
    for (SizeType d = 0; d < Dim; d++)
      if (coord[d] == 0 || coord[d] == potential.getSize(d) - 1)
        density *= 0.5; // half cell at a face
    return density * device.getCellVolume();
  }

  T getExpectedNrParticlesAtContact(const SizeVec &coord, const DeviceType &device) override {
    T expected = device.getCellVolume() * device.getDopingProfile().getDoping(coord);
    const auto extent = device.getGridExtent();
    for (SizeType d = 0; d < Dim; d++)
      if (coord[d] == 0 || coord[d] == extent[d] - 1)
        expected *= 0.5;
    return expected;
  }

  emcParticle<T> generateInitialParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    return create(coord, device, rng, initEnergyEV > T(0));
  }
  emcParticle<T> generateInjectedParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    return create(coord, device, rng, false);
  }

private:
  // draw order: valley, sub-valley, energy, cos(theta), phi, tau, grainTau
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

#endif
