// Electrons that behave as implemented in ViennaWD -- the particle type of the MOSFET example.
// Interface mirrored: reference examples/mosfet2D/electronVWD.hpp (same struct name).  Differences to emcElectron:
// the initial density always follows the equilibrium potential and particle numbers are rounded (:40-62); valley and
// sub-valley are drawn from U[0,1) (:25); the first free-flight time is looked up with the VALLEY index in the place of
// the region index (:74, :87).  Initial particles are created here on the host (seeded like the reference); the
// particles the contacts inject during the run are created on the device by the same rules
// (EMCGPU_PARTICLE_ELECTRON_VWD).
#ifndef ELECTRON_VWD_HPP
#define ELECTRON_VWD_HPP

#include <cmath>

#include <emcgpu.h>

#include <ParticleType/emcParticleType.hpp>
#include <emcConstants.hpp>
#include <emcParticleInitialization.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> struct electronVWD : public emcParticleType<T, DeviceType> {
  typedef typename DeviceType::ValueVec ValueVec;
  typedef typename DeviceType::SizeVec SizeVec;
  static const SizeType Dim = DeviceType::Dimension;

  std::uniform_real_distribution<T> dist{0, 1};

  electronVWD(SizeType inHandlerNrEnergyLevels = 1000, T inHandlerMaxEnergy = 4.)
      : emcParticleType<T, DeviceType>(inHandlerNrEnergyLevels, inHandlerMaxEnergy) {}

  std::string getName() const override { return "Electrons"; }
  T getMass() const override { return constants::me; }
  T getCharge() const override { return -constants::q; }
  bool isMoved() const override { return true; }
  bool isInjected() const override { return true; }
  int deviceParticleKind() const override { return EMCGPU_PARTICLE_ELECTRON_VWD; }

  T getInitialNrParticles(const SizeVec &coord, const DeviceType &device, const emcGrid<T, Dim> &potential) override {
    T density = std::exp(potential[coord]) * device.getMaterial().getNi();
    for (SizeType d = 0; d < Dim; d++)
      if (coord[d] == 0 || coord[d] == potential.getSize(d) - 1)
        density *= 0.5;
    return std::round(density * device.getCellVolume());
  }

  T getExpectedNrParticlesAtContact(const SizeVec &coord, const DeviceType &device) override {
    T expected = device.getCellVolume() * device.getDopingProfile().getDoping(coord);
    const auto extent = device.getGridExtent();
    for (SizeType d = 0; d < Dim; d++)
      if (coord[d] == 0 || coord[d] == extent[d] - 1)
        expected *= 0.5;
    return std::round(expected);
  }

  emcParticle<T> generateInitialParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    return create(coord, device, rng);
  }
  emcParticle<T> generateInjectedParticle(const SizeVec &coord, const DeviceType &device, emcRNG &rng) override {
    return create(coord, device, rng);
  }

private:
  emcParticle<T> create(const SizeVec &coord, const DeviceType &device, emcRNG &rng) {
    emcParticle<T> part;
    part.region = device.getDopingProfile().getDopingRegionIdx(coord);
    part.valley = std::floor(this->getNrValleys() * dist(rng));
    auto valley = this->getValley(part.valley);
    part.subValley = std::floor(valley->getDegeneracyFactor() * dist(rng));
    initParticleKSpaceMaxwellian(part, coord, device, valley, rng);
    part.tau = this->getNewTau(part.valley, part.valley, rng); // the valley index where the region belongs, as ViennaWD
    part.grainTau = this->getNewGrainTau(rng);
    return part;
  }
};

#endif
