// One species of simulated particles: its valleys, its scatter mechanisms and
// how its particles are created.
// Interface mirrored: reference include/ParticleType/emcParticleType.hpp (pure
// virtuals :47-93, addValley :125-132, addScatterMechanism :141-153,
// init/reinitScatterTables :168-171, getTau/getNewTau/getNewGrainTau :180-193,
// check :196-206).
// Not here: scatterParticle*/surface/grain dispatch (:97-117) -- the per-event CPU
// path of the reference; events are processed by the device kernels.  The draw of
// getNewGrainTau() is kept: it is consumed at every particle creation whether or
// not a grain mechanism is set.
#ifndef EMC_PARTICLE_TYPE_HPP
#define EMC_PARTICLE_TYPE_HPP

#include <cmath>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include <ValleyTypes/emcAbstractValley.hpp>
#include <emcGrid.hpp>
#include <emcMessage.hpp>
#include <emcParticle.hpp>
#include <emcScatterHandler.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> struct emcParticleType {
  static const SizeType Dim = DeviceType::Dimension;
  typedef typename DeviceType::ValueVec ValueVec;
  typedef typename DeviceType::SizeVec SizeVec;
  typedef emcAbstractValley<T> AbstractValley;
  typedef emcScatterMechanism<T> AbstractScatterMechanism;

  std::vector<std::unique_ptr<AbstractValley>> valleys;
  emcScatterHandler<T, DeviceType> scatterHandler;
  mutable std::uniform_real_distribution<T> uniDistLog{1e-6, 1.};

  emcParticleType(SizeType inHandlerNrEnergyLevels = 1000, T inHandlerMaxEnergy = 4.)
      : scatterHandler(inHandlerNrEnergyLevels, inHandlerMaxEnergy) {}
  virtual ~emcParticleType() = default;

  virtual std::string getName() const = 0;
  virtual T getCharge() const = 0;
  virtual bool isMoved() const = 0;
  virtual bool isInjected() const = 0;
  // expected number of particles of a grid cell at start (fractional part = probability)
  virtual T getInitialNrParticles(const SizeVec &coord, const DeviceType &device,
                                  const emcGrid<T, Dim> &potential) = 0;

  // --- additive: the creation rules the contacts use on the device (emcgpu_particle_kind); -1 = none, such a type
  // cannot be injected on the GPU path ---
  virtual int deviceParticleKind() const { return -1; }

  virtual T getMass() const { return unimplemented("getMass", "isMoved"), T(0); }
  virtual emcParticle<T> generateInitialParticle(const SizeVec &, const DeviceType &, emcRNG &) {
    return unimplemented("generateInitialParticle", "isMoved"), emcParticle<T>();
  }
  virtual T getExpectedNrParticlesAtContact(const SizeVec &, const DeviceType &) {
    return unimplemented("getExpectedNrParticlesAtContact", "isInjected"), T(0);
  }
  virtual emcParticle<T> generateInjectedParticle(const SizeVec &, const DeviceType &, emcRNG &) {
    return unimplemented("generateInjectedParticle", "isInjected"), emcParticle<T>();
  }

  SizeType getNrValleys() const { return valleys.size(); }
  auto getValley(SizeType idxValley) const {
    requireValley(idxValley);
    return valleys[idxValley].get();
  }

  template <class DerivedValley>
  typename std::enable_if<std::is_base_of<AbstractValley, DerivedValley>::value>::type
  addValley(std::unique_ptr<DerivedValley> &&newValleyType) {
    newValleyType->check();
    valleys.push_back(std::move(newValleyType));
  }

  // regions: doping-region indices in which the mechanism acts
  template <class DerivedScatterMechanism>
  typename std::enable_if<std::is_base_of<AbstractScatterMechanism, DerivedScatterMechanism>::value>::type
  addScatterMechanism(const std::vector<int> &regions, std::unique_ptr<DerivedScatterMechanism> &&newMechanism) {
    requireValley(newMechanism->getIdxValley());
    newMechanism->setPtrValley(valleys);
    newMechanism->check();
    scatterHandler.addScatterMechanism(std::move(newMechanism), regions);
  }

  template <class DerivedSurfaceScatterMechanism>
  typename std::enable_if<
      std::is_base_of<emcSurfaceScatterMechanism<T, DeviceType>, DerivedSurfaceScatterMechanism>::value>::type
  setSurfaceScatterMechanism(emcBoundaryPos boundaryPosition, std::unique_ptr<DerivedSurfaceScatterMechanism> &&newMechanism) {
    scatterHandler.setSurfaceScatterMechanism(std::move(newMechanism), boundaryPosition);
  }

  void setGrainScatterMechanism(std::unique_ptr<emcGrainScatterMechanism<T>> &&newMechanism) {
    scatterHandler.setGrainScatterMechanism(std::move(newMechanism));
  }

  void initScatterTables() { scatterHandler.initScatterTables(); }
  void reinitScatterTables() { scatterHandler.reinitScatterTables(); }

  T getGrainTau() const { return scatterHandler.getGrainTau(); }
  T getTau(SizeType idxValley, SizeType idxRegion) const { return scatterHandler.getTau(idxRegion, idxValley); }
  T getNewTau(SizeType idxValley, SizeType idxRegion, emcRNG &rng) const {
    return -std::log(uniDistLog(rng)) * getTau(idxValley, idxRegion);
  }
  T getNewGrainTau(emcRNG &rng) const { return -std::log(uniDistLog(rng)) * getGrainTau(); }

  void check() const {
    if (isMoved() && valleys.empty())
      emcMessage::getInstance()
          .addError("Moving Particle Type " + getName() + " has to at least have one added valley.")
          .print();
  }

private:
  void requireValley(SizeType idxValley) const {
    if (idxValley >= valleys.size())
      emcMessage::getInstance().addError("Used index for Valley for " + getName() + " is invalid.").print();
  }
  int unimplemented(const std::string &func, const std::string &switchFunc) const {
    emcMessage::getInstance()
        .addError("Function " + func + "() is not implemented for " + getName() + ". Either let " + switchFunc +
                  "() return false or implement it for this ParticleType.")
        .print();
    return 0;
  }
};

#endif
