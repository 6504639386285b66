// Parameters of a device run.  Interface mirrored: reference include/emcSimulationParameter.hpp
// (ctors :54-85, setters :87-121, addParticleType :123-129, step counts :133-146, print :148-156,
// check :158-180 and the private checks :183-214).  Same error texts, same defaults.
#ifndef EMC_SIMULATION_PARAMETER_HPP
#define EMC_SIMULATION_PARAMETER_HPP

#include <chrono>
#include <cstdlib>
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <string>

#include <ParticleType/emcParticleType.hpp>
#include <emcMessage.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> class emcSimulationParameter {
  typedef emcParticleType<T, DeviceType> ParticleType;
  typedef std::map<SizeType, std::unique_ptr<ParticleType>> MapIdxToParticleTypes;

  T simTime, stepTime, transientTime; // [s]
  SizeType nrCarriersPerPart = 1;
  SizeType nrStepsBetweenShowProgress = 100;
  SizeType nrStepsForFinalAvg = 100;
  std::string namePrefix;
  T (*adaptPotentialForWrite)(const T &, const DeviceType &) = nullptr;
  MapIdxToParticleTypes particleTypes;
  // unseeded runs differ from run to run, like the reference's
  // (the environment variable EMCGPU_SEED replaces the clock: an unmodified main() that never calls setSeed becomes
  // reproducible, which the statistical tests of the drop-in use)
  SizeType seedRNG = defaultSeed();
  static SizeType defaultSeed() {
    if (const char *e = std::getenv("EMCGPU_SEED"))
      if (*e)
        return static_cast<SizeType>(std::strtoull(e, nullptr, 10));
    return std::chrono::high_resolution_clock::now().time_since_epoch().count();
  }

  static void error(const char *text) { emcMessage::getInstance().addError(text).print(); }
  void checkTimes() const {
    if (stepTime < 0 || simTime < 0 || transientTime < 0)
      error("The time parameter can't be negative!");
    if (stepTime == 0)
      error("Step Time can't be zero!");
    if (stepTime > simTime || transientTime > simTime)
      error("stepTime, avgTime and transientTime have to be smaller than simTime!");
  }
  void checkNrCarriersPerParticles() const {
    if (nrCarriersPerPart < 1)
      error("Nr. of Carriers per simulated particle has to be at least 1!");
  }

public:
  emcSimulationParameter() : emcSimulationParameter(10e-12, 1.5e-16, 3e-12, 1) {}
  emcSimulationParameter(T inSimTime, T inStepTime, T inTransientTime)
      : emcSimulationParameter(inSimTime, inStepTime, inTransientTime, 1) {}
  emcSimulationParameter(T inSimTime, T inStepTime, T inTransientTime, SizeType inNrCarriersPerPart)
      : simTime(inSimTime), stepTime(inStepTime), transientTime(inTransientTime), nrCarriersPerPart(inNrCarriersPerPart) {
    checkTimes();
    checkNrCarriersPerParticles();
  }

  void setTimes(T inSimTime, T inStepTime, T inTransientTime) {
    simTime = inSimTime;
    stepTime = inStepTime;
    transientTime = inTransientTime;
    checkTimes();
  }
  void setSeed(SizeType inSeed) { seedRNG = inSeed; }
  void setSimTime(T inSimTime) { simTime = inSimTime; }
  void setStepTime(T inStepTime) { stepTime = inStepTime; }
  void setTransientTime(T inTransientTime) { transientTime = inTransientTime; }
  void setNrCarriersPerPart(SizeType inNrCarriersPerParticle) {
    nrCarriersPerPart = inNrCarriersPerParticle;
    checkNrCarriersPerParticles();
  }
  void setNamePrefix(std::string inNamePrefix) { namePrefix = inNamePrefix; }
  void setAdaptPotentialForWriteFunction(T (*inAdaptPotentialForWrite)(const T &, const DeviceType &)) {
    adaptPotentialForWrite = inAdaptPotentialForWrite;
  }
  void setNrStepsBetweenShowProgress(SizeType nrSteps) { nrStepsBetweenShowProgress = nrSteps; }
  void setNrStepsForFinalAvg(SizeType nrSteps) { nrStepsForFinalAvg = nrSteps; }

  // particle types are indexed in the order they are added
  template <class DerivedParticleType>
  typename std::enable_if<std::is_base_of<ParticleType, DerivedParticleType>::value>::type
  addParticleType(std::unique_ptr<DerivedParticleType> &&newParticleType) {
    newParticleType->check();
    const SizeType idx = particleTypes.size();
    particleTypes[idx] = std::move(newParticleType);
  }

  SizeType getNrParticleTypes() const { return particleTypes.size(); }
  SizeType getNrSteps() const { return std::ceil(simTime / stepTime); }
  SizeType getNrTransientSteps() const { return std::ceil(transientTime / stepTime); }
  SizeType getNrNonTransientSteps() const { return getNrSteps() - getNrTransientSteps(); }
  bool isTransientStep(SizeType nrStep) const { return nrStep < getNrTransientSteps(); }

  void print() const {
    std::cout << "Simulation Parameter ...\n"
              << "\tTotal Simulation Time:\t" << simTime << " s\n"
              << "\tStep Time:\t\t" << stepTime << " s\n"
              << "\tTransient Time:\t\t" << transientTime << " s\n"
              << "\tNr. of Steps:\t\t" << getNrSteps() << "\n"
              << "\tNr. of Steps for Avg:\t" << getNrNonTransientSteps() << "\n";
  }

  void check() const {
    checkTimes();
    if (nrStepsForFinalAvg > getNrSteps())
      error("nrStepsForFinalAvg has to be smaller than total step nr.!");
    if (particleTypes.empty())
      error("Add at least one particleType to simulation parameter!");
    bool anyMoved = false;
    for (const auto &entry : particleTypes)
      anyMoved = anyMoved || entry.second->isMoved();
    if (!anyMoved)
      error("Add at least one moving particleType to simulation parameter!");
  }

  template <class, class> friend class emcSimulationResults;
  template <class, class, class, class, class> friend class emcSimulation;
};

#endif
