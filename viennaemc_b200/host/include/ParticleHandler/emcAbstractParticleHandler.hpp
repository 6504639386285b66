// Interface of the particle handlers of a device run.  Interface mirrored: reference
// include/ParticleHandler/emcAbstractParticleHandler.hpp (public pure virtuals :111-174; the expected
// population of the contact cells :263-277).
#ifndef EMC_ABSTRACT_PARTICLE_HANDLER_HPP
#define EMC_ABSTRACT_PARTICLE_HANDLER_HPP

#include <chrono>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include <ParticleType/emcParticleType.hpp>
#include <emcGrid.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType, class PMScheme, SizeType Dim = DeviceType::Dimension>
class emcAbstractParticleHandler {
public:
  typedef emcParticleType<T, DeviceType> ParticleType;
  typedef std::map<SizeType, std::unique_ptr<ParticleType>> MapIdxToParticleTypes;
  typedef typename DeviceType::SizeVec SizeVec;
  typedef typename DeviceType::ValueVec ValueVec;
  typedef std::vector<std::vector<int>> NettoParticleCounter; // [particle type][contact]

protected:
  const DeviceType &device;
  PMScheme &pmScheme;
  const SizeType nrCarriersPerPart;
  MapIdxToParticleTypes &idxTypeToPartType;
  std::vector<emcGrid<T, Dim>> expNrPart; // per particle type: population the reservoir cells are kept at

  NettoParticleCounter initNettoParticleCounter() const {
    return NettoParticleCounter(idxTypeToPartType.size(), std::vector<int>(device.getSurface().getNrContacts(), 0));
  }

public:
  emcAbstractParticleHandler() = delete;
  emcAbstractParticleHandler(const DeviceType &inDevice, PMScheme &inPMScheme, SizeType inNrCarriersPerPart,
                             MapIdxToParticleTypes &inIdxToPartTypesMap)
      : device(inDevice), pmScheme(inPMScheme), nrCarriersPerPart(inNrCarriersPerPart),
        idxTypeToPartType(inIdxToPartTypesMap) {
    for (const auto &[idxType, type] : idxTypeToPartType) {
      (void)idxType;
      expNrPart.emplace_back(device.getGridExtent(), 0);
      if (!type->isInjected())
        continue;
      SizeVec coord;
      for (coord.fill(0); !device.isEndCoord(coord); device.advanceCoord(coord))
        if (device.getSurface().isReservoirContact(coord))
          expNrPart.back()[coord] = type->getExpectedNrParticlesAtContact(coord, device);
    }
  }
  virtual ~emcAbstractParticleHandler() = default;

  SizeType getNrParticleTypes() const { return idxTypeToPartType.size(); }
  virtual bool calcsPartPartInteraction() const = 0;
  virtual void printNrParticles() const = 0;
  virtual void generateInitialParticles(const emcGrid<T, Dim> &potential) = 0;
  virtual void assignParticlesToMesh(SizeType idxType, emcGrid<T, Dim> &gridNrParticles) = 0;
  virtual NettoParticleCounter driftScatterParticles(T tStep, std::vector<emcGrid<T, Dim>> &eField) = 0;
  virtual NettoParticleCounter handleOhmicContacts() = 0;
  virtual void print(std::string namePrefix, std::string nameSuffix) = 0;
  virtual T getParticlePotential(const ValueVec & /*position*/) { return 0; }
};

#endif
