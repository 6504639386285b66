// Scatter mechanisms of the reference's API that exist here by NAME ONLY: their classes can be constructed with the
// reference's argument lists (so that an unmodified driver which mentions them compiles), but they have no device final-state
// sampler yet -- and the GPU particle handlers never run a mechanism on the CPU.  Adding one to a particle type stops the
// program with an error that carries the mechanism's name (check() is called by emcParticleType::addScatterMechanism,
// reference include/ParticleType/emcParticleType.hpp:141-153).
#ifndef EMC_DETAIL_NO_DEVICE_SAMPLER_HPP
#define EMC_DETAIL_NO_DEVICE_SAMPLER_HPP

#include <string>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <emcMessage.hpp>

namespace emcdetail {

template <class T> class NoDeviceSamplerMechanism : public emcScatterMechanism<T> {
  std::string name;

public:
  NoDeviceSamplerMechanism(std::string inName, SizeType inValley) : emcScatterMechanism<T>(inValley), name(std::move(inName)) {}
  std::string getName() const override { return name; }
  T getScatterRate(T, SizeType) const override { return 0; }
  void scatterParticle(emcParticle<T> &, emcRNG &) const override {}
  void check() final {
    emcMessage::getInstance()
        .addError("Scatter mechanism '" + name +
                  "' has no device final-state sampler in this library; it cannot run on the GPU path and there is no CPU "
                  "fallback.")
        .print();
  }
};

} // namespace emcdetail

#endif
