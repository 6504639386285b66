// Plug-in interface of one bulk scattering mechanism.
// Interface mirrored: reference include/ScatterMechanisms/emcScatterMechanism.hpp:17-53
// (getScatterRate, scatterParticle, getName, check, setPtrValley, getIdxValley;
// protected idxValley / ptrValley).
//
// How a mechanism reaches the GPU: its RATES are tabulated on the host by
// emcScatterHandler through getScatterRate(), exactly as in the reference, and
// the cumulative tables are uploaded.  Its FINAL-STATE SAMPLER runs on the
// device as a hand-written device function selected by ID; the additive virtual
// deviceSampler() names that function and its parameters.  The default names
// none: such a mechanism is rejected by emcgpu_set_tables with an error that
// carries getName() (EMCGPU_E_UNSUPPORTED_MECHANISM) -- scatterParticle() is
// never used as a CPU fallback by the particle handlers.
#ifndef EMC_SCATTER_MECHANISM_HPP
#define EMC_SCATTER_MECHANISM_HPP

#include <map>
#include <memory>
#include <string>
#include <vector>

#include <ValleyTypes/emcAbstractValley.hpp>
#include <emcParticle.hpp>
#include <emcUtil.hpp>

// what emcgpu_mech_t (include/emcgpu.h) needs to know about a mechanism
struct emcDeviceSamplerDesc {
  int samplerId = 0; // EMCGPU_SAMPLER_*; 0 = no device sampler
  SizeType finalValley = 0;
  std::map<SizeType, std::vector<SizeType>> finalSubValleys; // initial sub-valley -> candidates
  double param[4] = {0., 0., 0., 0.};
};

template <class T> class emcPhononBath; // emcPhononBath.hpp

template <class T> struct emcScatterMechanism {
  typedef emcAbstractValley<T> AbstractValley;

protected:
  SizeType idxValley;                      // valley the mechanism is attached to
  std::vector<AbstractValley *> ptrValley; // all valleys of the particle type

public:
  explicit emcScatterMechanism(SizeType inIdxValley) : idxValley(inIdxValley) {}
  virtual ~emcScatterMechanism() = default;

  // scattering rate [1/s] at kinetic energy [eV] in a doping region
  virtual T getScatterRate(T energy, SizeType idxRegion) const = 0;
  // final state of one event (host reference implementation of the sampler)
  virtual void scatterParticle(emcParticle<T> &particle, emcRNG &rng) const = 0;
  virtual std::string getName() const = 0;
  virtual void check() {}

  // device final-state sampler of this mechanism in a region; valid after the tables were built
  virtual emcDeviceSamplerDesc deviceSampler(SizeType /*idxRegion*/) const { return emcDeviceSamplerDesc(); }

  // phonon bath whose event counters this mechanism feeds on the device (polar-optical hot-phonon mechanisms)
  virtual emcPhononBath<T> *devicePhononBath() const { return nullptr; }

  void setPtrValley(std::vector<std::unique_ptr<AbstractValley>> &inPtrValley) {
    for (auto &v : inPtrValley)
      ptrValley.emplace_back(v.get());
  }
  SizeType getIdxValley() const { return idxValley; }
};

#endif
