// libemchost.so: the drop-in host API instantiated for the silicon model, behind
// the small C interface of include/emchost.h.
#include <algorithm>
#include <map>
#include <memory>

#include <emchost.h>

#include <ParticleType/emcElectron.hpp>
#include <detail/emcBulkEnsembleBuilder.hpp>
#include <emcDevice.hpp>
#include <emcGpuBinding.hpp>

#include "../examples/SiliconModel.hpp"

namespace {

typedef emcDevice<double, 3> Device;
typedef emcElectron<double, Device> Electron;

struct Model {
  Device device;
  std::unique_ptr<Electron> electrons;
  explicit Model(const emchost_si_spec &s)
      : device(SiliconModel::material<double>(), {s.box[0], s.box[1], s.box[2]},
               {s.spacing[0], s.spacing[1], s.spacing[2]}, s.temperature),
        electrons(std::make_unique<Electron>(s.nLevels, s.maxEnergy, false)) {
    device.addConstantDopingRegion({0, 0, 0}, {s.box[0], s.box[1], s.box[2]}, s.doping);
    SiliconModel::addXValley<double>(electrons);
    SiliconModel::addScattering<double>(electrons, device, {0}, s.mechanisms, s.coulombSecond != 0);
    electrons->scatterHandler.writeRateFiles = false;
    electrons->scatterHandler.reportTau = false;
    electrons->initScatterTables();
  }
};

} // namespace

extern "C" {

int emchost_si_upload(emcgpu_ctx *ctx, const emchost_si_spec *spec) {
  if (!ctx || !spec)
    return EMCGPU_E_INVALID;
  Model m(*spec);
  emcgpu::uploadParticleType(ctx, *m.electrons);
  return EMCGPU_OK;
}

int emchost_si_tables(const emchost_si_spec *spec, double *cum, int64_t cumCapacity, double *tau, int32_t *nMech) {
  if (!spec)
    return EMCGPU_E_INVALID;
  Model m(*spec);
  const auto &sets = m.electrons->scatterHandler.getTableSets();
  const auto it = sets.find({0, 0});
  if (it == sets.end())
    return EMCGPU_E_INVALID;
  const auto &set = it->second;
  if (nMech)
    *nMech = static_cast<int32_t>(set.cum.size());
  if (tau)
    *tau = set.tau;
  if (cum) {
    if (cumCapacity < static_cast<int64_t>(set.cum.size()) * spec->nLevels)
      return EMCGPU_E_CAPACITY;
    for (size_t i = 0; i < set.cum.size(); i++)
      std::copy(set.cum[i].begin(), set.cum[i].end(), cum + i * spec->nLevels);
  }
  return EMCGPU_OK;
}

int64_t emchost_si_initial_ensemble(const emchost_si_spec *spec, uint64_t seed, int64_t capacity, double *const *soa,
                                    uint32_t *packed, double *grainTau) {
  if (!spec)
    return -1;
  Model m(*spec);
  emcRNG rng(seed);
  emcdetail::HostEnsemble h;
  emcdetail::generateBulkEnsemble(h, *m.electrons, m.device, rng);
  const int64_t n = static_cast<int64_t>(h.size());
  if (n > capacity && (soa || packed || grainTau))
    return -1;
  if (soa)
    for (int s = 0; s < EMCGPU_N_STREAMS; s++)
      if (soa[s])
        std::copy(h.stream[s].begin(), h.stream[s].end(), soa[s]);
  if (packed)
    std::copy(h.packed.begin(), h.packed.end(), packed);
  if (grainTau)
    std::copy(h.grainTau.begin(), h.grainTau.end(), grainTau);
  return n;
}

} // extern "C"
