// Host-side creation of the initial bulk ensemble in the SoA layout of the C ABI.
// Logic mirrored: reference examples/bulkSimulation/basicBulkParticleHandler.hpp:143-159
// (cells in storage order; floor(n) particles plus one more with probability
// frac(n), the comparison draw always consumed) and :587-596 (position first,
// then the particle type's generateInitialParticle).  Separate from the handler
// so that it can be exercised without a GPU.
#ifndef EMC_DETAIL_BULK_ENSEMBLE_BUILDER_HPP
#define EMC_DETAIL_BULK_ENSEMBLE_BUILDER_HPP

#include <cstdint>
#include <random>
#include <vector>

#include <emcgpu.h>

#include <ParticleType/emcParticleType.hpp>
#include <emcGrid.hpp>
#include <emcParticleInitialization.hpp>

namespace emcdetail {

// SoA staging buffer (the layout emcgpu_set_ensemble / emcgpu_get_ensemble use)
struct HostEnsemble {
  std::vector<double> stream[EMCGPU_N_STREAMS];
  std::vector<uint32_t> packed;
  std::vector<double> grainTau; // kept for completeness; the bulk kernels only need it with a grain mechanism
  size_t size() const { return packed.size(); }
  void clear() {
    for (auto &s : stream)
      s.clear();
    packed.clear();
    grainTau.clear();
  }
};

template <class T, class DeviceType>
void appendParticle(HostEnsemble &h, emcParticleType<T, DeviceType> &type, const DeviceType &device,
                    const typename DeviceType::SizeVec &coord, emcRNG &rng) {
  const auto pos = initParticlePos(coord, device.getGridExtent(), device.getSpacing(), rng);
  emcParticle<T> part;
  if (type.isMoved())
    part = type.generateInitialParticle(coord, device, rng);
  h.stream[EMCGPU_KX].push_back(part.k[0]);
  h.stream[EMCGPU_KY].push_back(part.k[1]);
  h.stream[EMCGPU_KZ].push_back(part.k[2]);
  h.stream[EMCGPU_ENERGY].push_back(part.energy);
  h.stream[EMCGPU_TAU].push_back(part.tau);
  h.stream[EMCGPU_X].push_back(pos[0]);
  h.stream[EMCGPU_Y].push_back(pos[1]);
  h.stream[EMCGPU_Z].push_back(DeviceType::Dimension > 2 ? pos[DeviceType::Dimension - 1] : 0.);
  h.packed.push_back(EMCGPU_PACK(part.valley, part.subValley, part.region));
  h.grainTau.push_back(part.grainTau);
}

template <class T, class DeviceType>
void generateBulkEnsemble(HostEnsemble &h, emcParticleType<T, DeviceType> &type, const DeviceType &device, emcRNG &rng) {
  emcGrid<T, DeviceType::Dimension> potential(device.getGridExtent(), 0);
  std::uniform_real_distribution<T> uniform(0., 1.);
  typename DeviceType::SizeVec coord;
  for (coord.fill(0); !device.isEndCoord(coord); device.advanceCoord(coord)) {
    auto toCreate = type.getInitialNrParticles(coord, device, potential);
    while (toCreate >= 1) {
      appendParticle(h, type, device, coord, rng);
      toCreate--;
    }
    if (uniform(rng) < toCreate)
      appendParticle(h, type, device, coord, rng);
  }
}

} // namespace emcdetail

#endif
