// Glue between the drop-in host API and the C ABI of the CUDA library
// (include/emcgpu.h): flattens a particle type -- valleys, cumulative rate tables,
// device sampler descriptors -- into the plain structs the ABI takes, and turns
// ABI failures into the API's error convention (emcMessage: print + abort,
// reference include/emcMessage.hpp:43-65).
//
// There is no CPU path behind this header: a valley class without a device
// dispersion or a scatter mechanism without a device sampler is an error that
// names the offender.
#ifndef EMC_GPU_BINDING_HPP
#define EMC_GPU_BINDING_HPP

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <emcgpu.h>

#include <ParticleType/emcParticleType.hpp>
#include <emcMessage.hpp>
#include <emcPhononBath.hpp>

namespace emcgpu {

inline void require(emcgpu_ctx *ctx, int status, const std::string &what) {
  if (status == EMCGPU_OK)
    return;
  emcMessage::getInstance()
      .addError(what + " failed (emcgpu status " + std::to_string(status) + "): " + emcgpu_last_error(ctx))
      .print();
}

// emcgpu_set_grain for the grain mechanism of a particle type (none: the clock is inert and stays on the host)
template <class T, class DeviceType> bool uploadGrainMechanism(emcgpu_ctx *ctx, const emcParticleType<T, DeviceType> &type) {
  const auto *grain = type.scatterHandler.getGrainScatterMechanism();
  require(ctx, emcgpu_set_grain(ctx, grain ? grain->getTransmissionProbability() : 0.5, grain ? grain->getScatterRate() : 0.),
          "emcgpu_set_grain");
  return grain != nullptr;
}

// the phonon baths the mechanisms of a particle type feed, in the order of first appearance (= device bath index)
template <class T, class DeviceType>
std::vector<emcPhononBath<T> *> collectPhononBaths(const emcParticleType<T, DeviceType> &type) {
  std::vector<emcPhononBath<T> *> baths;
  const auto &handler = type.scatterHandler;
  for (SizeType m = 0; m < handler.getNrMechanisms(); m++) {
    auto *bath = handler.getMechanism(m).devicePhononBath();
    if (bath && std::find(baths.begin(), baths.end(), bath) == baths.end())
      baths.push_back(bath);
  }
  return baths;
}

// emcgpu_set_phonon_baths: binning and the prefix sums the q-resolved polar angle is drawn from
template <class T> void uploadPhononBaths(emcgpu_ctx *ctx, const std::vector<emcPhononBath<T> *> &baths) {
  if (baths.empty())
    return;
  std::vector<double> cumW, cumWN;
  for (const auto *b : baths) {
    if (b->nrBins != baths[0]->nrBins || b->dq != baths[0]->dq)
      emcMessage::getInstance().addError("All phonon baths of a particle type must share one |q| binning on the GPU path.").print();
    cumW.insert(cumW.end(), b->cumW.begin(), b->cumW.end());
    cumWN.insert(cumWN.end(), b->cumWN.begin(), b->cumWN.end());
  }
  require(ctx,
          emcgpu_set_phonon_baths(ctx, static_cast<int>(baths.size()), static_cast<int>(baths[0]->nrBins), baths[0]->dq,
                                  cumW.data(), cumWN.data()),
          "emcgpu_set_phonon_baths");
}

// recordEmission / recordAbsorption of the step(s) since the last call: device counters -> nEm / nAbs of the baths
// sumOverRanks: sharded runs (several GPUs) add the counters of all ranks before the baths see them (nEm / nAbs
// join the per-step all-reduce); counts are integers far below 2^53, so the sum of doubles is exact and identical everywhere.
// hasParticles = false: this rank holds no particle of the type (it still takes part in the sum).
template <class T, class Reduce>
void collectPhononCounts(emcgpu_ctx *ctx, const std::vector<emcPhononBath<T> *> &baths, Reduce &&sumOverRanks,
                         bool hasParticles = true) {
  if (baths.empty())
    return;
  const SizeType bins = baths[0]->nrBins;
  std::vector<int64_t> em(baths.size() * bins, 0), ab(baths.size() * bins, 0);
  if (hasParticles)
    require(ctx, emcgpu_get_phonon_counts(ctx, em.data(), ab.data(), 1), "emcgpu_get_phonon_counts");
  std::vector<double> both(2 * em.size());
  for (SizeType i = 0; i < em.size(); i++) {
    both[i] = static_cast<double>(em[i]);
    both[em.size() + i] = static_cast<double>(ab[i]);
  }
  sumOverRanks(both);
  for (SizeType b = 0; b < baths.size(); b++)
    for (SizeType i = 0; i < bins; i++) {
      baths[b]->nEm[i] += static_cast<T>(both[b * bins + i]);
      baths[b]->nAbs[i] += static_cast<T>(both[em.size() + b * bins + i]);
    }
}
template <class T> void collectPhononCounts(emcgpu_ctx *ctx, const std::vector<emcPhononBath<T> *> &baths) {
  collectPhononCounts(ctx, baths, [](std::vector<double> &) {});
}

// emcgpu_set_valleys + emcgpu_set_phonon_baths + emcgpu_set_tables for one particle type.  To be called after
// init/reinitScatterTables(); may be repeated whenever the tables were rebuilt.
template <class T, class DeviceType> void uploadParticleType(emcgpu_ctx *ctx, const emcParticleType<T, DeviceType> &type) {
  const SizeType nValleys = type.getNrValleys();
  std::vector<emcgpu_valley_t> valleys(nValleys);
  for (SizeType v = 0; v < nValleys; v++) {
    const auto *valley = type.getValley(v);
    emcgpu_valley_t &out = valleys[v];
    std::memset(&out, 0, sizeof out);
    out.kind = valley->deviceValleyKind();
    if (out.kind < 0)
      emcMessage::getInstance()
          .addError("Valley " + std::to_string(v) + " of " + type.getName() +
                    " has no device dispersion (deviceValleyKind() < 0); it cannot run on the GPU path and "
                    "there is no CPU fallback.")
          .print();
    out.degeneracy = static_cast<int32_t>(valley->getDegeneracyFactor());
    if (out.degeneracy > EMCGPU_MAX_SUBVALLEYS)
      emcMessage::getInstance().addError("Valley degeneracy exceeds EMCGPU_MAX_SUBVALLEYS.").print();
    out.effMassCond = valley->getEffMassCond();
    out.effMassDOS = valley->getEffMassDOS();
    out.alpha = valley->getNonParabolicity();
    out.bottomEnergy = valley->getBottomEnergy();
    const auto &vogt = valley->getVogtTransformationFactor();
    for (int i = 0; i < 3; i++)
      out.vogt[i] = vogt[i];
    // column j of R_s = image of the unit vector e_j under the public transform
    for (int s = 0; s < EMCGPU_MAX_SUBVALLEYS; s++) {
      for (int j = 0; j < 3; j++) {
        std::array<T, 3> e = {0, 0, 0};
        e[j] = 1;
        const auto col = s < out.degeneracy ? valley->transformToEllipseCoord(s, e) : e;
        for (int i = 0; i < 3; i++)
          out.rot[s][3 * i + j] = col[i];
      }
    }
  }
  require(ctx, emcgpu_set_valleys(ctx, valleys.data(), static_cast<int>(nValleys)), "emcgpu_set_valleys");

  const auto &handler = type.scatterHandler;
  const auto baths = collectPhononBaths(type);
  uploadPhononBaths(ctx, baths);
  const SizeType nLevels = handler.getNrEnergyLevels();
  std::vector<emcgpu_tableset_t> sets;
  std::vector<std::vector<double>> cumStore;
  std::vector<std::vector<emcgpu_mech_t>> mechStore;
  for (const auto &[key, set] : handler.getTableSets()) {
    if (set.cum.empty())
      continue; // no mechanisms: the device uses the default tau as well
    cumStore.emplace_back();
    mechStore.emplace_back();
    auto &cum = cumStore.back();
    auto &mechs = mechStore.back();
    for (SizeType m = 0; m < set.cum.size(); m++) {
      cum.insert(cum.end(), set.cum[m].begin(), set.cum[m].end());
      const auto &mech = handler.getMechanism(set.mechanisms[m]);
      const emcDeviceSamplerDesc desc = mech.deviceSampler(key.second);
      emcgpu_mech_t out;
      std::memset(&out, 0, sizeof out);
      out.sampler = desc.samplerId;
      out.finalValley = static_cast<int32_t>(desc.finalValley);
      out.mechId = static_cast<int32_t>(set.mechanisms[m]);
      for (int i = 0; i < 4; i++)
        out.param[i] = desc.param[i];
      if (auto *bath = mech.devicePhononBath())
        out.param[2] = static_cast<double>(std::find(baths.begin(), baths.end(), bath) - baths.begin());
      std::strncpy(out.name, mech.getName().c_str(), EMCGPU_NAME_LEN - 1);
      if (!desc.finalSubValleys.empty()) {
        out.nFinal = static_cast<int32_t>(desc.finalSubValleys.begin()->second.size());
        if (out.nFinal > EMCGPU_MAX_FINAL)
          emcMessage::getInstance().addError(mech.getName() + ": too many final subvalleys for the device.").print();
        for (const auto &[from, to] : desc.finalSubValleys) {
          if (from >= EMCGPU_MAX_SUBVALLEYS || static_cast<int32_t>(to.size()) != out.nFinal)
            emcMessage::getInstance()
                .addError(mech.getName() + ": final subvalley lists must have equal length on the device.")
                .print();
          for (SizeType f = 0; f < to.size(); f++)
            out.finalSub[from][f] = static_cast<uint8_t>(to[f]);
        }
      }
      mechs.push_back(out);
    }
    emcgpu_tableset_t ts;
    std::memset(&ts, 0, sizeof ts);
    ts.valley = static_cast<int32_t>(key.first);
    ts.region = static_cast<int32_t>(key.second);
    ts.nMech = static_cast<int32_t>(set.cum.size());
    ts.tau = set.tau;
    sets.push_back(ts);
  }
  for (SizeType i = 0; i < sets.size(); i++) {
    sets[i].cum = cumStore[i].data();
    sets[i].mech = mechStore[i].data();
  }
  require(ctx, emcgpu_set_tables(ctx, sets.data(), static_cast<int>(sets.size()), static_cast<int>(nLevels),
                                 handler.getMaxEnergy()),
          "emcgpu_set_tables");
}

} // namespace emcgpu

#endif
