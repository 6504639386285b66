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
#include "../examples/hotPhononGa2O3/Ga2O3Model.hpp"

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

struct Ga2O3 {
  Device device;
  std::unique_ptr<Electron> electrons;
  std::vector<std::shared_ptr<emcPhononBath<double>>> baths;
  std::shared_ptr<emcPlasmonScreening<double>> screening;
  explicit Ga2O3(const emchost_ga2o3_spec &s)
      : device(Ga2O3Model::material<double>(), {s.box, s.box, s.box}, {s.box / 2., s.box / 2., s.box / 2.}, s.temperature),
        electrons(std::make_unique<Electron>(s.nLevels, s.maxEnergy, false)) {
    const Ga2O3Model::Parameters p;
    device.addConstantDopingRegion({0, 0, 0}, {s.box, s.box, s.box}, s.doping);
    electrons->scatterHandler.writeRateFiles = false;
    electrons->scatterHandler.reportTau = false;
    if (s.polar >= 2) {
      Ga2O3Model::Parameters q = p;
      q.tauLO = s.tauLO;
      q.tauAc = s.tauAc;
      auto setup = Ga2O3Model::addBandAndScattering<double>(electrons, device, s.polar == 3, s.multimode != 0, s.screening != 0,
                                                            s.qResolved != 0, s.qResolvedAngle != 0, s.impurity != 0,
                                                            s.acousticBath != 0, s.temperature, s.box * s.box * s.box, q);
      baths = setup.baths;
      screening = setup.screening;
    } else {
      // the unscreened classes, assembled like Ga2O3Functions.hpp:173-222 (full static permittivity for every mode)
      typedef std::map<SizeType, std::vector<SizeType>> SubValleyMap;
      electrons->addValley(std::make_unique<emcNonParabolicIsotropValley<double>>(p.relEffMass, electrons->getMass(), 1, p.alpha));
      const std::vector<int> regions = {0};
      electrons->addScatterMechanism(regions, std::make_unique<emcAcousticScatterMechanism<double>>(0, p.sigmaAc, device));
      const SubValleyMap same = {{0, {0}}};
      electrons->addScatterMechanism(regions, std::make_unique<emcZeroOrderInterValleyAbsorptionScatterMechanism<double>>(
                                                  "NPO", 0, same, p.defPotNPO, p.hwNPO, device));
      electrons->addScatterMechanism(regions, std::make_unique<emcZeroOrderInterValleyEmissionScatterMechanism<double>>(
                                                  "NPO", 0, same, p.defPotNPO, p.hwNPO, device));
      if (s.impurity)
        electrons->addScatterMechanism(regions, std::make_unique<emcCoulombScatterMechanism<double, Device>>(0, p.epsLo, device));
      const std::vector<double> modes = s.multimode ? p.hwModes : std::vector<double>{p.hwPOP};
      screening = std::make_shared<emcPlasmonScreening<double>>(p.epsLo, s.screening != 0);
      for (SizeType m = 0; m < modes.size(); m++) {
        const std::string suffix = "Ga2O3-" + std::to_string(m);
        if (s.polar == 1) {
          baths.push_back(std::make_shared<emcPhononBath<double>>(p.nrPhononBins, p.dqBin, s.tauLO, modes[m], s.temperature,
                                                                  s.box * s.box * s.box, s.acousticBath != 0, modes[m] / 2., s.tauAc));
          electrons->addScatterMechanism(regions, std::make_unique<emcHotPhononFroehlichAbsorption3D<double>>(
                                                      0, modes[m], p.relEffMass, p.epsHi, p.epsLo, baths[m], suffix));
          electrons->addScatterMechanism(regions, std::make_unique<emcHotPhononFroehlichEmission3D<double>>(
                                                      0, modes[m], p.relEffMass, p.epsHi, p.epsLo, baths[m], suffix));
        } else {
          electrons->addScatterMechanism(regions, std::make_unique<emcFroehlichAbsorption3D<double>>(
                                                      0, modes[m], p.relEffMass, p.epsHi, p.epsLo, s.temperature, "Ga2O3"));
          electrons->addScatterMechanism(regions, std::make_unique<emcFroehlichEmission3D<double>>(
                                                      0, modes[m], p.relEffMass, p.epsHi, p.epsLo, s.temperature, "Ga2O3"));
        }
      }
    }
    screening->update(s.doping, s.temperature);
    for (auto &b : baths)
      b->setScreeningQ2(screening->getQs2());
    electrons->initScatterTables();
  }
  void copyTables(double *cum, int nLevels) const {
    const auto &set = electrons->scatterHandler.getTableSets().at({0, 0});
    for (size_t i = 0; i < set.cum.size(); i++)
      std::copy(set.cum[i].begin(), set.cum[i].end(), cum + i * nLevels);
  }
};

} // namespace

extern "C" {

int emchost_ga2o3_host_loop(const emchost_ga2o3_spec *spec, int nSteps, double dt, const double *counts, const double *meanEnergy,
                            double *cumInitial, double *cumFinal, double *tauSeries, double *meanNq, double *finalNq) {
  if (!spec || nSteps < 0)
    return -EMCGPU_E_INVALID;
  Ga2O3 g(*spec);
  const size_t nB = g.baths.size(), bins = nB ? g.baths[0]->nrBins : 0;
  if (cumInitial)
    g.copyTables(cumInitial, spec->nLevels);
  for (int s = 0; s < nSteps; s++) {
    bool stale = false;
    if (spec->screening) {
      g.screening->update(spec->doping, 2. * meanEnergy[s] * constants::q / (3. * constants::kB));
      for (auto &b : g.baths)
        b->setScreeningQ2(g.screening->getQs2());
      stale = true;
    }
    for (size_t b = 0; b < nB; b++) {
      const double *em = counts + ((size_t)(s * nB + b) * 2 + 0) * bins, *ab = em + bins;
      for (size_t i = 0; i < bins; i++) {
        g.baths[b]->nEm[i] += em[i];
        g.baths[b]->nAbs[i] += ab[i];
      }
      g.baths[b]->update(dt);
      stale = true;
    }
    if (stale)
      g.electrons->reinitScatterTables();
    if (tauSeries)
      tauSeries[s] = g.electrons->getTau(0, 0);
    if (meanNq)
      for (size_t b = 0; b < nB; b++)
        meanNq[s * nB + b] = g.baths[b]->getMeanNq();
  }
  if (cumFinal)
    g.copyTables(cumFinal, spec->nLevels);
  if (finalNq)
    for (size_t b = 0; b < nB; b++)
      std::copy(g.baths[b]->Nq.begin(), g.baths[b]->Nq.end(), finalNq + b * bins);
  return static_cast<int>(g.electrons->scatterHandler.getTableSets().at({0, 0}).cum.size());
}

int emchost_ga2o3_upload(emcgpu_ctx *ctx, const emchost_ga2o3_spec *spec) {
  if (!ctx || !spec)
    return EMCGPU_E_INVALID;
  Ga2O3 g(*spec);
  emcgpu::uploadParticleType(ctx, *g.electrons);
  return EMCGPU_OK;
}

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
