// Host side of the null-scattering scheme: owns the scatter mechanisms of one
// particle type and tabulates their cumulative, normalised rates per (valley,
// doping region).  The tables and tau = 1/Gamma_max are what the GPU consumes
// (emcgpu_set_tables); selection and final-state sampling happen on the device.
//
// Interface mirrored: reference include/emcScatterHandler.hpp -- ctor :71-76,
// getTau :79-84, getGrainTau :87, initScatterTables :91-95, reinitScatterTables
// :100-108, addScatterMechanism :131-145, per-mechanism rate files :198-217,
// table fill :220-235 (rate at (level+1) dE, mechanisms in insertion order), cumulative
// sum + normalisation by the largest total rate :248-273, the "tau =" report :275-287,
// setSurfaceScatterMechanism :118-125 (one wall mechanism per face of the box).
// Not here on purpose: scatterParticle() (:148-170) -- the reference's CPU selection
// loop.  Its replacement is the device code behind emcgpu_bulk_step.
#ifndef EMC_SCATTER_HANDLER_HPP
#define EMC_SCATTER_HANDLER_HPP

#include <algorithm>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include <ScatterMechanisms/emcScatterMechanism.hpp>
#include <SurfaceScatterMechanisms/emcSurfaceScatterMechanism.hpp>
#include <emcGrainScatterMechanism.hpp>
#include <emcMessage.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> class emcScatterHandler {
public:
  typedef emcScatterMechanism<T> ScatterMechanism;
  // one (valley, region) key: mechanisms in insertion order, cum[m][level]
  struct TableSet {
    std::vector<SizeType> mechanisms; // indices into the handler's mechanism list
    std::vector<std::vector<T>> cum;  // filled by init/reinitScatterTables
    T tau = 2e-15;
  };
  typedef std::pair<SizeType, SizeType> ValleyRegion; // ordered by valley, then region

private:
  SizeType nrEnergyLevels;
  T maxEnergy;
  T dE;
  std::vector<std::unique_ptr<ScatterMechanism>> mechanismList;
  std::map<ValleyRegion, TableSet> sets;
  T grainTau = 1.;
  SizeType tableVersion = 0;
  std::unique_ptr<emcGrainScatterMechanism<T>> grainMechanism;
  std::vector<std::unique_ptr<emcSurfaceScatterMechanism<T, DeviceType>>> surfaceMechanisms; // [2 * Dim], null = specular

  void tabulate() {
    for (auto &[key, set] : sets) {
      set.cum.assign(set.mechanisms.size(), std::vector<T>());
      for (SizeType m = 0; m < set.mechanisms.size(); m++) {
        auto &row = set.cum[m];
        row.reserve(nrEnergyLevels);
        for (SizeType level = 0; level < nrEnergyLevels; level++)
          row.push_back(mechanismList[set.mechanisms[m]]->getScatterRate((level + 1) * dE, key.second));
      }
    }
    if (grainMechanism)
      grainTau = 1. / grainMechanism->getScatterRate();
  }

  void accumulateAndNormalise() {
    for (auto &[key, set] : sets) {
      (void)key;
      if (set.cum.empty()) {
        set.tau = 2e-15; // free flight time of a (valley, region) without mechanisms
        continue;
      }
      for (SizeType m = 1; m < set.cum.size(); m++)
        for (SizeType l = 0; l < nrEnergyLevels; l++)
          set.cum[m][l] = set.cum[m][l] + set.cum[m - 1][l];
      const T gammaMax = *std::max_element(set.cum.back().begin(), set.cum.back().end());
      for (auto &row : set.cum)
        for (auto &x : row)
          x /= gammaMax;
      set.tau = 1. / gammaMax;
    }
    if (reportTau) {
      std::cout << "Initialized ScatterHandler ...\n";
      for (const auto &[key, set] : sets)
        std::cout << "\tidxValley " << key.first << " idxRegion " << key.second << ": tau = " << set.tau << " s\n";
    }
  }

public:
  // the reference writes one "<Name><region><valley>ScatterMechanism.txt" per table into the
  // working directory on every initScatterTables() and prints tau on every (re)build; both can
  // be switched off (benchmarks, per-step table rebuilds)
  bool writeRateFiles = true;
  bool reportTau = true;

  emcScatterHandler() : emcScatterHandler(1000, 4.) {}
  emcScatterHandler(SizeType inNrEnergyLevels, T inMaxEnergy)
      : nrEnergyLevels(inNrEnergyLevels), maxEnergy(inMaxEnergy), dE(inMaxEnergy / inNrEnergyLevels),
        surfaceMechanisms(2 * DeviceType::Dimension) {}

  T getTau(SizeType idxRegion, SizeType idxValley) const {
    auto it = sets.find(ValleyRegion(idxValley, idxRegion));
    return it == sets.end() ? T(2e-15) : it->second.tau;
  }
  T getGrainTau() const { return grainTau; }

  template <class DerivedScatterMechanism>
  typename std::enable_if<std::is_base_of<ScatterMechanism, DerivedScatterMechanism>::value>::type
  addScatterMechanism(std::unique_ptr<DerivedScatterMechanism> &&mechanism, const std::vector<int> &regions) {
    const SizeType valley = mechanism->getIdxValley();
    const SizeType idx = mechanismList.size();
    mechanismList.push_back(std::move(mechanism));
    for (const int region : regions)
      sets[ValleyRegion(valley, static_cast<SizeType>(region))].mechanisms.push_back(idx);
  }

  void setGrainScatterMechanism(std::unique_ptr<emcGrainScatterMechanism<T>> &&newMechanism) {
    grainMechanism = std::move(newMechanism);
  }
  bool hasGrainScatterMechanism() const { return static_cast<bool>(grainMechanism); }
  const emcGrainScatterMechanism<T> *getGrainScatterMechanism() const { return grainMechanism.get(); }

  template <class DerivedSurfaceScatterMechanism>
  typename std::enable_if<std::is_base_of<emcSurfaceScatterMechanism<T, DeviceType>, DerivedSurfaceScatterMechanism>::value>::type
  setSurfaceScatterMechanism(std::unique_ptr<DerivedSurfaceScatterMechanism> &&newMechanism, emcBoundaryPos boundaryPosition) {
    const SizeType face = toUnderlying(boundaryPosition);
    if (face >= surfaceMechanisms.size())
      emcMessage::getInstance().addError("Index for Boundary is out of bounds.").print();
    newMechanism->setBoundaryPosition(boundaryPosition);
    surfaceMechanisms[face].reset(newMechanism.release());
  }
  // wall mechanism of a face or nullptr (default specular reflection)
  const emcSurfaceScatterMechanism<T, DeviceType> *getSurfaceScatterMechanism(SizeType face) const {
    return face < surfaceMechanisms.size() ? surfaceMechanisms[face].get() : nullptr;
  }

  // incremented by every (re)build of the tables: lets a GPU particle handler notice that it has to upload them again
  SizeType getTableVersion() const { return tableVersion; }

  void initScatterTables() {
    tableVersion++;
    tabulate();
    if (writeRateFiles)
      writeTablesToFiles();
    accumulateAndNormalise();
  }
  void reinitScatterTables() {
    tableVersion++;
    tabulate();
    accumulateAndNormalise();
  }

  // un-normalised rates, "energy rate" per line, default stream precision (as the reference)
  void writeTablesToFiles() const {
    for (const auto &[key, set] : sets) {
      for (SizeType m = 0; m < set.cum.size(); m++) {
        std::ofstream os(mechanismList[set.mechanisms[m]]->getName() + std::to_string(key.second) +
                         std::to_string(key.first) + "ScatterMechanism.txt");
        T energy = dE;
        for (const auto &rate : set.cum[m]) {
          os << energy << " " << rate << "\n";
          energy += dE;
        }
      }
    }
  }

  // --- read access for the GPU binding ------------------------------------
  SizeType getNrEnergyLevels() const { return nrEnergyLevels; }
  T getMaxEnergy() const { return maxEnergy; }
  const std::map<ValleyRegion, TableSet> &getTableSets() const { return sets; }
  const ScatterMechanism &getMechanism(SizeType idx) const { return *mechanismList[idx]; }
  SizeType getNrMechanisms() const { return mechanismList.size(); }
};

#endif
