// Results of a device run: terminal currents from the per-step contact counters, potential, field and
// carrier concentration (current and averaged over the last steps).
// Interface mirrored: reference include/emcSimulationResults.hpp (ctor :70-85, updateCurrent :125-151,
// writeCurrentResults :155-172, writeFinalResults :178-204, initPotential :208-213) -- same files, same
// formats (emcOutput.hpp).
//
// The grids live on the GPU during the run (potential, concentration, field, counts and the running
// sums of updateAverageCharacteristics :87-93 / updateCurrentParticleConcentrations :98-116 are
// device kernels: accumulateKernel, concentrationKernel).  This class keeps the HOST MIRRORS that the
// file writers read; emcSimulation refreshes them from the device whenever something is written.
#ifndef EMC_SIMULATION_RESULTS_HPP
#define EMC_SIMULATION_RESULTS_HPP

#include <cmath>
#include <vector>

#include <emcConstants.hpp>
#include <emcGrid.hpp>
#include <emcOutput.hpp>
#include <emcSimulationParameter.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType> class emcSimulationResults {
  static const SizeType Dim = DeviceType::Dimension;
  typedef std::vector<std::vector<T>> CurrentMeasurement;          // [type][contact]
  typedef std::vector<std::vector<int>> NettoParticleCounter;      // [type][contact]
  typedef emcGrid<T, Dim> GridType;

  const SizeType nrPartTypes, nrContacts, nrCurrSteps;
  const emcSimulationParameter<T, DeviceType> &param;
  GridType currPot, avgPot;                                  // normalised by Vt; avgPot = SUM over nrAvgSteps
  std::vector<GridType> currConc, avgConc, eField, nrPart;   // normalised by Ni / [V/m] / carriers per grid point
  SizeType nrCummSteps = 0, nrAvgSteps = 0;
  std::vector<NettoParticleCounter> nettoPart; // per non-transient step: left - (injected - deleted)
  std::vector<CurrentMeasurement> current;     // running mean current [A] after each non-transient step
  CurrentMeasurement nettoPartSum;
  std::vector<T> currentFactor; // carriers per particle * charge / dt

public:
  emcSimulationResults() = delete;
  emcSimulationResults(const DeviceType &device, const emcSimulationParameter<T, DeviceType> &inParam)
      : nrPartTypes(inParam.getNrParticleTypes()), nrContacts(device.getSurface().getNrContacts()),
        nrCurrSteps(inParam.getNrNonTransientSteps()), param(inParam), currPot(device.getGridExtent()), avgPot(currPot),
        currConc(nrPartTypes, GridType(currPot)), avgConc(nrPartTypes, GridType(currPot)), eField(Dim, GridType(currPot)),
        nrPart(nrPartTypes, GridType(currPot)),
        nettoPart(nrCurrSteps, NettoParticleCounter(nrPartTypes, std::vector<int>(nrContacts))),
        current(nrCurrSteps, CurrentMeasurement(nrPartTypes, std::vector<T>(nrContacts))),
        nettoPartSum(nrPartTypes, std::vector<T>(nrContacts, 0)), currentFactor(nrPartTypes) {
    // first guess of the potential: local charge neutrality, asinh(doping / 2 Ni)
    typename DeviceType::SizeVec coord;
    for (coord.fill(0); !currPot.isEndCoord(coord); currPot.advanceCoord(coord))
      currPot[coord] = std::asinh(0.5 * device.getDopingProfile().getDoping(coord, true));
    for (const auto &[idxType, partType] : param.particleTypes)
      currentFactor[idxType] = param.nrCarriersPerPart * partType->getCharge() / param.stepTime;
  }

  // one non-transient step: particles that left through each contact and injected - deleted ones
  void updateCurrent(const NettoParticleCounter &nrRemPart, const NettoParticleCounter &nrInjPart) {
    if (nrCummSteps >= nrCurrSteps)
      return;
    for (SizeType t = 0; t < nrPartTypes; t++)
      for (SizeType c = 0; c < nrContacts; c++) {
        // the reference forms the difference in T and stores it as int
        const int netto = static_cast<int>(static_cast<T>(nrRemPart[t][c]) - static_cast<T>(nrInjPart[t][c]));
        nettoPart[nrCummSteps][t][c] = netto;
        nettoPartSum[t][c] += netto;
        current[nrCummSteps][t][c] = nettoPartSum[t][c] / (nrCummSteps + 1) * currentFactor[t];
      }
    nrCummSteps++;
  }

  T getAvgCurrent(SizeType idxType, SizeType idxContact) const {
    return nrCummSteps ? current[nrCummSteps - 1][idxType][idxContact] : T(0);
  }

  void writeCurrentResults(std::string nameSuffix, const DeviceType &device) const {
    writeToFile(currPot, param.namePrefix + "Potential" + nameSuffix,
                param.adaptPotentialForWrite ? param.adaptPotentialForWrite : undoNormalizationPotential<T, DeviceType>,
                device);
    for (const auto &[idxType, partType] : param.particleTypes)
      writeToFile(currConc[idxType], param.namePrefix + partType->getName() + "Conc" + nameSuffix,
                  undoNormalizationConcentration<T, DeviceType>, device);
    static const char *axis[3] = {"X", "Y", "Z"};
    for (SizeType d = 0; d < Dim; d++)
      writeToFile(eField[d], param.namePrefix + "EField" + axis[d] + nameSuffix);
  }

  void writeFinalResults(const DeviceType &device) const {
    const auto mean = [this](const GridType &sum) {
      GridType out(sum);
      for (auto &v : out)
        v = v / nrAvgSteps;
      return out;
    };
    for (const auto &[idxType, partType] : param.particleTypes) {
      const std::string prefix = param.namePrefix + partType->getName();
      writeToFile(nettoPart, current, idxType, param.stepTime, param.transientTime, prefix + "Current");
      writeToFile(mean(avgConc[idxType]), prefix + "ConcAvg", undoNormalizationConcentration<T, DeviceType>, device);
    }
    writeToFile(mean(avgPot), param.namePrefix + "PotentialAvg",
                param.adaptPotentialForWrite ? param.adaptPotentialForWrite : undoNormalizationPotential<T, DeviceType>,
                device);
  }

  template <class, class, class, class, class> friend class emcSimulation;
};

#endif
