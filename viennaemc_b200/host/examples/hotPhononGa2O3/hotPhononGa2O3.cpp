// Velocity-field characteristic of bulk beta-Ga2O3 with equilibrium or non-equilibrium (hot) polar phonons on the
// GPU-resident drop-in API -- the scenario of the reference's examples/hotPhononGa2O3/hotPhononGa2O3.cpp: per time step
// particles move on the GPU, the phonon baths collect the emission / absorption events counted on the device, relax,
// and the Froehlich rate tables are rebuilt on the host and uploaded again.
//
//   hotPhononGa2O3 [--fields 10,50,100] (kV/cm) [--time s] [--dt s] [--box m] [--doping 1/m3] [--temp K] [--use-hpb 0|1]
//                  [--multimode 0|1] [--screening 0|1] [--qresolved 0|1] [--impurity 0|1] [--acoustic-bath 0|1]
//                  [--reinit-every n] [--steady-frac f] [--seed S] [--outdir dir] [--tag name]
// Output: <outdir>/ga2o3_vE_<tag>.txt with F[kV/cm] v[cm/s] <E>[eV] N_LO N_LO/N_0 T_LO[K] T_ac[K] (as the reference).
#include <chrono>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <ParticleType/emcElectron.hpp>
#include <basicBulkParticleHandler.hpp>
#include <emcDevice.hpp>

#include "Ga2O3Model.hpp"

using NumType = double;
using DeviceType = emcDevice<NumType, 3>;
using ParticleHandler = basicBulkParticleHandler<NumType, DeviceType>;

static NumType planckTemp(NumType energyEV, NumType N) {
  return N <= 0. ? 0. : constants::q * energyEV / (constants::kB * std::log(1. + 1. / N));
}

int main(int argc, char **argv) {
  std::string fieldList = "10,50,100,150,200,300,400", outdir = ".", tag;
  double time = 5e-12, dt = 1e-16, box = 3e-7, doping = 1e23, temperature = 300., steadyFrac = 0.4;
  long useHpb = 1, multimode = 0, screening = 0, qresolved = 0, qresAngle = 1, impurity = 0, acousticBath = 1, reinitEvery = 1;
  unsigned long seed = 1;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string key = argv[i], val = argv[i + 1];
    if (key == "--fields") fieldList = val;
    else if (key == "--time") time = std::stod(val);
    else if (key == "--dt") dt = std::stod(val);
    else if (key == "--box") box = std::stod(val);
    else if (key == "--doping") doping = std::stod(val);
    else if (key == "--temp") temperature = std::stod(val);
    else if (key == "--steady-frac") steadyFrac = std::stod(val);
    else if (key == "--use-hpb") useHpb = std::stol(val);
    else if (key == "--multimode") multimode = std::stol(val);
    else if (key == "--screening") screening = std::stol(val);
    else if (key == "--qresolved") qresolved = std::stol(val);
    else if (key == "--qres-angle") qresAngle = std::stol(val);
    else if (key == "--impurity") impurity = std::stol(val);
    else if (key == "--acoustic-bath") acousticBath = std::stol(val);
    else if (key == "--reinit-every") reinitEvery = std::max(1L, std::stol(val));
    else if (key == "--seed") seed = std::stoul(val);
    else if (key == "--outdir") outdir = val;
    else if (key == "--tag") tag = val;
    else {
      std::cerr << "unknown option " << key << "\n";
      return 2;
    }
  }
  if (tag.empty())
    tag = useHpb ? "hpb" : "eq";
  std::vector<NumType> fields;
  {
    std::stringstream ss(fieldList);
    std::string item;
    while (std::getline(ss, item, ','))
      if (!item.empty())
        fields.push_back(std::stod(item));
  }
  const Ga2O3Model::Parameters par;
  std::ofstream summary(outdir + "/ga2o3_vE_" + tag + ".txt");
  summary << "# beta-Ga2O3 velocity-field, " << (useHpb ? "non-equilibrium" : "equilibrium") << " phonons\n"
          << "# tau_LO=" << par.tauLO << " s  tau_ac=" << par.tauAc << " s  T=" << temperature << " K  n=" << doping << " m^-3\n"
          << "# F[kV/cm]  v[cm/s]  <E>[eV]  N_LO  N_LO/N_0  T_LO[K]  T_ac[K]\n";
  const auto start = std::chrono::steady_clock::now();
  double particleSteps = 0;
  for (const NumType fieldkVcm : fields) {
    const std::array<NumType, 3> maxPos = {box, box, box};
    const NumType Vsim = box * box * box;
    DeviceType device{Ga2O3Model::material<NumType>(par), maxPos, {box / 2., box / 2., box / 2.}, temperature};
    device.addConstantDopingRegion({0, 0, 0}, maxPos, doping);
    ParticleHandler::MapIdxToParticleTypes types;
    types[0] = std::make_unique<emcElectron<NumType, DeviceType>>(2000, 5., false);
    types[0]->scatterHandler.writeRateFiles = false;
    types[0]->scatterHandler.reportTau = false;
    auto polar = Ga2O3Model::addBandAndScattering<NumType>(types[0], device, useHpb != 0, multimode != 0, screening != 0,
                                                          qresolved != 0, qresAngle != 0, impurity != 0, acousticBath != 0,
                                                          temperature, Vsim, par);
    polar.screening->update(doping, temperature);
    for (auto &b : polar.baths)
      b->setScreeningQ2(polar.screening->getQs2());
    ParticleHandler handler(device, types, {-1, 0, 0});
    handler.setSeed(seed);
    handler.resetAppliedFieldStrength(fieldkVcm * 1e5);
    handler.generateInitialParticles();
    const SizeType nrSteps = std::ceil(time / dt);
    const SizeType firstSteady = static_cast<SizeType>((1. - steadyFrac) * nrSteps);
    NumType sumV = 0, sumE = 0, sumNq = 0, sumTac = 0, N0 = 0;
    SizeType nSteady = 0;
    for (SizeType m = 0; m < polar.modeEnergy.size(); m++)
      N0 += polar.modeWeight[m] / (std::exp(constants::q * polar.modeEnergy[m] / (constants::kB * temperature)) - 1.);
    for (SizeType step = 1; step <= nrSteps; step++) {
      handler.moveParticles(dt);
      const NumType v = handler.getAvgDriftVelocity(0)[0], e = handler.getAvgEnergy(0)[0];
      bool stale = false;
      if (screening) {
        polar.screening->update(doping, 2. * e * constants::q / (3. * constants::kB));
        for (auto &b : polar.baths)
          b->setScreeningQ2(polar.screening->getQs2());
        stale = true;
      }
      for (auto &b : polar.baths) {
        b->update(dt);
        stale = true;
      }
      if (stale && step % reinitEvery == 0)
        types[0]->reinitScatterTables();
      NumType nq = useHpb ? 0. : N0;
      for (SizeType m = 0; m < polar.baths.size(); m++)
        nq += polar.modeWeight[m] * polar.baths[m]->getMeanNq();
      if (step > firstSteady) {
        sumV += v;
        sumE += e;
        sumNq += nq;
        sumTac += useHpb ? polar.baths[0]->getAcousticTemp() : temperature;
        nSteady++;
      }
    }
    particleSteps += static_cast<double>(handler.getNrParticles(0)) * nrSteps;
    const NumType v = sumV / nSteady, e = sumE / nSteady, nq = sumNq / nSteady;
    summary << std::setw(8) << fieldkVcm << " " << std::scientific << std::setprecision(4) << std::abs(v) * 100. << " "
            << std::fixed << std::setprecision(4) << e << " " << nq << " " << nq / N0 << " " << std::setprecision(1)
            << planckTemp(polar.modeEnergy[0], nq) << " " << sumTac / nSteady << "\n";
    std::cout << "F = " << std::setw(6) << fieldkVcm << " kV/cm   v = " << std::scientific << std::setprecision(3)
              << std::abs(v) * 100. << " cm/s   <E> = " << std::fixed << e << " eV   N_LO = " << nq << " (x"
              << std::setprecision(2) << nq / N0 << " N_0)   " << handler.getNrParticles(0) << " particles\n";
    handler.deleteParticles();
  }
  const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
  std::cout << "wall time " << seconds << " s, " << particleSteps / seconds << " particle-steps/s\n";
  return 0;
}
