// Test program (tests/test_host_sl.py): the angle-resolved single-layer mechanism classes of the drop-in host API on a
// parabolic and a non-parabolic single-layer valley -- scattering rates on an energy grid and the device sampler
// descriptors, printed with 17 significant digits for the comparison with the oracle.
#include <cstdio>
#include <memory>
#include <vector>

#include <ScatterMechanisms/emc2DChargedImpurityScatterMechanism.hpp>
#include <ScatterMechanisms/emcFroehlichInteractionSingleLayer.hpp>
#include <ScatterMechanisms/emcPiezoelectricSingleLayerScatterMechanism.hpp>
#include <ScatterMechanisms/emcRemoteSurfaceOpticalPhononMechanism.hpp>
#include <ScatterMechanisms/emcScreenedIntravalleyOpticalMechanism.hpp>
#include <ScatterMechanisms/emcSurfaceRoughnessScatterMechanism.hpp>
#include <ValleyTypes/emcNonParabolicIsotropSingleLayerValley.hpp>
#include <ValleyTypes/emcParabolicIsotropSingleLayerValley.hpp>

using T = double;

int main() {
  std::vector<std::unique_ptr<emcAbstractValley<T>>> valleys;
  valleys.push_back(std::make_unique<emcParabolicIsotropSingleLayerValley<T>>(0.48, constants::me, 1));
  valleys.push_back(std::make_unique<emcNonParabolicIsotropSingleLayerValley<T>>(0.47, constants::me, 6, 0.94));
  const T temperature = 300., qs = 2.5e8, rho = 3.1e-6;
  std::vector<std::unique_ptr<emcScatterMechanism<T>>> mechs;
  for (SizeType v = 0; v < 2; v++) {
    mechs.push_back(std::make_unique<emc2DChargedImpurityScatterMechanism<T>>(v, 1e16, 4.0, qs, 4.0e-9, 1.0e-9, 2.0));
    mechs.push_back(std::make_unique<emc2DChargedImpurityScatterMechanism<T>>(v, 2e15, 1.0, 0.0));
    mechs.push_back(std::make_unique<emcSurfaceRoughnessScatterMechanism<T>>(v, 3e8, 3.0e-10, 1.5e-9, qs));
    mechs.push_back(std::make_unique<emcRemoteSurfaceOpticalPhononMechanism<T>>(v, 0.0484, 0.06, 5.0e-10, temperature, false, qs));
    mechs.push_back(std::make_unique<emcRemoteSurfaceOpticalPhononMechanism<T>>(v, 0.0484, 0.06, 5.0e-10, temperature, true, 0.0));
    mechs.push_back(std::make_unique<emcScreenedIntravalleyOpticalMechanism<T>>(v, 5.8e10, rho, temperature, 0.048, false, qs));
    mechs.push_back(std::make_unique<emcScreenedIntravalleyOpticalMechanism<T>>(v, 5.8e10, rho, temperature, 0.048, true, qs));
    mechs.push_back(std::make_unique<emcFroehlichInteractionAbsorptionSL<T>>(v, 0.048, 0.4e-10, 5.41e-10, temperature, "", qs));
    mechs.push_back(std::make_unique<emcFroehlichInteractionEmissionSL<T>>(v, 0.048, 0.4e-10, 5.41e-10, temperature, "", 0.0));
    mechs.push_back(std::make_unique<emcPiezoelectricSingleLayerMechanism<T>>(v, 3.0e-11, 5.41e-10, rho, 4.2e3, temperature, "TA", qs));
  }
  for (auto &m : mechs) {
    m->setPtrValley(valleys);
    const auto d = m->deviceSampler(0);
    std::printf("mech %s valley %zu sampler %d final %zu param %.17g %.17g %.17g\n", m->getName().c_str(), (size_t)m->getIdxValley(),
                d.samplerId, (size_t)d.finalValley, d.param[0], d.param[1], d.param[2]);
    for (int i = 0; i < 48; i++)
      std::printf("%.17g\n", m->getScatterRate((i + 1) * (0.48 / 48), 0));
  }
  return 0;
}
