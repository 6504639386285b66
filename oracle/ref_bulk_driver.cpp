// TEST INFRASTRUCTURE (oracle/). Not product code, never shipped, never timed
// as the product.
//
// Recorder around the UNMODIFIED reference bulk path.  It is compiled only in
// the development container (where /root/reference exists) by oracle/Makefile
// into oracle/_ref/, with
//   * -I oracle/_ref/overlay in front of -I /root/reference/include, where the
//     overlay holds a build-time generated emcUtil.hpp whose only change is
//     `typedef RecordingRNG emcRNG` (reference: include/emcUtil.hpp:15), and
//   * -fno-access-control so that the private ensemble / table members can be
//     dumped bit-exactly (reference: examples/bulkSimulation/
//     basicBulkParticleHandler.hpp:58-59, include/emcScatterHandler.hpp:57-60).
// The step loop that runs is the reference's own moveParticles()
// (basicBulkParticleHandler.hpp:181-225) and its own observables (:289-347).
//
// Output: a flat container of named arrays, read by oracle/make_golden.py.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <ParticleType/emcElectron.hpp>
#include <ScatterMechanisms/emcAcousticScatterMechanism.hpp>
#include <ScatterMechanisms/emcCoulombScatterMechanism.hpp>
#include <ScatterMechanisms/emcFirstOrderInterValleyScatterMechanism.hpp>
#include <ScatterMechanisms/emcZeroOrderInterValleyScatterMechanism.hpp>
#include <ValleyTypes/emcNonParabolicAnistropValley.hpp>
#include <ValleyTypes/emcNonParabolicIsotropValley.hpp>
#include <ValleyTypes/emcParabolicAnisotropValley.hpp>
#include <ValleyTypes/emcParabolicIsotropValley.hpp>
#include <emcDevice.hpp>
#include <emcGrainScatterMechanism.hpp>

#include <basicBulkParticleHandler.hpp> // -I $(REF)/examples/bulkSimulation
// single-layer MoS2 (examples/singleLayerMoS2): the example's own particle type and its default parameter set
#include <electron2D.hpp>       // -I $(REF)/examples/singleLayerMoS2
#include <parameterPilotto.hpp> //
#include <parameterKaasbjerg.hpp>

using T = double;
using DeviceType = emcDevice<T, 3>;
using Handler = basicBulkParticleHandler<T, DeviceType>;
using TypeMap = Handler::MapIdxToParticleTypes;
using SubMap = std::map<SizeType, std::vector<SizeType>>;

// ---------------------------------------------------------------- container
struct Blob {
  std::ofstream os;
  explicit Blob(const std::string &path) : os(path, std::ios::binary) {}
  void put(const std::string &name, char dtype, const void *data,
           const std::vector<std::uint64_t> &dims, size_t elemSize) {
    std::uint32_t nl = name.size();
    os.write((const char *)&nl, 4);
    os.write(name.data(), nl);
    os.write(&dtype, 1);
    std::uint32_t nd = dims.size();
    os.write((const char *)&nd, 4);
    std::uint64_t n = 1;
    for (auto d : dims) {
      os.write((const char *)&d, 8);
      n *= d;
    }
    os.write((const char *)data, n * elemSize);
  }
  void f64(const std::string &n, const std::vector<double> &v,
           std::vector<std::uint64_t> dims = {}) {
    if (dims.empty())
      dims = {v.size()};
    put(n, 'd', v.data(), dims, 8);
  }
  void i64(const std::string &n, const std::vector<std::int64_t> &v,
           std::vector<std::uint64_t> dims = {}) {
    if (dims.empty())
      dims = {v.size()};
    put(n, 'q', v.data(), dims, 8);
  }
  void u64(const std::string &n, const std::vector<std::uint64_t> &v) {
    put(n, 'Q', v.data(), {v.size()}, 8);
  }
};

// ------------------------------------------------------- logging decorator
// Wraps a reference mechanism through the public plugin ABI
// (include/ScatterMechanisms/emcScatterMechanism.hpp:17-53) and logs which
// particle was scattered by which table entry.
static std::int64_t g_step = 0;
static const emcParticle<T> *g_base = nullptr;
static std::vector<std::int64_t> g_events; // (step, particle, mechId) triples

struct LoggingMechanism : public emcScatterMechanism<T> {
  std::unique_ptr<emcScatterMechanism<T>> inner;
  std::int64_t id;
  LoggingMechanism(std::unique_ptr<emcScatterMechanism<T>> &&in,
                   std::int64_t inId)
      : emcScatterMechanism<T>(in->getIdxValley()), inner(std::move(in)),
        id(inId) {}
  T getScatterRate(T energy, SizeType idxRegion) const override {
    return inner->getScatterRate(energy, idxRegion);
  }
  void scatterParticle(emcParticle<T> &p, emcRNG &rng) const override {
    g_events.push_back(g_step);
    g_events.push_back(g_base ? (&p - g_base) : -1);
    g_events.push_back(id);
    inner->scatterParticle(p, rng);
  }
  std::string getName() const override { return inner->getName(); }
  void check() override {
    inner->ptrValley = this->ptrValley; // needs -fno-access-control
    inner->check();
  }
};

static std::int64_t g_nextMech = 0;
template <class M, class PT>
void addLogged(PT &type, const std::vector<int> &regions,
               std::unique_ptr<M> &&m) {
  std::unique_ptr<emcScatterMechanism<T>> base(m.release());
  type->addScatterMechanism(
      regions, std::make_unique<LoggingMechanism>(std::move(base), g_nextMech++));
}

// ------------------------------------------------------------- scenarios
struct Args {
  std::string out = "ref.bin", material = "si", mechs = "acoustic,zero,first";
  int cells = 2, steps = 100, levels = 1000, snapEvery = 0;
  double box = 1e-7, doping = 1e23, field = 1e6, dt = 1e-16, emax = 1.0,
         temperature = 300, grainRate = 0, grainProb = 0.5, sheetDensity = 0;
  double fdir[3] = {-1, 0, 0};
  unsigned long seed = 7;
};

static bool has(const std::string &list, const std::string &item) {
  std::stringstream ss(list);
  std::string tok;
  while (std::getline(ss, tok, ','))
    if (tok == item)
      return true;
  return false;
}

// Silicon numbers: examples/SiliconFunctions.hpp:21-51 (parameter source).
static emcMaterial<T> siMaterial() {
  return emcMaterial<T>{11.8, 2329., 1.45e16, 9040, 1.15};
}

template <class PT>
void buildSilicon(PT &type, DeviceType &device, const std::string &mechs) {
  using V = emcNonParabolicAnisotropValley<T>;
  auto v = std::make_unique<V>(std::array<T, 3>{0.916, 0.196, 0.196},
                               type->getMass(), 3, 0.5);
  v->setSubValleyEllipseCoordSystem(0, {1, 0, 0}, {0, 1, 0}, {0, 0, 1});
  v->setSubValleyEllipseCoordSystem(1, {0, 1, 0}, {1, 0, 0}, {0, 0, 1});
  v->setSubValleyEllipseCoordSystem(2, {0, 0, 1}, {0, 1, 0}, {1, 0, 0});
  type->addValley(std::move(v));
  const SubMap g = {{0, {0}}, {1, {1}}, {2, {2}}};
  const SubMap f = {{0, {1, 1, 2, 2}}, {1, {0, 0, 2, 2}}, {2, {0, 0, 1, 1}}};
  const std::vector<int> reg = {0};
  if (has(mechs, "acoustic"))
    addLogged(type, reg,
              std::make_unique<emcAcousticScatterMechanism<T>>(0, 9., device));
  if (has(mechs, "coulomb"))
    addLogged(type, reg,
              std::make_unique<emcCoulombScatterMechanism<T, DeviceType>>(
                  0, 11.8, device));
  if (has(mechs, "zero")) {
    using A = emcZeroOrderInterValleyAbsorptionScatterMechanism<T>;
    using E = emcZeroOrderInterValleyEmissionScatterMechanism<T>;
    addLogged(type, reg, std::make_unique<A>("F", 0, f, 5.23e10, 0.06, device));
    addLogged(type, reg, std::make_unique<E>("F", 0, f, 5.23e10, 0.06, device));
    addLogged(type, reg, std::make_unique<A>("G", 0, g, 5.23e10, 0.06, device));
    addLogged(type, reg, std::make_unique<E>("G", 0, g, 5.23e10, 0.06, device));
  }
  if (has(mechs, "first")) {
    using A = emcFirstOrderInterValleyAbsorptionScatterMechanism<T>;
    using E = emcFirstOrderInterValleyEmissionScatterMechanism<T>;
    addLogged(type, reg, std::make_unique<A>("F", 0, f, 2.5, 0.023, device));
    addLogged(type, reg, std::make_unique<E>("F", 0, f, 2.5, 0.023, device));
    addLogged(type, reg, std::make_unique<A>("G", 0, g, 4., 0.018, device));
    addLogged(type, reg, std::make_unique<E>("G", 0, g, 4., 0.018, device));
  }
}

// Synthetic four-valley material: one valley of every 3-D valley class
// (include/ValleyTypes/*.hpp) with non-zero valley offsets, so that valley
// changes, non-axis-aligned ellipsoid frames and every dispersion variant are
// exercised by the reference's own code.  Numbers are invented for coverage.
template <class PT>
void buildMixed(PT &type, DeviceType &device, const std::string &mechs) {
  const T m0 = type->getMass();
  type->addValley(
      std::make_unique<emcNonParabolicIsotropValley<T>>(0.067, m0, 1, 0.61, 0.));
  {
    auto v = std::make_unique<emcNonParabolicAnisotropValley<T>>(
        std::array<T, 3>{1.9, 0.075, 0.11}, m0, 4, 0.46, 0.05);
    v->setSubValleyEllipseCoordSystem(0, {1, 1, 1}, {-1, 1, 0}, {-1, -1, 2});
    v->setSubValleyEllipseCoordSystem(1, {-1, 1, 1}, {1, 1, 0}, {1, -1, 2});
    v->setSubValleyEllipseCoordSystem(2, {1, -1, 1}, {1, 1, 0}, {-1, 1, 2});
    v->setSubValleyEllipseCoordSystem(3, {1, 1, -1}, {1, 0, 1}, {-1, 2, 1});
    type->addValley(std::move(v));
  }
  type->addValley(
      std::make_unique<emcParabolicIsotropValley<T>>(0.3, m0, 2, 0.03));
  {
    auto v = std::make_unique<emcParabolicAnisotropValley<T>>(
        std::array<T, 3>{0.9, 0.2, 0.3}, m0, 3, 0.08);
    v->setSubValleyEllipseCoordSystem(0, {1, 0, 0}, {0, 1, 0}, {0, 0, 1});
    v->setSubValleyEllipseCoordSystem(1, {0, 1, 0}, {1, 0, 0}, {0, 0, 1});
    v->setSubValleyEllipseCoordSystem(2, {0, 0, 1}, {0, 1, 0}, {1, 0, 0});
    type->addValley(std::move(v));
  }
  const int deg[4] = {1, 4, 2, 3};
  const std::vector<int> reg = {0};
  using ZA = emcZeroOrderInterValleyAbsorptionScatterMechanism<T>;
  using ZE = emcZeroOrderInterValleyEmissionScatterMechanism<T>;
  using FA = emcFirstOrderInterValleyAbsorptionScatterMechanism<T>;
  using FE = emcFirstOrderInterValleyEmissionScatterMechanism<T>;
  for (SizeType vi = 0; vi < 4; vi++) {
    if (has(mechs, "acoustic"))
      addLogged(type, reg,
                std::make_unique<emcAcousticScatterMechanism<T>>(vi, 7. + vi,
                                                                 device));
    if (has(mechs, "coulomb"))
      addLogged(type, reg,
                std::make_unique<emcCoulombScatterMechanism<T, DeviceType>>(
                    vi, 11.8, device));
    for (SizeType vf = 0; vf < 4; vf++) {
      if (vf == vi)
        continue;
      // every initial sub-valley may reach every final sub-valley
      SubMap sm;
      for (int si = 0; si < deg[vi]; si++)
        for (int sf = 0; sf < deg[vf]; sf++)
          sm[si].push_back(sf);
      std::string sfx = std::to_string(vi) + std::to_string(vf);
      if (has(mechs, "zero")) {
        addLogged(type, reg,
                  std::make_unique<ZA>(sfx, vi, vf, sm, 6e10, 0.03, device));
        addLogged(type, reg,
                  std::make_unique<ZE>(sfx, vi, vf, sm, 6e10, 0.03, device));
      }
      if (has(mechs, "first") && ((vi + vf) % 2 == 1)) {
        addLogged(type, reg,
                  std::make_unique<FA>(sfx, vi, vf, sm, 3.0, 0.02, device));
        addLogged(type, reg,
                  std::make_unique<FE>(sfx, vi, vf, sm, 3.0, 0.02, device));
      }
    }
  }
}

// Single-layer MoS2 exactly as examples/singleLayerMoS2/singleLayerMoS2.cpp:97-104 sets it up with its default parameter
// set (Pilotto: K valleys isotropic, Q valleys anisotropic with in-plane frames, acoustic + zero-order intervalley
// mechanisms of the single-layer classes).  The mechanisms are added by the example's own helper functions; the logging
// decorators are put around them afterwards (private vector, -fno-access-control).
// sheetDensity > 0 (material mos2ps): the K -> K Gamma-phonon pair is the free-carrier-screened class
// (emcScreenedIntravalleyOpticalMechanism, parameterPilotto.hpp:114-131), the optional argument of the example's helper.
template <class PT> void buildMoS2Pilotto(PT &type, double temperature, double sheetDensity = 0) {
  MoS2Pilotto::addValleys(type);
  MoS2Pilotto::addAcousticScatterMechanisms(type, {0}, temperature);
  MoS2Pilotto::addZeroOrderIntervalleyScatterMechanisms(type, {0}, temperature, sheetDensity);
  auto &mechs = type->scatterHandler.scatterMechanisms;
  for (size_t i = 0; i < mechs.size(); i++) {
    std::unique_ptr<emcScatterMechanism<T>> inner(mechs[i].release());
    mechs[i] = std::make_unique<LoggingMechanism>(std::move(inner), (std::int64_t)i);
  }
}

// The subset of the example's Kaasbjerg parameter set whose mechanisms have device samplers: ONE parabolic single-layer
// valley with one sub-valley, two acoustic branches, and the zero-order optical mechanisms through the constructor WITHOUT a
// sub-valley map (the sub-valley index is kept, no draw for it) and the first-order mechanisms likewise --
// singleLayerMoS2.cpp:70-76 without the Froehlich and piezoelectric terms.
// full = true: the whole set of singleLayerMoS2.cpp:64-77 (setKaasbjergParameter), i.e. with the Froehlich and the piezoelectric
// mechanisms; sheetDensity > 0 screens those two by the 2-D carrier gas (the optional arguments of the example's helpers).
// supported = true (material mos2kx): a supported, doped film -- the example's three OPTIONAL extrinsic helpers on top of the whole
// set (parameterKaasbjerg.hpp:272-351): charged impurities (1e16 1/m^2, eps_avg 4), interface roughness and the two remote
// surface-optical modes of HfO2 (the helper's own table: eps_inf 5.03, eps_0 23, 12.4 / 48.4 meV), screened by sheetDensity.
template <class PT> void buildMoS2KaasbjergSubset(PT &type, double temperature, bool full = false, double sheetDensity = 0, bool supported = false) {
  MoS2Kaasbjerg::addValleys(type);
  MoS2Kaasbjerg::addAcousticScatterMechanisms(type, {0}, temperature);
  MoS2Kaasbjerg::addZeroOrderIntervalleyScatterMechanisms(type, {0}, temperature);
  MoS2Kaasbjerg::addFirstOrderIntervalleyScatterMechanisms(type, {0}, temperature);
  if (full) {
    MoS2Kaasbjerg::addFroehlichScatterMechanisms(type, {0}, temperature, sheetDensity);
    MoS2Kaasbjerg::addPiezoelectricScatterMechanisms(type, {0}, temperature, sheetDensity);
  }
  if (supported) {
    MoS2Kaasbjerg::addChargedImpurityScatterMechanism(type, {0}, temperature, 1e16, sheetDensity, 4.0);
    MoS2Kaasbjerg::addSurfaceRoughnessScatterMechanism(type, {0}, temperature, sheetDensity, 4.0);
    MoS2Kaasbjerg::addRemoteSurfaceOpticalPhonon(type, {0}, temperature, sheetDensity, 5.03, 23.0, 0.0124, 0.0484);
  }
  auto &mechs = type->scatterHandler.scatterMechanisms;
  for (size_t i = 0; i < mechs.size(); i++) {
    std::unique_ptr<emcScatterMechanism<T>> inner(mechs[i].release());
    mechs[i] = std::make_unique<LoggingMechanism>(std::move(inner), (std::int64_t)i);
  }
}

// ------------------------------------------------------------------ dumps
static void dumpEnsemble(Blob &b, const std::string &prefix, Handler &h) {
  auto &parts = h.particles[0];
  auto &pos = h.positionsParticles[0];
  const size_t n = parts.size();
  std::vector<double> k(3 * n), e(n), tau(n), gtau(n), x(3 * n);
  std::vector<std::int64_t> idx(3 * n);
  for (size_t i = 0; i < n; i++) {
    for (int d = 0; d < 3; d++) {
      k[3 * i + d] = parts[i].k[d];
      x[3 * i + d] = pos[i][d];
    }
    e[i] = parts[i].energy;
    tau[i] = parts[i].tau;
    gtau[i] = parts[i].grainTau;
    idx[3 * i + 0] = parts[i].valley;
    idx[3 * i + 1] = parts[i].subValley;
    idx[3 * i + 2] = parts[i].region;
  }
  b.f64(prefix + "k", k, {n, 3});
  b.f64(prefix + "pos", x, {n, 3});
  b.f64(prefix + "energy", e);
  b.f64(prefix + "tau", tau);
  b.f64(prefix + "grainTau", gtau);
  b.i64(prefix + "idx", idx, {n, 3});
}

int main(int argc, char **argv) {
  Args a;
  for (int i = 1; i + 1 < argc; i += 2) {
    std::string key = argv[i], val = argv[i + 1];
    if (key == "--out") a.out = val;
    else if (key == "--material") a.material = val;
    else if (key == "--mechs") a.mechs = val;
    else if (key == "--cells") a.cells = std::stoi(val);
    else if (key == "--steps") a.steps = std::stoi(val);
    else if (key == "--levels") a.levels = std::stoi(val);
    else if (key == "--snap-every") a.snapEvery = std::stoi(val);
    else if (key == "--box") a.box = std::stod(val);
    else if (key == "--doping") a.doping = std::stod(val);
    else if (key == "--field") a.field = std::stod(val);
    else if (key == "--dt") a.dt = std::stod(val);
    else if (key == "--emax") a.emax = std::stod(val);
    else if (key == "--temperature") a.temperature = std::stod(val);
    else if (key == "--seed") a.seed = std::stoul(val);
    else if (key == "--grain-rate") a.grainRate = std::stod(val); // emcGrainScatterMechanism, [1/s]; 0: none
    else if (key == "--grain-prob") a.grainProb = std::stod(val); // its transmission probability
    else if (key == "--sheet-density") a.sheetDensity = std::stod(val); // mos2kf: 2-D carrier density [1/m^2] that screens
    else if (key == "--fdir")
      std::sscanf(val.c_str(), "%lf,%lf,%lf", &a.fdir[0], &a.fdir[1], &a.fdir[2]);
    else {
      std::cerr << "unknown option " << key << "\n";
      return 2;
    }
  }

  std::vector<std::uint64_t> draws;
  RecordingRNG::sink() = &draws;

  const T h = a.box / a.cells;
  const bool mos2 = a.material == "mos2" || a.material == "mos2ps" || a.material == "mos2k" || a.material == "mos2kf" || a.material == "mos2kx";
  // mos2: one layer of 0.65 nm (singleLayerMoS2.cpp:44-45), the placeholder material and doping of :133-135
  const T boxZ = mos2 ? 0.65e-9 : a.box, hZ = mos2 ? 0.65e-9 : h;
  DeviceType device{mos2 ? emcMaterial<T>{1, 1, 1, 1, 1} : siMaterial(), {a.box, a.box, boxZ}, {h, h, hZ}, a.temperature};
  device.addConstantDopingRegion({0, 0, 0}, {a.box, a.box, boxZ}, mos2 ? 1 : a.doping);

  TypeMap types;
  if (mos2) {
    types[0] = std::make_unique<electron2D<T, DeviceType>>(); // 5000 levels up to 0.5 eV, 4 particles per grid point
    a.levels = 5000;
    a.emax = 0.5;
    if (a.material == "mos2" || a.material == "mos2ps")
      buildMoS2Pilotto(types[0], a.temperature, a.material == "mos2ps" ? a.sheetDensity : 0);
    else
      buildMoS2KaasbjergSubset(types[0], a.temperature, a.material != "mos2k", a.sheetDensity, a.material == "mos2kx");
  } else {
    types[0] = std::make_unique<emcElectron<T, DeviceType>>(a.levels, a.emax, false);
    if (a.material == "si")
      buildSilicon(types[0], device, a.mechs);
    else if (a.material == "mixed")
      buildMixed(types[0], device, a.mechs);
    else {
      std::cerr << "unknown material\n";
      return 2;
    }
  }

  if (a.grainRate > 0)
    types[0]->setGrainScatterMechanism(std::make_unique<emcGrainScatterMechanism<T>>(a.grainProb, a.grainRate));

  Handler handler(device, types, {a.fdir[0], a.fdir[1], a.fdir[2]}, a.field,
                  a.seed);
  Blob blob(a.out);

  // ---- valley constants as the reference computes them
  auto &type = types[0];
  const size_t nV = type->getNrValleys();
  {
    std::vector<double> vc, rot;
    std::vector<std::int64_t> deg;
    for (size_t v = 0; v < nV; v++) {
      auto val = type->getValley(v);
      auto vogt = val->getVogtTransformationFactor();
      vc.insert(vc.end(), {val->getEffMassCond(), val->getEffMassDOS(),
                           val->getNonParabolicity(), val->getBottomEnergy(),
                           vogt[0], vogt[1], vogt[2]});
      deg.push_back(val->getDegeneracyFactor());
      for (size_t s = 0; s < 8; s++) {
        for (int r = 0; r < 3; r++) {
          std::array<T, 3> unit = {0, 0, 0};
          unit[r] = 1;
          std::array<T, 3> row = {0, 0, 0};
          if (s < val->getDegeneracyFactor())
            row = val->transformToDeviceCoord(s, unit); // row r of R_s
          rot.insert(rot.end(), row.begin(), row.end());
        }
      }
    }
    blob.f64("valley_consts", vc, {nV, 7});
    blob.i64("valley_deg", deg);
    blob.f64("valley_rot", rot, {nV, 8, 9});
  }

  // ---- normalised cumulative tables + tau (private members)
  {
    auto &sh = type->scatterHandler;
    for (auto &[key, tables] : sh.scatterTables) {
      const size_t nM = tables.size();
      std::vector<double> flat;
      std::vector<std::int64_t> mechIds;
      for (size_t m = 0; m < nM; m++) {
        flat.insert(flat.end(), tables[m].begin(), tables[m].end());
        mechIds.push_back(sh.idxTableToIdxMech.at(key)[m]);
      }
      std::string sfx = "_v" + std::to_string(std::get<0>(key)) + "_r" +
                        std::to_string(std::get<1>(key));
      blob.f64("cum" + sfx, flat, {nM, (std::uint64_t)a.levels});
      blob.i64("mech" + sfx, mechIds);
      blob.f64("tau" + sfx, {sh.tau.at(key)});
    }
    // un-normalised single-mechanism rates through the public ABI
    const size_t nMech = sh.scatterMechanisms.size();
    std::vector<double> rates;
    const T dE = a.emax / a.levels;
    for (size_t m = 0; m < nMech; m++)
      for (int l = 0; l < a.levels; l++)
        rates.push_back(sh.scatterMechanisms[m]->getScatterRate((l + 1) * dE, 0));
    blob.f64("raw_rates", rates, {nMech, (std::uint64_t)a.levels});
  }

  // ---- initial ensemble
  handler.generateInitialParticles();
  g_base = handler.particles[0].data();
  const size_t n = handler.getNrParticles(0);
  blob.u64("draws_init_count", {draws.size()});
  dumpEnsemble(blob, "init_", handler);

  // ---- step loop: exactly bulkSimulation.cpp:150-157
  std::vector<double> obs; // per step: E[v], vd[v], occ[v]
  std::vector<std::uint64_t> drawCount;
  auto pushObs = [&]() {
    auto e = handler.getAvgEnergy(0);
    auto v = handler.getAvgDriftVelocity(0);
    auto o = handler.getValleyOccupationProbability(0);
    obs.insert(obs.end(), e.begin(), e.end());
    obs.insert(obs.end(), v.begin(), v.end());
    obs.insert(obs.end(), o.begin(), o.end());
  };
  pushObs();
  for (int step = 1; step <= a.steps; step++) {
    g_step = step;
    handler.moveParticles(a.dt);
    pushObs();
    drawCount.push_back(draws.size());
    if (a.snapEvery > 0 && step % a.snapEvery == 0 && step != a.steps)
      dumpEnsemble(blob, "snap" + std::to_string(step) + "_", handler);
  }
  dumpEnsemble(blob, "final_", handler);
  blob.f64("obs", obs, {(std::uint64_t)a.steps + 1, 3, nV});
  blob.u64("draw_count_after_step", drawCount);
  blob.u64("draws", draws);
  blob.i64("events", g_events, {g_events.size() / 3, 3});
  blob.f64("params", {a.box, (double)a.cells, a.doping, a.field, a.dt,
                      (double)a.steps, (double)a.levels, a.emax, a.temperature,
                      a.fdir[0], a.fdir[1], a.fdir[2], (double)a.seed,
                      (double)n});
  std::cout << "ref_bulk_driver: " << n << " particles, " << draws.size()
            << " draws, " << g_events.size() / 3 << " events -> " << a.out
            << "\n";
  return 0;
}
