// Nonlinear Poisson solver (successive over-relaxation) -- GPU resident.
// Interface mirrored: reference include/PoissonSolver/emcSORSolver.hpp (ctor: device, accuracy [V],
// omega; calcEquilibriumPotential :49-128, calcNonEquilibriumPotential :131-197).  Default: red-black ordering on a
// thread-block cluster (same equation, relaxation factor and stopping rule as the reference, 7x faster on the MOSFET grid,
// potential equal to the lexicographic one within the solver's accuracy).  setRedBlackOrdering(false) -- or the environment
// variable EMCGPU_SOR_ORDER=lexicographic -- selects the reference's own update order (pipelined wavefront sweep,
// emc_device_run.cuh sorRowsKernel): iterates and sweep counts are then the reference's up to the last bits of exp().
//
// Inside emcSimulation the solver works on the device-resident grids of the particle handler's
// context (attach()); called on its own it creates a private context and moves the grids there and
// back.  calcBackgroundPotential (:200-284) belongs to the FMM path, which is out of scope.
#ifndef EMC_SOR_SOLVER_HPP
#define EMC_SOR_SOLVER_HPP

#include <PoissonSolver/emcAbstractSolver.hpp>
#include <detail/emcDeviceFlatten.hpp>
#include <emcGpuBinding.hpp>
#include <emcSurface.hpp>

template <class T, class DeviceType, class ParticleHandler>
class emcSORSolver : public emcAbstractSolver<T, DeviceType, ParticleHandler> {
  static const SizeType Dim = DeviceType::Dimension;
  typedef emcGrid<T, Dim> GridType;

  T accuracyVolt;
  T omega;
  const DeviceType &device;
  emcgpu_ctx *ctx = nullptr;
  bool ownsContext = false;
  int lastSweeps = 0;
  bool redBlack = defaultRedBlack();

  static bool defaultRedBlack() {
    const char *e = std::getenv("EMCGPU_SOR_ORDER");
    return !(e && (e[0] == 'l' || e[0] == 'L' || e[0] == '0'));
  }

  void needContext() {
    if (ctx)
      return;
    const char *e = std::getenv("EMCGPU_DEVICE");
    if (emcgpu_create(e ? std::atoi(e) : 0, &ctx) != EMCGPU_OK)
      emcMessage::getInstance()
          .addError(std::string("emcSORSolver: cannot create the GPU context: ") + emcgpu_last_error(nullptr))
          .print();
    ownsContext = true;
    emcdetail::FlatDevice<T, Dim> flat(device);
    emcgpu::require(ctx, emcgpu_device_configure(ctx, &flat.desc, 0., 1., nullptr, EMCGPU_MATH_EXACT),
                    "emcgpu_device_configure");
  }
  void solve(GridType &pot, const GridType *eConc, bool resetBC) {
    needContext();
    emcgpu::require(ctx, emcgpu_device_set_grid(ctx, EMCGPU_GRID_POTENTIAL, pot.raw()), "emcgpu_device_set_grid");
    if (eConc)
      emcgpu::require(ctx, emcgpu_device_set_grid(ctx, EMCGPU_GRID_CONCENTRATION, eConc->raw()), "emcgpu_device_set_grid");
    int32_t sweeps = 0;
    emcgpu::require(ctx, emcgpu_set_option(ctx, "sor_order", redBlack ? 1 : 0), "emcgpu_set_option");
    emcgpu::require(ctx, emcgpu_device_poisson(ctx, eConc ? 0 : 1, accuracyVolt, omega, resetBC ? 1 : 0, &sweeps),
                    "emcgpu_device_poisson");
    lastSweeps = sweeps;
    emcgpu::require(ctx, emcgpu_device_get_grid(ctx, EMCGPU_GRID_POTENTIAL, pot.raw()), "emcgpu_device_get_grid");
  }

public:
  emcSORSolver(const DeviceType &inDevice, const T inAccuracy = 1e-5, const T inOmega = 1.5)
      : accuracyVolt(inAccuracy), omega(inOmega), device(inDevice) {}
  emcSORSolver(const emcSORSolver &) = delete;
  ~emcSORSolver() {
    if (ownsContext && ctx)
      emcgpu_destroy(ctx);
  }

  void calcEquilibriumPotential(GridType &pot, const DeviceType & /*device*/, bool resetBC = true) override {
    solve(pot, nullptr, resetBC);
  }
  void calcNonEquilibriumPotential(GridType &pot, const DeviceType & /*device*/, const GridType &eConc,
                                   bool resetBC = true) override {
    solve(pot, &eConc, resetBC);
  }
  void calcBackgroundPotential(GridType &, const DeviceType &, ParticleHandler &, bool = true) override {
    emcMessage::getInstance()
        .addError("emcSORSolver::calcBackgroundPotential belongs to the FMM particle-particle path, which has no GPU "
                  "implementation.")
        .print();
  }

  // --- additive: what emcSimulation needs to run the solver on the handler's device-resident grids ---
  T getAccuracy() const { return accuracyVolt; } // [V]
  T getOmega() const { return omega; }
  int getLastNrSweeps() const { return lastSweeps; }
  // order of the relaxation sweeps.  true (default): red-black ordering -- same equation, relaxation factor and stopping
  // rule, fully parallel; the potential agrees with the lexicographic one to about the accuracy of the solver.
  // false: the reference's lexicographic Gauss-Seidel order -- iterates and sweep counts are the reference's.
  void setRedBlackOrdering(bool on) { redBlack = on; }
  bool getRedBlackOrdering() const { return redBlack; }
  // use (not own) the context of a GPU particle handler that was configured for the same device
  void attach(emcgpu_ctx *shared) {
    if (ownsContext && ctx)
      emcgpu_destroy(ctx);
    ctx = shared;
    ownsContext = false;
  }
};

#endif
