// Plug-in interface of the Poisson solver.  Interface mirrored: reference
// include/PoissonSolver/emcAbstractSolver.hpp:9-58.
#ifndef EMC_ABSTRACT_SOLVER_HPP
#define EMC_ABSTRACT_SOLVER_HPP

#include <emcGrid.hpp>
#include <emcUtil.hpp>

template <class T, class DeviceType, class ParticleHandler> struct emcAbstractSolver {
  typedef emcGrid<T, DeviceType::Dimension> GridType;
  virtual ~emcAbstractSolver() = default;
  // potentials normalised by the thermal voltage, concentrations by Ni, lengths by the intrinsic Debye length
  virtual void calcEquilibriumPotential(GridType &pot, const DeviceType &device, bool resetBC = true) = 0;
  virtual void calcNonEquilibriumPotential(GridType &pot, const DeviceType &device, const GridType &eConc,
                                           bool resetBC = true) = 0;
  virtual void calcBackgroundPotential(GridType &pot, const DeviceType &device, ParticleHandler &handler,
                                       bool resetBC = true) = 0;
};

#endif
