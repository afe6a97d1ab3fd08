// PressureGradient.h -- host-side point evaluation of grad p at the staggered velocity points, for drivers that sample it
// outside the time step (test/full_test.cpp:143-146 builds the pressure-gradient error from it).  Same names, arguments and
// arithmetic as include/PressureGradient.h:9-24 of the reference: a one-sided difference over one cell.  Inside the time
// step the gradient is fused into the stage and correction kernels of libmifgpu.
#ifndef PRESSURE_GRADIENT_H
#define PRESSURE_GRADIENT_H

#include "StaggeredTensor.h"

namespace mif {

inline Real pressure_gradient_u(const StaggeredTensor &pressure, const size_t i, const size_t j, const size_t k) {
  return (pressure(i, j, k) - pressure(i - 1, j, k)) * pressure.constants.one_over_dx;
}
inline Real pressure_gradient_v(const StaggeredTensor &pressure, const size_t i, const size_t j, const size_t k) {
  return (pressure(i, j, k) - pressure(i, j - 1, k)) * pressure.constants.one_over_dy;
}
inline Real pressure_gradient_w(const StaggeredTensor &pressure, const size_t i, const size_t j, const size_t k) {
  return (pressure(i, j, k) - pressure(i, j, k - 1)) * pressure.constants.one_over_dz;
}

}  // namespace mif

#endif  // PRESSURE_GRADIENT_H
