// PressureEquation.h -- drop-in for include/PressureEquation.h:10-26.
#ifndef PRESSURE_EQUATION_H
#define PRESSURE_EQUATION_H

#include "PressureTensor.h"
#include "VelocityTensor.h"

namespace mif {

// lap(pressure) = div(velocity)/dt with homogeneous Neumann or periodic conditions, solved on the GPU
// (mifgpu_solve_pressure).  pressure_buffer is unused scratch, see PressureTensor.h.
void solve_pressure_equation_homogeneous_periodic(StaggeredTensor &pressure, PressureTensor &pressure_buffer,
                                                  const VelocityTensor &velocity, Real dt);

// The same with non-homogeneous Neumann data dp/dn = exact_pressure_gradient on all six faces.
void solve_pressure_equation_non_homogeneous_neumann(StaggeredTensor &pressure, PressureTensor &pressure_buffer,
                                                     const VelocityTensor &velocity,
                                                     const VectorFunction &exact_pressure_gradient, Real dt);

// Shift the pressure by the mean difference to the exact one (src/PressureEquation.cpp:288-343).
void adjust_pressure(StaggeredTensor &pressure, const std::function<Real(Real, Real, Real)> &exact_pressure);
// Extension: the same with the time-dependent function passed unfrozen.  When it is p_exact (Manufactured.h) the
// reduction and the shift run on the device (mifgpu_adjust_pressure) and the tensor is not downloaded.
void adjust_pressure(StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time);

}  // namespace mif

#endif  // PRESSURE_EQUATION_H
