// Timestep.h -- drop-in for include/Timestep.h:16-27 and include/TimestepVelocity.h:15-16.
#ifndef MIF_B200_TIMESTEP_H
#define MIF_B200_TIMESTEP_H

#include "PressureTensor.h"
#include "VelocityTensor.h"

namespace mif {

// One three-stage projection step on the GPU (mifgpu_timestep); same argument meaning as the reference:
// `velocity` and `pressure` hold the new solution on return, the other tensors are scratch.
void timestep(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &velocity_buffer_2,
              const TimeVectorFunction &exact_velocity, Real t_n, StaggeredTensor &pressure,
              StaggeredTensor &pressure_buffer, PressureTensor &pressure_solver_buffer);

// The variant with non-homogeneous Neumann conditions on the pressure increments.
void timestep_nhn(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &velocity_buffer_2,
                  const TimeVectorFunction &exact_velocity, const TimeVectorFunction &exact_pressure_gradient,
                  Real t_n, StaggeredTensor &pressure, StaggeredTensor &pressure_buffer,
                  PressureTensor &pressure_solver_buffer);

}  // namespace mif

#endif  // MIF_B200_TIMESTEP_H
