// TimestepVelocity.h -- drop-in for include/TimestepVelocity.h:15-16.
#ifndef MIF_B200_TIMESTEP_VELOCITY_H
#define MIF_B200_TIMESTEP_VELOCITY_H

#include "VelocityTensor.h"

namespace mif {

// One three-stage step of the momentum equation alone with the analytic forcing of ManufacturedVelocity.h, on the GPU
// (mifgpu_timestep_velocity).  Same argument meaning as the reference: `velocity` holds the new solution on return
// (the final swap_data of src/TimestepVelocity.cpp:89 is part of the call), the other two tensors are scratch.
void timestep_velocity(VelocityTensor &velocity, VelocityTensor &velocity_buffer, VelocityTensor &rhs_buffer,
                       const TimeVectorFunction &exact_velocity, Real t_n);

}  // namespace mif

#endif  // MIF_B200_TIMESTEP_VELOCITY_H
