// Norms.h -- drop-in for include/Norms.h:19-45 (host-side diagnostics over the downloaded fields).
#ifndef NORMS_H
#define NORMS_H

#include "VectorFunction.h"
#include "VelocityTensor.h"

namespace mif {

Real ErrorL1Norm(const VelocityTensor &velocity, const TimeVectorFunction &exact_velocity, Real time);
Real ErrorL2Norm(const VelocityTensor &velocity, const TimeVectorFunction &exact_velocity, Real time);
Real ErrorLInfNorm(const VelocityTensor &velocity, const TimeVectorFunction &exact_velocity, Real time);
Real ErrorL1Norm(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time);
Real ErrorL2Norm(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time);
Real ErrorLInfNorm(const StaggeredTensor &pressure, const std::function<Real(Real, Real, Real, Real)> &exact_pressure, Real time);

// Rank-0 accumulation of per-rank norms (src/Norms.cpp:122-163); with one rank they return local_error.
Real accumulate_error_mpi_l1(Real local_error, const Constants &constants);
Real accumulate_error_mpi_l2(Real local_error, const Constants &constants);
Real accumulate_error_mpi_linf(Real local_error, const Constants &constants);

}  // namespace mif

#endif  // NORMS_H
