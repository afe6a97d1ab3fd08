// Constants.h -- drop-in for the reference's mif::Constants (include/Constants.h:11-92): the same public
// const members, computed with the same formulas (src/Constants.cpp:58-101), plus the handle of the
// libmifgpu context that owns the device-side copy of them.
#ifndef CONSTANTS_H
#define CONSTANTS_H

#include <array>
#include <cstddef>

#include "Real.h"

struct mifgpu_ctx;

namespace mif {

class Constants {
public:
  const size_t Nx_global, Ny_global, Nz_global;
  const Real x_size, y_size_global, z_size_global;
  const Real min_x_global, min_y_global, min_z_global;
  const Real Re;
  const Real final_time;
  const unsigned int num_time_steps;
  const std::array<bool, 3> periodic_bc;

  const int Py, Pz, rank, y_rank, z_rank;

  const Real dt;
  const size_t Nx_domains, Ny_domains_global, Nz_domains_global;
  const Real dx, dy, dz;
  const Real one_over_2_dx, one_over_2_dy, one_over_2_dz;
  const Real one_over_8_dx, one_over_8_dy, one_over_8_dz;
  const Real one_over_dx2_Re, one_over_dy2_Re, one_over_dz2_Re;
  const Real dx_over_2, dy_over_2, dz_over_2;
  const Real one_over_dx, one_over_dy, one_over_dz;

  const int P;
  const size_t Ny_owner, Nz_owner;
  const size_t Nx, Ny, Nz;
  const size_t Nx_staggered, Ny_staggered, Nz_staggered;
  const int base_i, base_j, base_k;
  const int prev_proc_y, next_proc_y, prev_proc_z, next_proc_z;

  Constants(size_t Nx_global, size_t Ny_global, size_t Nz_global, Real x_size, Real y_size_global,
            Real z_size_global, Real min_x_global, Real min_y_global, Real min_z_global, Real Re, Real final_time,
            unsigned int num_time_steps, int Py, int Pz, int rank, const std::array<bool, 3> &periodic_bc);
  Constants(const Constants &) = delete;
  ~Constants();

  // The libmifgpu context of this geometry (created on first use on CUDA device $MIFGPU_DEVICE, default 0).
  // Throws std::runtime_error with mifgpu_last_error() if it cannot be created: there is no CPU fallback.
  mifgpu_ctx *gpu() const;

private:
  mutable mifgpu_ctx *gpu_ctx_ = nullptr;
};

}  // namespace mif

#endif  // CONSTANTS_H
