// mif_common.cuh -- shared definitions of libmifgpu (sm_100a).
//
// Device data layout ("uniform padded grid"): every field of a context (u, v, w of the three velocity
// triples, p, delta-p) is stored with the SAME pitches
//     idx(i, j, k) = i + j*PX + k*PX*PY,   PX = round_up(Nx_staggered, 16), PY = Ny_staggered,
// x fastest as in the reference (include/Tensor.h:232-238), so one index serves all arrays of a stencil
// and every x row starts on a 128-byte boundary.  Host arrays keep the reference's compact extents;
// upload/download convert (mif_api.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mifgpu {

// The scalar type of the fields: the reference's `Real` (include/Real.h:9-17).  libmifgpu.so is the USE_DOUBLE=1 build;
// -DMIFGPU_FP32 builds libmifgpu_f32.so, the USE_DOUBLE=0 build, from the same sources (fields, tables and arithmetic
// in float; the tuned FP64 transform kernels are left out and every sweep runs on the generic kernel).
#ifdef MIFGPU_FP32
typedef float real;
typedef float2 real2;
__host__ __device__ __forceinline__ real2 make_real2(real x, real y) { return make_float2(x, y); }
#else
typedef double real;
typedef double2 real2;
__host__ __device__ __forceinline__ real2 make_real2(real x, real y) { return make_double2(x, y); }
#endif
#ifdef MIFGPU_RC_PLAIN
#define RC(x) (x)
#else
#define RC(x) ((::mifgpu::real)(x))
#endif  // a floating literal in the build's scalar type

// Derived constants of mif::Constants (src/Constants.cpp:63-101), passed to kernels by value.
struct Geom {
  int sx[4], sy[4], sz[4];  // local extents of u, v, w, p tensors (src/StaggeredTensor.cpp:5-9)
  int Nx, Ny, Nz;           // local unstaggered extents (ghosts included)
  int PX, PY, PZ;           // device pitches / allocated z planes
  long long plane;          // PX*PY
  long long volume;         // PX*PY*PZ
  int periodic[3];
  int base_i, base_j, base_k;
  int prev_y, next_y, prev_z, next_z;  // neighbouring ranks or -1 (src/Constants.cpp:98-101)
  // owner range of pressure points (include/StaggeredTensorMacros.h:41-83)
  int own_lo[3], own_hi[3];
  real min_x, min_y, min_z;
  real dx, dy, dz;
  real dt;
  real one_over_dx, one_over_dy, one_over_dz;
  real one_over_2_dx, one_over_2_dy, one_over_2_dz;
  real one_over_8_dx, one_over_8_dy, one_over_8_dz;
  real one_over_dx2_Re, one_over_dy2_Re, one_over_dz2_Re;
  real dx_over_2, dy_over_2, dz_over_2;
};

__host__ __device__ inline long long gidx(const Geom &g, int i, int j, int k) {
  return (long long)i + (long long)j * g.PX + (long long)k * g.plane;
}

// Boundary-data descriptor handed to the face kernels.
struct BcDev {
  int kind;            // mifgpu_bc_kind
  real time;
  real Re;
  const real *tables[3][6];  // MIFGPU_BC_HOST_CALLBACK: device copies of the host-filled faces
};

}  // namespace mifgpu
