// mif_kernels.h -- host-callable launchers of the libmifgpu kernels (definitions in mif_stencil.cu and
// mif_poisson.cu).  Everything here is internal to the library; the public surface is include/mifgpu.h.
#pragma once

#include "mif_common.cuh"

namespace mifgpu {

struct Vec3 {
  real *c[3];
};
struct CVec3 {
  const real *c[3];
};

// ---- mif_stencil.cu ------------------------------------------------------------------------------

// A sub-range of the planes a stencil launcher covers (0-based within that launcher's own plane range); count < 0: all.
// Used to run the interior planes of a kernel while the halo exchange it depends on is still in flight (mif_api.cu).
struct PlaneRange {
  int first = 0, count = -1;
};

// RK stage kernels (src/Timestep.cpp:10-54).  stage = 1 (Y2), 2 (Y3), 3 (U*).
//   stage 1: in = velocity,          a = velocity_buffer (write Y2),           b = velocity_buffer_2 (write R1)
//   stage 2: in = velocity_buffer,   a = velocity_buffer_2 (read R1, write Y3), b = velocity (write a2*R2)
//   stage 3: in = velocity_buffer_2, a = velocity (read a2*R2, write U*),       b unused
void launch_stage(cudaStream_t stream, const Geom &g, int stage, CVec3 in, const real *pressure, Vec3 a, Vec3 b,
                  uint64_t *launches, PlaneRange planes = PlaneRange());

// Dirichlet faces of all three components (src/VelocityTensor.cpp:36-218), then the single-rank
// periodic ghost copies (src/StaggeredTensor.cpp:221-257).
void launch_apply_bc(cudaStream_t stream, const Geom &g, Vec3 vel, const BcDev &bc, uint64_t *launches);
void launch_periodic(cudaStream_t stream, const Geom &g, real *field, int comp, uint64_t *launches);

// rhs = div(velocity)/dt on owner points (src/PressureEquation.cpp:59-61, include/VelocityDivergence.h:9-20).
void launch_divergence(cudaStream_t stream, const Geom &g, CVec3 vel, real inv_dt_unused, real dt, real *rhs,
                       uint64_t *launches, PlaneRange planes = PlaneRange());
// rhs(face) +-= 2 g / h on the six faces (src/PressureEquation.cpp:10-56); tables as in BcDev.
void launch_nhn_rhs(cudaStream_t stream, const Geom &g, real *rhs, const BcDev &bc, uint64_t *launches);

// p += dp on all points and vel -= dt_s * grad(dp) on interior points (src/Timestep.cpp:66-81).
void launch_correct(cudaStream_t stream, const Geom &g, Vec3 vel, real *pressure, const real *dp, real dt_s,
                    uint64_t *launches, PlaneRange planes = PlaneRange());

// Stages of the velocity-only integrator with the manufactured forcing (src/TimestepVelocity.cpp:20-50):
//   stage 1: in = velocity,        rhs_buf written,       out = velocity_buffer
//   stage 2: in = velocity_buffer, rhs_buf read + written, out = velocity (read + written)
//   stage 3: in = velocity,        rhs_buf read,          out = velocity_buffer
void launch_velocity_stage(cudaStream_t stream, const Geom &g, int stage, CVec3 in, Vec3 rhs_buf, Vec3 out, real time,
                           real Re, uint64_t *launches);

// Diagnostics (src/Norms.cpp:11-118, src/PressureEquation.cpp:288-343) against an analytic family evaluated on the
// device.  `partial` receives 4 doubles per CTA (diag_blocks CTAs): velocity {sum |e|, sum |e|^2, max, 0}, pressure
// {sum |e|, sum e^2, max |e|, sum e} with e = exact - field; the host adds them up in CTA order.
int diag_blocks(const Geom &g, bool velocity);
void launch_velocity_error(cudaStream_t stream, const Geom &g, CVec3 vel, const BcDev &bc, real *partial,
                           uint64_t *launches);
void launch_pressure_error(cudaStream_t stream, const Geom &g, const real *p, const BcDev &bc, real *partial,
                           uint64_t *launches);
void launch_add_constant(cudaStream_t stream, const Geom &g, real *p, real difference, uint64_t *launches);

// ---- mif_poisson.cu ------------------------------------------------------------------------------

struct PoissonPlan;  // transform tables + eigenvalues for one context
PoissonPlan *poisson_plan_create(const Geom &g, const int n_points[3], const int periodic[3], const double h[3],
                                 const int n_global[3]);
void poisson_plan_destroy(PoissonPlan *plan);
// -1, or the direction whose line length the generic kernel cannot hold in shared memory (Bluestein transform lengths
// above 4096, i.e. more than 4097 points that are not 2^k + 1).
int poisson_plan_unsupported_direction(const PoissonPlan *plan);
// One sweep of the in-place spectral solve on the owner region of `field` (src/PressureEquation.cpp:65-264):
// dir = 0/1/2 (x/y/z); mode = 0 forward, 1 inverse + normalisation, 2 forward, eigenvalue division, inverse.
// The solve is the sequence (0,0) (1,0) (2,2) (1,1) (0,1).
void launch_poisson_sweep(cudaStream_t stream, const Geom &g, PoissonPlan *plan, real *field, int dir, int mode,
                          uint64_t *launches);
// The fused z sweep (mode 2) on a z pencil zbuf[z][y_local][x] (rows of g.PX doubles) that holds all z points of
// the y rows [y_offset, y_offset + ny_local) of the transform domain (multi-GPU slab decomposition).
void launch_poisson_zpencil(cudaStream_t stream, const Geom &g, PoissonPlan *plan, real *zbuf, int ny_local,
                            int y_offset, bool has_origin, uint64_t *launches);

// One sweep on a pencil buffer of the Py x Pz decomposition (rows of `pitch` doubles, nx_local of them used):
// dir = 1: y pencil buf[z_local][y (all)][x_local], n_outer = local z planes; dir = 2: z pencil
// buf[z (all)][y_local][x_local], n_outer = local y rows.  x_offset / y_offset: global transform indices of the
// first local x column / y row (eigenvalues, origin mode).
void launch_poisson_pencil(cudaStream_t stream, PoissonPlan *plan, real *buf, int dir, int mode, int nx_local, int pitch,
                           int n_outer, int x_offset, int y_offset, bool has_origin, uint64_t *launches);

// Peer-memory variant of the slab <-> pencil exchange: zbuf[r] / xfer[r] are the pencil and slab staging buffers of
// every rank, mapped into this process (CUDA IPC).  which = 0: forward y sweep whose results are stored straight into
// the z pencils of the owning GPUs; 1: fused z sweep on the local pencil, results stored into the slab staging of
// the owning GPUs; 2: inverse y sweep reading the local slab staging, writing `field`.
struct PeerLayout {
  int rank, nranks;
  int ylo[9], zlo[9];
  real *zbuf[8], *xfer[8];
};
bool poisson_peer_capable(const PoissonPlan *plan);
void launch_poisson_sweep_peer(cudaStream_t stream, const Geom &g, PoissonPlan *plan, real *field, const PeerLayout &peer,
                               int which, uint64_t *launches);

// Slab <-> z-pencil repacking (mif_stencil.cu).  ylo[r], ylo[r+1] delimit the y rows of rank r; the send buffer is
// ordered [dest][z_local][y in dest's range][x] with rows of g.PX doubles.
void launch_pack_slab(cudaStream_t stream, const Geom &g, const real *field, real *send, const int *ylo_dev,
                      int nranks, uint64_t *launches);
void launch_unpack_slab(cudaStream_t stream, const Geom &g, real *field, const real *recv, const int *ylo_dev,
                        int nranks, uint64_t *launches);

}  // namespace mifgpu
