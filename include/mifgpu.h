/* include/mifgpu.h -- C ABI of libmifgpu, the B200 (sm_100a) implementation of the projection-method
 * time step of FattiMei/mpi-incompressible-fluid.
 *
 * The reference has no FFI layer: its boundary is a set of C++ free functions in namespace mif that take
 * references to reference-defined classes.  This header is what a thin C++
 * forwarding layer with the reference's own class and function names binds to (that layer lives in
 * mpi-incompressible-fluid_b200/host/).  Each entry point names the reference
 * interface it replaces (paths relative to the reference repository root).
 *
 * Conventions: extern "C"; plain pointers and sizes only; every function that can fail returns 0 on
 * success or a negative mifgpu_status, and mifgpu_last_error() gives the message of the last failure on
 * the calling thread; all host pointers are caller-owned; one host thread per context; host arrays use
 * exactly the reference's ghosted, x-fastest layout  idx = i + j*sx + k*sx*sy  (include/Tensor.h:232-238)
 * with the extents of src/StaggeredTensor.cpp:5-9.  There is no CPU fallback: without a CUDA device every
 * compute entry point fails with MIFGPU_ERR_CUDA.
 */
#ifndef MIFGPU_H
#define MIFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MIFGPU_ABI_VERSION 1

typedef enum mifgpu_status {
  MIFGPU_OK = 0,
  MIFGPU_ERR_INVALID = -1,     /* bad argument (the reference would assert, src/Constants.cpp:104-119) */
  MIFGPU_ERR_CUDA = -2,        /* CUDA runtime failure or no device */
  MIFGPU_ERR_UNSUPPORTED = -3, /* valid reference configuration that this build does not cover yet */
  MIFGPU_ERR_COMM = -4         /* multi-GPU communication failure */
} mifgpu_status;

/* include/StaggeredTensor.h:13-15 (enum StaggeringDirection {x, y, z, none}) */
typedef enum mifgpu_staggering {
  MIFGPU_STAGGER_X = 0, /* u */
  MIFGPU_STAGGER_Y = 1, /* v */
  MIFGPU_STAGGER_Z = 2, /* w */
  MIFGPU_STAGGER_NONE = 3 /* pressure */
} mifgpu_staggering;

/* The reference's `Real` (include/Real.h:9-17): the element type of every FIELD buffer that crosses this interface
 * (tensor uploads / downloads, boundary face tables).  libmifgpu.so is the USE_DOUBLE=1 build; libmifgpu_f32.so is
 * the USE_DOUBLE=0 build of the same sources and exports the same symbols -- compile the caller with -DMIFGPU_FP32
 * and link that library instead (mifgpu_real_bytes() tells which one a process got).  Scalars -- times, dt, Re,
 * domain sizes, norms, the values of mifgpu_allreduce / mifgpu_gather -- are double in both builds. */
#ifdef MIFGPU_FP32
typedef float mifgpu_real;
#else
typedef double mifgpu_real;
#endif

/* The 16 constructor arguments of mif::Constants (include/Constants.h:86-90, src/Constants.cpp:58-62),
 * plus the CUDA device ordinal.  All derived quantities are recomputed by the library with the same
 * formulas (src/Constants.cpp:63-101). */
typedef struct mifgpu_params {
  uint64_t Nx_global, Ny_global, Nz_global; /* pressure POINTS per direction, walls included */
  double x_size, y_size_global, z_size_global;
  double min_x_global, min_y_global, min_z_global;
  double Re;
  double final_time;
  uint32_t num_time_steps; /* dt = final_time / num_time_steps */
  int32_t Py, Pz, rank;    /* pencil decomposition: rank = y_rank*Pz + z_rank */
  int32_t periodic_bc[3];
  int32_t device;          /* CUDA device ordinal for this rank */
} mifgpu_params;

typedef struct mifgpu_ctx mifgpu_ctx;       /* replaces Constants + PressureSolverStructures */
typedef struct mifgpu_tensor mifgpu_tensor; /* device-resident StaggeredTensor */

/* Analytic boundary / forcing data evaluated on the device.  The reference passes std::function
 * bundles (include/VectorFunction.h:16-54) that are evaluated point by point inside
 * VelocityTensor::apply_bc (src/VelocityTensor.cpp:36-218); a device cannot call them, so the known
 * analytic families are enumerated and anything else goes through the host-callback kind. */
typedef enum mifgpu_bc_kind {
  MIFGPU_BC_TEST_CASE_1 = 1,     /* include/TestCaseBoundaries.h:17-35  (v = 1 on the face x = 1)    */
  MIFGPU_BC_TEST_CASE_2 = 2,     /* include/TestCaseBoundaries.h:38-56  (v = 1 on the face x = -0.5) */
  MIFGPU_BC_ETHIER_STEINMAN = 3, /* generators/manufsol.py:31-72 (u_exact, v_exact, w_exact, dp_d*_exact) */
  MIFGPU_BC_HOST_CALLBACK = 4,   /* any other TimeVectorFunction: faces filled on the host */
  MIFGPU_BC_VELOCITY_TEST = 5    /* generators/manufsol_velocity.py:55-59 (u_exact_v_test, v_exact_v_test, w_exact_v_test) */
} mifgpu_bc_kind;

/* Host callback for MIFGPU_BC_HOST_CALLBACK.  The library asks for ONE face of ONE component at one
 * time; the callee writes the final boundary values (i.e. what src/VelocityTensor.cpp:36-218 would
 * store, including the half-cell extrapolation of the wall-normal component) into `values`, a dense
 * 2-D array over the FULL extent of that tensor on that face, first listed index fastest:
 *   face 0,1 (z-, z+): values[i + j*sx]     face 2,3 (y-, y+): values[i + k*sx]
 *   face 4,5 (x-, x+): values[j + k*sy]
 * `which` is 0 for the velocity itself and 1 for the pressure-gradient data g used by the
 * non-homogeneous Neumann variant (src/PressureEquation.cpp:10-56), evaluated at the unstaggered pressure
 * points of the face.  Inside mifgpu_timestep time = t_new and time_prev = t_old of the stage, and the
 * callee must return exactly what the reference hands to the solver there, namely
 * exact_pressure_gradient.get_difference_over_time(t_new, t_old) = g(t_old) - g(t_new)
 * (src/Timestep.cpp:89-93 with src/VectorFunction.cpp:52-60: the difference is "second minus first"). */
typedef void (*mifgpu_face_callback)(void *user, int which, double time, double time_prev, int component,
                                     int face, mifgpu_real *values);

typedef struct mifgpu_bc {
  int32_t kind; /* mifgpu_bc_kind */
  double Re;    /* the reference's global `Reynolds` read by the generated exact solutions */
  mifgpu_face_callback callback;
  void *user;
} mifgpu_bc;

/* ---- context ------------------------------------------------------------------------------------ */

/* Constants::Constants (src/Constants.cpp:58-120) + PressureSolverStructures::PressureSolverStructures
 * (src/PressureSolverStructures.cpp:13-70): geometry, decomposition, transform plans, eigenvalues. */
int mifgpu_create(const mifgpu_params *params, mifgpu_ctx **ctx);
void mifgpu_destroy(mifgpu_ctx *ctx);

/* Multi-GPU: one process (or thread) per GPU, rank = params->rank of params->Py * params->Pz ranks.  Replaces the
 * MPI set-up of the reference (MPI_Init + the Cartesian communicators of deps/2Decomp_C/C2Decomp.cpp:34-62): rank 0
 * calls mifgpu_comm_unique_id, the host program distributes the MIFGPU_UNIQUE_ID_BYTES bytes to all ranks by any
 * means (MPI_Bcast, torch.distributed, a file), and every rank calls mifgpu_create_distributed collectively.
 * Halos (src/StaggeredTensor.cpp:60-165) and pencil transposes (deps/2Decomp_C/Transpose*.cpp) then run inside the
 * library over NCCL.  Py = 1 (z slabs, Pz = number of GPUs) is the fast configuration on one NVSwitch box: its Y<->Z
 * transposes are fused into the sweep kernels over peer memory.  Py > 1 gives the reference's Py x Pz pencils
 * (rank = y_rank * Pz + z_rank, src/Constants.cpp:68): two-phase halos (y sheets, then whole z planes) and the four
 * 2Decomp transposes as grouped send/recv box exchanges.  Periodic y and z directions may be distributed: the neighbours
 * wrap around (src/Constants.cpp:98-101). */
#define MIFGPU_UNIQUE_ID_BYTES 128
int mifgpu_comm_unique_id(void *unique_id);
int mifgpu_create_distributed(const mifgpu_params *params, const void *unique_id, mifgpu_ctx **ctx);

/* Host-only helper (no GPU needed): the block distribution used for both the z slabs and the y ranges of the z
 * pencils, first[r] .. first[r+1] for r < parts, bigger blocks on the low ranks (src/Constants.cpp:78-79,
 * deps/2Decomp_C/C2Decomp.cpp:273-324).  `first` has parts + 1 entries. */
int mifgpu_slab_plan(uint64_t n_points, int32_t parts, int32_t *first);
const char *mifgpu_last_error(void);
/* sizeof(mifgpu_real) of the loaded library: 8 for libmifgpu.so, 4 for libmifgpu_f32.so. */
int mifgpu_real_bytes(void);
int mifgpu_abi_version(void);

/* Local extents {sx, sy, sz} of a tensor with the given staggering on this rank
 * (src/StaggeredTensor.cpp:5-9 with src/Constants.cpp:80-85). */
int mifgpu_tensor_extents(const mifgpu_ctx *ctx, int staggering, uint64_t extents[3]);

/* ---- tensors ------------------------------------------------------------------------------------ */

/* StaggeredTensor::StaggeredTensor (src/StaggeredTensor.cpp:5-36); zero-initialised like std::vector. */
int mifgpu_tensor_create(mifgpu_ctx *ctx, int staggering, mifgpu_tensor **tensor);
void mifgpu_tensor_destroy(mifgpu_tensor *tensor);
/* Host <-> device copies of a whole tensor in the reference layout (Tensor::raw_data(), include/Tensor.h:118). */
int mifgpu_tensor_upload(mifgpu_tensor *tensor, const mifgpu_real *host);
int mifgpu_tensor_download(const mifgpu_tensor *tensor, mifgpu_real *host);
/* The index box lo[d] <= index < hi[d] of a tensor as a compact array, x fastest:
 * host[(i - lo[0]) + (j - lo[1]) * bx + (k - lo[2]) * bx * by], bx = hi[0] - lo[0], by = hi[1] - lo[1].  This is what
 * the output path needs instead of whole fields: writeVTK reads three planes, writeDat one line with its
 * interpolation neighbours (src/VTKDatExport.cpp:115-312,342-583); the box is gathered on the device and crosses
 * PCIe as one contiguous block.  MIFGPU_ERR_INVALID if the box is empty or leaves the tensor. */
int mifgpu_tensor_download_box(const mifgpu_tensor *tensor, const int32_t lo[3], const int32_t hi[3], mifgpu_real *host);
/* Tensor::swap_data (include/Tensor.h:108-110) as used by VelocityTensor::swap_data (src/VelocityTensor.cpp:13-27). */
int mifgpu_tensor_swap(mifgpu_tensor *a, mifgpu_tensor *b);

/* The same transfers without blocking the host, for callers that stream independent jobs through the device (the
 * reference has no counterpart: its tensors live in host memory).  Each direction has its own link stream, its own
 * stream for the re-pitching copy on the device and two staging buffers (the link moves the next tensor while the last
 * one is re-pitched); ordering is by events only: a transfer starts after the last compute call that used the tensor and after
 * the tensor's previous transfers, and compute calls that use the tensor afterwards wait for it on the device.  So
 * upload(A) | timestep(B) | download(C) of three different tensor sets overlap, and PCIe runs in both directions at
 * once.  `host` should be page-locked; it must not be touched until mifgpu_synchronize has returned. */
int mifgpu_tensor_upload_async(mifgpu_tensor *tensor, const mifgpu_real *host);
int mifgpu_tensor_download_async(mifgpu_tensor *tensor, mifgpu_real *host);

/* ---- the hot path ------------------------------------------------------------------------------- */

/* mif::timestep / mif::timestep_nhn (include/Timestep.h:16-27, src/Timestep.cpp:97-156): one
 * three-stage projection step.  velocity, velocity_buffer, velocity_buffer_2 are {u, v, w} triples.
 * On return `velocity` and `pressure` hold the new solution; the other tensors hold the same scratch
 * contents the reference leaves in them.  nhn != 0 selects timestep_nhn (pressure-gradient data from
 * bc->callback with which = 1, or the Ethier-Steinman gradient). */
int mifgpu_timestep(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], mifgpu_tensor *const velocity_buffer[3],
                    mifgpu_tensor *const velocity_buffer_2[3], const mifgpu_bc *bc, double t_n,
                    mifgpu_tensor *pressure, mifgpu_tensor *pressure_buffer, int nhn);

/* mif::timestep_velocity (include/TimestepVelocity.h:15-16, src/TimestepVelocity.cpp:58-90): one three-stage step of
 * the momentum equation alone, with the analytic forcing forcing_{x,y,z} of generators/manufsol_velocity.py that
 * calculate_momentum_rhs_with_forcing_* (include/MomentumEquationForcing.h:11-33) always adds; bc->Re is the
 * reference's global `Reynolds` read by that forcing.  On return `velocity` holds the new solution (the data of
 * velocity and velocity_buffer are swapped as by VelocityTensor::swap_data, src/TimestepVelocity.cpp:89);
 * velocity_buffer and rhs_buffer hold the scratch contents the reference leaves in them. */
int mifgpu_timestep_velocity(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], mifgpu_tensor *const velocity_buffer[3],
                             mifgpu_tensor *const rhs_buffer[3], const mifgpu_bc *bc, double t_n);

/* VelocityTensor::apply_bc (src/VelocityTensor.cpp:36-233) at one fixed time. */
int mifgpu_apply_bc(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], const mifgpu_bc *bc, double time);

/* mif::solve_pressure_equation_homogeneous_periodic / _non_homogeneous_neumann
 * (include/PressureEquation.h:10-21, src/PressureEquation.cpp:266-286): pressure <- solution of
 * lap(p) = div(velocity)/dt.  nhn_bc may be NULL (homogeneous / periodic); otherwise its callback is
 * asked for the Neumann data with which = 1, time = nhn_time, time_prev = nhn_time (the caller's
 * callback returns g at nhn_time when both are equal). */
int mifgpu_solve_pressure(mifgpu_ctx *ctx, mifgpu_tensor *pressure, mifgpu_tensor *const velocity[3], double dt,
                          const mifgpu_bc *nhn_bc, double nhn_time);

/* ---- diagnostics ------------------------------------------------------------------------------------ */

/* ErrorL1Norm / ErrorL2Norm / ErrorLInfNorm of the velocity (src/Norms.cpp:11-86) and of a scalar tensor
 * (src/Norms.cpp:88-118) against an analytic family evaluated on the device -- `exact->kind` is one of the
 * device-evaluated kinds; its pressure is p_exact of generators/manufsol.py:58-72 for Ethier-Steinman and 0 for
 * the two lid-driven test cases.  norms[] = {L1, L2, LInf} of THIS rank's part of the domain, exactly what the
 * reference's functions return before accumulate_error_mpi_* (src/Norms.cpp:122-162) combines the ranks.  Only
 * a few kB of per-CTA partial sums cross PCIe instead of four whole fields.  MIFGPU_ERR_UNSUPPORTED for
 * MIFGPU_BC_HOST_CALLBACK (arbitrary std::functions: download the tensors, use the host layer's norms). */
int mifgpu_velocity_error_norms(mifgpu_ctx *ctx, mifgpu_tensor *const velocity[3], const mifgpu_bc *exact, double time,
                                double norms[3]);
int mifgpu_pressure_error_norms(mifgpu_ctx *ctx, const mifgpu_tensor *pressure, const mifgpu_bc *exact, double time,
                                double norms[3]);
/* mif::adjust_pressure (include/PressureEquation.h:25-26, src/PressureEquation.cpp:288-343): adds the mean of
 * (exact - pressure) over all owner points of all ranks to every point of the tensor, ghosts included.
 * Collective over the ranks of a distributed context. */
int mifgpu_adjust_pressure(mifgpu_ctx *ctx, mifgpu_tensor *pressure, const mifgpu_bc *exact, double time);

/* ---- host-side collectives --------------------------------------------------------------------------- */

/* What the reference's host code does with MPI around the path, for the ranks of a distributed context (single-rank
 * contexts: identity / copy).  mifgpu_allreduce: values[] <- sum (op = 0) or max (op = 1) over all ranks, in place, on
 * every rank.  mifgpu_gather: MPI_Gather of the counts + MPI_Gatherv of the data to rank 0 -- recv (rank 0 only; may
 * be NULL elsewhere) receives the ranks' arrays one after the other in rank order and counts[r] (every rank, nranks
 * entries) the number of values of rank r.  Used by accumulate_error_mpi_* (src/Norms.cpp:120-162) and by the
 * writers' MPI-IO offsets / gathers (src/VTKDatExport.cpp:54-69,219-311,519-554) in the host layer. */
int mifgpu_allreduce(mifgpu_ctx *ctx, double *values, int32_t count, int32_t op);
int mifgpu_gather(mifgpu_ctx *ctx, const double *send, uint64_t count, double *recv, uint64_t *counts);
int mifgpu_rank_count(const mifgpu_ctx *ctx);

/* How this context moves data between the y and the z sweeps of the Poisson solve (the 2Decomp transposes,
 * deps/2Decomp_C/Transpose*.cpp), as decided at creation -- e.g. whether the CUDA IPC mapping of the peers' buffers
 * succeeded.  -1 for a NULL context. */
enum {
  MIFGPU_TRANSPOSE_NONE = 0,           /* one rank: strided sweeps in place */
  MIFGPU_TRANSPOSE_PEER_FUSED = 1,     /* z slabs: the sweeps store into the peers' buffers over NVLink */
  MIFGPU_TRANSPOSE_NCCL_ALLTOALL = 2,  /* z slabs: pack, grouped ncclSend/ncclRecv, unpack */
  MIFGPU_TRANSPOSE_PENCIL_BOXES = 3    /* Py x Pz pencils: four box exchanges per solve */
};
int mifgpu_transpose_path(const mifgpu_ctx *ctx);

/* Blocks until all work queued by this context has finished (its compute stream and both copy streams). */
int mifgpu_synchronize(mifgpu_ctx *ctx);

/* The CUDA stream (cudaStream_t) the context launches on, for callers that time with CUDA events. */
void *mifgpu_stream(mifgpu_ctx *ctx);

/* Optional per-kernel timing (no counterpart in the reference, whose only timers are MPI_Wtime around the
 * time loop, test/full_test.cpp:116-138).  While enabled, every kernel group of the step is bracketed by
 * CUDA events on the context's stream; mifgpu_profile_read synchronises, returns the number of
 * categories (<= capacity) and fills, per category, a static name, the accumulated milliseconds and the
 * number of timed groups since the previous read. */
int mifgpu_profile_enable(mifgpu_ctx *ctx, int enable);
int mifgpu_profile_read(mifgpu_ctx *ctx, int capacity, const char **names, double *milliseconds, uint64_t *counts);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t mifgpu_launch_count(const mifgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif

#endif /* MIFGPU_H */
