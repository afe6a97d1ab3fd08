// mif_stencil.cu -- stencil-type kernels of the projection step (sm_100a, FP64):
//   RK stage kernels      src/Timestep.cpp:10-54 + include/MomentumEquation.h:42-255 + include/PressureGradient.h:9-24
//   Dirichlet faces       src/VelocityTensor.cpp:36-218
//   periodic ghost copies src/StaggeredTensor.cpp:221-257
//   divergence rhs        src/PressureEquation.cpp:59-61 + include/VelocityDivergence.h:9-20
//   p += dp, u -= dt grad(dp)   src/Timestep.cpp:66-81
// All fields use the uniform padded layout of mif_common.cuh, so one linear index addresses every array.
#include <math_constants.h>

#include <cstdlib>

#include "../../include/mifgpu.h"
#include "mif_kernels.h"

namespace mifgpu {

namespace {

__device__ __forceinline__ void prefetch_l2(const real *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }


// ------------------------------------------------------------------------------------------------
// RK stage kernel, two x-adjacent points per thread.  The pair (i, i+1), i even, is 16-byte aligned in every array
// (rows start on 128-byte boundaries), so the 31 values per point of stage_kernel become 20 128-bit loads plus 11
// scalar loads (the x-1 / x+2 ends of the rows) per PAIR: half the load instructions and about 70% of the L1
// wavefronts per point of the one-point kernel, which is bound by exactly those.  Arithmetic per point is identical.
// ------------------------------------------------------------------------------------------------
struct StageCoef {
  real a1, a2, a3, b;
};

template <int STAGE>
__device__ __forceinline__ void rk_combine(real c_val, real rhs, real p_grad, real a_old, const StageCoef &k,
                                           real &a_new, real &b_new) {
  if (STAGE == 1) {
    a_new = c_val + k.a1 * rhs - k.b * p_grad;  // src/Timestep.cpp:18
    b_new = rhs;                                // src/Timestep.cpp:19
  } else if (STAGE == 2) {
    const real rhs_1 = a_old;
    const real rhs_2_scaled = k.a2 * rhs;
    a_new = c_val + k.a1 * rhs_1 + rhs_2_scaled - k.b * p_grad;  // src/Timestep.cpp:35-36
    b_new = rhs_2_scaled;                                        // src/Timestep.cpp:37
  } else {
    const real rhs_2_scaled = -a_old;
    a_new = c_val + rhs_2_scaled + k.a3 * rhs - k.b * p_grad;  // src/Timestep.cpp:51-52
    b_new = RC(0.0);
  }
}

// Loads / stores of one component's pair: 128-bit when both points are written, scalar otherwise (the other point is
// a boundary value that must stay untouched).
__device__ __forceinline__ real2 load_pair_if(const real *ptr, bool any) {
  return any ? *reinterpret_cast<const real2 *>(ptr) : make_real2(RC(0.0), RC(0.0));
}
__device__ __forceinline__ void store_pair(real *ptr, bool w0, bool w1, real v0, real v1) {
  if (w0 && w1) *reinterpret_cast<real2 *>(ptr) = make_real2(v0, v1);
  else if (w0) ptr[0] = v0;
  else if (w1) ptr[1] = v1;
}

#ifndef MIFGPU_STAGE_CTAS
#define MIFGPU_STAGE_CTAS 4  // resident CTAs per SM the register budget is set for: 128 registers.  Measured with 5 (96
                             // registers) and 6 (80): 7 % and 22 % slower (profiles/r02_s7_ab_stage_occupancy.jsonl)
#endif
template <int STAGE>
__global__ void __launch_bounds__(128, MIFGPU_STAGE_CTAS)
stage_kernel_pair(const Geom g, const real *__restrict__ in_u, const real *__restrict__ in_v,
                  const real *__restrict__ in_w, const real *__restrict__ p, real *__restrict__ a_u,
                  real *__restrict__ a_v, real *__restrict__ a_w, real *__restrict__ b_u, real *__restrict__ b_v,
                  real *__restrict__ b_w, int prefetch_planes, int nk, int chunk_blocks_y, int k_shift) {
  const int y_chunk = blockIdx.z / nk;  // blockIdx.z = chunk * nk + plane: y chunks keep a plane's working set in L2
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j = (y_chunk * chunk_blocks_y + blockIdx.y) * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z - y_chunk * nk + 1 + k_shift;
  if (prefetch_planes > 0 && (threadIdx.x & 1) == 0 && k + prefetch_planes < g.PZ && i < g.PX && j < g.PY) {
    const long long ahead = gidx(g, i, j, k + prefetch_planes);
    prefetch_l2(in_u + ahead);
    prefetch_l2(in_v + ahead);
    prefetch_l2(in_w + ahead);
    prefetch_l2(p + ahead);
    if (STAGE >= 2) {
      prefetch_l2(a_u + ahead);
      prefetch_l2(a_v + ahead);
      prefetch_l2(a_w + ahead);
    }
  }
  if (i > max(g.sx[0], g.Nx) - 2) return;
  const bool in_j_v = j <= g.sy[1] - 2, in_j_o = j <= g.Ny - 2;
  if (!in_j_v && !in_j_o) return;
  const bool in_k_w = k <= g.sz[2] - 2, in_k_o = k <= g.Nz - 2;
  bool do_u[2], do_v[2], do_w[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int ii = i + e;
    const bool in_i_u = ii >= 1 && ii <= g.sx[0] - 2, in_i_o = ii >= 1 && ii <= g.Nx - 2;
    do_u[e] = in_i_u && in_j_o && in_k_o;
    do_v[e] = in_i_o && in_j_v && in_k_o;
    do_w[e] = in_i_o && in_j_o && in_k_w;
  }
  if (!(do_u[0] || do_u[1] || do_v[0] || do_v[1] || do_w[0] || do_w[1])) return;

  const long long c = gidx(g, i, j, k);
  const long long sj = g.PX, sk = g.plane;
  const bool has_left = i >= 2, has_right = i + 2 < g.PX;
  auto pair = [](const real *ptr) { return *reinterpret_cast<const real2 *>(ptr); };

  // rows of u, v, w, p (pairs) and the row ends one to the left / two to the right
  const real2 U_c = pair(in_u + c), U_ym = pair(in_u + c - sj), U_yp = pair(in_u + c + sj);
  const real2 U_zm = pair(in_u + c - sk), U_zp = pair(in_u + c + sk);
  const real u_l = has_left ? in_u[c - 1] : RC(0.0), u_r = has_right ? in_u[c + 2] : RC(0.0);
  const real u_r_ym = has_right ? in_u[c + 2 - sj] : RC(0.0), u_r_zm = has_right ? in_u[c + 2 - sk] : RC(0.0);
  const real2 V_c = pair(in_v + c), V_ym = pair(in_v + c - sj), V_yp = pair(in_v + c + sj);
  const real2 V_zm = pair(in_v + c - sk), V_zp = pair(in_v + c + sk), V_yp_zm = pair(in_v + c + sj - sk);
  const real v_l = has_left ? in_v[c - 1] : RC(0.0), v_r = has_right ? in_v[c + 2] : RC(0.0);
  const real v_l_yp = has_left ? in_v[c - 1 + sj] : RC(0.0);
  const real2 W_c = pair(in_w + c), W_ym = pair(in_w + c - sj), W_yp = pair(in_w + c + sj);
  const real2 W_zm = pair(in_w + c - sk), W_zp = pair(in_w + c + sk), W_ym_zp = pair(in_w + c - sj + sk);
  const real w_l = has_left ? in_w[c - 1] : RC(0.0), w_r = has_right ? in_w[c + 2] : RC(0.0);
  const real w_l_zp = has_left ? in_w[c - 1 + sk] : RC(0.0);
  const real2 P_c = pair(p + c), P_ym = pair(p + c - sj), P_zm = pair(p + c - sk);
  const real p_l = has_left ? p[c - 1] : RC(0.0);

  const real dt = g.dt;
  StageCoef coef;
  if (STAGE == 1) {
    coef.a1 = RC(64.0) / RC(120.0) * dt; coef.a2 = RC(0.0); coef.a3 = RC(0.0); coef.b = coef.a1;
  } else if (STAGE == 2) {
    coef.a1 = -RC(34.0) / RC(120.0) * dt; coef.a2 = RC(50.0) / RC(120.0) * dt; coef.a3 = RC(0.0); coef.b = coef.a1 + coef.a2;
  } else {
    coef.a1 = RC(0.0); coef.a2 = -RC(50.0) / RC(120.0) * dt; coef.a3 = RC(90.0) / RC(120.0) * dt; coef.b = coef.a2 + coef.a3;
  }

  const real2 A_u = (STAGE >= 2) ? load_pair_if(a_u + c, do_u[0] || do_u[1]) : make_real2(RC(0.0), RC(0.0));
  const real2 A_v = (STAGE >= 2) ? load_pair_if(a_v + c, do_v[0] || do_v[1]) : make_real2(RC(0.0), RC(0.0));
  const real2 A_w = (STAGE >= 2) ? load_pair_if(a_w + c, do_w[0] || do_w[1]) : make_real2(RC(0.0), RC(0.0));
  real na_u[2], nb_u[2], na_v[2], nb_v[2], na_w[2], nb_w[2];

#pragma unroll
  for (int e = 0; e < 2; e++) {
    // neighbourhood of point i + e (names as in stage_kernel)
    const real u_c = e ? U_c.y : U_c.x, u_xm = e ? U_c.x : u_l, u_xp = e ? u_r : U_c.y;
    const real u_ym = e ? U_ym.y : U_ym.x, u_yp = e ? U_yp.y : U_yp.x, u_zm = e ? U_zm.y : U_zm.x, u_zp = e ? U_zp.y : U_zp.x;
    const real u_xp_ym = e ? u_r_ym : U_ym.y, u_xp_zm = e ? u_r_zm : U_zm.y;
    const real v_c = e ? V_c.y : V_c.x, v_xm = e ? V_c.x : v_l, v_xp = e ? v_r : V_c.y;
    const real v_ym = e ? V_ym.y : V_ym.x, v_yp = e ? V_yp.y : V_yp.x, v_zm = e ? V_zm.y : V_zm.x, v_zp = e ? V_zp.y : V_zp.x;
    const real v_xm_yp = e ? V_yp.x : v_l_yp, v_yp_zm = e ? V_yp_zm.y : V_yp_zm.x;
    const real w_c = e ? W_c.y : W_c.x, w_xm = e ? W_c.x : w_l, w_xp = e ? w_r : W_c.y;
    const real w_ym = e ? W_ym.y : W_ym.x, w_yp = e ? W_yp.y : W_yp.x, w_zm = e ? W_zm.y : W_zm.x, w_zp = e ? W_zp.y : W_zp.x;
    const real w_xm_zp = e ? W_zp.x : w_l_zp, w_ym_zp = e ? W_ym_zp.y : W_ym_zp.x;
    const real p_c = e ? P_c.y : P_c.x, p_xm = e ? P_c.x : p_l, p_ym = e ? P_ym.y : P_ym.x, p_zm = e ? P_zm.y : P_zm.x;
    {
      // include/MomentumEquation.h:50-96
      const real convection = -u_c * (u_xp - u_xm) * g.one_over_2_dx -
                                (v_yp + v_c + v_xm_yp + v_xm) * (u_yp - u_ym) * g.one_over_8_dy -
                                (w_zp + w_c + w_xm_zp + w_xm) * (u_zp - u_zm) * g.one_over_8_dz;
      const real diffusion = (u_xp - 2 * u_c + u_xm) * g.one_over_dx2_Re + (u_yp - 2 * u_c + u_ym) * g.one_over_dy2_Re +
                               (u_zp - 2 * u_c + u_zm) * g.one_over_dz2_Re;
      const real p_grad = (p_c - p_xm) * g.one_over_dx;  // include/PressureGradient.h:9-12
      rk_combine<STAGE>(u_c, convection + diffusion, p_grad, e ? A_u.y : A_u.x, coef, na_u[e], nb_u[e]);
    }
    {
      // include/MomentumEquation.h:126-165
      const real convection = -(u_xp + u_c + u_xp_ym + u_ym) * (v_xp - v_xm) * g.one_over_8_dx -
                                v_c * (v_yp - v_ym) * g.one_over_2_dy -
                                (w_zp + w_c + w_ym_zp + w_ym) * (v_zp - v_zm) * g.one_over_8_dz;
      const real diffusion = (v_xp - 2 * v_c + v_xm) * g.one_over_dx2_Re + (v_yp - 2 * v_c + v_ym) * g.one_over_dy2_Re +
                               (v_zp - 2 * v_c + v_zm) * g.one_over_dz2_Re;
      const real p_grad = (p_c - p_ym) * g.one_over_dy;  // include/PressureGradient.h:15-18
      rk_combine<STAGE>(v_c, convection + diffusion, p_grad, e ? A_v.y : A_v.x, coef, na_v[e], nb_v[e]);
    }
    {
      // include/MomentumEquation.h:196-236
      const real convection = -(u_xp + u_c + u_xp_zm + u_zm) * (w_xp - w_xm) * g.one_over_8_dx -
                                (v_yp + v_c + v_yp_zm + v_zm) * (w_yp - w_ym) * g.one_over_8_dy -
                                w_c * (w_zp - w_zm) * g.one_over_2_dz;
      const real diffusion = (w_xp - 2 * w_c + w_xm) * g.one_over_dx2_Re + (w_yp - 2 * w_c + w_ym) * g.one_over_dy2_Re +
                               (w_zp - 2 * w_c + w_zm) * g.one_over_dz2_Re;
      const real p_grad = (p_c - p_zm) * g.one_over_dz;  // include/PressureGradient.h:21-24
      rk_combine<STAGE>(w_c, convection + diffusion, p_grad, e ? A_w.y : A_w.x, coef, na_w[e], nb_w[e]);
    }
  }
  store_pair(a_u + c, do_u[0], do_u[1], na_u[0], na_u[1]);
  store_pair(a_v + c, do_v[0], do_v[1], na_v[0], na_v[1]);
  store_pair(a_w + c, do_w[0], do_w[1], na_w[0], na_w[1]);
  if (STAGE != 3) {
    store_pair(b_u + c, do_u[0], do_u[1], nb_u[0], nb_u[1]);
    store_pair(b_v + c, do_v[0], do_v[1], nb_v[0], nb_v[1]);
    store_pair(b_w + c, do_w[0], do_w[1], nb_w[0], nb_w[1]);
  }
}

// ------------------------------------------------------------------------------------------------
// Analytic boundary data on the device.
// ------------------------------------------------------------------------------------------------
// The reference's generated functions (generators/manufsol*.py -> C with double literals and `double Reynolds`) evaluate
// in double whatever `Real` is and round the result once; so do these: arguments and result in the build's scalar
// type, arithmetic in double.
__device__ __forceinline__ real exact_velocity(int kind, int comp, real t_r, real x_r, real y_r, real z_r,
                                               real Re_r) {
  const double t = t_r, x = x_r, y = y_r, z = z_r, Re = Re_r;
  if (kind == MIFGPU_BC_ETHIER_STEINMAN) {
    // generators/manufsol.py:31-57 with a = pi/4, d = pi/2.
    const double a = CUDART_PI / 4.0, d = CUDART_PI / 2.0;
    const double decay = exp(-d * d * t / Re);
    if (comp == 0) return -a * (exp(a * x) * sin(a * y + d * z) + exp(a * z) * cos(a * x + d * y)) * decay;
    if (comp == 1) return -a * (exp(a * y) * sin(a * z + d * x) + exp(a * x) * cos(a * y + d * z)) * decay;
    return -a * (exp(a * z) * sin(a * x + d * y) + exp(a * y) * cos(a * z + d * x)) * decay;
  }
  if (kind == MIFGPU_BC_VELOCITY_TEST) {
    // generators/manufsol_velocity.py:55-59
    if (comp == 0) return sin(x) * cos(y) * sin(z) * sin(t);
    if (comp == 1) return cos(x) * sin(y) * sin(z) * sin(t);
    return 2 * cos(x) * cos(y) * cos(z) * sin(t);
  }
  // include/TestCaseBoundaries.h:17-56: only v is non-zero, and only on one x face.
  if (comp != 1) return 0.0;
  // compared in double in the float build too, as the reference's `x < 1.0 + exact_solution_precision` promotes
  const double face = (kind == MIFGPU_BC_TEST_CASE_1) ? 1.0 : -0.5;
  const double precision = (double)RC(1e-12);  // `constexpr Real exact_solution_precision = 1e-12`
  return (x < face + precision && x > face - precision) ? 1.0 : 0.0;
}

// Exact pressure of the analytic family (generators/manufsol.py:58-72, Ethier-Steinman); the lid-driven test cases have
// p = 0 as their reference pressure (include/TestCaseBoundaries.h: exact_p_initial_t*).
__device__ __forceinline__ real exact_pressure(int kind, real t_r, real x_r, real y_r, real z_r, real Re_r) {
  const double t = t_r, x = x_r, y = y_r, z = z_r, Re = Re_r;
  if (kind != MIFGPU_BC_ETHIER_STEINMAN) return 0.0;
  const double a = CUDART_PI / 4.0, d = CUDART_PI / 2.0;
  return -a * a / 2.0 *
         (exp(2 * a * x) + exp(2 * a * y) + exp(2 * a * z) + 2 * sin(a * x + d * y) * cos(a * z + d * x) * exp(a * (y + z)) +
          2 * sin(a * y + d * z) * cos(a * x + d * y) * exp(a * (z + x)) +
          2 * sin(a * z + d * x) * cos(a * y + d * z) * exp(a * (x + y))) *
         exp(-2 * d * d * t / Re);
}

// f_c evaluated at the staggered coordinate of component `at` index (i, j, k)
// (evaluate_function_at_index, include/VelocityTensor.h:16-22,40-46,64-70) or, with at = 3, at the
// unstaggered pressure point (include/StaggeredTensor.h:113-119).
__device__ __forceinline__ real eval_at(const Geom &g, const BcDev &bc, int f_comp, int at, int i, int j, int k) {
  real x = g.min_x + g.dx * (g.base_i + i);
  real y = g.min_y + g.dy * (g.base_j + j);
  real z = g.min_z + g.dz * (g.base_k + k);
  if (at == 0) x -= g.dx_over_2;
  if (at == 1) y -= g.dy_over_2;
  if (at == 2) z -= g.dz_over_2;
  return exact_velocity(bc.kind, f_comp, bc.time, x, y, z, bc.Re);
}

__device__ __forceinline__ bool face_active(const Geom &g, int face) {
  switch (face) {
    case 0: return g.prev_z == -1 && !g.periodic[2];
    case 1: return g.next_z == -1 && !g.periodic[2];
    case 2: return g.prev_y == -1 && !g.periodic[1];
    case 3: return g.next_y == -1 && !g.periodic[1];
    default: return !g.periodic[0];
  }
}

// ------------------------------------------------------------------------------------------------
// Velocity-only integrator with analytic forcing: mif::timestep_velocity (src/TimestepVelocity.cpp:20-90,
// include/MomentumEquationForcing.h:11-33).
// ------------------------------------------------------------------------------------------------
// forcing_{x,y,z} of generators/manufsol_velocity.py:38-48, f = d_t c + (u . grad) c - lap(c) / Re for the manufactured
// field of MIFGPU_BC_VELOCITY_TEST (same closed form as oracle/mif_oracle.c forcing(), checked against sympy there).
__device__ __forceinline__ real manufactured_forcing(int comp, real t_r, real x_r, real y_r, real z_r, real Re_r) {
  const double t = t_r, x = x_r, y = y_r, z = z_r, Re = Re_r;  // evaluated in double like the reference's generated code
  double sx, cx, sy, cy, sz, cz, st, ct;
  sincos(x, &sx, &cx);
  sincos(y, &sy, &cy);
  sincos(z, &sz, &cz);
  sincos(t, &st, &ct);
  if (comp == 0) return sx * cy * sz * ct + st * st * sx * cx * (2.0 - 2.0 * sy * sy - sz * sz) + 3.0 * sx * cy * sz * st / Re;
  if (comp == 1) return cx * sy * sz * ct + st * st * sy * cy * (2.0 - 2.0 * sx * sx - sz * sz) + 3.0 * cx * sy * sz * st / Re;
  return 2.0 * cx * cy * cz * ct + 2.0 * st * st * sz * cz * (sx * sx + sy * sy - 2.0) + 6.0 * cx * cy * cz * st / Re;
}

// One thread per grid index, all three components (loads shared as in stage_kernel).
//   STAGE 1 (Y2):  S = velocity;         rhs_buf = R;  out = S + dt a21 R                      (out = velocity_buffer)
//   STAGE 2 (Y3):  S = velocity_buffer;  r = rhs_buf;  rhs_buf = out + dt b1 r;  out = out + dt (a31 r + a32 R)   (out = velocity)
//   STAGE 3 (U*):  S = velocity;         r = rhs_buf;  out = r + dt b3 R                        (out = velocity_buffer)
// with R = momentum rhs of S + forcing at the stage time, evaluated at the staggered coordinate of the component.
template <int STAGE>
__global__ void __launch_bounds__(256)
velocity_stage_kernel(const Geom g, const real *__restrict__ in_u, const real *__restrict__ in_v,
                      const real *__restrict__ in_w, real *__restrict__ r_u, real *__restrict__ r_v,
                      real *__restrict__ r_w, real *__restrict__ o_u, real *__restrict__ o_v, real *__restrict__ o_w,
                      real time, real Re) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z + 1;
  const bool in_i_u = i <= g.sx[0] - 2, in_i_o = i <= g.Nx - 2;
  const bool in_j_v = j <= g.sy[1] - 2, in_j_o = j <= g.Ny - 2;
  const bool in_k_w = k <= g.sz[2] - 2, in_k_o = k <= g.Nz - 2;
  const bool do_u = in_i_u && in_j_o && in_k_o;
  const bool do_v = in_i_o && in_j_v && in_k_o;
  const bool do_w = in_i_o && in_j_o && in_k_w;
  if (!do_u && !do_v && !do_w) return;
  const long long c = gidx(g, i, j, k);
  const long long sj = g.PX, sk = g.plane;
  const real u_c = in_u[c], u_xm = in_u[c - 1], u_xp = in_u[c + 1];
  const real u_ym = in_u[c - sj], u_yp = in_u[c + sj], u_zm = in_u[c - sk], u_zp = in_u[c + sk];
  const real u_xp_ym = in_u[c + 1 - sj], u_xp_zm = in_u[c + 1 - sk];
  const real v_c = in_v[c], v_xm = in_v[c - 1], v_xp = in_v[c + 1];
  const real v_ym = in_v[c - sj], v_yp = in_v[c + sj], v_zm = in_v[c - sk], v_zp = in_v[c + sk];
  const real v_xm_yp = in_v[c - 1 + sj], v_yp_zm = in_v[c + sj - sk];
  const real w_c = in_w[c], w_xm = in_w[c - 1], w_xp = in_w[c + 1];
  const real w_ym = in_w[c - sj], w_yp = in_w[c + sj], w_zm = in_w[c - sk], w_zp = in_w[c + sk];
  const real w_xm_zp = in_w[c - 1 + sk], w_ym_zp = in_w[c - sj + sk];
  const real x = g.min_x + g.dx * (g.base_i + i), y = g.min_y + g.dy * (g.base_j + j), z = g.min_z + g.dz * (g.base_k + k);
  const real dt = g.dt;
  constexpr real a21 = RC(8.0) / RC(15.0), a31 = RC(1.0) / RC(4.0), a32 = RC(5.0) / RC(12.0), b1 = RC(1.0) / RC(4.0), b3 = RC(3.0) / RC(4.0);  // src/TimestepVelocity.cpp:11-17

  auto combine = [&](real rhs, real center, real *r, real *o) {
    if (STAGE == 1) {
      r[c] = rhs;                        // :25
      o[c] = center + dt * a21 * rhs;    // :26
    } else if (STAGE == 2) {
      const real prev = r[c], vel = o[c];
      r[c] = vel + dt * (b1 * prev);                  // :35
      o[c] = vel + dt * (a31 * prev + a32 * rhs);     // :36-38
    } else {
      o[c] = r[c] + dt * (b3 * rhs);     // :46-48
    }
  };
  if (do_u) {
    const real convection = -u_c * (u_xp - u_xm) * g.one_over_2_dx -
                              (v_yp + v_c + v_xm_yp + v_xm) * (u_yp - u_ym) * g.one_over_8_dy -
                              (w_zp + w_c + w_xm_zp + w_xm) * (u_zp - u_zm) * g.one_over_8_dz;
    const real diffusion = (u_xp - 2 * u_c + u_xm) * g.one_over_dx2_Re + (u_yp - 2 * u_c + u_ym) * g.one_over_dy2_Re +
                             (u_zp - 2 * u_c + u_zm) * g.one_over_dz2_Re;
    combine(convection + diffusion + manufactured_forcing(0, time, x - g.dx_over_2, y, z, Re), u_c, r_u, o_u);
  }
  if (do_v) {
    const real convection = -(u_xp + u_c + u_xp_ym + u_ym) * (v_xp - v_xm) * g.one_over_8_dx -
                              v_c * (v_yp - v_ym) * g.one_over_2_dy -
                              (w_zp + w_c + w_ym_zp + w_ym) * (v_zp - v_zm) * g.one_over_8_dz;
    const real diffusion = (v_xp - 2 * v_c + v_xm) * g.one_over_dx2_Re + (v_yp - 2 * v_c + v_ym) * g.one_over_dy2_Re +
                             (v_zp - 2 * v_c + v_zm) * g.one_over_dz2_Re;
    combine(convection + diffusion + manufactured_forcing(1, time, x, y - g.dy_over_2, z, Re), v_c, r_v, o_v);
  }
  if (do_w) {
    const real convection = -(u_xp + u_c + u_xp_zm + u_zm) * (w_xp - w_xm) * g.one_over_8_dx -
                              (v_yp + v_c + v_yp_zm + v_zm) * (w_yp - w_ym) * g.one_over_8_dy -
                              w_c * (w_zp - w_zm) * g.one_over_2_dz;
    const real diffusion = (w_xp - 2 * w_c + w_xm) * g.one_over_dx2_Re + (w_yp - 2 * w_c + w_ym) * g.one_over_dy2_Re +
                             (w_zp - 2 * w_c + w_zm) * g.one_over_dz2_Re;
    combine(convection + diffusion + manufactured_forcing(2, time, x, y, z - g.dz_over_2, Re), w_c, r_w, o_w);
  }
}

// One thread per (component, face, point of the face).  The reference writes the faces in the order
// z-, z+, y-, y+, x-, x+ over the full 2-D extent of each tensor, so on edges and corners the last face
// wins (src/VelocityTensor.cpp:47-217); here a thread simply does not write where a later active face
// owns the point, which gives the same final state without ordering launches.
__global__ void __launch_bounds__(256) bc_face_kernel(const Geom g, real *u, real *v, real *w, const BcDev bc) {
  const int comp = blockIdx.y / 6, face = blockIdx.y % 6;
  if (!face_active(g, face)) return;
  const int sx = g.sx[comp], sy = g.sy[comp], sz = g.sz[comp];
  const int dir = 2 - face / 2;  // normal direction of the face: z, z, y, y, x, x
  const int na = (dir == 0) ? sy : sx;
  const int nb = (dir == 2) ? sy : sz;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)na * nb) return;
  const int a = (int)(t % na), b2 = (int)(t / na);
  const bool upper = face & 1;
  int i, j, k;
  if (dir == 2) {
    i = a; j = b2; k = upper ? sz - 1 : 0;
  } else if (dir == 1) {
    i = a; k = b2; j = upper ? sy - 1 : 0;
  } else {
    j = a; k = b2; i = upper ? sx - 1 : 0;
  }
  // Later faces that also contain this point take precedence.
  for (int later = face + 1; later < 6; later++) {
    if (!face_active(g, later)) continue;
    bool on;
    switch (later) {
      case 1: on = (k == sz - 1); break;
      case 2: on = (j == 0); break;
      case 3: on = (j == sy - 1); break;
      case 4: on = (i == 0); break;
      default: on = (i == sx - 1); break;
    }
    if (on) return;
  }
  real value;
  if (bc.kind == MIFGPU_BC_HOST_CALLBACK) {
    value = bc.tables[comp][face][t];
  } else if (comp == dir) {
    // Wall-normal component: value on the wall plus half a cell of -(tangential divergence) of the
    // analytic field (src/VelocityTensor.cpp:51-60,78-90,109-118,136-148,167-176,194-206).
    // `q` is the unstaggered index of the wall along the normal direction.
    if (dir == 2) {
      const int q = upper ? g.Nz - 1 : 0;
      const real at_wall = eval_at(g, bc, 2, 3, i, j, q);
      const real du_dx = (eval_at(g, bc, 0, 0, i + 1, j, q) - eval_at(g, bc, 0, 0, i, j, q)) * g.one_over_dx;
      const real dv_dy = (eval_at(g, bc, 1, 1, i, j + 1, q) - eval_at(g, bc, 1, 1, i, j, q)) * g.one_over_dy;
      value = upper ? at_wall - g.dz_over_2 * (du_dx + dv_dy) : at_wall + g.dz_over_2 * (du_dx + dv_dy);
    } else if (dir == 1) {
      const int q = upper ? g.Ny - 1 : 0;
      const real at_wall = eval_at(g, bc, 1, 3, i, q, k);
      const real du_dx = (eval_at(g, bc, 0, 0, i + 1, q, k) - eval_at(g, bc, 0, 0, i, q, k)) * g.one_over_dx;
      const real dw_dz = (eval_at(g, bc, 2, 2, i, q, k + 1) - eval_at(g, bc, 2, 2, i, q, k)) * g.one_over_dz;
      value = upper ? at_wall - g.dy_over_2 * (du_dx + dw_dz) : at_wall + g.dy_over_2 * (du_dx + dw_dz);
    } else {
      const int q = upper ? g.Nx - 1 : 0;
      const real at_wall = eval_at(g, bc, 0, 3, q, j, k);
      const real dv_dy = (eval_at(g, bc, 1, 1, q, j + 1, k) - eval_at(g, bc, 1, 1, q, j, k)) * g.one_over_dy;
      const real dw_dz = (eval_at(g, bc, 2, 2, q, j, k + 1) - eval_at(g, bc, 2, 2, q, j, k)) * g.one_over_dz;
      value = upper ? at_wall - g.dx_over_2 * (dv_dy + dw_dz) : at_wall + g.dx_over_2 * (dv_dy + dw_dz);
    }
  } else {
    // Tangential component: the analytic value at the staggered coordinate of the face point
    // (src/VelocityTensor.cpp:66-67,96-98,124-125,154-156,182-183,212-214).
    value = eval_at(g, bc, comp, comp, i, j, k);
  }
  real *field = comp == 0 ? u : (comp == 1 ? v : w);
  field[gidx(g, i, j, k)] = value;
}

// ghost(0) <- slice(s-2), ghost(s-1) <- slice(1) along `dir` (src/StaggeredTensor.cpp:221-257).
__global__ void __launch_bounds__(256) periodic_copy_kernel(const Geom g, real *field, int sx, int sy, int sz, int dir) {
  const int na = (dir == 0) ? sy : sx;
  const int nb = (dir == 2) ? sy : sz;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)na * nb) return;
  const int a = (int)(t % na), b = (int)(t / na);
  if (dir == 0) {
    field[gidx(g, 0, a, b)] = field[gidx(g, sx - 2, a, b)];
    field[gidx(g, sx - 1, a, b)] = field[gidx(g, 1, a, b)];
  } else if (dir == 1) {
    field[gidx(g, a, 0, b)] = field[gidx(g, a, sy - 2, b)];
    field[gidx(g, a, sy - 1, b)] = field[gidx(g, a, 1, b)];
  } else {
    field[gidx(g, a, b, 0)] = field[gidx(g, a, b, sz - 2)];
    field[gidx(g, a, b, sz - 1)] = field[gidx(g, a, b, 1)];
  }
}

// Two x-adjacent points per thread, 128-bit accesses (see correct_kernel below).
__global__ void __launch_bounds__(256)
divergence_kernel(const Geom g, const real *__restrict__ u, const real *__restrict__ v,
                  const real *__restrict__ w, real dt, real *__restrict__ rhs, int k_shift) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j = blockIdx.y * blockDim.y + threadIdx.y + g.own_lo[1];
  const int k = blockIdx.z + g.own_lo[2] + k_shift;
  if (i >= g.own_hi[0] || j >= g.own_hi[1]) return;
  const bool do0 = i >= g.own_lo[0], do1 = i + 1 < g.own_hi[0];
  const long long c = gidx(g, i, j, k);
  auto pair = [](const real *ptr) { return *reinterpret_cast<const real2 *>(ptr); };
  const real2 U = pair(u + c), V0 = pair(v + c), V1 = pair(v + c + g.PX), W0 = pair(w + c), W1 = pair(w + c + g.plane);
  real r0 = RC(0.0), r1 = RC(0.0);
  if (do0) {
    const real du_dx = (U.y - U.x) * g.one_over_dx;
    const real dv_dy = (V1.x - V0.x) * g.one_over_dy;
    const real dw_dz = (W1.x - W0.x) * g.one_over_dz;
    r0 = (du_dx + dv_dy + dw_dz) / dt;
  }
  if (do1) {
    const real du_dx = (u[c + 2] - U.y) * g.one_over_dx;
    const real dv_dy = (V1.y - V0.y) * g.one_over_dy;
    const real dw_dz = (W1.y - W0.y) * g.one_over_dz;
    r1 = (du_dx + dv_dy + dw_dz) / dt;
  }
  if (do0 && do1) *reinterpret_cast<real2 *>(rhs + c) = make_real2(r0, r1);
  else if (do0) rhs[c] = r0;
  else if (do1) rhs[c + 1] = r1;
}

// Two x-adjacent points per thread with 128-bit accesses (rows start on 128-byte boundaries and PX is even, so the
// pair (i, i+1), i even, is 16-byte aligned in every array); arithmetic per point as in the reference.
__global__ void __launch_bounds__(256)
correct_kernel(const Geom g, real *__restrict__ u, real *__restrict__ v, real *__restrict__ w,
               real *__restrict__ p, const real *__restrict__ dp, real dt_s, int k_shift) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z + k_shift;
  if (i >= g.sx[0] || j >= g.sy[1]) return;
  const long long c = gidx(g, i, j, k);
  const real2 raw = *reinterpret_cast<const real2 *>(dp + c);
  const bool in_jk = j < g.Ny && k < g.Nz;
  const bool in_p0 = in_jk && i < g.Nx, in_p1 = in_jk && i + 1 < g.Nx;
  const real d0 = in_p0 ? raw.x : RC(0.0), d1 = in_p1 ? raw.y : RC(0.0);
  if (in_p0) {  // src/Timestep.cpp:79-81 (all points, ghosts included)
    real2 pv = *reinterpret_cast<real2 *>(p + c);
    pv.x += d0;
    if (in_p1) pv.y += d1;
    *reinterpret_cast<real2 *>(p + c) = pv;
  }
  const bool int_j_o = j >= 1 && j <= g.Ny - 2, int_k_o = k >= 1 && k <= g.Nz - 2;
  const bool int_i_o0 = i >= 1 && i <= g.Nx - 2, int_i_o1 = i + 1 <= g.Nx - 2;
  // src/Timestep.cpp:66-72 (interior points of each component)
  if (int_j_o && int_k_o) {
    const bool do0 = i >= 1 && i <= g.sx[0] - 2, do1 = i + 1 <= g.sx[0] - 2;
    if (do0 || do1) {
      real2 uv = *reinterpret_cast<real2 *>(u + c);
      if (do0) uv.x -= (d0 - dp[c - 1]) * g.one_over_dx * dt_s;
      if (do1) uv.y -= (d1 - raw.x) * g.one_over_dx * dt_s;
      *reinterpret_cast<real2 *>(u + c) = uv;
    }
  }
  if ((int_i_o0 || int_i_o1) && j >= 1 && j <= g.sy[1] - 2 && int_k_o) {
    const real2 low = *reinterpret_cast<const real2 *>(dp + c - g.PX);
    real2 vv = *reinterpret_cast<real2 *>(v + c);
    if (int_i_o0) vv.x -= (d0 - low.x) * g.one_over_dy * dt_s;
    if (int_i_o1) vv.y -= (d1 - low.y) * g.one_over_dy * dt_s;
    *reinterpret_cast<real2 *>(v + c) = vv;
  }
  if ((int_i_o0 || int_i_o1) && int_j_o && k >= 1 && k <= g.sz[2] - 2) {
    const real2 low = *reinterpret_cast<const real2 *>(dp + c - g.plane);
    real2 wv = *reinterpret_cast<real2 *>(w + c);
    if (int_i_o0) wv.x -= (d0 - low.x) * g.one_over_dz * dt_s;
    if (int_i_o1) wv.y -= (d1 - low.y) * g.one_over_dz * dt_s;
    *reinterpret_cast<real2 *>(w + c) = wv;
  }
}

// rhs(face) +-= 2 g_n / h with g given as host-filled face tables (src/PressureEquation.cpp:10-56).
// The six faces are applied one launch at a time in the reference order because edge points receive
// the contribution of every face they lie on.
__global__ void __launch_bounds__(256) nhn_face_kernel(const Geom g, real *rhs, const real *table, int face) {
  const int dir = 2 - face / 2;
  const int na = (dir == 0) ? g.Ny : g.Nx;
  const int nb = (dir == 2) ? g.Ny : g.Nz;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)na * nb) return;
  const int a = (int)(t % na), b = (int)(t / na);
  const bool upper = face & 1;
  int i, j, k;
  real scale;
  if (dir == 2) {
    i = a; j = b; k = upper ? g.Nz - 1 : 0; scale = RC(2.0) * g.one_over_dz;
  } else if (dir == 1) {
    i = a; k = b; j = upper ? g.Ny - 1 : 0; scale = RC(2.0) * g.one_over_dy;
  } else {
    j = a; k = b; i = upper ? g.Nx - 1 : 0; scale = RC(2.0) * g.one_over_dx;
  }
  const real term = table[t] * scale;
  const long long c = gidx(g, i, j, k);
  if (upper) rhs[c] -= term;
  else rhs[c] += term;
}


// ------------------------------------------------------------------------------------------------
// Diagnostics on the device (src/Norms.cpp:11-118, adjust_pressure src/PressureEquation.cpp:288-343): per-CTA
// partial results, summed on the host in CTA order (deterministic; the reference's serial summation order differs,
// SURVEY section 8a asks for ~1e-12 relative agreement only).
// ------------------------------------------------------------------------------------------------
constexpr int kDiagThreads = 256;

// Block-wide sums of up to 3 values and a maximum; the result is valid in thread 0.
__device__ __forceinline__ void block_reduce(real &s0, real &s1, real &s2, real &mx) {
  __shared__ real red[4][kDiagThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_down_sync(0xffffffffu, s0, o);
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    red[0][warp] = s0; red[1][warp] = s1; red[2][warp] = s2; red[3][warp] = mx;
  }
  __syncthreads();
  if (tid == 0) {
    s0 = s1 = s2 = RC(0.0);
    mx = RC(0.0);
    for (int wi = 0; wi < kDiagThreads / 32; wi++) {
      s0 += red[0][wi]; s1 += red[1][wi]; s2 += red[2][wi]; mx = fmax(mx, red[3][wi]);
    }
  }
}

// compute_error for the velocity (src/Norms.cpp:11-47): components averaged to the pressure points, interior points
// only.  partial[4 b + {0,1,2}] = sum of |e|_2, sum of |e|_2^2, max of the component errors over the column block b.
__global__ void __launch_bounds__(kDiagThreads)
velocity_error_kernel(const Geom g, const real *__restrict__ u, const real *__restrict__ v, const real *__restrict__ w,
                      const BcDev bc, real *__restrict__ partial) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  real l1 = RC(0.0), l2 = RC(0.0), unused = RC(0.0), linf = RC(0.0);
  if (i <= g.Nx - 2 && j <= g.Ny - 2) {
    const real x = g.min_x + (g.base_i + i) * g.dx, y = g.min_y + (g.base_j + j) * g.dy;
    for (int k = 1; k <= g.Nz - 2; k++) {
      const real z = g.min_z + (g.base_k + k) * g.dz;
      const long long c = gidx(g, i, j, k);
      const real eu = exact_velocity(bc.kind, 0, bc.time, x, y, z, bc.Re) - (u[c] + u[c + 1]) / RC(2.0);
      const real ev = exact_velocity(bc.kind, 1, bc.time, x, y, z, bc.Re) - (v[c] + v[c + g.PX]) / RC(2.0);
      const real ew = exact_velocity(bc.kind, 2, bc.time, x, y, z, bc.Re) - (w[c] + w[c + g.plane]) / RC(2.0);
      const real sq = eu * eu + ev * ev + ew * ew;
      l1 += sqrt(sq);
      l2 += sq;
      linf = fmax(linf, fmax(fabs(eu), fmax(fabs(ev), fabs(ew))));
    }
  }
  block_reduce(l1, l2, unused, linf);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    real *out = partial + 4 * (blockIdx.y * gridDim.x + blockIdx.x);
    out[0] = l1; out[1] = l2; out[2] = linf; out[3] = RC(0.0);
  }
}

// compute_error for a scalar (src/Norms.cpp:88-101) and the sum of adjust_pressure (src/PressureEquation.cpp:294-296):
// owner points.  partial[4 b + {0,1,2,3}] = sum |e|, sum e^2, max |e|, sum e  with e = exact - p.
__global__ void __launch_bounds__(kDiagThreads)
pressure_error_kernel(const Geom g, const real *__restrict__ p, const BcDev bc, real *__restrict__ partial) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + g.own_lo[0], j = blockIdx.y * blockDim.y + threadIdx.y + g.own_lo[1];
  real l1 = RC(0.0), l2 = RC(0.0), sum = RC(0.0), linf = RC(0.0);
  if (i < g.own_hi[0] && j < g.own_hi[1]) {
    const real x = g.min_x + (g.base_i + i) * g.dx, y = g.min_y + (g.base_j + j) * g.dy;
    for (int k = g.own_lo[2]; k < g.own_hi[2]; k++) {
      const real z = g.min_z + (g.base_k + k) * g.dz;
      const real e = exact_pressure(bc.kind, bc.time, x, y, z, bc.Re) - p[gidx(g, i, j, k)];
      l1 += fabs(e);
      l2 += e * e;
      sum += e;
      linf = fmax(linf, fabs(e));
    }
  }
  block_reduce(l1, l2, sum, linf);
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    real *out = partial + 4 * (blockIdx.y * gridDim.x + blockIdx.x);
    out[0] = l1; out[1] = l2; out[2] = linf; out[3] = sum;
  }
}

// pressure += difference on all points of the tensor, ghosts included (src/PressureEquation.cpp:335-342).
__global__ void __launch_bounds__(256) add_constant_kernel(const Geom g, real *__restrict__ p, real difference) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
  if (i < g.Nx && j < g.Ny) p[gidx(g, i, j, k)] += difference;
}

// Slab (all y, local z) <-> blocks per destination rank (its y rows, local z), whole padded x rows.
template <bool PACK>
__global__ void __launch_bounds__(256) slab_pack_kernel(const Geom g, real *field, real *buf, const int *__restrict__ ylo,
                                                        int nranks) {
  const int nz = g.own_hi[2] - g.own_lo[2];
  const int y = blockIdx.y, zl = blockIdx.z;  // y: row index inside the owner region
  int r = 0;
  while (r + 1 < nranks && y >= ylo[r + 1]) r++;
  const int ny_r = ylo[r + 1] - ylo[r];
  const long long block = (long long)nz * g.PX * ylo[r];
  real *row_buf = buf + block + ((long long)zl * ny_r + (y - ylo[r])) * g.PX;
  real *row_field = field + gidx(g, 0, g.own_lo[1] + y, g.own_lo[2] + zl);
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < g.PX; x += gridDim.x * blockDim.x) {
    if (PACK) row_buf[x] = row_field[x];
    else row_field[x] = row_buf[x];
  }
}

inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

}  // namespace

void launch_stage(cudaStream_t stream, const Geom &g, int stage, CVec3 in, const real *pressure, Vec3 a, Vec3 b,
                  uint64_t *launches, PlaneRange planes) {
  const int ni = max(g.sx[0], g.Nx) - 2, nj = max(g.sy[1], g.Ny) - 2, nk_all = max(g.sz[2], g.Nz) - 2;
  // interior planes 1 .. nk_all, or the sub-range [planes.first, planes.first + planes.count) of them (0-based)
  const int k_shift = planes.count < 0 ? 0 : planes.first;
  const int nk = planes.count < 0 ? nk_all : min(planes.count, nk_all - planes.first);
  if (ni <= 0 || nj <= 0 || nk <= 0) return;
  static const int prefetch_planes = getenv("MIFGPU_STAGE_PREFETCH") ? atoi(getenv("MIFGPU_STAGE_PREFETCH")) : 1;
  const dim3 block(64, 4, 1);
  // y chunks of about 2.2 MB per array and plane (the plane size of the 513^3 case, where one launch reads every
  // input exactly once from HBM); MIFGPU_STAGE_CHUNK_MB overrides.
  static const real chunk_mb = getenv("MIFGPU_STAGE_CHUNK_MB") ? atof(getenv("MIFGPU_STAGE_CHUNK_MB")) : RC(2.2);
  const int blocks_y = (int)cdiv(nj, block.y);
  int chunk_blocks_y = (int)(chunk_mb * RC(1e6) / ((real)g.PX * sizeof(real) * block.y));
  chunk_blocks_y = max(1, min(chunk_blocks_y, blocks_y));
  int n_chunks = (int)cdiv(blocks_y, chunk_blocks_y);
  while ((long long)n_chunks * nk > 65535 && chunk_blocks_y < blocks_y) {  // gridDim.z limit
    chunk_blocks_y *= 2;
    n_chunks = (int)cdiv(blocks_y, chunk_blocks_y);
  }
  chunk_blocks_y = (int)cdiv(blocks_y, n_chunks);  // equal chunks
  const dim3 pblock(32, 4, 1);  // 64 x-points by 4 rows per CTA
  const dim3 pgrid(cdiv(ni + 1, 2 * pblock.x), chunk_blocks_y, nk * n_chunks);
  if (stage == 1)
    stage_kernel_pair<1><<<pgrid, pblock, 0, stream>>>(g, in.c[0], in.c[1], in.c[2], pressure, a.c[0], a.c[1], a.c[2], b.c[0],
                                                       b.c[1], b.c[2], prefetch_planes, nk, chunk_blocks_y, k_shift);
  else if (stage == 2)
    stage_kernel_pair<2><<<pgrid, pblock, 0, stream>>>(g, in.c[0], in.c[1], in.c[2], pressure, a.c[0], a.c[1], a.c[2], b.c[0],
                                                       b.c[1], b.c[2], prefetch_planes, nk, chunk_blocks_y, k_shift);
  else
    stage_kernel_pair<3><<<pgrid, pblock, 0, stream>>>(g, in.c[0], in.c[1], in.c[2], pressure, a.c[0], a.c[1], a.c[2], b.c[0],
                                                       b.c[1], b.c[2], prefetch_planes, nk, chunk_blocks_y, k_shift);
  ++*launches;
}

void launch_periodic(cudaStream_t stream, const Geom &g, real *field, int comp, uint64_t *launches) {
  const int sx = g.sx[comp], sy = g.sy[comp], sz = g.sz[comp];
  // x, then y, then z: later copies read the ghosts written by earlier ones (src/StaggeredTensor.cpp:224-256).
  if (g.periodic[0]) {
    periodic_copy_kernel<<<cdiv((long long)sy * sz, 256), 256, 0, stream>>>(g, field, sx, sy, sz, 0);
    ++*launches;
  }
  if (g.periodic[1] && g.prev_y == -1) {  // periodic_bc[1] && Py == 1
    periodic_copy_kernel<<<cdiv((long long)sx * sz, 256), 256, 0, stream>>>(g, field, sx, sy, sz, 1);
    ++*launches;
  }
  if (g.periodic[2] && g.prev_z == -1) {  // periodic_bc[2] && Pz == 1
    periodic_copy_kernel<<<cdiv((long long)sx * sy, 256), 256, 0, stream>>>(g, field, sx, sy, sz, 2);
    ++*launches;
  }
}

void launch_apply_bc(cudaStream_t stream, const Geom &g, Vec3 vel, const BcDev &bc, uint64_t *launches) {
  long long max_area = 0;
  for (int c = 0; c < 3; c++) {
    const long long sx = g.sx[c], sy = g.sy[c], sz = g.sz[c];
    max_area = max(max_area, max(sx * sy, max(sx * sz, sy * sz)));
  }
  if (!(g.periodic[0] && g.periodic[1] && g.periodic[2])) {
    const dim3 grid(cdiv(max_area, 256), 18, 1);
    bc_face_kernel<<<grid, 256, 0, stream>>>(g, vel.c[0], vel.c[1], vel.c[2], bc);
    ++*launches;
  }
  for (int c = 0; c < 3; c++) launch_periodic(stream, g, vel.c[c], c, launches);
}

void launch_divergence(cudaStream_t stream, const Geom &g, CVec3 vel, real, real dt, real *rhs,
                       uint64_t *launches, PlaneRange planes) {
  const int nj = g.own_hi[1] - g.own_lo[1], nk_all = g.own_hi[2] - g.own_lo[2];
  const int k_shift = planes.count < 0 ? 0 : planes.first;  // owner planes, 0-based
  const int nk = planes.count < 0 ? nk_all : min(planes.count, nk_all - planes.first);
  if (nk <= 0) return;
  const dim3 block(64, 4, 1);
  const dim3 grid(cdiv(g.own_hi[0], 2 * block.x), cdiv(nj, block.y), nk);  // pairs (i, i+1), i even, from i = 0
  divergence_kernel<<<grid, block, 0, stream>>>(g, vel.c[0], vel.c[1], vel.c[2], dt, rhs, k_shift);
  ++*launches;
}

void launch_nhn_rhs(cudaStream_t stream, const Geom &g, real *rhs, const BcDev &bc, uint64_t *launches) {
  for (int face = 0; face < 6; face++) {
    const int dir = 2 - face / 2;
    const long long na = (dir == 0) ? g.Ny : g.Nx, nb = (dir == 2) ? g.Ny : g.Nz;
    nhn_face_kernel<<<cdiv(na * nb, 256), 256, 0, stream>>>(g, rhs, bc.tables[dir][face], face);
    ++*launches;
  }
}

void launch_pack_slab(cudaStream_t stream, const Geom &g, const real *field, real *send, const int *ylo_dev,
                      int nranks, uint64_t *launches) {
  const dim3 grid(cdiv(g.PX, 256), g.own_hi[1] - g.own_lo[1], g.own_hi[2] - g.own_lo[2]);
  slab_pack_kernel<true><<<grid, 256, 0, stream>>>(g, const_cast<real *>(field), send, ylo_dev, nranks);
  ++*launches;
}

void launch_unpack_slab(cudaStream_t stream, const Geom &g, real *field, const real *recv, const int *ylo_dev,
                        int nranks, uint64_t *launches) {
  const dim3 grid(cdiv(g.PX, 256), g.own_hi[1] - g.own_lo[1], g.own_hi[2] - g.own_lo[2]);
  slab_pack_kernel<false><<<grid, 256, 0, stream>>>(g, field, const_cast<real *>(recv), ylo_dev, nranks);
  ++*launches;
}

void launch_correct(cudaStream_t stream, const Geom &g, Vec3 vel, real *pressure, const real *dp, real dt_s,
                    uint64_t *launches, PlaneRange planes) {
  const int k_shift = planes.count < 0 ? 0 : planes.first;  // planes of the (ghosted) tensors, 0-based
  const int nk = planes.count < 0 ? g.sz[2] : min(planes.count, g.sz[2] - planes.first);
  if (nk <= 0) return;
  const dim3 block(64, 4, 1);
  const dim3 grid(cdiv(g.sx[0], 2 * block.x), cdiv(g.sy[1], block.y), nk);
  correct_kernel<<<grid, block, 0, stream>>>(g, vel.c[0], vel.c[1], vel.c[2], pressure, dp, dt_s, k_shift);
  ++*launches;
}

void launch_velocity_stage(cudaStream_t stream, const Geom &g, int stage, CVec3 in, Vec3 rhs_buf, Vec3 out, real time,
                           real Re, uint64_t *launches) {
  const int ni = max(g.sx[0], g.Nx) - 2, nj = max(g.sy[1], g.Ny) - 2, nk = max(g.sz[2], g.Nz) - 2;
  if (ni <= 0 || nj <= 0 || nk <= 0) return;
  const dim3 block(64, 4, 1), grid(cdiv(ni, block.x), cdiv(nj, block.y), nk);
  if (stage == 1)
    velocity_stage_kernel<1><<<grid, block, 0, stream>>>(g, in.c[0], in.c[1], in.c[2], rhs_buf.c[0], rhs_buf.c[1], rhs_buf.c[2],
                                                         out.c[0], out.c[1], out.c[2], time, Re);
  else if (stage == 2)
    velocity_stage_kernel<2><<<grid, block, 0, stream>>>(g, in.c[0], in.c[1], in.c[2], rhs_buf.c[0], rhs_buf.c[1], rhs_buf.c[2],
                                                         out.c[0], out.c[1], out.c[2], time, Re);
  else
    velocity_stage_kernel<3><<<grid, block, 0, stream>>>(g, in.c[0], in.c[1], in.c[2], rhs_buf.c[0], rhs_buf.c[1], rhs_buf.c[2],
                                                         out.c[0], out.c[1], out.c[2], time, Re);
  ++*launches;
}

int diag_blocks(const Geom &g, bool velocity) {
  const int ni = velocity ? g.Nx - 2 : g.own_hi[0] - g.own_lo[0], nj = velocity ? g.Ny - 2 : g.own_hi[1] - g.own_lo[1];
  if (ni <= 0 || nj <= 0) return 0;
  return (int)(cdiv(ni, 64) * cdiv(nj, 4));
}

void launch_velocity_error(cudaStream_t stream, const Geom &g, CVec3 vel, const BcDev &bc, real *partial,
                           uint64_t *launches) {
  const int ni = g.Nx - 2, nj = g.Ny - 2;
  if (ni <= 0 || nj <= 0) return;
  const dim3 block(64, 4, 1), grid(cdiv(ni, 64), cdiv(nj, 4), 1);
  velocity_error_kernel<<<grid, block, 0, stream>>>(g, vel.c[0], vel.c[1], vel.c[2], bc, partial);
  ++*launches;
}

void launch_pressure_error(cudaStream_t stream, const Geom &g, const real *p, const BcDev &bc, real *partial,
                           uint64_t *launches) {
  const int ni = g.own_hi[0] - g.own_lo[0], nj = g.own_hi[1] - g.own_lo[1];
  if (ni <= 0 || nj <= 0) return;
  const dim3 block(64, 4, 1), grid(cdiv(ni, 64), cdiv(nj, 4), 1);
  pressure_error_kernel<<<grid, block, 0, stream>>>(g, p, bc, partial);
  ++*launches;
}

void launch_add_constant(cudaStream_t stream, const Geom &g, real *p, real difference, uint64_t *launches) {
  const dim3 block(64, 4, 1), grid(cdiv(g.Nx, 64), cdiv(g.Ny, 4), g.Nz);
  add_constant_kernel<<<grid, block, 0, stream>>>(g, p, difference);
  ++*launches;
}

}  // namespace mifgpu
