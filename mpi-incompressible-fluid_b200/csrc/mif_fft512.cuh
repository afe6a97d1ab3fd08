// mif_fft512.cuh -- the 512-point complex FFT behind the 513-point DCT-I lines, as 16 x 32 (sm_100a, FP64).
//
// The sweeps of a 513^3 grid are bound by the shared-memory / shuffle data pipe, not by HBM (ncu: 70-90 % of the LSU
// wavefront peak, profiles/r02_*), so the transform is organised to cross that pipe as rarely as possible: 16 values
// per lane, radix-16 butterflies in registers, ONE exchange through shared memory and ONE exchange between two lanes
// (the radix-8 Stockham version of mif_fft_warp.cuh needs two round trips).
//
//   n = n1 + 32 n2 (n1 < 32, n2 < 16),   k = 16 k1 + k2 (k1 < 32, k2 < 16)
//   C[16 k1 + k2] = sum_n1 W_32^(n1 k1) { W_512^(n1 k2) [ sum_n2 c[n1 + 32 n2] W_16^(n2 k2) ] }
//
//   phase A  the thread that holds c[n1 + 32 s], s < 16: 16-point DFT over s, times W_512^(n1 k2); the 16 results go
//            to shared memory, slot n1 + 33 k2 of the line's region (conflict free for 16-byte accesses);
//   phase B  lane L = k2 + 16 p of the line's warp reads the 16 values n1 = 2 m + p of its k2, 16-point DFT over m,
//            lanes p = 1 multiply by W_32^k, and the last radix-2 step pairs lanes L and L ^ 16 (each lane sends 8
//            values and receives 8).  Register r of lane L ends up with C_k,
//                k = k2 + 16 (r & 7) + 128 p + 256 (r >> 3).
//
// DCT-I unpack (same formula as mif_fft_warp.cuh): the partner C_{M-k} of register r sits in register 15 - r of lane
// 32 - L (lanes 0 and 16 pair with each other, register 16 - r), so one round of shuffles does it, and the unpack
// twiddle of register r is one table value times a constant rotation.
#pragma once

#include <cuda_runtime.h>

#include "mif_fft_fast.cuh"  // cadd / csub / cmul / dft4

namespace mifgpu {
namespace fft512 {

using fast::cadd;
using fast::cmul;
using fast::csub;

constexpr int kM = 512;
constexpr int kRowStride = 33;             // complex slots between consecutive k2 rows of a line region
constexpr int kRegion = 16 * kRowStride;   // 528 slots used per line
constexpr int kLinePitch = kRegion + 1;    // odd: the 8 line regions of a tile start on different banks
constexpr int kTwiddles = 4 * 32;          // T[32 e + n1] = W_512^(n1 2^e), e < 4

// W_N^q = exp(-2 pi i q / N) as compile-time constants
__device__ __forceinline__ double2 w16(int q) {
  constexpr double c[4] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};
  // cos(pi q / 8), -sin(pi q / 8) for q = 0..15 from the first quadrant
  const int quad = (q >> 2) & 3, rem = q & 3;
  const double co = c[rem], si = (rem == 0) ? 0.0 : c[4 - rem];
  // exp(-i (pi/2) quad) * (co - i si)
  if (quad == 0) return make_double2(co, -si);
  if (quad == 1) return make_double2(-si, -co);
  if (quad == 2) return make_double2(-co, si);
  return make_double2(si, co);
}
__device__ __forceinline__ double2 w32(int q) {  // exp(-i pi q / 16), q = 0..15
  constexpr double c[9] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708,
                           0.70710678118654752440, 0.55557023301960222474, 0.38268343236508977173,
                           0.19509032201612826785, 0.0};
  return (q <= 8) ? make_double2(c[q], -c[8 - q]) : make_double2(-c[16 - q], -c[q - 8]);
}
// (cos, sin)(pi t / 32), t = 0..7: rotation between the unpack twiddles of registers r and r + 1
__device__ __forceinline__ double2 rot32(int t) {
  constexpr double co[8] = {1.0, 0.99518472667219688624, 0.98078528040323044913, 0.95694033573220886494,
                            0.92387953251128675613, 0.88192126434835502971, 0.83146961230254523708,
                            0.77301045336273696081};
  constexpr double si[8] = {0.0, 0.09801714032956060199, 0.19509032201612826785, 0.29028467725446236764,
                            0.38268343236508977173, 0.47139673682599764856, 0.55557023301960222474,
                            0.63439328416364549822};
  return make_double2(co[t], si[t]);
}
// (cos, sin)(pi s / 32), s = 0..15
__device__ __forceinline__ double2 rot32x(int s) {
  if (s < 8) return rot32(s);
  if (s == 8) return make_double2(0.70710678118654752440, 0.70710678118654752440);
  const double2 r = rot32(16 - s);  // cos(pi s / 32) = sin(pi (16 - s) / 32)
  return make_double2(r.y, r.x);
}
// (cos, sin)(pi s / 16), s = 0..15
__device__ __forceinline__ double2 rot16(int s) {
  const double2 w = w32(s);
  return make_double2(w.x, -w.y);
}

// Forward 16-point DFT in registers, natural order in and out (4 x 4).
__device__ __forceinline__ void dft16(double2 *a) {
#pragma unroll
  for (int s1 = 0; s1 < 4; s1++) fast::dft4(a[s1], a[s1 + 4], a[s1 + 8], a[s1 + 12]);  // a[s1 + 4 k2] = sum_s2 a[s1 + 4 s2] W_4^(s2 k2)
#pragma unroll
  for (int s1 = 1; s1 < 4; s1++)
#pragma unroll
    for (int k2 = 1; k2 < 4; k2++) {
      const int e = s1 * k2;
      if (e == 4) a[s1 + 4 * k2] = fast::mul_neg_i(a[s1 + 4 * k2]);
      else a[s1 + 4 * k2] = cmul(a[s1 + 4 * k2], w16(e));
    }
#pragma unroll
  for (int k2 = 0; k2 < 4; k2++) fast::dft4(a[4 * k2], a[4 * k2 + 1], a[4 * k2 + 2], a[4 * k2 + 3]);  // a[k1 + 4 k2] = X[4 k1 + k2]
  // natural order: X[4 k1 + k2] sits at a[k1 + 4 k2]
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int jj = i + 1; jj < 4; jj++) {
      const double2 t = a[i + 4 * jj];
      a[i + 4 * jj] = a[jj + 4 * i];
      a[jj + 4 * i] = t;
    }
}

// Twiddle table of the kernel's shared memory from the plan's full table tw[q * STRIDE] = exp(-2 pi i q / 512)
// (STRIDE = 2: the table of a 1024-point transform, mif_poisson_tma.cuh tma_dct1024_kernel).
template <int THREADS, int STRIDE = 1>
__device__ __forceinline__ void load_twiddles(double2 *T, const double2 *__restrict__ tw) {
  for (int idx = threadIdx.x; idx < kTwiddles; idx += THREADS) {
    const int e = idx >> 5, n1 = idx & 31;
    T[idx] = __ldg(&tw[((n1 << e) & (kM - 1)) * STRIDE]);
  }
}

// Phase A on v[s] = c[n1 + 32 s]: leaves v[k2] = W_512^(n1 k2) sum_s v[s] W_16^(s k2).
__device__ __forceinline__ void phase_a(double2 *v, int n1, const double2 *T) {
  dft16(v);
  const double2 w1 = T[n1], w2 = T[32 + n1], w4 = T[64 + n1], w8 = T[96 + n1];
  const double2 w3 = cmul(w1, w2), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
  v[1] = cmul(v[1], w1);
  v[2] = cmul(v[2], w2);
  v[3] = cmul(v[3], w3);
  v[4] = cmul(v[4], w4);
  v[5] = cmul(v[5], w5);
  v[6] = cmul(v[6], w6);
  v[7] = cmul(v[7], w7);
  v[8] = cmul(v[8], w8);
  v[9] = cmul(v[9], cmul(w8, w1));
  v[10] = cmul(v[10], cmul(w8, w2));
  v[11] = cmul(v[11], cmul(w8, w3));
  v[12] = cmul(v[12], cmul(w8, w4));
  v[13] = cmul(v[13], cmul(w8, w5));
  v[14] = cmul(v[14], cmul(w8, w6));
  v[15] = cmul(v[15], cmul(w8, w7));
}
__device__ __forceinline__ void store_a(double2 *region, int n1, const double2 *v) {
#pragma unroll
  for (int k2 = 0; k2 < 16; k2++) region[n1 + kRowStride * k2] = v[k2];
}

// Phase B by lane L of the line's warp (every lane of the warp must call it): v[r] = C_k,
// k = (L & 15) + 16 (r & 7) + 128 (L >> 4) + 256 (r >> 3).
__device__ __forceinline__ void phase_b(const double2 *region, int L, double2 *v) {
  const int k2 = L & 15, p = L >> 4;
  const double2 *src = region + p + kRowStride * k2;
#pragma unroll
  for (int m = 0; m < 16; m++) v[m] = src[2 * m];
  dft16(v);
  if (p) {
#pragma unroll
    for (int q = 1; q < 16; q++) v[q] = cmul(v[q], w32(q));
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const double2 keep = p ? v[8 + i] : v[i], send = p ? v[i] : v[8 + i];
    const double2 got = make_double2(__shfl_xor_sync(0xffffffffu, send.x, 16), __shfl_xor_sync(0xffffffffu, send.y, 16));
    const double2 sum = cadd(keep, got), dif = csub(keep, got);
    v[i] = sum;                                               // k1 = 8 p + i
    v[8 + i] = p ? make_double2(-dif.x, -dif.y) : dif;        // k1 = 16 + 8 p + i: (even part) - (odd part)
  }
}

// Index of the spectrum element in register r of lane L.
__device__ __forceinline__ int k_of(int L, int r) { return (L & 15) + 16 * (r & 7) + 128 * (L >> 4) + 256 * (r >> 3); }

__device__ __forceinline__ double2 shuffle_from(double2 value, int src) {
  return make_double2(__shfl_sync(0xffffffffu, value.x, src), __shfl_sync(0xffffffffu, value.y, src));
}

// DCT-I unpack in registers: out[r] = E_k for k = k_of(L, r); E_M is returned in e_last (valid in lane 0).
// cs[k] = (cos, sin)(pi k / 512).
// HALF = 1 / 2: the warp holds the even (C_{2k}) / odd (C_{2k+1}) half of a radix-2 split 1024-point transform and
// produces E_{2k} / E_{2k+1} of a 1025-point line; cs is then the table of the long transform, (cos, sin)(pi q / 1024).
// Odd half: the partner of C_{2k+1} is C_{2(511-k)+1}, register 15 - r of lane 31 - L, with no special lanes.
template <int HALF = 0>
__device__ __forceinline__ void unpack_dct(const double2 *v, int L, const double2 *__restrict__ cs, double *out, double &e_last) {
  // E_k and E_{M-k} come from the same pair (A, B) = (C_k, C_{M-k}): with (wx, wy) = (cos, sin)(pi k / M),
  //   S = A.x + B.x,  T = wx (A.y + B.y) - wy (A.x - B.x),  E_k = (S + T) / 2,  E_{M-k} = (S - T) / 2
  // (the twiddle of M - k is (-wx, wy)).  Every lane therefore unpacks only its registers r < 8, for which it fetches the
  // partner's register 15 - r (>= 8), and hands the second result back to the partner, whose register 15 - r it is:
  // half the twiddle rotations and products of unpacking all sixteen registers, and 24 shuffles instead of 32.
  // Lanes 0 and 16 (k2 = 0) pair with each other, register r with register 16 - r; their register 0 pairs inside the lane:
  // lane 0 has k = 0 (partner itself, second result E_M) and k = 256 in register 8 (its own partner, E_256 = Re C_256),
  // lane 16 has k = 128 with partner k = 384 in its own register 8.
  const int k2 = L & 15, p = L >> 4;
  const bool special = (HALF != 2) && (k2 == 0);
  const int src = (HALF == 2) ? 31 - L : (special ? (L ^ 16) : 32 - L);
  const double2 base = __ldg(&cs[HALF == 0 ? k2 + 128 * p : 2 * (k2 + 128 * p) + (HALF == 2 ? 1 : 0)]);
  double back[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const double2 send = special ? v[(16 - r) & 15] : v[15 - r];
    double2 B = shuffle_from(send, src);
    if (special && r == 0) B = p ? v[8] : v[0];
    const double2 A = v[r];
    const double2 rt = rot32(r);
    const double wx = (r == 0) ? base.x : base.x * rt.x - base.y * rt.y, wy = (r == 0) ? base.y : base.x * rt.y + base.y * rt.x;
    const double half_s = 0.5 * (A.x + B.x);
    const double t = wx * (A.y + B.y) - wy * (A.x - B.x);
    out[r] = 0.5 * t + half_s;
    const double image = half_s - 0.5 * t;  // E_{M-k}: register 15 - r (special lanes: 16 - r) of lane src
    back[r] = __shfl_sync(0xffffffffu, image, src);
    if (r == 0) e_last = image;  // lane 0: E_M = Re C_0 - Im C_0 (meaningful in lane 0 only)
  }
  // what the partner computed for this lane's upper registers
#pragma unroll
  for (int r = 0; r < 7; r++) out[15 - r] = special ? back[r + 1] : back[r];
  // register 8: general lanes the partner's last image; lane 16 its own pair (k = 128, 384); lane 0 the self-paired k = 256
  out[8] = special ? (p ? e_last : v[8].x) : back[7];
}

// conj Z_k for a real spectrum (see mif_poisson_tma.cuh): X_k = xr, X_{M-k} = yr, (c, sn) = (cos, sin)(pi k / M).
__device__ __forceinline__ double2 pack_input(double xr, double yr, double c, double sn) {
  const double pr = xr + yr, dr = xr - yr;
  return make_double2(pr - sn * dr, -(c * dr));
}

// First-pass inputs of the inverse transform from a real spectrum in the natural layout: nat[s] = E_k, k = L + 32 s,
// e_last = E_M (needed in lane 0); v[s] = conj Z_k.  The partners E_{M-k} come from lane 32 - L, register 15 - s
// (lane 0: its own register 16 - s, and E_M for s = 0).
__device__ __forceinline__ void pack_natural(const double *nat, double e_last, int L, const double2 *__restrict__ cs, double2 *v) {
  const int src = (32 - L) & 31;
  const double2 base = __ldg(&cs[L]);
#pragma unroll
  for (int s = 0; s < 16; s++) {
    double yr = __shfl_sync(0xffffffffu, nat[15 - s], src);
    if (L == 0) yr = (s == 0) ? e_last : nat[(16 - s) & 15];  // k = 32 s: M - k = 32 (16 - s)
    const double2 rt = rot16(s);
    const double c = base.x * rt.x - base.y * rt.y, sn = base.x * rt.y + base.y * rt.x;
    v[s] = pack_input(nat[s], yr, c, sn);
  }
}

// From the spectrum layout of phase B (spec[r] = E_k, k = k_of(L, r); e_last = E_M in lane 0) to the first-pass inputs
// of the inverse transform in the natural layout, v[s] = conj Z_k for k = L + 32 s: the registers whose k is not
// congruent to L modulo 32 change places with lane L ^ 16, then the partners E_{M-k} come from lane 32 - L.
__device__ __forceinline__ void repack_for_inverse(const double *spec, double e_last, int L, const double2 *__restrict__ cs, double2 *v) {
  const int p = L >> 4;
  double nat[16];  // nat[s] = E_k, k = L + 32 s
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double even = spec[2 * a + 8 * c], odd = spec[2 * a + 1 + 8 * c];  // k = k2 + 32 a + 128 p + 256 c (+ 16)
      const double keep = p ? odd : even, send = p ? even : odd;
      const double got = __shfl_xor_sync(0xffffffffu, send, 16);
      nat[a + 8 * c] = p ? got : keep;      // s = a + 8 c: from the lane with p = 0
      nat[a + 4 + 8 * c] = p ? keep : got;  // s = a + 4 + 8 c: from the lane with p = 1
    }
  pack_natural(nat, e_last, L, cs, v);
}

}  // namespace fft512
}  // namespace mifgpu
