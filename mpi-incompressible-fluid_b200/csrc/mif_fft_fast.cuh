// mif_fft_fast.cuh -- register-blocked power-of-two line transforms for the Poisson sweeps (sm_100a, FP64).
//
// One CTA transforms a tile of 8 lines.  A line of N = M+1 real points (DCT-I, FFTW_REDFT00) is the real
// part of the DFT of its even extension of period 2M; that extension is packed two reals per complex
// (c_j = e_{2j} + i e_{2j+1}, j < M) and transformed by ONE complex FFT of length M = 2^LOGM, followed by
// the usual real-FFT unpack, which also folds the evenness:
//     E_k = 1/2 [ (C_k + conj C_{M-k}) - i exp(-i pi k / M) (C_k - conj C_{M-k}) ]   (real),  k = 0..M.
// (Same algorithm as oracle/fft_cpu.h mo_r2r_exec; no prefix sums or 1/sin factors, so the error stays
// at the O(log M) level of a plain FFT.)
//
// Layout in shared memory: complex slot q of line l lives at S[q*9 + l] (8 lines + 1 pad), so the eight
// lanes of a quarter warp -- always the eight lines at the same slot -- touch 128 contiguous bytes and
// every access pattern of the FFT is bank-conflict free; the pad makes the transposing accesses of the
// x sweeps (consecutive slots of one line) conflict free as well.  The eight lines share their twiddles,
// which are therefore broadcast loads.
//
// FFT: Stockham autosort, M/8 threads per line, 8 complex values per thread in registers, radix-8 passes
// (plus one radix-4 or radix-2 pass when log2 M is not a multiple of 3), natural order in and out.
#pragma once

#include <cuda_runtime.h>

namespace mifgpu {
namespace fast {

constexpr int kLines = 8;       // lines per CTA
constexpr int kSlotPitch = 9;   // complex elements per slot row (8 lines + 1 pad)

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 mul_neg_i(double2 a) { return make_double2(a.y, -a.x); }  // a * (-i)

// Forward 8-point DFT in registers (decimation in frequency), natural order in and out.
__device__ __forceinline__ void dft8(double2 *a) {
  const double h = 0.70710678118654752440;
  double2 b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
  double2 b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
  double2 b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]);
  double2 b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
  b5 = make_double2((b5.x + b5.y) * h, (b5.y - b5.x) * h);    // * exp(-i pi/4)
  b6 = mul_neg_i(b6);                                         // * exp(-i pi/2)
  b7 = make_double2((b7.y - b7.x) * h, -(b7.x + b7.y) * h);   // * exp(-3 i pi/4)
  const double2 c0 = cadd(b0, b2), c2 = csub(b0, b2), c1 = cadd(b1, b3), c3 = mul_neg_i(csub(b1, b3));
  const double2 c4 = cadd(b4, b6), c6 = csub(b4, b6), c5 = cadd(b5, b7), c7 = mul_neg_i(csub(b5, b7));
  a[0] = cadd(c0, c1); a[4] = csub(c0, c1); a[2] = cadd(c2, c3); a[6] = csub(c2, c3);
  a[1] = cadd(c4, c5); a[5] = csub(c4, c5); a[3] = cadd(c6, c7); a[7] = csub(c6, c7);
}

__device__ __forceinline__ void dft4(double2 &a0, double2 &a1, double2 &a2, double2 &a3) {
  const double2 c0 = cadd(a0, a2), c2 = csub(a0, a2), c1 = cadd(a1, a3), c3 = mul_neg_i(csub(a1, a3));
  a0 = cadd(c0, c1); a2 = csub(c0, c1); a1 = cadd(c2, c3); a3 = csub(c2, c3);
}

// One Stockham pass of radix R on the 8 values v[s] = x[j + s*T] held by thread j of a line.
//   NS = product of the radices of the previous passes.  tw[q] = exp(-2 pi i q / M).
// For R < 8 the thread performs 8/R butterflies: butterfly u uses v[u + (8/R) t], t < R.
template <int LOGM, int R, int NS>
__device__ __forceinline__ void stockham_pass(double2 *S, int line, int j, const double2 *__restrict__ tw,
                                              double2 *v) {
  constexpr int M = 1 << LOGM, T = M / 8, G = 8 / R;  // G butterflies per thread
#pragma unroll
  for (int s = 0; s < 8; s++) v[s] = S[(j + s * T) * kSlotPitch + line];
  if (NS > 1) {
#pragma unroll
    for (int u = 0; u < G; u++) {
      const int jj = j + u * T;                 // butterfly index in [0, M/R)
      const int k = jj & (NS - 1);
#pragma unroll
      for (int t = 1; t < R; t++) {
        const double2 w = __ldg(&tw[(k * t) * (M / (NS * R))]);
        v[u + G * t] = cmul(v[u + G * t], w);
      }
    }
  }
  if (R == 8) {
    dft8(v);
  } else if (R == 4) {
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
  } else {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double2 a = v[u], b = v[u + 4];
      v[u] = cadd(a, b);
      v[u + 4] = csub(a, b);
    }
  }
  __syncthreads();  // every thread has read its inputs
#pragma unroll
  for (int u = 0; u < G; u++) {
    const int jj = j + u * T;
    const int k = jj & (NS - 1);
    const int base = ((jj - k) * R) + k;
#pragma unroll
    for (int t = 0; t < R; t++) S[(base + t * NS) * kSlotPitch + line] = v[u + G * t];
  }
  __syncthreads();
}

// Forward complex FFT of length M of the 8 lines held in S (natural order in and out).
template <int LOGM>
__device__ __forceinline__ void fft_lines(double2 *S, int line, int j, const double2 *__restrict__ tw) {
  double2 v[8];
  static_assert(LOGM >= 6 && LOGM <= 10, "fast path covers M = 64 .. 1024");
  stockham_pass<LOGM, 8, 1>(S, line, j, tw, v);
  stockham_pass<LOGM, 8, 8>(S, line, j, tw, v);
  if (LOGM == 7) stockham_pass<LOGM, 2, 64>(S, line, j, tw, v);
  if (LOGM == 8) stockham_pass<LOGM, 4, 64>(S, line, j, tw, v);
  if (LOGM >= 9) stockham_pass<LOGM, 8, 64>(S, line, j, tw, v);
  if (LOGM == 10) stockham_pass<LOGM, 2, 512>(S, line, j, tw, v);
}

// Packed position of real element e (0 <= e <= 2M-1) of the even extension: slot e>>1, component e&1.
__device__ __forceinline__ void put_packed(double *Sd, int M, int e, int line, double value) {
  Sd[((e >> 1) * kSlotPitch + line) * 2 + (e & 1)] = value;
  if (e > 0 && e < M) {
    const int r = 2 * M - e;  // mirror image, same parity as e
    Sd[((r >> 1) * kSlotPitch + line) * 2 + (r & 1)] = value;
  }
}

// DCT-I unpack: thread j of a line produces E_k and E_{M-k} for k = j + T s, s < 4 (k < M/2), from the pair
// (C_k, C_{M-k}); k = 0 yields (E_0, E_M); thread 0 also produces E_{M/2}.  cs[k] = (cos, sin)(pi k / M).
template <int LOGM>
__device__ __forceinline__ void dct_unpack(const double2 *S, int line, int j, const double2 *__restrict__ cs,
                                           double *lo, double *hi, double &mid) {
  constexpr int M = 1 << LOGM, T = M / 8;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = j + T * s;
    const double2 A = S[k * kSlotPitch + line];
    if (k == 0) {
      lo[s] = A.x + A.y;
      hi[s] = A.x - A.y;
    } else {
      const double2 B = S[(M - k) * kSlotPitch + line];
      const double2 w = __ldg(&cs[k]);
      const double sum_r = A.x + B.x, dif_r = A.x - B.x, sum_i = A.y + B.y;
      const double rot = w.x * sum_i - w.y * dif_r;
      lo[s] = 0.5 * (sum_r + rot);
      hi[s] = 0.5 * (sum_r - rot);
    }
  }
  mid = (j == 0) ? S[(M / 2) * kSlotPitch + line].x : 0.0;
}

}  // namespace fast
}  // namespace mifgpu
