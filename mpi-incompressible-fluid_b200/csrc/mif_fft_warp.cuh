// mif_fft_warp.cuh -- warp-per-line DCT-I sweeps for the Poisson solve (sm_100a, FP64).
//
// Same mathematics as mif_fft_fast.cuh (even extension packed two reals per complex, one complex FFT of
// length M = 2^LOGM, real-FFT unpack), different mapping: every line of a tile of 8 lines belongs to ONE warp
// (two warps for M = 1024, four for M = 2048), which keeps 16 complex values per lane in registers and exchanges them through
// its own shared-memory region between the radix-8 Stockham passes.  Because no other warp touches that
// region, the passes need only __syncwarp(): the 16 resident warps of an SM run their lines completely
// asynchronously, which is what hides the shared-memory and global latencies (the CTA-synchronous variant
// in mif_fft_fast.cuh spends most of its time waiting at barriers).  CTA-wide barriers remain only around
// the coalesced global load / store of the strided (y, z) sweeps; x sweeps have none.
//
// Shared layout of one line: complex slot q at position q + (q >> 3) (one pad every 8 slots), lines
// 8 * (M + M/8) + 1 complex apart.  With lanes along the line this makes every pass access conflict free:
// reads are contiguous, and the stride-8 writes of the first pass become stride-9.  The odd line pitch keeps
// the line-fastest accesses of the strided load/store phases conflict free.
#pragma once

#include <cuda_runtime.h>

#include "mif_fft_fast.cuh"  // cadd / csub / cmul / dft8 / dft4

namespace mifgpu {
namespace warpfft {

using fast::cadd;
using fast::cmul;
using fast::csub;

constexpr int kLines = 8;

template <int LOGM>
struct Cfg {
  static constexpr int M = 1 << LOGM;
  static constexpr int WPL = (LOGM >= 11) ? 4 : ((LOGM >= 10) ? 2 : 1);  // warps per line
  static constexpr int TL = 32 * WPL;                       // threads per line
  static constexpr int EPT = M / TL;                        // complex values per thread
  static constexpr int LINES = (LOGM >= 11) ? 4 : kLines;   // lines per CTA (M = 2048: 4 lines x 128 threads)
  static constexpr int THREADS = LINES * TL;
  static constexpr int LINE_PITCH = M + M / 8 + 1;          // complex elements between lines
  static constexpr int TW_PASS2 = 3 * 8;                    // twiddles t = 1, 2, 4 of the NS = 8 pass
  static constexpr int TW_PASS3 = 3 * 64;                   // ... of the NS = 64 pass
  static constexpr int TW_PASS4 = (LOGM == 10) ? 512 : ((LOGM == 11) ? 2 * 512 : 0);  // NS = 512 pass: radix 2 (M = 1024), radix 4 (M = 2048)
  static constexpr int TW_TOTAL = TW_PASS2 + TW_PASS3 + TW_PASS4;
  static constexpr size_t SMEM = (size_t)(LINES * LINE_PITCH + TW_TOTAL) * sizeof(double2);
};

__device__ __forceinline__ int pad(int q) { return q + (q >> 3); }

template <int WPL>
__device__ __forceinline__ void line_sync(int line) {
  if (WPL == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(line + 1), "r"(32 * WPL) : "memory");
}

// Copy the twiddles the passes need into shared memory: T2[ti][k] = W_64^(k t), T3[ti][k] = W_(64 R3)^(k t)
// with t = 1, 2, 4 (ti = 0, 1, 2), T4[k] = W_1024^k.  tw[q] = exp(-2 pi i q / M).
// TW_STRIDE > 1: tw is the table of the TW_STRIDE times longer transform (tw[q * TW_STRIDE] = exp(-2 pi i q / M)).
template <int LOGM, int TW_STRIDE = 1, int THREADS = Cfg<LOGM>::THREADS>
__device__ __forceinline__ void load_twiddles(double2 *T, const double2 *__restrict__ tw) {
  using C = Cfg<LOGM>;
  constexpr int M = C::M;
  constexpr int R3 = (LOGM == 7) ? 2 : (LOGM == 8 ? 4 : 8);  // radix of the NS = 64 pass
  for (int idx = threadIdx.x; idx < C::TW_TOTAL; idx += THREADS) {
    int q;
    if (idx < C::TW_PASS2) {
      const int ti = idx / 8, k = idx - ti * 8;
      q = (k << ti) * (M / 64);
    } else if (idx < C::TW_PASS2 + C::TW_PASS3) {
      const int r = idx - C::TW_PASS2, ti = r / 64, k = r - ti * 64;
      q = (k << ti) * (M / (64 * R3));
    } else if (LOGM == 11) {
      const int r = idx - C::TW_PASS2 - C::TW_PASS3, ti = r / 512, k = r - ti * 512;  // NS = 512, R = 4: W_2048^(k t), t = 1, 2
      q = k << ti;
    } else {
      q = idx - C::TW_PASS2 - C::TW_PASS3;  // NS = 512, R = 2: W_1024^k
    }
    T[idx] = __ldg(&tw[(q & (M - 1)) * TW_STRIDE]);
  }
}

// One Stockham pass of radix R over the EPT values v[s] = x[j + s*TL] of lane j (NS = product of previous radices).
// LOAD = false: v[] was filled by the caller (first pass of the x sweeps, straight from global memory);
// STORE = false: the results stay in registers (last pass, followed by the shuffle unpack).
template <int LOGM, int R, int NS, bool LOAD = true, bool STORE = true>
__device__ __forceinline__ void pass(double2 *S, const double2 *T, int j, int line, double2 *v) {
  using C = Cfg<LOGM>;
  constexpr int TL = C::TL, EPT = C::EPT, G = EPT / R;
  if (LOAD) {
#pragma unroll
    for (int s = 0; s < EPT; s++) v[s] = S[pad(j + s * TL)];
  }
#pragma unroll
  for (int u = 0; u < G; u++) {
    const int jj = j + u * TL;
    if (NS > 1) {
      const int k = jj & (NS - 1);
      if (R == 8) {
        const double2 w1 = T[k], w2 = T[NS + k], w4 = T[2 * NS + k];
        const double2 w3 = cmul(w1, w2);
        v[u + G * 1] = cmul(v[u + G * 1], w1);
        v[u + G * 2] = cmul(v[u + G * 2], w2);
        v[u + G * 3] = cmul(v[u + G * 3], w3);
        v[u + G * 4] = cmul(v[u + G * 4], w4);
        v[u + G * 5] = cmul(v[u + G * 5], cmul(w4, w1));
        v[u + G * 6] = cmul(v[u + G * 6], cmul(w4, w2));
        v[u + G * 7] = cmul(v[u + G * 7], cmul(w4, w3));
      } else if (R == 4) {
        const double2 w1 = T[k], w2 = T[NS + k];
        v[u + G * 1] = cmul(v[u + G * 1], w1);
        v[u + G * 2] = cmul(v[u + G * 2], w2);
        v[u + G * 3] = cmul(v[u + G * 3], cmul(w1, w2));
      } else {
        v[u + G] = cmul(v[u + G], T[k]);
      }
    }
    if (R == 8) {
      double2 a[8];
#pragma unroll
      for (int t = 0; t < 8; t++) a[t] = v[u + G * t];
      fast::dft8(a);
#pragma unroll
      for (int t = 0; t < 8; t++) v[u + G * t] = a[t];
    } else if (R == 4) {
      fast::dft4(v[u], v[u + G], v[u + 2 * G], v[u + 3 * G]);
    } else {
      const double2 a = v[u], b = v[u + G];
      v[u] = cadd(a, b);
      v[u + G] = csub(a, b);
    }
  }
  if (!STORE) return;
  if (LOAD) line_sync<C::WPL>(line);  // every lane of the line has read its inputs
#pragma unroll
  for (int u = 0; u < G; u++) {
    const int jj = j + u * TL;
    const int k = jj & (NS - 1);
    const int base = (jj - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; t++) S[pad(base + t * NS)] = v[u + G * t];
  }
  line_sync<C::WPL>(line);
}

// Radix of the last pass and the number of butterflies per lane in it.
template <int LOGM>
struct LastPass {
  static constexpr int R = (LOGM == 7 || LOGM == 10) ? 2 : ((LOGM == 8 || LOGM == 11) ? 4 : 8);
  static constexpr int NS = (1 << LOGM) / R;
  static constexpr int G = Cfg<LOGM>::EPT / R;
};

// Forward complex FFT of one line (padded, natural order in S); T = twiddle tables of load_twiddles.
//   FROM_REGS: v[] already holds x[j + s*TL] (the first pass does not read shared memory);
//   KEEP:      the spectrum stays in registers, element (u, t) of the last pass, v[u + G t] = X[j + TL u + NS t].
//   SKIP_FIRST: the first pass has already been done and its output stored (strided sweeps, first_pass_in_place).
template <int LOGM, bool FROM_REGS, bool KEEP, bool SKIP_FIRST = false>
__device__ __forceinline__ void fft_line(double2 *S, const double2 *T, int j, int line, double2 *v) {
  using C = Cfg<LOGM>;
  const double2 *T2 = T, *T3 = T + C::TW_PASS2, *T4 = T3 + C::TW_PASS3;
  if (!SKIP_FIRST) pass<LOGM, 8, 1, !FROM_REGS, true>(S, T, j, line, v);
  pass<LOGM, 8, 8>(S, T2, j, line, v);
  if (LOGM == 7) pass<LOGM, 2, 64, true, !KEEP>(S, T3, j, line, v);
  if (LOGM == 8) pass<LOGM, 4, 64, true, !KEEP>(S, T3, j, line, v);
  if (LOGM == 9) pass<LOGM, 8, 64, true, !KEEP>(S, T3, j, line, v);
  if (LOGM == 10) {
    pass<LOGM, 8, 64>(S, T3, j, line, v);
    pass<LOGM, 2, 512, true, !KEEP>(S, T4, j, line, v);
  }
  if (LOGM == 11) {
    pass<LOGM, 8, 64>(S, T3, j, line, v);
    pass<LOGM, 4, 512, true, !KEEP>(S, T4, j, line, v);
  }
}

// First radix-8 pass (no twiddles) on values v[s] = c[b + s*TL] that the caller loaded straight from global memory,
// executed by the thread that loaded them -- in the strided sweeps that is thread (line l, butterfly index b) of the
// line-fastest mapping, not a lane of the line's warp.  Stores the pass output into line l's region; the caller
// separates this from the following passes with __syncthreads().
template <int LOGM>
__device__ __forceinline__ void first_pass_in_place(double2 *S_line, int b, double2 *v) {
  using C = Cfg<LOGM>;
  constexpr int TL = C::TL, G = C::EPT / 8;
#pragma unroll
  for (int u = 0; u < G; u++) {
    double2 a[8];
#pragma unroll
    for (int t = 0; t < 8; t++) a[t] = v[u + G * t];
    fast::dft8(a);
    const int jj = b + u * TL;
#pragma unroll
    for (int t = 0; t < 8; t++) S_line[pad(jj * 8 + t)] = a[t];
  }
}

// The same in two steps, for kernels that want the butterflies done before the line regions may be written
// (mif_poisson_tma.cuh: the regions alias an output stage that the copy engine may still be reading).
template <int LOGM>
__device__ __forceinline__ void first_pass_compute(double2 *v) {
  constexpr int G = Cfg<LOGM>::EPT / 8;
#pragma unroll
  for (int u = 0; u < G; u++) {
    double2 a[8];
#pragma unroll
    for (int t = 0; t < 8; t++) a[t] = v[u + G * t];
    fast::dft8(a);
#pragma unroll
    for (int t = 0; t < 8; t++) v[u + G * t] = a[t];
  }
}
template <int LOGM>
__device__ __forceinline__ void first_pass_store(double2 *S_line, int b, const double2 *v) {
  constexpr int TL = Cfg<LOGM>::TL, G = Cfg<LOGM>::EPT / 8;
#pragma unroll
  for (int u = 0; u < G; u++)
#pragma unroll
    for (int t = 0; t < 8; t++) S_line[pad((b + u * TL) * 8 + t)] = v[u + G * t];
}

// DCT-I unpack in registers (one warp per line): after the last pass lane j holds C_k for k = j + 32 u + NS t.
// Its partners C_{M-k} all live in lane 32 - j (butterfly G-1-u, output R-1-t), so one round of shuffles
// replaces the shared-memory round trip; lane 0 is its own partner with a slightly different index map.
// out[u + G t] = E_k for the same k; E_M is returned separately (valid in lane 0).
//
// SPLIT (lines of 2M + 1 points, mif_poisson.cu warp_dct_split_kernel): the warp holds one half of a radix-2
// decimation-in-frequency split of the length-2M transform, C_{2k} (ODD = false) or C_{2k+1} (ODD = true), and
// produces E_{2k} resp. E_{2k+1} of the long line.  cs is then the table of the long transform, (cos, sin)(pi q / 2M).
// Even half: partners and angles are those of the short transform (cs[2k]).  Odd half: the partner of C_{2k+1} is
// C_{2M-2k-1} = C_{2(M-1-k)+1}, i.e. index M-1-k, which lives in lane 31 - j (butterfly G-1-u, output R-1-t) with no
// special case, and the angle is pi (2k+1) / 2M (cs[2k+1]).
template <int LOGM, bool SPLIT = false, bool ODD = false>
__device__ __forceinline__ void unpack_regs(const double2 *v, int j, const double2 *__restrict__ cs, double *out,
                                            double &e_last) {
  using L = LastPass<LOGM>;
  constexpr int R = L::R, G = L::G;
  static_assert(Cfg<LOGM>::WPL == 1, "shuffle unpack needs the whole line in one warp");
  static_assert(SPLIT || !ODD, "ODD only exists in the split scheme");
  const int src = ODD ? 31 - j : (32 - j) & 31;
  // cos / sin of pi t / 8, t = 0..7
  constexpr double kRotCos[8] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173,
                                 0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128675613};
  constexpr double kRotSin[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128675613,
                                 1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};
  double2 base_w[G];
#pragma unroll
  for (int u = 0; u < G; u++) base_w[u] = __ldg(&cs[SPLIT ? 2 * (j + 32 * u) + (ODD ? 1 : 0) : j + 32 * u]);
#pragma unroll
  for (int u = 0; u < G; u++)
#pragma unroll
    for (int t = 0; t < R; t++) {
      const double2 mine = v[(G - 1 - u) + G * (R - 1 - t)];
      double2 B = make_double2(__shfl_sync(0xffffffffu, mine.x, src), __shfl_sync(0xffffffffu, mine.y, src));
      // lane 0: k = 32 u + NS t.  u = 0: partner NS (R - t) in the same butterfly (t = 0 pairs with itself);
      // u >= 1: partner in butterfly G - u, output R - 1 - t.
      if (!ODD && j == 0) B = (u == 0) ? v[G * ((R - t) % R)] : v[(G - u) + G * (R - 1 - t)];
      // exp(i pi k / M) for k = k0 + NS t is exp(i pi k0 / M) times the constant exp(i pi t / R): one table load per
      // butterfly instead of one per output (the loads go through the same L1/shared pipe that bounds this kernel).
      const double2 A = v[u + G * t];
      const double2 w0 = base_w[u];
      const double cr = kRotCos[(8 / R) * t], sr = kRotSin[(8 / R) * t];
      const double wx = w0.x * cr - w0.y * sr, wy = w0.x * sr + w0.y * cr;
      out[u + G * t] = 0.5 * ((A.x + B.x) + wx * (A.y + B.y) - wy * (A.x - B.x));
    }
  e_last = v[0].x - v[0].y;  // E_M = Re C_0 - Im C_0 (meaningful in lane 0)
}

// cos / sin of pi t / 8, t = 0..7: exp(i pi k / M) for k = k0 + NS t is exp(i pi k0 / M) times exp(i pi t / R).
__device__ __forceinline__ double2 rot8(int t) {
  constexpr double kRotCos[8] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173,
                                 0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128675613};
  constexpr double kRotSin[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128675613,
                                 1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};
  return make_double2(kRotCos[t], kRotSin[t]);
}

// Real FFT of n = 2M points (FFTW_R2HC) from the length-M complex FFT of the packed line c_q = x_{2q} + i x_{2q+1}:
//   X_k = 1/2 [ (C_k + conj C_{M-k}) - i exp(-2 pi i k / n) (C_k - conj C_{M-k}) ],  k = 0 .. M  (C_M = C_0).
// Same register / lane layout as unpack_regs: lane j holds C_k for k = j + 32 u + NS t and receives C_{M-k} from lane
// 32 - j.  re/im[u + G t] = X_k; X_0 and X_M are real, X_M is returned in x_last (lane 0).  cs[k] = (cos, sin)(2 pi k / n).
template <int LOGM>
__device__ __forceinline__ void unpack_r2hc_regs(const double2 *v, int j, const double2 *__restrict__ cs, double *re,
                                                 double *im, double &x_last) {
  using L = LastPass<LOGM>;
  constexpr int R = L::R, G = L::G;
  static_assert(Cfg<LOGM>::WPL == 1, "shuffle unpack needs the whole line in one warp");
  const int src = (32 - j) & 31;
  double2 base_w[G];
#pragma unroll
  for (int u = 0; u < G; u++) base_w[u] = __ldg(&cs[j + 32 * u]);
#pragma unroll
  for (int u = 0; u < G; u++)
#pragma unroll
    for (int t = 0; t < R; t++) {
      const double2 mine = v[(G - 1 - u) + G * (R - 1 - t)];
      double2 B = make_double2(__shfl_sync(0xffffffffu, mine.x, src), __shfl_sync(0xffffffffu, mine.y, src));
      if (j == 0) B = (u == 0) ? v[G * ((R - t) % R)] : v[(G - u) + G * (R - 1 - t)];
      const double2 A = v[u + G * t];
      const double2 w0 = base_w[u], r = rot8((8 / R) * t);
      const double c = w0.x * r.x - w0.y * r.y, sn = w0.x * r.y + w0.y * r.x;
      const double sum_r = A.x + B.x, dif_i = A.y - B.y, dif_r = A.x - B.x, sum_i = A.y + B.y;
      re[u + G * t] = 0.5 * (sum_r + c * sum_i - sn * dif_r);
      im[u + G * t] = 0.5 * (dif_i - c * dif_r - sn * sum_i);
    }
  x_last = v[0].x - v[0].y;  // X_M = Re C_0 - Im C_0 (meaningful in lane 0)
}

// Input of the inverse (FFTW_HC2R, unnormalised) as ONE complex FFT:  Z_k = (X_k + conj X_{M-k}) + i exp(+2 pi i k / n)
// (X_k - conj X_{M-k}), z = inverse FFT_M(Z) = conj(FFT_M(conj Z)),  x_{2q} = Re z_q, x_{2q+1} = Im z_q.
// Returns conj(Z_k) from X_k = (xr, xi), X_{M-k} = (yr, yi) and (c, sn) = (cos, sin)(2 pi k / n).
__device__ __forceinline__ double2 hc2r_input(double xr, double xi, double yr, double yi, double c, double sn) {
  const double pr = xr + yr, pi = xi - yi, dr = xr - yr, di = xi + yi;
  return make_double2(pr - sn * dr - c * di, -(pi + c * dr - sn * di));
}

// The same from the registers left by unpack_r2hc_regs (after the eigenvalue scaling): the partners X_{M-k} come
// from lane 32 - j again; the results are the first-pass inputs v[s] of the next transform because
// k = j + 32 (u + G t) is exactly slot s = u + G t of lane j.
template <int LOGM>
__device__ __forceinline__ void pack_hc2r_regs(const double *re, const double *im, double x_last, int j,
                                               const double2 *__restrict__ cs, double2 *v) {
  using L = LastPass<LOGM>;
  constexpr int R = L::R, G = L::G;
  const int src = (32 - j) & 31;
  double2 base_w[G];
#pragma unroll
  for (int u = 0; u < G; u++) base_w[u] = __ldg(&cs[j + 32 * u]);
#pragma unroll
  for (int u = 0; u < G; u++)
#pragma unroll
    for (int t = 0; t < R; t++) {
      const int partner = (G - 1 - u) + G * (R - 1 - t);
      double yr = __shfl_sync(0xffffffffu, re[partner], src), yi = __shfl_sync(0xffffffffu, im[partner], src);
      if (j == 0) {
        if (u == 0 && t == 0) {
          yr = x_last;  // X_M
          yi = 0.0;
        } else {
          const int own = (u == 0) ? G * (R - t) : (G - u) + G * (R - 1 - t);
          yr = re[own];
          yi = im[own];
        }
      }
      const double2 w0 = base_w[u], r = rot8((8 / R) * t);
      const double c = w0.x * r.x - w0.y * r.y, sn = w0.x * r.y + w0.y * r.x;
      const double xi = (j == 0 && u == 0 && t == 0) ? 0.0 : im[u + G * t];
      v[u + G * t] = hc2r_input(re[u + G * t], xi, yr, yi, c, sn);
    }
}

// Store real element e (0 <= e <= M) of the even extension and its mirror image 2M - e into the packed line.
__device__ __forceinline__ void put_packed(double *Sd, int M, int e, double value) {
  Sd[pad(e >> 1) * 2 + (e & 1)] = value;
  if (e > 0 && e < M) {
    const int r = 2 * M - e;
    Sd[pad(r >> 1) * 2 + (r & 1)] = value;
  }
}

// DCT-I unpack: lane j produces E_k and E_{M-k} for k = j + TL*s, s < EPT/2 (k < M/2); k = 0 gives (E_0, E_M);
// lane 0 also produces E_{M/2}.  cs[k] = (cos, sin)(pi k / M).
template <int LOGM>
__device__ __forceinline__ void unpack_line(const double2 *S, int j, const double2 *__restrict__ cs, double *lo, double *hi,
                                            double &mid) {
  using C = Cfg<LOGM>;
  constexpr int M = C::M, TL = C::TL, PAIRS = C::EPT / 2;
#pragma unroll
  for (int s = 0; s < PAIRS; s++) {
    const int k = j + TL * s;
    const double2 A = S[pad(k)];
    if (k == 0) {
      lo[s] = A.x + A.y;
      hi[s] = A.x - A.y;
    } else {
      const double2 B = S[pad(M - k)];
      const double2 w = __ldg(&cs[k]);
      const double sum_r = A.x + B.x, dif_r = A.x - B.x, sum_i = A.y + B.y;
      const double rot = w.x * sum_i - w.y * dif_r;
      lo[s] = 0.5 * (sum_r + rot);
      hi[s] = 0.5 * (sum_r - rot);
    }
  }
  mid = (j == 0) ? S[pad(M / 2)].x : 0.0;
}

}  // namespace warpfft
}  // namespace mifgpu
