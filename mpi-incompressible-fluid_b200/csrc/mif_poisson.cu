// mif_poisson.cu -- spectral pressure Poisson solve on the device (sm_100a, FP64), replacing
// mif::solve_pressure_equation (src/PressureEquation.cpp:65-264), the FFTW plans of
// PressureSolverStructures (src/PressureSolverStructures.cpp:23-41) and the 2Decomp pencil transposes
// (deps/2Decomp_C/Transpose*.cpp), which on one GPU become strided tile accesses: every sweep reads a
// tile of L lines straight out of the x-fastest array into shared memory, transforms them there and
// writes them back in place, so no global transpose and no eigenvalue array exist.
//
// Transforms (FFTW definitions, see oracle/fft_cpu.h):
//   non-periodic direction: FFTW_REDFT00 (DCT-I) on N_global points, forward and inverse
//   periodic direction:     FFTW_R2HC forward / FFTW_HC2R inverse on N_global-1 points
// each reduced to ONE complex DFT per line (length N-1 for DCT-I, n/2 for even-length real FFTs, n for
// odd), executed as a power-of-two shared-memory FFT or, for other lengths, Bluestein's chirp-z on a
// power-of-two FFT.  The sweep along z does forward transform, division by the eigenvalues
// (src/PressureEquation.cpp:158-163), inverse transform and normalisation in one kernel.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#include <algorithm>

#include "mif_kernels.h"
#ifndef MIFGPU_FP32  // the tuned transform kernels are FP64 (tile geometry, register FFTs); the float build runs sweep_kernel
#include "mif_fft512.cuh"
#include "mif_fft_fast.cuh"
#include "mif_fft_warp.cuh"
#include "mif_poisson_tma.cuh"
#endif

namespace mifgpu {

namespace {

const double kPi = 3.14159265358979323846264338327950288;

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel AND per plan: the attribute belongs to the device
// the plan's context runs on, so a process that drives several GPUs sets it on each of them.
struct SmemAttrOnce {
  std::set<const void *> done;
  template <class K>
  void ensure(K kernel, size_t bytes) {
    if (done.insert(reinterpret_cast<const void *>(kernel)).second)
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  }
};

struct DirPlanDev {
  int periodic;   // 0: DCT-I, 1: halfcomplex real FFT
  int n;          // real points per line
  int m;          // complex DFT length
  int P, logP;    // executed power-of-two FFT size
  int bluestein;  // m is not a power of two
  const real2 *tw;      // P/2: exp(-2 pi i q / P)
  const real2 *tw_full; // P:   exp(-2 pi i q / P), all q (fast path)
  const real2 *chirp;   // m:   exp(-i pi j^2 / m)                         (Bluestein)
  const real2 *filt;    // P:   FFT_P(wrapped conj chirp) / P, bit-reversed (Bluestein)
  const real2 *unpack;  // (cos, sin)(pi k / m) for DCT-I, (cos, sin)(2 pi k / n) for real FFTs
  const real *lambda;   // n eigenvalues of the 1-D second difference in FFTW output order
  real inv_norm;        // 1 / (2 (N_global - 1)) or 1 / (N_global - 1)
};

struct SweepJob {
  DirPlanDev plan;
  int dir;             // 0 = x, 1 = y, 2 = z
  int L;               // lines per CTA
  int n_tile_lines;    // number of lines along the tiled dimension
  long long lstride;   // element offset between consecutive lines of a tile
  long long estride;   // element offset between consecutive points of a line
  int r_pitch;         // shared-memory pitch of one real line
  int mode;            // 0 forward, 1 inverse (+normalise), 2 forward * eigen, inverse (+normalise)
  const real *lam_a, *lam_b;  // eigenvalues of the two other directions (mode 2)
  bool has_origin;              // this launch holds the (0,0,0) mode at tile line 0, outer index 0
};

__device__ __forceinline__ real2 cmul(real2 a, real2 b) {
  return make_real2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// Radix-2 decimation-in-frequency stages, natural order in, bit-reversed order out -- executed two stages per pass
// over shared memory (a radix-4 butterfly on w[i0 + {0, 1, 2, 3} q]): the same butterflies and table twiddles as two
// single-stage passes, half the shared-memory round trips and barriers.  A last single stage remains when logP is odd.
__device__ void fft_dif(real2 *W, int L, int P, int logP, const real2 *__restrict__ tw, bool conj_tw) {
  const int halfP = P >> 1, quarterP = P >> 2;
  int s = logP - 1;
  for (; s >= 1; s -= 2) {
    const int q = 1 << (s - 1);      // stage s pairs (i, i + 2q), stage s - 1 pairs (i, i + q)
    const int tstride = halfP >> s;  // twiddle stride of stage s; stage s - 1 uses 2 * tstride
    for (int item = threadIdx.x; item < L * quarterP; item += blockDim.x) {
      const int l = item / quarterP, g = item - l * quarterP;
      const int k = g & (q - 1);
      const int i0 = ((g >> (s - 1)) << (s + 1)) + k;
      real2 *w = W + (size_t)l * P;
      const real2 a = w[i0], b = w[i0 + q], c = w[i0 + 2 * q], d = w[i0 + 3 * q];
      real2 t0 = __ldg(&tw[k * tstride]), t1 = __ldg(&tw[(k + q) * tstride]), t2 = __ldg(&tw[2 * k * tstride]);
      if (conj_tw) {
        t0.y = -t0.y;
        t1.y = -t1.y;
        t2.y = -t2.y;
      }
      const real2 a1 = make_real2(a.x + c.x, a.y + c.y), c1 = cmul(make_real2(a.x - c.x, a.y - c.y), t0);
      const real2 b1 = make_real2(b.x + d.x, b.y + d.y), d1 = cmul(make_real2(b.x - d.x, b.y - d.y), t1);
      w[i0] = make_real2(a1.x + b1.x, a1.y + b1.y);
      w[i0 + q] = cmul(make_real2(a1.x - b1.x, a1.y - b1.y), t2);
      w[i0 + 2 * q] = make_real2(c1.x + d1.x, c1.y + d1.y);
      w[i0 + 3 * q] = cmul(make_real2(c1.x - d1.x, c1.y - d1.y), t2);
    }
    __syncthreads();
  }
  if (s == 0) {  // half = 1, twiddle 1
    for (int item = threadIdx.x; item < L * halfP; item += blockDim.x) {
      const int l = item / halfP, i0 = (item - l * halfP) << 1;
      real2 *w = W + (size_t)l * P;
      const real2 a = w[i0], b = w[i0 + 1];
      real2 t = __ldg(&tw[0]);
      if (conj_tw) t.y = -t.y;
      w[i0] = make_real2(a.x + b.x, a.y + b.y);
      w[i0 + 1] = cmul(make_real2(a.x - b.x, a.y - b.y), t);
    }
    __syncthreads();
  }
}

// Radix-2 decimation-in-time stages, bit-reversed order in, natural order out, two stages per pass like fft_dif.
__device__ void fft_dit(real2 *W, int L, int P, int logP, const real2 *__restrict__ tw, bool conj_tw) {
  const int halfP = P >> 1, quarterP = P >> 2;
  int s = 0;
  for (; s + 1 < logP; s += 2) {
    const int h = 1 << s;                    // stage s pairs (i, i + h), stage s + 1 pairs (i, i + 2h)
    const int tstride = halfP >> (s + 1);    // twiddle stride of stage s + 1; stage s uses 2 * tstride
    for (int item = threadIdx.x; item < L * quarterP; item += blockDim.x) {
      const int l = item / quarterP, g = item - l * quarterP;
      const int k = g & (h - 1);
      const int i0 = ((g >> s) << (s + 2)) + k;
      real2 *w = W + (size_t)l * P;
      real2 t0 = __ldg(&tw[2 * k * tstride]), t1 = __ldg(&tw[k * tstride]), t2 = __ldg(&tw[(k + h) * tstride]);
      if (conj_tw) {
        t0.y = -t0.y;
        t1.y = -t1.y;
        t2.y = -t2.y;
      }
      const real2 a = w[i0], b = cmul(w[i0 + h], t0), c = w[i0 + 2 * h], d = cmul(w[i0 + 3 * h], t0);
      const real2 a1 = make_real2(a.x + b.x, a.y + b.y), b1 = make_real2(a.x - b.x, a.y - b.y);
      const real2 c1 = cmul(make_real2(c.x + d.x, c.y + d.y), t1), d1 = cmul(make_real2(c.x - d.x, c.y - d.y), t2);
      w[i0] = make_real2(a1.x + c1.x, a1.y + c1.y);
      w[i0 + 2 * h] = make_real2(a1.x - c1.x, a1.y - c1.y);
      w[i0 + h] = make_real2(b1.x + d1.x, b1.y + d1.y);
      w[i0 + 3 * h] = make_real2(b1.x - d1.x, b1.y - d1.y);
    }
    __syncthreads();
  }
  if (s < logP) {  // last single stage, half = P / 2
    const int half = 1 << s;
    const int tstride = halfP >> s;
    for (int item = threadIdx.x; item < L * halfP; item += blockDim.x) {
      const int l = item / halfP, bfly = item - l * halfP;
      const int k = bfly & (half - 1);
      const int i0 = ((bfly >> s) << (s + 1)) + k;
      real2 *w = W + (size_t)l * P;
      real2 t = __ldg(&tw[k * tstride]);
      if (conj_tw) t.y = -t.y;
      const real2 a = w[i0], b = cmul(w[i0 + half], t);
      w[i0] = make_real2(a.x + b.x, a.y + b.y);
      w[i0 + half] = make_real2(a.x - b.x, a.y - b.y);
    }
    __syncthreads();
  }
}

// Complex DFT of length m of every line of the tile.  Input: W[l][0..m) in natural order.
// Output: element k of the spectrum is at W[l][spec_pos(k)].  inverse = unnormalised exp(+...).
__device__ void cdft_tile(real2 *W, int L, const DirPlanDev &pl, bool inverse) {
  const int P = pl.P, m = pl.m;
  if (!pl.bluestein) {
    fft_dif(W, L, P, pl.logP, pl.tw, inverse);
    return;
  }
  // Bluestein: X_k = w_k sum_j (x_j w_j) conj(w_{k-j}), w_j = exp(-i pi j^2/m); the inverse transform is
  // conj(DFT(conj x)).
  for (int item = threadIdx.x; item < L * P; item += blockDim.x) {
    const int l = item / P, j = item - l * P;
    real2 *w = W + (size_t)l * P;
    if (j < m) {
      real2 x = w[j];
      if (inverse) x.y = -x.y;
      w[j] = cmul(x, __ldg(&pl.chirp[j]));
    } else {
      w[j] = make_real2(RC(0.0), RC(0.0));
    }
  }
  __syncthreads();
  fft_dif(W, L, P, pl.logP, pl.tw, false);
  for (int item = threadIdx.x; item < L * P; item += blockDim.x) {
    const int l = item / P, j = item - l * P;
    real2 *w = W + (size_t)l * P;
    w[j] = cmul(w[j], __ldg(&pl.filt[j]));
  }
  __syncthreads();
  fft_dit(W, L, P, pl.logP, pl.tw, true);
  for (int item = threadIdx.x; item < L * m; item += blockDim.x) {
    const int l = item / m, k = item - l * m;
    real2 *w = W + (size_t)l * P;
    real2 x = cmul(w[k], __ldg(&pl.chirp[k]));
    if (inverse) x.y = -x.y;
    w[k] = x;
  }
  __syncthreads();
}

__device__ __forceinline__ int spec_pos(const DirPlanDev &pl, int k) {
  if (pl.bluestein) return k;
  return pl.logP == 0 ? 0 : (int)(__brev((unsigned)k) >> (32 - pl.logP));
}

// Forward real transform of every line: R (n reals, natural order) -> R (FFTW output order).
__device__ void real_forward(real *R, real2 *W, int L, int r_pitch, const DirPlanDev &pl) {
  const int n = pl.n, m = pl.m, P = pl.P;
  // pack
  for (int item = threadIdx.x; item < L * m; item += blockDim.x) {
    const int l = item / m, j = item - l * m;
    const real *r = R + (size_t)l * r_pitch;
    real2 c;
    if (!pl.periodic) {
      const int q0 = 2 * j, q1 = 2 * j + 1;  // even extension of period 2m
      c = make_real2(r[q0 <= m ? q0 : 2 * m - q0], r[q1 <= m ? q1 : 2 * m - q1]);
    } else if ((n & 1) == 0) {
      c = make_real2(r[2 * j], r[2 * j + 1]);
    } else {
      c = make_real2(r[j], RC(0.0));
    }
    W[(size_t)l * P + j] = c;
  }
  __syncthreads();
  cdft_tile(W, L, pl, false);
  // unpack
  if (!pl.periodic) {
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int l = item / n, k = item - l * n;
      const real2 *w = W + (size_t)l * P;
      const int k0 = (k == m) ? 0 : k, k1 = (k == 0) ? 0 : m - k;
      const real2 A = w[spec_pos(pl, k0)], B = w[spec_pos(pl, k1)];
      const real2 cs = __ldg(&pl.unpack[k]);
      R[(size_t)l * r_pitch + k] = RC(0.5) * ((A.x + B.x) + cs.x * (A.y + B.y) - cs.y * (A.x - B.x));
    }
  } else if ((n & 1) == 0) {
    const int h = m;
    for (int item = threadIdx.x; item < L * (h + 1); item += blockDim.x) {
      const int l = item / (h + 1), k = item - l * (h + 1);
      const real2 *w = W + (size_t)l * P;
      const int k0 = (k == h) ? 0 : k, k1 = (k == 0) ? 0 : h - k;
      const real2 A = w[spec_pos(pl, k0)], B = w[spec_pos(pl, k1)];
      const real2 cs = __ldg(&pl.unpack[k]);
      real *r = R + (size_t)l * r_pitch;
      r[k] = RC(0.5) * ((A.x + B.x) + cs.x * (A.y + B.y) - cs.y * (A.x - B.x));
      if (k > 0 && k < h) r[n - k] = RC(0.5) * ((A.y - B.y) - cs.x * (A.x - B.x) - cs.y * (A.y + B.y));
    }
  } else {
    const int h = n / 2;
    for (int item = threadIdx.x; item < L * (h + 1); item += blockDim.x) {
      const int l = item / (h + 1), k = item - l * (h + 1);
      const real2 A = W[(size_t)l * P + spec_pos(pl, k)];
      real *r = R + (size_t)l * r_pitch;
      r[k] = A.x;
      if (k > 0) r[n - k] = A.y;
    }
  }
  __syncthreads();
}

// Inverse real transform of every line (unnormalised): for DCT-I the same transform, for periodic
// directions FFTW_HC2R.  R (FFTW order) -> R (natural order).
__device__ void real_inverse(real *R, real2 *W, int L, int r_pitch, const DirPlanDev &pl) {
  if (!pl.periodic) {
    real_forward(R, W, L, r_pitch, pl);
    return;
  }
  const int n = pl.n, m = pl.m, P = pl.P;
  if ((n & 1) == 0) {
    const int h = m;
    for (int item = threadIdx.x; item < L * h; item += blockDim.x) {
      const int l = item / h, k = item - l * h;
      const real *r = R + (size_t)l * r_pitch;
      const int kk = h - k;
      const real xr = r[k], xi = (k == 0) ? RC(0.0) : r[n - k];
      const real yr = r[kk], yi = (kk == h) ? RC(0.0) : r[n - kk];
      const real pr = xr + yr, pi = xi - yi, dr = xr - yr, di = xi + yi;
      const real2 cs = __ldg(&pl.unpack[k]);
      W[(size_t)l * P + k] = make_real2(pr - cs.y * dr - cs.x * di, pi + cs.x * dr - cs.y * di);
    }
    __syncthreads();
    cdft_tile(W, L, pl, true);
    for (int item = threadIdx.x; item < L * h; item += blockDim.x) {
      const int l = item / h, j = item - l * h;
      const real2 A = W[(size_t)l * P + spec_pos(pl, j)];
      real *r = R + (size_t)l * r_pitch;
      r[2 * j] = A.x;
      r[2 * j + 1] = A.y;
    }
  } else {
    const int h = n / 2;
    for (int item = threadIdx.x; item < L * (h + 1); item += blockDim.x) {
      const int l = item / (h + 1), k = item - l * (h + 1);
      const real *r = R + (size_t)l * r_pitch;
      real2 *w = W + (size_t)l * P;
      if (k == 0) {
        w[0] = make_real2(r[0], RC(0.0));
      } else {
        w[k] = make_real2(r[k], r[n - k]);
        w[n - k] = make_real2(r[k], -r[n - k]);
      }
    }
    __syncthreads();
    cdft_tile(W, L, pl, true);
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int l = item / n, j = item - l * n;
      R[(size_t)l * r_pitch + j] = W[(size_t)l * P + spec_pos(pl, j)].x;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) sweep_kernel(const SweepJob job, real *__restrict__ field, long long origin,
                                                    long long tile_stride, long long outer_stride) {
  extern __shared__ real2 smem2[];
  const DirPlanDev &pl = job.plan;
  const int L = job.L, n = pl.n, r_pitch = job.r_pitch;
  real2 *W = smem2;
  real *R = reinterpret_cast<real *>(W + (size_t)L * pl.P);
  const int first_line = blockIdx.x * L;
  const int lines = min(L, job.n_tile_lines - first_line);
  real *base = field + origin + (long long)first_line * tile_stride + (long long)blockIdx.y * outer_stride;

  // load: element-fastest when points of a line are contiguous (x sweeps), line-fastest otherwise
  if (job.estride == 1) {
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int l = item / n, e = item - l * n;
      R[(size_t)l * r_pitch + e] = (l < lines) ? base[(long long)l * job.lstride + e] : RC(0.0);
    }
  } else {
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int e = item / L, l = item - e * L;
      R[(size_t)l * r_pitch + e] = (l < lines) ? base[(long long)l * job.lstride + (long long)e * job.estride] : RC(0.0);
    }
  }
  __syncthreads();

  if (job.mode == 0 || job.mode == 2) real_forward(R, W, L, r_pitch, pl);
  if (job.mode == 2) {
    // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z); mode (0,0,0) := 0  (src/PressureEquation.cpp:158-163,
    // table of src/PressureSolverStructures.cpp:52-69).  For the z sweep a tile line is x index
    // first_line + l and blockIdx.y is the y index.
    const real lam_y = job.lam_b[blockIdx.y];
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int l = item / n, e = item - l * n;
      if (l < lines) {
        const int ix = first_line + l;
        const real lam_x = job.lam_a[ix];
        real scale = RC(1.0) / (lam_x + lam_y + pl.lambda[e]);
        if (job.has_origin && ix == 0 && blockIdx.y == 0 && e == 0) scale = RC(0.0);
        R[(size_t)l * r_pitch + e] *= scale;
      }
    }
    __syncthreads();
  }
  if (job.mode == 1 || job.mode == 2) {
    real_inverse(R, W, L, r_pitch, pl);
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int l = item / n, e = item - l * n;
      R[(size_t)l * r_pitch + e] *= pl.inv_norm;
    }
    __syncthreads();
  }

  if (job.estride == 1) {
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int l = item / n, e = item - l * n;
      if (l < lines) base[(long long)l * job.lstride + e] = R[(size_t)l * r_pitch + e];
    }
  } else {
    for (int item = threadIdx.x; item < L * n; item += blockDim.x) {
      const int e = item / L, l = item - e * L;
      if (l < lines) base[(long long)l * job.lstride + (long long)e * job.estride] = R[(size_t)l * r_pitch + e];
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Fast path: DCT-I sweeps with N - 1 = M = 2^LOGM (64 <= M <= 1024), 8 lines per CTA, M threads.
// ------------------------------------------------------------------------------------------------
// Piecewise-strided addressing of a line for the multi-GPU sweeps: the points [lo[r], lo[r+1]) of every line live
// in segment r (a block of a staging buffer, possibly in the memory of peer GPU r mapped over NVLink).  The buffers
// are blocked in x by tiles of 8 doubles, with the line direction next:
//   address(e, outer, x) = base[r] + (x >> 3) * xtile_stride[r] + outer * outer_stride[r] + (e - lo[r]) * estride[r] + (x & 7)
// and estride = 8 wherever a sweep STORES: the 8 lines of a CTA times consecutive points of the lines are then one
// contiguous run in the destination, so a warp-wide store is 256 contiguous bytes (4 points x 8 lines) instead of
// four separate 64-byte pieces -- NVLink moves small writes at well under half its bandwidth.
constexpr int kMaxRanks = 8;
struct SegMap {
  int n = 0;  // 0: plain strided addressing
  int lo[kMaxRanks + 1];
  real *base[kMaxRanks];
  long long estride[kMaxRanks], outer_stride[kMaxRanks], xtile_stride[kMaxRanks];
};
#ifndef MIFGPU_FP32
__device__ __forceinline__ double *seg_address(const SegMap &m, int e, int outer, int x) {
  int r = 0;
  while (r + 1 < m.n && e >= m.lo[r + 1]) r++;
  return m.base[r] + (long long)(x >> 3) * m.xtile_stride[r] + (long long)outer * m.outer_stride[r] +
         (long long)(e - m.lo[r]) * m.estride[r] + (x & 7);
}

struct FastJob {
  SegMap load_map, store_map;  // strided (y / z) sweeps of the warp-per-line kernel only
  long long origin, lstride, estride, tile_stride, outer_stride;
  int n_tile_lines;
  int mode;  // 0 forward, 1 inverse (+normalise), 2 forward, eigenvalues, inverse (+normalise)
  const double2 *tw, *cs;
  const double *lam_x, *lam_y, *lam_z;
  double inv_norm;
  bool has_origin;  // the tile at (blockIdx.x, blockIdx.y) = (0, 0) contains the (0,0,0) mode
};

template <int LOGM, bool CONTIG>
__global__ void __launch_bounds__(1 << LOGM, (LOGM <= 9 ? 2 : 1)) fast_dct_kernel(const FastJob job, double *__restrict__ field) {
  using namespace fast;
  constexpr int M = 1 << LOGM, T = M / 8, THREADS = M, NPTS = M + 1;
  extern __shared__ double2 smem2[];
  double2 *S = smem2;
  double *Sd = reinterpret_cast<double *>(smem2);
  const int tid = threadIdx.x, line = tid & 7, j = tid >> 3;
  const int first_line = blockIdx.x * kLines;
  const int lines = min(kLines, job.n_tile_lines - first_line);
  double *base = field + job.origin + (long long)first_line * job.tile_stride + (long long)blockIdx.y * job.outer_stride;

  // load and pack the even extension (each value goes to its slot and to the slot of its mirror image)
  if (CONTIG) {
    for (int idx = tid; idx < kLines * NPTS; idx += THREADS) {
      const int l = idx / NPTS, e = idx - l * NPTS;
      const double value = (l < lines) ? base[(long long)l * job.lstride + e] : 0.0;
      put_packed(Sd, M, e, l, value);
    }
  } else {
    for (int e = j; e < NPTS; e += T) {
      const double value = (line < lines) ? base[(long long)line * job.lstride + (long long)e * job.estride] : 0.0;
      put_packed(Sd, M, e, line, value);
    }
  }
  __syncthreads();
  fft_lines<LOGM>(S, line, j, job.tw);
  double lo[4], hi[4], mid;
  dct_unpack<LOGM>(S, line, j, job.cs, lo, hi, mid);

  if (job.mode == 2) {
    // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z); mode (0,0,0) := 0 (src/PressureEquation.cpp:158-163)
    const int ix = min(first_line + line, job.n_tile_lines - 1);
    const double lam_xy = job.lam_x[ix] + job.lam_y[blockIdx.y];
    const bool origin_line = job.has_origin && (first_line + line == 0) && (blockIdx.y == 0);
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int k = j + T * s;
      const double scale_lo = (origin_line && k == 0) ? 0.0 : 1.0 / (lam_xy + job.lam_z[k]);
      lo[s] *= scale_lo;
      hi[s] *= 1.0 / (lam_xy + job.lam_z[M - k]);
    }
    mid *= 1.0 / (lam_xy + job.lam_z[M / 2]);
    __syncthreads();  // everyone has finished reading the spectrum
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int k = j + T * s;
      put_packed(Sd, M, k, line, lo[s]);
      put_packed(Sd, M, M - k, line, hi[s]);
    }
    if (j == 0) put_packed(Sd, M, M / 2, line, mid);
    __syncthreads();
    fft_lines<LOGM>(S, line, j, job.tw);
    dct_unpack<LOGM>(S, line, j, job.cs, lo, hi, mid);
  }
  const double scale = (job.mode == 0) ? 1.0 : job.inv_norm;

  if (!CONTIG) {
    if (line < lines) {
      double *out = base + (long long)line * job.lstride;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const int k = j + T * s;
        out[(long long)k * job.estride] = lo[s] * scale;
        out[(long long)(M - k) * job.estride] = hi[s] * scale;
      }
      if (j == 0) out[(long long)(M / 2) * job.estride] = mid * scale;
    }
  } else {
    // transpose through shared memory so that the global stores run along the contiguous x lines
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int k = j + T * s;
      Sd[k * kSlotPitch + line] = lo[s] * scale;
      Sd[(M - k) * kSlotPitch + line] = hi[s] * scale;
    }
    if (j == 0) Sd[(M / 2) * kSlotPitch + line] = mid * scale;
    __syncthreads();
    for (int idx = tid; idx < kLines * NPTS; idx += THREADS) {
      const int l = idx / NPTS, e = idx - l * NPTS;
      if (l < lines) base[(long long)l * job.lstride + e] = Sd[e * kSlotPitch + l];
    }
  }
}

template <int LOGM>
void launch_fast(cudaStream_t stream, SmemAttrOnce &attrs, const FastJob &job, bool contig, dim3 grid, double *field) {
  constexpr int M = 1 << LOGM;
  const size_t smem = (size_t)M * fast::kSlotPitch * sizeof(double2);
  attrs.ensure(fast_dct_kernel<LOGM, true>, smem);
  attrs.ensure(fast_dct_kernel<LOGM, false>, smem);
  if (contig) fast_dct_kernel<LOGM, true><<<grid, M, smem, stream>>>(job, field);
  else fast_dct_kernel<LOGM, false><<<grid, M, smem, stream>>>(job, field);
}


// ------------------------------------------------------------------------------------------------
// Warp-per-line path: DCT-I sweeps with M = 2^LOGM, 256 <= M <= 1024 (see mif_fft_warp.cuh).
// ------------------------------------------------------------------------------------------------
// SEG (strided sweeps only): 0 = launches that use a segment map (the LSU variant of the multi-GPU fused transposes); 1 =
// launches that use neither map: an instantiation without the map code -- no per-element map test, no indexed constant
// loads of the map arrays, a third fewer instructions.
template <int LOGM, bool CONTIG, int SEG = 0>
__global__ void __launch_bounds__(warpfft::Cfg<LOGM>::THREADS, (LOGM >= 10 ? 1 : 2)) warp_dct_kernel(const FastJob job, double *__restrict__ field) {
  using namespace warpfft;
  using C = Cfg<LOGM>;
  using L = LastPass<LOGM>;
  constexpr int M = C::M, TL = C::TL, NPTS = M + 1, EPT = C::EPT, PAIRS = EPT / 2;
  constexpr bool SHUFFLE = (C::WPL == 1);  // whole line in one warp: unpack with shuffles, spectrum stays in registers
  extern __shared__ double2 smem2[];
  double2 *T = smem2 + C::LINES * C::LINE_PITCH;  // twiddle tables behind the line regions
  const int tid = threadIdx.x;
  const int line = tid / TL, j = tid - line * TL;  // FFT mapping: a line is one warp (two for M = 1024)
  double2 *S = smem2 + line * C::LINE_PITCH;
  double *Sd = reinterpret_cast<double *>(S);
  const int first_line = blockIdx.x * C::LINES;
  const int lines = min(C::LINES, job.n_tile_lines - first_line);
  double *base = field + job.origin + (long long)first_line * job.tile_stride + (long long)blockIdx.y * job.outer_stride;
  double2 v[EPT];

  load_twiddles<LOGM>(T, job.tw);
  if (CONTIG) {
    // x sweep: lane j loads its first-pass inputs c[j + s*TL] = (e[2q], e[2q+1]) straight from global memory;
    // slots q >= M/2 are the mirror images (x[2M-2q], x[2M-2q-1]).  No shared-memory staging, no CTA barrier.
    const double *src = base + (long long)line * job.lstride;
    const bool live = line < lines;
#pragma unroll
    for (int s = 0; s < EPT; s++) {
      const int q = j + s * TL;
      if (!live) v[s] = make_double2(0.0, 0.0);
      else if (s < EPT / 2) v[s] = *reinterpret_cast<const double2 *>(src + 2 * q);
      else v[s] = make_double2(src[2 * M - 2 * q], src[2 * M - 2 * q - 1]);
    }
    __syncthreads();  // twiddle tables are in place
  } else {
    // y / z sweep: line-fastest mapping (the 8 lines are 8 consecutive x, so every request is a set of 64-byte
    // segments).  Thread (l, b) loads the inputs of the first-pass butterflies jj = b + u*TL of line l straight from
    // global memory -- slots jj + (M/8) t, i.e. elements (2q, 2q+1) or their mirror images (2M-2q, 2M-2q-1) -- runs
    // that pass in registers and stores its output into line l's region: the packed even extension never exists
    // in shared memory.
    const int l = tid % C::LINES, b = tid / C::LINES;
    const bool live = l < lines;
    const double *src = base + (long long)l * job.lstride;
    auto element = [&](int e) -> double {
      if (!live) return 0.0;
      if constexpr (SEG == 1) return src[(long long)e * job.estride];
      else return job.load_map.n ? *seg_address(job.load_map, e, blockIdx.y, first_line + l) : src[(long long)e * job.estride];
    };
    {
#pragma unroll
      for (int s = 0; s < EPT; s++) {
        const int q = b + s * TL;
        if (s < EPT / 2) v[s] = make_double2(element(2 * q), element(2 * q + 1));
        else v[s] = make_double2(element(2 * M - 2 * q), element(2 * M - 2 * q - 1));
      }
    }
    first_pass_in_place<LOGM>(smem2 + l * C::LINE_PITCH, b, v);
    __syncthreads();  // all lines and the twiddle tables are in place
  }

  double lo[PAIRS], hi[PAIRS], mid = 0.0, e_last = 0.0;
  double spec[EPT];  // shuffle path: spec[u + G t] = E_k, k = j + 32 u + NS t
  if (!CONTIG) fft_line<LOGM, false, SHUFFLE, true>(S, T, j, line, v);                     // first pass already done
  else fft_line<LOGM, true, SHUFFLE>(S, T, j, line, v);  // first pass from registers
  if constexpr (SHUFFLE) unpack_regs<LOGM>(v, j, job.cs, spec, e_last);
  else unpack_line<LOGM>(S, j, job.cs, lo, hi, mid);

  if (job.mode == 2) {
    // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z); mode (0,0,0) := 0 (src/PressureEquation.cpp:158-163)
    const int ix = min(first_line + line, job.n_tile_lines - 1);
    const double lam_xy = job.lam_x[ix] + job.lam_y[blockIdx.y];
    const bool origin_line = job.has_origin && (first_line + line == 0) && (blockIdx.y == 0);
    if constexpr (SHUFFLE) {
#pragma unroll
      for (int u = 0; u < L::G; u++)
#pragma unroll
        for (int t = 0; t < L::R; t++) {
          const int k = j + 32 * u + L::NS * t;
          spec[u + L::G * t] *= (origin_line && k == 0) ? 0.0 : 1.0 / (lam_xy + job.lam_z[k]);
        }
      e_last *= 1.0 / (lam_xy + job.lam_z[M]);
      line_sync<C::WPL>(line);  // all lanes are done with the previous contents of the line region
#pragma unroll
      for (int u = 0; u < L::G; u++)
#pragma unroll
        for (int t = 0; t < L::R; t++) put_packed(Sd, M, j + 32 * u + L::NS * t, spec[u + L::G * t]);
      if (j == 0) put_packed(Sd, M, M, e_last);
      line_sync<C::WPL>(line);
      fft_line<LOGM, false, true>(S, T, j, line, v);
      unpack_regs<LOGM>(v, j, job.cs, spec, e_last);
    } else {
#pragma unroll
      for (int s = 0; s < PAIRS; s++) {
        const int k = j + TL * s;
        lo[s] *= (origin_line && k == 0) ? 0.0 : 1.0 / (lam_xy + job.lam_z[k]);
        hi[s] *= 1.0 / (lam_xy + job.lam_z[M - k]);
      }
      mid *= 1.0 / (lam_xy + job.lam_z[M / 2]);
      line_sync<C::WPL>(line);  // the whole line has been unpacked into registers
#pragma unroll
      for (int s = 0; s < PAIRS; s++) {
        const int k = j + TL * s;
        put_packed(Sd, M, k, lo[s]);
        put_packed(Sd, M, M - k, hi[s]);
      }
      if (j == 0) put_packed(Sd, M, M / 2, mid);
      line_sync<C::WPL>(line);
      fft_line<LOGM, false, false>(S, T, j, line, v);
      unpack_line<LOGM>(S, j, job.cs, lo, hi, mid);
    }
  }
  const double scale = (job.mode == 0) ? 1.0 : job.inv_norm;

  if constexpr (CONTIG && SHUFFLE) {
    // x sweep: for fixed (u, t) the 32 lanes hold 32 consecutive outputs -> coalesced stores from registers.
    if (line < lines) {
      double *out = base + (long long)line * job.lstride;
#pragma unroll
      for (int u = 0; u < L::G; u++)
#pragma unroll
        for (int t = 0; t < L::R; t++) out[j + 32 * u + L::NS * t] = spec[u + L::G * t] * scale;
      if (j == 0) out[M] = e_last * scale;
    }
    return;
  }
  // Results go back through the line's own region as plain reals R[e], then out with coalesced stores.
  line_sync<C::WPL>(line);
  if constexpr (SHUFFLE) {
#pragma unroll
    for (int u = 0; u < L::G; u++)
#pragma unroll
      for (int t = 0; t < L::R; t++) Sd[j + 32 * u + L::NS * t] = spec[u + L::G * t] * scale;
    if (j == 0) Sd[M] = e_last * scale;
  } else {
#pragma unroll
    for (int s = 0; s < PAIRS; s++) {
      const int k = j + TL * s;
      Sd[k] = lo[s] * scale;
      Sd[M - k] = hi[s] * scale;
    }
    if (j == 0) Sd[M / 2] = mid * scale;
  }
  if (CONTIG) {
    line_sync<C::WPL>(line);
    if (line < lines) {
      double *out = base + (long long)line * job.lstride;
      for (int e = j; e < NPTS; e += TL) out[e] = Sd[e];
    }
  } else {
    __syncthreads();
    const int l = tid % C::LINES, q0 = tid / C::LINES;
    constexpr int QSTEP = C::THREADS / C::LINES;
    if (l < lines) {
      const double *srcl = reinterpret_cast<const double *>(smem2 + l * C::LINE_PITCH);
      double *out = base + (long long)l * job.lstride;
      if constexpr (SEG == 1) {
        for (int e = q0; e < NPTS; e += QSTEP) out[(long long)e * job.estride] = srcl[e];
      } else if (job.store_map.n) {
        // Fused transpose: the results go straight into the pencil / slab buffers of the owning GPUs (peer stores
        // over NVLink for r != this rank), so no pack kernel and no separate all-to-all copy exist.
        for (int e = q0; e < NPTS; e += QSTEP) *seg_address(job.store_map, e, blockIdx.y, first_line + l) = srcl[e];
      } else {
        for (int e = q0; e < NPTS; e += QSTEP) out[(long long)e * job.estride] = srcl[e];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// x sweeps of 513-point DCT-I lines on the 16 x 32 transform (mif_fft512.cuh): one warp per line, inputs straight
// from global memory (coalesced along the line), ONE pass through the warp's shared-memory region, results stored
// from registers.  MODE 0: forward.  MODE 1: inverse + normalisation as the half-complex -> real transform of the
// real spectrum (see mif_poisson_tma.cuh): 513 values are read once, no unpack pass, the mirror half of the output
// is never formed.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256, 2) x_dct512_kernel(const FastJob job, double *__restrict__ field) {
  constexpr int M = 512;
  extern __shared__ double2 smem2[];
  double2 *T = smem2 + warpfft::kLines * fft512::kLinePitch;
  const int tid = threadIdx.x, line = tid >> 5, L = tid & 31;
  double2 *Sline = smem2 + line * fft512::kLinePitch;
  const int first_line = blockIdx.x * warpfft::kLines;
  const int lines = min(warpfft::kLines, job.n_tile_lines - first_line);
  fft512::load_twiddles<256>(T, job.tw);
  __syncthreads();
  if (line >= lines) return;  // whole warps: nothing below synchronises across warps
  double *row = field + job.origin + (long long)first_line * job.tile_stride + (long long)blockIdx.y * job.outer_stride +
                (long long)line * job.lstride;
  double2 v[16];
  if (MODE == 0) {
    // packed even extension c_q = (e(2q), e(2q+1)), q = L + 32 s; slots q >= M/2 are the mirror images
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const int q = L + 32 * s;
      if (s < 8) v[s] = *reinterpret_cast<const double2 *>(row + 2 * q);
      else v[s] = make_double2(row[2 * M - 2 * q], row[2 * M - 2 * q - 1]);
    }
  } else {
    double nat[16];
#pragma unroll
    for (int s = 0; s < 16; s++) nat[s] = row[L + 32 * s];
    const double e_last = row[M];  // same address in all lanes: one broadcast request
    fft512::pack_natural(nat, e_last, L, job.cs, v);
  }
  fft512::phase_a(v, L, T);
  fft512::store_a(Sline, L, v);
  __syncwarp();
  fft512::phase_b(Sline, L, v);
  if (MODE == 0) {
    double spec[16], e_last;
    fft512::unpack_dct(v, L, job.cs, spec, e_last);
#pragma unroll
    for (int r = 0; r < 16; r++) row[fft512::k_of(L, r)] = spec[r];  // 16 consecutive lanes = 128 contiguous bytes
    if (L == 0) row[M] = e_last;
  } else {
    // register r < 8 holds z_q = conj(v[r]), q = k2 + 128 p + 16 r: x(2q) = Re z_q, x(2q+1) = Im z_q
    const double scale = job.inv_norm;
    const int q0 = (L & 15) + 128 * (L >> 4);
#pragma unroll
    for (int r = 0; r < 8; r++)
      *reinterpret_cast<double2 *>(row + 2 * (q0 + 16 * r)) = make_double2(v[r].x * scale, -v[r].y * scale);
    if (L == 0) row[M] = v[8].x * scale;  // q = M/2
  }
}

void launch_x512(cudaStream_t stream, SmemAttrOnce &attrs, const FastJob &job, int mode, int outer, double *field) {
  const size_t smem = (size_t)(warpfft::kLines * fft512::kLinePitch + fft512::kTwiddles) * sizeof(double2);
  attrs.ensure(x_dct512_kernel<0>, smem);
  attrs.ensure(x_dct512_kernel<1>, smem);
  const dim3 grid((job.n_tile_lines + warpfft::kLines - 1) / warpfft::kLines, outer, 1);
  if (mode == 0) x_dct512_kernel<0><<<grid, 256, smem, stream>>>(job, field);
  else x_dct512_kernel<1><<<grid, 256, smem, stream>>>(job, field);
}

// ------------------------------------------------------------------------------------------------
// Periodic directions: FFTW_R2HC / FFTW_HC2R on n = 2M real points (M = 256 or 512) with the same warp-per-line
// machinery -- the line is packed two reals per complex, transformed by ONE length-M complex FFT and unpacked to
// FFTW's halfcomplex order r_0 .. r_{n/2}, i_{n/2-1} .. i_1 with shuffles.  The fused z sweep never leaves the
// registers: unpack, eigenvalue scaling in halfcomplex order, repack (the unpacked index k = j + 32 (u + G t) is
// the first-pass slot of the next transform), inverse FFT as conj(FFT(conj .)), normalisation.
// ------------------------------------------------------------------------------------------------
// MODE as a template parameter: the three paths keep different arrays alive (the 1024-point instantiation spilled 400-600
// bytes while the mode was a run-time value).
template <int LOGM, bool CONTIG, int MODE>
__global__ void __launch_bounds__(warpfft::Cfg<LOGM>::THREADS, 2) warp_rfft_kernel(const FastJob job, double *__restrict__ field) {
  using namespace warpfft;
  using C = Cfg<LOGM>;
  using L = LastPass<LOGM>;
  constexpr int M = C::M, n = 2 * M, TL = C::TL, EPT = C::EPT;
  static_assert(C::WPL == 1, "one warp per line");
  extern __shared__ double2 smem2[];
  double2 *T = smem2 + kLines * C::LINE_PITCH;
  const int tid = threadIdx.x;
  const int line = tid / TL, j = tid - line * TL;
  double2 *S = smem2 + line * C::LINE_PITCH;
  const int first_line = blockIdx.x * kLines;
  const int lines = min(kLines, job.n_tile_lines - first_line);
  double *base = field + job.origin + (long long)first_line * job.tile_stride + (long long)blockIdx.y * job.outer_stride;
  double2 v[EPT];

  load_twiddles<LOGM>(T, job.tw);
  // loader thread: (line, j) along the line for x sweeps, line-fastest (l, b) for the strided sweeps
  const int ll = CONTIG ? line : (tid & 7), lb = CONTIG ? j : (tid >> 3);
  const bool live = ll < lines;
  const double *src = base + (long long)ll * job.lstride;
  auto element = [&](int e) -> double { return live ? src[(long long)e * job.estride] : 0.0; };
  if (MODE == 1) {
    // halfcomplex input: X_k = (in[k], in[n-k]), partner X_{M-k} = (in[M-k], in[M+k]); X_0 and X_M are real
#pragma unroll
    for (int s = 0; s < EPT; s++) {
      const int k = lb + 32 * s;
      const double xr = element(k), yr = element(M - k);
      const double xi = (k == 0) ? 0.0 : element(n - k), yi = (k == 0) ? 0.0 : element(M + k);
      const double2 w = __ldg(&job.cs[k]);
      v[s] = hc2r_input(xr, xi, yr, yi, w.x, w.y);
    }
  } else {
#pragma unroll
    for (int s = 0; s < EPT; s++) {
      const int q = lb + 32 * s;
      v[s] = make_double2(element(2 * q), element(2 * q + 1));
    }
  }
  if (CONTIG) {
    __syncthreads();  // twiddle tables are in place
    fft_line<LOGM, true, true>(S, T, j, line, v);
  } else {
    first_pass_in_place<LOGM>(smem2 + ll * C::LINE_PITCH, lb, v);
    __syncthreads();
    fft_line<LOGM, false, true, true>(S, T, j, line, v);
  }

  double re[EPT], im[EPT], x_last = 0.0;
  if (MODE != 1) {
    unpack_r2hc_regs<LOGM>(v, j, job.cs, re, im, x_last);
    if (MODE == 2) {
      // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z) in FFTW output order; mode (0,0,0) := 0
      // (src/PressureEquation.cpp:158-163)
      const int ix = min(first_line + line, job.n_tile_lines - 1);
      const double lam_xy = job.lam_x[ix] + job.lam_y[blockIdx.y];
      const bool origin_line = job.has_origin && (first_line + line == 0) && (blockIdx.y == 0);
#pragma unroll
      for (int u = 0; u < L::G; u++)
#pragma unroll
        for (int t = 0; t < L::R; t++) {
          const int k = j + 32 * u + L::NS * t;
          re[u + L::G * t] *= (origin_line && k == 0) ? 0.0 : 1.0 / (lam_xy + job.lam_z[k]);
          if (k > 0) im[u + L::G * t] *= 1.0 / (lam_xy + job.lam_z[n - k]);
        }
      x_last *= 1.0 / (lam_xy + job.lam_z[M]);
      pack_hc2r_regs<LOGM>(re, im, x_last, j, job.cs, v);
      __syncwarp();
      fft_line<LOGM, true, true>(S, T, j, line, v);
    }
  }

  if (MODE == 0) {
    // halfcomplex output
    if (CONTIG) {
      if (line < lines) {
        double *out = base + (long long)line * job.lstride;
#pragma unroll
        for (int u = 0; u < L::G; u++)
#pragma unroll
          for (int t = 0; t < L::R; t++) {
            const int k = j + 32 * u + L::NS * t;
            out[k] = re[u + L::G * t];
            if (k > 0) out[n - k] = im[u + L::G * t];
          }
        if (j == 0) out[M] = x_last;
      }
      return;
    }
    double *Sd = reinterpret_cast<double *>(S);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < L::G; u++)
#pragma unroll
      for (int t = 0; t < L::R; t++) {
        const int k = j + 32 * u + L::NS * t;
        Sd[k] = re[u + L::G * t];
        if (k > 0) Sd[n - k] = im[u + L::G * t];
      }
    if (j == 0) Sd[M] = x_last;
  } else {
    // x_{2k} = Re z_k, x_{2k+1} = Im z_k with z = conj(FFT(conj Z)); normalised (src/PressureEquation.cpp:167-195)
    const double scale = job.inv_norm;
    if (CONTIG) {
      if (line < lines) {
        double *out = base + (long long)line * job.lstride;
#pragma unroll
        for (int s = 0; s < EPT; s++) {
          const int k = j + 32 * s;
          out[2 * k] = v[s].x * scale;
          out[2 * k + 1] = -v[s].y * scale;
        }
      }
      return;
    }
    __syncwarp();
#pragma unroll
    for (int s = 0; s < EPT; s++) S[j + 32 * s] = make_double2(v[s].x * scale, -v[s].y * scale);
  }
  // strided sweeps: the line's region now holds the n output reals; store them line-fastest
  __syncthreads();
  if (live) {
    const double *srcl = reinterpret_cast<const double *>(smem2 + ll * C::LINE_PITCH);
    double *out = base + (long long)ll * job.lstride;
    for (int e = lb; e < n; e += C::THREADS / 8) out[(long long)e * job.estride] = srcl[e];
  }
}

template <int LOGM, int MODE>
void launch_rfft_mode(cudaStream_t stream, SmemAttrOnce &attrs, const FastJob &job, bool contig, dim3 grid, double *field) {
  using C = warpfft::Cfg<LOGM>;
  if (contig) {
    attrs.ensure(warp_rfft_kernel<LOGM, true, MODE>, C::SMEM);
    warp_rfft_kernel<LOGM, true, MODE><<<grid, C::THREADS, C::SMEM, stream>>>(job, field);
  } else {
    attrs.ensure(warp_rfft_kernel<LOGM, false, MODE>, C::SMEM);
    warp_rfft_kernel<LOGM, false, MODE><<<grid, C::THREADS, C::SMEM, stream>>>(job, field);
  }
}
template <int LOGM>
void launch_rfft(cudaStream_t stream, SmemAttrOnce &attrs, const FastJob &job, bool contig, dim3 grid, double *field) {
  if (job.mode == 0) launch_rfft_mode<LOGM, 0>(stream, attrs, job, contig, grid, field);
  else if (job.mode == 1) launch_rfft_mode<LOGM, 1>(stream, attrs, job, contig, grid, field);
  else launch_rfft_mode<LOGM, 2>(stream, attrs, job, contig, grid, field);
}

// ------------------------------------------------------------------------------------------------
// Split path for lines of 1025 points (M = 1024): the length-1024 complex FFT is split by one radix-2
// decimation-in-frequency step into two independent length-512 transforms,
//     a_q = c_q + c_{q+512}  ->  C_{2k},        b_q = (c_q - c_{q+512}) W_1024^q  ->  C_{2k+1},      q, k < 512,
// and each of them is run by ONE warp like a 513-point line of warp_dct_kernel<9>: radix-8 passes through the
// warp's own shared-memory region with __syncwarp() only, spectrum kept in registers, DCT-I unpack by shuffles
// (even half: the pairs of the short transform; odd half: C_{2k+1} pairs with C_{2(511-k)+1}).
// Loading: thread jj (0..63) of a line owns first-pass butterfly jj of BOTH halves -- it loads the pairs
// (c_q, c_{q+512}), q = jj + 64 t, t < 8, exactly once, forms a_q and b_q, runs the two radix-8 butterflies in
// registers and stores them into the two regions.  For x sweeps thread jj is lane jj & 31 of the line's warp
// jj >> 5 (coalesced along the line); for the strided sweeps it is thread (l, jj) of the line-fastest mapping.
// The two warps of a line meet at a 64-thread named barrier after that first pass and where the 1025 results
// are gathered (eigenvalue step of the fused z sweep, coalesced store).  The previous M = 1024 variant (two warps
// sharing every pass through bar.sync and a shared-memory unpack, one 512-thread CTA per SM) stays available with
// MIFGPU_FFT_NO_SPLIT=1.
// ------------------------------------------------------------------------------------------------
namespace split {
constexpr int kFull = 1024;
using C9 = warpfft::Cfg<9>;
constexpr size_t smem_bytes(int lines) { return (size_t)(2 * lines * C9::LINE_PITCH + C9::TW_TOTAL) * sizeof(double2); }
// exp(-2 pi i t / 16) = W_1024^(64 t), t < 8
__device__ __forceinline__ double2 rot16(int t) {
  constexpr double kCos[8] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173,
                              0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128675613};
  constexpr double kSin[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128675613,
                              1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};
  return make_double2(kCos[t], -kSin[t]);
}
// 64-thread named barrier of one line's two warps.  Immediate barrier numbers: with a register operand ptxas
// reserves all 16 hardware barriers for the CTA, which limits the CTAs per SM.
__device__ __forceinline__ void pair_sync(int line) {
  switch (line) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    case 3: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
    case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
    case 5: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
    case 6: asm volatile("bar.sync 7, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 8, 64;" ::: "memory"); break;
  }
}
// First pass of both halves for butterfly jj: a[t], b[t] hold the pairs' sums / differences for q = jj + 64 t on
// entry (differences not yet multiplied by W_1024^q); results go to slots 8 jj + t of the two regions.
__device__ __forceinline__ void first_pass_both(double2 *even_region, double2 *odd_region, int jj, double2 wjj,
                                                double2 *a, double2 *b) {
#pragma unroll
  for (int t = 0; t < 8; t++) b[t] = warpfft::cmul(b[t], warpfft::cmul(wjj, rot16(t)));
  fast::dft8(a);
  fast::dft8(b);
#pragma unroll
  for (int t = 0; t < 8; t++) {
    even_region[warpfft::pad(jj * 8 + t)] = a[t];
    odd_region[warpfft::pad(jj * 8 + t)] = b[t];
  }
}
}  // namespace split

// LINES lines per CTA (64 threads each): 8 lines fill the register file with ONE CTA per SM, whose 16 warps then load,
// transform and store in lockstep; 4 lines give two independent CTAs per SM that overlap each other's phases.
// SEG as in warp_dct_kernel: 0 = segment maps in use, 1 = instantiation without the map code
template <bool CONTIG, int LINES, int SEG = 0>
__global__ void __launch_bounds__(64 * LINES, 8 / LINES) warp_dct_split_kernel(const FastJob job, double *__restrict__ field) {
  using namespace warpfft;
  using namespace split;
  using L = LastPass<9>;
  constexpr int EPT = C9::EPT, NPTS = kFull + 1, kThreads = 64 * LINES, kLines = LINES;
  extern __shared__ double2 smem2[];
  double2 *T = smem2 + 2 * kLines * C9::LINE_PITCH;
  const int tid = threadIdx.x, warp = tid >> 5, j = tid & 31;
  const int line = warp >> 1;
  const bool odd = warp & 1;
  double2 *S = smem2 + warp * C9::LINE_PITCH;                                        // this warp's FFT region
  double *Rl = reinterpret_cast<double *>(smem2 + (2 * line) * C9::LINE_PITCH);      // the line's 1025 reals (aliases S of the even warp)
  const int first_line = blockIdx.x * kLines;
  const int lines = min(kLines, job.n_tile_lines - first_line);
  double *base = field + job.origin + (long long)first_line * job.tile_stride + (long long)blockIdx.y * job.outer_stride;
  double2 v[EPT];

  load_twiddles<9, 2, kThreads>(T, job.tw);
  {
    // loader thread: butterfly jj of line ll
    const int ll = CONTIG ? line : tid % LINES, jj = CONTIG ? (odd ? 32 : 0) + j : tid / LINES;
    const bool live = ll < lines;
    const double *src = base + (long long)ll * job.lstride;
    auto element = [&](int e) -> double {
      if (!live) return 0.0;
      if (CONTIG) return src[e];
      if constexpr (SEG == 1) return src[(long long)e * job.estride];
      else return job.load_map.n ? *seg_address(job.load_map, e, blockIdx.y, first_line + ll) : src[(long long)e * job.estride];
    };
    double2 *a = v, *b = v + 8;
    {
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const int q = jj + 64 * t;
        double2 lo;
        if (CONTIG) lo = live ? *reinterpret_cast<const double2 *>(src + 2 * q) : make_double2(0.0, 0.0);
        else lo = make_double2(element(2 * q), element(2 * q + 1));
        const double2 hi = make_double2(element(kFull - 2 * q), element(kFull - 1 - 2 * q));
        a[t] = cadd(lo, hi);
        b[t] = csub(lo, hi);
      }
    }
    first_pass_both(smem2 + (2 * ll) * C9::LINE_PITCH, smem2 + (2 * ll + 1) * C9::LINE_PITCH, jj, __ldg(&job.tw[jj]), a, b);
  }
  __syncthreads();  // first passes of all lines and the twiddle tables are in place
  fft_line<9, false, true, true>(S, T, j, line, v);

  double spec[EPT], e_last = 0.0;  // spec[u + G t] = E_(2k + odd), k = j + 32 u + 64 t
  if (odd) unpack_regs<9, true, true>(v, j, job.cs, spec, e_last);
  else unpack_regs<9, true, false>(v, j, job.cs, spec, e_last);

  if (job.mode == 2) {
    // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z); mode (0,0,0) := 0 (src/PressureEquation.cpp:158-163)
    const int ix = min(first_line + line, job.n_tile_lines - 1);
    const double lam_xy = job.lam_x[ix] + job.lam_y[blockIdx.y];
    const bool origin_line = job.has_origin && (first_line + line == 0) && (blockIdx.y == 0);
#pragma unroll
    for (int u = 0; u < L::G; u++)
#pragma unroll
      for (int t = 0; t < L::R; t++) {
        const int k = 2 * (j + 32 * u + L::NS * t) + (odd ? 1 : 0);
        spec[u + L::G * t] *= (origin_line && k == 0) ? 0.0 : 1.0 / (lam_xy + job.lam_z[k]);
      }
    e_last *= 1.0 / (lam_xy + job.lam_z[kFull]);
    pair_sync(line);  // both warps are done with their regions
#pragma unroll
    for (int u = 0; u < L::G; u++)
#pragma unroll
      for (int t = 0; t < L::R; t++) Rl[2 * (j + 32 * u + L::NS * t) + (odd ? 1 : 0)] = spec[u + L::G * t];
    if (!odd && j == 0) Rl[kFull] = e_last;
    pair_sync(line);
    {
      const int jj = (odd ? 32 : 0) + j;
      double2 *a = v, *b = v + 8;
#pragma unroll
      for (int t = 0; t < 8; t++) {
        const int q = jj + 64 * t;
        const double2 lo = *reinterpret_cast<const double2 *>(Rl + 2 * q);
        const double2 hi = make_double2(Rl[kFull - 2 * q], Rl[kFull - 1 - 2 * q]);
        a[t] = cadd(lo, hi);
        b[t] = csub(lo, hi);
      }
      pair_sync(line);  // the whole line has been read before the regions (which alias it) are overwritten
      first_pass_both(smem2 + (2 * line) * C9::LINE_PITCH, smem2 + (2 * line + 1) * C9::LINE_PITCH, jj, __ldg(&job.tw[jj]), a, b);
    }
    pair_sync(line);
    fft_line<9, false, true, true>(S, T, j, line, v);
    if (odd) unpack_regs<9, true, true>(v, j, job.cs, spec, e_last);
    else unpack_regs<9, true, false>(v, j, job.cs, spec, e_last);
  }
  const double scale = (job.mode == 0) ? 1.0 : job.inv_norm;

  // gather the line as plain reals, then coalesced stores
  pair_sync(line);
#pragma unroll
  for (int u = 0; u < L::G; u++)
#pragma unroll
    for (int t = 0; t < L::R; t++) Rl[2 * (j + 32 * u + L::NS * t) + (odd ? 1 : 0)] = spec[u + L::G * t] * scale;
  if (!odd && j == 0) Rl[kFull] = e_last * scale;
  if (CONTIG) {
    pair_sync(line);
    if (line < lines) {
      double *out = base + (long long)line * job.lstride;
      for (int e = (odd ? 32 : 0) + j; e < NPTS; e += 64) out[e] = Rl[e];
    }
  } else {
    __syncthreads();
    const int l = tid % LINES, q0 = tid / LINES;
    if (l < lines) {
      const double *srcl = reinterpret_cast<const double *>(smem2 + (2 * l) * C9::LINE_PITCH);
      double *out = base + (long long)l * job.lstride;
      if constexpr (SEG == 1) {
        for (int e = q0; e < NPTS; e += 64) out[(long long)e * job.estride] = srcl[e];
      } else if (job.store_map.n) {
        for (int e = q0; e < NPTS; e += 64) *seg_address(job.store_map, e, blockIdx.y, first_line + l) = srcl[e];
      } else {
        for (int e = q0; e < NPTS; e += 64) out[(long long)e * job.estride] = srcl[e];
      }
    }
  }
}

// x sweeps of 1025-point lines: the radix-2 split of mif_poisson_tma.cuh (tma_dct1024_kernel) with the inputs straight
// from global memory -- two warps per line (even / odd half of the 1024-point transform), 4 lines per CTA.  Both warps read
// the whole line and each writes its half of the results, so a 64-thread barrier separates a line's loads from its stores.
template <int MODE>
__global__ void __launch_bounds__(256, 2) x_dct1024_kernel(const FastJob job, double *__restrict__ field) {
  constexpr int M = 1024, H = 512, kLinesX = 4;
  extern __shared__ double2 smem2[];
  double2 *T = smem2 + 2 * kLinesX * fft512::kLinePitch;
  const int tid = threadIdx.x, warp = tid >> 5, L = tid & 31;
  const int line = warp >> 1, half = warp & 1;
  double2 *Sline = smem2 + warp * fft512::kLinePitch;
  const int first_line = blockIdx.x * kLinesX;
  const int lines = min(kLinesX, job.n_tile_lines - first_line);
  fft512::load_twiddles<256, 2>(T, job.tw);
  __syncthreads();
  if (line >= lines) return;  // both warps of a line leave together; nothing below synchronises across lines
  double *row = field + job.origin + (long long)first_line * job.tile_stride + (long long)blockIdx.y * job.outer_stride +
                (long long)line * job.lstride;
  const double2 wL = __ldg(&job.tw[L]);  // W_1024^L;  W_1024^(L + 32 s) = W_1024^L W_32^s
  double2 v[16];
  if (MODE == 0) {
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const int q = L + 32 * s;
      const double2 lo = *reinterpret_cast<const double2 *>(row + 2 * q);          // c_q = (e(2q), e(2q+1))
      const double2 hi = make_double2(row[M - 2 * q], row[M - 2 * q - 1]);         // c_{q+512}: mirror image
      v[s] = tmasweep::split_input(lo, hi, half, fft512::cmul(wL, fft512::w32(s)));
    }
  } else {
    const double2 cs_L = __ldg(&job.cs[L]);
#pragma unroll
    for (int s = 0; s < 16; s++) {
      const int k = L + 32 * s;
      const double2 rt = fft512::rot32x(s);  // cs[k] = cs[L] rotated by pi s / 32, cs[k + 512] = cs[k] rotated by pi / 2
      const double c = cs_L.x * rt.x - cs_L.y * rt.y, sn = cs_L.x * rt.y + cs_L.y * rt.x;
      const double2 lo = fft512::pack_input(row[k], row[M - k], c, sn);
      const double2 hi = fft512::pack_input(row[k + H], row[H - k], -sn, c);
      v[s] = tmasweep::split_input(lo, hi, half, fft512::cmul(wL, fft512::w32(s)));
    }
  }
  fft512::phase_a(v, L, T);
  fft512::store_a(Sline, L, v);
  __syncwarp();
  fft512::phase_b(Sline, L, v);
  if (MODE == 0) {
    double spec[16], e_last;
    if (half) fft512::unpack_dct<2>(v, L, job.cs, spec, e_last);
    else fft512::unpack_dct<1>(v, L, job.cs, spec, e_last);
    split::pair_sync(line);  // both warps have read the line
#pragma unroll
    for (int r = 0; r < 16; r++) row[2 * fft512::k_of(L, r) + half] = spec[r];
    if (half == 0 && L == 0) row[M] = e_last;
  } else {
    // register r < 8 holds z_q = conj(v[r]), q = 2 (k2 + 128 p + 16 r) + half: x(2q) = Re z_q, x(2q+1) = Im z_q
    const double scale = job.inv_norm;
    const int q0 = 2 * ((L & 15) + 128 * (L >> 4)) + half;
    split::pair_sync(line);
#pragma unroll
    for (int r = 0; r < 8; r++)
      *reinterpret_cast<double2 *>(row + 2 * (q0 + 32 * r)) = make_double2(v[r].x * scale, -v[r].y * scale);
    if (half == 0 && L == 0) row[M] = v[8].x * scale;  // q = M/2
  }
}

void launch_x1024(cudaStream_t stream, SmemAttrOnce &attrs, const FastJob &job, int mode, int outer, double *field) {
  const size_t smem = (size_t)(8 * fft512::kLinePitch + fft512::kTwiddles) * sizeof(double2);
  attrs.ensure(x_dct1024_kernel<0>, smem);
  attrs.ensure(x_dct1024_kernel<1>, smem);
  const dim3 grid((job.n_tile_lines + 3) / 4, outer, 1);
  if (mode == 0) x_dct1024_kernel<0><<<grid, 256, smem, stream>>>(job, field);
  else x_dct1024_kernel<1><<<grid, 256, smem, stream>>>(job, field);
}

template <int LINES>
void launch_split(cudaStream_t stream, SmemAttrOnce &attrs, const FastJob &job, bool contig, int outer, double *field) {
  const size_t smem = split::smem_bytes(LINES);
  const dim3 grid((job.n_tile_lines + LINES - 1) / LINES, outer, 1);
  if (contig) {
    attrs.ensure(warp_dct_split_kernel<true, LINES>, smem);
    warp_dct_split_kernel<true, LINES><<<grid, 64 * LINES, smem, stream>>>(job, field);
  } else if (job.load_map.n == 0 && job.store_map.n == 0) {
    attrs.ensure(warp_dct_split_kernel<false, LINES, 1>, smem);
    warp_dct_split_kernel<false, LINES, 1><<<grid, 64 * LINES, smem, stream>>>(job, field);
  } else {
    attrs.ensure(warp_dct_split_kernel<false, LINES>, smem);
    warp_dct_split_kernel<false, LINES><<<grid, 64 * LINES, smem, stream>>>(job, field);
  }
}

template <int LOGM>
void launch_warp(cudaStream_t stream, SmemAttrOnce &attrs, const FastJob &job, bool contig, dim3 grid, double *field) {
  using C = warpfft::Cfg<LOGM>;
  grid.x = (job.n_tile_lines + C::LINES - 1) / C::LINES;
  if (contig) {
    attrs.ensure(warp_dct_kernel<LOGM, true>, C::SMEM);
    warp_dct_kernel<LOGM, true><<<grid, C::THREADS, C::SMEM, stream>>>(job, field);
  } else if (job.load_map.n == 0 && job.store_map.n == 0) {
    // strided sweeps without the segment-map code when no map is in use (profiles/r01_ab_plain_strided.json)
    attrs.ensure(warp_dct_kernel<LOGM, false, 1>, C::SMEM);
    warp_dct_kernel<LOGM, false, 1><<<grid, C::THREADS, C::SMEM, stream>>>(job, field);
  } else {
    attrs.ensure(warp_dct_kernel<LOGM, false>, C::SMEM);
    warp_dct_kernel<LOGM, false><<<grid, C::THREADS, C::SMEM, stream>>>(job, field);
  }
}

#endif  // !MIFGPU_FP32

template <typename T>
T *to_device(const std::vector<T> &host) {
  if (host.empty()) return nullptr;
  T *dev = nullptr;
  cudaMalloc(&dev, host.size() * sizeof(T));
  cudaMemcpy(dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice);
  return dev;
}
// The plan tables are computed in double and stored in the build's scalar type.
#ifdef MIFGPU_FP32
real2 *to_device(const std::vector<double2> &host) {
  std::vector<real2> narrow(host.size());
  for (size_t i = 0; i < host.size(); i++) narrow[i] = make_real2((real)host[i].x, (real)host[i].y);
  return to_device(narrow);
}
real *to_device(const std::vector<double> &host) {
  std::vector<real> narrow(host.begin(), host.end());
  return to_device(narrow);
}
#endif

// exp(-2 pi i num / den) with the argument reduced in integers.
double2 unit_root(long long num, long long den) {
  num %= den;
  const double a = -2.0 * kPi * (double)num / (double)den;
  return make_double2(std::cos(a), std::sin(a));
}

void host_fft_pow2(std::vector<double2> &z, bool inverse) {  // small recursive helper for plan tables
  const size_t n = z.size();
  if (n <= 1) return;
  std::vector<double2> even(n / 2), odd(n / 2);
  for (size_t i = 0; i < n / 2; i++) {
    even[i] = z[2 * i];
    odd[i] = z[2 * i + 1];
  }
  host_fft_pow2(even, inverse);
  host_fft_pow2(odd, inverse);
  for (size_t k = 0; k < n / 2; k++) {
    double2 t = unit_root((long long)k, (long long)n);
    if (inverse) t.y = -t.y;
    const double2 o = make_double2(odd[k].x * t.x - odd[k].y * t.y, odd[k].x * t.y + odd[k].y * t.x);
    z[k] = make_double2(even[k].x + o.x, even[k].y + o.y);
    z[k + n / 2] = make_double2(even[k].x - o.x, even[k].y - o.y);
  }
}

}  // namespace

struct PoissonPlan {
  DirPlanDev dir[3];
  int fast_logm[3];  // > 0: DCT-I with N - 1 = 2^fast_logm handled by fast_dct_kernel
  int rfft_logm[3];  // > 0: periodic direction with n / 2 = 2^rfft_logm handled by warp_rfft_kernel
  int L[3];
  size_t smem[3];
  std::vector<void *> allocations;
#ifndef MIFGPU_FP32
  tmasweep::Cache tma;  // tensor maps of the TMA-staged strided sweeps, per field / direction
#endif
  SmemAttrOnce attrs;   // dynamic shared-memory opt-in per kernel, for this plan's device
};

PoissonPlan *poisson_plan_create(const Geom &g, const int n_points[3], const int periodic[3], const double h[3],
                                 const int n_global[3]) {
  (void)g;
  PoissonPlan *plan = new PoissonPlan();
  for (int d = 0; d < 3; d++) {
    DirPlanDev &pl = plan->dir[d];
    const int n = n_points[d];
    pl.periodic = periodic[d];
    pl.n = n;
    pl.m = !periodic[d] ? n - 1 : ((n % 2 == 0) ? n / 2 : n);
    if (pl.m < 1) pl.m = 1;
    const bool pow2 = (pl.m & (pl.m - 1)) == 0;
    pl.bluestein = !pow2;
    int P = 1;
    if (pow2) P = pl.m;
    else while (P < 2 * pl.m - 1) P <<= 1;
    pl.P = P;
    pl.logP = 0;
    while ((1 << pl.logP) < P) pl.logP++;

    std::vector<double2> tw(std::max(P / 2, 1));
    for (int q = 0; q < P / 2; q++) tw[q] = unit_root(q, P);
    pl.tw = to_device(tw);
    plan->allocations.push_back((void *)pl.tw);
    std::vector<double2> tw_full(P);
    for (int q = 0; q < P; q++) tw_full[q] = unit_root(q, P);
    pl.tw_full = to_device(tw_full);
    plan->allocations.push_back((void *)pl.tw_full);

    pl.chirp = nullptr;
    pl.filt = nullptr;
    if (pl.bluestein) {
      const int m = pl.m;
      std::vector<double2> chirp(m), filt(P, make_double2(0.0, 0.0));
      for (int j = 0; j < m; j++) chirp[j] = unit_root(((long long)j * j) % (2LL * m), 2LL * m);
      for (int j = 0; j < m; j++) {
        filt[j] = make_double2(chirp[j].x, -chirp[j].y);
        if (j > 0) filt[P - j] = filt[j];
      }
      host_fft_pow2(filt, false);
      std::vector<double2> filt_br(P);
      for (int i = 0; i < P; i++) {
        unsigned r = 0;
        for (int b = 0; b < pl.logP; b++)
          if (i & (1 << b)) r |= 1u << (pl.logP - 1 - b);
        filt_br[i] = make_double2(filt[r].x / P, filt[r].y / P);
      }
      pl.chirp = to_device(chirp);
      pl.filt = to_device(filt_br);
      plan->allocations.push_back((void *)pl.chirp);
      plan->allocations.push_back((void *)pl.filt);
    }

    // unpack twiddles: DCT-I exp(-i pi k / m), k <= m; real FFT exp(-2 pi i k / n), k <= n/2 (stored as cos, sin)
    std::vector<double2> unpack;
    if (!periodic[d]) {
      unpack.resize(n);
      for (int k = 0; k < n; k++) {
        const double2 r = unit_root(k, 2LL * std::max(pl.m, 1));
        unpack[k] = make_double2(r.x, -r.y);
      }
    } else {
      unpack.resize(n / 2 + 1);
      for (int k = 0; k <= n / 2; k++) {
        const double2 r = unit_root(k, n);
        unpack[k] = make_double2(r.x, -r.y);
      }
    }
    pl.unpack = to_device(unpack);
    plan->allocations.push_back((void *)pl.unpack);

    // eigenvalues, formulas and evaluation order of src/PressureSolverStructures.cpp:5-11
    std::vector<double> lambda(n);
    for (int k = 0; k < n; k++) {
      if (periodic[d]) lambda[k] = 2.0 * (std::cos(2.0 * M_PI * k / n) - 1.0) / (h[d] * h[d]);
      else lambda[k] = 2.0 * (std::cos(M_PI * k / (n - 1)) - 1.0) / (h[d] * h[d]);
    }
    pl.lambda = to_device(lambda);
    plan->allocations.push_back((void *)pl.lambda);
    // src/PressureEquation.cpp:167,200,234: N_domains_global * (periodic ? 1 : 2)
    pl.inv_norm = (real)(1.0 / ((double)(n_global[d] - 1) * (periodic[d] ? 1.0 : 2.0)));

#ifdef MIFGPU_FP32
    plan->fast_logm[d] = plan->rfft_logm[d] = 0;  // every line length on sweep_kernel
#else
    plan->fast_logm[d] = (!periodic[d] && pow2 && pl.logP >= 6 && pl.logP <= 11) ? pl.logP : 0;
    static const bool no_rfft = getenv("MIFGPU_NO_WARP_RFFT") != nullptr;  // A/B switch for profiling
    plan->rfft_logm[d] = (periodic[d] && n % 2 == 0 && pow2 && (pl.logP == 8 || pl.logP == 9) && !no_rfft) ? pl.logP : 0;
#endif
    // lines per CTA: the largest power of two <= 8 that fits a ~100 KB shared-memory budget (two CTAs per SM)
    const size_t per_line = (size_t)P * sizeof(real2) + (size_t)(n | 1) * sizeof(real);
    int L = 8;
    while (L > 1 && per_line * L > 100 * 1024) L >>= 1;
    plan->L[d] = L;
    plan->smem[d] = per_line * L;
  }
  return plan;
}

// Directions whose lines go through the generic kernel need one line (FFT work array + real line) in shared memory.
int poisson_plan_unsupported_direction(const PoissonPlan *plan) {
  for (int d = 0; d < 3; d++)
    if (plan->fast_logm[d] == 0 && plan->rfft_logm[d] == 0 && plan->smem[d] > 200 * 1024) return d;
  return -1;
}

void poisson_plan_destroy(PoissonPlan *plan) {
  if (!plan) return;
  for (void *p : plan->allocations) cudaFree(p);
  delete plan;
}

namespace {

// Where the lines of one sweep live: `origin` is the first point of the first line, consecutive points of a line
// are `estride` apart, consecutive lines of a tile `lstride`, tiles `tile_stride`, outer index `outer_stride`.
struct SweepLayout {
  long long origin, lstride, estride, tile_stride, outer_stride;
  int n_tile_lines, outer;
  bool contig;       // estride == 1 (x sweeps)
  int lam_y_offset;  // first global y index of outer index 0 (z sweeps on a y-distributed pencil)
  int lam_x_offset = 0;  // first global x index of tile line 0 (pencils whose x range is distributed, Py > 1)
  bool has_origin;   // this rank holds the (0,0,0) mode
  SegMap load_map, store_map;
  bool peer = false;  // peer-memory sweeps: tiles of exactly 8 lines (the x tiles of the blocked buffers)
};

#ifndef MIFGPU_FP32
void launch_tma_modes(cudaStream_t stream, tmasweep::Cache &cache, const tmasweep::MapSet &maps, const tmasweep::Job &job, int logm,
                      int mode);

// TMA-staged strided sweep (mif_poisson_tma.cuh) when the geometry allows it: 257- or 513-point DCT-I lines along y or
// z, tiles of 8 consecutive x, plain strided addressing with 16-byte aligned rows.  Returns false when the launch has
// to take the LSU path (MIFGPU_NO_TMA=1 forces that for A/B runs; MIFGPU_REQUIRE_TMA=1 turns a refusal into an abort
// so that tests cannot pass on the fallback by accident).
bool launch_tma_sweep(cudaStream_t stream, PoissonPlan *plan, double *field, int d, int mode, const SweepLayout &lay) {
  static const bool disabled = getenv("MIFGPU_NO_TMA") != nullptr;
  static const bool required = getenv("MIFGPU_REQUIRE_TMA") != nullptr;
  const int logm = plan->fast_logm[d];
  if (disabled || lay.contig || (logm != 8 && logm != 9 && logm != 10) || lay.load_map.n || lay.store_map.n || lay.peer) return false;
  auto refuse = [&](const char *why) {
    if (required) {
      fprintf(stderr, "libmifgpu: MIFGPU_REQUIRE_TMA=1 but the sweep along %d cannot use TMA: %s\n", d, why);
      abort();
    }
    return false;
  };
  if (lay.lstride != 1 || lay.tile_stride != 1 || (lay.estride & 1) || (lay.outer_stride & 1)) return refuse("strides");
  tmasweep::Cache &cache = plan->tma;
  if (cache.device < 0) {
    cudaGetDevice(&cache.device);
    cudaDeviceGetAttribute(&cache.sms, cudaDevAttrMultiProcessorCount, cache.device);
  }
  static const bool swizzle = getenv("MIFGPU_TMA_NO_SWIZZLE") == nullptr;
  static const bool radix8_fft = getenv("MIFGPU_FFT_RADIX8") != nullptr;  // A/B: 513-point lines on the radix-8 Stockham passes
  static const int promo = getenv("MIFGPU_TMA_L2PROMO") ? atoi(getenv("MIFGPU_TMA_L2PROMO")) : 2;
  const int x_off = (int)(lay.origin & 1);
  const tmasweep::TensorDesc where{field + lay.origin - x_off, x_off + lay.n_tile_lines, lay.estride, lay.outer_stride, lay.outer};
  const tmasweep::MapSet *maps = tmasweep::maps_for(cache, where, where, plan->dir[d].n, swizzle, promo);
  if (!maps) return refuse("tensor map");
  tmasweep::Job job;
  job.n_xtiles = (lay.n_tile_lines + warpfft::kLines - 1) / warpfft::kLines;
  job.n_outer = lay.outer;
  job.n_lines = lay.n_tile_lines;
  job.x_off = x_off;
  job.tw = plan->dir[d].tw_full;
  job.cs = plan->dir[d].unpack;
  job.lam_x = plan->dir[0].lambda + lay.lam_x_offset;
  job.lam_y = plan->dir[1].lambda + lay.lam_y_offset;
  job.lam_z = plan->dir[2].lambda;
  job.inv_norm = plan->dir[d].inv_norm;
  job.has_origin = lay.has_origin ? 1 : 0;
  job.swz = swizzle ? 3u : 0u;
  job.outer_fastest = 0;
  job.in_x_tiled = 1;
  job.in_c2_mult = 0;
  job.in_perm_base = -1;
  job.out.n = 0;
  launch_tma_modes(stream, cache, *maps, job, logm, mode);
  return true;
}

void launch_tma_modes(cudaStream_t stream, tmasweep::Cache &cache, const tmasweep::MapSet &maps_ref, const tmasweep::Job &job_in, int logm,
                      int mode) {
  static const bool radix8_fft = getenv("MIFGPU_FFT_RADIX8") != nullptr;  // A/B: 513-point lines on the radix-8 Stockham passes
  const tmasweep::MapSet *maps = &maps_ref;
  const tmasweep::Job &job = job_in;
  if (logm == 10) {
    if (mode == 0) tmasweep::launch_1024<0>(stream, cache, *maps, job);
    else if (mode == 1) tmasweep::launch_1024<1>(stream, cache, *maps, job);
    else tmasweep::launch_1024<2>(stream, cache, *maps, job);
  } else if (logm == 8) {
    if (mode == 0) tmasweep::launch_one<8, 0>(stream, cache, *maps, job);
    else if (mode == 1) tmasweep::launch_one<8, 1>(stream, cache, *maps, job);
    else tmasweep::launch_one<8, 2>(stream, cache, *maps, job);
  } else if (radix8_fft) {
    if (mode == 0) tmasweep::launch_one<9, 0>(stream, cache, *maps, job);
    else if (mode == 1) tmasweep::launch_one<9, 1>(stream, cache, *maps, job);
    else tmasweep::launch_one<9, 2>(stream, cache, *maps, job);
  } else {
    if (mode == 0) tmasweep::launch_512<0>(stream, cache, *maps, job);
    else if (mode == 1) tmasweep::launch_512<1>(stream, cache, *maps, job);
    else tmasweep::launch_512<2>(stream, cache, *maps, job);
  }
}

#endif  // !MIFGPU_FP32

void launch_sweep(cudaStream_t stream, PoissonPlan *plan, real *field, int d, int mode, const SweepLayout &lay,
                  uint64_t *launches) {
  SmemAttrOnce &attrs = plan->attrs;
  attrs.ensure(sweep_kernel, 200 * 1024);
#ifndef MIFGPU_FP32
  if (plan->fast_logm[d] > 0 || plan->rfft_logm[d] > 0) {
    FastJob fj;
    fj.origin = lay.origin; fj.lstride = lay.lstride; fj.estride = lay.estride;
    fj.tile_stride = lay.tile_stride; fj.outer_stride = lay.outer_stride;
    fj.n_tile_lines = lay.n_tile_lines; fj.mode = mode;
    fj.tw = plan->dir[d].tw_full; fj.cs = plan->dir[d].unpack;
    fj.lam_x = plan->dir[0].lambda + lay.lam_x_offset; fj.lam_y = plan->dir[1].lambda + lay.lam_y_offset; fj.lam_z = plan->dir[2].lambda;
    fj.inv_norm = plan->dir[d].inv_norm;
    fj.has_origin = lay.has_origin;
    fj.load_map = lay.load_map; fj.store_map = lay.store_map;
    const dim3 fgrid((lay.n_tile_lines + fast::kLines - 1) / fast::kLines, lay.outer, 1);
    if (plan->rfft_logm[d] > 0) {
      if (plan->rfft_logm[d] == 8) launch_rfft<8>(stream, attrs, fj, lay.contig, fgrid, field);
      else launch_rfft<9>(stream, attrs, fj, lay.contig, fgrid, field);
      ++*launches;
      return;
    }
    if (launch_tma_sweep(stream, plan, field, d, mode, lay)) {  // strided 257- / 513- / 1025-point lines
      ++*launches;
      return;
    }
    // A/B switch: 513- and 1025-point lines on the radix-8 Stockham passes of mif_fft_warp.cuh instead of the 16 x 32 transform
    static const bool radix8_fft = getenv("MIFGPU_FFT_RADIX8") != nullptr;
    const bool aligned = ((lay.origin | lay.lstride | lay.tile_stride | lay.outer_stride) & 1) == 0;
    switch (plan->fast_logm[d]) {
      case 6: launch_fast<6>(stream, attrs, fj, lay.contig, fgrid, field); break;
      case 7: launch_fast<7>(stream, attrs, fj, lay.contig, fgrid, field); break;
      case 8: launch_warp<8>(stream, attrs, fj, lay.contig, fgrid, field); break;
      case 9:
        if (lay.contig && !radix8_fft && aligned && mode != 2) launch_x512(stream, attrs, fj, mode, lay.outer, field);
        else launch_warp<9>(stream, attrs, fj, lay.contig, fgrid, field);
        break;
      case 11:  // 2049-point lines (BASELINE configs[4]): four warps per line, 4-line CTAs
        launch_warp<11>(stream, attrs, fj, lay.contig, fgrid, field);
        break;
      default:  // 1025-point lines
        if (lay.contig && !radix8_fft && aligned && mode != 2) launch_x1024(stream, attrs, fj, mode, lay.outer, field);
        else if (lay.peer) launch_split<8>(stream, attrs, fj, lay.contig, lay.outer, field);  // 8-line tiles of the blocked buffers
        else launch_split<4>(stream, attrs, fj, lay.contig, lay.outer, field);
        break;
    }
    ++*launches;
    return;
  }
#endif
  SweepJob job;
  job.plan = plan->dir[d];
  job.dir = d;
  job.L = plan->L[d];
  job.r_pitch = plan->dir[d].n | 1;
  job.mode = mode;
  job.lam_a = plan->dir[0].lambda + lay.lam_x_offset;
  job.lam_b = plan->dir[1].lambda + lay.lam_y_offset;
  job.has_origin = lay.has_origin;
  job.n_tile_lines = lay.n_tile_lines;
  job.lstride = lay.lstride;
  job.estride = lay.estride;
  const dim3 grid((job.n_tile_lines + job.L - 1) / job.L, lay.outer, 1);
  sweep_kernel<<<grid, 256, plan->smem[d], stream>>>(job, field, lay.origin, lay.tile_stride, lay.outer_stride);
  ++*launches;
}

}  // namespace

void launch_poisson_sweep(cudaStream_t stream, const Geom &g, PoissonPlan *plan, real *field, int d, int mode,
                          uint64_t *launches) {
  const int nx = g.own_hi[0] - g.own_lo[0], ny = g.own_hi[1] - g.own_lo[1], nz = g.own_hi[2] - g.own_lo[2];
  SweepLayout lay;
  lay.origin = gidx(g, g.own_lo[0], g.own_lo[1], g.own_lo[2]);
  lay.lam_y_offset = 0;
  lay.has_origin = true;
  lay.contig = (d == 0);
  if (d == 0) {  // lines along x, tile over y, outer z
    lay.n_tile_lines = ny; lay.lstride = g.PX; lay.estride = 1;
    lay.tile_stride = g.PX; lay.outer_stride = g.plane; lay.outer = nz;
  } else if (d == 1) {  // lines along y, tile over x, outer z
    lay.n_tile_lines = nx; lay.lstride = 1; lay.estride = g.PX;
    lay.tile_stride = 1; lay.outer_stride = g.plane; lay.outer = nz;
  } else {  // lines along z, tile over x, outer y
    lay.n_tile_lines = nx; lay.lstride = 1; lay.estride = g.plane;
    lay.tile_stride = 1; lay.outer_stride = g.PX; lay.outer = ny;
  }
  launch_sweep(stream, plan, field, d, mode, lay, launches);
}

bool poisson_peer_capable(const PoissonPlan *plan) {
#ifdef MIFGPU_FP32
  (void)plan;
  return false;  // multi-GPU float runs take the pack -> all-to-all -> unpack transposes
#endif
  return plan->fast_logm[1] >= 8 && plan->fast_logm[1] <= 10 && plan->fast_logm[2] >= 8 && plan->fast_logm[2] <= 10;
}

#ifndef MIFGPU_FP32
// Blocked buffer layouts of the peer path (nxt = PX / 8 x tiles; all extents in doubles):
//   zbuf[r]  z pencil of rank r:          [z (all N_z)][x tile][y in r's range][8]
//   xfer[r]  slab staging of rank r:      [x tile][y (all N_y)][z in r's slab][8]
// A tile of the forward y sweep (8 x, all y, one z) is then one contiguous run per owning rank in zbuf[r], a tile of the
// fused z sweep (8 x, one y, all z) one contiguous run per owning rank in xfer[r], and the inverse y sweep reads its own
// xfer with one row stride for all y.
static void fill_map(SegMap &m, const PeerLayout &p, int which, const Geom &g) {
  const int me = p.rank, P = p.nranks;
  const long long nxt = g.PX / 8;
  const long long nz_me = p.zlo[me + 1] - p.zlo[me], ny_all = p.ylo[P];
  if (which == 2) {  // inverse y sweep <- own slab staging: one segment
    m.n = 1;
    m.lo[0] = 0;
    m.lo[1] = p.ylo[P];
    m.base[0] = p.xfer[me];
    m.xtile_stride[0] = ny_all * nz_me * 8;
    m.estride[0] = nz_me * 8;
    m.outer_stride[0] = 8;
    return;
  }
  m.n = P;
  for (int r = 0; r < P; r++) {
    const long long ny_r = p.ylo[r + 1] - p.ylo[r], nz_r = p.zlo[r + 1] - p.zlo[r];
    if (which == 0) {         // forward y sweep -> z pencils: segments are y ranges, outer index is the local z plane
      m.lo[r] = p.ylo[r];
      m.xtile_stride[r] = ny_r * 8;
      m.estride[r] = 8;
      m.outer_stride[r] = nxt * ny_r * 8;
      m.base[r] = p.zbuf[r] + (long long)p.zlo[me] * m.outer_stride[r];
    } else {                  // fused z sweep -> slab staging of the owners: segments are z ranges, outer is local y
      m.lo[r] = p.zlo[r];
      m.xtile_stride[r] = ny_all * nz_r * 8;
      m.outer_stride[r] = nz_r * 8;
      m.estride[r] = 8;
      m.base[r] = p.xfer[r] + (long long)p.ylo[me] * nz_r * 8;
    }
  }
  m.lo[P] = (which == 1) ? p.zlo[P] : p.ylo[P];
}

// The three sweeps of the peer path on the TMA kernels (257-, 513- or 1025-point lines along y and z): the forward y sweep
// reads the slab and leaves its tiles as bulk copies in the z pencils of the owners, the fused z sweep reads its pencil
// (tiles = (x tile, local y), rows along z) and leaves its tiles in the slab staging of the owners, the inverse y sweep
// reads its staging buffer (tiles = (x tile, local z), rows along y) and stores into the field.  The output stage is
// copied as it is, i.e. swizzled: the consumer un-permutes the column pairs (Job::in_perm_base).
static bool launch_tma_peer(cudaStream_t stream, const Geom &g, PoissonPlan *plan, double *field, const PeerLayout &peer, int which) {
  static const bool disabled = getenv("MIFGPU_NO_TMA") != nullptr || getenv("MIFGPU_NO_TMA_PEER") != nullptr;
  const int logm_y = plan->fast_logm[1], logm_z = plan->fast_logm[2];
  if (disabled || logm_y < 8 || logm_y > 10 || logm_z < 8 || logm_z > 10) return false;
  tmasweep::Cache &cache = plan->tma;
  if (cache.device < 0) {
    cudaGetDevice(&cache.device);
    cudaDeviceGetAttribute(&cache.sms, cudaDevAttrMultiProcessorCount, cache.device);
  }
  static const bool swizzle = getenv("MIFGPU_TMA_NO_SWIZZLE") == nullptr;
  static const int promo = getenv("MIFGPU_TMA_L2PROMO") ? atoi(getenv("MIFGPU_TMA_L2PROMO")) : 2;
  const int me = peer.rank, P = peer.nranks;
  const int nx = g.own_hi[0] - g.own_lo[0], nz_me = g.own_hi[2] - g.own_lo[2];
  const long long nxt = g.PX / 8, ny_all = peer.ylo[P], nz_all = peer.zlo[P], ny_me = peer.ylo[me + 1] - peer.ylo[me];
  const long long origin = gidx(g, g.own_lo[0], g.own_lo[1], g.own_lo[2]);
  const int x_off = (int)(origin & 1);
  if (ny_me <= 0 || nz_me <= 0) return false;
  const tmasweep::TensorDesc slab{field + origin - x_off, x_off + nx, g.PX, g.plane, nz_me};
  tmasweep::Job job;
  job.n_xtiles = (nx + warpfft::kLines - 1) / warpfft::kLines;
  job.n_lines = nx;
  job.x_off = x_off;
  job.lam_x = plan->dir[0].lambda;
  job.lam_y = plan->dir[1].lambda;
  job.lam_z = plan->dir[2].lambda;
  job.has_origin = 0;
  job.swz = swizzle ? 3u : 0u;
  job.out.n = 0;
  const int d = (which == 1) ? 2 : 1;
  job.tw = plan->dir[d].tw_full;
  job.cs = plan->dir[d].unpack;
  job.inv_norm = plan->dir[d].inv_norm;
  const tmasweep::MapSet *maps = nullptr;
  int mode;
  if (which == 0) {
    maps = tmasweep::maps_for(cache, slab, slab, plan->dir[1].n, swizzle, promo);
    job.n_outer = nz_me;
    job.outer_fastest = 0; job.in_x_tiled = 1; job.in_c2_mult = 0; job.in_perm_base = -1;
    job.out.n = P;
    for (int r = 0; r < P; r++) {
      const long long ny_r = peer.ylo[r + 1] - peer.ylo[r];
      job.out.lo[r] = peer.ylo[r];
      job.out.xtile_stride[r] = ny_r * 8;
      job.out.outer_stride[r] = nxt * ny_r * 8;
      job.out.base[r] = peer.zbuf[r] + (long long)peer.zlo[me] * job.out.outer_stride[r];
    }
    job.out.lo[P] = peer.ylo[P];
    mode = 0;
  } else if (which == 1) {
    // zbuf[me][z][x tile][y local][8]: coordinate 2 = x tile * ny_me + local y
    const tmasweep::TensorDesc pencil{peer.zbuf[me], 8, nxt * ny_me * 8, 8, nxt * ny_me};
    maps = tmasweep::maps_for(cache, pencil, slab, plan->dir[2].n, swizzle, promo);
    job.n_outer = (int)ny_me;
    job.outer_fastest = 1; job.in_x_tiled = 0; job.in_c2_mult = ny_me; job.in_perm_base = peer.ylo[me];
    job.lam_y = plan->dir[1].lambda + peer.ylo[me];
    job.has_origin = peer.ylo[me] == 0;
    job.out.n = P;
    for (int r = 0; r < P; r++) {
      const long long nz_r = peer.zlo[r + 1] - peer.zlo[r];
      job.out.lo[r] = peer.zlo[r];
      job.out.xtile_stride[r] = ny_all * nz_r * 8;
      job.out.outer_stride[r] = nz_r * 8;
      job.out.base[r] = peer.xfer[r] + (long long)peer.ylo[me] * nz_r * 8;
    }
    job.out.lo[P] = peer.zlo[P];
    mode = 2;
  } else {
    // xfer[me][x tile][y][z local][8]: coordinate 2 = x tile * (N_y * nz_me) + local z
    const tmasweep::TensorDesc staging{peer.xfer[me], 8, (long long)nz_me * 8, 8, (long long)(job.n_xtiles - 1) * ny_all * nz_me + nz_me};
    maps = tmasweep::maps_for(cache, staging, slab, plan->dir[1].n, swizzle, promo);
    job.n_outer = nz_me;
    job.outer_fastest = 1; job.in_x_tiled = 0; job.in_c2_mult = ny_all * nz_me; job.in_perm_base = peer.zlo[me];
    mode = 1;
  }
  (void)nz_all;
  if (!maps) {
    if (getenv("MIFGPU_REQUIRE_TMA")) {
      fprintf(stderr, "libmifgpu: MIFGPU_REQUIRE_TMA=1 but peer sweep %d cannot use TMA\n", which);
      abort();
    }
    return false;
  }
  if (getenv("MIFGPU_TMA_VERBOSE")) fprintf(stderr, "libmifgpu: rank %d peer sweep %d on the TMA path (%d-point lines)\n", me, which, plan->dir[d].n);
  launch_tma_modes(stream, cache, *maps, job, which == 1 ? logm_z : logm_y, mode);
  return true;
}

#endif  // !MIFGPU_FP32

void launch_poisson_sweep_peer(cudaStream_t stream, const Geom &g, PoissonPlan *plan, real *field, const PeerLayout &peer,
                               int which, uint64_t *launches) {
#ifdef MIFGPU_FP32
  (void)stream; (void)g; (void)plan; (void)field; (void)peer; (void)which; (void)launches;
  fprintf(stderr, "libmifgpu_f32: the peer-memory sweeps are FP64 only (poisson_peer_capable is false in this build)\n");
  abort();
#else
  if (launch_tma_peer(stream, g, plan, field, peer, which)) {
    ++*launches;
    return;
  }
  const int me = peer.rank;
  const int nx = g.own_hi[0] - g.own_lo[0], nz = g.own_hi[2] - g.own_lo[2];
  SweepLayout lay;
  lay.lam_y_offset = 0;
  lay.has_origin = false;
  lay.contig = false;
  lay.n_tile_lines = nx;
  lay.lstride = 1;
  lay.peer = true;  // 8-line tiles: the x tiles of the blocked buffers
  if (which == 0 || which == 2) {
    // y sweeps on the local slab: lines along y, tile over x, outer over the local owner z planes
    lay.origin = gidx(g, g.own_lo[0], g.own_lo[1], g.own_lo[2]);
    lay.tile_stride = 1;
    lay.estride = g.PX;
    lay.outer_stride = g.plane;
    lay.outer = nz;
    if (which == 0) fill_map(lay.store_map, peer, 0, g);
    else fill_map(lay.load_map, peer, 2, g);
    launch_sweep(stream, plan, field, 1, which == 0 ? 0 : 1, lay, launches);
  } else {
    // fused z sweep on the local z pencil zbuf[z][x tile][y_local][8]: the tile of 8 lines starting at x = 8 t begins
    // at t * ny_local * 8, i.e. first_line * ny_local
    const long long ny_local = peer.ylo[me + 1] - peer.ylo[me];
    lay.origin = 0;
    lay.tile_stride = ny_local;
    lay.estride = (long long)(g.PX / 8) * ny_local * 8;
    lay.outer_stride = 8;
    lay.outer = (int)ny_local;
    lay.lam_y_offset = peer.ylo[me];
    lay.has_origin = peer.ylo[me] == 0;
    fill_map(lay.store_map, peer, 1, g);
    launch_sweep(stream, plan, peer.zbuf[me], 2, 2, lay, launches);
  }
#endif
}

void launch_poisson_zpencil(cudaStream_t stream, const Geom &g, PoissonPlan *plan, real *zbuf, int ny_local,
                            int y_offset, bool has_origin, uint64_t *launches) {
  // zbuf[z][y_local][x]: x rows of PX doubles, ny_local rows per z plane, all N_z transform points.
  SweepLayout lay;
  lay.origin = g.own_lo[0];
  lay.n_tile_lines = g.own_hi[0] - g.own_lo[0];
  lay.lstride = 1;
  lay.tile_stride = 1;
  lay.estride = (long long)ny_local * g.PX;
  lay.outer_stride = g.PX;
  lay.outer = ny_local;
  lay.contig = false;
  lay.lam_y_offset = y_offset;
  lay.has_origin = has_origin;
  launch_sweep(stream, plan, zbuf, 2, 2, lay, launches);
}

// One sweep on a pencil buffer of the Py x Pz decomposition (rows of `pitch` doubles, nx_local of them used):
//   dir = 1: y pencil buf[z_local][y (all points)][x_local], n_outer = local z planes: lines along y;
//   dir = 2: z pencil buf[z (all points)][y_local][x_local], n_outer = local y rows: lines along z (mode 2: fused).
// x_offset / y_offset: global transform indices of local x = 0 and (dir = 2) local y = 0, for the eigenvalues.
void launch_poisson_pencil(cudaStream_t stream, PoissonPlan *plan, real *buf, int dir, int mode, int nx_local, int pitch,
                           int n_outer, int x_offset, int y_offset, bool has_origin, uint64_t *launches) {
  if (nx_local <= 0 || n_outer <= 0) return;  // this rank holds no line of the pencil (fewer rows than ranks)
  SweepLayout lay;
  lay.origin = 0;
  lay.n_tile_lines = nx_local;
  lay.lstride = 1;
  lay.tile_stride = 1;
  lay.contig = false;
  lay.outer = n_outer;
  if (dir == 1) {
    lay.estride = pitch;
    lay.outer_stride = (long long)plan->dir[1].n * pitch;
  } else {
    lay.estride = (long long)n_outer * pitch;
    lay.outer_stride = pitch;
  }
  lay.lam_x_offset = x_offset;
  lay.lam_y_offset = y_offset;
  lay.has_origin = has_origin;
  launch_sweep(stream, plan, buf, dir, mode, lay, launches);
}

}  // namespace mifgpu
