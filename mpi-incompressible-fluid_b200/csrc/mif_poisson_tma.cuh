// mif_poisson_tma.cuh -- strided (y / z) DCT-I sweeps of the Poisson solve staged by TMA (sm_100a, FP64).
// Included by mif_poisson.cu.
//
// Replaces, for lines of 2^k + 1 points (k = 8, 9) along y or z, the reference's 1-D REDFT00 sweeps and the 2Decomp
// pack / transpose / unpack passes around them (src/PressureEquation.cpp:106-229,
// deps/2Decomp_C/MemSplitMerge.cpp:59-86,116-143): the lines stay where they are in the x-fastest array and a tile of
// 8 lines (8 consecutive x, i.e. 64-byte rows) is moved by the copy engine instead of by the warps.
//
//   * persistent CTAs (2 per SM, 8 warps, one warp per line), tiles handed out round-robin, x tile fastest;
//   * load: cp.async.bulk.tensor boxes of 8 x 130 doubles into a shared stage, completion on an mbarrier.  The stage
//     is free again as soon as every thread has taken its first-pass inputs into registers, so the boxes of the NEXT
//     tile are in flight while the current one is transformed -- no warp ever waits for HBM in steady state;
//   * forward sweeps read the even extension e(2q), e(2q+1) and its mirror image: two tensor maps over the even and
//     the odd rows land them in two separate areas, which makes the line-fastest first-pass reads (8 lines x 4
//     consecutive rows per warp = 256 contiguous bytes) bank-conflict free;
//   * inverse sweeps do not transform the even extension again.  DCT-I of a real sequence E is the unnormalised
//     half-complex -> real transform of the real spectrum E, and that is ONE length-M complex FFT of
//     conj Z_k = (E_k + E_{M-k}) - i exp(i pi k / M) (E_k - E_{M-k}),  x_{2q} = Re z_q, x_{2q+1} = -Im z_q:
//     513 points are read once (natural rows), no unpack pass follows, and the mirror half of the output is never
//     formed.  In the fused z sweep the scaled spectrum is repacked in registers (partners by shuffles) and is the
//     first-pass input of the second FFT without touching shared memory;
//   * store: every warp writes its line's column into a swizzled output stage (aliasing the FFT work area)
//     and one thread issues cp.async.bulk.tensor stores; they drain while the next tile's inputs are read.
#pragma once

#include <map>
#include <tuple>
#include <vector>

#include "mif_fft512.cuh"
#include "mif_fft_warp.cuh"
#include "mif_tma.cuh"

namespace mifgpu {
namespace tmasweep {

using namespace warpfft;

constexpr int kBoxRows = 130;       // rows of one load box (x 8 doubles = 8320 bytes: box bases stay 128-byte aligned)
constexpr int kStoreRows = 128;     // rows of one store box (8192 bytes, one swizzle period = 8 rows)
constexpr int kThreads = 256;

template <int LOGM>
struct Layout {
  using C = Cfg<LOGM>;
  static constexpr int M = C::M;
  static constexpr int kLoadBoxes = 2 * ((M / 2 + 1 + kBoxRows - 1) / kBoxRows);  // even rows + odd rows; also covers M + 1 natural rows
  static constexpr int kOddRow0 = (kLoadBoxes / 2) * kBoxRows;                    // first row of the odd area
  static constexpr int kStoreBoxes = (M + 1 + kStoreRows - 1) / kStoreRows;
  static constexpr size_t kWorkBytes = (size_t)kLines * C::LINE_PITCH * sizeof(double2);
  static constexpr size_t kOutBytes = (size_t)kStoreBoxes * kStoreRows * 64;
  static constexpr size_t kTwBytes = (size_t)C::TW_TOTAL * sizeof(double2);
  static constexpr size_t kStageBytes = (size_t)kLoadBoxes * kBoxRows * 64;
  static constexpr size_t kWorkArea = ((kWorkBytes > kOutBytes ? kWorkBytes : kOutBytes) + 127) / 128 * 128;
  // [work / output stage][twiddles][input stage][mbarrier], after aligning the dynamic window to 1024 bytes
  static constexpr size_t kSmem = 1024 + kWorkArea + (kTwBytes + 127) / 128 * 128 + kStageBytes + 64;
  static_assert(kLoadBoxes * kBoxRows >= M + 1, "natural rows fit the stage");
};

// Multi-GPU sweeps whose results go to the pencil / slab buffers of the owning ranks (fused 2Decomp transposes): rows
// [lo[r], lo[r+1]) of every tile are one contiguous run in rank r's buffer (blocked layouts of mif_poisson.cu), so they
// leave the output stage as one bulk copy per rank -- over NVLink for r != this rank.
struct PeerOut {
  int n;  // 0: one tensor store through map_out
  int lo[9];
  double *base[8];
  long long xtile_stride[8], outer_stride[8];
};

struct Job {
  int n_xtiles, n_outer, n_lines, x_off;
  const double2 *tw, *cs;
  const double *lam_x, *lam_y, *lam_z;
  double inv_norm;
  int has_origin;
  unsigned swz;  // 3: the output stage is swizzled (CU_TENSOR_MAP_SWIZZLE_64B), 0: plain
  // input side
  int outer_fastest;     // tile order: 0 = x tile fastest (plain arrays), 1 = outer index fastest (blocked buffers)
  int in_x_tiled;        // load coordinate 0 = x_off + 8 xt (plain arrays) or 0 (blocked buffers: x tile folded into coordinate 2)
  long long in_c2_mult;  // load coordinate 2 = xt * in_c2_mult + outer
  int in_perm_base;      // >= 0: the input rows left a swizzled output stage in which this tile's lines were row in_perm_base +
                         // outer, so the 16-byte pairs of its 8 columns arrive XOR-permuted by ((row >> 1) & 3); -1: in order
  PeerOut out;
#ifdef MIFGPU_PHASE_TRACE
  unsigned long long *trace;  // diagnostic build only (make trace): SM-clock stamps at the phase boundaries of two CTAs
#endif
};

#ifdef MIFGPU_PHASE_TRACE
constexpr int kTraceTiles = 24, kTraceSlots = 16;
#define MIF_TRACE(id)                                                                                                          \
  do {                                                                                                                          \
    if (trace_cta >= 0 && j == 0 && n_traced < kTraceTiles)                                                                     \
      job.trace[((size_t)(trace_cta * 8 + line) * kTraceTiles + n_traced) * kTraceSlots + (id)] = clock64();                    \
  } while (0)
#else
#define MIF_TRACE(id)
#endif

struct Tile {
  int xt, outer, x_in, c2, perm;
};
__device__ __forceinline__ Tile decode_tile(const Job &job, int tile) {
  Tile t;
  if (job.outer_fastest) {
    t.xt = tile / job.n_outer;
    t.outer = tile - t.xt * job.n_outer;
  } else {
    t.outer = tile / job.n_xtiles;
    t.xt = tile - t.outer * job.n_xtiles;
  }
  t.x_in = job.in_x_tiled ? job.x_off + t.xt * kLines : 0;
  t.c2 = (int)(t.xt * job.in_c2_mult) + t.outer;
  t.perm = (job.in_perm_base >= 0 && job.swz) ? (((job.in_perm_base + t.outer) >> 1) & 3) << 1 : 0;
  return t;
}

// The finished output stage O (rows of 64 bytes) leaves the CTA: one thread, one bulk group.
template <int STORE_BOXES>
__device__ __forceinline__ void store_tile(const Job &job, const CUtensorMap *map_out, const unsigned char *O, const Tile &t) {
  if (job.out.n == 0) {
    const int x0 = job.x_off + t.xt * kLines;
#pragma unroll
    for (int c = 0; c < STORE_BOXES; c++) tma::store_3d(map_out, O + c * kStoreRows * 64, x0, c * kStoreRows, t.outer);
  } else {
    for (int r = 0; r < job.out.n; r++) {
      const int rows = job.out.lo[r + 1] - job.out.lo[r];
      if (rows <= 0) continue;
      double *dst = job.out.base[r] + (long long)t.xt * job.out.xtile_stride[r] + (long long)t.outer * job.out.outer_stride[r];
      tma::store_bulk(dst, O + (size_t)job.out.lo[r] * 64, (unsigned)rows * 64u);
    }
  }
  tma::commit_group();
}

__device__ __forceinline__ double2 dct_pack_input(double xr, double yr, double c, double sn) {
  // conj Z_k for a real spectrum: X_k = xr, X_{M-k} = yr, (c, sn) = (cos, sin)(pi k / M); see hc2r_input
  const double pr = xr + yr, dr = xr - yr;
  return make_double2(pr - sn * dr, -(c * dr));
}

// conj Z_k from the registers left by unpack_regs (spec[u + G t] = E_k, k = j + 32 u + NS t, e_last = E_M in lane 0):
// the partners E_{M-k} come from lane 32 - j, and slot u + G t is exactly first-pass slot s of the next transform.
template <int LOGM>
__device__ __forceinline__ void pack_dct_regs(const double *spec, double e_last, int j, const double2 *__restrict__ cs,
                                              double2 *v) {
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
      double yr = __shfl_sync(0xffffffffu, spec[partner], src);
      if (j == 0) {
        if (u == 0 && t == 0) yr = e_last;  // E_M
        else yr = spec[(u == 0) ? G * (R - t) : (G - u) + G * (R - 1 - t)];
      }
      const double2 w0 = base_w[u], r = rot8((8 / R) * t);
      const double c = w0.x * r.x - w0.y * r.y, sn = w0.x * r.y + w0.y * r.x;
      v[u + G * t] = dct_pack_input(spec[u + G * t], yr, c, sn);
    }
}

// MODE 0: forward; 1: inverse + normalisation; 2: forward, eigenvalue division, inverse + normalisation.
template <int LOGM, int MODE>
__global__ void __launch_bounds__(kThreads, 2)
    tma_dct_kernel(const __grid_constant__ CUtensorMap map_in_a, const __grid_constant__ CUtensorMap map_in_b,
                   const __grid_constant__ CUtensorMap map_out, const Job job) {
  using C = Cfg<LOGM>;
  using L = LastPass<LOGM>;
  using Y = Layout<LOGM>;
  constexpr int M = C::M, EPT = C::EPT, HALF = M / 2;
  static_assert(C::WPL == 1 && C::LINES == kLines && C::THREADS == kThreads, "one warp per line, 8 lines");
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024 - (tma::swizzle_address(smem_raw) & 1023)) & 1023);
  double2 *W = reinterpret_cast<double2 *>(smem);                                   // FFT work area, one region per line
  unsigned char *O = smem;                                                          // output stage (aliases W)
  double2 *T = reinterpret_cast<double2 *>(smem + Y::kWorkArea);                    // twiddle tables
  double *S = reinterpret_cast<double *>(smem + Y::kWorkArea + (Y::kTwBytes + 127) / 128 * 128);  // input stage
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(S) + Y::kStageBytes);

  const int tid = threadIdx.x;
  const int line = tid >> 5, j = tid & 31;  // FFT mapping: warp = line
  const int l = tid & 7, b = tid >> 3;      // stage mapping: line fastest
  double2 *Sline = W + line * C::LINE_PITCH;
  const int n_tiles = job.n_xtiles * job.n_outer;

  auto issue_load = [&](int tile) {
    const Tile nt = decode_tile(job, tile);
    tma::mbar_arrive_expect_tx(full, (unsigned)Y::kStageBytes);
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < Y::kLoadBoxes; i++) tma::load_3d(S + i * kBoxRows * 8, &map_in_a, full, nt.x_in, i * kBoxRows, nt.c2);
    } else {
#pragma unroll
      for (int i = 0; i < Y::kLoadBoxes / 2; i++) {
        tma::load_3d(S + i * kBoxRows * 8, &map_in_a, full, nt.x_in, i * kBoxRows, nt.c2);
        tma::load_3d(S + (Y::kOddRow0 + i * kBoxRows) * 8, &map_in_b, full, nt.x_in, i * kBoxRows, nt.c2);
      }
    }
  };

  if (tid == 0) {
    tma::prefetch_map(&map_in_a);
    tma::prefetch_map(&map_in_b);
    tma::prefetch_map(&map_out);
    tma::mbar_init(full, 1);
    tma::fence_barrier_init();
    tma::fence_proxy_async();
  }
  load_twiddles<LOGM>(T, job.tw);
  __syncthreads();
  int tile = blockIdx.x;
  if (tid == 0 && tile < n_tiles) issue_load(tile);
  unsigned parity = 0;

  for (; tile < n_tiles; tile += gridDim.x) {
    const Tile t = decode_tile(job, tile);
    const int xt = t.xt, outer = t.outer, col = line ^ t.perm;  // col: this line's x offset inside the tile
    double2 v[EPT];

    // ---- first-pass inputs from the stage ---------------------------------------------------------------------
    tma::mbar_wait(full, parity);
    parity ^= 1;
    if (MODE == 1) {
      const double *N = S + l;  // natural rows: element e of line l at N[8 e]
#pragma unroll
      for (int s = 0; s < EPT; s++) {
        const int k = b + 32 * s;
        const double2 w = __ldg(&job.cs[k]);
        v[s] = dct_pack_input(N[8 * k], N[8 * (M - k)], w.x, w.y);
      }
    } else {
      const double *Ev = S + l, *Od = S + Y::kOddRow0 * 8 + l;  // e(2q) at Ev[8 q], e(2q+1) at Od[8 q]
#pragma unroll
      for (int s = 0; s < EPT; s++) {
        const int q = b + 32 * s;
        if (s < EPT / 2) v[s] = make_double2(Ev[8 * q], Od[8 * q]);
        else v[s] = make_double2(Ev[8 * (M - q)], Od[8 * (M - q - 1)]);  // mirror image: e(2M-2q), e(2M-2q-1)
      }
    }
    // The first radix-8 butterflies run in registers BEFORE the barrier: the bulk stores of the previous tile (which read
    // the output stage inside W) get that time to drain, and the stage is released to the next tile's boxes right after.
    first_pass_compute<LOGM>(v);
    if (tid == 0) tma::wait_stores_read();  // the previous tile's output stage (in W) has been read out
    __syncthreads();                        // the stage has been consumed, W is free
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) issue_load(tile + gridDim.x);

    // ---- transform --------------------------------------------------------------------------------------------
    first_pass_store<LOGM>(W + l * C::LINE_PITCH, b, v);
    __syncthreads();
    fft_line<LOGM, false, true, true>(Sline, T, j, line, v);

    double spec[EPT], e_last = 0.0;
    if (MODE != 1) unpack_regs<LOGM>(v, j, job.cs, spec, e_last);
    if (MODE == 2) {
      // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z); mode (0,0,0) := 0 (src/PressureEquation.cpp:158-163)
      const int ix = min(xt * kLines + col, job.n_lines - 1);
      const double lam_xy = job.lam_x[ix] + job.lam_y[outer];
      const bool origin_line = job.has_origin && (xt * kLines + col == 0) && (outer == 0);
#pragma unroll
      for (int u = 0; u < L::G; u++)
#pragma unroll
        for (int t = 0; t < L::R; t++) {
          const int k = j + 32 * u + L::NS * t;
          spec[u + L::G * t] *= (origin_line && k == 0) ? 0.0 : 1.0 / (lam_xy + job.lam_z[k]);
        }
      e_last *= 1.0 / (lam_xy + job.lam_z[M]);
      pack_dct_regs<LOGM>(spec, e_last, j, job.cs, v);
      __syncwarp();  // every lane has finished reading the line region (last pass of the forward transform)
      fft_line<LOGM, true, true>(Sline, T, j, line, v);
    }
    __syncthreads();  // all warps are done with the work area: it becomes the output stage

    // ---- results into the swizzled output stage: row e of the tile at 64 e, column `line` -------------------------
    // A column access of one warp (32 rows, one 8-byte column) can spread over at most 8 bank positions: the row parity
    // picks the 64-byte half of a 128-byte line and the swizzle one of 4 chunks in it -- a 4-way conflict, half the
    // rate of a conflict-free access and still cheaper than a transposition through the line regions.
    if (MODE == 0) {
      // lane j holds E_k, k = j + 32 u + NS t; address bits 7-8 (row >> 1) do not depend on (u, t)
      const unsigned off = tma::swizzle_offset((unsigned)(j * 64 + col * 8), job.swz);
#pragma unroll
      for (int u = 0; u < L::G; u++)
#pragma unroll
        for (int t = 0; t < L::R; t++)
          *reinterpret_cast<double *>(O + off + (32 * u + L::NS * t) * 64) = spec[u + L::G * t];
      if (j == 0) *reinterpret_cast<double *>(O + tma::swizzle_offset((unsigned)(M * 64 + col * 8), job.swz)) = e_last;
    } else {
      // lane j holds z_q = conj(v), q = j + 32 u + NS t: x(2q) = Re z_q, x(2q+1) = Im z_q; only q <= M/2 is stored.
      // Rows 2q of all lanes would sit in the same half of their 128-byte lines (4 bank positions): lanes with bit 2 of
      // j set store the odd row first, so that every instruction covers both halves (8 positions).
      const unsigned off = tma::swizzle_offset((unsigned)(j * 128 + col * 8), job.swz);
      const bool odd_first = (j >> 2) & 1;
      const unsigned first = off + (odd_first ? 64u : 0u), second = first ^ 64u;
      const double scale = job.inv_norm;
#pragma unroll
      for (int u = 0; u < L::G; u++)
#pragma unroll
        for (int t = 0; t < L::R / 2; t++) {
          const unsigned at = (32 * u + L::NS * t) * 128;
          const double even = v[u + L::G * t].x * scale, odd = -v[u + L::G * t].y * scale;
          *reinterpret_cast<double *>(O + first + at) = odd_first ? odd : even;
          *reinterpret_cast<double *>(O + second + at) = odd_first ? even : odd;
        }
      if (j == 0)  // q = M/2 = NS * R/2: slot u = 0, t = R/2 of lane 0
        *reinterpret_cast<double *>(O + tma::swizzle_offset((unsigned)(M * 64 + col * 8), job.swz)) = v[L::G * (L::R / 2)].x * scale;
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) store_tile<Y::kStoreBoxes>(job, &map_out, O, t);
  }
  if (tid == 0) tma::wait_stores_done();
}


// ---- 513-point lines on the 16 x 32 transform of mif_fft512.cuh ------------------------------------------------------------
// Same staging, tile loop and barriers as tma_dct_kernel; the transform crosses shared memory once instead of twice:
// the stage threads (line fastest) run phase A -- a 16-point DFT and the W_512 twiddles in registers -- while the bulk
// stores of the previous tile drain, the line's warp runs phase B, and in the fused z sweep the scaled spectrum goes back
// to first-pass order by one exchange between lane pairs before the second transform.
struct Layout512 {
  static constexpr int M = 512;
  static constexpr int kLoadBoxes = Layout<9>::kLoadBoxes, kOddRow0 = Layout<9>::kOddRow0, kStoreBoxes = Layout<9>::kStoreBoxes;
  static constexpr size_t kWorkBytes = (size_t)kLines * fft512::kLinePitch * sizeof(double2);
  static constexpr size_t kOutBytes = Layout<9>::kOutBytes;
  static constexpr size_t kTwBytes = (size_t)fft512::kTwiddles * sizeof(double2);
  static constexpr size_t kStageBytes = Layout<9>::kStageBytes;
  static constexpr size_t kWorkArea = ((kWorkBytes > kOutBytes ? kWorkBytes : kOutBytes) + 127) / 128 * 128;
  static constexpr size_t kSmem = 1024 + kWorkArea + (kTwBytes + 127) / 128 * 128 + kStageBytes + 64;
};

// Scale the 16 spectrum values of a lane by 1 / (lam_xy + lam_z[k]) (src/PressureEquation.cpp:158-163) with ONE division
// per four values: prefix products, one reciprocal, back-substitution -- 9 multiplications replace three of the four
// divisions (each a MUFU.RCP64H plus two Newton steps and a guarded slow path).  The products stay far from the FP64
// range (|lambda sums| <= 12 / h^2); the (0,0,0) mode, whose sum is zero, takes the divisor 1 and the factor 0.
template <class KOf>
__device__ __forceinline__ void scale_by_inverse_eigenvalues(double *spec, double lam_xy, const double *__restrict__ lam_z,
                                                             bool origin_line, KOf k_of_register) {
#pragma unroll
  for (int q = 0; q < 4; q++) {
    double d[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int k = k_of_register(4 * q + e);
      d[e] = (origin_line && k == 0) ? 1.0 : lam_xy + lam_z[k];
    }
    const double p01 = d[0] * d[1], p012 = p01 * d[2];
    double inv = 1.0 / (p012 * d[3]);
    const double r3 = inv * p012;
    inv *= d[3];
    const double r2 = inv * p01;
    inv *= d[2];
    const double r1 = inv * d[0], r0 = inv * d[1];
    spec[4 * q] *= (origin_line && k_of_register(4 * q) == 0) ? 0.0 : r0;
    spec[4 * q + 1] *= r1;
    spec[4 * q + 2] *= r2;
    spec[4 * q + 3] *= r3;
  }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2)
    tma_dct512_kernel(const __grid_constant__ CUtensorMap map_in_a, const __grid_constant__ CUtensorMap map_in_b,
                      const __grid_constant__ CUtensorMap map_out, const Job job) {
  using Y = Layout512;
  constexpr int M = 512;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024 - (tma::swizzle_address(smem_raw) & 1023)) & 1023);
  double2 *W = reinterpret_cast<double2 *>(smem);                                   // phase A -> phase B exchange, one region per line
  unsigned char *O = smem;                                                          // output stage (aliases W)
  double2 *T = reinterpret_cast<double2 *>(smem + Y::kWorkArea);                    // W_512^(n1 2^e)
  double *S = reinterpret_cast<double *>(smem + Y::kWorkArea + (Y::kTwBytes + 127) / 128 * 128);  // input stage
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(S) + Y::kStageBytes);

  const int tid = threadIdx.x;
  const int line = tid >> 5, j = tid & 31;  // transform mapping: warp = line, lane j = k2 + 16 p
  const int l = tid & 7, b = tid >> 3;      // stage mapping: line fastest, b = n1
  double2 *Sline = W + line * fft512::kLinePitch;
  const int n_tiles = job.n_xtiles * job.n_outer;

  auto issue_load = [&](int tile) {
    const Tile nt = decode_tile(job, tile);
    tma::mbar_arrive_expect_tx(full, (unsigned)Y::kStageBytes);
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < Y::kLoadBoxes; i++) tma::load_3d(S + i * kBoxRows * 8, &map_in_a, full, nt.x_in, i * kBoxRows, nt.c2);
    } else {
#pragma unroll
      for (int i = 0; i < Y::kLoadBoxes / 2; i++) {
        tma::load_3d(S + i * kBoxRows * 8, &map_in_a, full, nt.x_in, i * kBoxRows, nt.c2);
        tma::load_3d(S + (Y::kOddRow0 + i * kBoxRows) * 8, &map_in_b, full, nt.x_in, i * kBoxRows, nt.c2);
      }
    }
  };

  if (tid == 0) {
    tma::prefetch_map(&map_in_a);
    tma::prefetch_map(&map_in_b);
    tma::prefetch_map(&map_out);
    tma::mbar_init(full, 1);
    tma::fence_barrier_init();
    tma::fence_proxy_async();
  }
  fft512::load_twiddles<kThreads>(T, job.tw);
  __syncthreads();
  int tile = blockIdx.x;
  if (tid == 0 && tile < n_tiles) issue_load(tile);
  unsigned parity = 0;
#ifdef MIFGPU_PHASE_TRACE
  const int trace_cta = (job.trace && blockIdx.x % 148 == 0 && blockIdx.x < 296) ? (int)(blockIdx.x / 148) : -1;
  int n_traced = 0;
  if (trace_cta >= 0 && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    job.trace[(size_t)2 * 8 * kTraceTiles * kTraceSlots + trace_cta] = smid;
  }
#endif

  for (; tile < n_tiles; tile += gridDim.x) {
    const Tile t = decode_tile(job, tile);
    const int xt = t.xt, outer = t.outer, col = line ^ t.perm;  // col: this line's x offset inside the tile
    double2 v[16];

    // ---- phase A by the stage threads: inputs c[b + 32 s] of line l -------------------------------------------------
    MIF_TRACE(0);
    tma::mbar_wait(full, parity);
    parity ^= 1;
    MIF_TRACE(1);
    if (MODE == 1) {
      const double *N = S + l;  // natural rows: element e of line l at N[8 e]
      const double2 base = __ldg(&job.cs[b]);
#pragma unroll
      for (int s = 0; s < 16; s++) {
        const int k = b + 32 * s;
        const double2 rt = fft512::rot16(s);  // (cos, sin)(pi k / M) = cs[b] rotated by pi s / 16
        const double c = base.x * rt.x - base.y * rt.y, sn = base.x * rt.y + base.y * rt.x;
        v[s] = fft512::pack_input(N[8 * k], N[8 * (M - k)], c, sn);
      }
    } else {
      const double *Ev = S + l, *Od = S + Y::kOddRow0 * 8 + l;  // e(2q) at Ev[8 q], e(2q+1) at Od[8 q]
#pragma unroll
      for (int s = 0; s < 16; s++) {
        const int q = b + 32 * s;
        if (s < 8) v[s] = make_double2(Ev[8 * q], Od[8 * q]);
        else v[s] = make_double2(Ev[8 * (M - q)], Od[8 * (M - q - 1)]);  // mirror image: e(2M-2q), e(2M-2q-1)
      }
    }
    MIF_TRACE(2);
    fft512::phase_a(v, b, T);
    MIF_TRACE(3);
    if (tid == 0) tma::wait_stores_read();  // the previous tile's output stage (in W) has been read out
    __syncthreads();                        // the stage has been consumed, W is free
    MIF_TRACE(4);
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) issue_load(tile + gridDim.x);
    fft512::store_a(W + l * fft512::kLinePitch, b, v);
    MIF_TRACE(5);
    __syncthreads();
    MIF_TRACE(6);

    // ---- phase B by the line's warp ------------------------------------------------------------------------------------
    fft512::phase_b(Sline, j, v);
    MIF_TRACE(7);
    double spec[16], e_last = 0.0;
    if (MODE != 1) fft512::unpack_dct(v, j, job.cs, spec, e_last);
    MIF_TRACE(8);
    if (MODE == 2) {
      // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z); mode (0,0,0) := 0 (src/PressureEquation.cpp:158-163)
      const int ix = min(xt * kLines + col, job.n_lines - 1);
      const double lam_xy = job.lam_x[ix] + job.lam_y[outer];
      const bool origin_line = job.has_origin && (xt * kLines + col == 0) && (outer == 0);
      scale_by_inverse_eigenvalues(spec, lam_xy, job.lam_z, origin_line, [j](int r) { return fft512::k_of(j, r); });
      e_last *= 1.0 / (lam_xy + job.lam_z[M]);
      fft512::repack_for_inverse(spec, e_last, j, job.cs, v);
      fft512::phase_a(v, j, T);
      __syncwarp();  // every lane has finished reading the line region (phase B of the forward transform)
      fft512::store_a(Sline, j, v);
      __syncwarp();
      fft512::phase_b(Sline, j, v);
    }
    MIF_TRACE(9);
    __syncthreads();  // all warps are done with the work area: it becomes the output stage
    MIF_TRACE(10);

    // ---- results into the swizzled output stage (see tma_dct_kernel): row e of the tile at 64 e, column `line` ----------
    const int k2 = j & 15, p = j >> 4;
    if (MODE == 0) {
      // register r holds E_k, k = k2 + 128 p + 16 (r & 7) + 256 (r >> 3): the register part is a multiple of 1024 bytes
      const unsigned off = tma::swizzle_offset((unsigned)((k2 + 128 * p) * 64 + col * 8), job.swz);
#pragma unroll
      for (int r = 0; r < 16; r++)
        *reinterpret_cast<double *>(O + off + (16 * (r & 7) + 256 * (r >> 3)) * 64) = spec[r];
      if (j == 0) *reinterpret_cast<double *>(O + tma::swizzle_offset((unsigned)(M * 64 + col * 8), job.swz)) = e_last;
    } else {
      // register r < 8 holds z_q = conj(v[r]), q = k2 + 128 p + 16 r: x(2q) = Re z_q, x(2q+1) = Im z_q; q = 256 is register 8
      // of lane 0.  Lanes with bit 2 of k2 set store the odd row first (both halves of the 128-byte lines per instruction).
      const unsigned off = tma::swizzle_offset((unsigned)((k2 + 128 * p) * 128 + col * 8), job.swz);
      const bool odd_first = (k2 >> 2) & 1;
      const unsigned first = off + (odd_first ? 64u : 0u), second = first ^ 64u;
      const double scale = job.inv_norm;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const double even = v[r].x * scale, odd = -v[r].y * scale;
        *reinterpret_cast<double *>(O + first + r * 16 * 128) = odd_first ? odd : even;
        *reinterpret_cast<double *>(O + second + r * 16 * 128) = odd_first ? even : odd;
      }
      if (j == 0) *reinterpret_cast<double *>(O + tma::swizzle_offset((unsigned)(M * 64 + col * 8), job.swz)) = v[8].x * scale;
    }
    MIF_TRACE(11);
    tma::fence_proxy_async();
    __syncthreads();
    MIF_TRACE(12);
    if (tid == 0) store_tile<Y::kStoreBoxes>(job, &map_out, O, t);
#ifdef MIFGPU_PHASE_TRACE
    n_traced++;
#endif
  }
  if (tid == 0) tma::wait_stores_done();
}

// ---- 1025-point lines: radix-2 split into two 16 x 32 transforms -----------------------------------------------------------
// The 1024-point FFT of the packed even extension splits by one decimation-in-frequency step,
//     a_q = c_q + c_{q+512}  ->  C_{2k},        b_q = (c_q - c_{q+512}) W_1024^q  ->  C_{2k+1},      q, k < 512,
// into two independent 512-point transforms (mif_fft512.cuh), one warp each: 16 warps per tile of 8 lines, one CTA per
// SM.  Stage thread (l, b, h) forms the phase-A inputs of half h of line l (the pairs (c_q, c_{q+512}), q = b + 32 s, are
// read by both halves); warp (h, line) runs phase B and the unpack of its half: E_{2k} / E_{2k+1} with partners inside
// the warp (even half as for 513-point lines, odd half: C_{2k+1} pairs with C_{2(511-k)+1} in lane 31 - j).  In the
// fused z sweep the two warps of a line gather the scaled spectrum as plain reals in the line's first region (64-thread
// named barrier) and form the inputs of the second transform from there.
struct Layout1024 {
  static constexpr int M = 1024;
  static constexpr int kThreads = 512;
  static constexpr int kLoadBoxes = 2 * ((M / 2 + 1 + kBoxRows - 1) / kBoxRows);  // 8 boxes: 4 of even rows + 4 of odd rows, or 8 of natural rows
  static constexpr int kOddRow0 = (kLoadBoxes / 2) * kBoxRows;
  static constexpr int kStoreBoxes = (M + 1 + kStoreRows - 1) / kStoreRows;
  static constexpr size_t kWorkBytes = (size_t)2 * kLines * fft512::kLinePitch * sizeof(double2);
  static constexpr size_t kOutBytes = (size_t)kStoreBoxes * kStoreRows * 64;
  static constexpr size_t kTwBytes = (size_t)fft512::kTwiddles * sizeof(double2);
  static constexpr size_t kStageBytes = (size_t)kLoadBoxes * kBoxRows * 64;
  static constexpr size_t kWorkArea = ((kWorkBytes > kOutBytes ? kWorkBytes : kOutBytes) + 127) / 128 * 128;
  static constexpr size_t kSmem = 1024 + kWorkArea + (kTwBytes + 127) / 128 * 128 + kStageBytes + 64;
  static_assert(kLoadBoxes * kBoxRows >= M + 1, "natural rows fit the stage");
  static_assert(kSmem <= 227 * 1024, "one CTA per SM");
};

// 64-thread named barrier of the two warps of a line (immediate barrier numbers, see pair_sync in mif_poisson.cu).
__device__ __forceinline__ void line_pair_sync(int line) {
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

// Phase-A input of half h from the pair (lo, hi) = (c_q, c_{q+512}); wq = W_1024^q.
__device__ __forceinline__ double2 split_input(double2 lo, double2 hi, int h, double2 wq) {
  return h ? fft512::cmul(fft512::csub(lo, hi), wq) : fft512::cadd(lo, hi);
}

template <int MODE>
__global__ void __launch_bounds__(Layout1024::kThreads, 1)
    tma_dct1024_kernel(const __grid_constant__ CUtensorMap map_in_a, const __grid_constant__ CUtensorMap map_in_b,
                       const __grid_constant__ CUtensorMap map_out, const Job job) {
  using Y = Layout1024;
  constexpr int M = 1024, H = 512;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = smem_raw + ((1024 - (tma::swizzle_address(smem_raw) & 1023)) & 1023);
  double2 *W = reinterpret_cast<double2 *>(smem);                                   // region (h, line) at (8 h + line) * kLinePitch
  unsigned char *O = smem;                                                          // output stage (aliases W)
  double2 *T = reinterpret_cast<double2 *>(smem + Y::kWorkArea);                    // W_512^(n1 2^e)
  double *S = reinterpret_cast<double *>(smem + Y::kWorkArea + (Y::kTwBytes + 127) / 128 * 128);  // input stage
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(S) + Y::kStageBytes);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, j = tid & 31;
  const int line = warp & 7, half = warp >> 3;             // transform mapping: warp = (half, line), lane j = k2 + 16 p
  const int l = tid & 7, b = (tid >> 3) & 31, h = tid >> 8;  // stage mapping: line fastest, b = n1, h = half (warp uniform)
  double2 *Sline = W + (8 * half + line) * fft512::kLinePitch;
  double *Rl = reinterpret_cast<double *>(W + line * fft512::kLinePitch);  // the line's 1025 reals (first region of the pair)
  const int n_tiles = job.n_xtiles * job.n_outer;

  auto issue_load = [&](int tile) {
    const Tile nt = decode_tile(job, tile);
    tma::mbar_arrive_expect_tx(full, (unsigned)Y::kStageBytes);
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < Y::kLoadBoxes; i++) tma::load_3d(S + i * kBoxRows * 8, &map_in_a, full, nt.x_in, i * kBoxRows, nt.c2);
    } else {
#pragma unroll
      for (int i = 0; i < Y::kLoadBoxes / 2; i++) {
        tma::load_3d(S + i * kBoxRows * 8, &map_in_a, full, nt.x_in, i * kBoxRows, nt.c2);
        tma::load_3d(S + (Y::kOddRow0 + i * kBoxRows) * 8, &map_in_b, full, nt.x_in, i * kBoxRows, nt.c2);
      }
    }
  };
  // conj Z_k and conj Z_{k+512} of a real spectrum E given as a function, combined to the input of half hh; cs[k] =
  // (cos, sin)(pi k / 1024) comes as cs[n1] rotated by pi s / 32, and cs[k + 512] is cs[k] rotated by pi / 2
  auto inverse_input = [&](auto E, int n1, int s, int hh, double2 cs_n1, double2 w_n1) {
    const int k = n1 + 32 * s;
    const double2 rt = fft512::rot32x(s);
    const double c = cs_n1.x * rt.x - cs_n1.y * rt.y, sn = cs_n1.x * rt.y + cs_n1.y * rt.x;
    const double2 lo = fft512::pack_input(E(k), E(M - k), c, sn);
    const double2 hi = fft512::pack_input(E(k + H), E(H - k), -sn, c);
    return split_input(lo, hi, hh, fft512::cmul(w_n1, fft512::w32(s)));
  };

  if (tid == 0) {
    tma::prefetch_map(&map_in_a);
    tma::prefetch_map(&map_in_b);
    tma::prefetch_map(&map_out);
    tma::mbar_init(full, 1);
    tma::fence_barrier_init();
    tma::fence_proxy_async();
  }
  fft512::load_twiddles<Y::kThreads, 2>(T, job.tw);
  __syncthreads();
  int tile = blockIdx.x;
  if (tid == 0 && tile < n_tiles) issue_load(tile);
  unsigned parity = 0;

  for (; tile < n_tiles; tile += gridDim.x) {
    const Tile t = decode_tile(job, tile);
    const int xt = t.xt, outer = t.outer, col = line ^ t.perm;  // col: this line's x offset inside the tile
    double2 v[16];

    // ---- phase A by the stage threads: inputs of half h of line l -------------------------------------------------------
    tma::mbar_wait(full, parity);
    parity ^= 1;
    {
      const double2 wb = __ldg(&job.tw[b]);  // W_1024^b;  W_1024^(b + 32 s) = W_1024^b W_32^s
      if (MODE == 1) {
        const double *N = S + l;  // natural rows: element e of line l at N[8 e]
        const double2 cs_b = __ldg(&job.cs[b]);
#pragma unroll
        for (int s = 0; s < 16; s++) v[s] = inverse_input([&](int e) { return N[8 * e]; }, b, s, h, cs_b, wb);
      } else {
        const double *Ev = S + l, *Od = S + Y::kOddRow0 * 8 + l;  // e(2q) at Ev[8 q], e(2q+1) at Od[8 q]
#pragma unroll
        for (int s = 0; s < 16; s++) {
          const int q = b + 32 * s;
          const double2 lo = make_double2(Ev[8 * q], Od[8 * q]);                    // c_q
          const double2 hi = make_double2(Ev[8 * (H - q)], Od[8 * (H - q - 1)]);    // c_{q+512} = (e(2M-2q-1024), e(2M-2q-1025))
          v[s] = split_input(lo, hi, h, fft512::cmul(wb, fft512::w32(s)));
        }
      }
    }
    fft512::phase_a(v, b, T);
    if (tid == 0) tma::wait_stores_read();  // the previous tile's output stage (in W) has been read out
    __syncthreads();                        // the stage has been consumed, W is free
    if (tid == 0 && tile + (int)gridDim.x < n_tiles) issue_load(tile + gridDim.x);
    fft512::store_a(W + (8 * h + l) * fft512::kLinePitch, b, v);
    __syncthreads();

    // ---- phase B by warp (half, line) ----------------------------------------------------------------------------------
    fft512::phase_b(Sline, j, v);
    double spec[16], e_last = 0.0;  // spec[r] = E_(2 k + half), k = k_of(j, r)
    if (MODE != 1) {
      if (half) fft512::unpack_dct<2>(v, j, job.cs, spec, e_last);
      else fft512::unpack_dct<1>(v, j, job.cs, spec, e_last);
    }
    if (MODE == 2) {
      // pressure_hat *= 1 / (lambda_x + lambda_y + lambda_z); mode (0,0,0) := 0 (src/PressureEquation.cpp:158-163)
      const int ix = min(xt * kLines + col, job.n_lines - 1);
      const double lam_xy = job.lam_x[ix] + job.lam_y[outer];
      const bool origin_line = job.has_origin && (xt * kLines + col == 0) && (outer == 0);
      scale_by_inverse_eigenvalues(spec, lam_xy, job.lam_z, origin_line, [j, half](int r) { return 2 * fft512::k_of(j, r) + half; });
      e_last *= 1.0 / (lam_xy + job.lam_z[M]);
      line_pair_sync(line);  // both warps of the line are done with their regions
#pragma unroll
      for (int r = 0; r < 16; r++) Rl[2 * fft512::k_of(j, r) + half] = spec[r];
      if (half == 0 && j == 0) Rl[M] = e_last;
      line_pair_sync(line);
      {
        const double2 wj = __ldg(&job.tw[j]), cs_j = __ldg(&job.cs[j]);
#pragma unroll
        for (int s = 0; s < 16; s++) v[s] = inverse_input([&](int e) { return Rl[e]; }, j, s, half, cs_j, wj);
      }
      fft512::phase_a(v, j, T);
      line_pair_sync(line);  // the whole line has been read before the regions (which alias it) are overwritten
      fft512::store_a(Sline, j, v);
      __syncwarp();
      fft512::phase_b(Sline, j, v);
    }
    __syncthreads();  // all warps are done with the work area: it becomes the output stage

    // ---- results into the swizzled output stage: row e of the tile at 64 e, column col ----------------------------------
    // (all rows of one warp have the parity of its half, i.e. sit in the same half of their 128-byte lines: 4 bank
    // positions per store instruction instead of the 8 of the 513-point kernels)
    const int k2 = j & 15, p = j >> 4;
    if (MODE == 0) {
      // register r holds E_e, e = 2 (k2 + 128 p + 16 (r & 7) + 256 (r >> 3)) + half
      const unsigned off = tma::swizzle_offset((unsigned)((2 * (k2 + 128 * p) + half) * 64 + col * 8), job.swz);
#pragma unroll
      for (int r = 0; r < 16; r++)
        *reinterpret_cast<double *>(O + off + (32 * (r & 7) + 512 * (r >> 3)) * 64) = spec[r];
      if (half == 0 && j == 0) *reinterpret_cast<double *>(O + tma::swizzle_offset((unsigned)(M * 64 + col * 8), job.swz)) = e_last;
    } else {
      // register r < 8 holds z_q = conj(v[r]), q = 2 (k2 + 128 p + 16 r) + half: x(2q) = Re z_q, x(2q+1) = Im z_q; q = 512
      // is register 8 of lane 0 of the even half
      const unsigned off = tma::swizzle_offset((unsigned)((2 * (k2 + 128 * p) + half) * 128 + col * 8), job.swz);
      const bool odd_first = (k2 >> 1) & 1;
      const unsigned first = off + (odd_first ? 64u : 0u), second = first ^ 64u;
      const double scale = job.inv_norm;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        const double even = v[r].x * scale, odd = -v[r].y * scale;
        *reinterpret_cast<double *>(O + first + r * 32 * 128) = odd_first ? odd : even;
        *reinterpret_cast<double *>(O + second + r * 32 * 128) = odd_first ? even : odd;
      }
      if (half == 0 && j == 0) *reinterpret_cast<double *>(O + tma::swizzle_offset((unsigned)(M * 64 + col * 8), job.swz)) = v[8].x * scale;
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) store_tile<Y::kStoreBoxes>(job, &map_out, O, t);
  }
  if (tid == 0) tma::wait_stores_done();
}

// ---- host -----------------------------------------------------------------------------------------------------
// Where the tiles of a sweep live: element (x, row, c2) at base + x + row * estride + c2 * outer_stride (doubles).
struct TensorDesc {
  const double *base;
  int dim0;                    // valid x columns (coordinate 0 beyond it: zero fill on load, clipped on store)
  long long estride, outer_stride;
  long long outer_extent;      // extent of coordinate 2
  bool operator<(const TensorDesc &o) const {
    return std::tie(base, dim0, estride, outer_stride, outer_extent) < std::tie(o.base, o.dim0, o.estride, o.outer_stride, o.outer_extent);
  }
};
struct MapKey {
  TensorDesc in, out;
  int n, kind, promo;
  bool operator<(const MapKey &o) const { return std::tie(in, out, n, kind, promo) < std::tie(o.in, o.out, o.n, o.kind, o.promo); }
};
struct MapSet {
  CUtensorMap even, odd, natural, out;
  bool ok;
};

struct Cache {
  std::map<MapKey, MapSet> sets;
  int device = -1, sms = 0;
  bool attr[2][3] = {};
  bool attr512[3] = {};
  bool attr1024[3] = {};
};

// Tensor maps of one sweep: even rows / odd rows / natural rows of the input (boxes of 8 x kBoxRows), natural rows of the
// output (boxes of 8 x kStoreRows, swizzled if the output stage is).  n = points per line.
inline const MapSet *maps_for(Cache &cache, const TensorDesc &in, const TensorDesc &out, int n, bool swizzle, int promo) {
  const MapKey key{in, out, n, swizzle ? 1 : 0, promo};
  auto it = cache.sets.find(key);
  if (it != cache.sets.end()) return it->second.ok ? &it->second : nullptr;
  MapSet set;
  const uint64_t row = (uint64_t)in.estride * 8, outer_b = (uint64_t)in.outer_stride * 8;
  const int half = (n - 1) / 2;
  const char *why = nullptr;
  set.ok = tma::encode_map(&set.even, in.base, (uint64_t)in.dim0, (uint64_t)half + 1, (uint64_t)in.outer_extent, 2 * row, outer_b, kLines, kBoxRows, false, promo, &why) &&
           tma::encode_map(&set.odd, in.base + in.estride, (uint64_t)in.dim0, (uint64_t)half, (uint64_t)in.outer_extent, 2 * row, outer_b, kLines, kBoxRows, false, promo, &why) &&
           tma::encode_map(&set.natural, in.base, (uint64_t)in.dim0, (uint64_t)n, (uint64_t)in.outer_extent, row, outer_b, kLines, kBoxRows, false, promo, &why) &&
           tma::encode_map(&set.out, out.base, (uint64_t)out.dim0, (uint64_t)n, (uint64_t)out.outer_extent, (uint64_t)out.estride * 8,
                           (uint64_t)out.outer_stride * 8, kLines, kStoreRows, swizzle, 0, &why);
  if (!set.ok && getenv("MIFGPU_TMA_VERBOSE")) fprintf(stderr, "libmifgpu: no tensor map for this sweep: %s\n", why ? why : "?");
  auto ins = cache.sets.emplace(key, set);
  return set.ok ? &ins.first->second : nullptr;
}

template <int LOGM, int MODE>
void launch_one(cudaStream_t stream, Cache &cache, const MapSet &maps, const Job &job) {
  using Y = Layout<LOGM>;
  bool &attr = cache.attr[LOGM - 8][MODE];
  if (!attr) {
    cudaFuncSetAttribute(tma_dct_kernel<LOGM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y::kSmem);
    attr = true;
  }
  static const int per_sm = getenv("MIFGPU_TMA_CTAS_PER_SM") ? atoi(getenv("MIFGPU_TMA_CTAS_PER_SM")) : 2;
  const int n_tiles = job.n_xtiles * job.n_outer;
  const int grid = std::min(n_tiles, std::max(1, cache.sms * per_sm));
  tma_dct_kernel<LOGM, MODE><<<grid, kThreads, Y::kSmem, stream>>>(MODE == 1 ? maps.natural : maps.even, maps.odd, maps.out, job);
}

template <int MODE>
void launch_512(cudaStream_t stream, Cache &cache, const MapSet &maps, const Job &job) {
  using Y = Layout512;
  if (!cache.attr512[MODE]) {
    cudaFuncSetAttribute(tma_dct512_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y::kSmem);
    cache.attr512[MODE] = true;
  }
  static const int per_sm = getenv("MIFGPU_TMA_CTAS_PER_SM") ? atoi(getenv("MIFGPU_TMA_CTAS_PER_SM")) : 2;
  const int n_tiles = job.n_xtiles * job.n_outer;
  const int grid = std::min(n_tiles, std::max(1, cache.sms * per_sm));
#ifdef MIFGPU_PHASE_TRACE
  // diagnostic build: the launch number MIFGPU_TRACE_LAUNCH of this mode writes its stamps to MIFGPU_TRACE_FILE
  static int launch_no[3] = {0, 0, 0};
  static unsigned long long *trace_dev = nullptr;
  const size_t words = (size_t)2 * 8 * kTraceTiles * kTraceSlots + 2;
  Job traced = job;
  traced.trace = nullptr;
  const char *file = getenv("MIFGPU_TRACE_FILE");
  const int want_mode = getenv("MIFGPU_TRACE_MODE") ? atoi(getenv("MIFGPU_TRACE_MODE")) : 0;
  const int want_launch = getenv("MIFGPU_TRACE_LAUNCH") ? atoi(getenv("MIFGPU_TRACE_LAUNCH")) : 12;
  const bool dump = file && MODE == want_mode && launch_no[MODE]++ == want_launch;
  if (dump) {
    if (!trace_dev) cudaMalloc(&trace_dev, words * 8);
    cudaMemsetAsync(trace_dev, 0, words * 8, stream);
    traced.trace = trace_dev;
  }
  tma_dct512_kernel<MODE><<<grid, kThreads, Y::kSmem, stream>>>(MODE == 1 ? maps.natural : maps.even, maps.odd, maps.out, traced);
  if (dump) {
    std::vector<unsigned long long> host(words);
    cudaStreamSynchronize(stream);
    cudaMemcpy(host.data(), trace_dev, words * 8, cudaMemcpyDeviceToHost);
    if (FILE *f = fopen(file, "wb")) {
      fwrite(host.data(), 8, words, f);
      fclose(f);
    }
  }
#else
  tma_dct512_kernel<MODE><<<grid, kThreads, Y::kSmem, stream>>>(MODE == 1 ? maps.natural : maps.even, maps.odd, maps.out, job);
#endif
}

template <int MODE>
void launch_1024(cudaStream_t stream, Cache &cache, const MapSet &maps, const Job &job) {
  using Y = Layout1024;
  if (!cache.attr1024[MODE]) {
    cudaFuncSetAttribute(tma_dct1024_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y::kSmem);
    cache.attr1024[MODE] = true;
  }
  const int n_tiles = job.n_xtiles * job.n_outer;
  const int grid = std::min(n_tiles, std::max(1, cache.sms));
  tma_dct1024_kernel<MODE><<<grid, Y::kThreads, Y::kSmem, stream>>>(MODE == 1 ? maps.natural : maps.even, maps.odd, maps.out, job);
}

}  // namespace tmasweep
}  // namespace mifgpu
