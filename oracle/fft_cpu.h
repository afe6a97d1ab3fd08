/* oracle/fft_cpu.h -- TEST INFRASTRUCTURE ONLY (CPU checker / CPU baseline; never linked into the product).
 *
 * Plain-C restatement of the three 1-D real transforms the reference asks FFTW for
 * (/root/reference/src/PressureSolverStructures.cpp:23-41, executed at
 * /root/reference/src/PressureEquation.cpp:85,112,139,174,207,241):
 *
 *   FFTW_REDFT00 (DCT-I), n points:  Y[k] = X[0] + (-1)^k X[n-1] + 2 sum_{j=1}^{n-2} X[j] cos(pi j k/(n-1))
 *   FFTW_R2HC, n points:             C[k] = sum_j X[j] exp(-2 pi i j k/n); output r0 r1 .. r_{n/2} i_{(n+1)/2-1} .. i1
 *   FFTW_HC2R, n points:             unnormalised inverse of R2HC (HC2R(R2HC(x)) = n x)
 *
 * These are FFTW's documented definitions (FFTW 3 manual, "1d Real-even DFTs" / "The Halfcomplex-format
 * DFT"); FFTW itself is a third-party dependency that is absent from /root/reference and from this
 * image (no version is pinned by the reference: CMakeLists.txt:13-19, Makefile:53-57), so the transforms
 * are pinned by definition and cross-checked in tests/ against scipy.fft.dct(type=1) and numpy.fft.rfft.
 *
 * Two implementations are provided for each transform: a direct O(n^2) sum that reads exactly like the
 * definition (mo_*_direct, the oracle of the oracle), and an O(n log n) version built on one complex FFT
 * (radix-4 Stockham passes for powers of two, Bluestein's chirp-z otherwise) that is used by the fftw3.h shim so the
 * compiled reference doubles as the CPU timing baseline.
 */
#ifndef MIF_ORACLE_FFT_CPU_H
#define MIF_ORACLE_FFT_CPU_H

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef MO_PI
#define MO_PI 3.14159265358979323846264338327950288
#endif

/* ---------------------------------------------------------------------------------------------- */
/* Direct O(n^2) definitions.                                                                      */
/* ---------------------------------------------------------------------------------------------- */

static inline void mo_redft00_direct(int n, const double *x, double *y) {
  /* y may not alias x */
  for (int k = 0; k < n; k++) {
    double acc = x[0] + ((k & 1) ? -x[n - 1] : x[n - 1]);
    for (int j = 1; j < n - 1; j++) {
      /* reduce j*k modulo the period 2(n-1) in integers so the cosine argument stays accurate */
      const long long r = ((long long)j * k) % (2LL * (n - 1));
      acc += 2.0 * x[j] * cos(MO_PI * (double)r / (double)(n - 1));
    }
    y[k] = acc;
  }
}

static inline void mo_r2hc_direct(int n, const double *x, double *y) {
  for (int k = 0; k <= n / 2; k++) {
    double re = 0.0, im = 0.0;
    for (int j = 0; j < n; j++) {
      const long long r = ((long long)j * k) % n;
      const double a = 2.0 * MO_PI * (double)r / (double)n;
      re += x[j] * cos(a);
      im -= x[j] * sin(a);
    }
    y[k] = re;
    if (k > 0 && k < n - k) y[n - k] = im;
  }
}

static inline void mo_hc2r_direct(int n, const double *y, double *x) {
  for (int j = 0; j < n; j++) {
    double acc = y[0];
    for (int k = 1; k < n - k; k++) {
      const long long r = ((long long)j * k) % n;
      const double a = 2.0 * MO_PI * (double)r / (double)n;
      acc += 2.0 * (y[k] * cos(a) - y[n - k] * sin(a));
    }
    if (n % 2 == 0) acc += (j & 1) ? -y[n / 2] : y[n / 2];
    x[j] = acc;
  }
}

/* ---------------------------------------------------------------------------------------------- */
/* Complex FFT of arbitrary length (interleaved re,im).                                            */
/* ---------------------------------------------------------------------------------------------- */

typedef struct mo_cfft_plan {
  int n;
  int is_pow2;
  /* power-of-two path: radix-4 Stockham autosort passes (one radix-2 pass last when log2 n is odd) */
  double *tw;   /* per radix-4 pass of length len = n, n/4, ...: 3 * len/4 complex twiddles w^p, w^2p, w^3p (forward) */
  double *ping; /* n complex scratch */
  /* Bluestein path */
  int m;
  struct mo_cfft_plan *sub;
  double *chirp; /* w_j = exp(-i pi j^2 / n), j < n */
  double *filt;  /* FFT_m of the wrapped conjugate chirp */
  double *work;  /* m complex scratch */
} mo_cfft_plan;

static inline mo_cfft_plan *mo_cfft_create(int n);
static inline void mo_cfft_destroy(mo_cfft_plan *p);

/* One radix-4 decimation-in-frequency Stockham pass: the sequence of length len (stride s, n = len * s) is split
 * into four interleaved sub-sequences of length len/4 (stride 4 s).  sg = +1 forward, -1 inverse. */
static inline void mo_stockham4_pass(int len, int s, const double *tw, double sg, const double *__restrict__ x,
                                     double *__restrict__ y) {
  const int n1 = len / 4;
  if (s == 1) {
    for (int p = 0; p < n1; p++) {
      const double w1r = tw[6 * p], w1i = sg * tw[6 * p + 1];
      const double w2r = tw[6 * p + 2], w2i = sg * tw[6 * p + 3];
      const double w3r = tw[6 * p + 4], w3i = sg * tw[6 * p + 5];
      const double ar = x[2 * p], ai = x[2 * p + 1];
      const double br = x[2 * (p + n1)], bi = x[2 * (p + n1) + 1];
      const double cr = x[2 * (p + 2 * n1)], ci = x[2 * (p + 2 * n1) + 1];
      const double dr = x[2 * (p + 3 * n1)], di = x[2 * (p + 3 * n1) + 1];
      const double apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
      const double bpdr = br + dr, bpdi = bi + di;
      const double jr = -sg * (bi - di), ji = sg * (br - dr); /* sg * i * (b - d) */
      double *o = y + 8 * p;
      o[0] = apcr + bpdr;
      o[1] = apci + bpdi;
      const double t1r = amcr - jr, t1i = amci - ji;
      o[2] = t1r * w1r - t1i * w1i;
      o[3] = t1r * w1i + t1i * w1r;
      const double t2r = apcr - bpdr, t2i = apci - bpdi;
      o[4] = t2r * w2r - t2i * w2i;
      o[5] = t2r * w2i + t2i * w2r;
      const double t3r = amcr + jr, t3i = amci + ji;
      o[6] = t3r * w3r - t3i * w3i;
      o[7] = t3r * w3i + t3i * w3r;
    }
    return;
  }
  for (int p = 0; p < n1; p++) {
    const double w1r = tw[6 * p], w1i = sg * tw[6 * p + 1];
    const double w2r = tw[6 * p + 2], w2i = sg * tw[6 * p + 3];
    const double w3r = tw[6 * p + 4], w3i = sg * tw[6 * p + 5];
    const double *xa = x + 2 * (size_t)s * p, *xb = xa + 2 * (size_t)s * n1, *xc = xb + 2 * (size_t)s * n1,
                 *xd = xc + 2 * (size_t)s * n1;
    double *y0 = y + 2 * (size_t)s * (4 * p), *y1 = y0 + 2 * (size_t)s, *y2 = y1 + 2 * (size_t)s, *y3 = y2 + 2 * (size_t)s;
    for (int q = 0; q < s; q++) {
      const double ar = xa[2 * q], ai = xa[2 * q + 1], br = xb[2 * q], bi = xb[2 * q + 1];
      const double cr = xc[2 * q], ci = xc[2 * q + 1], dr = xd[2 * q], di = xd[2 * q + 1];
      const double apcr = ar + cr, apci = ai + ci, amcr = ar - cr, amci = ai - ci;
      const double bpdr = br + dr, bpdi = bi + di;
      const double jr = -sg * (bi - di), ji = sg * (br - dr);
      y0[2 * q] = apcr + bpdr;
      y0[2 * q + 1] = apci + bpdi;
      const double t1r = amcr - jr, t1i = amci - ji;
      y1[2 * q] = t1r * w1r - t1i * w1i;
      y1[2 * q + 1] = t1r * w1i + t1i * w1r;
      const double t2r = apcr - bpdr, t2i = apci - bpdi;
      y2[2 * q] = t2r * w2r - t2i * w2i;
      y2[2 * q + 1] = t2r * w2i + t2i * w2r;
      const double t3r = amcr + jr, t3i = amci + ji;
      y3[2 * q] = t3r * w3r - t3i * w3i;
      y3[2 * q + 1] = t3r * w3i + t3i * w3r;
    }
  }
}

static inline void mo_cfft_pow2_exec(const mo_cfft_plan *p, double *z, int inverse) {
  const int n = p->n;
  const double sg = inverse ? -1.0 : 1.0;
  double *x = z, *y = p->ping;
  const double *tw = p->tw;
  int len = n, s = 1;
  while (len >= 4) {
    mo_stockham4_pass(len, s, tw, sg, x, y);
    tw += 6 * (len / 4);
    len /= 4;
    s *= 4;
    double *t = x;
    x = y;
    y = t;
  }
  if (len == 2) { /* last pass: n/2 butterflies without twiddles, stride s = n/2 */
    for (int q = 0; q < s; q++) {
      const double ar = x[2 * q], ai = x[2 * q + 1], br = x[2 * (q + s)], bi = x[2 * (q + s) + 1];
      y[2 * q] = ar + br;
      y[2 * q + 1] = ai + bi;
      y[2 * (q + s)] = ar - br;
      y[2 * (q + s) + 1] = ai - bi;
    }
    double *t = x;
    x = y;
    y = t;
  }
  if (x != z) memcpy(z, x, sizeof(double) * 2 * (size_t)n);
}

/* Forward (inverse=0: exp(-2 pi i jk/n)) or unnormalised inverse (inverse=1) DFT, in place. */
static inline void mo_cfft_exec(const mo_cfft_plan *p, double *z, int inverse) {
  const int n = p->n;
  if (n <= 1) return;
  if (p->is_pow2) {
    mo_cfft_pow2_exec(p, z, inverse);
    return;
  }
  /* Bluestein: X_k = w_k sum_j (x_j w_j) conj(w_{k-j}); the inverse transform is conj(DFT(conj x)). */
  const int m = p->m;
  double *a = p->work;
  const double cs = inverse ? -1.0 : 1.0;
  for (int j = 0; j < n; j++) {
    const double xr = z[2 * j], xi = cs * z[2 * j + 1];
    const double wr = p->chirp[2 * j], wi = p->chirp[2 * j + 1];
    a[2 * j] = xr * wr - xi * wi;
    a[2 * j + 1] = xr * wi + xi * wr;
  }
  memset(a + 2 * n, 0, sizeof(double) * 2 * (size_t)(m - n));
  mo_cfft_pow2_exec(p->sub, a, 0);
  for (int k = 0; k < m; k++) {
    const double ar = a[2 * k], ai = a[2 * k + 1];
    const double fr = p->filt[2 * k], fi = p->filt[2 * k + 1];
    a[2 * k] = ar * fr - ai * fi;
    a[2 * k + 1] = ar * fi + ai * fr;
  }
  mo_cfft_pow2_exec(p->sub, a, 1);
  const double inv_m = 1.0 / (double)m;
  for (int k = 0; k < n; k++) {
    const double ar = a[2 * k] * inv_m, ai = a[2 * k + 1] * inv_m;
    const double wr = p->chirp[2 * k], wi = p->chirp[2 * k + 1];
    z[2 * k] = ar * wr - ai * wi;
    z[2 * k + 1] = cs * (ar * wi + ai * wr);
  }
}

static inline mo_cfft_plan *mo_cfft_create(int n) {
  mo_cfft_plan *p = (mo_cfft_plan *)calloc(1, sizeof(mo_cfft_plan));
  p->n = n;
  p->is_pow2 = (n > 0) && ((n & (n - 1)) == 0);
  if (n <= 1) return p;
  if (p->is_pow2) {
    p->tw = (double *)malloc(sizeof(double) * 2 * (size_t)n + 64);
    p->ping = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    double *tw = p->tw;
    for (int len = n; len >= 4; len /= 4) {
      for (int q = 0; q < len / 4; q++)
        for (int r = 1; r <= 3; r++) {
          const double a = -2.0 * MO_PI * (double)(q * r) / (double)len;
          *tw++ = cos(a);
          *tw++ = sin(a);
        }
    }
    return p;
  }
  int m = 1;
  while (m < 2 * n - 1) m <<= 1;
  p->m = m;
  p->sub = mo_cfft_create(m);
  p->chirp = (double *)malloc(sizeof(double) * 2 * (size_t)n);
  p->filt = (double *)calloc(2 * (size_t)m, sizeof(double));
  p->work = (double *)malloc(sizeof(double) * 2 * (size_t)m);
  for (int j = 0; j < n; j++) {
    const long long r = ((long long)j * j) % (2LL * n);
    const double a = -MO_PI * (double)r / (double)n;
    p->chirp[2 * j] = cos(a);
    p->chirp[2 * j + 1] = sin(a);
  }
  for (int j = 0; j < n; j++) {
    p->filt[2 * j] = p->chirp[2 * j];
    p->filt[2 * j + 1] = -p->chirp[2 * j + 1];
    if (j > 0) {
      p->filt[2 * (m - j)] = p->chirp[2 * j];
      p->filt[2 * (m - j) + 1] = -p->chirp[2 * j + 1];
    }
  }
  mo_cfft_pow2_exec(p->sub, p->filt, 0);
  return p;
}

static inline void mo_cfft_destroy(mo_cfft_plan *p) {
  if (!p) return;
  free(p->tw);
  free(p->ping);
  free(p->chirp);
  free(p->filt);
  free(p->work);
  mo_cfft_destroy(p->sub);
  free(p);
}

/* ---------------------------------------------------------------------------------------------- */
/* Fast real transforms built on one complex FFT.                                                  */
/* ---------------------------------------------------------------------------------------------- */

enum { MO_REDFT00 = 0, MO_R2HC = 1, MO_HC2R = 2 };

typedef struct mo_r2r_plan {
  int n;    /* number of real points */
  int kind; /* MO_REDFT00 / MO_R2HC / MO_HC2R */
  int nc;   /* length of the complex FFT */
  mo_cfft_plan *c;
  double *cs, *sn; /* unpack twiddles */
  double *z;       /* nc complex scratch */
} mo_r2r_plan;

static inline mo_r2r_plan *mo_r2r_create(int n, int kind) {
  mo_r2r_plan *p = (mo_r2r_plan *)calloc(1, sizeof(mo_r2r_plan));
  p->n = n;
  p->kind = kind;
  if (kind == MO_REDFT00) {
    /* Even extension of period 2(n-1), packed two reals per complex: FFT length n-1. */
    p->nc = n - 1;
    p->cs = (double *)malloc(sizeof(double) * (size_t)n);
    p->sn = (double *)malloc(sizeof(double) * (size_t)n);
    for (int k = 0; k < n; k++) {
      p->cs[k] = cos(MO_PI * (double)k / (double)(n - 1));
      p->sn[k] = sin(MO_PI * (double)k / (double)(n - 1));
    }
  } else if (n % 2 == 0) {
    p->nc = n / 2;
    p->cs = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1));
    p->sn = (double *)malloc(sizeof(double) * (size_t)(n / 2 + 1));
    for (int k = 0; k <= n / 2; k++) {
      p->cs[k] = cos(2.0 * MO_PI * (double)k / (double)n);
      p->sn[k] = sin(2.0 * MO_PI * (double)k / (double)n);
    }
  } else {
    p->nc = n;
  }
  p->c = mo_cfft_create(p->nc);
  p->z = (double *)malloc(sizeof(double) * 2 * (size_t)(p->nc > 0 ? p->nc : 1));
  return p;
}

static inline void mo_r2r_destroy(mo_r2r_plan *p) {
  if (!p) return;
  mo_cfft_destroy(p->c);
  free(p->cs);
  free(p->sn);
  free(p->z);
  free(p);
}

/* in and out may alias (the reference transforms in place, src/PressureEquation.cpp:84-85). */
static inline void mo_r2r_exec(const mo_r2r_plan *p, const double *in, double *out) {
  const int n = p->n;
  double *z = p->z;
  if (p->kind == MO_REDFT00) {
    const int m = n - 1; /* complex length; even extension e has period 2m */
    if (m == 0) {
      out[0] = in[0];
      return;
    }
    for (int j = 0; j < m; j++) {
      const int i0 = 2 * j, i1 = 2 * j + 1;
      z[2 * j] = in[i0 <= m ? i0 : 2 * m - i0];
      z[2 * j + 1] = in[i1 <= m ? i1 : 2 * m - i1];
    }
    mo_cfft_exec(p->c, z, 0);
    for (int k = 0; k <= m; k++) {
      const int k0 = k % m, k1 = (m - k) % m;
      const double a = z[2 * k0], b = z[2 * k0 + 1], c = z[2 * k1], d = z[2 * k1 + 1];
      out[k] = 0.5 * ((a + c) + p->cs[k] * (b + d) - p->sn[k] * (a - c));
    }
    return;
  }
  if (p->kind == MO_R2HC) {
    if (n % 2 == 0) {
      const int h = n / 2;
      for (int j = 0; j < h; j++) {
        z[2 * j] = in[2 * j];
        z[2 * j + 1] = in[2 * j + 1];
      }
      mo_cfft_exec(p->c, z, 0);
      for (int k = 0; k <= h; k++) {
        const int k0 = k % h, k1 = (h - k) % h;
        const double a = z[2 * k0], b = z[2 * k0 + 1], c = z[2 * k1], d = z[2 * k1 + 1];
        const double re = 0.5 * ((a + c) + p->cs[k] * (b + d) - p->sn[k] * (a - c));
        const double im = 0.5 * ((b - d) - p->cs[k] * (a - c) - p->sn[k] * (b + d));
        out[k] = re; /* all of `in` was consumed into z above, so in == out is safe */
        if (k > 0 && k < h) out[n - k] = im;
      }
    } else {
      for (int j = 0; j < n; j++) {
        z[2 * j] = in[j];
        z[2 * j + 1] = 0.0;
      }
      mo_cfft_exec(p->c, z, 0);
      for (int k = 0; k <= n / 2; k++) {
        out[k] = z[2 * k];
        if (k > 0) out[n - k] = z[2 * k + 1];
      }
    }
    return;
  }
  /* MO_HC2R */
  if (n % 2 == 0) {
    const int h = n / 2;
    /* X_k for k = 0..h from the halfcomplex input, then C'_k = P_k + Q_k with
     *   P_k = X_k + conj X_{h-k},  Q_k = i conj(w^k) (X_k - conj X_{h-k}),  w = exp(-2 pi i/n). */
    for (int k = 0; k < h; k++) {
      const int kk = h - k;
      const double xr = in[k], xi = (k == 0) ? 0.0 : in[n - k];
      const double yr = in[kk], yi = (kk == h) ? 0.0 : in[n - kk];
      const double pr = xr + yr, pi_ = xi - yi;
      const double dr = xr - yr, di = xi + yi;
      /* i conj(w^k) = i (cos + i sin) = -sin + i cos */
      const double qr = -p->sn[k] * dr - p->cs[k] * di;
      const double qi = p->cs[k] * dr - p->sn[k] * di;
      z[2 * k] = pr + qr;
      z[2 * k + 1] = pi_ + qi;
    }
    mo_cfft_exec(p->c, z, 1);
    for (int j = 0; j < h; j++) {
      out[2 * j] = z[2 * j];
      out[2 * j + 1] = z[2 * j + 1];
    }
  } else {
    z[0] = in[0];
    z[1] = 0.0;
    for (int k = 1; k <= n / 2; k++) {
      z[2 * k] = in[k];
      z[2 * k + 1] = in[n - k];
      z[2 * (n - k)] = in[k];
      z[2 * (n - k) + 1] = -in[n - k];
    }
    mo_cfft_exec(p->c, z, 1);
    for (int j = 0; j < n; j++) out[j] = z[2 * j];
  }
}

#ifdef __cplusplus
}
#endif

#endif /* MIF_ORACLE_FFT_CPU_H */
