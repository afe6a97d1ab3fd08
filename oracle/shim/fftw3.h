// oracle/shim/fftw3.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Stand-in for the FFTW3 subset the reference calls (FFTW is an external dependency that is not
// vendored under /root/reference and is not installed in this image; no version is pinned by the
// reference, CMakeLists.txt:13-19 / Makefile:53-57).  Call sites covered:
//   plans     /root/reference/src/PressureSolverStructures.cpp:23-41   fftw_plan_r2r_1d(n, buf, buf, kind, FFTW_ESTIMATE)
//   executes  /root/reference/src/PressureEquation.cpp:85,112,139,174,207,241   fftw_execute_r2r(plan, ptr, ptr)
// The kinds are implemented from FFTW's documented definitions by oracle/fft_cpu.h (O(n log n)), so
// every number produced through this shim is "reference stencils/transposes + in-repo FFT", not FFTW.
#ifndef MIF_SHIM_FFTW3_H
#define MIF_SHIM_FFTW3_H

#include <cstdlib>

#include "../fft_cpu.h"

typedef enum { FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_DHT = 2, FFTW_REDFT00 = 3 } fftw_r2r_kind;
#define FFTW_ESTIMATE (1U << 6)
#define FFTW_MEASURE (0U)

struct mifshim_fftw_plan {
  mo_r2r_plan *r2r;
  double *in;
  double *out;
};
typedef mifshim_fftw_plan *fftw_plan;

static inline fftw_plan fftw_plan_r2r_1d(int n, double *in, double *out, fftw_r2r_kind kind, unsigned) {
  int mo_kind;
  switch (kind) {
    case FFTW_REDFT00: mo_kind = MO_REDFT00; break;
    case FFTW_R2HC: mo_kind = MO_R2HC; break;
    case FFTW_HC2R: mo_kind = MO_HC2R; break;
    default: std::abort();
  }
  fftw_plan p = new mifshim_fftw_plan();
  p->r2r = mo_r2r_create(n, mo_kind);
  p->in = in;
  p->out = out;
  return p;
}
static inline void fftw_execute(const fftw_plan p) { mo_r2r_exec(p->r2r, p->in, p->out); }
static inline void fftw_execute_r2r(const fftw_plan p, double *in, double *out) { mo_r2r_exec(p->r2r, in, out); }
static inline void fftw_destroy_plan(fftw_plan p) {
  if (p) {
    mo_r2r_destroy(p->r2r);
    delete p;
  }
}
static inline void *fftw_malloc(size_t bytes) { return std::aligned_alloc(64, (bytes + 63) / 64 * 64); }
static inline void fftw_free(void *ptr) { std::free(ptr); }
static inline int fftw_alignment_of(double *) { return 0; }
static inline void fftw_cleanup() {}

// Single-precision entry points (the reference's USE_DOUBLE=0 build, include/Real.h:9-17; call sites
// /root/reference/src/PressureSolverStructures.cpp:30-41,85-94 and src/PressureEquation.cpp:94,121,148,187,220,254:
// fftwf_plan_r2r_1d + fftwf_execute on the plan's own float buffer).  The stand-in widens the line to double, runs the
// double transform of fft_cpu.h and narrows the result: float storage, transform round-off below float epsilon --
// i.e. at least as accurate as fftwf would be.  Used only for the FP32 goldens (oracle/make_golden.py).
#include <vector>
struct mifshim_fftwf_plan {
  mo_r2r_plan *r2r;
  float *in;
  float *out;
  int n;
};
typedef mifshim_fftwf_plan *fftwf_plan;
static inline fftwf_plan fftwf_plan_r2r_1d(int n, float *in, float *out, fftw_r2r_kind kind, unsigned) {
  int mo_kind;
  switch (kind) {
    case FFTW_REDFT00: mo_kind = MO_REDFT00; break;
    case FFTW_R2HC: mo_kind = MO_R2HC; break;
    case FFTW_HC2R: mo_kind = MO_HC2R; break;
    default: std::abort();
  }
  fftwf_plan p = new mifshim_fftwf_plan();
  p->r2r = mo_r2r_create(n, mo_kind);
  p->in = in;
  p->out = out;
  p->n = n;
  return p;
}
static inline void fftwf_execute(const fftwf_plan p) {
  std::vector<double> wide(p->in, p->in + p->n);
  mo_r2r_exec(p->r2r, wide.data(), wide.data());
  for (int i = 0; i < p->n; i++) p->out[i] = (float)wide[i];
}
static inline void fftwf_destroy_plan(fftwf_plan p) {
  if (p) {
    mo_r2r_destroy(p->r2r);
    delete p;
  }
}
static inline void *fftwf_malloc(size_t bytes) { return std::aligned_alloc(64, (bytes + 63) / 64 * 64); }
static inline void fftwf_free(void *ptr) { std::free(ptr); }
static inline void fftwf_cleanup() {}

#endif  // MIF_SHIM_FFTW3_H
