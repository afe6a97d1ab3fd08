// tests/simt_emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// Stand-in for the CUDA runtime header that lets the UNCHANGED kernel sources of libmifgpu (csrc/*.cu, *.cuh) be
// compiled for the host by g++ and executed by the SIMT interpreter in emu_runtime.cpp: every CUDA thread of a block
// is a fiber; __syncthreads / __syncwarp / named barriers / warp shuffles are scheduling points.  Blocks run one
// after the other.  This exists to catch indexing and arithmetic mistakes in kernels on a machine without a GPU; it
// says nothing about races or performance, it is never built into or loaded by the product, and no parity claim
// rests on it (those are the `-m gpu` tests on a real B200).
#ifndef MIF_SIMT_EMU_CUDA_RUNTIME_H
#define MIF_SIMT_EMU_CUDA_RUNTIME_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

#define MIF_SIMT_EMU 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static

struct alignas(16) double2 { double x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline void sincos(float a, float *s, float *c) { sincosf(a, s, c); }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
static const int warpSize = 32;

// ---- runtime API subset --------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 999 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaPitchedPtr { void *ptr; size_t pitch, xsize, ysize; };
struct cudaPos { size_t x, y, z; };
struct cudaExtent { size_t width, height, depth; };
struct cudaMemcpy3DParms {
  void *srcArray; cudaPos srcPos; cudaPitchedPtr srcPtr;
  void *dstArray; cudaPos dstPos; cudaPitchedPtr dstPtr;
  cudaExtent extent; cudaMemcpyKind kind;
};
static inline cudaPitchedPtr make_cudaPitchedPtr(void *d, size_t p, size_t xsz, size_t ysz) { return cudaPitchedPtr{d, p, xsz, ysz}; }
static inline cudaPos make_cudaPos(size_t x, size_t y, size_t z) { return cudaPos{x, y, z}; }
static inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return cudaExtent{w, h, d}; }

cudaError_t emu_malloc(void **ptr, size_t bytes);
template <class T> static inline cudaError_t cudaMalloc(T **ptr, size_t bytes) { return emu_malloc((void **)ptr, bytes); }
template <class T> static inline cudaError_t cudaMallocHost(T **ptr, size_t bytes) { return emu_malloc((void **)ptr, bytes); }
cudaError_t cudaFree(void *ptr);
cudaError_t cudaFreeHost(void *ptr);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t stream = nullptr);
cudaError_t cudaMemset(void *dst, int value, size_t bytes);
cudaError_t cudaMemsetAsync(void *dst, int value, size_t bytes, cudaStream_t stream = nullptr);
cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms *p, cudaStream_t stream = nullptr);
cudaError_t cudaGetDeviceCount(int *count);
cudaError_t cudaSetDevice(int device);
cudaError_t cudaGetDevice(int *device);
enum { cudaDevAttrMultiProcessorCount = 16 };
cudaError_t cudaDeviceGetAttribute(int *value, int attr, int device);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t err);
cudaError_t cudaStreamCreate(cudaStream_t *stream);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *stream, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *stream, unsigned flags, int priority);
cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest);
cudaError_t cudaStreamDestroy(cudaStream_t stream);
cudaError_t cudaStreamSynchronize(cudaStream_t stream);
cudaError_t cudaEventCreate(cudaEvent_t *event);
enum { cudaEventDisableTiming = 2 };
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *event, unsigned flags);
cudaError_t cudaStreamWaitEvent(cudaStream_t stream, cudaEvent_t event, unsigned flags = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t event);
cudaError_t cudaEventDestroy(cudaEvent_t event);
cudaError_t cudaEventRecord(cudaEvent_t event, cudaStream_t stream = nullptr);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t start, cudaEvent_t stop);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *handle, void *ptr);
cudaError_t cudaIpcOpenMemHandle(void **ptr, cudaIpcMemHandle_t handle, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *ptr);
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// ---- SIMT interpreter ----------------------------------------------------------------------------------------
namespace emu {
void launch(dim3 grid, dim3 block, size_t dynamic_smem_bytes, cudaStream_t stream, const std::function<void()> &thread_body);
void *dynamic_smem();
void tma_block_end();
void block_barrier();
void warp_barrier();
void named_barrier(int id, int count);
void *shuffle_slot(int lane, int parity);
int lane_id();
int &shuffle_parity();
}  // namespace emu

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
[[noreturn]] static inline void __trap() { abort(); }
template <class T> static inline T __ldg(const T *ptr) { return *ptr; }
static inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int b = 0; b < 32; b++) r |= ((v >> b) & 1u) << (31 - b);
  return r;
}

// Warp shuffles: every lane publishes its value, one warp-wide scheduling point, every lane reads its source.  The
// slots are double buffered, so one barrier per shuffle is enough (a lane can only overwrite buffer p again after
// every lane has passed the barrier of the following shuffle, i.e. after they all read buffer p).
template <class T> static inline T emu_shuffle(T value, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  int &parity = emu::shuffle_parity();
  const int p = parity;
  parity ^= 1;
  memcpy(emu::shuffle_slot(emu::lane_id(), p), &value, sizeof(T));
  emu::warp_barrier();
  T out;
  memcpy(&out, emu::shuffle_slot(src_lane, p), sizeof(T));
  return out;
}
template <class T> static inline T __shfl_sync(unsigned, T value, int src_lane, int width = 32) {
  const int lane = emu::lane_id();
  return emu_shuffle(value, (lane / width) * width + (src_lane & (width - 1)));
}
template <class T> static inline T __shfl_down_sync(unsigned, T value, unsigned delta, int width = 32) {
  const int lane = emu::lane_id();
  const int src = lane + (int)delta;
  return emu_shuffle(value, (src / width == lane / width && src < 32) ? src : lane);
}
template <class T> static inline T __shfl_xor_sync(unsigned, T value, int mask, int width = 32) {
  (void)width;
  return emu_shuffle(value, emu::lane_id() ^ mask);
}

// CUDA's unqualified min / max in device code
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }
static inline double min(double a, double b) { return fmin(a, b); }
static inline double max(double a, double b) { return fmax(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }

#endif  // MIF_SIMT_EMU_CUDA_RUNTIME_H
