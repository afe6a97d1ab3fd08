// mif_tma.cuh -- Blackwell bulk-tensor copies (TMA) and mbarriers for the Poisson sweeps (sm_100a).
//
// Device side: thin wrappers around the PTX the strided sweeps use -- cp.async.bulk.tensor.3d (global <-> shared
// boxes described by a CUtensorMap kernel parameter), mbarrier init / arrive.expect_tx / try_wait.parity, bulk
// commit / wait groups and the generic -> async proxy fence.  Host side: encode_map() builds a rank-3 FP64 tensor
// map through cuTensorMapEncodeTiled, fetched from the driver with cudaGetDriverEntryPoint so that libmifgpu.so does
// not link against libcuda.
//
// Under MIF_SIMT_EMU (tests/simt_emu, the CPU interpreter of these kernel sources -- test infrastructure) the same
// names are served by tests/simt_emu/emu_runtime.cpp: copies are performed lazily (loads when the barrier is first
// waited on, stores when the group is waited on) so that missing waits and early buffer reuse show up as NaNs.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#ifndef MIF_SIMT_EMU
#include <cuda.h>
#else
struct alignas(64) CUtensorMap {  // emulated descriptor: rank 3, FP64
  void *base;
  uint64_t dim[3];
  uint64_t stride_bytes[3];  // stride_bytes[0] = 8
  uint32_t box[3];
  uint32_t swizzle_mask;  // 0: none, 3: CU_TENSOR_MAP_SWIZZLE_64B
  uint64_t pad_[5];
};
#define __grid_constant__
namespace emu {
void tma_load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2);
void tma_store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2);
void tma_store_bulk(void *global_dst, const void *smem_src, unsigned bytes);
void tma_commit_group();
void tma_wait_group(int pending_allowed, bool read_only);
void mbar_init(uint64_t *bar, int count);
void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes);
bool mbar_test(uint64_t *bar, unsigned parity);
void spin_yield();
}  // namespace emu
#endif

namespace mifgpu {
namespace tma {

#ifndef MIF_SIMT_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// Makes the barrier initialisation visible to the async proxy (the TMA unit) before the first copy names it.
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MIF_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MIF_DONE_%=;\n"
      "bra MIF_WAIT_%=;\n"
      "MIF_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared box copy; completion is signalled on `bar` with the box size in bytes (zero-filled outside the
// tensor's extents, which still count).
__device__ __forceinline__ void load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// shared -> global box copy (clipped at the tensor's extents); joins the current bulk group.
__device__ __forceinline__ void store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// shared -> global copy of a contiguous run of bytes (multiple of 16, both ends 16-byte aligned); the destination may be
// memory of a peer GPU mapped into this process (NVLink).  Joins the current bulk group.
__device__ __forceinline__ void store_bulk(void *global_dst, const void *smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(global_dst)),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until the bulk stores of this thread have finished READING shared memory (the source may be overwritten).
__device__ __forceinline__ void wait_stores_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// Wait until the bulk stores of this thread are complete.
__device__ __forceinline__ void wait_stores_done() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Orders the generic-proxy shared-memory writes of this thread before later async-proxy (TMA) reads of them.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_map(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// Address bits the 128-byte swizzle mixes: shared-window address on the device.
__device__ __forceinline__ uintptr_t swizzle_address(const void *p) { return (uintptr_t)smem_u32(p); }
#else
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { emu::mbar_init(bar, count); }
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) { emu::mbar_arrive_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  while (!emu::mbar_test(bar, parity)) emu::spin_yield();
}
__device__ __forceinline__ void load_3d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  emu::tma_load_3d(smem_dst, map, bar, c0, c1, c2);
}
__device__ __forceinline__ void store_3d(const CUtensorMap *map, const void *smem_src, int c0, int c1, int c2) {
  emu::tma_store_3d(map, smem_src, c0, c1, c2);
}
__device__ __forceinline__ void store_bulk(void *global_dst, const void *smem_src, unsigned bytes) {
  emu::tma_store_bulk(global_dst, smem_src, bytes);
}
__device__ __forceinline__ void commit_group() { emu::tma_commit_group(); }
__device__ __forceinline__ void wait_stores_read() { emu::tma_wait_group(0, true); }
__device__ __forceinline__ void wait_stores_done() { emu::tma_wait_group(0, false); }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void prefetch_map(const CUtensorMap *) {}
__device__ __forceinline__ uintptr_t swizzle_address(const void *p) { return reinterpret_cast<uintptr_t>(p); }
#endif

// Byte offset inside a buffer of dense 64-byte rows (8 doubles) under CU_TENSOR_MAP_SWIZZLE_64B: the 16-byte chunk
// index inside each 64-byte row (address bits 4-5) is XORed with address bits 7-8, i.e. with the index of the
// 128-byte line modulo 4; `mask` = 3 (swizzled map) or 0 (plain map); the buffer starts on a 1024-byte boundary.
// Measured on a B200 with scripts/probes/tma_probe.cu.  (CU_TENSOR_MAP_SWIZZLE_128B is not usable for 64-byte rows:
// the copy engine then pads every row to a 128-byte line in shared memory.)
__device__ __forceinline__ unsigned swizzle_offset(unsigned off, unsigned mask) { return off ^ (((off >> 7) & mask) << 4); }

// ---- host -------------------------------------------------------------------------------------------------------
// Rank-3 FP64 tensor map: element (c0, c1, c2) lives at base + c0 * 8 + c1 * stride1_bytes + c2 * stride2_bytes;
// boxes of box0 x box1 x 1 elements.  Returns false (with a message in `why`) when the driver refuses the geometry.
inline bool encode_map(CUtensorMap *map, const void *base, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1_bytes,
                       uint64_t stride2_bytes, uint32_t box0, uint32_t box1, bool swizzle64, int l2_promotion,
                       const char **why) {
  static const char *dummy;
  if (!why) why = &dummy;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15) || stride1_bytes >= (1ull << 40) ||
      stride2_bytes >= (1ull << 40) || box0 == 0 || box1 == 0 || box0 > 256 || box1 > 256 || ((box0 * 8) & 15) || dim0 == 0 ||
      dim1 == 0 || dim2 == 0 || (swizzle64 && box0 * 8 > 64)) {
    *why = "geometry outside the tensor-map limits";
    return false;
  }
#ifndef MIF_SIMT_EMU
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  static bool looked_up = false;
  if (!looked_up) {
    looked_up = true;
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult status;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &status) == cudaSuccess &&
        status == cudaDriverEntryPointSuccess)
      encode = reinterpret_cast<EncodeFn>(fn);
    else
      (void)cudaGetLastError();
  }
  if (!encode) {
    *why = "cuTensorMapEncodeTiled is not available from this driver";
    return false;
  }
  const cuuint64_t dims[3] = {dim0, dim1, dim2};
  const cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  const cuuint32_t box[3] = {box0, box1, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapL2promotion promo = l2_promotion == 0   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : l2_promotion == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : l2_promotion == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                           : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                             promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    *why = "cuTensorMapEncodeTiled rejected the geometry";
    return false;
  }
  return true;
#else
  (void)l2_promotion;
  map->base = const_cast<void *>(base);
  map->dim[0] = dim0; map->dim[1] = dim1; map->dim[2] = dim2;
  map->stride_bytes[0] = 8; map->stride_bytes[1] = stride1_bytes; map->stride_bytes[2] = stride2_bytes;
  map->box[0] = box0; map->box[1] = box1; map->box[2] = 1;
  map->swizzle_mask = swizzle64 ? 3 : 0;
  return true;
#endif
}

}  // namespace tma
}  // namespace mifgpu
