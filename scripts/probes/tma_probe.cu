// tma_probe.cu -- prints where cp.async.bulk.tensor puts the elements of a box of 8-double (64-byte) rows in shared
// memory under each swizzle mode (diagnostic for csrc/mif_poisson_tma.cuh; not part of the product).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, double *out, int n_out, int box_bytes) {
  extern __shared__ unsigned char raw[];
  unsigned char *smem = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
  double *S = reinterpret_cast<double *>(smem);
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 32768);
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) S[i] = -1.0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(box_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(S)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(bar)), "r"(0), "r"(0), "r"(0) : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) out[i] = S[i];
}

int main() {
  typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult st;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess || !fn) { printf("no encode\n"); return 1; }
  EncodeFn encode = (EncodeFn)fn;
  const int rows = 32, cols = 8;
  std::vector<double> h(rows * cols);
  for (int i = 0; i < rows * cols; i++) h[i] = i;
  double *g, *out;
  const int n_out = 1024;  // 8 KB of shared memory dumped
  cudaMalloc(&g, h.size() * 8); cudaMalloc(&out, n_out * 8);
  cudaMemcpy(g, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const CUtensorMapSwizzle modes[4] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
  const char *names[4] = {"NONE", "32B", "64B", "128B"};
  for (int m = 0; m < 4; m++) {
    CUtensorMap map;
    const cuuint64_t dims[3] = {cols, rows, 1};
    const cuuint64_t strides[2] = {cols * 8, (cuuint64_t)cols * 8 * rows};
    const cuuint32_t box[3] = {cols, rows, 1}, es[3] = {1, 1, 1};
    CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, modes[m],
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("mode %s: encode failed %d\n", names[m], (int)rc); continue; }
    probe<<<1, 128, 40000>>>(map, out, n_out, rows * cols * 8);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %s: kernel failed: %s\n", names[m], cudaGetErrorString(e)); return 1; }
    std::vector<double> o(n_out);
    cudaMemcpy(o.data(), out, n_out * 8, cudaMemcpyDeviceToHost);
    printf("mode %s: smem 16-byte chunk c of each 128-byte line holds source element (row*8+col) / 2 pairs; -1 = untouched\n", names[m]);
    for (int line = 0; line < 24; line++) {
      printf("  smem[%4d..]:", line * 128);
      for (int c = 0; c < 8; c++) printf(" %4d", (int)o[line * 16 + c * 2]);
      printf("\n");
    }
    // check against the address-based model: off ^= ((off >> 7) & mask) << 4
    const unsigned mask = m == 0 ? 0 : (m == 1 ? 1 : (m == 2 ? 3 : 7));
    int bad = 0;
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        unsigned off = r * 64 + c * 8;
        off ^= ((off >> 7) & mask) << 4;
        if (o[off / 8] != r * 8 + c) bad++;
      }
    printf("  address-based model (mask %u): %d mismatches\n", mask, bad);
  }
  return 0;
}
