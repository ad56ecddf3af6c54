// Probe of the sm_100 TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4): which tensor-map box shape the
// instruction wants, what lands where in shared memory, and how many bytes the mbarrier must expect.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/gather4_probe scripts/gather4_probe.cu && /tmp/gather4_probe
// Table: uint32 [N rows][W cols], element (r, c) = r * 1000 + c.  One warp gathers rows {3, 7, 1, 5} starting at column c0.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint32_t* out, int c0, int expect_bytes, int n_words) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint32_t* dst = reinterpret_cast<uint32_t*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
  const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(bar);
  const uint32_t dst_u32 = (uint32_t)__cvta_generic_to_shared(dst);
  for (int i = threadIdx.x; i < n_words; i += 32) dst[i] = 0xdeadbeefu;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u32) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u32), "r"(expect_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
        ::"r"(dst_u32), "l"(&tm), "r"(c0), "r"(3), "r"(7), "r"(1), "r"(5), "r"(bar_u32) : "memory");
  }
  // bounded wait: a wrong expect_tx must not hang the box
  uint32_t ok = 0;
  for (int it = 0; it < 2000000 && !ok; ++it) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar_u32) : "memory");
  }
  __syncwarp();
  if (threadIdx.x == 0) out[n_words] = ok;
  for (int i = threadIdx.x; i < n_words; i += 32) out[i] = dst[i];
}

int main() {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || !ptr) { printf("no encode entry point\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  const int N = 64, W = 256;
  std::vector<uint32_t> h((size_t)N * W);
  for (int r = 0; r < N; ++r) for (int c = 0; c < W; ++c) h[(size_t)r * W + c] = r * 1000 + c;
  uint32_t *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int n_words = 4 * W;
  cudaMalloc(&o, (n_words + 1) * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int rows_opts[2] = {1, 4};
  for (int bi = 0; bi < 2; ++bi) {
    for (int bw : {256, 64}) {
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)N};
      cuuint64_t strides[1] = {(cuuint64_t)W * 4};
      cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)rows_opts[bi]};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("box {%d cols, %d rows}: encode -> %d\n", bw, rows_opts[bi], (int)r);
      if (r != CUDA_SUCCESS) continue;
      for (int expect : {4 * bw * 4, bw * 4}) {
        cudaMemset(o, 0, (n_words + 1) * 4);
        probe<<<1, 32, 40000>>>(tm, o, bw == 64 ? 64 : 0, expect, n_words);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint32_t> res(n_words + 1);
        cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
        printf("  expect_tx %5d: sync=%s barrier_completed=%u | smem words [0]=%u [1]=%u [%d]=%u [%d]=%u [%d]=%u [%d]=%u\n", expect,
               cudaGetErrorString(e), res[n_words], res[0], res[1], bw, res[bw], 2 * bw, res[2 * bw], 3 * bw, res[3 * bw], 3 * bw + 1, res[3 * bw + 1]);
        if (e != cudaSuccess) { printf("  (context poisoned, stopping)\n"); return 0; }
      }
    }
  }
  return 0;
}
