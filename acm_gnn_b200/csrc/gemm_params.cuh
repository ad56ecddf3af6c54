// Shared between gemm.cu (C-ABI entry points) and gemm_simt.cu (CUDA-core kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acm {

// C[m,n] = sum_k A[m,k] B[k,n];  element (i,k) of A at a + i*a_rs + k*a_cs (elements), same for B.
struct GemmParams {
  const void* a; int64_t a_rs, a_cs;
  const void* b; int64_t b_rs, b_cs;
  void* c0; int64_t ldc0, ncols0;   // output columns [0, ncols0)
  void* c1; int64_t ldc1;           // output columns [ncols0, n)   (may be null when ncols0 == n)
  int64_t m, n, k;
  int64_t k_chunk;     // K range per blockIdx.z (set by gemm_simt)
  int c_bf16;          // output storage: 1 bf16, 0 fp32
  int relu_cols;       // relu on output columns < relu_cols
  int atomic;          // accumulate into fp32 c0 with atomicAdd (split-K)
};

int gemm_simt(int dtype, GemmParams p, int splits, cudaStream_t st);

}  // namespace acm
