// CUDA-core (fp32 FMA) tiled GEMM with arbitrary operand strides: the exact-fp32 parity
// path of the three dense contractions of the ACM layer and the shape-agnostic fallback
// of the tcgen05 path (gemm_tc.cu) for shapes its TMA descriptors cannot express.
//   forward   H = X . [W_low|W_high|W_mlp]      (ACM-Pytorch/models/layers.py:163-165,179-194)
//   backward  dWcat = X^T . dH ,  dX = dH . Wcat^T      (autograd of the above)
#include "acm_common.cuh"
#include "gemm_params.cuh"

namespace acm {


template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T, bool A_KFAST, bool B_NFAST>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * p.k_chunk;
  const int64_t kend = min(p.k, kbeg + p.k_chunk);
  const T* A = reinterpret_cast<const T*>(p.a);
  const T* B = reinterpret_cast<const T*>(p.b);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      int mm, kk;
      if (A_KFAST) { mm = idx >> 4; kk = idx & 15; } else { kk = idx >> 6; mm = idx & 63; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.m && gk < kend) ? ld_as_float<T>(A + gm * p.a_rs + gk * p.a_cs) : 0.f;
      int nn, kb;
      if (B_NFAST) { kb = idx >> 6; nn = idx & 63; } else { nn = idx >> 4; kb = idx & 15; }
      const int64_t gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < p.n && gkb < kend) ? ld_as_float<T>(B + gkb * p.b_rs + gn * p.b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gm = m0 + ty * 4 + i;
    if (gm >= p.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t gn = n0 + tx * 4 + j;
      if (gn >= p.n) continue;
      float v = acc[i][j];
      if (gn < p.relu_cols) v = fmaxf(v, 0.f);
      if (p.atomic) {
        atomicAdd(reinterpret_cast<float*>(p.c0) + gm * p.ldc0 + gn, v);
      } else {
        void* base = (gn < p.ncols0) ? p.c0 : p.c1;
        const int64_t off = (gn < p.ncols0) ? gm * p.ldc0 + gn : gm * p.ldc1 + (gn - p.ncols0);
        if (p.c_bf16) reinterpret_cast<__nv_bfloat16*>(base)[off] = __float2bfloat16_rn(v);
        else reinterpret_cast<float*>(base)[off] = v;
      }
    }
  }
}

int gemm_simt(int dtype, GemmParams p, int splits, cudaStream_t st) {
  if (p.m == 0 || p.n == 0) return 0;
  if (splits < 1) splits = 1;
  int64_t chunk = (p.k + splits - 1) / splits;
  chunk = ((chunk + BK - 1) / BK) * BK;
  if (chunk == 0) chunk = BK;
  splits = (int)((p.k + chunk - 1) / chunk);
  if (splits < 1) splits = 1;
  p.k_chunk = chunk;
  p.atomic = splits > 1 ? 1 : p.atomic;
  // M tiles on blockIdx.x (2^31 limit: node dimension), N tiles on y, split-K slices on z
  const int64_t gx = (p.m + BM - 1) / BM, gy = (p.n + BN - 1) / BN;
  ACM_CHECK_ARG(gx < (1ll << 31) && gy <= 65535 && splits <= 65535, "gemm_simt: grid too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)splits);
  const bool akf = (p.a_cs == 1), bnf = (p.b_cs == 1);
#define ACM_G(TT)                                                                         \
  if (akf && bnf) gemm_simt_kernel<TT, true, true><<<grid, 256, 0, st>>>(p);              \
  else if (akf) gemm_simt_kernel<TT, true, false><<<grid, 256, 0, st>>>(p);               \
  else if (bnf) gemm_simt_kernel<TT, false, true><<<grid, 256, 0, st>>>(p);               \
  else gemm_simt_kernel<TT, false, false><<<grid, 256, 0, st>>>(p);
  if (dtype == ACM_BF16) { ACM_G(__nv_bfloat16) } else { ACM_G(float) }
#undef ACM_G
  ACM_LAUNCH_CHECK("gemm_simt");
  return 0;
}

}  // namespace acm
