// Inter-layer glue of the two-layer stack in ONE pass per direction:
//   fea1 = F.dropout(F.relu(fea1), p, training) [+ xX]     (ACM-Pytorch/models/models.py:160-164,
//                                                           ACM-Geometric/models.py:70-74)
// The reference runs relu, dropout (value + bool mask) and the add as three ATen launches forward and
// two backward (threshold_backward, mul) over [N, hidden]: ~7.5 passes of the activation matrix plus a
// byte-per-element mask.  Here: forward reads x (+ add) and writes y and ONE BIT per element; backward
// reads g and the bits and writes dx.
//
// Random bits: Philox4x32-10 (Salmon et al., SC'11; the generator behind torch's CUDA dropout), keyed by
// the 64-bit seed, counter = (element_index / 4, offset): element e is kept iff
//   philox(seed, [e/4, offset])[e % 4] >= floor(p * 2^32),
// so the mask is a pure function of (seed, offset, e) -- independent of the launch geometry and restated
// on the CPU by oracle/philox_oracle.py for the bit-exact parity test.  seed/offset live in DEVICE memory
// (rng_state[0], rng_state[1]) and the offset is advanced by a trailing one-thread launch, so a captured
// CUDA graph draws a fresh mask on every replay.  This is NOT the reference's random stream (torch keeps
// its own Philox offsets per launch geometry): callers opt in (functional.inter_layer_glue).
#include "acm_common.cuh"

namespace acm {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<__nv_bfloat16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <typename T> __device__ __forceinline__ float load1(const T* p);
template <> __device__ __forceinline__ float load1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void store1(T* p, float v);
template <> __device__ __forceinline__ void store1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// one thread = 8 consecutive elements = one mask byte.  T = type of x, TY = type of add and y (TY = float
// with T = bf16 is torch's type promotion of "bf16 activations + fp32 xX")
template <typename T, typename TY>
__global__ void __launch_bounds__(256)
glue_fwd_kernel(const T* __restrict__ x, const TY* __restrict__ add, TY* __restrict__ y, uint8_t* __restrict__ mask,
                int64_t total, int relu, uint32_t thr, float scale, const uint64_t* __restrict__ rng_state) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t base = t * 8;
  if (base >= total) return;
  uint32_t keep = 0xffu;
  if (thr) {
    const uint64_t seed = rng_state[0], off = rng_state[1];
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t g0 = (uint64_t)t * 2;
    const uint4 r0 = philox4x32_10(make_uint4((uint32_t)g0, (uint32_t)(g0 >> 32), (uint32_t)off, (uint32_t)(off >> 32)), key);
    const uint4 r1 = philox4x32_10(make_uint4((uint32_t)(g0 + 1), (uint32_t)((g0 + 1) >> 32), (uint32_t)off, (uint32_t)(off >> 32)), key);
    keep = (r0.x >= thr ? 1u : 0u) | (r0.y >= thr ? 2u : 0u) | (r0.z >= thr ? 4u : 0u) | (r0.w >= thr ? 8u : 0u) |
           (r1.x >= thr ? 16u : 0u) | (r1.y >= thr ? 32u : 0u) | (r1.z >= thr ? 64u : 0u) | (r1.w >= thr ? 128u : 0u);
  }
  float v[8], a[8];
  const bool full = base + 8 <= total;
  if (full) {
    Slice8<T> s;
    s.load(x + base);
    s.to_float(v);
    if (add) {
      Slice8<TY> sa;
      sa.load(add + base);
      sa.to_float(a);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = (base + j < total) ? load1<T>(x + base + j) : 0.f;
      a[j] = (add && base + j < total) ? load1<TY>(add + base + j) : 0.f;
    }
  }
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool on = ((keep >> j) & 1u) && (!relu || v[j] > 0.f);
    bits |= on ? (1u << j) : 0u;
    // the reference rounds after the dropout scaling and again after the add (two ATen ops)
    float o = on ? round_to<T>(__fmul_rn(v[j], scale)) : 0.f;   // no contraction with the add below
    if (add) o = __fadd_rn(o, a[j]);
    v[j] = o;
  }
  if (full) {
    Slice8<TY>::store(y + base, v);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (base + j < total) store1<TY>(y + base + j, v[j]);
  }
  if (mask) mask[t] = (uint8_t)bits;
}

__global__ void glue_bump_kernel(uint64_t* rng_state) { rng_state[1] += 1; }

// TY = type of g (that of y), T = type of dx (that of x): autograd casts g to T first, then masks and scales
template <typename T, typename TY>
__global__ void __launch_bounds__(256)
glue_bwd_kernel(const TY* __restrict__ g, const uint8_t* __restrict__ mask, T* __restrict__ dx, int64_t total, float scale) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t base = t * 8;
  if (base >= total) return;
  const uint32_t bits = mask[t];
  float v[8];
  if (base + 8 <= total) {
    Slice8<TY> s;
    s.load(g + base);
    s.to_float(v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ((bits >> j) & 1u) ? round_to<T>(v[j]) * scale : 0.f;
    Slice8<T>::store(dx + base, v);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (base + j < total) store1<T>(dx + base + j, ((bits >> j) & 1u) ? round_to<T>(load1<TY>(g + base + j)) * scale : 0.f);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace acm

extern "C" int acm_glue_fwd(int dtype, int out_dtype, const void* x, const void* add, void* y, uint8_t* mask, int64_t total,
                            int relu, float p, uint64_t* rng_state, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == 0 || dtype == 1, "glue_fwd: dtype must be ACM_F32 or ACM_BF16");
  ACM_CHECK_ARG(out_dtype == dtype || (out_dtype == 0 && add), "glue_fwd: y has the type of x, or fp32 when an fp32 `add` promotes it");
  ACM_CHECK_ARG(x && y && total >= 0, "glue_fwd: null pointer");
  ACM_CHECK_ARG(p >= 0.f && p < 1.f, "glue_fwd: dropout probability must be in [0, 1) (got %g)", (double)p);
  ACM_CHECK_ARG(p == 0.f || rng_state, "glue_fwd: p > 0 needs the device rng state {seed, offset}");
  ACM_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(add), "glue_fwd: x, add and y must be 16-byte aligned, contiguous");
  if (total == 0) return 0;
  const int64_t blocks = ((total + 7) / 8 + 255) / 256;
  ACM_CHECK_ARG(blocks < (1ll << 31), "glue_fwd: too many elements");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // keep iff r >= thr: P(drop) = thr / 2^32; thr == 0 also switches the generator off
  const double thr_d = (double)p * 4294967296.0;
  const uint32_t thr = thr_d >= 4294967295.0 ? 0xffffffffu : (uint32_t)thr_d;
  const float scale = thr ? 1.f / (1.f - p) : 1.f;
  if (dtype == 0)
    glue_fwd_kernel<float, float><<<(unsigned)blocks, 256, 0, st>>>((const float*)x, (const float*)add, (float*)y, mask, total, relu, thr, scale, rng_state);
  else if (out_dtype == 0)
    glue_fwd_kernel<__nv_bfloat16, float><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const float*)add, (float*)y, mask, total, relu, thr, scale, rng_state);
  else
    glue_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)add, (__nv_bfloat16*)y, mask, total, relu, thr, scale, rng_state);
  ACM_LAUNCH_CHECK("glue_fwd");
  if (thr) {
    glue_bump_kernel<<<1, 1, 0, st>>>(rng_state);
    ACM_LAUNCH_CHECK("glue_bump");
  }
  return 0;
}

extern "C" int acm_glue_bwd(int dtype, int out_dtype, const void* g, const uint8_t* mask, void* dx, int64_t total, float scale, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == 0 || dtype == 1, "glue_bwd: dtype must be ACM_F32 or ACM_BF16");
  ACM_CHECK_ARG(out_dtype == dtype || out_dtype == 0, "glue_bwd: g has the type of dx, or fp32");
  ACM_CHECK_ARG(g && mask && dx && total >= 0, "glue_bwd: null pointer");
  ACM_CHECK_ARG(aligned16(g) && aligned16(dx), "glue_bwd: g and dx must be 16-byte aligned, contiguous");
  if (total == 0) return 0;
  const int64_t blocks = ((total + 7) / 8 + 255) / 256;
  ACM_CHECK_ARG(blocks < (1ll << 31), "glue_bwd: too many elements");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == 0)
    glue_bwd_kernel<float, float><<<(unsigned)blocks, 256, 0, st>>>((const float*)g, mask, (float*)dx, total, scale);
  else if (out_dtype == 0)
    glue_bwd_kernel<__nv_bfloat16, float><<<(unsigned)blocks, 256, 0, st>>>((const float*)g, mask, (__nv_bfloat16*)dx, total, scale);
  else
    glue_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)g, mask, (__nv_bfloat16*)dx, total, scale);
  ACM_LAUNCH_CHECK("glue_bwd");
  return 0;
}
