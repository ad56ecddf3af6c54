// Shared device/host helpers for libacm_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/acm_b200.h"

namespace acm {

// ---- error plumbing (no exceptions across the C boundary) ---------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define ACM_CHECK_ARG(cond, ...)             \
  do {                                       \
    if (!(cond)) {                           \
      ::acm::set_error(__VA_ARGS__);         \
      return ACM_ERR_BAD_ARG;                \
    }                                        \
  } while (0)

#define ACM_LAUNCH_CHECK(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    ::acm::count_launch();                                                       \
    if (e__ != cudaSuccess) {                                                    \
      ::acm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
      return (int)e__;                                                           \
    }                                                                            \
  } while (0)

// ---- 8-feature slices: every lane of a row group owns 8 consecutive features ----------
// bf16: one 16-byte vector; fp32: two.

__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// relu of 8 packed bf16 values (4 HMNMX2 instead of 8 FMNMX after the unpack)
__device__ __forceinline__ uint4 relu_bf16x8(uint4 v) {
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
  const __nv_bfloat162 z = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __hmax2(h[i], z);
  return v;
}

// 8 consecutive floats from a 32-byte aligned shared-memory address as two LDS.128 (scalar reads
// at a 32-byte lane stride are 8-way bank conflicts)
__device__ __forceinline__ void load_smem8(const float* s, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(s);
  const float4 b = *reinterpret_cast<const float4*>(s + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// 16-byte read-only load with the L2 prefetch-size hint pinned to 64 B (narrow-row gathers: a
// 64-byte table row per edge, see acm_set_narrow_row_hint)
__device__ __forceinline__ uint4 ldg_l2_64(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// gather load of one slice: HINT = 1 -> L2::64B prefetch-size hint (narrow rows)
template <int HINT, typename S, typename T>
__device__ __forceinline__ void gather_load(S& s, const T* p) {
  if (HINT) s.load_l2_64(p); else s.load(p);
}

extern int g_narrow_row_hint;   // csr.cu; set by acm_set_narrow_row_hint

template <typename T>
struct Slice8;  // raw (register) form of 8 consecutive features of storage type T

template <>
struct Slice8<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void load_plain(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void load_l2_64(const __nv_bfloat16* p) { v = ldg_l2_64(p); }
  __device__ __forceinline__ void to_float(float* f) const { unpack_bf16x8(v, f); }
  __device__ __forceinline__ static void store(__nv_bfloat16* p, const float* f) {
    *reinterpret_cast<uint4*>(p) = pack_bf16x8(f);
  }
  __device__ __forceinline__ void zero() { v = make_uint4(0, 0, 0, 0); }
};

template <>
struct Slice8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void load_plain(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void load_l2_64(const float* p) {
    const uint4 x = ldg_l2_64(p), y = ldg_l2_64(p + 4);
    a = make_float4(__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w));
    b = make_float4(__uint_as_float(y.x), __uint_as_float(y.y), __uint_as_float(y.z), __uint_as_float(y.w));
  }
  __device__ __forceinline__ void to_float(float* f) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  __device__ __forceinline__ static void store(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *(reinterpret_cast<float4*>(p) + 1) = make_float4(f[4], f[5], f[6], f[7]);
  }
  __device__ __forceinline__ void zero() { a = make_float4(0, 0, 0, 0); b = a; }
};

// ---- cp.async (LDGSTS): global -> shared without register staging ---------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// one 8-feature slice (16 B in bf16, 32 B in fp32)
template <typename T> __device__ __forceinline__ void cp_async_slice(uint32_t dst, const T* src) {
  cp_async16(dst, src);
  if (sizeof(T) == 4) cp_async16(dst + 16, reinterpret_cast<const char*>(src) + 16);
}

// ---- mbarrier + TMA row gather (sm_100) -------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
// cp.async.bulk.tensor ... tile::gather4: FOUR rows of a 2-D tensor, given by four row indices, in one request of the
// TMA engine; they land back to back at `dst` (4 x box bytes) and complete `bar`.  Tensor map: box {row_words x 1}
// (scripts/gather4_probe.cu: a 4-row box is an illegal instruction), no swizzle.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(map), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
// host: cached tensor-map encoding (gemm_tc.cu).  kind 0 = bf16 elements with 128-byte swizzle, 1 = uint32 without swizzle
int tma_encode_2d(CUtensorMap* map, int kind, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                  uint32_t box_inner, uint32_t box_outer, const char* what);
int tma_encode_2d_u32(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                      uint32_t box_inner, uint32_t box_outer, const char* what);

// ---- packed fp32 arithmetic (sm_100: FFMA2 / FMUL2, two IEEE fp32 results per issue slot) ----------------------
// The row kernels are ISSUE bound (mix_bwd: 72 % of the issue slots for 88 % of the HBM peak); their inner loops are
// per-feature FMAs over 8-feature slices, i.e. natural pairs.  Each lane of a pair is a plain fma.rn.f32: bit-identical
// to the scalar form.  Measured effect on mix_bwd at the headline size: none beyond noise (6.53 vs 6.59 ms on boxes
// running at 1.72 / 1.73 GHz) -- ptxas spends most of the saved slots on the register-pair moves; kept because it is
// never slower and halves the FMA count of the rank-1 gather.
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
// acc[0..8) += w * f[0..8)   (four FFMA2)
__device__ __forceinline__ void axpy8(float (&acc)[8], const float w, const float (&f)[8]) {
  const float2 w2 = make_float2(w, w);
#pragma unroll
  for (int t = 0; t < 8; t += 2) {
    const float2 r = ffma2(w2, make_float2(f[t], f[t + 1]), make_float2(acc[t], acc[t + 1]));
    acc[t] = r.x;
    acc[t + 1] = r.y;
  }
}
// sum_t a[t] * b[t] over 8 features: two-lane partial sums (four FFMA2) + one add.  NOTE: a different summation
// order than the 8-long scalar FMA chain -- used where the value feeds a reduction anyway.
__device__ __forceinline__ float dot8(const float (&a)[8], const float (&b)[8]) {
  float2 s = make_float2(0.f, 0.f);
#pragma unroll
  for (int t = 0; t < 8; t += 2) s = ffma2(make_float2(a[t], a[t + 1]), make_float2(b[t], b[t + 1]), s);
  return s.x + s.y;
}

extern int g_gather_mode;       // spmm_fwd.cu; set by acm_set_gather_mode

// reduce over the LANES lanes of one row group (LANES is a power of two <= 32; groups are
// aligned, so xor-shuffles stay inside the group)
template <int LANES>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float sigmoidf_acc(float z) { return 1.0f / (1.0f + expf(-z)); }

// ---- long rows --------------------------------------------------------------------------------
// Rows with more than kLongRow stored edges are aggregated by a separate segment-parallel pass
// (spmm_long_rows_kernel) into an fp32 side buffer; the row kernels then read the finished sums
// instead of walking thousands of edges with one lane group (degree skew: Squirrel 1904,
// twitch-gamers ~35k).  long_rows is sorted ascending.
constexpr int kLongRow = 256;

struct LongRows {
  const int32_t* rows;   // [n_long] local row ids, ascending (nullptr: none)
  const float* acc;      // [n_long, width] finished fp32 sums
  int n_long;
};

__device__ __forceinline__ int find_long_row(const LongRows& lr, int64_t row) {
  int lo = 0, hi = lr.n_long;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (lr.rows[mid] < row) lo = mid + 1; else hi = mid;
  }
  return lo;  // caller guarantees presence
}

// ---- fused exchange: push to peer memory --------------------------------------------------------
// Under a 1-D row partition the operand table of an aggregation is needed on EVERY rank.  Instead
// of writing the own rows locally and calling an all-gather afterwards, the producing kernel
// (GEMM epilogue forward, mix_bwd backward) stores each finished row straight into all ranks'
// tables through NVLink peer mappings (symmetric memory): the transfer overlaps the compute and
// the all-gather disappears as a separate step.  tables[r] = base of rank r's [N_pad, width]
// table; rows are written at global index row_off + local row.
//
// NVSwitch multicast (NVLS): when the symmetric allocation also has a multicast mapping, `mc` is
// the multicast address of the same table and ONE multimem.st per 16 bytes replaces the n peer
// stores -- the switch replicates the write into every rank's copy (the own one included), so a
// rank's NVLink egress drops from (n-1) x to 1 x the bytes it produced.
constexpr int kMaxPeers = 8;
struct PeerTables {
  void* tables[kMaxPeers];
  void* mc;            // multicast address of the table, or nullptr -> unicast peer stores
  int n;               // 0 = no push (single GPU / NCCL path)
  int64_t row_off;
};

__device__ __forceinline__ void multimem_st16(void* addr, const uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}
// 16 bytes to byte offset `off` of every rank's table
__device__ __forceinline__ void peer_store16(const PeerTables& pt, const int64_t off, const uint4 v) {
  if (pt.mc) {
    multimem_st16(reinterpret_cast<char*>(pt.mc) + off, v);
  } else {
#pragma unroll 1
    for (int r = 0; r < pt.n; ++r) *reinterpret_cast<uint4*>(reinterpret_cast<char*>(pt.tables[r]) + off) = v;
  }
}

// layout of the attention parameter pack (see acm_b200.h)
__host__ __device__ __forceinline__ int pack_off_a(int fp, int k) { return k * fp; }
__host__ __device__ __forceinline__ int pack_off_avec(int fp) { return 4 * fp; }
__host__ __device__ __forceinline__ int pack_off_gamma(int fp, int k) { return 4 * fp + 16 + k * fp; }
__host__ __device__ __forceinline__ int pack_off_beta(int fp, int k) { return 8 * fp + 16 + k * fp; }
// derived tail of the VALUE pack (written by acm_pack_params when LayerNorm is live; not part of
// the gradient pack): gamma*a per feature and the per-channel sums the row kernels need, so that
// no CTA has to recompute them (1.25 M CTAs did, ~10 % of the fused forward in LayerNorm mode)
__host__ __device__ __forceinline__ int pack_off_ga(int fp, int k) { return 12 * fp + 16 + k * fp; }
__host__ __device__ __forceinline__ int pack_off_sum_ba(int fp) { return 16 * fp + 16; }      // [4] sum_f beta*a
__host__ __device__ __forceinline__ int pack_off_sum_ga(int fp) { return 16 * fp + 16 + 4; }  // [4] sum_f gamma*a

constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default, ACM-Geometric/layers.py:21-22

// dispatch helpers -----------------------------------------------------------------------
#define ACM_DISPATCH_FP(fp, ...)                         \
  switch (fp) {                                          \
    case 8: { constexpr int FP = 8; __VA_ARGS__; } break;     \
    case 16: { constexpr int FP = 16; __VA_ARGS__; } break;   \
    case 32: { constexpr int FP = 32; __VA_ARGS__; } break;   \
    case 64: { constexpr int FP = 64; __VA_ARGS__; } break;   \
    case 128: { constexpr int FP = 128; __VA_ARGS__; } break; \
    case 256: { constexpr int FP = 256; __VA_ARGS__; } break; \
    default:                                             \
      ::acm::set_error("padded feature width %d not in {8,16,32,64,128,256}", (int)(fp)); \
      return ACM_ERR_UNSUPPORTED;                        \
  }

}  // namespace acm
