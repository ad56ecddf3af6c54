// Fused forward of an aggregate-first ACM layer (DESIGN "what next" item 1):
//
//     [S_L | S_H | HI] = [Z W_L | D W_H | X W_I]          three tcgen05 GEMMs sharing one CTA tile of 128 rows
//     O_k = relu(.) ; z_k = O_k . a_k ; s = sigmoid(z) ; att = softmax(s Avec / 3) ; Y = c sum_k att_k O_k
//
// in ONE launch: the fp32 accumulators never leave the SM.  Replaces the three `tn` GEMM launches plus
// the pre-aggregated epilogue launch of spmm_mix_fwd_kernel, i.e. torch.mm x3 + relu + attention3 + mix of
// ACM-Pytorch/models/layers.py:163-165,188-204 and 94-119, and deletes the 15.4 GB write + 15.4 GB
// read-back of [S_L|S_H|HI] between them at the headline size (out_features = 256 only: the row of
// three 256-wide channels is 768 fp32 columns, TMEM has 512).
//
// TMEM plan (512 columns = two regions of 256):   R0 = cols [0,256)   R1 = cols [256,512)
//   MMA warp     : HI -> R0, S_L -> R1, (wait: R0 drained) S_H -> R0
//   epilogue     : drain HI from R0 (identity logit; written to h_i as bf16, the form the backward reads) while
//                  S_L is being accumulated; logits of S_L while S_H is being accumulated; logits of S_H;
//                  softmax; second pass over R1/R0 for Y, with the thread's HI values read back from the h_i
//                  rows this same warp has just written (L2 hits, one full 32-byte sector per load pair).
//   In the TMEM accumulator layout one thread owns one row (lane = row), so the three dot products,
//   the softmax and the mix are thread-local; the FOUR warps of a lane quadrant split the 256 columns
//   of every channel in quarters and exchange their three partial logits through shared memory.
//   (A first version with two warps per quadrant that kept HI in 64 registers per thread was latency
//   bound -- 8 warps, 3.5 k dependent instructions each per tile, issue slots 20 % busy: 11.7 ms against
//   10.5 ms for the unfused launches, profiles/r2_fused_fwd_v1_ncu_*.  Sixteen epilogue warps hide those latencies.)
//
// Warps: 0 = TMA producer (4-stage ring of {A 128x64, B 256x64} bf16 tiles, 128B swizzle),
//        1 = MMA issuer + TMEM owner, 2..17 = epilogue (warp w: lane quadrant w & 3, column quarter (w-2) >> 2).
// Global stores: every thread writes 32-byte pieces of ITS OWN row with one 256-bit store (st.global.v8.b32,
// SASS STG.E.256 -- new on sm_100): each store instruction fills 32 whole sectors, so no shared-memory
// transposition tile is needed (the v2 kernel's tile cost ~450 instructions and 512 KB of shared-memory
// traffic per CTA tile next to the 1.1 MB the TMA ring and the MMAs already move).
#include <cuda.h>

#include "acm_common.cuh"

namespace acm {
namespace fused {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int FP = 256;                   // out_features (padded): the only width this kernel is built for
constexpr int kStages = 4;
constexpr int kParts = 4;                 // column parts per channel = epilogue warps per TMEM lane quadrant
constexpr int kEpiWarps = 4 * kParts;
constexpr int CW = FP / kParts;           // columns of every channel owned by one epilogue warp
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kABytes = BM * BK * 2;           // 16 KB
constexpr uint32_t kBBytes = FP * BK * 2;           // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// issue only: the destination registers are valid after tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]),
        "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]),
        "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  tmem_ld32_issue(taddr, r);
  tmem_wait_ld();
}
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, float* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
      : "r"(taddr) : "memory");
}
// sm_100 UMMA shared-memory descriptor, K-major operand, 128B swizzle (same encoding as gemm_tc.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((16u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ __forceinline__ uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void named_sync(int id, int n_threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

struct Params {
  int64_t n;            // rows
  int k;                // padded input width (<= 256)
  int f;                // true out_features (<= 256)
  int64_t m_tiles;
  const float* pack;    // attention parameter pack (acm_b200.h layout, fp = 256)
  float out_scale;
  void* y; int y_bf16; int64_t ldy;
  __nv_bfloat16* s_lh;  // [n, 512] = [S_L | S_H] (pre-relu), or nullptr (inference)
  __nv_bfloat16* h_i;   // [n, 256] pre-relu HI (always written: the second pass reads it back)
  float* att; float* sig;
};

// 32 bytes of this thread's own row, one 256-bit access (dst / src 32-byte aligned)
__device__ __forceinline__ void st_row32(void* dst, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ld_row32(const void* src, uint32_t* w) {
  asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(src) : "memory");
}

// sum_j relu(acc[row, j]) * a[j] over this warp's CW columns of one channel (thread = row); the raw (pre-relu)
// accumulators go to `dst` (this thread's row of the backward's table, bf16; nullptr: not stored) on the way
template <bool KEEP>
__device__ __forceinline__ float row_dot(uint32_t taddr, const float* a, __nv_bfloat16* dst, uint32_t* keep = nullptr) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;       // four independent FMA chains
#pragma unroll
  for (int cc = 0; cc < CW / 32; ++cc) {
    float v[32];
    tmem_ld32(taddr + (uint32_t)(cc * 32), v);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 av = *reinterpret_cast<const float4*>(a + cc * 32 + j);
      a0 = fmaf(fmaxf(v[j], 0.f), av.x, a0);
      a1 = fmaf(fmaxf(v[j + 1], 0.f), av.y, a1);
      a2 = fmaf(fmaxf(v[j + 2], 0.f), av.z, a2);
      a3 = fmaf(fmaxf(v[j + 3], 0.f), av.w, a3);
    }
    if (KEEP) {       // the bf16 pairs stay in registers for the second pass (early release of the TMEM region)
      uint32_t* w16 = keep + cc * 16;
#pragma unroll
      for (int j = 0; j < 16; ++j) w16[j] = pack2(v[2 * j], v[2 * j + 1]);
      if (dst) {
        st_row32(dst + cc * 32, w16);
        st_row32(dst + cc * 32 + 16, w16 + 8);
      }
    } else if (dst) {
      uint32_t w16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w16[j] = pack2(v[2 * j], v[2 * j + 1]);
      st_row32(dst + cc * 32, w16);
      st_row32(dst + cc * 32 + 16, w16 + 8);
    }
  }
  return (a0 + a1) + (a2 + a3);
}

// EARLY: the epilogue keeps bf16(S_H) of its columns in registers during the logit pass, so region R0 is released BEFORE
// the second pass and the HI MMA of the next tile runs under it (in the late variant the epilogue warps spent 12 % of
// their samples waiting for that MMA); the second pass then mixes bf16-rounded S_H -- the value the backward sees in the
// table anyway, like HI.  MEASURED SLOWER: 8.65 / 8.69 ms against 7.98 ms for the late variant in the same run (the 32
// extra live registers at the 96-register cap cost more in the second pass than the hidden MMA wait gives back).
// Opt-in A/B switch (acm_set_gemm_direct_store bit 2); the late variant is the default.
template <bool EARLY>
__global__ void __launch_bounds__(kThreads, 1)
fused_agg_fwd_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmD,
                     const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // ring | a_k [3][256] f32 | avec [16] f32 | z exchange [2][kParts][128][4] f32 | barriers
  const uint32_t off_a = kStages * kStageBytes;
  const uint32_t off_avec = off_a + 3 * FP * 4;
  const uint32_t off_z = off_avec + 64;
  const uint32_t off_bar = off_z + 2 * kParts * BM * 16;
  const uint32_t bar_full = base + off_bar;                // [kStages]
  const uint32_t bar_empty = bar_full + 8 * kStages;       // [kStages]
  const uint32_t bar_acc = bar_empty + 8 * kStages;        // [3]: HI, S_L, S_H accumulators complete
  const uint32_t bar_r0 = bar_acc + 24;                    // R0 drained (HI in registers)
  const uint32_t bar_done = bar_r0 + 8;                    // epilogue finished with TMEM (EARLY: with R1)
  const uint32_t bar_r0free = bar_done + 8;                // EARLY: epilogue finished with R0 (S_H kept in registers)
  const uint32_t tmem_slot = bar_r0free + 8;
  float* s_a = reinterpret_cast<float*>(gbase + off_a);
  float* s_avec = reinterpret_cast<float*>(gbase + off_avec);
  float* s_z = reinterpret_cast<float*>(gbase + off_z);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_total = (p.k + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int c = 0; c < 3; ++c) mbar_init(bar_acc + 8 * c, 1);
    mbar_init(bar_r0, kEpiWarps);
    mbar_init(bar_done, kEpiWarps);
    mbar_init(bar_r0free, kEpiWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 3 * FP; i += blockDim.x) s_a[i] = p.pack[i];
  if (threadIdx.x < 16) s_avec[threadIdx.x] = p.pack[pack_off_avec(FP) + threadIdx.x];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ---------------- TMA producer: per tile, channel order HI (X, W rows 512..), S_L (Z, 0..), S_H (D, 256..)
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmZ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
        const int m0 = (int)(tile * BM);
#pragma unroll 1
        for (int ch = 0; ch < 3; ++ch) {
          const CUtensorMap* ma = ch == 0 ? &tmX : ch == 1 ? &tmZ : &tmD;
          const int w0 = ch == 0 ? 2 * FP : ch == 1 ? 0 : FP;
          for (int kb = 0; kb < kb_total; ++kb) {
            mbar_wait(bar_empty + 8 * s, ph ^ 1);
            const uint32_t full = bar_full + 8 * s;
            mbar_expect_tx(full, kStageBytes);
            tma_load_2d(base + s * kStageBytes, ma, kb * BK, m0, full);
            tma_load_2d(base + s * kStageBytes + kABytes, &tmW, kb * BK, w0, full);
            if (++s == kStages) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, FP);
      int s = 0;
      uint32_t ph = 0;
      int t = 0;
      for (int64_t tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++t) {
        const uint32_t tp = (uint32_t)(t & 1);
        // the epilogue of the previous tile has released TMEM: all of it (late) / region R0 (EARLY)
        mbar_wait(EARLY ? bar_r0free : bar_done, tp ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int ch = 0; ch < 3; ++ch) {
          if (ch == 1 && EARLY) {          // S_L overwrites R1: the second pass of the previous tile must be over
            mbar_wait(bar_done, tp ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          if (ch == 2) {                   // S_H overwrites R0: HI must have been drained
            mbar_wait(bar_r0, tp);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          const uint32_t d_tmem = tmem_base + (ch == 1 ? (uint32_t)FP : 0u);
          for (int kb = 0; kb < kb_total; ++kb) {
            mbar_wait(bar_full + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = base + s * kStageBytes, sb = sa + kABytes;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16(d_tmem, make_desc(sa + k * UMMA_K * 2), make_desc(sb + k * UMMA_K * 2), idesc, (kb | k) ? 1u : 0u);
            umma_commit(bar_empty + 8 * s);
            if (++s == kStages) { s = 0; ph ^= 1; }
          }
          umma_commit(bar_acc + 8 * ch);
        }
      }
    }
  } else {
    // ---------------- epilogue: thread = row (TMEM lane), the kParts warps of a quadrant split the columns
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int c_lo = part * CW;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const float inv_k = 1.f / 3.f;
    int t = 0;
    for (int64_t tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++t) {
      const uint32_t tp = (uint32_t)(t & 1);
      const int64_t row = tile * BM + q * 32 + lane;
      const bool row_ok = row < p.n;

      // ---- HI: R0 -> h_i (bf16), identity logit
      mbar_wait(bar_acc + 0, tp);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __nv_bfloat16* hrow = p.h_i + (row_ok ? row : 0) * FP + c_lo;
      const float zI = row_dot<false>(lane_addr + (uint32_t)c_lo, s_a + 2 * FP + c_lo, row_ok ? hrow : nullptr);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_r0);

      // ---- logits of S_L (R1) while S_H is being accumulated, then of S_H (R0); raw accumulators -> table
      __nv_bfloat16* trow = (p.s_lh && row_ok) ? p.s_lh + row * (2 * FP) + c_lo : nullptr;
      mbar_wait(bar_acc + 8, tp);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const float zL = row_dot<false>(lane_addr + (uint32_t)(FP + c_lo), s_a + c_lo, trow);
      // this thread's HI values for the second pass: issued now, consumed after the S_H pass (L2 latency hidden)
      // (EARLY keeps bf16(S_H) in those registers instead and streams HI through a two-deep prefetch in the second pass)
      uint32_t hi[EARLY ? 1 : CW / 2];
      if (!EARLY) {
#pragma unroll
        for (int j = 0; j < CW / 16; ++j) ld_row32(hrow + j * 16, hi + j * 8);
      }
      mbar_wait(bar_acc + 16, tp);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t sh[EARLY ? CW / 2 : 1];
      const float zH = row_dot<EARLY>(lane_addr + (uint32_t)c_lo, s_a + FP + c_lo, trow ? trow + FP : nullptr, sh);
      if (EARLY) {                         // R0 is free: the next tile's HI accumulation may start
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_r0free);
      }

      // ---- combine the column parts (fixed order: identical sums in every warp of the quadrant), attention
      *reinterpret_cast<float4*>(s_z + ((tp * kParts + part) * BM + q * 32 + lane) * 4) = make_float4(zL, zH, zI, 0.f);
      named_sync(1 + q, 32 * kParts);
      float z0 = 0.f, z1 = 0.f, z2 = 0.f;
#pragma unroll
      for (int pp = 0; pp < kParts; ++pp) {
        const float4 zo = *reinterpret_cast<const float4*>(s_z + ((tp * kParts + pp) * BM + q * 32 + lane) * 4);
        z0 += zo.x;
        z1 += zo.y;
        z2 += zo.z;
      }
      float sgm[3], al[3];
      sgm[0] = __fdividef(1.f, 1.f + __expf(-z0));
      sgm[1] = __fdividef(1.f, 1.f + __expf(-z1));
      sgm[2] = __fdividef(1.f, 1.f + __expf(-z2));
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float l = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) l = fmaf(sgm[j], s_avec[j * 4 + k], l);
        al[k] = l * inv_k;
        mx = fmaxf(mx, al[k]);
      }
      float den = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        al[k] = __expf(al[k] - mx);
        den += al[k];
      }
      const float rden = __fdividef(1.f, den);
#pragma unroll
      for (int k = 0; k < 3; ++k) al[k] *= rden;
      if (row_ok && part < 2) {
        float* dst = part == 0 ? p.att : p.sig;
        if (dst) {
          dst[row * 3 + 0] = part == 0 ? al[0] : sgm[0];
          dst[row * 3 + 1] = part == 0 ? al[1] : sgm[1];
          dst[row * 3 + 2] = part == 0 ? al[2] : sgm[2];
        }
      }
      const float cL = p.out_scale * al[0], cH = p.out_scale * al[1], cI = p.out_scale * al[2];

      // ---- second pass: Y = c (att_L relu(S_L) + att_H relu(S_H) + att_I relu(HI)), 16 columns at a time
      // (columns >= f are padding with zero weights and are never stored; f % 16 == 0 for bf16 Y, % 8 for fp32)
      uint32_t hp[2][8];                  // EARLY: HI of the current / next 16 columns
      if (EARLY) ld_row32(hrow, hp[0]);
#pragma unroll
      for (int hc = 0; hc < CW / 16; ++hc) {
        const int c0 = c_lo + hc * 16;
        if (EARLY && hc + 1 < CW / 16) ld_row32(hrow + (hc + 1) * 16, hp[(hc + 1) & 1]);
        float vl[16], vh[16];
        tmem_ld16_issue(lane_addr + (uint32_t)(FP + c0), vl);
        if (!EARLY) tmem_ld16_issue(lane_addr + (uint32_t)c0, vh);
        tmem_wait_ld();
        if (EARLY) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 s2 = unpack2(sh[hc * 8 + j]);
            vh[2 * j] = s2.x;
            vh[2 * j + 1] = s2.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 h2 = unpack2(EARLY ? hp[hc & 1][j] : hi[hc * 8 + j]);
          vl[2 * j] = fmaf(cL, fmaxf(vl[2 * j], 0.f), fmaf(cH, fmaxf(vh[2 * j], 0.f), cI * fmaxf(h2.x, 0.f)));
          vl[2 * j + 1] = fmaf(cL, fmaxf(vl[2 * j + 1], 0.f), fmaf(cH, fmaxf(vh[2 * j + 1], 0.f), cI * fmaxf(h2.y, 0.f)));
        }
        if (row_ok) {
          if (p.y_bf16) {
            if (c0 < p.f) {
              uint32_t w8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w8[j] = pack2(vl[2 * j], vl[2 * j + 1]);
              st_row32(reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy + c0, w8);
            }
          } else {
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              if (c0 + h8 * 8 < p.f) {
                uint32_t w8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w8[j] = __float_as_uint(vl[h8 * 8 + j]);
                st_row32(reinterpret_cast<float*>(p.y) + row * p.ldy + c0 + h8 * 8, w8);
              }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_done);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------
static int make_map(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                    uint32_t box_inner, uint32_t box_outer, const char* what) {
  return tma_encode_2d(map, 0, ptr, inner, outer, pitch_elems * 2, box_inner, box_outer, what);   // cached (gemm_tc.cu)
}

int g_fused_early = 0;   // acm_set_gemm_direct_store bit 2: early release of TMEM region R0 in the fused forward (A/B)

}  // namespace fused
}  // namespace acm

extern "C" int acm_fused_agg_fwd(const void* z, const void* d, const void* x, int64_t ldx,
                                 const void* wcat_t, int64_t ldw, const float* pack,
                                 int64_t n_rows, int k, int f, int fp, float out_scale,
                                 void* y, int y_dtype, int64_t ldy, void* s_lh, void* h_i,
                                 float* att, float* sig, void* stream) {
  using namespace acm;
  using namespace acm::fused;
  if (fp != FP) { set_error("fused_agg_fwd: built for a padded out_features of %d (got %d)", FP, fp); return ACM_ERR_UNSUPPORTED; }
  ACM_CHECK_ARG(z && d && x && wcat_t && pack && y && att && h_i, "fused_agg_fwd: null pointer (h_i is required: the epilogue reads HI back from it)");
  ACM_CHECK_ARG(k >= 8 && k <= 256 && k % 8 == 0 && ldx >= k && ldw >= k, "fused_agg_fwd: need 8 <= k <= 256, k %% 8 == 0, ldx, ldw >= k");
  ACM_CHECK_ARG(f >= 1 && f <= fp, "fused_agg_fwd: need 1 <= f <= fp");
  ACM_CHECK_ARG(y_dtype == ACM_F32 || y_dtype == ACM_BF16, "fused_agg_fwd: bad y dtype %d", y_dtype);
  ACM_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 31) == 0 && (ldy * (y_dtype == ACM_BF16 ? 2 : 4)) % 32 == 0 &&
                    f % (y_dtype == ACM_BF16 ? 16 : 8) == 0,
                "fused_agg_fwd: y must be 32-byte aligned with a 32-byte multiple row pitch and f a multiple of 16 (bf16) / 8 (fp32)");
  ACM_CHECK_ARG(((reinterpret_cast<uintptr_t>(s_lh) | reinterpret_cast<uintptr_t>(h_i)) & 31) == 0, "fused_agg_fwd: s_lh / h_i must be 32-byte aligned");
  if (n_rows == 0) return 0;
  ACM_CHECK_ARG(n_rows < (1ll << 31) - BM, "fused_agg_fwd: more than 2^31 rows");
  Params p{};
  p.n = n_rows; p.k = k; p.f = f; p.m_tiles = (n_rows + BM - 1) / BM;
  p.pack = pack; p.out_scale = out_scale;
  p.y = y; p.y_bf16 = (y_dtype == ACM_BF16); p.ldy = ldy;
  p.s_lh = reinterpret_cast<__nv_bfloat16*>(s_lh); p.h_i = reinterpret_cast<__nv_bfloat16*>(h_i);
  p.att = att; p.sig = sig;
  CUtensorMap mz, md, mx, mw;
  int rc;
  if ((rc = make_map(&mz, z, (uint64_t)k, (uint64_t)n_rows, (uint64_t)ldx, BK, BM, "Z"))) return rc;
  if ((rc = make_map(&md, d, (uint64_t)k, (uint64_t)n_rows, (uint64_t)ldx, BK, BM, "D"))) return rc;
  if ((rc = make_map(&mx, x, (uint64_t)k, (uint64_t)n_rows, (uint64_t)ldx, BK, BM, "X"))) return rc;
  if ((rc = make_map(&mw, wcat_t, (uint64_t)k, (uint64_t)(3 * FP), (uint64_t)ldw, BK, FP, "Wcat^T"))) return rc;
  const size_t smem = (size_t)kStages * kStageBytes + 3 * FP * 4 + 64 + 2 * kParts * BM * 16 + 128 + 1024;
  auto kern = g_fused_early ? fused_agg_fwd_kernel<true> : fused_agg_fwd_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("fused_agg_fwd: smem attribute (%zu bytes): %s", smem, cudaGetErrorString(e)); return (int)e; }
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t grid = p.m_tiles < sms ? p.m_tiles : sms;
  kern<<<(unsigned)grid, kThreads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(mz, md, mx, mw, p);
  ACM_LAUNCH_CHECK("fused_agg_fwd");
  return 0;
}
