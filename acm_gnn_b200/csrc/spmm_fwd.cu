// Fused forward of the ACM layer: CSR row-parallel aggregation of the [HL|HH] table
// (one gather per stored edge serves the low-pass AND the high-pass channel) + relu +
// per-node channel attention (optional LayerNorm, sigmoid, KxK mix, softmax) + weighted sum.
//
// Replaces, in one launch, the ~25 ATen launches of
//   ACM-Pytorch/models/layers.py:176-204 + attention3/attention4 (94-152)
//   (ACM-Geometric/layers.py:57-116; the LayerNorm branch is live there).
//
// Mapping: a row is owned by a group of LANES = FP/8 lanes; each lane owns the same 8
// features of every channel, so the whole epilogue is lane-local except K (or 3K with
// LayerNorm) group reductions.  Per edge a lane issues two 16-byte loads (bf16) from the
// neighbour's table row: features [8g,8g+8) of HL and of HH.  Accumulation is fp32.
#include <cuda.h>

#include "acm_common.cuh"

namespace acm {

struct FwdParams {
  int64_t n_rows, row0;
  const int64_t* rowptr;
  const int32_t* col;
  const float* val;
  const float* rowscale;
  const void* table;
  const void* h_i;
  const void* o_s;
  const float* pack;
  int k, ln, variant, f, vec_y, pre_agg, y_bf16;
  float out_scale;
  float* y;
  int64_t ldy;
  void* o_save;
  float* att;
  float* sig;
  LongRows lr;
  const int32_t* row_order;   // optional: slot -> local row (degree-sorted processing order), or nullptr
};

constexpr int kFwdWarps = 8;
constexpr int kUnroll = 4;

// ---- asynchronous gather (cp.async ring in shared memory) ---------------------------------------
// Each lane copies its own two 8-feature slices of a neighbour row straight from global to shared
// memory (LDGSTS, no register staging) kAsyncBytes/row-slot deep, so the number of neighbour rows
// in flight per lane is fixed by the ring depth instead of by ptxas' register-pressure driven load
// scheduling (which, capped at 64 registers, interleaved the LDGs with the FMAs and left ~2 rows
// in flight).  A lane only ever reads back the bytes it copied itself: no warp synchronisation.
template <typename T> struct AsyncCfg { static constexpr int kStages = sizeof(T) == 2 ? 8 : 4; };

// ---- bulk-copy gather (gather mode 2, FP = 256 only: one row per warp) -------------------------
// The neighbour's whole [HL|HH] table row (1 KB in bf16) is one contiguous span, so ONE
// cp.async.bulk (the TMA engine's linear-copy form, SASS UBLKCP) issued by lane 0 replaces the
// 64 per-lane LDGSTS of mode 1; completion is tracked per ring slot by an mbarrier
// (arrive.expect_tx by the issuing lane, complete_tx by the copy engine).
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- TMA row gather (gather mode 3, FP = 256 in bf16: one row per warp) --------------------------------------
// sm_100's cp.async.bulk.tensor ... tile::gather4 fetches FOUR rows of a 2-D tensor, given by four row indices, with one
// instruction of the TMA engine: the "TMA-staged feature tiles" of the north star.  The [N, 2*FP] bf16 table is
// described to the engine as a uint32 [N, 256] tensor with a {256 x 1} box (established with scripts/gather4_probe.cu:
// a box of 4 rows is an illegal instruction, the four rows land back to back in shared memory, 4 KB per request).
// One elected lane issues a request per group of four edges into a kTmaStages-deep ring of 4-KB stages, completion
// by one mbarrier per stage; every lane then reads its own two 16-byte slices of each of the four rows.
constexpr int kTmaStages = 2;   // x 4 rows x 1 KB per warp: 64 KB per CTA -> 3 CTAs per SM (3 stages / 2 CTAs per SM measured 39.0 vs 34.6 ms)

// ---- epilogue: attention (optional LayerNorm, sigmoid, KxK mix, softmax), weighted sum, stores ----
// o[k][t]: the channel outputs O_k of this lane's 8 features (already relu'd as the variant asks;
// zeros for invalid rows).  Every lane of the warp must call it (group shuffles).
template <typename T, int FP, int MODE>
__device__ __forceinline__ void fwd_epilogue(const FwdParams& p, const int64_t row, const bool valid, const int gl,
                                             float (&o)[(MODE & 2) ? 4 : 3][8], const float* s_a,
                                             const float* s_avec, const float* s_ga, const float* s_sc) {
  constexpr int LANES = FP / 8;
  constexpr bool LN = (MODE & 1) != 0;
  constexpr int KMAX = (MODE & 2) ? 4 : 3;
  constexpr int K = KMAX;
  constexpr int TW = 2 * FP;
  // bf16 storage mode: the six exponentials and three reciprocals of the per-row attention are a
  // fifth of this (issue-bound) epilogue's instructions; the approximate forms (ex2.approx /
  // rcp.approx, ~2 ulp) are far below the bf16 rounding of the tables they are applied to.  The
  // fp32 parity mode keeps the IEEE forms.  The backward reads the saved att / sig, so forward and
  // backward stay consistent either way.
  constexpr bool FAST = sizeof(T) == 2;
  float z[KMAX];
  constexpr bool ln = LN;
  const float inv_f = 1.f / (float)p.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float dot = 0.f;
    if (!ln) {
      float ak[8];
      load_smem8(s_a + k * FP + gl * 8, ak);
#pragma unroll
      for (int t = 0; t < 8; ++t) dot = fmaf(o[k][t], ak[t], dot);
      z[k] = group_sum<LANES>(dot);
    } else {
      float s1 = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) s1 += o[k][t];
      const float mu = group_sum<LANES>(s1) * inv_f;
      float s2 = 0.f, gak[8];
      load_smem8(s_ga + k * FP + gl * 8, gak);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float d = (gl * 8 + t < p.f) ? o[k][t] - mu : 0.f;
        s2 = fmaf(d, d, s2);
        dot = fmaf(d, gak[t], dot);
      }
      const float var = group_sum<LANES>(s2) * inv_f;
      dot = group_sum<LANES>(dot);
      z[k] = (FAST ? dot * rsqrtf(var + kLnEps) : dot / sqrtf(var + kLnEps)) + s_sc[k];
    }
  }
  float s[KMAX], a[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) s[k] = FAST ? __fdividef(1.f, 1.f + __expf(-z[k])) : sigmoidf_acc(z[k]);
  float mx = -INFINITY;
  const float inv_k = 1.f / (float)K;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j)
      if (j < K) l = fmaf(s[j], s_avec[j * 4 + k], l);
    a[k] = (k < K) ? l * inv_k : -INFINITY;
    mx = fmaxf(mx, a[k]);
  }
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    a[k] = (k < K) ? (FAST ? __expf(a[k] - mx) : expf(a[k] - mx)) : 0.f;
    den += a[k];
  }
  const float rden = FAST ? __fdividef(1.f, den) : 1.f / den;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) a[k] *= rden;

  if (!valid) return;

  float yv[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float acc = a[0] * o[0][t];
#pragma unroll
    for (int k = 1; k < KMAX; ++k) acc = fmaf(a[k], o[k][t], acc);
    yv[t] = p.out_scale * acc;
  }
  const int f0 = gl * 8;
  if (p.y_bf16) {
    // bf16 inter-layer activations (the next layer's bf16 cast folded in here)
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy + f0;
    if (p.vec_y && f0 + 8 <= p.f) {
      *reinterpret_cast<uint4*>(yb) = pack_bf16x8(yv);
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (f0 + t < p.f) yb[t] = __float2bfloat16_rn(yv[t]);
    }
  }
  float* yr = p.y + row * p.ldy + f0;
  if (p.y_bf16) {
  } else if (p.vec_y && f0 + 8 <= p.f) {
    *reinterpret_cast<float4*>(yr) = make_float4(yv[0], yv[1], yv[2], yv[3]);
    *reinterpret_cast<float4*>(yr + 4) = make_float4(yv[4], yv[5], yv[6], yv[7]);
  } else {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      if (f0 + t < p.f) yr[t] = yv[t];
  }
  if (p.o_save) {
    T* os = reinterpret_cast<T*>(p.o_save) + row * TW + f0;
    Slice8<T>::store(os, o[0]);
    Slice8<T>::store(os + FP, o[1]);
  }
  if (gl == 0) {
    for (int k = 0; k < K; ++k) {
      p.att[row * K + k] = a[k];
      if (p.sig) p.sig[row * K + k] = s[k];
    }
  }
}

// GM: 0 register-staged LDG gather, 1 cp.async (LDGSTS) ring, 2 cp.async.bulk ring (FP = 256),
//     3 LDG gather with the L2::64B prefetch-size hint (narrow rows, FP <= 32),
//     4 no gather at all: pre-aggregated mode (own kernel instantiation, so that its prefetch
//       registers do not count against the occupancy of the gather kernels)
template <typename T, int FP, int MODE, int GM>
__device__ __forceinline__ void fwd_row_block(const FwdParams& p, const int64_t rb, const float* s_a,
                                              const float* s_avec, const float* s_ga, const float* s_sc,
                                              uint8_t* s_ring, const CUtensorMap* tmap = nullptr) {
  constexpr int LANES = FP / 8;
  constexpr int RPW = 32 / LANES;
  constexpr bool LN = (MODE & 1) != 0;   // LayerNorm of the attention logits live (ACM-Geometric flavour)
  constexpr bool K4 = (MODE & 2) != 0;   // 4th (structure) channel
  constexpr int KMAX = K4 ? 4 : 3;
  constexpr int K = KMAX;
  constexpr int TW = 2 * FP;  // table row width (elements)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / LANES;
  const int gl = lane % LANES;
  const int64_t slot_id = (rb * kFwdWarps + warp) * RPW + sub;
  const bool valid = slot_id < p.n_rows;
  // several rows share a warp (RPW > 1): the warp walks max(degree) edges, so the caller may hand
  // in a processing order that puts rows of equal degree next to each other (acm_b200.h)
  const int64_t row = (RPW > 1 && valid && p.row_order) ? (int64_t)__ldg(p.row_order + slot_id) : slot_id;

  int64_t e = 0, e1 = 0;
  if (valid) {
    e = __ldg(p.rowptr + row);
    e1 = __ldg(p.rowptr + row + 1);
  }
  const T* __restrict__ tab = reinterpret_cast<const T*>(p.table) + gl * 8;
  const int32_t* __restrict__ col = p.col;
  const float* __restrict__ val = p.val;

  float accL[8], accH[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) accL[t] = accH[t] = 0.f;

  if (p.lr.rows != nullptr && e1 - e > kLongRow) {
    // long row: sums were produced by the segment-parallel pass
    const float* a = p.lr.acc + (int64_t)find_long_row(p.lr, row) * TW + gl * 8;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      accL[t] = a[t];
      accH[t] = a[FP + t];
    }
    e = e1;
  }

  if (GM == 2 && LANES == 32 && e < e1) {
    constexpr int ST = AsyncCfg<T>::kStages;
    constexpr uint32_t ROWB = TW * (uint32_t)sizeof(T);   // one table row = one ring slot
    constexpr int SB = 8 * (int)sizeof(T);
    uint8_t* ring = s_ring + warp * (ST * ROWB);
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    // the mbarriers follow the kFwdWarps rings
    const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(s_ring + kFwdWarps * (ST * ROWB)) + warp * (ST * 8);
    const T* __restrict__ tab0 = reinterpret_cast<const T*>(p.table);
    const int n_e = (int)(e1 - e);
    // column indices / weights travel in registers, 32 edges per coalesced load, broadcast by shuffle
    int32_t ccur = (lane < n_e) ? __ldg(col + e + lane) : 0;
    float wcur = (val && lane < n_e) ? __ldg(val + e + lane) : 1.f;
#pragma unroll
    for (int st = 0; st < ST; ++st) {
      const int32_t c = __shfl_sync(0xffffffffu, ccur, st);
      if (lane == 0 && st < n_e) {
        mbar_expect_tx(bar_u32 + st * 8, ROWB);
        bulk_copy_g2s(ring_u32 + st * ROWB, tab0 + (int64_t)c * TW, ROWB, bar_u32 + st * 8);
      }
    }
    for (int i = 0; i < n_e; ++i) {
      const int slot = i & (ST - 1);
      if ((i & 31) == 0 && i) wcur = (val && i + lane < n_e) ? __ldg(val + e + i + lane) : 1.f;
      const int j = i + ST;                            // edge whose copy re-fills this slot
      if ((j & 31) == 0) ccur = (j + lane < n_e) ? __ldg(col + e + j + lane) : 0;
      const float w = __shfl_sync(0xffffffffu, wcur, i & 31);
      const int32_t cn = __shfl_sync(0xffffffffu, ccur, j & 31);
      mbar_wait(bar_u32 + slot * 8, (i / ST) & 1);
      Slice8<T> vl, vh;
      vl.load_plain(reinterpret_cast<const T*>(ring + slot * ROWB + lane * SB));
      vh.load_plain(reinterpret_cast<const T*>(ring + slot * ROWB + FP * sizeof(T) + lane * SB));
      float fl[8], fh[8];
      vl.to_float(fl);
      vh.to_float(fh);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        accL[t] = fmaf(w, fl[t], accL[t]);
        accH[t] = fmaf(w, fh[t], accH[t]);
      }
      __syncwarp();                                    // every lane has read the slot
      if (lane == 0 && j < n_e) {
        mbar_expect_tx(bar_u32 + slot * 8, ROWB);
        bulk_copy_g2s(ring_u32 + slot * ROWB, tab0 + (int64_t)cn * TW, ROWB, bar_u32 + slot * 8);
      }
    }
    e = e1;
  }

  if (GM == 5 && LANES == 32 && sizeof(T) == 2 && e < e1) {
    constexpr int ST = kTmaStages;
    constexpr uint32_t STAGE = 4 * TW * (uint32_t)sizeof(T);      // four table rows = 4 KB
    uint8_t* ring = s_ring + warp * (ST * STAGE);
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(s_ring + kFwdWarps * (ST * STAGE)) + warp * (ST * 8);
    const int n_e = (int)(e1 - e);
    const int n_grp = (n_e + 3) >> 2;
    // column indices / weights travel in registers, 32 edges (8 groups) per coalesced load, broadcast by shuffle;
    // edges past the end of the row gather row 0 with weight 0
    int32_t ccur = (lane < n_e) ? __ldg(col + e + lane) : 0;
    float wcur = (lane < n_e) ? (val ? __ldg(val + e + lane) : 1.f) : 0.f;
    int32_t cnext = ccur;                                          // indices of the 32 edges the issue side is in
    auto issue = [&](int grp, int stage) {                         // all lanes call it (shuffles), lane 0 issues
      const int b = (grp * 4) & 31;
      const int r0 = __shfl_sync(0xffffffffu, cnext, b), r1 = __shfl_sync(0xffffffffu, cnext, b + 1);
      const int r2 = __shfl_sync(0xffffffffu, cnext, b + 2), r3 = __shfl_sync(0xffffffffu, cnext, b + 3);
      if (lane == 0) {
        mbar_expect_tx(bar_u32 + stage * 8, STAGE);
        tma_gather4(ring_u32 + stage * STAGE, tmap, r0, r1, r2, r3, bar_u32 + stage * 8);
      }
    };
#pragma unroll
    for (int st = 0; st < ST; ++st)
      if (st < n_grp) issue(st, st);                               // ST * 4 <= 32: all inside the first index chunk
    for (int g = 0; g < n_grp; ++g) {
      const int stage = g % ST;
      if (((g * 4) & 31) == 0 && g) {
        const int i0 = g * 4;
        wcur = (i0 + lane < n_e) ? (val ? __ldg(val + e + i0 + lane) : 1.f) : 0.f;
      }
      mbar_wait(bar_u32 + stage * 8, (g / ST) & 1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = __shfl_sync(0xffffffffu, wcur, (g * 4 + u) & 31);
        Slice8<T> vl, vh;
        vl.load_plain(reinterpret_cast<const T*>(ring + stage * STAGE + u * (TW * sizeof(T)) + lane * 16));
        vh.load_plain(reinterpret_cast<const T*>(ring + stage * STAGE + u * (TW * sizeof(T)) + FP * sizeof(T) + lane * 16));
        float fl[8], fh[8];
        vl.to_float(fl);
        vh.to_float(fh);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          accL[t] = fmaf(w, fl[t], accL[t]);
          accH[t] = fmaf(w, fh[t], accH[t]);
        }
      }
      __syncwarp();                                                // every lane has read the stage
      const int j = g + ST;                                        // group whose rows re-fill this stage
      if (j < n_grp) {
        if (((j * 4) & 31) == 0) cnext = ((j * 4) + lane < n_e) ? __ldg(col + e + j * 4 + lane) : 0;
        issue(j, stage);
      }
    }
    e = e1;
  }

  if (GM == 1 && e < e1) {
    constexpr int ST = AsyncCfg<T>::kStages;
    constexpr int SB = 8 * (int)sizeof(T);           // bytes of one 8-feature slice
    constexpr int STAGE_BYTES = 64 * SB;             // per warp: 32 L slices then 32 H slices
    uint8_t* ring = s_ring + warp * (ST * STAGE_BYTES);
    const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring) + lane * SB;
    const int n_e = (int)(e1 - e);
#pragma unroll
    for (int st = 0; st < ST; ++st) {
      if (st < n_e) {
        const T* r = tab + (int64_t)__ldg(col + e + st) * TW;
        cp_async_slice<T>(ring_u32 + st * STAGE_BYTES, r);
        cp_async_slice<T>(ring_u32 + st * STAGE_BYTES + 32 * SB, r + FP);
      }
      cp_async_commit();
    }
    for (int i = 0; i < n_e; ++i) {
      cp_async_wait<ST - 1>();                        // the oldest group (edge i) has landed
      const float w = val ? __ldg(val + e + i) : 1.f;
      const int slot = i & (ST - 1);
      Slice8<T> vl, vh;
      vl.load_plain(reinterpret_cast<const T*>(ring + slot * STAGE_BYTES + lane * SB));
      vh.load_plain(reinterpret_cast<const T*>(ring + slot * STAGE_BYTES + 32 * SB + lane * SB));
      float fl[8], fh[8];
      vl.to_float(fl);
      vh.to_float(fh);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        accL[t] = fmaf(w, fl[t], accL[t]);
        accH[t] = fmaf(w, fh[t], accH[t]);
      }
      if (i + ST < n_e) {
        const T* r = tab + (int64_t)__ldg(col + e + i + ST) * TW;
        cp_async_slice<T>(ring_u32 + slot * STAGE_BYTES, r);
        cp_async_slice<T>(ring_u32 + slot * STAGE_BYTES + 32 * SB, r + FP);
      }
      cp_async_commit();
    }
    e = e1;
  }

  // ---- gather loop: kUnroll independent neighbour rows in flight per lane ----------------
  for (; e + kUnroll <= e1; e += kUnroll) {
    int32_t c[kUnroll];
    float w[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      c[u] = __ldg(col + e + u);
      w[u] = val ? __ldg(val + e + u) : 1.f;
    }
    Slice8<T> vl[kUnroll], vh[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const T* r = tab + (int64_t)c[u] * TW;
      gather_load<GM == 3>(vl[u], r);
      gather_load<GM == 3>(vh[u], r + FP);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float fl[8], fh[8];
      vl[u].to_float(fl);
      vh[u].to_float(fh);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        accL[t] = fmaf(w[u], fl[t], accL[t]);
        accH[t] = fmaf(w[u], fh[t], accH[t]);
      }
    }
  }
  for (; e < e1; ++e) {
    const int32_t c = __ldg(col + e);
    const float w = val ? __ldg(val + e) : 1.f;
    const T* r = tab + (int64_t)c * TW;
    Slice8<T> vl, vh;
    gather_load<GM == 3>(vl, r);
    gather_load<GM == 3>(vh, r + FP);
    float fl[8], fh[8];
    vl.to_float(fl);
    vh.to_float(fh);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      accL[t] = fmaf(w, fl[t], accL[t]);
      accH[t] = fmaf(w, fh[t], accH[t]);
    }
  }

  // ---- epilogue: channels, attention, mix --------------------------------------------------
  float o[KMAX][8];
  {
    float hs[8], hi[8];
    if (valid) {
      Slice8<T> a, b;
      a.load(tab + (p.row0 + row) * TW + FP);
      b.load(reinterpret_cast<const T*>(p.h_i) + row * FP + gl * 8);
      a.to_float(hs);
      b.to_float(hi);
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) hs[t] = hi[t] = 0.f;
    }
    const float rs = (valid && p.rowscale) ? __ldg(p.rowscale + row) : 1.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float sl = rs * accL[t];
      float sh = hs[t] - rs * accH[t];
      if (!p.variant) {
        sl = fmaxf(sl, 0.f);
        sh = fmaxf(sh, 0.f);
      }
      o[0][t] = sl;
      o[1][t] = sh;
      o[2][t] = fmaxf(hi[t], 0.f);
    }
    if (K4) {
      if (valid) {
        Slice8<T> s4;
        s4.load(reinterpret_cast<const T*>(p.o_s) + row * FP + gl * 8);
        s4.to_float(o[KMAX - 1]);
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) o[KMAX - 1][t] = 0.f;
      }
    }
  }

  fwd_epilogue<T, FP, MODE>(p, row, valid, gl, o, s_a, s_avec, s_ga, s_sc);
}

template <typename T, int FP, int MODE, int GM>
__global__ void __launch_bounds__(kFwdWarps * 32, GM == 4 ? 3 : 1) spmm_mix_fwd_kernel(const FwdParams p) {
  constexpr bool LN = (MODE & 1) != 0;
  constexpr int KMAX = (MODE & 2) ? 4 : 3;

  extern __shared__ float smem[];
  float* s_a = smem;                 // [KMAX][FP]   a_k
  float* s_avec = s_a + KMAX * FP;   // [16]
  float* s_ga = s_avec + 16;         // LN: [KMAX][FP] gamma*a
  float* s_sc = s_ga + (LN ? KMAX * FP : 0);  // LN: [8] sum(beta*a) per channel

  for (int i = threadIdx.x; i < KMAX * FP; i += blockDim.x) s_a[i] = p.pack[i];
  if (threadIdx.x < 16) s_avec[threadIdx.x] = p.pack[pack_off_avec(FP) + threadIdx.x];
  if (LN) {
    // gamma*a and sum(beta*a) come precomputed in the derived tail of the pack (acm_pack_params)
    for (int i = threadIdx.x; i < KMAX * FP; i += blockDim.x) s_ga[i] = p.pack[pack_off_ga(FP, 0) + i];
    if (threadIdx.x < 4) s_sc[threadIdx.x] = p.pack[pack_off_sum_ba(FP) + threadIdx.x];
  }
  __syncthreads();

  constexpr int RPWk = 32 / (FP / 8);
  if constexpr (GM == 4) {
    // pre-aggregated (aggregate-first) mode: the own rows of `table` hold [S_L|S_H]; no gather, pure
    // streaming.  A block's rows are too little work to amortise the parameter-pack load above ->
    // grid-stride over row blocks, with the NEXT row's three slices prefetched into registers
    // before the epilogue of the current one (without it each warp exposes a full DRAM latency
    // per row: 3.6 TB/s measured; the epilogue is ~300 instructions deep).
    constexpr int LANESk = FP / 8;
    constexpr bool K4k = (MODE & 2) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LANESk, gl = lane % LANESk;
    const T* tab = reinterpret_cast<const T*>(p.table) + gl * 8;
    const T* hib = reinterpret_cast<const T*>(p.h_i) + gl * 8;
    const T* osb = reinterpret_cast<const T*>(p.o_s) + gl * 8;
    const int64_t n_blocks = (p.n_rows + (int64_t)kFwdWarps * RPWk - 1) / ((int64_t)kFwdWarps * RPWk);
    Slice8<T> nL, nH, nI, nS;
    int64_t nrow = 0;
    bool nvalid = false;
    auto fetch = [&](const int64_t rb) {
      nrow = (rb * kFwdWarps + warp) * RPWk + sub;
      nvalid = nrow < p.n_rows;
      if (nvalid) {
        const T* r = tab + (p.row0 + nrow) * (2 * FP);
        nL.load(r);
        nH.load(r + FP);
        nI.load(hib + nrow * FP);
        if (K4k) nS.load(osb + nrow * FP);
      }
    };
    int64_t rb = blockIdx.x;
    if (rb < n_blocks) fetch(rb);
    for (; rb < n_blocks; rb += gridDim.x) {
      const Slice8<T> cL = nL, cH = nH, cI = nI, cS = nS;
      const int64_t row = nrow;
      const bool valid = nvalid;
      if (rb + gridDim.x < n_blocks) fetch(rb + gridDim.x);
      float o[KMAX][8];
      if (valid) {
        cL.to_float(o[0]);
        cH.to_float(o[1]);
        cI.to_float(o[2]);
        if (K4k) cS.to_float(o[KMAX - 1]);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (!p.variant) {
            o[0][t] = fmaxf(o[0][t], 0.f);
            o[1][t] = fmaxf(o[1][t], 0.f);
          }
          o[2][t] = fmaxf(o[2][t], 0.f);
        }
      } else {
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
#pragma unroll
          for (int t = 0; t < 8; ++t) o[k][t] = 0.f;
      }
      fwd_epilogue<T, FP, MODE>(p, row, valid, gl, o, s_a, s_avec, s_ga, s_sc);
    }
  } else {
    // gather mode: one row block per CTA; the hardware block scheduler balances the degrees
    // the cp.async ring follows the parameter pack in dynamic shared memory (16-byte aligned)
    constexpr int kPackFloats = KMAX * FP + 16 + (LN ? KMAX * FP + 8 : 0);
    uint8_t* s_ring = reinterpret_cast<uint8_t*>(smem + ((kPackFloats + 3) & ~3));
    if (GM == 2) {
      // one mbarrier per ring slot per warp; the __syncthreads below publishes the inits
      constexpr int ST = AsyncCfg<T>::kStages;
      if (threadIdx.x < kFwdWarps * ST)
        mbar_init((uint32_t)__cvta_generic_to_shared(s_ring + kFwdWarps * ST * (2 * FP * sizeof(T))) + threadIdx.x * 8, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      __syncthreads();
    }
    fwd_row_block<T, FP, MODE, GM>(p, blockIdx.x, s_a, s_avec, s_ga, s_sc, s_ring);
  }
}

// gather mode 3: same row block, neighbour rows staged by the TMA engine (tile::gather4); FP = 256, bf16 tables
template <typename T, int FP, int MODE>
__global__ void __launch_bounds__(kFwdWarps * 32, 1)
spmm_mix_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmap, const FwdParams p) {
  constexpr bool LN = (MODE & 1) != 0;
  constexpr int KMAX = (MODE & 2) ? 4 : 3;
  extern __shared__ __align__(128) uint8_t smem_tma[];
  // the dynamic shared window is only guaranteed 16-byte aligned: round up to the 128 bytes a TMA destination needs
  float* smem = reinterpret_cast<float*>(smem_tma + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_tma) & 127u)) & 127u));
  float* s_a = smem;
  float* s_avec = s_a + KMAX * FP;
  float* s_ga = s_avec + 16;
  float* s_sc = s_ga + (LN ? KMAX * FP : 0);
  for (int i = threadIdx.x; i < KMAX * FP; i += blockDim.x) s_a[i] = p.pack[i];
  if (threadIdx.x < 16) s_avec[threadIdx.x] = p.pack[pack_off_avec(FP) + threadIdx.x];
  if (LN) {
    for (int i = threadIdx.x; i < KMAX * FP; i += blockDim.x) s_ga[i] = p.pack[pack_off_ga(FP, 0) + i];
    if (threadIdx.x < 4) s_sc[threadIdx.x] = p.pack[pack_off_sum_ba(FP) + threadIdx.x];
  }
  constexpr int kPackFloats = KMAX * FP + 16 + (LN ? KMAX * FP + 8 : 0);
  uint8_t* s_ring = reinterpret_cast<uint8_t*>(smem + ((kPackFloats + 31) & ~31));     // 128-byte aligned stages
  constexpr uint32_t STAGE = 4 * 2 * FP * (uint32_t)sizeof(T);
  if (threadIdx.x < kFwdWarps * kTmaStages)
    mbar_init((uint32_t)__cvta_generic_to_shared(s_ring + kFwdWarps * kTmaStages * STAGE) + threadIdx.x * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  fwd_row_block<T, FP, MODE, 5>(p, blockIdx.x, s_a, s_avec, s_ga, s_sc, s_ring, &tmap);
}

int g_gather_mode = 3;  // 0: LDG register staging, 1: cp.async ring, 2: cp.async.bulk ring, 3 (default): TMA tile::gather4 where it applies (width 256, bf16), else 1

template <typename K>
static int raise_smem(K kernel, size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("spmm_mix_fwd: cannot raise dynamic shared memory to %zu: %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

template <typename T, int FP, int MODE>
static int launch_fwd(const FwdParams& p, cudaStream_t st) {
  constexpr int LANES = FP / 8;
  constexpr int RPB = (32 / LANES) * kFwdWarps;
  int64_t blocks = (p.n_rows + RPB - 1) / RPB;
  if (blocks == 0) return 0;
  if (p.pre_agg && blocks > 148 * 16) blocks = 148 * 16;
  ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_mix_fwd: too many rows for one launch");
  constexpr int KMAXh = (MODE & 2) ? 4 : 3;
  constexpr int kPackFloats = KMAXh * FP + 16 + ((MODE & 1) ? KMAXh * FP + 8 : 0);
  // the ring pays off for wide rows (FP = 256: 97 % vs 91 % of HBM peak); for narrow rows (FP = 16,
  // two lanes per row) the per-edge commit/wait overhead costs more than it hides (measured 5.4 vs
  // 4.8 ms), so they keep the register-staged LDG loop
  if constexpr (FP == 256 && sizeof(T) == 2) {
    if (!p.pre_agg && g_gather_mode >= 3) {
      // north-star wording: neighbour rows staged in shared memory by the TMA engine (4 rows per request)
      CUtensorMap tm;
      // the row extent only bounds the coordinates the engine accepts: column ids are < the number of table rows by
      // construction of the CSR, so the largest extent the descriptor can hold is as good as the true one
      if (int rc = tma_encode_2d_u32(&tm, p.table, 256, 0x7fffffffull, 1024, 256, 1, "gather table")) return rc;
      size_t smem_t = sizeof(float) * ((kPackFloats + 31) & ~31) + (size_t)kFwdWarps * kTmaStages * (4 * 2 * FP * sizeof(T) + 8) + 128;
      if (int rc = raise_smem(spmm_mix_fwd_tma_kernel<T, FP, MODE>, smem_t)) return rc;
      spmm_mix_fwd_tma_kernel<T, FP, MODE><<<(unsigned)blocks, kFwdWarps * 32, smem_t, st>>>(tm, p);
      ACM_LAUNCH_CHECK("spmm_mix_fwd (TMA gather)");
      return 0;
    }
  }
  const int gm = p.pre_agg ? 4
                 : (g_gather_mode == 2 && FP == 256) ? 2
                 : (g_gather_mode >= 1 && FP >= 64) ? 1 : 0;
  size_t smem = sizeof(float) * ((kPackFloats + 3) & ~3);
  if (gm == 2) {
    constexpr int GMB = FP == 256 ? 2 : 1;   // the bulk ring is only instantiated for FP = 256
    smem += (size_t)kFwdWarps * AsyncCfg<T>::kStages * (2 * FP * sizeof(T) + 8);
    if (int rc = raise_smem(spmm_mix_fwd_kernel<T, FP, MODE, GMB>, smem)) return rc;
    spmm_mix_fwd_kernel<T, FP, MODE, GMB><<<(unsigned)blocks, kFwdWarps * 32, smem, st>>>(p);
  } else if (gm == 1) {
    smem += (size_t)kFwdWarps * AsyncCfg<T>::kStages * 64 * 8 * sizeof(T);
    if (int rc = raise_smem(spmm_mix_fwd_kernel<T, FP, MODE, 1>, smem)) return rc;
    spmm_mix_fwd_kernel<T, FP, MODE, 1><<<(unsigned)blocks, kFwdWarps * 32, smem, st>>>(p);
  } else if (gm == 4) {
    spmm_mix_fwd_kernel<T, FP, MODE, 4><<<(unsigned)blocks, kFwdWarps * 32, smem, st>>>(p);
  } else if (FP <= 32 && g_narrow_row_hint) {
    constexpr int GMH = FP <= 32 ? 3 : 0;    // the hinted loop is only instantiated for narrow rows
    spmm_mix_fwd_kernel<T, FP, MODE, GMH><<<(unsigned)blocks, kFwdWarps * 32, smem, st>>>(p);
  } else {
    spmm_mix_fwd_kernel<T, FP, MODE, 0><<<(unsigned)blocks, kFwdWarps * 32, smem, st>>>(p);
  }
  ACM_LAUNCH_CHECK("spmm_mix_fwd");
  return 0;
}

template <typename T, int FP>
static int launch_fwd_mode(const FwdParams& p, int mode, cudaStream_t st) {
  switch (mode) {
    case 0: return launch_fwd<T, FP, 0>(p, st);
    case 1: return launch_fwd<T, FP, 1>(p, st);
    case 2: return launch_fwd<T, FP, 2>(p, st);
    default: return launch_fwd<T, FP, 3>(p, st);
  }
}

}  // namespace acm

extern "C" int acm_spmm_mix_fwd(int dtype, int fp, int f, int64_t n_rows, int64_t row0,
                                const int64_t* rowptr, const int32_t* col, const float* val, const float* rowscale,
                                const void* table, const void* h_i, const void* o_s,
                                const float* pack, int k_channels, int ln_live, int variant, float out_scale,
                                void* y, int y_dtype, int64_t ldy, void* o_save, float* att, float* sig,
                                const int32_t* long_rows, int n_long, const float* long_acc,
                                const int32_t* row_order, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "spmm_mix_fwd: bad dtype %d", dtype);
  ACM_CHECK_ARG(k_channels == 3 || k_channels == 4, "spmm_mix_fwd: k_channels must be 3 or 4");
  ACM_CHECK_ARG(f >= 1 && f <= fp, "spmm_mix_fwd: need 1 <= f <= fp");
  ACM_CHECK_ARG(k_channels == 3 || o_s != nullptr, "spmm_mix_fwd: 4 channels need o_s");
  ACM_CHECK_ARG(table && h_i && pack && y && att, "spmm_mix_fwd: null pointer");
  ACM_CHECK_ARG((rowptr && col) || (!rowptr && !col), "spmm_mix_fwd: rowptr and col must both be given or both be NULL");
  FwdParams p;
  p.n_rows = n_rows; p.row0 = row0; p.rowptr = rowptr; p.col = col; p.val = val; p.rowscale = rowscale;
  p.table = table; p.h_i = h_i; p.o_s = o_s; p.pack = pack; p.k = k_channels; p.ln = ln_live;
  ACM_CHECK_ARG(y_dtype == ACM_F32 || y_dtype == ACM_BF16, "spmm_mix_fwd: bad y dtype %d", y_dtype);
  p.variant = variant; p.f = f; p.out_scale = out_scale; p.y = reinterpret_cast<float*>(y); p.ldy = ldy; p.o_save = o_save;
  p.y_bf16 = (y_dtype == ACM_BF16);
  p.att = att; p.sig = sig;
  p.row_order = row_order;
  p.pre_agg = (rowptr == nullptr);
  ACM_CHECK_ARG(!p.pre_agg || (rowscale == nullptr && n_long == 0), "spmm_mix_fwd: the pre-aggregated mode takes no rowscale / long rows");
  p.lr.rows = n_long > 0 ? long_rows : nullptr; p.lr.acc = long_acc; p.lr.n_long = n_long;
  ACM_CHECK_ARG(n_long == 0 || (long_rows && long_acc), "spmm_mix_fwd: long rows need long_rows and long_acc");
  p.vec_y = p.y_bf16 ? ((f % 8 == 0) && (ldy % 8 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0))
                     : ((f % 4 == 0) && (ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int mode = (ln_live ? 1 : 0) | (k_channels == 4 ? 2 : 0);
  if (dtype == ACM_BF16) {
    ACM_DISPATCH_FP(fp, return launch_fwd_mode<__nv_bfloat16, FP>(p, mode, st));
  } else {
    ACM_DISPATCH_FP(fp, return launch_fwd_mode<float, FP>(p, mode, st));
  }
  return 0;
}

extern "C" int acm_set_gather_mode(int mode) {
  if (mode < 0 || mode > 4) {
    acm::set_error("gather mode must be 0 (LDG), 1 (cp.async ring), 2 (cp.async.bulk ring), 3 (TMA tile::gather4) or 4 (3 + aggregate-first gather)");
    return ACM_ERR_BAD_ARG;
  }
  acm::g_gather_mode = mode;
  return 0;
}
