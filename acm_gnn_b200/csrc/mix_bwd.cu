// Row-local backward of the channel attention / mix / relu part of the ACM layer
// (autograd of ACM-Pytorch/models/layers.py:94-152,185-204; math spec in DESIGN.md):
//   d_alpha_k = c * <G, O_k> ; d_logit = alpha * (d_alpha - <alpha, d_alpha>) ;
//   d_s = d_logit . att_vec^T / K ; d_att_vec += s^T d_logit / K ; d_z = d_s s (1-s)
//   d_O_k = c alpha_k G + d_z_k a_k   (through LayerNorm backward when it is live)
//   d_a_k += O_k^T d_z_k   (LayerNorm(O_k) when live; plus d_gamma, d_beta)
//   variant 0: d_S_k = d_O_k * [O_k > 0]          variant 1: d_S_k = d_O_k
// Same lane mapping as the forward kernel (LANES = FP/8 lanes per row, 8 features per lane), no
// gather: pure streaming, HBM bound.  Parameter gradients are accumulated in registers over a
// grid-stride loop, reduced through shared memory and flushed with one global atomicAdd per
// element per block.
//
// LayerNorm (ACM-Geometric flavour): with xhat = (O - mu)/sigma, y = gamma xhat + beta, z = y.a
// the three per-feature parameter gradients need only ONE accumulator T_k[f] = sum_i dz_k,i xhat_i,f
// and one scalar S_k = sum_i dz_k,i per channel:
//   d_a_k[f] = gamma_f T_k[f] + beta_f S_k ;  d_gamma_k[f] = a_f T_k[f] ;  d_beta_k[f] = a_f S_k
// so the LayerNorm mode costs the same registers as the plain mode.
//
// MODE bit 0: LayerNorm live; bit 1: 4 channels (structure channel).
#include "acm_common.cuh"

namespace acm {

struct BwdParams {
  int64_t n_rows;
  const float* g;
  int64_t ldg;
  const void* o_lh;
  const void* h_i;
  const void* o_s;
  const float* att;
  const float* sig;
  const float* pack;
  int k, ln, variant, f, vec_g, g_bf16, ring;
  int gtab;           // table_mode 1: t_lh = T g[table_rows][FP] then float4 {c att_L, c att_H, dz_L, dz_H}[table_rows]
  int64_t table_rows; //   (spmm_t_rank1_kernel)
  float out_scale;
  void* t_lh;
  void* dh_all;
  void* dos_pre;
  float* dpack;
  PeerTables peers;   // optional push of the [dS_L|dS_H] rows into every rank's table
};

constexpr int kBwdWarps = 8;

// cp.async ring for the streaming inputs (bf16 tables, 3-channel modes): every lane prefetches
// its own slices of G, O_L, O_H and HI for the next kRingStages-1 rows of its lane group straight
// into shared memory, so ~4 rows per group are in flight instead of the one row that the ~125
// registers of this kernel leave room for (measured 9.8 -> 7.7 ms at the headline config).
constexpr int kRingStages = 4;
constexpr int kRingAsOff = 32 * (32 + 3 * 16);        // per warp: G (<=32 B/lane) | O_L | O_H | HI (16 B/lane each)
constexpr int kRingStageBytes = kRingAsOff + 4 * 32;  // ... | per row group: att[0..3] (16 B) sig[0..3] (16 B), <= 4 groups

__device__ __forceinline__ void bwd_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void bwd_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bwd_cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
template <int N> __device__ __forceinline__ void bwd_cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int FP, int MODE>
struct BwdSmem {
  static constexpr bool LN = (MODE & 1) != 0;
  static constexpr int KMAX = (MODE & 2) ? 4 : 3;
  // s_a [KMAX][FP] | s_avec [16] | s_dav [16] | s_da [KMAX][FP] | LN: gamma, beta [KMAX][FP] each | sums [16]
  static constexpr int kFloats = KMAX * FP * 2 + 32 + (LN ? 2 * KMAX * FP : 0) + 16;
  static constexpr int kFloatsPadded = (kFloats + 3) & ~3;
};

// MINB = minimum resident CTAs per SM asked of ptxas (2: ~125 registers, no spills)
// RING  = 0: inputs loaded in place; 1 / 2: cp.async input ring with fp32 / bf16 G (bf16 tables,
//         3 channels).  A compile-time choice: with both paths in one function every instruction
//         after their merge point waited on the scoreboards of EITHER path's loads.
template <typename T, int FP, int MODE, int MINB, int RING>
__global__ void __launch_bounds__(kBwdWarps * 32, MINB) mix_bwd_kernel(const BwdParams p) {
  constexpr int LANES = FP / 8;
  constexpr int RPW = 32 / LANES;
  constexpr int RPB = RPW * kBwdWarps;
  constexpr bool LN = (MODE & 1) != 0;
  constexpr bool K4 = (MODE & 2) != 0;
  constexpr int KMAX = K4 ? 4 : 3;
  constexpr int K = KMAX;
  constexpr int TW = 2 * FP;
  using SM = BwdSmem<FP, MODE>;

  extern __shared__ float smem[];
  float* s_a = smem;                            // [KMAX][FP]
  float* s_avec = s_a + KMAX * FP;              // [16]
  float* s_dav = s_avec + 16;                   // [16]  d att_vec
  float* s_da = s_dav + 16;                     // [KMAX][FP]  d a_k  (LN: T_k)
  float* s_gam = s_da + KMAX * FP;              // LN: gamma [KMAX][FP]
  float* s_bet = s_gam + (LN ? KMAX * FP : 0);  // LN: beta  [KMAX][FP]
  float* s_sum = s_bet + (LN ? KMAX * FP : 0);  // [0..4) sum_f gamma*a ; [4..8) S_k = sum_i dz_k,i

  for (int i = threadIdx.x; i < KMAX * FP; i += blockDim.x) {
    s_a[i] = p.pack[i];
    s_da[i] = 0.f;
    if (LN) {
      s_gam[i] = p.pack[pack_off_gamma(FP, 0) + i];
      s_bet[i] = p.pack[pack_off_beta(FP, 0) + i];
    }
  }
  if (threadIdx.x < 16) {
    s_avec[threadIdx.x] = p.pack[pack_off_avec(FP) + threadIdx.x];
    s_dav[threadIdx.x] = 0.f;
    s_sum[threadIdx.x] = 0.f;
  }
  __syncthreads();
  if (LN) {
    if (threadIdx.x < 4) s_sum[threadIdx.x] = p.pack[pack_off_sum_ga(FP) + threadIdx.x];  // precomputed (acm_pack_params)
    __syncthreads();
  }

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / LANES;
  const int gl = lane % LANES;
  const int f0 = gl * 8;
  const float c = p.out_scale;
  const float inv_k = 1.f / (float)K;
  const float inv_f = 1.f / (float)p.f;

  float da[KMAX][8];     // plain: sum_i dz O ; LN: T_k = sum_i dz xhat
  float dsum[KMAX];      // LN: S_k = sum_i dz_k,i (identical on every lane of a group)
  float dav[16];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    dsum[k] = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) da[k][t] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) dav[i] = 0.f;

  static_assert(RING == 0 || (!K4 && sizeof(T) == 2), "the input ring is for bf16 tables, 3 channels");
  constexpr bool ring = RING != 0;
  // ring mode: att / sig of a row (2K floats) travel through the ring too -- lane j of the row's
  // group copies att[j] (j < K) or sig[j - K] (K <= j < 2K) -- instead of being loaded in place,
  // where they were a dependent DRAM-latency stall per row
  constexpr bool AS_RING = ring && LANES >= 8;
  const bool g_bf16 = RING == 2 ? true : RING == 1 ? false : (p.g_bf16 != 0);
  const int64_t stride = (int64_t)gridDim.x * RPB;
  const int64_t base0 = (int64_t)blockIdx.x * RPB;
  const int64_t n_iter = base0 < p.n_rows ? (p.n_rows - base0 + stride - 1) / stride : 0;
  uint8_t* ring_w = nullptr;
  uint32_t ring_u32 = 0;
  auto ring_issue = [&](int64_t r, int slot) {
    const uint32_t d = ring_u32 + slot * kRingStageBytes;
    if (AS_RING) {
      if (gl < K) bwd_cp_async4(d + kRingAsOff + sub * 32 + gl * 4, p.att + r * K + gl);
      else if (gl < 2 * K) bwd_cp_async4(d + kRingAsOff + sub * 32 + 16 + (gl - K) * 4, p.sig + r * K + (gl - K));
    }
    if (g_bf16) {
      bwd_cp_async16(d + lane * 32, reinterpret_cast<const __nv_bfloat16*>(p.g) + r * p.ldg + f0);
    } else {
      const float* gr = p.g + r * p.ldg + f0;
      bwd_cp_async16(d + lane * 32, gr);
      bwd_cp_async16(d + lane * 32 + 16, gr + 4);
    }
    const T* ol = reinterpret_cast<const T*>(p.o_lh) + r * TW + f0;
    bwd_cp_async16(d + 1024 + lane * 16, ol);
    bwd_cp_async16(d + 1536 + lane * 16, ol + FP);
    bwd_cp_async16(d + 2048 + lane * 16, reinterpret_cast<const T*>(p.h_i) + r * FP + f0);
  };
  uint8_t* tail = reinterpret_cast<uint8_t*>(smem + SM::kFloatsPadded);  // ring, then push staging
  if (ring) {
    ring_w = tail + warp * (kRingStages * kRingStageBytes);
    ring_u32 = (uint32_t)__cvta_generic_to_shared(ring_w);
    tail += kBwdWarps * kRingStages * kRingStageBytes;
#pragma unroll
    for (int st = 0; st < kRingStages; ++st) {
      const int64_t r = base0 + st * stride + warp * RPW + sub;
      if (st < n_iter && r < p.n_rows) ring_issue(r, st);
      bwd_cp_async_commit();
    }
  }
  // staging tile of the narrow-row peer push
  T* push_stage = (p.peers.n > 0 && LANES < 32) ? reinterpret_cast<T*>(tail) : nullptr;

  for (int64_t it = 0; it < n_iter; ++it) {
    const int64_t row = base0 + it * stride + warp * RPW + sub;
    const bool valid = row < p.n_rows;
    float G[8], o[KMAX][8], al[KMAX], sg[KMAX];
#pragma unroll
    for (int t = 0; t < 8; ++t) G[t] = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      al[k] = 0.f;
      sg[k] = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) o[k][t] = 0.f;
    }
    if (ring) {
      bwd_cp_async_wait<kRingStages - 1>();
      if (AS_RING) __syncwarp();   // att / sig were copied by other lanes of the group
      const int slot = (int)(it & (kRingStages - 1));
      if (valid) {
        const uint8_t* d = ring_w + slot * kRingStageBytes;
        if (AS_RING) {
          const float4 a4 = *reinterpret_cast<const float4*>(d + kRingAsOff + sub * 32);
          const float4 s4 = *reinterpret_cast<const float4*>(d + kRingAsOff + sub * 32 + 16);
          al[0] = a4.x; al[1] = a4.y; al[2] = a4.z;
          sg[0] = s4.x; sg[1] = s4.y; sg[2] = s4.z;
        }
        if (g_bf16) {
          unpack_bf16x8(*reinterpret_cast<const uint4*>(d + lane * 32), G);
        } else {
          const float4 a = *reinterpret_cast<const float4*>(d + lane * 32);
          const float4 b = *reinterpret_cast<const float4*>(d + lane * 32 + 16);
          G[0] = a.x; G[1] = a.y; G[2] = a.z; G[3] = a.w;
          G[4] = b.x; G[5] = b.y; G[6] = b.z; G[7] = b.w;
        }
        uint4 vl = *reinterpret_cast<const uint4*>(d + 1024 + lane * 16);
        uint4 vh = *reinterpret_cast<const uint4*>(d + 1536 + lane * 16);
        if (!p.variant) {   // o_lh may hold the pre-relu [S_L|S_H] (aggregate-first order): relu on load
          vl = relu_bf16x8(vl);
          vh = relu_bf16x8(vh);
        }
        unpack_bf16x8(vl, o[0]);
        unpack_bf16x8(vh, o[1]);
        unpack_bf16x8(*reinterpret_cast<const uint4*>(d + 2048 + lane * 16), o[2]);
      }
      const int64_t rn = base0 + (it + kRingStages) * stride + warp * RPW + sub;
      if (it + kRingStages < n_iter && rn < p.n_rows) ring_issue(rn, slot);
      bwd_cp_async_commit();
    } else if (valid) {
      const float* gr = p.g + row * p.ldg + f0;
      if (g_bf16) {
        const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(p.g) + row * p.ldg + f0;
        if (p.vec_g && f0 + 8 <= p.f) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(gb));
          unpack_bf16x8(v, G);
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (f0 + t < p.f) G[t] = __bfloat162float(gb[t]);
        }
      } else if (p.vec_g && f0 + 8 <= p.f) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(gr));
        const float4 b = __ldg(reinterpret_cast<const float4*>(gr) + 1);
        G[0] = a.x; G[1] = a.y; G[2] = a.z; G[3] = a.w;
        G[4] = b.x; G[5] = b.y; G[6] = b.z; G[7] = b.w;
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t)
          if (f0 + t < p.f) G[t] = __ldg(gr + t);
      }
      Slice8<T> a, b, d;
      const T* ol = reinterpret_cast<const T*>(p.o_lh) + row * TW + f0;
      a.load(ol);
      b.load(ol + FP);
      d.load(reinterpret_cast<const T*>(p.h_i) + row * FP + f0);
      a.to_float(o[0]);
      b.to_float(o[1]);
      d.to_float(o[2]);
      if (!p.variant) {     // see the ring path: o_lh may be the pre-relu table
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          o[0][t] = fmaxf(o[0][t], 0.f);
          o[1][t] = fmaxf(o[1][t], 0.f);
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int t = 0; t < 8; ++t) o[2][t] = fmaxf(o[2][t], 0.f);
      if (K4) {
        Slice8<T> s4;
        s4.load(reinterpret_cast<const T*>(p.o_s) + row * FP + f0);
        s4.to_float(o[KMAX - 1]);
      }
      if (!AS_RING) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          al[k] = __ldg(p.att + row * K + k);
          sg[k] = __ldg(p.sig + row * K + k);
        }
      }
    }
    // d_alpha
    float dal[KMAX], dlog[KMAX], dz[KMAX];
    float adot = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      dal[k] = c * group_sum<LANES>(dot8(G, o[k]));
      adot = fmaf(al[k], dal[k], adot);
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) dlog[k] = al[k] * (dal[k] - adot);
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      float ds = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) ds = fmaf(dlog[k], s_avec[j * 4 + k], ds);
      ds *= inv_k;
      dz[j] = ds * sg[j] * (1.f - sg[j]);
      if (gl == 0 && valid) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) dav[j * 4 + k] = fmaf(sg[j], dlog[k] * inv_k, dav[j * 4 + k]);
      }
    }
    // d_O_k (+ parameter gradient partials)
    float dO[KMAX][8];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (!LN) {
        float ak[8];
        load_smem8(s_a + k * FP + f0, ak);
        axpy8(da[k], dz[k], o[k]);
        const float2 ca2 = make_float2(c * al[k], c * al[k]), dz2 = make_float2(dz[k], dz[k]);
#pragma unroll
        for (int t = 0; t < 8; t += 2) {
          const float2 r = ffma2(ca2, make_float2(G[t], G[t + 1]), fmul2(dz2, make_float2(ak[t], ak[t + 1])));
          dO[k][t] = r.x;
          dO[k][t + 1] = r.y;
        }
      } else {
        float s1 = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) s1 += o[k][t];
        const float mu = group_sum<LANES>(s1) * inv_f;
        float xh[8], s2 = 0.f;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          xh[t] = (f0 + t < p.f) ? o[k][t] - mu : 0.f;
          s2 = fmaf(xh[t], xh[t], s2);
        }
        const float rstd = 1.f / sqrtf(group_sum<LANES>(s2) * inv_f + kLnEps);
        float gx = 0.f, gak[8], ak[8];
        load_smem8(s_gam + k * FP + f0, gak);
        load_smem8(s_a + k * FP + f0, ak);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          gak[t] *= ak[t];
          xh[t] *= rstd;
          gx = fmaf(gak[t], xh[t], gx);
        }
        const float m1 = dz[k] * s_sum[k] * inv_f;
        const float m2 = dz[k] * group_sum<LANES>(gx) * inv_f;
        dsum[k] += dz[k];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float ga = gak[t];
          da[k][t] = fmaf(dz[k], xh[t], da[k][t]);  // T_k
          const float dln = (f0 + t < p.f) ? rstd * (dz[k] * ga - m1 - xh[t] * m2) : 0.f;
          dO[k][t] = fmaf(c * al[k], G[t], dln);
        }
      }
    }
    if (p.gtab) {
      // rank-structured table (variant 1, no LayerNorm): the G row + four scalars instead of [dO_L | dO_H]
      if (valid) {
        const float4 sc = make_float4(c * al[0], c * al[1], dz[0], dz[1]);
        const int64_t scal_off = p.table_rows * FP * (int64_t)sizeof(T);   // byte offset of the scalar region
        if (p.peers.n > 0) {
          const int64_t grow = p.peers.row_off + row;
          if (p.peers.mc) {
            uint4 pk[sizeof(T) == 2 ? 1 : 2];
            Slice8<T>::store(reinterpret_cast<T*>(pk), G);
#pragma unroll
            for (int h = 0; h < (int)(sizeof(pk) / 16); ++h)
              multimem_st16(reinterpret_cast<char*>(p.peers.mc) + (grow * FP + f0) * (int64_t)sizeof(T) + h * 16, pk[h]);
            if (gl == 0) multimem_st16(reinterpret_cast<char*>(p.peers.mc) + scal_off + grow * 16,
                                       make_uint4(__float_as_uint(sc.x), __float_as_uint(sc.y), __float_as_uint(sc.z), __float_as_uint(sc.w)));
          } else {
#pragma unroll 1
            for (int r = 0; r < p.peers.n; ++r) {
              char* tb = reinterpret_cast<char*>(p.peers.tables[r]);
              Slice8<T>::store(reinterpret_cast<T*>(tb) + grow * FP + f0, G);
              if (gl == 0) *reinterpret_cast<float4*>(tb + scal_off + grow * 16) = sc;
            }
          }
        } else {
          char* tb = reinterpret_cast<char*>(p.t_lh);
          Slice8<T>::store(reinterpret_cast<T*>(tb) + row * FP + f0, G);
          if (gl == 0) *reinterpret_cast<float4*>(tb + scal_off + row * 16) = sc;
        }
      }
    } else if (p.peers.n > 0 && LANES < 32) {
      // fused all-gather, narrow rows: a lane's 16-byte slices of one row are only 32..256 B apart
      // from the next lane group's, which would reach NVLink as small fragments.  The warp's RPW
      // consecutive rows (512*sizeof(T) bytes) are staged in shared memory and pushed to every
      // peer as fully contiguous 512-byte warp stores.
      T* stg = push_stage + warp * (RPW * TW);
      if (valid) {
        float out[8];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int t = 0; t < 8; ++t) out[t] = (p.variant || o[k][t] > 0.f) ? dO[k][t] : 0.f;
          Slice8<T>::store(stg + sub * TW + k * FP + f0, out);
        }
      }
      __syncwarp();
      const int64_t first_row = base0 + it * stride + warp * RPW;
      int64_t rows_valid = p.n_rows - first_row;
      if (rows_valid > RPW) rows_valid = RPW;
      if (rows_valid > 0) {
        const int n_vec = (int)(rows_valid * TW * sizeof(T) / 16);
        const uint4* src = reinterpret_cast<const uint4*>(stg);
        const int64_t off = (p.peers.row_off + first_row) * TW * (int64_t)sizeof(T);
        if (p.peers.mc) {
          for (int v = lane; v < n_vec; v += 32) multimem_st16(reinterpret_cast<char*>(p.peers.mc) + off + v * 16, src[v]);
        } else {
#pragma unroll 1
          for (int r = 0; r < p.peers.n; ++r) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<char*>(p.peers.tables[r]) + off);
            for (int v = lane; v < n_vec; v += 32) dst[v] = src[v];
          }
        }
      }
      __syncwarp();
    }
    if (valid) {
      float out[8];
      if (!p.gtab && !(p.peers.n > 0 && LANES < 32)) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int t = 0; t < 8; ++t) out[t] = (p.variant || o[k][t] > 0.f) ? dO[k][t] : 0.f;
          if (p.peers.n > 0) {
            // fused all-gather of the backward operand table (see PeerTables): wide rows, every
            // warp store already covers 512 contiguous bytes
            const int64_t off = (p.peers.row_off + row) * TW + f0 + k * FP;
            if (p.peers.mc) {
              // stage the slice in registers in storage format, then one (bf16) / two (fp32) multicast stores
              uint4 pk[sizeof(T) == 2 ? 1 : 2];
              Slice8<T>::store(reinterpret_cast<T*>(pk), out);
#pragma unroll
              for (int h = 0; h < (int)(sizeof(pk) / 16); ++h)
                multimem_st16(reinterpret_cast<char*>(p.peers.mc) + off * (int64_t)sizeof(T) + h * 16, pk[h]);
            } else {
#pragma unroll 1
              for (int r = 0; r < p.peers.n; ++r) Slice8<T>::store(reinterpret_cast<T*>(p.peers.tables[r]) + off, out);
            }
          } else {
            Slice8<T>::store(reinterpret_cast<T*>(p.t_lh) + row * TW + f0 + k * FP, out);
          }
        }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) out[t] = (o[2][t] > 0.f) ? dO[2][t] : 0.f;
      Slice8<T>::store(reinterpret_cast<T*>(p.dh_all) + row * (3 * FP) + 2 * FP + f0, out);
      if (K4) {
#pragma unroll
        for (int t = 0; t < 8; ++t) out[t] = (o[KMAX - 1][t] > 0.f) ? dO[KMAX - 1][t] : 0.f;
        Slice8<T>::store(reinterpret_cast<T*>(p.dos_pre) + row * FP + f0, out);
      }
    }
  }

  // ---- flush parameter gradients: registers -> shared -> one global atomic per element ----
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
#pragma unroll
    for (int t = 0; t < 8; ++t) atomicAdd(&s_da[k * FP + f0 + t], da[k][t]);
    if (LN && gl == 0) atomicAdd(&s_sum[4 + k], dsum[k]);
  }
  if (gl == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) atomicAdd(&s_dav[i], dav[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KMAX * FP; i += blockDim.x) {
    const float tv = s_da[i];
    if (!LN) {
      if (tv != 0.f) atomicAdd(p.dpack + i, tv);
    } else {
      const int k = i / FP, fidx = i - k * FP;
      const float sk = s_sum[4 + k];
      const float av = s_a[i];
      const float d_a = (fidx < p.f) ? fmaf(s_gam[i], tv, s_bet[i] * sk) : 0.f;
      const float d_g = av * tv, d_b = av * sk;
      if (d_a != 0.f) atomicAdd(p.dpack + i, d_a);
      if (d_g != 0.f) atomicAdd(p.dpack + pack_off_gamma(FP, 0) + i, d_g);
      if (d_b != 0.f) atomicAdd(p.dpack + pack_off_beta(FP, 0) + i, d_b);
    }
  }
  if (threadIdx.x < 16) {
    const float v = s_dav[threadIdx.x];
    if (v != 0.f) atomicAdd(p.dpack + pack_off_avec(FP) + threadIdx.x, v);
  }
}

static int g_mix_bwd_minb = 2;
static int g_mix_bwd_ring = 1;

template <typename T, int FP, int MODE, int MINB, int RING>
static int launch_bwd_ring(const BwdParams& p, cudaStream_t st) {
  constexpr int LANES = FP / 8;
  constexpr int RPB = (32 / LANES) * kBwdWarps;
  constexpr bool K4 = (MODE & 2) != 0;
  int64_t blocks = (p.n_rows + RPB - 1) / RPB;
  if (blocks == 0) return 0;
  // persistent-style grid: a few CTAs per SM, grid-stride over rows, so the number of
  // global atomics per parameter element stays ~ #CTAs
  const int64_t cap = 148 * (MINB == 3 ? 6 : 4);
  if (blocks > cap) blocks = cap;
  size_t smem = sizeof(float) * BwdSmem<FP, MODE>::kFloatsPadded;
  if (RING) smem += (size_t)kBwdWarps * kRingStages * kRingStageBytes;
  if (p.peers.n > 0 && LANES < 32) smem += (size_t)kBwdWarps * 512 * sizeof(T);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mix_bwd_kernel<T, FP, MODE, MINB, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mix_bwd: cannot raise dynamic shared memory to %zu: %s", smem, cudaGetErrorString(e));
      return (int)e;
    }
  }
  mix_bwd_kernel<T, FP, MODE, MINB, RING><<<(unsigned)blocks, kBwdWarps * 32, smem, st>>>(p);
  ACM_LAUNCH_CHECK("mix_bwd");
  return 0;
}

template <typename T, int FP, int MODE, int MINB>
static int launch_bwd_impl(const BwdParams& p, cudaStream_t st) {
  if constexpr (sizeof(T) == 2 && (MODE & 2) == 0) {
    if (p.ring) return p.g_bf16 ? launch_bwd_ring<T, FP, MODE, MINB, 2>(p, st) : launch_bwd_ring<T, FP, MODE, MINB, 1>(p, st);
  }
  return launch_bwd_ring<T, FP, MODE, MINB, 0>(p, st);
}

template <typename T, int FP>
static int launch_bwd(const BwdParams& p, int mode, cudaStream_t st) {
  switch (mode) {
    case 0: return g_mix_bwd_minb == 3 ? launch_bwd_impl<T, FP, 0, 3>(p, st) : launch_bwd_impl<T, FP, 0, 2>(p, st);
    case 1: return launch_bwd_impl<T, FP, 1, 2>(p, st);
    case 2: return launch_bwd_impl<T, FP, 2, 2>(p, st);
    default: return launch_bwd_impl<T, FP, 3, 2>(p, st);
  }
}

}  // namespace acm

extern "C" int acm_mix_bwd(int dtype, int fp, int f, int64_t n_rows,
                           const void* g, int g_dtype, int64_t ldg, const void* o_lh, const void* h_i, const void* o_s,
                           const float* att, const float* sig, const float* pack,
                           int k_channels, int ln_live, int variant, float out_scale,
                           void* t_lh, int table_mode, int64_t table_rows, void* dh_all, void* dos_pre, float* dpack,
                           void* const* peer_tables, int n_peers, int64_t peer_row_off, void* multicast_table,
                           void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(n_peers >= 0 && n_peers <= kMaxPeers, "mix_bwd: 0 <= n_peers <= %d", kMaxPeers);
  ACM_CHECK_ARG(n_peers == 0 || peer_tables, "mix_bwd: peer push needs peer_tables");
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "mix_bwd: bad dtype %d", dtype);
  ACM_CHECK_ARG(g_dtype == ACM_F32 || g_dtype == ACM_BF16, "mix_bwd: bad g dtype %d", g_dtype);
  ACM_CHECK_ARG(k_channels == 3 || k_channels == 4, "mix_bwd: k_channels must be 3 or 4");
  ACM_CHECK_ARG(f >= 1 && f <= fp, "mix_bwd: need 1 <= f <= fp");
  ACM_CHECK_ARG(k_channels == 3 || (o_s && dos_pre), "mix_bwd: 4 channels need o_s and dos_pre");
  ACM_CHECK_ARG(g && o_lh && h_i && att && sig && pack && (t_lh || n_peers > 0) && dh_all && dpack, "mix_bwd: null pointer");
  BwdParams p;
  p.n_rows = n_rows; p.g = reinterpret_cast<const float*>(g); p.ldg = ldg; p.g_bf16 = (g_dtype == ACM_BF16);
  p.o_lh = o_lh; p.h_i = h_i; p.o_s = o_s; p.att = att; p.sig = sig;
  p.pack = pack; p.k = k_channels; p.ln = ln_live; p.variant = variant; p.f = f; p.out_scale = out_scale;
  ACM_CHECK_ARG(table_mode == 0 || (table_mode == 1 && variant && !ln_live && fp >= 64),
                "mix_bwd: table_mode 1 (rank-structured table) needs variant 1 without LayerNorm and fp >= 64");
  ACM_CHECK_ARG(table_mode == 0 || table_rows >= peer_row_off + n_rows, "mix_bwd: table_rows must cover the rows written");
  p.t_lh = t_lh; p.dh_all = dh_all; p.dos_pre = dos_pre; p.dpack = dpack; p.gtab = table_mode; p.table_rows = table_rows;
  p.peers = PeerTables{};
  p.peers.n = n_peers; p.peers.row_off = peer_row_off; p.peers.mc = n_peers > 0 ? multicast_table : nullptr;
  for (int r = 0; r < n_peers; ++r) p.peers.tables[r] = peer_tables[r];
  // cp.async ring: needs 16-byte aligned, unpadded rows (f == fp) of every streamed input
  const bool g_ok = p.g_bf16 ? (ldg % 8 == 0) : (ldg % 4 == 0);
  p.ring = (k_channels == 3 && f == fp && g_ok && ((reinterpret_cast<uintptr_t>(g) & 15) == 0) &&
            ((reinterpret_cast<uintptr_t>(o_lh) & 15) == 0) && ((reinterpret_cast<uintptr_t>(h_i) & 15) == 0))
               ? g_mix_bwd_ring : 0;
  p.vec_g = p.g_bf16 ? ((f % 8 == 0) && (ldg % 8 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0))
                     : ((f % 4 == 0) && (ldg % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int mode = (ln_live ? 1 : 0) | (k_channels == 4 ? 2 : 0);
  if (dtype == ACM_BF16) {
    ACM_DISPATCH_FP(fp, return launch_bwd<__nv_bfloat16, FP>(p, mode, st));
  } else {
    ACM_DISPATCH_FP(fp, return launch_bwd<float, FP>(p, mode, st));
  }
  return 0;
}

extern "C" int acm_set_mix_bwd_ring(int on) {
  acm::g_mix_bwd_ring = on ? 1 : 0;
  return 0;
}

extern "C" int acm_set_mix_bwd_occupancy(int min_blocks_per_sm) {
  if (min_blocks_per_sm != 2 && min_blocks_per_sm != 3) {
    acm::set_error("mix_bwd occupancy must be 2 or 3 CTAs per SM");
    return ACM_ERR_BAD_ARG;
  }
  acm::g_mix_bwd_minb = min_blocks_per_sm;
  return 0;
}
