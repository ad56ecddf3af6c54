// Transposed aggregation for the backward pass and the plain single-table aggregation
// used by the structure channel.
//
//  acm_spmm_t_bwd : autograd of torch.spmm(adj_low, .) and torch.spmm(adj_high, .)
//                   (ACM-Pytorch/models/layers.py:178-193):
//                     dHL = A_low^T dS_L ,  dHH = dS_H - A_low^T dS_H
//                   one gather of the [dS_L|dS_H] row per stored edge of A_low^T.
//  acm_spmm_plain : relu(mm(adj_low_unnormalized, struc_low)) (layers.py:207-209) and the
//                   transpose product of its backward.
#include "acm_common.cuh"

namespace acm {

constexpr int kTWarps = 8;
constexpr int kTUnroll = 4;

template <typename T, int FP, int HINT>
__global__ void __launch_bounds__(kTWarps * 32)
spmm_t_kernel(int64_t n_rows, int64_t row0, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
              const float* __restrict__ val, const T* __restrict__ table, const T* __restrict__ ptab,
              T* __restrict__ dh_all, const LongRows lr, const int32_t* __restrict__ row_order) {
  constexpr int LANES = FP / 8;
  constexpr int RPW = 32 / LANES;
  constexpr int TW = 2 * FP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LANES, gl = lane % LANES;
  const int64_t slot_id = ((int64_t)blockIdx.x * kTWarps + warp) * RPW + sub;
  if (slot_id >= n_rows) return;  // no cross-lane traffic in this kernel
  // degree-sorted processing order (rows of one warp walk max(degree) edges), see acm_b200.h
  const int64_t row = (RPW > 1 && row_order) ? (int64_t)__ldg(row_order + slot_id) : slot_id;
  int64_t e = __ldg(rowptr + row);
  const int64_t e1 = __ldg(rowptr + row + 1);
  const T* tab = table + gl * 8;
  float accL[8], accH[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) accL[t] = accH[t] = 0.f;
  if (lr.rows != nullptr && e1 - e > kLongRow) {
    const float* a = lr.acc + (int64_t)find_long_row(lr, row) * TW + gl * 8;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      accL[t] = a[t];
      accH[t] = a[FP + t];
    }
    e = e1;
  }
  for (; e + kTUnroll <= e1; e += kTUnroll) {
    int32_t c[kTUnroll];
    float w[kTUnroll];
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u) {
      c[u] = __ldg(col + e + u);
      w[u] = val ? __ldg(val + e + u) : 1.f;
    }
    Slice8<T> vl[kTUnroll], vh[kTUnroll];
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u) {
      const T* r = tab + (int64_t)c[u] * TW;
      gather_load<HINT>(vl[u], r);
      gather_load<HINT>(vh[u], r + FP);
    }
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u) {
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
    gather_load<HINT>(vl, r);
    gather_load<HINT>(vh, r + FP);
    float fl[8], fh[8];
    vl.to_float(fl);
    vh.to_float(fh);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      accL[t] = fmaf(w, fl[t], accL[t]);
      accH[t] = fmaf(w, fh[t], accH[t]);
    }
  }
  float self[8];
  {
    Slice8<T> s;
    s.load(tab + (row0 + row) * TW + FP);
    s.to_float(self);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) accH[t] = self[t] - accH[t];
  if (ptab) {  // variant 1: relu sits before the aggregation -> mask with the forward table
    Slice8<T> a, b;
    float pl[8], ph[8];
    a.load(ptab + row * TW + gl * 8);
    b.load(ptab + row * TW + FP + gl * 8);
    a.to_float(pl);
    b.to_float(ph);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (!(pl[t] > 0.f)) accL[t] = 0.f;
      if (!(ph[t] > 0.f)) accH[t] = 0.f;
    }
  }
  T* out = dh_all + row * (3 * FP) + gl * 8;
  Slice8<T>::store(out, accL);
  Slice8<T>::store(out + FP, accH);
}

// Variant 1 without LayerNorm ("rank-structured" backward table).  With the relu BEFORE the aggregation the
// gradient rows handed to the transposed aggregation are
//     dO_k[j,:] = c att_k[j] G[j,:] + dz_k[j] a_k^T              (k = L, H; mix_bwd, no relu mask in between)
// so   (A^T dO_k)[i,:] = sum_j w_ji (c att_k[j]) G[j,:]  +  (sum_j w_ji dz_k[j]) a_k^T :
// ONE gather of the G row (F wide) plus four scalars per stored edge serves BOTH channels -- half the
// bytes of gathering [dO_L | dO_H] (2F wide), and half the exchange under a row partition.
// Table (written by mix_bwd in table_mode 1), one allocation of table_rows * (FP*sizeof(T) + 16) bytes:
//     T g[table_rows][FP]   followed by   float4 {c att_L, c att_H, dz_L, dz_H}[table_rows]
// (the G rows keep their natural 512-byte / 1-KB alignment: a first layout that appended the scalars to each
// row -- 528-byte rows straddling DRAM pages and 128-byte lines -- ran at a third of the HBM peak).
//
// Gather = per-lane cp.async (LDGSTS) ring in shared memory, kRank1Stages neighbour rows in flight per lane, like
// the fused forward kernel: a register-staged loop left the load scheduling to ptxas, which interleaved the loads
// with the FMAs (two rows in flight, 50 % of the HBM peak measured).  Every lane copies its own 8-feature slice and
// reads back only what it copied itself; the row's 16 bytes of scalars are copied by the first lane of the row's
// lane group and broadcast by shuffle.
constexpr int kRank1Stages = 8;
template <typename T> struct Rank1Slot { static constexpr int kBytes = 8 * (int)sizeof(T) + 16; };

template <typename T, int FP>
__global__ void __launch_bounds__(kTWarps * 32)
spmm_t_rank1_kernel(int64_t n_rows, int64_t row0, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    const float* __restrict__ val, const T* __restrict__ table, const float4* __restrict__ scal,
                    const float* __restrict__ pack, const T* __restrict__ ptab, T* __restrict__ dh_all) {
  constexpr int LANES = FP / 8;
  constexpr int RPW = 32 / LANES;
  constexpr int ST = kRank1Stages;
  constexpr int SB = 8 * (int)sizeof(T);              // bytes of one 8-feature slice
  constexpr int SLOT = Rank1Slot<T>::kBytes;
  constexpr int STAGE = 32 * SLOT;                    // per warp and stage
  extern __shared__ __align__(16) uint8_t rank1_ring[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LANES, gl = lane % LANES;
  const int64_t row = ((int64_t)blockIdx.x * kTWarps + warp) * RPW + sub;
  if (row >= n_rows) return;                          // whole lane groups leave together
  const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
  const int64_t e = __ldg(rowptr + row);
  const int n_e = (int)(__ldg(rowptr + row + 1) - e);
  float accL[8], accH[8], sL = 0.f, sH = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t) accL[t] = accH[t] = 0.f;
  uint8_t* ring = rank1_ring + warp * (ST * STAGE) + lane * SLOT;
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
#pragma unroll
  for (int st = 0; st < ST; ++st) {
    if (st < n_e) {
      const int64_t c = __ldg(col + e + st);
      cp_async_slice<T>(ring_u32 + st * STAGE, table + c * FP + gl * 8);
      if (gl == 0) cp_async16(ring_u32 + st * STAGE + SB, scal + c);
    }
    cp_async_commit();
  }
  for (int i = 0; i < n_e; ++i) {
    cp_async_wait<ST - 1>();                          // the oldest group (edge i) has landed
    const float w = val ? __ldg(val + e + i) : 1.f;
    const int slot = i & (ST - 1);
    Slice8<T> v;
    v.load_plain(reinterpret_cast<const T*>(ring + slot * STAGE));
    float4 sc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gl == 0) sc = *reinterpret_cast<const float4*>(ring + slot * STAGE + SB);
    sc.x = __shfl_sync(gmask, sc.x, sub * LANES);
    sc.y = __shfl_sync(gmask, sc.y, sub * LANES);
    sc.z = __shfl_sync(gmask, sc.z, sub * LANES);
    sc.w = __shfl_sync(gmask, sc.w, sub * LANES);
    float f[8];
    v.to_float(f);
    sL = fmaf(w, sc.z, sL);
    sH = fmaf(w, sc.w, sH);
    axpy8(accL, w * sc.x, f);
    axpy8(accH, w * sc.y, f);
    if (i + ST < n_e) {
      const int64_t c = __ldg(col + e + i + ST);
      cp_async_slice<T>(ring_u32 + slot * STAGE, table + c * FP + gl * 8);
      if (gl == 0) cp_async16(ring_u32 + slot * STAGE + SB, scal + c);
    }
    cp_async_commit();
  }
  // own row: dO_H[i,:] = c att_H[i] G[i,:] + dz_H[i] a_H
  float g[8], aL[8], aH[8];
  {
    Slice8<T> s;
    s.load(table + (row0 + row) * FP + gl * 8);
    s.to_float(g);
    const float4 sc = __ldg(scal + row0 + row);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      aL[t] = __ldg(pack + pack_off_a(FP, 0) + gl * 8 + t);
      aH[t] = __ldg(pack + pack_off_a(FP, 1) + gl * 8 + t);
      accL[t] = fmaf(sL, aL[t], accL[t]);
      accH[t] = fmaf(sc.y, g[t], sc.w * aH[t]) - fmaf(sH, aH[t], accH[t]);
    }
  }
  if (ptab) {  // relu before the aggregation -> mask with the (relu'd) forward table
    constexpr int TW = 2 * FP;
    Slice8<T> a, b;
    float pl[8], ph[8];
    a.load(ptab + row * TW + gl * 8);
    b.load(ptab + row * TW + FP + gl * 8);
    a.to_float(pl);
    b.to_float(ph);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (!(pl[t] > 0.f)) accL[t] = 0.f;
      if (!(ph[t] > 0.f)) accH[t] = 0.f;
    }
  }
  T* out = dh_all + row * (3 * FP) + gl * 8;
  Slice8<T>::store(out, accL);
  Slice8<T>::store(out + FP, accH);
}

// Transposed aggregation at width 256 in bf16 with the [dS_L|dS_H] rows (1 KB) staged by the TMA engine
// (tile::gather4, 4 KB per request, 2-stage ring per warp, 3 CTAs per SM) -- the configuration that beat the cp.async
// ring in the fused forward kernel (spmm_fwd.cu gather mode 3).  Same accumulation order as spmm_t_kernel: bit-identical.
constexpr int kTTmaStages = 2;
__global__ void __launch_bounds__(kTWarps * 32)
spmm_t_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_rows, int64_t row0, const int64_t* __restrict__ rowptr,
                  const int32_t* __restrict__ col, const float* __restrict__ val, const __nv_bfloat16* __restrict__ table,
                  const __nv_bfloat16* __restrict__ ptab, __nv_bfloat16* __restrict__ dh_all, const LongRows lr) {
  using T = __nv_bfloat16;
  constexpr int FP = 256, TW = 2 * FP;
  constexpr int ST = kTTmaStages;
  constexpr uint32_t ROWB = TW * sizeof(T);           // 1 KB
  constexpr uint32_t STAGE = 4 * ROWB;
  extern __shared__ __align__(128) uint8_t t_smem[];
  uint8_t* base = t_smem + ((128u - ((uint32_t)__cvta_generic_to_shared(t_smem) & 127u)) & 127u);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* ring = base + warp * (ST * STAGE);
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(base + kTWarps * (ST * STAGE)) + warp * (ST * 8);
  if (lane < ST) mbar_init(bar_u32 + lane * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int64_t row = (int64_t)blockIdx.x * kTWarps + warp;
  if (row >= n_rows) return;
  int64_t e = __ldg(rowptr + row);
  const int64_t e1 = __ldg(rowptr + row + 1);
  float accL[8], accH[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) accL[t] = accH[t] = 0.f;
  if (lr.rows != nullptr && e1 - e > kLongRow) {
    const float* a = lr.acc + (int64_t)find_long_row(lr, row) * TW + lane * 8;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      accL[t] = a[t];
      accH[t] = a[FP + t];
    }
    e = e1;
  }
  if (e < e1) {
    const int n_e = (int)(e1 - e);
    const int n_grp = (n_e + 3) >> 2;
    float wcur = (lane < n_e) ? (val ? __ldg(val + e + lane) : 1.f) : 0.f;
    int32_t cnext = (lane < n_e) ? __ldg(col + e + lane) : 0;
    auto issue = [&](int grp, int stage) {
      const int b = (grp * 4) & 31;
      const int r0 = __shfl_sync(0xffffffffu, cnext, b), r1 = __shfl_sync(0xffffffffu, cnext, b + 1);
      const int r2 = __shfl_sync(0xffffffffu, cnext, b + 2), r3 = __shfl_sync(0xffffffffu, cnext, b + 3);
      if (lane == 0) {
        mbar_expect_tx(bar_u32 + stage * 8, STAGE);
        tma_gather4(ring_u32 + stage * STAGE, &tmap, r0, r1, r2, r3, bar_u32 + stage * 8);
      }
    };
#pragma unroll
    for (int st = 0; st < ST; ++st)
      if (st < n_grp) issue(st, st);
    for (int g = 0; g < n_grp; ++g) {
      const int stage = g & (ST - 1);
      if (((g * 4) & 31) == 0 && g) wcur = (g * 4 + lane < n_e) ? (val ? __ldg(val + e + g * 4 + lane) : 1.f) : 0.f;
      mbar_wait(bar_u32 + stage * 8, (g / ST) & 1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = __shfl_sync(0xffffffffu, wcur, (g * 4 + u) & 31);
        Slice8<T> vl, vh;
        vl.load_plain(reinterpret_cast<const T*>(ring + stage * STAGE + u * ROWB + lane * 16));
        vh.load_plain(reinterpret_cast<const T*>(ring + stage * STAGE + u * ROWB + FP * sizeof(T) + lane * 16));
        float fl[8], fh[8];
        vl.to_float(fl);
        vh.to_float(fh);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          accL[t] = fmaf(w, fl[t], accL[t]);
          accH[t] = fmaf(w, fh[t], accH[t]);
        }
      }
      __syncwarp();
      const int j = g + ST;
      if (j < n_grp) {
        if (((j * 4) & 31) == 0) cnext = (j * 4 + lane < n_e) ? __ldg(col + e + j * 4 + lane) : 0;
        issue(j, stage);
      }
    }
  }
  float self[8];
  {
    Slice8<T> s;
    s.load(table + (row0 + row) * TW + FP + lane * 8);
    s.to_float(self);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) accH[t] = self[t] - accH[t];
  if (ptab) {  // variant 1: relu sits before the aggregation -> mask with the forward table
    Slice8<T> a, b;
    float pl[8], ph[8];
    a.load(ptab + row * TW + lane * 8);
    b.load(ptab + row * TW + FP + lane * 8);
    a.to_float(pl);
    b.to_float(ph);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (!(pl[t] > 0.f)) accL[t] = 0.f;
      if (!(ph[t] > 0.f)) accH[t] = 0.f;
    }
  }
  T* out = dh_all + row * (3 * FP) + lane * 8;
  Slice8<T>::store(out, accL);
  Slice8<T>::store(out + FP, accH);
}

template <typename T, typename TO, int FP>
__global__ void __launch_bounds__(kTWarps * 32)
spmm_plain_kernel(int64_t n_rows, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                  const float* __restrict__ val, const T* __restrict__ table, TO* __restrict__ out,
                  int64_t ld_out, int f_out, int relu) {
  constexpr int LANES = FP / 8;
  constexpr int RPW = 32 / LANES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LANES, gl = lane % LANES;
  const int64_t row = ((int64_t)blockIdx.x * kTWarps + warp) * RPW + sub;
  if (row >= n_rows) return;
  int64_t e = __ldg(rowptr + row);
  const int64_t e1 = __ldg(rowptr + row + 1);
  const T* tab = table + gl * 8;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  for (; e + kTUnroll <= e1; e += kTUnroll) {
    int32_t c[kTUnroll];
    float w[kTUnroll];
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u) {
      c[u] = __ldg(col + e + u);
      w[u] = val ? __ldg(val + e + u) : 1.f;
    }
    Slice8<T> v[kTUnroll];
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u) v[u].load(tab + (int64_t)c[u] * FP);
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u) {
      float f[8];
      v[u].to_float(f);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(w[u], f[t], acc[t]);
    }
  }
  for (; e < e1; ++e) {
    const int32_t c = __ldg(col + e);
    const float w = val ? __ldg(val + e) : 1.f;
    Slice8<T> v;
    v.load(tab + (int64_t)c * FP);
    float f[8];
    v.to_float(f);
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = fmaf(w, f[t], acc[t]);
  }
  TO* o = out + row * ld_out + gl * 8;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float v = relu ? fmaxf(acc[t], 0.f) : acc[t];
    if (gl * 8 + t < f_out) {
      if constexpr (sizeof(TO) == 2) o[t] = __float2bfloat16_rn(v);
      else o[t] = v;
    }
  }
}

// Aggregate-first order (A(XW) = (AX)W): Z = A.X and D = X - Z for the
// own rows, one gather of the (narrower) INPUT row per stored edge.
// A cp.async shared-memory ring like the fused forward kernel's was tried here and lost: 21.8 ms
// (87 % of the HBM peak) against 20.9 ms (91 %) for this register-staged loop in the same run --
// with 40 registers this kernel already keeps 48 warps x 8 rows x 512 B in flight per SM.
template <typename T, int FP>
__global__ void __launch_bounds__(kTWarps * 32)
spmm_agg_first_kernel(int64_t n_rows, int64_t row0, const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                      const float* __restrict__ val, const T* __restrict__ table, T* __restrict__ z_out,
                      T* __restrict__ d_out, const LongRows lr) {
  constexpr int LANES = FP / 8;
  constexpr int RPW = 32 / LANES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LANES, gl = lane % LANES;
  const int64_t row = ((int64_t)blockIdx.x * kTWarps + warp) * RPW + sub;
  if (row >= n_rows) return;
  int64_t e = __ldg(rowptr + row);
  const int64_t e1 = __ldg(rowptr + row + 1);
  const T* tab = table + gl * 8;
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  constexpr int U = 8;  // narrower rows than the [HL|HH] table: keep the same bytes in flight
  if (lr.rows != nullptr && e1 - e > kLongRow) {
    const float* a = lr.acc + (int64_t)find_long_row(lr, row) * FP + gl * 8;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = a[t];
    e = e1;
  }
  for (; e + U <= e1; e += U) {
    int32_t c[U];
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      c[u] = __ldg(col + e + u);
      w[u] = val ? __ldg(val + e + u) : 1.f;
    }
    Slice8<T> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u].load(tab + (int64_t)c[u] * FP);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float f[8];
      v[u].to_float(f);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaf(w[u], f[t], acc[t]);
    }
  }
  for (; e < e1; ++e) {
    const int32_t c = __ldg(col + e);
    const float w = val ? __ldg(val + e) : 1.f;
    Slice8<T> v;
    v.load(tab + (int64_t)c * FP);
    float f[8];
    v.to_float(f);
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = fmaf(w, f[t], acc[t]);
  }
  float self[8], d[8];
  {
    Slice8<T> s;
    s.load(tab + (row0 + row) * FP);
    s.to_float(self);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) d[t] = self[t] - acc[t];
  Slice8<T>::store(z_out + row * FP + gl * 8, acc);
  Slice8<T>::store(d_out + row * FP + gl * 8, d);
}

// The same aggregation with the neighbour rows staged by the TMA engine (tile::gather4, four 512-byte rows per request
// into a 4-stage ring of 2-KB stages per warp; see spmm_fwd.cu gather mode 3, where the technique beat the cp.async ring
// 34.6 vs 37.4 ms).  Input width 256 in bf16: one row per warp, a lane owns 16 bytes of every row.
// MEASURED SLOWER here: 24.4 vs 20.6 ms in the same run (profiles/r2_tma_gather_ab.txt).  With 512-byte rows a request
// carries 2 KB instead of 4 KB, and at 52.5 M requests per launch the engine's request rate (~1 per 130 cycles per SM),
// not the bytes, sets the pace.  Opt-in only (gather mode 4); the register-staged loop stays the default of this kernel.
constexpr int kAggTmaStages = 4;
__global__ void __launch_bounds__(kTWarps * 32)
spmm_agg_first_tma_kernel(const __grid_constant__ CUtensorMap tmap, int64_t n_rows, int64_t row0,
                          const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                          const __nv_bfloat16* __restrict__ table, __nv_bfloat16* __restrict__ z_out,
                          __nv_bfloat16* __restrict__ d_out, const LongRows lr) {
  using T = __nv_bfloat16;
  constexpr int FP = 256;
  constexpr int ST = kAggTmaStages;
  constexpr uint32_t ROWB = FP * sizeof(T);           // 512
  constexpr uint32_t STAGE = 4 * ROWB;
  extern __shared__ __align__(128) uint8_t agg_smem[];
  uint8_t* base = agg_smem + ((128u - ((uint32_t)__cvta_generic_to_shared(agg_smem) & 127u)) & 127u);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* ring = base + warp * (ST * STAGE);
  const uint32_t ring_u32 = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t bar_u32 = (uint32_t)__cvta_generic_to_shared(base + kTWarps * (ST * STAGE)) + warp * (ST * 8);
  if (lane < ST) mbar_init(bar_u32 + lane * 8, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const int64_t row = (int64_t)blockIdx.x * kTWarps + warp;
  if (row >= n_rows) return;                            // whole warps leave together
  int64_t e = __ldg(rowptr + row);
  const int64_t e1 = __ldg(rowptr + row + 1);
  float acc[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) acc[t] = 0.f;
  if (lr.rows != nullptr && e1 - e > kLongRow) {
    const float* a = lr.acc + (int64_t)find_long_row(lr, row) * FP + lane * 8;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = a[t];
    e = e1;
  }
  if (e < e1) {
    const int n_e = (int)(e1 - e);
    const int n_grp = (n_e + 3) >> 2;
    // indices / weights in registers, 32 edges per coalesced load, broadcast by shuffle; padding edges gather row 0 with weight 0
    float wcur = (lane < n_e) ? (val ? __ldg(val + e + lane) : 1.f) : 0.f;
    int32_t cnext = (lane < n_e) ? __ldg(col + e + lane) : 0;
    auto issue = [&](int grp, int stage) {
      const int b = (grp * 4) & 31;
      const int r0 = __shfl_sync(0xffffffffu, cnext, b), r1 = __shfl_sync(0xffffffffu, cnext, b + 1);
      const int r2 = __shfl_sync(0xffffffffu, cnext, b + 2), r3 = __shfl_sync(0xffffffffu, cnext, b + 3);
      if (lane == 0) {
        mbar_expect_tx(bar_u32 + stage * 8, STAGE);
        tma_gather4(ring_u32 + stage * STAGE, &tmap, r0, r1, r2, r3, bar_u32 + stage * 8);
      }
    };
#pragma unroll
    for (int st = 0; st < ST; ++st)
      if (st < n_grp) issue(st, st);                    // ST * 4 <= 32: inside the first index chunk
    for (int g = 0; g < n_grp; ++g) {
      const int stage = g & (ST - 1);
      if (((g * 4) & 31) == 0 && g) wcur = (g * 4 + lane < n_e) ? (val ? __ldg(val + e + g * 4 + lane) : 1.f) : 0.f;
      mbar_wait(bar_u32 + stage * 8, (g / ST) & 1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = __shfl_sync(0xffffffffu, wcur, (g * 4 + u) & 31);
        Slice8<T> v;
        v.load_plain(reinterpret_cast<const T*>(ring + stage * STAGE + u * ROWB + lane * 16));
        float f[8];
        v.to_float(f);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] = fmaf(w, f[t], acc[t]);
      }
      __syncwarp();                                     // every lane has read the stage
      const int j = g + ST;
      if (j < n_grp) {
        if (((j * 4) & 31) == 0) cnext = (j * 4 + lane < n_e) ? __ldg(col + e + j * 4 + lane) : 0;
        issue(j, stage);
      }
    }
  }
  float self[8], d[8];
  {
    Slice8<T> s;
    s.load(table + (row0 + row) * FP + lane * 8);
    s.to_float(self);
  }
#pragma unroll
  for (int t = 0; t < 8; ++t) d[t] = self[t] - acc[t];
  Slice8<T>::store(z_out + row * FP + lane * 8, acc);
  Slice8<T>::store(d_out + row * FP + lane * 8, d);
}

// Segment-parallel aggregation of the long rows: one lane group per segment of <= kLongRow
// edges, partial sums added atomically into acc[long_index, :] (fp32, zeroed by the caller).
// HALVES = 2: table rows are [L | H] pairs of FP features (row width 2*FP); 1: single FP-wide rows.
template <typename T, int FP, int HALVES>
__global__ void __launch_bounds__(kTWarps * 32)
spmm_long_rows_kernel(int64_t n_seg, const int32_t* __restrict__ seg_long, const int64_t* __restrict__ seg_e0,
                      const int64_t* __restrict__ seg_e1, const int32_t* __restrict__ col,
                      const float* __restrict__ val, const T* __restrict__ table, float* __restrict__ acc_out) {
  constexpr int LANES = FP / 8;
  constexpr int RPW = 32 / LANES;
  constexpr int TW = HALVES * FP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane / LANES, gl = lane % LANES;
  const int64_t seg = ((int64_t)blockIdx.x * kTWarps + warp) * RPW + sub;
  if (seg >= n_seg) return;
  int64_t e = seg_e0[seg];
  const int64_t e1 = seg_e1[seg];
  const T* tab = table + gl * 8;
  float acc[HALVES][8];
#pragma unroll
  for (int h = 0; h < HALVES; ++h)
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[h][t] = 0.f;
  for (; e + kTUnroll <= e1; e += kTUnroll) {
    int32_t c[kTUnroll];
    float w[kTUnroll];
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u) {
      c[u] = __ldg(col + e + u);
      w[u] = val ? __ldg(val + e + u) : 1.f;
    }
    Slice8<T> v[kTUnroll][HALVES];
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u)
#pragma unroll
      for (int h = 0; h < HALVES; ++h) v[u][h].load(tab + (int64_t)c[u] * TW + h * FP);
#pragma unroll
    for (int u = 0; u < kTUnroll; ++u)
#pragma unroll
      for (int h = 0; h < HALVES; ++h) {
        float f[8];
        v[u][h].to_float(f);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[h][t] = fmaf(w[u], f[t], acc[h][t]);
      }
  }
  for (; e < e1; ++e) {
    const int32_t c = __ldg(col + e);
    const float w = val ? __ldg(val + e) : 1.f;
#pragma unroll
    for (int h = 0; h < HALVES; ++h) {
      Slice8<T> v;
      v.load(tab + (int64_t)c * TW + h * FP);
      float f[8];
      v.to_float(f);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[h][t] = fmaf(w, f[t], acc[h][t]);
    }
  }
  float* o = acc_out + (int64_t)seg_long[seg] * TW + gl * 8;
#pragma unroll
  for (int h = 0; h < HALVES; ++h)
#pragma unroll
    for (int t = 0; t < 8; ++t) atomicAdd(o + h * FP + t, acc[h][t]);
}

}  // namespace acm

extern "C" int acm_spmm_long_rows(int dtype, int fp, int halves, int64_t n_seg, const int32_t* seg_long,
                                  const int64_t* seg_e0, const int64_t* seg_e1, const int32_t* col, const float* val,
                                  const void* table, float* acc_out, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "spmm_long_rows: bad dtype %d", dtype);
  ACM_CHECK_ARG(halves == 1 || halves == 2, "spmm_long_rows: halves must be 1 or 2");
  ACM_CHECK_ARG(seg_long && seg_e0 && seg_e1 && col && table && acc_out, "spmm_long_rows: null pointer");
  if (n_seg == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ACM_L_LAUNCH(TT, HH)                                                                    \
  ACM_DISPATCH_FP(fp, {                                                                         \
    constexpr int RPB = (32 / (FP / 8)) * kTWarps;                                              \
    const int64_t blocks = (n_seg + RPB - 1) / RPB;                                             \
    ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_long_rows: too many segments");                   \
    spmm_long_rows_kernel<TT, FP, HH><<<(unsigned)blocks, kTWarps * 32, 0, st>>>(               \
        n_seg, seg_long, seg_e0, seg_e1, col, val, (const TT*)table, acc_out);                  \
  })
  if (dtype == ACM_BF16) {
    if (halves == 2) { ACM_L_LAUNCH(__nv_bfloat16, 2); } else { ACM_L_LAUNCH(__nv_bfloat16, 1); }
  } else {
    if (halves == 2) { ACM_L_LAUNCH(float, 2); } else { ACM_L_LAUNCH(float, 1); }
  }
#undef ACM_L_LAUNCH
  ACM_LAUNCH_CHECK("spmm_long_rows");
  return 0;
}

extern "C" int acm_spmm_agg_first(int dtype, int fp, int64_t n_rows, int64_t row0,
                                  const int64_t* rowptr, const int32_t* col, const float* val,
                                  const void* table, void* z_out, void* d_out,
                                  const int32_t* long_rows, int n_long, const float* long_acc, void* stream) {
  using namespace acm;
  LongRows lr{n_long > 0 ? long_rows : nullptr, long_acc, n_long};
  ACM_CHECK_ARG(n_long == 0 || (long_rows && long_acc), "spmm_agg_first: long rows need long_rows and long_acc");
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "spmm_agg_first: bad dtype %d", dtype);
  ACM_CHECK_ARG(rowptr && col && table && z_out && d_out, "spmm_agg_first: null pointer");
  if (n_rows == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == ACM_BF16 && fp == 256 && g_gather_mode == 4) {
    // neighbour rows staged by the TMA engine (tile::gather4); row extent: see spmm_fwd.cu
    CUtensorMap tm;
    if (int rc = tma_encode_2d_u32(&tm, table, 128, 0x7fffffffull, 512, 128, 1, "input table")) return rc;
    const int64_t blocks = (n_rows + kTWarps - 1) / kTWarps;
    ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_agg_first: too many rows");
    const size_t smem = (size_t)kTWarps * kAggTmaStages * (4 * 512 + 8) + 128;
    cudaError_t e_ = cudaFuncSetAttribute(spmm_agg_first_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e_ != cudaSuccess) { set_error("spmm_agg_first: smem attribute: %s", cudaGetErrorString(e_)); return (int)e_; }
    spmm_agg_first_tma_kernel<<<(unsigned)blocks, kTWarps * 32, smem, st>>>(
        tm, n_rows, row0, rowptr, col, val, (const __nv_bfloat16*)table, (__nv_bfloat16*)z_out, (__nv_bfloat16*)d_out, lr);
    ACM_LAUNCH_CHECK("spmm_agg_first (TMA gather)");
    return 0;
  }
#define ACM_A_LAUNCH(TT)                                                                          \
  ACM_DISPATCH_FP(fp, {                                                                           \
    constexpr int RPB = (32 / (FP / 8)) * kTWarps;                                                \
    const int64_t blocks = (n_rows + RPB - 1) / RPB;                                              \
    ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_agg_first: too many rows");                         \
    spmm_agg_first_kernel<TT, FP><<<(unsigned)blocks, kTWarps * 32, 0, st>>>(                     \
        n_rows, row0, rowptr, col, val, (const TT*)table, (TT*)z_out, (TT*)d_out, lr);            \
  })
  if (dtype == ACM_BF16) { ACM_A_LAUNCH(__nv_bfloat16); } else { ACM_A_LAUNCH(float); }
#undef ACM_A_LAUNCH
  ACM_LAUNCH_CHECK("spmm_agg_first");
  return 0;
}

extern "C" int acm_spmm_t_bwd(int dtype, int fp, int64_t n_rows, int64_t row0,
                              const int64_t* rowptr_t, const int32_t* col_t, const float* val_t,
                              const void* t_table, const void* p_table, void* dh_all,
                              const int32_t* long_rows, int n_long, const float* long_acc,
                              const int32_t* row_order, void* stream) {
  using namespace acm;
  LongRows lr{n_long > 0 ? long_rows : nullptr, long_acc, n_long};
  ACM_CHECK_ARG(n_long == 0 || (long_rows && long_acc), "spmm_t_bwd: long rows need long_rows and long_acc");
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "spmm_t_bwd: bad dtype %d", dtype);
  ACM_CHECK_ARG(rowptr_t && col_t && t_table && dh_all, "spmm_t_bwd: null pointer");
  if (n_rows == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == ACM_BF16 && fp == 256 && g_gather_mode >= 3) {
    CUtensorMap tm;     // [dS_L|dS_H] table as uint32 [*, 256], box {256 x 1}; row extent: see spmm_fwd.cu
    if (int rc = tma_encode_2d_u32(&tm, t_table, 256, 0x7fffffffull, 1024, 256, 1, "backward table")) return rc;
    const int64_t blocks = (n_rows + kTWarps - 1) / kTWarps;
    ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_t_bwd: too many rows");
    const size_t smem = (size_t)kTWarps * kTTmaStages * (4 * 1024 + 8) + 128;
    cudaError_t e_ = cudaFuncSetAttribute(spmm_t_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e_ != cudaSuccess) { set_error("spmm_t_bwd: smem attribute: %s", cudaGetErrorString(e_)); return (int)e_; }
    spmm_t_tma_kernel<<<(unsigned)blocks, kTWarps * 32, smem, st>>>(
        tm, n_rows, row0, rowptr_t, col_t, val_t, (const __nv_bfloat16*)t_table, (const __nv_bfloat16*)p_table,
        (__nv_bfloat16*)dh_all, lr);
    ACM_LAUNCH_CHECK("spmm_t_bwd (TMA gather)");
    return 0;
  }
#define ACM_T_LAUNCH(TT)                                                                           \
  ACM_DISPATCH_FP(fp, {                                                                            \
    constexpr int RPB = (32 / (FP / 8)) * kTWarps;                                                 \
    const int64_t blocks = (n_rows + RPB - 1) / RPB;                                               \
    ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_t_bwd: too many rows");                              \
    if (FP <= 32 && g_narrow_row_hint) {                                                           \
      constexpr int H = FP <= 32 ? 1 : 0;                                                          \
      spmm_t_kernel<TT, FP, H><<<(unsigned)blocks, kTWarps * 32, 0, st>>>(                         \
          n_rows, row0, rowptr_t, col_t, val_t, (const TT*)t_table, (const TT*)p_table, (TT*)dh_all, lr, row_order); \
    } else {                                                                                       \
      spmm_t_kernel<TT, FP, 0><<<(unsigned)blocks, kTWarps * 32, 0, st>>>(                         \
          n_rows, row0, rowptr_t, col_t, val_t, (const TT*)t_table, (const TT*)p_table, (TT*)dh_all, lr, row_order); \
    }                                                                                              \
  })
  if (dtype == ACM_BF16) { ACM_T_LAUNCH(__nv_bfloat16); } else { ACM_T_LAUNCH(float); }
#undef ACM_T_LAUNCH
  ACM_LAUNCH_CHECK("spmm_t_bwd");
  return 0;
}

extern "C" int acm_spmm_t_bwd_rank1(int dtype, int fp, int64_t n_rows, int64_t row0,
                                    const int64_t* rowptr_t, const int32_t* col_t, const float* val_t,
                                    const void* g_table, int64_t table_rows, const float* pack, const void* p_table,
                                    void* dh_all, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "spmm_t_bwd_rank1: bad dtype %d", dtype);
  ACM_CHECK_ARG(rowptr_t && col_t && g_table && pack && dh_all, "spmm_t_bwd_rank1: null pointer");
  ACM_CHECK_ARG(fp >= 64, "spmm_t_bwd_rank1: built for padded widths >= 64 (got %d)", fp);
  ACM_CHECK_ARG(table_rows >= row0 + n_rows, "spmm_t_bwd_rank1: table_rows must cover the own rows");
  const size_t el = dtype == ACM_BF16 ? 2 : 4;
  const float4* scal = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(g_table) + (size_t)table_rows * fp * el);
  if (n_rows == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ACM_R_LAUNCH(TT)                                                                           \
  switch (fp) {                                                                                    \
    case 64: { constexpr int FP = 64; ACM_R_BODY(TT) } break;                                      \
    case 128: { constexpr int FP = 128; ACM_R_BODY(TT) } break;                                    \
    case 256: { constexpr int FP = 256; ACM_R_BODY(TT) } break;                                    \
    default: set_error("spmm_t_bwd_rank1: padded width %d not in {64,128,256}", fp); return ACM_ERR_UNSUPPORTED; \
  }
#define ACM_R_BODY(TT)                                                                             \
    constexpr int RPB = (32 / (FP / 8)) * kTWarps;                                                 \
    const int64_t blocks = (n_rows + RPB - 1) / RPB;                                               \
    ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_t_bwd_rank1: too many rows");                        \
    constexpr size_t smem = (size_t)kTWarps * kRank1Stages * 32 * Rank1Slot<TT>::kBytes;           \
    if (smem > 48 * 1024) {                                                                        \
      cudaError_t e_ = cudaFuncSetAttribute(spmm_t_rank1_kernel<TT, FP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e_ != cudaSuccess) { set_error("spmm_t_bwd_rank1: smem attribute: %s", cudaGetErrorString(e_)); return (int)e_; } \
    }                                                                                              \
    spmm_t_rank1_kernel<TT, FP><<<(unsigned)blocks, kTWarps * 32, smem, st>>>(                     \
        n_rows, row0, rowptr_t, col_t, val_t, (const TT*)g_table, scal, pack, (const TT*)p_table, (TT*)dh_all);
  if (dtype == ACM_BF16) { ACM_R_LAUNCH(__nv_bfloat16) } else { ACM_R_LAUNCH(float) }
#undef ACM_R_BODY
#undef ACM_R_LAUNCH
  ACM_LAUNCH_CHECK("spmm_t_bwd_rank1");
  return 0;
}

extern "C" int acm_spmm_plain(int dtype, int out_dtype, int fp, int64_t n_rows,
                              const int64_t* rowptr, const int32_t* col, const float* val,
                              const void* table, void* out, int64_t ld_out, int f_out, int relu, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "spmm_plain: bad dtype %d", dtype);
  ACM_CHECK_ARG(out_dtype == ACM_F32 || out_dtype == ACM_BF16, "spmm_plain: bad out dtype %d", out_dtype);
  ACM_CHECK_ARG(rowptr && col && table && out, "spmm_plain: null pointer");
  ACM_CHECK_ARG(f_out >= 1 && f_out <= fp, "spmm_plain: need 1 <= f_out <= fp");
  if (n_rows == 0) return 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ACM_P_LAUNCH(TT, TO)                                                                  \
  ACM_DISPATCH_FP(fp, {                                                                       \
    constexpr int RPB = (32 / (FP / 8)) * kTWarps;                                            \
    const int64_t blocks = (n_rows + RPB - 1) / RPB;                                          \
    ACM_CHECK_ARG(blocks < (1ll << 31), "spmm_plain: too many rows");                         \
    spmm_plain_kernel<TT, TO, FP><<<(unsigned)blocks, kTWarps * 32, 0, st>>>(                 \
        n_rows, rowptr, col, val, (const TT*)table, (TO*)out, ld_out, f_out, relu);           \
  })
  if (dtype == ACM_BF16) {
    if (out_dtype == ACM_BF16) { ACM_P_LAUNCH(__nv_bfloat16, __nv_bfloat16); } else { ACM_P_LAUNCH(__nv_bfloat16, float); }
  } else {
    if (out_dtype == ACM_BF16) { ACM_P_LAUNCH(float, __nv_bfloat16); } else { ACM_P_LAUNCH(float, float); }
  }
#undef ACM_P_LAUNCH
  ACM_LAUNCH_CHECK("spmm_plain");
  return 0;
}
