// tcgen05 (5th-generation tensor core) bf16 GEMMs of the ACM layer, hand-written for sm_100a:
// TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory -> tcgen05.mma kind::f16 with the
// fp32 accumulator in TMEM -> tcgen05.ld epilogue.  Warp-specialised: warp 0 = TMA producer,
// warp 1 = MMA issuer (one elected lane) + TMEM owner, warps 2..5 = epilogue (one per
// 32-lane TMEM quadrant).  mbarrier full/empty ring between producer and MMA,
// tcgen05.commit to release stages and to publish the accumulator.
//
//   tn kernel :  C[M,N] = A[M,K] . B[N,K]^T      both operands K-major (row-major, K contiguous)
//       forward   [HL|HH|HI] = X . Wcat        A = X [n,fin],   B = Wcat^T [3fp, fin]
//                 (ACM-Pytorch/models/layers.py:163-165,179-194, three torch.mm fused)
//       backward  dX = dH . Wcat^T             A = dH [n,3fp],  B = Wcat [fin, 3fp]
//   nt kernel :  C[M,N] += A[K,M]^T . B[K,N]     both operands MN-major, split-K over CTAs
//       backward  dWcat = X^T . dH             A = X [n,fin],   B = dH [n,3fp]
#include <cuda.h>

#include "acm_common.cuh"

namespace acm {
namespace tc {

constexpr int BM = 128;       // UMMA_M (cta_group::1)
constexpr int BK = 64;        // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (sm_100 UMMA): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), majors (15,16), N>>3 (17..22), M>>4 (24..28)
__host__ __device__ __forceinline__ uint32_t make_idesc(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// 32 bytes of this thread's own output row in one 256-bit store (STG.E.256, sm_100): a whole sector per thread
__device__ __forceinline__ void st_row32(void* dst, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct TnParams {
  int64_t m, n, k;       // problem
  int bn;                // UMMA_N / B box rows (multiple of 16, <= 256)
  int n_tiles, stages, tmem_cols, acc_cols;
  int64_t total_tiles;
  // MODE 0 (forward): bf16 outputs split at ncols0
  __nv_bfloat16* c0; int64_t ldc0; int ncols0;
  __nv_bfloat16* c1; int64_t ldc1;
  int relu_cols;
  // MODE 1 (dX): fp32 output
  float* cf; int64_t ldcf; int vec_ok;
  // every output row segment of 16 bf16 / 8 fp32 columns is 32-byte aligned: the epilogue writes it with ONE
  // 256-bit store per thread straight from the TMEM registers (thread = row), no shared-memory transposition
  int direct;
  // MODE 0: optional push of the c0 part ([HL|HH] rows) into every rank's table (peer memory)
  PeerTables peers;
  // MODE 0: optional fp32 bias per output column, added before the relu (nn.Linear of the MLP helper)
  const float* bias;
};

constexpr int kStgLd = 36;                           // staging row stride in words (32 + 4 pad)
// EPI = epilogue warps: 8 (two per 32-lane TMEM quadrant) or, for store-bound products with a short K loop where
// the epilogue is the critical path of the tile loop, 16 (four per quadrant)
template <int EPI> struct TnCfg { static constexpr int kThreads = 64 + 32 * EPI; };   // + producer warp + MMA warp

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent kernel: one CTA per SM walks tiles (n fastest, so neighbouring CTAs share the A
// tile in L2); the TMA/MMA ring runs continuously across tiles and the accumulator is double
// buffered in TMEM, so the epilogue of tile t overlaps the main loop of tile t+1.
template <int MODE, int EPI>
__global__ void __launch_bounds__(TnCfg<EPI>::kThreads, 1)
tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
  constexpr int kTnEpiWarps = EPI;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = BM * BK * 2;
  const uint32_t b_bytes = (uint32_t)p.bn * BK * 2;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);
  const uint32_t bar_base = base + p.stages * stage_bytes;        // full[s] | empty[s]
  const uint32_t tfull = bar_base + 16 * p.stages;                // tmem_full[2]
  const uint32_t tempty = tfull + 16;                             // tmem_empty[2]
  const uint32_t tmem_slot = tempty + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb_total = (int)((p.k + BK - 1) / BK);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (p.stages + s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull + 8 * a, 1);
      mbar_init(tempty + 8 * a, kTnEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      int s = 0;
      uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n0 = (int)(tile % p.n_tiles) * p.bn;
        const int m0 = (int)((tile / p.n_tiles) * BM);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_base + 8 * (p.stages + s), ph ^ 1);
          const uint32_t full = bar_base + 8 * s;
          mbar_expect_tx(full, a_bytes + b_bytes);
          tma_load_2d(base + s * stage_bytes, &tmA, kb * BK, m0, full);
          tma_load_2d(base + s * stage_bytes + a_bytes, &tmB, kb * BK, n0, full);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BM, p.bn, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int t = 0;
      for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
        const int a = t & 1;
        mbar_wait(tempty + 8 * a, ((t >> 1) & 1) ^ 1);   // epilogue drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(a * p.acc_cols);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_base + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * stage_bytes, sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major, 128B swizzle: rows are 128 B apart, 8-row groups 1024 B apart (SBO);
            // stepping K by 16 bf16 = +32 B inside the swizzle row
            const uint64_t ad = make_desc(sa + k * UMMA_K * 2, 16, 1024);
            const uint64_t bd = make_desc(sb + k * UMMA_K * 2, 16, 1024);
            umma_bf16(d_tmem, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(bar_base + 8 * (p.stages + s));  // frees the smem stage when the MMAs retire
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(tfull + 8 * a);  // accumulator complete
      }
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); the two warps of a
    // quadrant take alternating 32-column chunks
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    float* stg = reinterpret_cast<float*>(smem_raw + (tmem_slot + 16 - smem_u32(smem_raw))) + (warp - 2) * (32 * kStgLd);
    int t = 0;
    for (int64_t tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++t) {
      const int a = t & 1;
      const int n0 = (int)(tile % p.n_tiles) * p.bn;
      const int64_t m0 = (tile / p.n_tiles) * BM;
      mbar_wait(tfull + 8 * a, (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c = part * 32; c < p.bn; c += 32 * (EPI / 4)) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * p.acc_cols + c), r);
        if (p.direct) {
          const int64_t row = m0 + q * 32 + lane;
          if (row < p.m) {
            if (MODE == 0) {
#pragma unroll
              for (int g2 = 0; g2 < 2; ++g2) {            // two groups of 16 columns
                const int j = n0 + c + g2 * 16;
                if (c + g2 * 16 >= p.bn || j >= p.n) continue;
                uint32_t w[8];
                const bool relu = j < p.relu_cols;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                  float x0 = __uint_as_float(r[g2 * 16 + 2 * u]), x1 = __uint_as_float(r[g2 * 16 + 2 * u + 1]);
                  if (p.bias) { x0 += __ldg(p.bias + j + 2 * u); x1 += __ldg(p.bias + j + 2 * u + 1); }
                  if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                  w[u] = pack2_bf16(x0, x1);
                }
                if (j < p.ncols0) st_row32(p.c0 + row * p.ldc0 + j, w);
                else st_row32(p.c1 + row * p.ldc1 + (j - p.ncols0), w);
              }
            } else {
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) {            // four groups of 8 fp32 columns
                const int j = n0 + c + g4 * 8;
                if (c + g4 * 8 >= p.bn || j >= p.n) continue;
                st_row32(p.cf + row * p.ldcf + j, r + g4 * 8);
              }
            }
          }
          continue;
        }
        // transpose through a per-warp staging tile (row stride 36 words: conflict-free v4
        // stores) so that global stores are row-contiguous full 32-byte sectors
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4)
          *reinterpret_cast<uint4*>(stg + lane * kStgLd + g4 * 4) = make_uint4(r[g4 * 4], r[g4 * 4 + 1], r[g4 * 4 + 2], r[g4 * 4 + 3]);
        __syncwarp();
        if (MODE == 0) {
          // 4 lanes x 16 B (8 bf16) per row, 8 rows per pass
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + (lane >> 2);
            const int cc = (lane & 3) * 8;
            const int64_t row = m0 + q * 32 + rr;
            const int j = n0 + c + cc;
            if (row >= p.m || c + cc >= p.bn || j >= p.n) continue;
            const float4 v0 = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cc);
            const float4 v1 = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cc + 4);
            float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
            if (p.bias) {                                  // n % 8 == 0: the 8 columns are all valid
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] += __ldg(p.bias + j + u);
            }
            if (j < p.relu_cols) {
#pragma unroll
              for (int u = 0; u < 8; ++u) v[u] = fmaxf(v[u], 0.f);
            }
            const uint4 packed = pack_bf16x8(v);
            if (j < p.ncols0) {
              if (p.peers.n > 0) {
                // fused all-gather: the finished row segment goes to every rank's table over NVLink
                const int64_t off = (p.peers.row_off + row) * p.ldc0 + j;
                peer_store16(p.peers, off * (int64_t)sizeof(__nv_bfloat16), packed);
              } else {
                *reinterpret_cast<uint4*>(p.c0 + row * p.ldc0 + j) = packed;
              }
            } else {
              *reinterpret_cast<uint4*>(p.c1 + row * p.ldc1 + (j - p.ncols0)) = packed;
            }
          }
        } else {
          // 8 lanes x 16 B (4 fp32) per row, 4 rows per pass
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3);
            const int cc = (lane & 7) * 4;
            const int64_t row = m0 + q * 32 + rr;
            const int j = n0 + c + cc;
            if (row >= p.m || c + cc >= p.bn || j >= p.n) continue;
            const float4 v = *reinterpret_cast<const float4*>(stg + rr * kStgLd + cc);
            float* dst = p.cf + row * p.ldcf + j;
            if (p.vec_ok && j + 4 <= p.n) {
              *reinterpret_cast<float4*>(dst) = v;
            } else {
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (j + u < p.n) dst[u] = vv[u];
            }
          }
        }
        __syncwarp();
      }
      // all of this warp's TMEM reads have completed (tcgen05.wait::ld in tmem_ld32)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + 8 * a);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

struct NtParams {
  int64_t m, n, k;          // C[m,n] += sum over k rows
  int bn, n_chunks;         // UMMA_N and number of 64-wide column chunks of the B tile
  int m_tiles, n_tiles, stages, tmem_cols;
  int kb_per_split;
  float* c; int64_t ldc;
};

__global__ void __launch_bounds__(kThreads)
nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const NtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  constexpr uint32_t chunk_bytes = BK * 128;  // 64 k-rows x 128 B (64 bf16 along M/N)
  const uint32_t a_bytes = 2 * chunk_bytes;
  const uint32_t b_bytes = (uint32_t)p.n_chunks * chunk_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bar_base = base + p.stages * stage_bytes;
  const uint32_t tmem_full = bar_base + 16 * p.stages;
  const uint32_t tmem_slot = tmem_full + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = p.m_tiles * p.n_tiles;
  const int tile = blockIdx.x % tiles;
  const int split = blockIdx.x / tiles;
  const int m0 = (tile / p.n_tiles) * BM;
  const int n0 = (tile % p.n_tiles) * p.bn;
  const int kb_all = (int)((p.k + BK - 1) / BK);
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(kb_all, kb0 + p.kb_per_split);
  const int kb_total = max(0, kb1 - kb0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (p.stages + s), 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (kb_total > 0) {
    if (warp == 0) {
      if (lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        int s = 0;
        uint32_t ph = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_base + 8 * (p.stages + s), ph ^ 1);
          const uint32_t full = bar_base + 8 * s;
          mbar_expect_tx(full, a_bytes + b_bytes);
          const uint32_t sa = base + s * stage_bytes;
          // box = {64 columns (inner), 64 node rows}; one box per 64-wide column chunk
          tma_load_2d(sa, &tmA, m0, kb * BK, full);
          tma_load_2d(sa + chunk_bytes, &tmA, m0 + 64, kb * BK, full);
          for (int c = 0; c < p.n_chunks; ++c)
            tma_load_2d(sa + a_bytes + c * chunk_bytes, &tmB, n0 + c * 64, kb * BK, full);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc(BM, p.bn, 1, 1);
        int s = 0;
        uint32_t ph = 0;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(bar_base + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + s * stage_bytes, sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // MN-major, 128B swizzle: 64 M/N elements contiguous (128 B), k-rows 128 B apart,
            // 8-row groups 1024 B apart (SBO), 64-wide M/N chunks chunk_bytes apart (LBO);
            // stepping K by 16 rows = +2048 B
            const uint64_t ad = make_desc(sa + k * UMMA_K * 128, chunk_bytes, 1024);
            const uint64_t bd = make_desc(sb + k * UMMA_K * 128, chunk_bytes, 1024);
            umma_bf16(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(bar_base + 8 * (p.stages + s));
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(tmem_full);
      }
    } else {
      const int q = warp & 3;
      mbar_wait(tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.m;
      for (int c = 0; c < p.bn; c += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
        if (!row_ok) continue;
        float* dst = p.c + (int64_t)row * p.ldc + n0 + c;
#pragma unroll
        for (int t = 0; t < 32; ++t)
          if (c + t < p.bn && n0 + c + t < p.n) atomicAdd(dst + t, __uint_as_float(r[t]));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------

}  // namespace tc

// ---- tensor-map encoding, shared by every TMA user of the library (declared in acm_common.cuh) -------------------
// A CUtensorMap is a pure function of (address, extents, pitch, box, element type, swizzle): the layer calls hand the
// same buffers of the caching allocator back every step, so the encoded descriptors are kept in a small cache keyed by
// exactly those values (SURVEY 8b "cached CUtensorMaps") instead of calling the driver 2-4 times per launch.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

namespace {
struct MapKey {
  const void* ptr; uint64_t inner, outer, pitch_bytes; uint32_t box_inner, box_outer; int kind;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && pitch_bytes == o.pitch_bytes &&
           box_inner == o.box_inner && box_outer == o.box_outer && kind == o.kind;
  }
};
constexpr int kMapCache = 64;
struct MapCache {
  MapKey key[kMapCache];
  CUtensorMap map[kMapCache];
  int used = 0, next = 0;
};
thread_local MapCache t_maps;    // per thread: no locking, the library is called from one host thread per device
}  // namespace

// kind 0: bf16 elements, 128-byte swizzle (GEMM operands); kind 1: uint32 elements, no swizzle (TMA row gather)
int tma_encode_2d(CUtensorMap* map, int kind, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                  uint32_t box_inner, uint32_t box_outer, const char* what) {
  const MapKey k{ptr, inner, outer, pitch_bytes, box_inner, box_outer, kind};
  MapCache& c = t_maps;
  for (int i = 0; i < c.used; ++i)
    if (c.key[i] == k) { *map = c.map[i]; return 0; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("TMA: cuTensorMapEncodeTiled entry point unavailable"); return ACM_ERR_UNSUPPORTED; }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || pitch_bytes % 16) {
    set_error("TMA: %s must be 16-byte aligned with a row pitch that is a multiple of 16 bytes", what);
    return ACM_ERR_BAD_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   kind == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("TMA: cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r); return ACM_ERR_BAD_ARG; }
  const int slot = c.used < kMapCache ? c.used++ : (c.next = (c.next + 1) % kMapCache);
  c.key[slot] = k;
  c.map[slot] = *map;
  return 0;
}

int tma_encode_2d_u32(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                      uint32_t box_inner, uint32_t box_outer, const char* what) {
  return tma_encode_2d(map, 1, ptr, inner, outer, pitch_bytes, box_inner, box_outer, what);
}

namespace tc {

// 2-D bf16 tensor [outer][inner] with row pitch `pitch_elems`; box {box_inner, box_outer}; 128B swizzle; OOB -> 0
static int make_map(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                    uint32_t box_inner, uint32_t box_outer, const char* what) {
  return tma_encode_2d(map, 0, ptr, inner, outer, pitch_elems * 2, box_inner, box_outer, what);
}

static int pow2_cols(int n) { int c = 32; while (c < n) c <<= 1; return c; }

static int sm_count() {
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

int g_tn_direct = 1;   // acm_set_gemm_direct_store: 1 = 256-bit row stores when the output layout allows, 0 = staging tile

template <int MODE, int EPI>
static int launch_tn_epi(const void* a, int64_t lda, const void* b, int64_t ldb, TnParams p, cudaStream_t st) {
  constexpr int kTnEpiWarps = EPI;
  constexpr int kTnThreads = TnCfg<EPI>::kThreads;
  if (p.m == 0) return 0;
  int bn = (int)((p.n < 256 ? p.n : 256));
  bn = (bn + 15) / 16 * 16;
  p.bn = bn;
  p.n_tiles = (int)((p.n + bn - 1) / bn);
  p.acc_cols = pow2_cols(bn);
  p.tmem_cols = 2 * p.acc_cols;
  const uint32_t stage_bytes = BM * BK * 2 + (((uint32_t)bn * BK * 2 + 1023u) & ~1023u);
  const size_t stg_bytes = (size_t)kTnEpiWarps * 32 * kStgLd * 4;
  int stages = (int)((226 * 1024 - 1024 - 128 - stg_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 16 * stages + 64 + stg_bytes + 1024;
  CUtensorMap ma, mb;
  int rc = make_map(&ma, a, (uint64_t)p.k, (uint64_t)p.m, (uint64_t)lda, BK, BM, "A operand");
  if (rc) return rc;
  rc = make_map(&mb, b, (uint64_t)p.k, (uint64_t)p.n, (uint64_t)ldb, BK, (uint32_t)bn, "B operand");
  if (rc) return rc;
  // Direct 256-bit row stores pay off when the epilogue's LATENCY is what limits the tile loop (K >= 128: forward
  // X.Wcat 5.3 -> 4.7 ms at the headline size); a store-bound product with a short K loop (layer-1 dX: K = 48,
  // 5 GB written) is faster through the transposition tile, whose store instructions cover 8 rows x 64 contiguous
  // bytes instead of 32 rows x 32 bytes (measured 1.65 vs 2.2 ms) -> keep the tile there.
  const bool k_long = p.k >= 128;
  if (MODE == 0) {
    const bool a0 = (reinterpret_cast<uintptr_t>(p.c0) & 31) == 0 && (p.ldc0 * 2) % 32 == 0 && p.ncols0 % 16 == 0;
    const bool a1 = p.c1 == nullptr || ((reinterpret_cast<uintptr_t>(p.c1) & 31) == 0 && (p.ldc1 * 2) % 32 == 0);
    p.direct = g_tn_direct && k_long && p.peers.n == 0 && a0 && a1 && p.n % 16 == 0 && p.relu_cols % 16 == 0;
  } else {
    p.direct = g_tn_direct && k_long && (reinterpret_cast<uintptr_t>(p.cf) & 31) == 0 && (p.ldcf * 4) % 32 == 0 && p.n % 8 == 0;
  }
  const int64_t m_tiles = (p.m + BM - 1) / BM;
  ACM_CHECK_ARG(m_tiles * BM < (1ll << 31), "tcgen05 GEMM: more than 2^31 rows");
  p.total_tiles = m_tiles * p.n_tiles;
  const int64_t grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  cudaError_t e = cudaFuncSetAttribute(tn_kernel<MODE, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tcgen05 GEMM: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  tn_kernel<MODE, EPI><<<(unsigned)grid, kTnThreads, smem, st>>>(ma, mb, p);
  ACM_LAUNCH_CHECK("tcgen05 gemm_tn");
  return 0;
}

int g_tn_epi16 = 1;   // store-bound short-K products: 16 epilogue warps (acm_set_gemm_direct_store bit 1 clears it)

template <int MODE>
static int launch_tn(const void* a, int64_t lda, const void* b, int64_t ldb, const TnParams& p, cudaStream_t st) {
  // a short K loop (layer-1 dX: K = 48) makes the tile loop epilogue-bound: measured 1.65 ms = 3.7 TB/s for 6 GB with
  // 8 epilogue warps -> four warps per TMEM quadrant there
  if (g_tn_epi16 && p.k < 128 && p.n >= 128 && p.peers.n == 0) return launch_tn_epi<MODE, 16>(a, lda, b, ldb, p, st);
  return launch_tn_epi<MODE, 8>(a, lda, b, ldb, p, st);
}

}  // namespace tc

int tc_gemm_fwd(const void* x, int64_t ldx, const void* wcat_t, void* h_lh, void* h_i, int64_t n, int64_t fin,
                int64_t fp, int relu_lh, const PeerTables* peers, cudaStream_t st) {
  tc::TnParams p{};
  if (peers) p.peers = *peers;
  p.m = n; p.n = 3 * fp; p.k = fin;
  p.c0 = (__nv_bfloat16*)h_lh; p.ldc0 = 2 * fp; p.ncols0 = (int)(2 * fp);
  p.c1 = (__nv_bfloat16*)h_i; p.ldc1 = fp;
  p.relu_cols = relu_lh ? (int)(2 * fp) : 0;
  return tc::launch_tn<0>(x, ldx, wcat_t, ldx, p, st);
}

int tc_gemm_dx(const void* dh, const void* wcat, float* dx, int64_t lddx, int64_t n, int64_t fin, int64_t fp,
               cudaStream_t st) {
  tc::TnParams p{};
  p.m = n; p.n = fin; p.k = 3 * fp;
  p.cf = dx; p.ldcf = lddx;
  p.vec_ok = (lddx % 4 == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0);
  return tc::launch_tn<1>(dh, 3 * fp, wcat, 3 * fp, p, st);
}

// C[m, ncols] (fp32, row stride ldc, atomically accumulated) += A[k_rows, m]^T . B[k_rows, ncols]
int tc_gemm_atb(const void* a, int64_t lda, const void* b, int64_t ldb, float* c, int64_t ldc,
                int64_t k_rows, int64_t m, int64_t ncols, cudaStream_t st) {
  using namespace tc;
  if (k_rows == 0 || m == 0 || ncols == 0) return 0;
  NtParams p{};
  p.m = m; p.n = ncols; p.k = k_rows;
  int bn = (int)(p.n < 256 ? p.n : 256);
  bn = (bn + 15) / 16 * 16;
  p.bn = bn;
  p.n_chunks = (bn + 63) / 64;
  p.m_tiles = (int)((p.m + BM - 1) / BM);
  p.n_tiles = (int)((p.n + bn - 1) / bn);
  p.tmem_cols = pow2_cols(bn);
  p.c = c; p.ldc = ldc;
  const uint32_t stage_bytes = (2 + p.n_chunks) * BK * 128;
  int stages = (int)(190 * 1024 / stage_bytes);
  if (stages > 8) stages = 8;
  p.stages = stages;
  const int tiles = p.m_tiles * p.n_tiles;
  const int kb_all = (int)((k_rows + BK - 1) / BK);
  // one CTA per SM (190 KB of smem each): never more CTAs than SMs, or the doubly loaded SMs
  // set the critical path (measured: 150 CTAs on 148 SMs ran at half speed)
  int splits = sm_count() / tiles;
  if (splits > kb_all) splits = kb_all;
  if (splits < 1) splits = 1;
  p.kb_per_split = (kb_all + splits - 1) / splits;
  splits = (kb_all + p.kb_per_split - 1) / p.kb_per_split;
  const size_t smem = (size_t)stages * stage_bytes + 16 * stages + 16 + 1024;
  CUtensorMap ma, mb;
  // MN-major operands: inner dimension = feature columns, outer = node rows (K)
  int rc = make_map(&ma, a, (uint64_t)m, (uint64_t)k_rows, (uint64_t)lda, 64, BK, "A^T operand");
  if (rc) return rc;
  rc = make_map(&mb, b, (uint64_t)ncols, (uint64_t)k_rows, (uint64_t)ldb, 64, BK, "B operand (MN-major)");
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("tcgen05 GEMM: smem attribute: %s", cudaGetErrorString(e)); return (int)e; }
  nt_kernel<<<(unsigned)(tiles * splits), kThreads, smem, st>>>(ma, mb, p);
  ACM_LAUNCH_CHECK("tcgen05 gemm_nt");
  return 0;
}

int tc_gemm_dw(const void* x, int64_t ldx, const void* dh, float* dwcat, int64_t n, int64_t fin, int64_t fp,
               cudaStream_t st) {
  return tc_gemm_atb(x, ldx, dh, 3 * fp, dwcat, 3 * fp, n, fin, 3 * fp, st);
}

// C[m, n] (bf16, row stride ldc) = A[m,k] . B^T with B given K-major as b_nk [n, k]
int tc_gemm_ab(const void* a, int64_t lda, const void* b_nk, int64_t ldb, void* c, int64_t ldc,
               int64_t m, int64_t n, int64_t k, int relu, cudaStream_t st) {
  tc::TnParams p{};
  p.m = m; p.n = n; p.k = k;
  p.c0 = (__nv_bfloat16*)c; p.ldc0 = ldc; p.ncols0 = (int)n;
  p.c1 = nullptr; p.ldc1 = 0;
  p.relu_cols = relu ? (int)n : 0;
  return tc::launch_tn<0>(a, lda, b_nk, ldb, p, st);
}

// y[m,n] (bf16) = relu?(x[m,k] . w[n,k]^T + bias[n])
int tc_gemm_linear(const void* x, int64_t ldx, const void* w_nk, int64_t ldw, const float* bias, void* y, int64_t ldy,
                   int64_t m, int64_t n, int64_t k, int relu, cudaStream_t st) {
  tc::TnParams p{};
  p.m = m; p.n = n; p.k = k;
  p.c0 = (__nv_bfloat16*)y; p.ldc0 = ldy; p.ncols0 = (int)n;
  p.c1 = nullptr; p.ldc1 = 0;
  p.relu_cols = relu ? (int)n : 0;
  p.bias = bias;
  return tc::launch_tn<0>(x, ldx, w_nk, ldw, p, st);
}

}  // namespace acm

namespace acm { namespace fused { extern int g_fused_early; } }

extern "C" int acm_set_gemm_direct_store(int on) {
  acm::tc::g_tn_direct = (on & 1) ? 1 : 0;
  acm::tc::g_tn_epi16 = (on & 2) ? 0 : 1;     // bit 1 set: keep 8 epilogue warps for short-K products too (A/B switch)
  acm::fused::g_fused_early = (on & 4) ? 1 : 0;   // bit 2 set: early TMEM release in the fused forward (A/B switch)
  return 0;
}
