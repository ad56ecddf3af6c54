// Parameter packing / gradient unpacking in ONE launch each.  The layer keeps the reference's
// parameter tensors (weight_low/high/mlp [fin,f], att_vec_* [f,1], att_vec [K,K], LayerNorm
// gamma/beta [f]; ACM-Pytorch/models/layers.py:41-67) and the kernels consume them as
//   wcat   [fin, 3*fp]  T   = [W_low | W_high | W_mlp], zero padded columns
//   wcat_t [3*fp, ldt]  T   = its transpose (K-major B operand of the tcgen05 GEMM), zero padded
//   pack   [12*fp+16]  fp32 (layout in acm_b200.h)
// Doing this with torch slicing costs ~25 micro-launches per layer call, which dominates the
// step on the small reference graphs (Cora: 2 708 nodes).
#include "acm_common.cuh"

namespace acm {

struct PackArgs {
  const float* w[3];
  const float* a[4];
  const float* gamma[4];
  const float* beta[4];
  const float* att_vec;
  int fin, f, fp, k, ln, ldt, t_bf16;
  void* wcat;
  void* wcat_t;
  float* pack;
};

__device__ __forceinline__ void store_t(void* base, int64_t idx, float v, int bf16) {
  if (bf16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(base)[idx] = v;
}

__global__ void pack_params_kernel(const PackArgs p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w3 = 3 * p.fp;
  const int64_t n_w = (int64_t)p.fin * w3;
  const int64_t n_wt = p.wcat_t ? (int64_t)w3 * p.ldt : 0;
  const int64_t n_pack = 16 * p.fp + 32;
  if (p.ln && blockIdx.x == 0) {
    // per-channel sums of the derived tail (tiny: one warp per channel)
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (w < p.k) {
      float sb = 0.f, sg = 0.f;
      for (int j = l; j < p.f; j += 32) {
        sb += p.beta[w][j] * p.a[w][j];
        sg += p.gamma[w][j] * p.a[w][j];
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
      }
      if (l == 0) {
        p.pack[pack_off_sum_ba(p.fp) + w] = sb;
        p.pack[pack_off_sum_ga(p.fp) + w] = sg;
      }
    }
  }
  if (i < n_w) {
    const int r = (int)(i / w3), c = (int)(i % w3);
    const int k = c / p.fp, j = c % p.fp;
    store_t(p.wcat, i, j < p.f ? p.w[k][(int64_t)r * p.f + j] : 0.f, p.t_bf16);
  } else if (i < n_w + n_wt) {
    const int64_t q = i - n_w;
    const int c = (int)(q / p.ldt), r = (int)(q % p.ldt);
    const int k = c / p.fp, j = c % p.fp;
    store_t(p.wcat_t, q, (r < p.fin && j < p.f) ? p.w[k][(int64_t)r * p.f + j] : 0.f, p.t_bf16);
  } else if (i < n_w + n_wt + n_pack) {
    const int q = (int)(i - n_w - n_wt);
    float v = 0.f;
    if (q < 4 * p.fp) {
      const int k = q / p.fp, j = q % p.fp;
      if (k < p.k && j < p.f) v = p.a[k][j];
    } else if (q < 4 * p.fp + 16) {
      const int jj = (q - 4 * p.fp) / 4, kk = (q - 4 * p.fp) % 4;
      if (jj < p.k && kk < p.k) v = p.att_vec[jj * p.k + kk];
    } else if (q >= 16 * p.fp + 16) {
      // sums: written above by block 0 (LayerNorm) -- zero the unused slots only
      const int e = q - (16 * p.fp + 16);
      if (p.ln && e < 8 && (e & 3) < p.k) return;
    } else if (p.ln) {
      const int q2 = q - 4 * p.fp - 16;
      const int sec = q2 / (4 * p.fp);            // 0 gamma, 1 beta, 2 gamma*a
      const int q3 = q2 - sec * 4 * p.fp;
      const int k = q3 / p.fp, j = q3 % p.fp;
      if (k < p.k && j < p.f) v = sec == 0 ? p.gamma[k][j] : (sec == 1 ? p.beta[k][j] : p.gamma[k][j] * p.a[k][j]);
    }
    p.pack[q] = v;
  }
}

struct UnpackArgs {
  const float* dwcat;   // [fin, 3fp]
  const float* dpack;   // [12fp+16]
  float* dw[3];         // [fin, f]
  float* da[4];         // [f]
  float* dgamma[4];
  float* dbeta[4];
  float* datt_vec;      // [k,k]
  int fin, f, fp, k, ln;
};

__global__ void unpack_grads_kernel(const UnpackArgs p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_w = (int64_t)3 * p.fin * p.f;
  if (i < n_w) {
    const int k = (int)(i / ((int64_t)p.fin * p.f));
    const int64_t q = i - (int64_t)k * p.fin * p.f;
    const int r = (int)(q / p.f), j = (int)(q % p.f);
    p.dw[k][q] = p.dwcat[(int64_t)r * 3 * p.fp + k * p.fp + j];
  } else {
    const int q = (int)(i - n_w);
    const int per = p.f;  // per-channel vectors
    if (q < p.k * per) {
      const int k = q / per, j = q % per;
      if (p.da[k]) p.da[k][j] = p.dpack[k * p.fp + j];
      if (p.ln && p.dgamma[k]) {
        p.dgamma[k][j] = p.dpack[4 * p.fp + 16 + k * p.fp + j];
        p.dbeta[k][j] = p.dpack[8 * p.fp + 16 + k * p.fp + j];
      }
    } else if (q < p.k * per + p.k * p.k) {
      const int e = q - p.k * per;
      p.datt_vec[e] = p.dpack[4 * p.fp + (e / p.k) * 4 + (e % p.k)];
    }
  }
}

}  // namespace acm

extern "C" int acm_pack_params(int dtype, int fin, int f, int fp, int k_channels, int ln_live, int ldt,
                               const float* w_low, const float* w_high, const float* w_mlp,
                               const float* const* a_vecs, const float* att_vec,
                               const float* const* ln_gamma, const float* const* ln_beta,
                               void* wcat, void* wcat_t, float* pack, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "pack_params: bad dtype %d", dtype);
  ACM_CHECK_ARG(w_low && w_high && w_mlp && a_vecs && att_vec && wcat && pack, "pack_params: null pointer");
  ACM_CHECK_ARG(k_channels == 3 || k_channels == 4, "pack_params: k_channels must be 3 or 4");
  ACM_CHECK_ARG(!ln_live || (ln_gamma && ln_beta), "pack_params: LayerNorm needs gamma and beta");
  ACM_CHECK_ARG(!wcat_t || ldt >= fin, "pack_params: ldt < fin");
  PackArgs p{};
  p.w[0] = w_low; p.w[1] = w_high; p.w[2] = w_mlp;
  for (int k = 0; k < k_channels; ++k) {
    p.a[k] = a_vecs[k];
    ACM_CHECK_ARG(p.a[k], "pack_params: null att vector %d", k);
    if (ln_live) { p.gamma[k] = ln_gamma[k]; p.beta[k] = ln_beta[k]; }
  }
  p.att_vec = att_vec; p.fin = fin; p.f = f; p.fp = fp; p.k = k_channels; p.ln = ln_live; p.ldt = ldt;
  p.t_bf16 = (dtype == ACM_BF16); p.wcat = wcat; p.wcat_t = wcat_t; p.pack = pack;
  const int64_t total = (int64_t)fin * 3 * fp + (wcat_t ? (int64_t)3 * fp * ldt : 0) + 16 * fp + 32;
  pack_params_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  ACM_LAUNCH_CHECK("pack_params");
  return 0;
}

extern "C" int acm_unpack_grads(int fin, int f, int fp, int k_channels, int ln_live,
                                const float* dwcat, const float* dpack,
                                float* dw_low, float* dw_high, float* dw_mlp,
                                float* const* da, float* datt_vec, float* const* dgamma, float* const* dbeta,
                                void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dwcat && dpack && dw_low && dw_high && dw_mlp && da && datt_vec, "unpack_grads: null pointer");
  ACM_CHECK_ARG(k_channels == 3 || k_channels == 4, "unpack_grads: k_channels must be 3 or 4");
  UnpackArgs p{};
  p.dwcat = dwcat; p.dpack = dpack; p.dw[0] = dw_low; p.dw[1] = dw_high; p.dw[2] = dw_mlp;
  for (int k = 0; k < k_channels; ++k) {
    p.da[k] = da[k];
    if (ln_live && dgamma && dbeta) { p.dgamma[k] = dgamma[k]; p.dbeta[k] = dbeta[k]; }
  }
  p.datt_vec = datt_vec; p.fin = fin; p.f = f; p.fp = fp; p.k = k_channels; p.ln = ln_live && dgamma && dbeta;
  const int64_t total = (int64_t)3 * fin * f + (int64_t)k_channels * f + k_channels * k_channels;
  unpack_grads_kernel<<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  ACM_LAUNCH_CHECK("unpack_grads");
  return 0;
}
