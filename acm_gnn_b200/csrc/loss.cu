// Fused log_softmax + NLL loss on the training rows, forward and gradient in ONE pass.
// Replaces (for callers that opt in) the reference's loss glue
//   output = F.log_softmax(output, dim=1); loss = criterion(output[idx_train], labels[idx_train])
//   (ACM-Pytorch/utils.py:567-568, ACM-Geometric/train.py:133-134) and its autograd:
//   log_softmax, index_select, nll_loss, nll_loss_backward, index_put, log_softmax_backward
// = 6 ATen launches moving ~10x the bytes.  SURVEY.md 8(f) rank 3 ("fold log_softmax+NLL").
//   loss      = scale * sum_{i in train} ( logsumexp(x_i) - x_i[label_i] )
//   dlogits_i = scale * (softmax(x_i) - onehot(label_i))   for train rows, 0 otherwise
#include "acm_common.cuh"

namespace acm {

template <int CMAX>
__global__ void __launch_bounds__(256)
nll_kernel(const float* __restrict__ x, int64_t ld, int64_t n, int c, const int64_t* __restrict__ labels,
           const uint8_t* __restrict__ mask, float scale, float* __restrict__ loss, float* __restrict__ dx, int64_t lddx,
           int vec, int vec_d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float li = 0.f;
  if (i < n) {
    // a row whose label is outside [0, c) (e.g. the -1 of unlabeled nodes) is treated as unselected:
    // torch's nll_loss would raise; reading x[label] out of bounds is never an option
    const int64_t lab64 = labels[i];
    const bool on = (mask ? (mask[i] != 0) : true) && lab64 >= 0 && lab64 < c;
    float v[CMAX];
    const float* xr = x + i * ld;
    if (on) {
      float mx = -INFINITY;
      if (vec) {  // c % 4 == 0, 16-byte aligned rows: 128-bit loads
#pragma unroll
        for (int j = 0; j < CMAX; j += 4) {
          if (j < c) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(xr + j));
            v[j] = t.x; v[j + 1] = t.y; v[j + 2] = t.z; v[j + 3] = t.w;
          } else {
            v[j] = v[j + 1] = v[j + 2] = v[j + 3] = -INFINITY;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < CMAX; ++j) v[j] = (j < c) ? __ldg(xr + j) : -INFINITY;
      }
#pragma unroll
      for (int j = 0; j < CMAX; ++j) mx = fmaxf(mx, v[j]);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < CMAX; ++j) {
        v[j] = (j < c) ? expf(v[j] - mx) : 0.f;
        s += v[j];
      }
      const int lab = (int)lab64;
      const float xl = __ldg(xr + lab);  // L1 hit
      li = (logf(s) + mx - xl) * scale;
      if (dx) {
        const float rs = scale / s;
        float* d = dx + i * lddx;
        if (vec_d) {
#pragma unroll
          for (int j = 0; j < CMAX; j += 4)
            if (j < c)
              *reinterpret_cast<float4*>(d + j) = make_float4(v[j] * rs - (j == lab ? scale : 0.f), v[j + 1] * rs - (j + 1 == lab ? scale : 0.f),
                                                              v[j + 2] * rs - (j + 2 == lab ? scale : 0.f), v[j + 3] * rs - (j + 3 == lab ? scale : 0.f));
        } else {
#pragma unroll
          for (int j = 0; j < CMAX; ++j)
            if (j < c) d[j] = v[j] * rs - (j == lab ? scale : 0.f);
        }
      }
    } else if (dx) {
      float* d = dx + i * lddx;
      if (vec_d) {
#pragma unroll
        for (int j = 0; j < CMAX; j += 4)
          if (j < c) *reinterpret_cast<float4*>(d + j) = make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < CMAX; ++j)
          if (j < c) d[j] = 0.f;
      }
    }
  }
  // block reduction of the loss, one atomic per block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) li += __shfl_xor_sync(0xffffffffu, li, o);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = li;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = part[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(loss, t);
  }
}

}  // namespace acm

extern "C" int acm_nll_log_softmax(const float* logits, int64_t ld, int64_t n_rows, int n_classes,
                                   const int64_t* labels, const uint8_t* mask, float scale,
                                   float* loss_sum, float* dlogits, int64_t ld_d, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(logits && labels && loss_sum, "nll_log_softmax: null pointer");
  ACM_CHECK_ARG(n_classes >= 1 && n_classes <= 64, "nll_log_softmax: 1 <= classes <= 64 supported (got %d)", n_classes);
  if (n_rows == 0) return 0;
  const int64_t blocks = (n_rows + 255) / 256;
  ACM_CHECK_ARG(blocks < (1ll << 31), "nll_log_softmax: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int vec = (n_classes % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
  const int vec_d = dlogits && (n_classes % 4 == 0) && (ld_d % 4 == 0) && ((reinterpret_cast<uintptr_t>(dlogits) & 15) == 0);
  if (n_classes <= 8) nll_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(logits, ld, n_rows, n_classes, labels, mask, scale, loss_sum, dlogits, ld_d, vec, vec_d);
  else if (n_classes <= 16) nll_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(logits, ld, n_rows, n_classes, labels, mask, scale, loss_sum, dlogits, ld_d, vec, vec_d);
  else if (n_classes <= 32) nll_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(logits, ld, n_rows, n_classes, labels, mask, scale, loss_sum, dlogits, ld_d, vec, vec_d);
  else nll_kernel<64><<<(unsigned)blocks, 256, 0, st>>>(logits, ld, n_rows, n_classes, labels, mask, scale, loss_sum, dlogits, ld_d, vec, vec_d);
  ACM_LAUNCH_CHECK("nll_log_softmax");
  return 0;
}

extern "C" int acm_set_l2_fetch_granularity(int bytes) {
  cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
  if (e != cudaSuccess) {
    acm::set_error("cudaLimitMaxL2FetchGranularity=%d: %s", bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}
