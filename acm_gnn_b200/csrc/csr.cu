// Operator construction helpers (integer / degree work, bit-exact targets) and the
// library's error plumbing.
//
// Stand in for the reference's dense/scipy preprocessing:
//   ACM-Pytorch/utils.py:421-438 (normalize_tensor), 626-628 (adj_low / adj_high)
//   ACM-Geometric/utils.py:5-19, train.py:76-80
#include <stdarg.h>

#include <atomic>

#include "acm_common.cuh"

namespace acm {

static thread_local char g_err[512] = "no error";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int g_narrow_row_hint = 1;   // default on: measured 1-5 % faster on the layer-1 gathers, bit-identical

// rowptr[r] = first position e with row[e] >= r   (row sorted ascending)
__global__ void rowptr_kernel(const int64_t* __restrict__ row, int64_t nnz, int64_t n, int64_t* __restrict__ rowptr) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > nnz) return;
  const int64_t lo = (e == 0) ? -1 : row[e - 1];
  const int64_t hi = (e == nnz) ? n : row[e];
  for (int64_t r = lo + 1; r <= hi; ++r) rowptr[r] = e;  // rows (lo, hi] start at e
}

// one thread per row: sequential fp32 sum in column order, IEEE reciprocal and product
__global__ void degree_kernel(const int64_t* __restrict__ rowptr, const float* __restrict__ mult, int64_t n,
                              float* __restrict__ rowsum, float* __restrict__ rinv, float* __restrict__ w) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t e0 = rowptr[r], e1 = rowptr[r + 1];
  float s = 0.f;
  for (int64_t e = e0; e < e1; ++e) s = __fadd_rn(s, mult[e]);
  float ri = __fdiv_rn(1.0f, s);
  if (isinf(ri)) ri = 0.f;
  if (rowsum) rowsum[r] = s;
  if (rinv) rinv[r] = ri;
  if (w)
    for (int64_t e = e0; e < e1; ++e) w[e] = __fmul_rn(ri, mult[e]);
}

// warp per row; each lane handles edges (i,j) of row i and binary-searches i in row j
__global__ void transpose_values_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                        const float* __restrict__ w, int64_t n, float* __restrict__ wt,
                                        int* __restrict__ not_sym) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int64_t e0 = rowptr[i], e1 = rowptr[i + 1];
  for (int64_t e = e0 + lane; e < e1; e += 32) {
    const int64_t j = col[e];
    int64_t lo = rowptr[j], hi = rowptr[j + 1];
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (col[mid] < i) lo = mid + 1; else hi = mid;
    }
    if (lo < rowptr[j + 1] && col[lo] == i) wt[lo] = w[e];
    else *not_sym = 1;
  }
}

template <typename TO>
__global__ void cast_pad_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t ld_src,
                                TO* __restrict__ dst, int64_t ld_dst) {
  // one thread per 4 consecutive destination elements
  const int64_t per_row = ld_dst >> 2;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * per_row) return;
  const int64_t r = idx / per_row;
  const int64_t c = (idx - r * per_row) << 2;
  float v[4];
  const float* s = src + r * ld_src + c;
  if (c + 4 <= cols && ((ld_src & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(s));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = (c + t < cols) ? __ldg(s + t) : 0.f;
  }
  TO* d = dst + r * ld_dst + c;
  if constexpr (sizeof(TO) == 2) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(d) = o;
  } else {
    *reinterpret_cast<float4*>(d) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

}  // namespace acm

extern "C" int acm_version(void) { return 104; }  // 0.2.2: + acm_linear_fwd; 103: + acm_glue_fwd / acm_glue_bwd (102: acm_fused_agg_fwd, acm_spmm_t_bwd_rank1, table_mode of acm_mix_bwd)
extern "C" const char* acm_last_error_string(void) { return acm::g_err; }
extern "C" int64_t acm_launch_count(void) { return acm::g_launches.load(std::memory_order_relaxed); }

extern "C" int acm_csr_rowptr(const int64_t* row_sorted, int64_t nnz, int64_t n, int64_t* rowptr, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(rowptr && (row_sorted || nnz == 0), "csr_rowptr: null pointer");
  ACM_CHECK_ARG(n >= 0 && nnz >= 0, "csr_rowptr: negative size");
  const int64_t threads = nnz + 1;
  const int64_t blocks = (threads + 255) / 256;
  rowptr_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(row_sorted, nnz, n, rowptr);
  ACM_LAUNCH_CHECK("csr_rowptr");
  return 0;
}

extern "C" int acm_degree_normalise(const int64_t* rowptr, const float* mult, int64_t n,
                                    float* rowsum, float* rinv, float* w, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(rowptr && mult, "degree_normalise: null pointer");
  if (n == 0) return 0;
  degree_kernel<<<(unsigned)((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rowptr, mult, n, rowsum, rinv, w);
  ACM_LAUNCH_CHECK("degree_normalise");
  return 0;
}

extern "C" int acm_csr_transpose_values(const int64_t* rowptr, const int32_t* col, const float* w, int64_t n,
                                        float* w_t, int* not_symmetric, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(rowptr && col && w && w_t && not_symmetric, "csr_transpose_values: null pointer");
  if (n == 0) return 0;
  const int64_t blocks = (n * 32 + 255) / 256;
  transpose_values_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rowptr, col, w, n, w_t, not_symmetric);
  ACM_LAUNCH_CHECK("csr_transpose_values");
  return 0;
}

extern "C" int acm_cast_pad(const float* src, int64_t rows, int64_t cols, int64_t ld_src,
                            void* dst, int dst_dtype, int64_t ld_dst, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(src && dst, "cast_pad: null pointer");
  ACM_CHECK_ARG(ld_dst % 4 == 0 && ld_dst >= cols, "cast_pad: ld_dst must be a multiple of 4 and >= cols");
  ACM_CHECK_ARG(dst_dtype == ACM_F32 || dst_dtype == ACM_BF16, "cast_pad: bad dtype");
  const int64_t total = rows * (ld_dst >> 2);
  if (total == 0) return 0;
  const int64_t blocks = (total + 255) / 256;
  ACM_CHECK_ARG(blocks < (1ll << 31), "cast_pad: tensor too large for one launch");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dst_dtype == ACM_BF16)
    cast_pad_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(src, rows, cols, ld_src, (__nv_bfloat16*)dst, ld_dst);
  else
    cast_pad_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(src, rows, cols, ld_src, (float*)dst, ld_dst);
  ACM_LAUNCH_CHECK("cast_pad");
  return 0;
}

extern "C" int acm_set_narrow_row_hint(int on) {
  acm::g_narrow_row_hint = on ? 1 : 0;
  return 0;
}
