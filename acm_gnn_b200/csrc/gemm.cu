// C-ABI entry points of the dense feature transforms (forward + both backward products),
// dispatching to the CUDA-core fp32 path (gemm_simt.cu) or the tcgen05 path (gemm_tc.cu).
// Reference lines replaced: ACM-Pytorch/models/layers.py:163-165,179-194 (three torch.mm
// sharing the left operand) and their autograd.
#include "acm_common.cuh"
#include "gemm_params.cuh"

namespace acm {

// gemm_tc.cu
int tc_gemm_fwd(const void* x, int64_t ldx, const void* wcat_t, void* h_lh, void* h_i, int64_t n, int64_t fin,
                int64_t fp, int relu_lh, const PeerTables* peers, cudaStream_t st);
int tc_gemm_dw(const void* x, int64_t ldx, const void* dh, float* dwcat, int64_t n, int64_t fin, int64_t fp,
               cudaStream_t st);
int tc_gemm_dx(const void* dh, const void* wcat, float* dx, int64_t lddx, int64_t n, int64_t fin, int64_t fp,
               cudaStream_t st);
int tc_gemm_atb(const void* a, int64_t lda, const void* b, int64_t ldb, float* c, int64_t ldc,
                int64_t k_rows, int64_t m, int64_t ncols, cudaStream_t st);
int tc_gemm_ab(const void* a, int64_t lda, const void* b_nk, int64_t ldb, void* c, int64_t ldc,
               int64_t m, int64_t n, int64_t k, int relu, cudaStream_t st);
int tc_gemm_linear(const void* x, int64_t ldx, const void* w_nk, int64_t ldw, const float* bias, void* y, int64_t ldy,
                   int64_t m, int64_t n, int64_t k, int relu, cudaStream_t st);
}  // namespace acm

extern "C" int acm_gemm_ab(int impl, int dtype, const void* a, int64_t lda, const void* b_kn, int64_t ldb_kn,
                           const void* b_nk, int64_t ldb_nk, void* c, int64_t ldc,
                           int64_t m, int64_t n, int64_t k, int relu, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "gemm_ab: bad dtype %d", dtype);
  ACM_CHECK_ARG(a && c, "gemm_ab: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == ACM_GEMM_TCGEN05) {
    ACM_CHECK_ARG(dtype == ACM_BF16 && b_nk, "gemm_ab: the tcgen05 path needs bf16 storage and the K-major B operand");
    ACM_CHECK_ARG(n % 8 == 0, "gemm_ab: tcgen05 path needs n %% 8 == 0");
    return tc_gemm_ab(a, lda, b_nk, ldb_nk, c, ldc, m, n, k, relu, st);
  }
  ACM_CHECK_ARG(impl == ACM_GEMM_SIMT && b_kn, "gemm_ab: SIMT path needs the [k,n] B operand");
  GemmParams p{};
  p.a = a; p.a_rs = lda; p.a_cs = 1;
  p.b = b_kn; p.b_rs = ldb_kn; p.b_cs = 1;
  p.c0 = c; p.ldc0 = ldc; p.ncols0 = n; p.c1 = nullptr; p.ldc1 = 0;
  p.m = m; p.n = n; p.k = k;
  p.c_bf16 = (dtype == ACM_BF16); p.relu_cols = relu ? (int)n : 0; p.atomic = 0;
  return gemm_simt(dtype, p, 1, st);
}

// nn.Linear (+ optional relu) of the reference's MLP helper (ACM-Pytorch/models/layers.py:245-285; the acmgcn++
// branch xX = relu(mlpX(x)), models.py:116-122) on the tcgen05 path: bf16 operands, fp32 accumulation, fp32 bias.
extern "C" int acm_linear_fwd(const void* x, int64_t ldx, const void* w_nk, int64_t ldw, const float* bias,
                              void* y, int64_t ldy, int64_t m, int64_t n, int64_t k, int relu, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(x && w_nk && y, "linear_fwd: null pointer");
  ACM_CHECK_ARG(m >= 0 && n >= 8 && n % 8 == 0 && k >= 8 && k % 8 == 0, "linear_fwd: need n %% 8 == 0 and k %% 8 == 0 (got n=%lld k=%lld)", (long long)n, (long long)k);
  ACM_CHECK_ARG(ldx >= k && ldw >= k && ldy >= n && ldx % 8 == 0 && ldw % 8 == 0 && ldy % 8 == 0, "linear_fwd: row strides must be multiples of 8 elements and cover the rows");
  return tc_gemm_linear(x, ldx, w_nk, ldw, bias, y, ldy, m, n, k, relu, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int acm_gemm_atb(int impl, int dtype, const void* a, int64_t lda, const void* b, int64_t ldb,
                            float* c, int64_t ldc, int64_t k_rows, int64_t m, int64_t n, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "gemm_atb: bad dtype %d", dtype);
  ACM_CHECK_ARG(a && b && c, "gemm_atb: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == ACM_GEMM_TCGEN05) {
    ACM_CHECK_ARG(dtype == ACM_BF16, "gemm_atb: the tcgen05 path needs bf16 storage");
    return tc_gemm_atb(a, lda, b, ldb, c, ldc, k_rows, m, n, st);
  }
  ACM_CHECK_ARG(impl == ACM_GEMM_SIMT, "gemm_atb: unknown impl %d", impl);
  GemmParams p{};
  p.a = a; p.a_rs = 1; p.a_cs = lda;
  p.b = b; p.b_rs = ldb; p.b_cs = 1;
  p.c0 = c; p.ldc0 = ldc; p.ncols0 = n; p.c1 = nullptr; p.ldc1 = 0;
  p.m = m; p.n = n; p.k = k_rows;
  p.c_bf16 = 0; p.relu_cols = 0; p.atomic = 1;
  const int64_t tiles = ((m + 63) / 64) * ((n + 63) / 64);
  int64_t splits = (148 * 8 + tiles - 1) / tiles;
  const int64_t max_splits = (k_rows + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return gemm_simt(dtype, p, (int)splits, st);
}

extern "C" int acm_gemm_xw_fwd(int impl, int dtype, const void* x, int64_t ldx, const void* wcat, const void* wcat_t,
                               void* h_lh, void* h_i, int64_t n, int64_t fin, int64_t fp, int relu_lh, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "gemm_xw_fwd: bad dtype %d", dtype);
  ACM_CHECK_ARG(x && h_lh && h_i, "gemm_xw_fwd: null pointer");
  ACM_CHECK_ARG(ldx >= fin && fin >= 1, "gemm_xw_fwd: need ldx >= fin >= 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == ACM_GEMM_TCGEN05) {
    ACM_CHECK_ARG(dtype == ACM_BF16, "gemm_xw_fwd: the tcgen05 path computes in bf16; storage dtype must be bf16");
    ACM_CHECK_ARG(wcat_t, "gemm_xw_fwd: tcgen05 path needs wcat_t");
    return tc_gemm_fwd(x, ldx, wcat_t, h_lh, h_i, n, fin, fp, relu_lh, nullptr, st);
  }
  ACM_CHECK_ARG(impl == ACM_GEMM_SIMT, "gemm_xw_fwd: unknown impl %d", impl);
  ACM_CHECK_ARG(wcat, "gemm_xw_fwd: SIMT path needs wcat");
  GemmParams p{};
  p.a = x; p.a_rs = ldx; p.a_cs = 1;
  p.b = wcat; p.b_rs = 3 * fp; p.b_cs = 1;
  p.c0 = h_lh; p.ldc0 = 2 * fp; p.ncols0 = 2 * fp;
  p.c1 = h_i; p.ldc1 = fp;
  p.m = n; p.n = 3 * fp; p.k = fin;
  p.c_bf16 = (dtype == ACM_BF16); p.relu_cols = relu_lh ? (int)(2 * fp) : 0; p.atomic = 0;
  return gemm_simt(dtype, p, 1, st);
}

extern "C" int acm_gemm_bwd_dw(int impl, int dtype, const void* x, int64_t ldx, const void* dh,
                               float* dwcat, int64_t n, int64_t fin, int64_t fp, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "gemm_bwd_dw: bad dtype %d", dtype);
  ACM_CHECK_ARG(x && dh && dwcat, "gemm_bwd_dw: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == ACM_GEMM_TCGEN05) {
    ACM_CHECK_ARG(dtype == ACM_BF16, "gemm_bwd_dw: the tcgen05 path needs bf16 storage");
    return tc_gemm_dw(x, ldx, dh, dwcat, n, fin, fp, st);
  }
  ACM_CHECK_ARG(impl == ACM_GEMM_SIMT, "gemm_bwd_dw: unknown impl %d", impl);
  GemmParams p{};
  p.a = x; p.a_rs = 1; p.a_cs = ldx;          // A = X^T : [fin, n]
  p.b = dh; p.b_rs = 3 * fp; p.b_cs = 1;      // B = dH  : [n, 3fp]
  p.c0 = dwcat; p.ldc0 = 3 * fp; p.ncols0 = 3 * fp; p.c1 = nullptr; p.ldc1 = 0;
  p.m = fin; p.n = 3 * fp; p.k = n;
  p.c_bf16 = 0; p.relu_cols = 0; p.atomic = 1;  // dwcat is zeroed by the caller
  const int64_t tiles = ((fin + 63) / 64) * ((3 * fp + 63) / 64);
  int64_t splits = (148 * 8 + tiles - 1) / tiles;
  const int64_t max_splits = (n + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  return gemm_simt(dtype, p, (int)splits, st);
}

extern "C" int acm_gemm_bwd_dx(int impl, int dtype, const void* dh, const void* wcat, const void* wcat_t, int64_t ldwt,
                               float* dx, int64_t lddx, int64_t n, int64_t fin, int64_t fp, void* stream) {
  using namespace acm;
  (void)wcat_t; (void)ldwt;
  ACM_CHECK_ARG(dtype == ACM_F32 || dtype == ACM_BF16, "gemm_bwd_dx: bad dtype %d", dtype);
  ACM_CHECK_ARG(dh && wcat && dx, "gemm_bwd_dx: null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (impl == ACM_GEMM_TCGEN05) {
    ACM_CHECK_ARG(dtype == ACM_BF16, "gemm_bwd_dx: the tcgen05 path needs bf16 storage");
    return tc_gemm_dx(dh, wcat, dx, lddx, n, fin, fp, st);
  }
  ACM_CHECK_ARG(impl == ACM_GEMM_SIMT, "gemm_bwd_dx: unknown impl %d", impl);
  GemmParams p{};
  p.a = dh; p.a_rs = 3 * fp; p.a_cs = 1;       // A = dH : [n, 3fp]
  p.b = wcat; p.b_rs = 1; p.b_cs = 3 * fp;     // B = Wcat^T : [3fp, fin], element (k,j) = wcat[j*3fp + k]
  p.c0 = dx; p.ldc0 = lddx; p.ncols0 = fin; p.c1 = nullptr; p.ldc1 = 0;
  p.m = n; p.n = fin; p.k = 3 * fp;
  p.c_bf16 = 0; p.relu_cols = 0; p.atomic = 0;
  return gemm_simt(dtype, p, 1, st);
}

extern "C" int acm_gemm_xw_fwd_push(const void* x, int64_t ldx, const void* wcat_t, void* const* peer_tables, int n_peers,
                                    int64_t row_off, void* multicast_table, void* h_i, int64_t n, int64_t fin,
                                    int64_t fp, int relu_lh, void* stream) {
  using namespace acm;
  ACM_CHECK_ARG(x && wcat_t && peer_tables && h_i, "gemm_xw_fwd_push: null pointer");
  ACM_CHECK_ARG(n_peers >= 1 && n_peers <= kMaxPeers, "gemm_xw_fwd_push: 1 <= n_peers <= %d", kMaxPeers);
  PeerTables pt{};
  pt.n = n_peers; pt.row_off = row_off; pt.mc = multicast_table;
  for (int r = 0; r < n_peers; ++r) pt.tables[r] = peer_tables[r];
  return tc_gemm_fwd(x, ldx, wcat_t, nullptr, h_i, n, fin, fp, relu_lh, &pt, reinterpret_cast<cudaStream_t>(stream));
}
