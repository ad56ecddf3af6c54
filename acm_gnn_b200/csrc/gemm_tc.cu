// tcgen05 (5th-gen tensor core) bf16 GEMM path -- placeholder until the kernel lands.
#include "acm_common.cuh"

namespace acm {
int tc_gemm_fwd(const void*, int64_t, const void*, void*, void*, int64_t, int64_t, int64_t, int, cudaStream_t) {
  set_error("tcgen05 GEMM path not built yet");
  return ACM_ERR_UNSUPPORTED;
}
int tc_gemm_dw(const void*, int64_t, const void*, float*, int64_t, int64_t, int64_t, cudaStream_t) {
  set_error("tcgen05 GEMM path not built yet");
  return ACM_ERR_UNSUPPORTED;
}
int tc_gemm_dx(const void*, const void*, float*, int64_t, int64_t, int64_t, int64_t, cudaStream_t) {
  set_error("tcgen05 GEMM path not built yet");
  return ACM_ERR_UNSUPPORTED;
}
}  // namespace acm
