// gemm_dispatch.cpp — picks the product kernel by shape (north_star item 2: tcgen05 TF32 only where the
// shape is a true dense GEMM, FP32 SIMT otherwise; tiny products use the bit-exact strict kernel).
#include "graph.h"

namespace cg {

void launch_gemm(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb,
                 int ldc, const GemmEpilogue& epi, bool allow_tf32, cudaStream_t stream) {
  Runtime& rt = Runtime::get();
  const int mx = m > n ? (m > k ? m : k) : (n > k ? n : k);
  if (mx <= rt.strict_gemm_max) {
    launch_gemm_f32(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, /*strict=*/true, epi, stream);
  } else if (allow_tf32 && gemm_tf32_eligible(m, n, k, ta, tb) && gemm_tf32_aligned(A, B, lda, ldb)) {
    launch_gemm_tf32(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
  } else {
    launch_gemm_f32(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, /*strict=*/false, epi, stream);
  }
}

}  // namespace cg
