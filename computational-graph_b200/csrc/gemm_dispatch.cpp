// gemm_dispatch.cpp — picks the product kernel by shape (north_star item 2: tcgen05 TF32 only where the
// shape is a true dense GEMM, FP32 SIMT otherwise; tiny products use the bit-exact strict kernel).
// precision (cg_set_tf32): 0 = FP32 SIMT, 1 = TF32 tensor core, 2 = FP32-accurate on the tensor core (3 x TF32 split).
#include "graph.h"

namespace cg {

void launch_gemm(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb,
                 int ldc, const GemmEpilogue& epi, int precision, cudaStream_t stream) {
  Runtime& rt = Runtime::get();
  const int mx = m > n ? (m > k ? m : k) : (n > k ? n : k);
  if (mx <= rt.strict_gemm_max) {
    launch_gemm_f32(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, /*strict=*/true, epi, stream);
  } else if (precision == 1 && gemm_tf32_eligible(m, n, k, ta, tb) && gemm_tf32_aligned(A, B, lda, ldb)) {
    launch_gemm_tf32(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
  } else if (precision == 2 && gemm_tf32x3_eligible(m, n, k, ta, tb) && gemm_tf32_aligned(A, B, lda, ldb)) {
    launch_gemm_tf32x3(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, epi, stream);
  } else {
    launch_gemm_f32(A, ta, B, tb, C, m, n, k, lda, ldb, ldc, /*strict=*/false, epi, stream);
  }
}

}  // namespace cg
