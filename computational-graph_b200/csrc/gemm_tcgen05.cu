// gemm_tcgen05.cu — TF32 tensor-core product (tcgen05 + TMA + TMEM). Placeholder until the kernel lands:
// nothing is eligible, so every product takes the FP32 SIMT path.
#include "cg_common.h"

namespace cg {
bool gemm_tf32_eligible(int, int, int, bool, bool) { return false; }
void launch_gemm_tf32(const float*, bool, const float*, bool, float*, int, int, int, int, int, int, const GemmEpilogue&,
                      cudaStream_t) {
  throw Error("tcgen05 TF32 product is not built");
}
}  // namespace cg
