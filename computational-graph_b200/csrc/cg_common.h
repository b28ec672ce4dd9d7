// cg_common.h — shared declarations for libcg_b200.so (host + device).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace cg {

// Host-side failure = the reference's `panic!` / `check()`; the C ABI turns it into an error code.
struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s (%s) at %s:%d", cudaGetErrorString(e), what, file, line);
    throw Error(buf);
  }
}
#define CG_CUDA(x) ::cg::cuda_check((x), #x, __FILE__, __LINE__)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ------------------------------------------------------------------------------------------------
// Fused-elementwise bytecode (executed by ew_program_kernel, kernels_ew.cu).
//
// An accumulator machine over per-thread vectors of EW_VEC floats: one accumulator `acc` and EW_REGS
// vector registers r[0..EW_REGS). Inputs are prefetched into r[0..n_in) before the first instruction
// (all loads in flight at once); every instruction is (opcode << 4) | operand, operand < 16.
// Every arithmetic opcode is a single IEEE f32 operation (no FMA contraction), so a fused chain is
// bit-identical to the reference's one-kernel-per-node evaluation (src/cpu/ops.cpp:68-86).
// ------------------------------------------------------------------------------------------------
constexpr int EW_REGS = 12;
constexpr int EW_MAX_IN = 12;
constexpr int EW_MAX_OUT = 12;
constexpr int EW_MAX_INS = 160;
constexpr int EW_MAX_IMM = 8;

enum EwOpcode : uint8_t {
  OP_END = 0,
  OP_LDR,      // acc = r[k]
  OP_STR,      // r[k] = acc
  OP_LDI,      // acc = imm[k]
  OP_ADD,      // acc = acc + r[k]
  OP_SUB,      // acc = acc - r[k]
  OP_RSUB,     // acc = r[k] - acc
  OP_MUL,      // acc = acc * r[k]
  OP_MULI,     // acc = acc * imm[k]
  OP_NEG,      // acc = -acc
  OP_SQ,       // acc = acc * acc
  OP_ABS_M,    // matrix path: x < 0 ? -x : x            (cpu/ops.cpp:70; abs(-0) = -0)
  OP_SGN_M,    // matrix path: x < 0 ? -1 : 1            (cpu/ops.cpp:71; signum(+-0) = +1)
  OP_ABS_S,    // scalar path: f32::abs                  (ops.rs:331; abs(-0) = +0)
  OP_SGN_S,    // scalar path: f32::signum               (ops.rs:332; signum(-0) = -1, NaN -> NaN)
  OP_SIGMOID,  // 1.0f / (1.0f + expf(-x))               (cpu/ops.cpp:72)
  OP_TANH,     // tanhf(x)                               (cpu/ops.cpp:73)
  OP_ONE_MINUS,// 1.0f - x                               (cpu/ops.cpp:74)
  OP_STG,      // out[k][i] = acc
  OP_COUNT
};

struct EwProgram {
  uint64_t n;                       // elements in the iteration domain
  uint32_t n_in, n_out, n_ins;
  uint32_t in_scalar_mask;          // bit k: input k is a 1-element buffer broadcast to every element
  const float* in[EW_MAX_IN];
  float* out[EW_MAX_OUT];
  float imm[EW_MAX_IMM];
  uint16_t code[EW_MAX_INS];        // (opcode << 4) | operand
};

inline uint16_t ew_encode(EwOpcode op, int k = 0) { return (uint16_t)((op << 4) | (k & 0xf)); }

// ------------------------------------------------------------------------------------------------
// Kernel launchers (all asynchronous on `stream`). Column-major f32 everywhere.
// ------------------------------------------------------------------------------------------------
void launch_ew_program(const EwProgram& p, cudaStream_t stream);

// Plan-time specialisation of an elementwise program into straight-line sm_100a code (ew_jit.cpp, NVRTC).
// ew_jit_get returns nullptr when libnvrtc is not available (the interpreter kernel above is used instead).
struct EwKernel;
bool ew_jit_available();
const EwKernel* ew_jit_get(const EwProgram& p);
void launch_ew_jit(const EwKernel* k, const EwProgram& p, cudaStream_t stream);
std::string ew_jit_source(const EwProgram& p);
size_t ew_jit_compile_only(const EwProgram& p);  // compiles without loading (no GPU needed); returns cubin bytes
size_t ew_jit_cache_size();

// dst (1 element) = reduce_sum(src[0..n)) / n   — the reference's `avg` (ops.rs:438-442, cpu/ops.cpp:138-145).
// n <= kStrictReduceMax: one thread, sequential f32 sum (bit-identical to the reference's loop);
// larger: float4 loads -> warp shuffle -> block -> ordered final pass (deterministic, pairwise rounding).
constexpr size_t kStrictReduceMax = 4096;
void launch_reduce_avg(const float* src, size_t n, float* dst, float* workspace, cudaStream_t stream);
// dst = reduce_sum(src) / denom (denom = 1 gives the plain reduce_sum of cpu/ops.cpp:138-145)
void launch_reduce_div(const float* src, size_t n, float denom, float* dst, float* workspace, cudaStream_t stream);
size_t reduce_workspace_floats();

// flag[0] = all(src == x) / all(src < x)  (cpu/ops.cpp:118-136); flag is a device int
void launch_all_equal(const float* src, size_t n, float x, int* flag, cudaStream_t stream);
void launch_all_less_than(const float* src, size_t n, float x, int* flag, cudaStream_t stream);

// dst(column-major h x w) <- src(row-major h x w)   (set_array(transpose=true), cpu/matrix.cpp:26-36)
void launch_rowmajor_to_colmajor(const float* src_rowmajor, float* dst_colmajor, unsigned h, unsigned w,
                                 cudaStream_t stream);

// C(m x n) = op(A) . op(B), column-major; lda/ldb/ldc = stored heights.
//   strict = true : every output element is accumulated k-ascending with separate f32 multiply and add
//                   roundings — bit-identical to the reference's naive gemm (cpu/ops.cpp:103-114).
//   strict = false: tiled FP32 SIMT with FMA accumulation.
struct GemmEpilogue {
  // mode 0: C = acc.
  // mode 1 (fused SGD, optimize.rs:149-152): C = C - (acc * lr), two roundings (never an FMA).
  // mode 2 (dag-mode error accumulation): C = C + acc.
  int mode = 0;
  float lr = 0.0f;
};
void launch_gemm_f32(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb,
                     int ldc, bool strict, const GemmEpilogue& epi, cudaStream_t stream);

// TF32 tensor-core path (tcgen05 + TMA + TMEM, gemm_tcgen05.cu). Returns false when the shape is not
// eligible (caller falls back to launch_gemm_f32).
bool gemm_tf32_eligible(int m, int n, int k, bool ta, bool tb);
bool gemm_tf32_aligned(const float* A, const float* B, int lda, int ldb);
void gemm_tf32_set_max_ctas(int n);  // persistent grid of the TF32 product (default: all 148 SMs)
void gemm_tf32_set_pair(int on);  // 1: CTA-pair (cta_group::2) kernel for shapes with m % 256 == 0 and n % 64 == 0
void launch_gemm_tf32(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda,
                      int ldb, int ldc, const GemmEpilogue& epi, cudaStream_t stream);
// FP32-accurate product on the tensor core: every operand split in shared memory into the part TF32 keeps and the part
// it drops, three MMAs per k-step (CTA-pair kernel; m % 256 == 0).
bool gemm_tf32x3_eligible(int m, int n, int k, bool ta, bool tb);
void launch_gemm_tf32x3(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda,
                        int ldb, int ldc, const GemmEpilogue& epi, cudaStream_t stream);

// Shape-based dispatch: strict (max(m,n,k) <= Runtime::strict_gemm_max) | tcgen05 TF32 (precision 1 and eligible) |
// tcgen05 3 x TF32 (precision 2 and eligible) | tiled FP32 SIMT.
void launch_gemm(const float* A, bool ta, const float* B, bool tb, float* C, int m, int n, int k, int lda, int ldb,
                 int ldc, const GemmEpilogue& epi, int precision, cudaStream_t stream);

// Data-parallel sum-allreduce of a weight gradient over NCCL (dp.cpp). In place.
void dp_allreduce_sum(float* buf, size_t n, cudaStream_t stream);
bool dp_enabled();

// Fused all-reduce + SGD update over NVLink peer memory (dp_kernels.cu). Buffers are "windows": allocations every
// rank has opened through CUDA IPC, so rank r can address the same offset in every peer's copy.
constexpr int DP_MAX_WORLD = 8;
constexpr int DP_THREADS = 512;
constexpr int DP_DEFAULT_CTAS = 64;  // CG_DP_CTAS overrides
int dp_max_ctas();
constexpr int DP_MC_DEFAULT_CTAS = 32;  // in-switch exchange; CG_DP_MC_CTAS overrides
int dp_mc_max_ctas();
constexpr unsigned long long DP_WAIT_TIMEOUT_NS = 20000000000ull;  // default 20 s (CG_DP_TIMEOUT_MS): a lost rank raises an error, never hangs the GPU
enum { DP_FLAG_A = 0, DP_FLAG_B = 16, DP_FLAG_EPOCH = 32, DP_FLAG_TICKET = 33, DP_FLAG_ERR = 34, DP_FLAG_WORDS = 64 };
struct DpReduceArgs {
  const float* g[DP_MAX_WORLD];  // the partial gradient in every rank (index = rank)
  float* w[DP_MAX_WORLD];        // the weight in every rank
  uint32_t* flags[DP_MAX_WORLD]; // every rank's flag block (DP_FLAG_WORDS words)
  float lr;
  size_t n;
  int rank, world;
  unsigned long long timeout_ns;  // bound of every flag wait
};
void launch_dp_reduce_update(const DpReduceArgs& a, cudaStream_t stream);
// The in-switch (NVLS multimem) form of the same exchange: buffers in the symmetric multicast heap (symm.cpp).
struct DpMcArgs {
  const float* g_mc;   // the gradient slot through the multicast object: a load sums it over every rank in the switch
  const float* w_uc;   // this rank's weight, ordinary (unicast) address
  float* w_mc;         // the weight through the multicast object: a store lands in every rank's copy
  uint32_t* flags_uc;  // this rank's counter block (DP_FLAG_WORDS words, in the heap)
  uint32_t* flags_mc;  // the same block through the multicast object
  float lr;
  size_t n;
  int rank, world;
  unsigned long long timeout_ns;
};
void launch_dp_multimem_update(const DpMcArgs& a, cudaStream_t stream);
// Symmetric multicast heap (symm.cpp). symm_init / symm_alloc are collective over the data-parallel ranks.
bool symm_init(unsigned long long job_id, int device_ordinal);
void symm_shutdown();
bool symm_available();
const char* symm_unavailable_reason();
void* symm_alloc(size_t bytes);
void symm_free(void* uc);
bool symm_locate(const void* uc, unsigned long long* seg, unsigned long long* off, void** mc);
size_t symm_bytes_mapped();
// Whether plans built now use the in-switch exchange (heap available and selected for this world size).
bool dp_multimem_active();
void dp_fill_mc_flags(DpMcArgs& a);
std::vector<unsigned char> dp_allgather_bytes(const void* mine, size_t bytes);  // control messages over the communicator
// Collective (every rank, same order): makes the cudaMalloc allocations that contain these pointers addressable from
// every rank. Returns false when peer access is not available (plans then keep ncclAllReduce + a separate update).
bool dp_register_windows(const std::vector<const void*>& ptrs);
// A device buffer is about to be freed: drops its registrations; true when dp.cpp has taken over the cudaFree (a peer
// may still have the allocation open — it is freed at the next collective point).
bool dp_retire(void* ptr, size_t bytes);
size_t dp_fused_min_elems();  // weights smaller than this stay on ncclAllReduce (CG_DP_FUSED_MIN)
// Fills the peer pointers for `local` (inside a registered window); false if it is not in one.
bool dp_peer_pointers(const void* local, void* peers[DP_MAX_WORLD]);
bool dp_fused_available();
void dp_fill_flags(DpReduceArgs& a);
void dp_check_errors();  // throws if a fused exchange timed out waiting for a peer

}  // namespace cg
