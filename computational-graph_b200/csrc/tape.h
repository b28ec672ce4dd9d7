// tape.h — the linear IR ("tape") one `minimize` iteration compiles to, and the compiled plan.
//
// The reference walks the Function tree recursively every iteration, issuing one FFI call (= one
// eager kernel) per node (src/optimize.rs:84-204). Here the walk happens ONCE, at minimize() entry:
// it emits SSA micro-ops over immutable values; the tape is then dead-code-eliminated, the dW product
// is fused with its SGD update, runs of same-shaped elementwise micro-ops are fused into register
// bytecode programs (one HBM pass per sub-tree), every surviving value gets a slot in one device
// arena sized from the tape (zero allocations inside the loop), and the iteration is captured as a
// CUDA graph.
#pragma once
#include <unordered_map>
#include <vector>

#include "graph.h"

namespace cg {

struct Shape {
  bool m = false;
  unsigned h = 0, w = 0;
  size_t n() const { return m ? (size_t)h * w : 1; }
  bool same(const Shape& o) const { return m == o.m && (!m || (h == o.h && w == o.w)); }
};

enum class IOp : uint8_t {
  // elementwise micro-ops (fusable)
  Neg, Sq, AbsM, SgnM, AbsS, SgnS, Sigmoid, Tanh, OneMinus, Add, Sub, Mul, MulI, Fill, Copy,
  // kernel-boundary ops
  Avg,        // dst(scalar) = reduce_sum(a) / size(a)            (ops.rs:438-442)
  Gemm,       // dst = op(a) . op(b) [epilogue]                   (ops.rs:411-422)
  AllReduce,  // dst = sum over data-parallel ranks of a (in place)
  DpUpdate,   // dst = c_in - (sum over ranks of a) * imm, written into every rank's copy: all-reduce + SGD update
              // of a replicated weight as one kernel over NVLink peer memory (dp_kernels.cu)
};
inline bool is_ew(IOp op) { return op <= IOp::Copy; }

struct Val {
  Shape sh;
  float* fixed = nullptr;  // storage owned elsewhere (leaf / constant / input / root value buffers)
  BufRef keep;
  bool persistent = false;  // must be in memory after the iteration (leaf versions, root value)
  // planning
  int nuses = 0;
  int def_step = -1, last_step = -1;
  bool needs_mem = false;
  int alias_of = -1;  // shares memory with this value (in-place update of a temp)
  float* ptr = nullptr;
};

struct Instr {
  IOp op;
  int dst = -1, a = -1, b = -1;
  float imm = 0.0f;
  bool ta = false, tb = false;  // Gemm
  int epi = 0;                  // Gemm epilogue: 0 store, 1 fused SGD (c_in - acc*imm), 2 accumulate (c_in + acc)
  int c_in = -1;
  bool dead = false;
};

struct EwGroup {
  size_t n = 0;
  std::vector<int> in, out;  // value ids
  std::vector<uint16_t> code;
  std::vector<float> imm;
};

struct Step {
  enum Kind { EW, AVG, GEMM, ALLREDUCE, DPUPDATE } kind;
  EwGroup ew;
  Instr ins;  // AVG / GEMM / ALLREDUCE
  DpReduceArgs dp;  // DPUPDATE, filled after memory assignment (peer pointers)
  DpMcArgs mc;      // DPUPDATE through the NVSwitch (use_mc): multicast addresses in the symmetric heap
  bool use_mc = false;
  EwProgram prog;  // EW, filled after memory assignment
  const EwKernel* jit = nullptr;  // EW: the program specialised into straight-line code (large domains only)
};

// One step of the resident tape executor (vm_kernel.cu), in device memory.
struct VmStep {
  enum { EW = 0, AVG = 1, GEMM = 2 };
  int kind;
  int prog;                      // EW: index into the program array
  int m, n, k, lda, ldb, ldc;    // GEMM: C(m x n) = op(A) . op(B); AVG: n = element count
  int ta, tb, epi;
  float lr;
  const float* a;                // GEMM: A; AVG: source
  const float* b;
  float* c;                      // GEMM: C; AVG: the 1-element destination
};
void launch_tape_vm(const VmStep* steps, const EwProgram* progs, int nsteps, int iters, const float* root, size_t root_n,
                    bool root_is_leaf, float* trace, cudaStream_t stream);

// The whole plan compiled into one resident kernel with every value in shared memory (ew_jit.cpp, NVRTC); nullptr when
// the plan does not fit in one SM's shared memory or libnvrtc is absent. load = false compiles without a device.
struct TapeKernel;
struct Plan;
const TapeKernel* tape_jit_build(const Plan& p, bool load);
void tape_jit_free(const TapeKernel* k);
size_t tape_jit_smem_bytes(const TapeKernel* k);
size_t tape_jit_cache_size();  // distinct compiled modules (structurally identical plans share one)
void launch_tape_jit(const TapeKernel* k, int iters, float* trace, cudaStream_t stream);

struct PlanStats {
  size_t tape_instrs = 0, live_instrs = 0, ew_groups = 0, gemms = 0, gemms_fused_sgd = 0, reduces = 0;
  size_t arena_bytes = 0, kernels_per_iter = 0, ew_specialised = 0, allreduces = 0, dp_fused_updates = 0, vm_resident = 0, tape_smem_bytes = 0, quiet_variant = 0;
  size_t dp_multimem_updates = 0;  // of dp_fused_updates: done inside the NVSwitch (multimem)
  double gemm_flop_per_iter = 0.0, ew_bytes_per_iter = 0.0, quiet_ew_bytes_per_iter = 0.0;
};

// The iteration graph with the host transfers of one `minimize(&args, lr, 1)` call inside it: H2D copies of the
// Input values (pinned host memory -> staging), their layout kernels, the iteration, and the D2H read of the root
// value — so the copies overlap the compute instead of preceding / following it (exec.cpp).
struct IoGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  std::vector<cudaGraphNode_t> h2d_nodes;  // one per input leaf, Plan::input_leaves order
  std::vector<const float*> h2d_src;
  std::vector<float*> h2d_dst;
  std::vector<size_t> h2d_bytes;
  cudaGraphNode_t d2h_node = nullptr;
  float* d2h_dst = nullptr;
  size_t layout_kernels = 0;
  // External-copy form (default): the H2D copies are enqueued on a copy stream right before the graph launch and the
  // graph holds one event-wait node per input instead of a memcpy node (exec.cpp launch_io_graph).
  std::vector<size_t> order;  // input indices, most urgent first (longest chain of work behind their first consumer)
  bool external_h2d = false;
  std::vector<cudaEvent_t> in_ev;  // one per input leaf: "this call's copy of input i has landed"
};

struct Plan {
  std::vector<Val> vals;
  std::vector<Instr> tape;
  std::vector<Step> steps;
  BufRef arena;
  BufRef dp_arena;   // gradient slots of the in-switch exchange: symmetric multicast heap (symm.cpp)
  BufRef eval_root;  // forward-only plans (`eval`): the root value, in a buffer of the plan's own
  PlanStats stats;
  float lr = 0.0f;
  int mode = 0, tf32 = 0, fix_act_grad = 0, fuse = 1, dp_epoch = 0, hoist = 0;
  long long ew_jit_min = -1;
  bool root_is_leaf = false;
  float* root_ptr = nullptr;
  size_t root_n = 1;
  std::vector<Function*> input_leaves;
  std::vector<Function*> param_leaves;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  std::vector<cudaEvent_t> events;  // one per captured node (+1 fork event) for the multi-stream capture
  IoGraph io[2];                    // [0] without, [1] with the D2H read of the root value
  std::vector<BufRef> io_staging;   // row-major landing buffer per matrix input
  cudaEvent_t io_done = nullptr;    // recorded behind every transfer-graph launch: the staging buffers are free again
  // resident tape executor (vm_kernel.cu): the steps and programs in device memory, when the plan is small enough
  // The same iteration without the store of the root value: only the last iteration's value is observable (Q12), and
  // for an elementwise-only graph (C2) that store is 4 of the 28 bytes per element the fused kernel moves.
  std::shared_ptr<Plan> quiet;
  const TapeKernel* tape_kernel = nullptr;  // the compiled form (shared-memory resident), preferred when present
  VmStep* vm_steps = nullptr;
  EwProgram* vm_progs = nullptr;
  int vm_nsteps = 0;
  ~Plan();
  void release_graphs();  // destroy the captured CUDA graphs (they are re-captured on the next use)
};

std::shared_ptr<Plan> build_plan(Function& root, float lr, bool keep_root = true, bool forward_only = false);
// Destroys the CUDA graphs of every live plan. NCCL keeps a communicator alive while a captured graph still holds one
// of its collectives (ncclCommDestroy would wait forever), so cg_dp_shutdown calls this first.
void release_all_plan_graphs();
void plan_launch_iteration(Plan& p, cudaStream_t stream);  // enqueue the kernels of one iteration
void plan_launch_step(Plan& p, size_t index, cudaStream_t stream);
double plan_step_flop(const Plan& p, size_t index);
std::string plan_describe(const Plan& p);  // one line per kernel of the iteration, in launch order

}  // namespace cg
