// graph.h — host-side graph types mirroring the reference's Rust host (src/function.rs, src/constant.rs).
//
//   Value     <-> `Constant` (constant.rs:5-8) held on the device: a matrix is column-major f32 in HBM,
//                 a scalar is a 1-element device buffer (so whole iterations are CUDA-graph capturable).
//   Function  <-> `struct Function {value, params, body: Rc<Expr>, pool}` (function.rs:10-16); `clone`
//                 deep-copies the value and shares the body (Q1: every use site of a param owns a copy).
//   Expr      <-> `enum Expr` (function.rs:18-33); children are owned by value.
// The per-node matrix object pool (function.rs:7-8,58-76) is not a member: placeholders are sized and
// placed by the tape compiler (plan.cpp) inside one device arena, with zero allocations in the loop.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "cg_common.h"

namespace cg {

struct DevBuf {
  float* ptr = nullptr;
  size_t n = 0;
  std::shared_ptr<void> keepalive;  // arena that owns `ptr`, when not owned here
  bool owned = false;
  bool symm = false;         // lives in the symmetric multicast heap (symm.cpp): data-parallel weights and gradient slots
  bool placeholder = false;  // plan-only mode: an address that is never dereferenced
  // plan-only mode: the INITIAL contents (storage order) this buffer would have been given — never computed on, only
  // written out next to the generated resident-kernel source (CG_TAPE_JIT_DUMP) for the CPU-side emulator test
  std::shared_ptr<std::vector<float>> shadow;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf();
  static std::shared_ptr<DevBuf> alloc(size_t n);
  static std::shared_ptr<DevBuf> alloc_symm(size_t n);  // collective over the data-parallel ranks
};
using BufRef = std::shared_ptr<DevBuf>;

struct Value {
  bool is_matrix = false;
  unsigned h = 0, w = 0;
  // Oracle rule: values of Dot/Input nodes are uninitialised memory in the reference;
  // they are "undefined" and never match an identity in the construction-time elimination (Q4).
  bool defined = true;
  // Host-side summary used by the value-based identity elimination (ops.rs:45-53,255-263): valid until
  // the device overwrites the buffer (minimize), after which all_equal runs a device reduction.
  bool summary_valid = true;
  bool uniform = false;
  float uval = 0.0f;
  BufRef buf;  // device storage: always present for leaves, attached by the plan for evaluated nodes
  size_t size() const { return is_matrix ? (size_t)h * w : 1; }
};

enum class Kind { Constant, Input, Param, Neg, Sq, Abs, Signum, Sigmoid, Tanh, Add, Sub, Mul, Dot };

struct Expr;
struct Plan;

struct Function {
  Value value;
  bool has_params = false;  // !params.is_empty() (function.rs:12; Input names count, function.rs:107)
  std::shared_ptr<Expr> body;
  std::shared_ptr<Plan> plan;  // compiled tape, cached on the root the user optimises
  std::shared_ptr<Plan> eval_plan;  // forward-only tape of `eval` (optimize.rs:8-32)
  std::shared_ptr<Function> negated;  // `-self`, built by the first `maximize` (optimize.rs:74-81) and reused
  bool negated_identity = false;      // whether that construction hit the identity elimination (`-0 == 0`)
  int plan_mode = -1;
};

struct Expr {
  Kind kind;
  std::string name;
  std::unique_ptr<Function> f1, f2;
  bool t1 = false, t2 = false;
};

inline bool is_leaf(Kind k) { return k == Kind::Constant || k == Kind::Input || k == Kind::Param; }

const std::vector<float>* plan_only_shadow(const float* placeholder);  // plan-only: initial contents of a buffer, or null
bool plan_only();
void set_plan_only(bool on);

// ---- global runtime state -----------------------------------------------------------------------
struct Runtime {
  cudaStream_t stream = nullptr;
  int mode = 0;          // 0 = exact (the reference's tree walk), 1 = dag (memoised)
  int tf32 = 0;          // 0 = FP32 SIMT products, 1 = tcgen05 TF32 where eligible, 2 = FP32-accurate 3 x TF32 on the tensor core
  int fix_act_grad = 0;  // 0 = reproduce Q3
  int fix_broadcast_sum = 0;  // 0 = reproduce Q5 in the backward pass (scalar <- matrix error = mean); 1: sum
  int fix_sub_sign = 0;       // 0 = reproduce Q4's `0 - x -> x`; 1: `0 - x -> -x`
  int fix_tied_params = 0;    // 0 = reproduce Q1 (every use site of a param drifts on its own); 1: one parameter per name
  int fuse = 1;          // 0 = one kernel per IR instruction (debug / ablation)
  int use_graph = 1;     // capture each iteration as a CUDA graph
  int graph_streams = 4; // side streams used when an iteration is captured as a DAG (0 = single stream)
  long long ew_jit_min = 32768;  // elementwise programs over >= this many elements are specialised (ew_jit.cpp); < 0: never
  int hoist = 1;         // dag mode: move loss-independent backward work next to the forward ops it depends on (plan.cpp)
  int vm_max_dim = 48;   // plans whose every matrix dimension is <= this run on the resident tape executor (vm_kernel.cu); 0: never
  int elide_root = 1;    // iterations whose root value is unobservable (Q12) skip its store (elementwise-only graphs)
  int tape_jit = 1;      // resident plans are compiled into one shared-memory-resident kernel when they fit (ew_jit.cpp)
  int io_graph = 1;      // put the H2D / D2H transfers of a `minimize(&args, lr, 1)` call inside the iteration graph
  std::vector<cudaStream_t> side_streams;
  int dp_epoch = 0;      // bumped whenever the data-parallel configuration changes (invalidates plans)
  int strict_gemm_max = 64;  // products with max(m,n,k) <= this use the bit-exact strict kernel
  unsigned long long launches = 0;  // kernels launched by this library (bench.py's gpu_launches)
  int* flag = nullptr;   // device int for all_equal / all_less_than
  float* red_ws = nullptr;
  BufRef staging;        // grow-only device staging buffer for row-major uploads (never freed in a loop)
  static Runtime& get();
  static bool get_if_initialised();
  void init();
};

// ---- constructors (function.rs:100-155, ops.rs:36-54,199-264,395-407) --------------------------------
Function* make_scalar_leaf(Kind kind, const char* name, float v);
Function* make_matrix_leaf(Kind kind, const char* name, unsigned h, unsigned w, const float* rowmajor_vals);
Function* make_fill_leaf(Kind kind, const char* name, unsigned h, unsigned w, float v);
Function* make_input(const char* name, bool is_matrix, unsigned h, unsigned w);
Function* clone_function(const Function& f);
Function* make_unary(Kind kind, const Function& f);
Function* make_binary(Kind kind, const Function& a, const Function& b);
Function* make_dot(const Function& a, bool t1, const Function& b, bool t2);

bool value_all_equal(const Value& v, float x);
bool value_all_less_than(const Value& v, float x);
void value_download(const Value& v, float* dst);  // storage (column-major) order, as get_array
void value_upload_rowmajor(Value& v, const float* src_rowmajor, bool sync = true);

struct Arg {
  std::string name;
  const float* vals;  // row-major host floats (matrix) or one float (scalar)
  unsigned h, w;      // 0,0 for a scalar
};

// optimize.rs:59-81
// reuse_inputs: keep the Input values uploaded by an earlier call (device-resident replays).
void minimize(Function& f, const std::vector<Arg>& args, float lr, int iters, int print_freq, float* trace,
              bool synchronize, bool reuse_inputs = false);
std::vector<Function*> collect_param_leaves(Function& root);
// optimize.rs:8-56
Function* function_grad(const Function& f, const std::string& param);        // symbolic df/dparam
void function_eval(Function& f, const std::vector<Arg>& args, float* dst);   // forward value, storage order; no cached state changes

}  // namespace cg
