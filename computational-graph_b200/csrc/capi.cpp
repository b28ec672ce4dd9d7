// capi.cpp — the C ABI of libcg_b200.so: include/cg_graph.h (graph level) and include/cg_matrix.h (the
// reference's 26 matrix symbols). No torch types, no C++ types in any signature.
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cg_graph.h"
#include "../../include/cg_matrix.h"
#include "tape.h"

using namespace cg;

static thread_local std::string g_last_error;

static Function* F(cg_function* h) { return reinterpret_cast<Function*>(h); }
static const Function* F(const cg_function* h) { return reinterpret_cast<const Function*>(h); }
static cg_function* H(Function* f) { return reinterpret_cast<cg_function*>(f); }

template <typename Fn>
static cg_function* guard_new(Fn&& fn) {
  try {
    g_last_error.clear();
    return H(fn());
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return nullptr;
  }
}
template <typename Fn>
static int guard_rc(Fn&& fn) {
  try {
    g_last_error.clear();
    fn();
    return 0;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}

extern "C" {

const char* cg_last_error(void) { return g_last_error.c_str(); }
int cg_dp_set_error(const char* msg) {
  g_last_error = msg;
  return -1;
}

// ---- leaves ---------------------------------------------------------------------------------------------
cg_function* cg_scalar_param(const char* name, float v) { return guard_new([&] { return make_scalar_leaf(Kind::Param, name, v); }); }
cg_function* cg_matrix_param(const char* name, unsigned h, unsigned w, const float* vals) {
  return guard_new([&] { return make_matrix_leaf(Kind::Param, name, h, w, vals); });
}
cg_function* cg_single_val_param(const char* name, unsigned h, unsigned w, float v) {
  return guard_new([&] { return make_fill_leaf(Kind::Param, name, h, w, v); });
}
cg_function* cg_scalar(float v) { return guard_new([&] { return make_scalar_leaf(Kind::Constant, nullptr, v); }); }
cg_function* cg_matrix(unsigned h, unsigned w, const float* vals) {
  return guard_new([&] { return make_matrix_leaf(Kind::Constant, nullptr, h, w, vals); });
}
cg_function* cg_single_val_matrix(unsigned h, unsigned w, float v) {
  return guard_new([&] { return make_fill_leaf(Kind::Constant, nullptr, h, w, v); });
}
cg_function* cg_input(const char* name, int is_matrix, unsigned h, unsigned w) {
  return guard_new([&] { return make_input(name, is_matrix != 0, h, w); });
}
cg_function* cg_clone(const cg_function* f) { return guard_new([&] { return clone_function(*F(f)); }); }
void cg_release(cg_function* f) { delete F(f); }

// ---- ops --------------------------------------------------------------------------------------------------
cg_function* cg_neg(const cg_function* f) { return guard_new([&] { return make_unary(Kind::Neg, *F(f)); }); }
cg_function* cg_sq(const cg_function* f) { return guard_new([&] { return make_unary(Kind::Sq, *F(f)); }); }
cg_function* cg_abs(const cg_function* f) { return guard_new([&] { return make_unary(Kind::Abs, *F(f)); }); }
cg_function* cg_signum(const cg_function* f) { return guard_new([&] { return make_unary(Kind::Signum, *F(f)); }); }
cg_function* cg_sigmoid(const cg_function* f) { return guard_new([&] { return make_unary(Kind::Sigmoid, *F(f)); }); }
cg_function* cg_tanh(const cg_function* f) { return guard_new([&] { return make_unary(Kind::Tanh, *F(f)); }); }
cg_function* cg_add(const cg_function* a, const cg_function* b) { return guard_new([&] { return make_binary(Kind::Add, *F(a), *F(b)); }); }
cg_function* cg_sub(const cg_function* a, const cg_function* b) { return guard_new([&] { return make_binary(Kind::Sub, *F(a), *F(b)); }); }
cg_function* cg_mul(const cg_function* a, const cg_function* b) { return guard_new([&] { return make_binary(Kind::Mul, *F(a), *F(b)); }); }
cg_function* cg_dot(const cg_function* a, int t1, const cg_function* b, int t2) {
  return guard_new([&] { return make_dot(*F(a), t1 != 0, *F(b), t2 != 0); });
}

// ---- optimisation ---------------------------------------------------------------------------------------------
static std::vector<Arg> pack_args(int nargs, const char** names, const float** vals, const unsigned* hs, const unsigned* ws) {
  std::vector<Arg> args;
  for (int i = 0; i < nargs; i++) args.push_back(Arg{names[i], vals[i], hs[i], ws[i]});
  return args;
}
int cg_minimize(cg_function* f, int nargs, const char** names, const float** arg_vals, const unsigned* arg_h,
                const unsigned* arg_w, float lr, int iters, int print_freq, float* trace) {
  return guard_rc([&] { minimize(*F(f), pack_args(nargs, names, arg_vals, arg_h, arg_w), lr, iters, print_freq, trace, true); });
}
int cg_maximize(cg_function* f, int nargs, const char** names, const float** arg_vals, const unsigned* arg_h,
                const unsigned* arg_w, float lr, int iters, int print_freq, float* trace) {
  // optimize.rs:74-81: (-self).minimize(...). The negated root shares f's body (and with it every leaf), so it is built
  // once and kept on f: repeated calls reuse its compiled plan instead of re-planning (and re-compiling) each time.
  return guard_rc([&] {
    Function& fn = *F(f);
    // (the constructor's value-based identity elimination, ops.rs:45-53, is re-decided per call as in the reference:
    // `-f` IS `f` while f's cached value is all zeros)
    const bool identity = value_all_equal(fn.value, 0.0f);
    if (!fn.negated || fn.negated_identity != identity) {
      fn.negated.reset(make_unary(Kind::Neg, fn));
      fn.negated_identity = identity;
    }
    minimize(*fn.negated, pack_args(nargs, names, arg_vals, arg_h, arg_w), lr, iters, print_freq, trace, true);
  });
}
// optimize.rs:8-56: forward evaluation and the symbolic ("naive") gradient
int cg_eval(cg_function* f, int nargs, const char** names, const float** arg_vals, const unsigned* arg_h, const unsigned* arg_w,
            float* dst) {
  return guard_rc([&] { function_eval(*F(f), pack_args(nargs, names, arg_vals, arg_h, arg_w), dst); });
}
cg_function* cg_grad(const cg_function* f, const char* param) {
  cg_function* out = nullptr;
  guard_rc([&] { out = reinterpret_cast<cg_function*>(function_grad(*F(f), param ? param : "")); });
  return out;
}
void* cg_host_alloc(size_t bytes) {
  void* ptr = nullptr;
  guard_rc([&] {
    Runtime::get();
    CG_CUDA(cudaHostAlloc(&ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  });
  return ptr;
}
void cg_host_free(void* ptr) {
  if (ptr) cudaFreeHost(ptr);
}
int cg_minimize_async(cg_function* f, float lr, int iters) {
  return guard_rc([&] { minimize(*F(f), {}, lr, iters, 0, nullptr, false, /*reuse_inputs=*/true); });
}
int cg_synchronize(void) {
  return guard_rc([&] {
    CG_CUDA(cudaStreamSynchronize(Runtime::get().stream));
    if (dp_fused_available()) dp_check_errors();
  });
}

// ---- read-back ----------------------------------------------------------------------------------------------------
int cg_value_dims(const cg_function* f, unsigned* h, unsigned* w) {
  const Value& v = F(f)->value;
  *h = v.is_matrix ? v.h : 0;
  *w = v.is_matrix ? v.w : 0;
  return v.is_matrix ? 1 : 0;
}
void cg_value_get(const cg_function* f, float* dst) {
  guard_rc([&] { value_download(F(f)->value, dst); });
}
bool cg_all_less_than(const cg_function* f, float x) {
  bool r = false;
  guard_rc([&] { r = value_all_less_than(F(f)->value, x); });
  return r;
}
bool cg_all_equal(const cg_function* f, float x) {
  bool r = false;
  guard_rc([&] { r = value_all_equal(F(f)->value, x); });
  return r;
}
int cg_num_leaves(cg_function* f) { return (int)collect_param_leaves(*F(f)).size(); }
int cg_leaf_info(cg_function* f, int i, char* name, int name_cap, unsigned* h, unsigned* w) {
  auto leaves = collect_param_leaves(*F(f));
  if (i < 0 || i >= (int)leaves.size()) return -1;
  snprintf(name, (size_t)name_cap, "%s", leaves[i]->body->name.c_str());
  const Value& v = leaves[i]->value;
  *h = v.is_matrix ? v.h : 0;
  *w = v.is_matrix ? v.w : 0;
  return v.is_matrix ? 1 : 0;
}
int cg_leaf_get(cg_function* f, int i, float* dst) {
  auto leaves = collect_param_leaves(*F(f));
  if (i < 0 || i >= (int)leaves.size()) return -1;
  return guard_rc([&] { value_download(leaves[i]->value, dst); });
}
int cg_leaf_set(cg_function* f, int i, const float* src) {
  auto leaves = collect_param_leaves(*F(f));
  if (i < 0 || i >= (int)leaves.size()) return -1;
  return guard_rc([&] {
    value_upload_rowmajor(leaves[i]->value, src);
    leaves[i]->value.summary_valid = false;
  });
}
unsigned long long cg_leaf_address(cg_function* f, int i) {
  auto leaves = collect_param_leaves(*F(f));
  if (i < 0 || i >= (int)leaves.size() || !leaves[i]->value.buf) return 0;
  return (unsigned long long)(uintptr_t)leaves[i]->value.buf->ptr;
}
int cg_kind(const cg_function* f) { return (int)F(f)->body->kind; }

// ---- switches --------------------------------------------------------------------------------------------------------
void cg_set_mode(int m) { guard_rc([&] { Runtime::get().mode = m; }); }
void cg_set_tf32(int v) { guard_rc([&] { Runtime::get().tf32 = v; }); }
void cg_set_fix_act_grad(int v) { guard_rc([&] { Runtime::get().fix_act_grad = v; }); }
void cg_set_fuse(int v) { guard_rc([&] { Runtime::get().fuse = v; }); }
void cg_set_fix_broadcast_sum(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.fix_broadcast_sum != v) rt.dp_epoch++;  // the backward tape changes: invalidate cached plans
    rt.fix_broadcast_sum = v;
  });
}
void cg_set_fix_tied_params(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.fix_tied_params != v) rt.dp_epoch++;  // the backward tape changes: invalidate cached plans
    rt.fix_tied_params = v;
  });
}
void cg_set_fix_sub_sign(int v) { guard_rc([&] { Runtime::get().fix_sub_sign = v; }); }  // construction-time rule
// (these three choices are baked into a compiled plan — its captured graphs, its kernel choice and its resident-executor
// eligibility — so a change invalidates the cached plans, like every other setter below)
void cg_set_cuda_graph(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.use_graph != v) rt.dp_epoch++;
    rt.use_graph = v;
  });
}
void cg_set_strict_gemm_max(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.strict_gemm_max != v) rt.dp_epoch++;
    rt.strict_gemm_max = v;
  });
}
void cg_set_ew_jit_min(long long v) { guard_rc([&] { Runtime::get().ew_jit_min = v; }); }
int cg_ew_jit_available(void) { return ew_jit_available() ? 1 : 0; }
// Compiles (NVRTC -> sm_100a cubin, no GPU needed) a program that uses every opcode; returns the cubin size or -1.
long long cg_ew_jit_selftest(void) {
  long long bytes = -1;
  guard_rc([&] {
    EwProgram p{};
    p.n = 1 << 20;
    p.n_in = 3;
    p.n_out = 2;
    p.in_scalar_mask = 1u << 2;
    const EwOpcode ops[] = {OP_LDR, OP_SIGMOID, OP_STR, OP_LDR, OP_SQ, OP_RSUB, OP_TANH, OP_STG, OP_ABS_M, OP_SGN_M, OP_ABS_S,
                            OP_SGN_S, OP_ONE_MINUS, OP_NEG, OP_MULI, OP_ADD, OP_SUB, OP_MUL, OP_LDI, OP_MUL, OP_STG};
    const int operand[] = {0, 0, 3, 1, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 1, 2, 1, 3, 0, 0, 1};
    for (size_t k = 0; k < sizeof ops / sizeof ops[0]; k++) p.code[p.n_ins++] = ew_encode(ops[k], operand[k]);
    bytes = (long long)ew_jit_compile_only(p);
  });
  return bytes;
}
void cg_set_tf32_pair(int v) {
  guard_rc([&] {
    gemm_tf32_set_pair(v);
    Runtime::get().dp_epoch++;  // captured graphs hold the kernel choice: invalidate cached plans
  });
}
void cg_set_io_graph(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.io_graph != v) rt.dp_epoch++;
    rt.io_graph = v;
  });
}
int cg_set_plan_only(int on) {
  return guard_rc([&] {
    if (Runtime::get_if_initialised()) throw Error("cg_set_plan_only must be the first call into the library");
    set_plan_only(on != 0);
  });
}
void cg_set_vm_max_dim(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.vm_max_dim != v) rt.dp_epoch++;  // the executor choice is part of a compiled plan
    rt.vm_max_dim = v;
  });
}
void cg_set_elide_root(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.elide_root != v) rt.dp_epoch++;
    rt.elide_root = v;
  });
}
void cg_set_tape_jit(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.tape_jit != v) rt.dp_epoch++;  // the executor choice is part of a compiled plan
    rt.tape_jit = v;
  });
}
void cg_set_hoist(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.hoist != v) rt.dp_epoch++;  // the tape order is part of a compiled plan: invalidate cached plans
    rt.hoist = v;
  });
}
void cg_set_graph_streams(int v) {
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    if (rt.graph_streams != v) rt.dp_epoch++;  // captured graphs depend on it: invalidate cached plans
    rt.graph_streams = v;
  });
}

// ---- introspection ---------------------------------------------------------------------------------------------------
static Plan& ensure_plan(Function& f, float lr) {
  Runtime& rt = Runtime::get();
  if (!f.plan || f.plan->lr != lr || f.plan->mode != rt.mode || f.plan->tf32 != rt.tf32 ||
      f.plan->fix_act_grad != rt.fix_act_grad || f.plan->fuse != rt.fuse || f.plan->dp_epoch != rt.dp_epoch || f.plan->ew_jit_min != rt.ew_jit_min)
    minimize(f, {}, lr, 0, 0, nullptr, true, true);  // compiles (and captures) without running an iteration
  return *f.plan;
}
int cg_plan_stats_get(cg_function* f, float lr, cg_plan_stats* out) {
  return guard_rc([&] {
    const PlanStats& s = ensure_plan(*F(f), lr).stats;
    out->tape_instrs = s.tape_instrs;
    out->live_instrs = s.live_instrs;
    out->kernels_per_iter = s.kernels_per_iter;
    out->ew_kernels = s.ew_groups;
    out->gemms = s.gemms;
    out->gemms_fused_sgd = s.gemms_fused_sgd;
    out->reduces = s.reduces;
    out->arena_bytes = s.arena_bytes;
    out->gemm_flop_per_iter = s.gemm_flop_per_iter;
    out->ew_bytes_per_iter = s.ew_bytes_per_iter;
    out->ew_specialised = s.ew_specialised;
    out->allreduces = s.allreduces;
    out->dp_fused_updates = s.dp_fused_updates;
    out->vm_resident = s.vm_resident;
    out->tape_smem_bytes = s.tape_smem_bytes;
    out->quiet_variant = s.quiet_variant;
    out->quiet_ew_bytes_per_iter = s.quiet_ew_bytes_per_iter;
    out->dp_multimem_updates = s.dp_multimem_updates;
  });
}
long long cg_plan_describe(cg_function* f, float lr, char* buf, unsigned long long cap) {
  long long need = -1;
  guard_rc([&] {
    const std::string d = plan_describe(ensure_plan(*F(f), lr));
    need = (long long)d.size() + 1;
    if (buf && cap) {
      const size_t n = std::min<size_t>(d.size(), (size_t)cap - 1);
      memcpy(buf, d.data(), n);
      buf[n] = 0;
    }
  });
  return need;
}
unsigned long long cg_launch_count(void) {
  unsigned long long n = 0;
  guard_rc([&] { n = Runtime::get().launches; });
  return n;
}
double cg_time_iterations(cg_function* f, float lr, int warmup, int reps) {
  double ms = -1.0;
  guard_rc([&] {
    Runtime& rt = Runtime::get();
    Function& fn = *F(f);
    ensure_plan(fn, lr);
    minimize(fn, {}, lr, warmup, 0, nullptr, true, true);
    cudaEvent_t e0, e1;
    CG_CUDA(cudaEventCreate(&e0));
    CG_CUDA(cudaEventCreate(&e1));
    CG_CUDA(cudaEventRecord(e0, rt.stream));
    minimize(fn, {}, lr, reps, 0, nullptr, false, true);
    CG_CUDA(cudaEventRecord(e1, rt.stream));
    CG_CUDA(cudaEventSynchronize(e1));
    float t = 0.0f;
    CG_CUDA(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ms = (double)t / (reps > 0 ? reps : 1);
  });
  return ms;
}
int cg_set_device(int device) {
  return guard_rc([&] { CG_CUDA(cudaSetDevice(device)); });
}

// ==========================================================================================================
// The reference's 26-symbol matrix ABI (include/cg_matrix.h). Failure = "ERROR: msg" + exit(EXIT_FAILURE),
// exactly like src/cpu/util.cpp:5-10.
// ==========================================================================================================
static void die(const char* msg) {
  fprintf(stderr, "ERROR: %s\n", msg);
  exit(EXIT_FAILURE);
}
#define LEGACY_TRY(body)                \
  try {                                 \
    body                                \
  } catch (const std::exception& e) {   \
    die(e.what());                      \
  }
static size_t msize(const Matrix* m) { return (size_t)m->width * m->height; }
static void check_same(const Matrix* a, const Matrix* b, const char* hmsg, const char* wmsg) {
  if (a->height != b->height) die(hmsg);
  if (a->width != b->width) die(wmsg);
}

static void run_ew(size_t n, EwOpcode op, const float* a, const float* b, bool use_imm, float imm, bool imm_first, float* out) {
  // out = op(a)            (unary)
  // out = a op b           (elementwise)
  // out = a op imm / imm op a   (broadcast / broadcast_rev)
  Runtime& rt = Runtime::get();
  EwProgram p{};
  p.n = n;
  int c = 0;
  if (op == OP_LDI) {  // fill
    p.imm[0] = imm;
    p.code[c++] = ew_encode(OP_LDI, 0);
  } else {
    p.in[0] = a;
    p.n_in = 1;
    if (b) {
      p.in[1] = b;
      p.n_in = 2;
      p.code[c++] = ew_encode(OP_LDR, 0);
      p.code[c++] = ew_encode(op, 1);
    } else if (use_imm) {
      // the immediate is staged in register 1 so that a - imm and imm - a share one code path
      p.imm[0] = imm;
      p.code[c++] = ew_encode(OP_LDI, 0);
      p.code[c++] = ew_encode(OP_STR, 1);
      if (imm_first) {
        p.code[c++] = ew_encode(op, 0);  // acc(imm) op a
      } else {
        p.code[c++] = ew_encode(OP_LDR, 0);
        p.code[c++] = ew_encode(op, 1);  // a op imm
      }
    } else {
      p.code[c++] = ew_encode(OP_LDR, 0);
      if (op != OP_LDR) p.code[c++] = ew_encode(op);
    }
  }
  p.out[0] = out;
  p.n_out = 1;
  p.code[c++] = ew_encode(OP_STG, 0);
  p.n_ins = (uint32_t)c;
  launch_ew_program(p, rt.stream);
  rt.launches++;
}

void alloc_matrix(Matrix* m) {
  LEGACY_TRY(Runtime::get(); size_t n = msize(m); CG_CUDA(cudaMalloc(&m->array, (n ? n : 1) * sizeof(float)));)
}
void free_matrix(Matrix* m) {
  LEGACY_TRY(CG_CUDA(cudaStreamSynchronize(Runtime::get().stream)); CG_CUDA(cudaFree(m->array)); m->array = nullptr;)
}
void fill_matrix(Matrix* m, float v) { LEGACY_TRY(run_ew(msize(m), OP_LDI, nullptr, nullptr, true, v, false, m->array);) }
void copy_matrix(const Matrix* src, Matrix* dst) {
  check_same(src, dst, "copy_matrix: heights differ", "copy_matrix: widths differ");
  LEGACY_TRY(run_ew(msize(src), OP_LDR, src->array, nullptr, false, 0.f, false, dst->array);)
}
void get_array(const Matrix* m, float* dst) {
  LEGACY_TRY(Runtime& rt = Runtime::get();
             CG_CUDA(cudaMemcpyAsync(dst, m->array, msize(m) * sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
             CG_CUDA(cudaStreamSynchronize(rt.stream));)
}
void set_array(Matrix* m, const float* src, bool transpose) {
  LEGACY_TRY(
      Runtime& rt = Runtime::get(); const size_t n = msize(m);
      if (!transpose || m->height <= 1 || m->width <= 1) {
        CG_CUDA(cudaMemcpyAsync(m->array, src, n * sizeof(float), cudaMemcpyHostToDevice, rt.stream));
      } else {
        BufRef staging = DevBuf::alloc(n);
        CG_CUDA(cudaMemcpyAsync(staging->ptr, src, n * sizeof(float), cudaMemcpyHostToDevice, rt.stream));
        launch_rowmajor_to_colmajor(staging->ptr, m->array, m->height, m->width, rt.stream);
        rt.launches++;
        CG_CUDA(cudaStreamSynchronize(rt.stream));
      } CG_CUDA(cudaStreamSynchronize(rt.stream));)
}

#define LEGACY_MAP(name, OPC)                                                                          \
  void map_##name(const Matrix* m, Matrix* r) {                                                        \
    check_same(m, r, "m->height must equal result->height", "m->width must equal result->width");      \
    LEGACY_TRY(run_ew(msize(m), OPC, m->array, nullptr, false, 0.f, false, r->array);)                 \
  }
LEGACY_MAP(neg, OP_NEG)
LEGACY_MAP(sq, OP_SQ)
LEGACY_MAP(abs, OP_ABS_M)
LEGACY_MAP(signum, OP_SGN_M)
LEGACY_MAP(sigmoid, OP_SIGMOID)
LEGACY_MAP(tanh, OP_TANH)
LEGACY_MAP(one_minus, OP_ONE_MINUS)

#define LEGACY_BIN(name, OPC)                                                                          \
  void elemwise_##name(const Matrix* a, const Matrix* b, Matrix* r) {                                  \
    check_same(a, b, "m1->height must equal m2->height", "m1->width must equal m2->width");            \
    check_same(a, r, "m1->height must equal result->height", "m1->width must equal result->width");    \
    LEGACY_TRY(run_ew(msize(a), OPC, a->array, b->array, false, 0.f, false, r->array);)                \
  }                                                                                                    \
  void broadcast_##name(const Matrix* m, float v, Matrix* r) {                                         \
    check_same(m, r, "m->height must equal result->height", "m->width must equal result->width");      \
    LEGACY_TRY(run_ew(msize(m), OPC, m->array, nullptr, true, v, false, r->array);)                    \
  }                                                                                                    \
  void broadcast_##name##_rev(float v, const Matrix* m, Matrix* r) {                                   \
    check_same(m, r, "m->height must equal result->height", "m->width must equal result->width");      \
    LEGACY_TRY(run_ew(msize(m), OPC, m->array, nullptr, true, v, true, r->array);)                     \
  }
LEGACY_BIN(add, OP_ADD)
LEGACY_BIN(sub, OP_SUB)
LEGACY_BIN(mul, OP_MUL)

void gemm(const Matrix* m1, bool t1, const Matrix* m2, bool t2, Matrix* r) {
  const unsigned h1 = t1 ? m1->width : m1->height, inner = t1 ? m1->height : m1->width;
  const unsigned h2 = t2 ? m2->width : m2->height, w2 = t2 ? m2->height : m2->width;
  if (h1 != r->height) die("height1 must equal result->height");
  if (inner != h2) die("inner_dim must equal height2");
  if (w2 != r->width) die("width2 must equal result->width");
  LEGACY_TRY(Runtime& rt = Runtime::get();
             launch_gemm(m1->array, t1, m2->array, t2, r->array, (int)h1, (int)w2, (int)inner, (int)m1->height,
                         (int)m2->height, (int)r->height, GemmEpilogue{}, rt.tf32, rt.stream);
             rt.launches++;)
}
static bool legacy_cmp(bool less, const Matrix* m, float x) {
  Runtime& rt = Runtime::get();
  if (less) launch_all_less_than(m->array, msize(m), x, rt.flag, rt.stream);
  else launch_all_equal(m->array, msize(m), x, rt.flag, rt.stream);
  rt.launches++;
  int violation = 0;
  CG_CUDA(cudaMemcpyAsync(&violation, rt.flag, sizeof(int), cudaMemcpyDeviceToHost, rt.stream));
  CG_CUDA(cudaStreamSynchronize(rt.stream));
  return violation == 0;
}
bool all_equal(const Matrix* m, float x) {
  bool r = false;
  LEGACY_TRY(r = legacy_cmp(false, m, x);)
  return r;
}
bool all_less_than(const Matrix* m, float x) {
  bool r = false;
  LEGACY_TRY(r = legacy_cmp(true, m, x);)
  return r;
}
float reduce_sum(const Matrix* m) {
  float out = 0.0f;
  LEGACY_TRY(Runtime& rt = Runtime::get(); float* d = reinterpret_cast<float*>(rt.flag + 4);
             launch_reduce_div(m->array, msize(m), 1.0f, d, rt.red_ws, rt.stream); rt.launches++;
             CG_CUDA(cudaMemcpyAsync(&out, d, sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
             CG_CUDA(cudaStreamSynchronize(rt.stream));)
  return out;
}

}  // extern "C"

// Instrumented pass: one iteration launched eagerly with a CUDA event between every step, on the library
// stream. kind_ms / kind_count are indexed by Step::Kind {EW, AVG, GEMM, ALLREDUCE}; max_gemm_* describe the
// largest product (by flop) for the roofline line of bench.py.
extern "C" int cg_profile_iteration(cg_function* f, float lr, double kind_ms[4], unsigned long long kind_count[4],
                                    double* max_gemm_flop, double* max_gemm_ms_avg, unsigned long long* max_gemm_count) {
  return guard_rc([&] {
    Runtime& rt = Runtime::get();
    Plan& p = ensure_plan(*F(f), lr);
    const size_t n = p.steps.size();
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) CG_CUDA(cudaEventCreate(&e));
    const int saved = rt.use_graph;
    CG_CUDA(cudaEventRecord(ev[0], rt.stream));
    for (size_t i = 0; i < n; i++) {
      plan_launch_step(p, i, rt.stream);
      CG_CUDA(cudaEventRecord(ev[i + 1], rt.stream));
    }
    rt.use_graph = saved;
    rt.launches += n;
    CG_CUDA(cudaStreamSynchronize(rt.stream));
    for (int k = 0; k < 4; k++) kind_ms[k] = 0.0, kind_count[k] = 0;
    double best_flop = 0.0;
    for (size_t i = 0; i < n; i++)
      if (p.steps[i].kind == Step::GEMM) best_flop = std::max(best_flop, plan_step_flop(p, i));
    double best_ms = 0.0;
    unsigned long long best_cnt = 0;
    for (size_t i = 0; i < n; i++) {
      float ms = 0.0f;
      CG_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      const int kind = p.steps[i].kind == Step::DPUPDATE ? (int)Step::ALLREDUCE : (int)p.steps[i].kind;  // the exchange step
      kind_ms[kind] += ms;
      kind_count[kind]++;
      if (p.steps[i].kind == Step::GEMM && plan_step_flop(p, i) == best_flop) best_ms += ms, best_cnt++;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    *max_gemm_flop = best_flop;
    *max_gemm_ms_avg = best_cnt ? best_ms / (double)best_cnt : 0.0;
    *max_gemm_count = best_cnt;
  });
}
