// graph.cpp — graph construction with the reference's semantics (function.rs, ops.rs constructors).
#include "graph.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <unordered_set>

namespace cg {

// Plan-only mode (cg_set_plan_only): graphs and plans are built with placeholder addresses so that the tape
// compiler — fusion, hoisting, arena assignment, kernel counts — can be inspected and tested on a machine without a
// GPU. Nothing is ever computed in this mode: every call that would touch the device raises.
static bool g_plan_only = false;
bool plan_only() { return g_plan_only; }
void set_plan_only(bool on) { g_plan_only = on; }
static void no_device(const char* what) {
  throw Error(std::string("plan-only mode: ") + what + " needs a CUDA device (this backend has no CPU path)");
}

static std::unordered_map<const float*, std::weak_ptr<std::vector<float>>> g_shadows;  // plan-only: placeholder -> contents
const std::vector<float>* plan_only_shadow(const float* ptr) {
  auto it = g_shadows.find(ptr);
  if (it == g_shadows.end()) return nullptr;
  auto sp = it->second.lock();
  return sp ? sp.get() : nullptr;  // (the owning DevBuf keeps it alive for as long as the plan can name the buffer)
}

DevBuf::~DevBuf() {
  if (placeholder) g_shadows.erase(ptr);
  if (owned && ptr && !placeholder) {
    if (symm) symm_free(ptr);  // back to the multicast heap's free list (symm.cpp); never unmapped under a peer
    else if (!dp_retire(ptr, (n ? n : 1) * sizeof(float))) cudaFree(ptr);  // peers may hold this buffer open (dp.cpp)
  }
}
BufRef DevBuf::alloc_symm(size_t n) {
  if (g_plan_only) return alloc(n);
  auto b = std::make_shared<DevBuf>();
  b->n = n;
  b->owned = true;
  b->symm = true;
  b->ptr = static_cast<float*>(symm_alloc((n ? n : 1) * sizeof(float)));
  return b;
}
BufRef DevBuf::alloc(size_t n) {
  auto b = std::make_shared<DevBuf>();
  b->n = n;
  b->owned = true;
  if (g_plan_only) {
    // distinct, never dereferenced, 256-byte aligned addresses from a range no allocator hands out
    static uintptr_t next = (uintptr_t)1 << 56;
    b->ptr = reinterpret_cast<float*>(next);
    next += (((n ? n : 1) * sizeof(float)) + 255) / 256 * 256;
    b->placeholder = true;
    if (n <= 65536) {  // resident plans only
      b->shadow = std::make_shared<std::vector<float>>(n ? n : 1, 0.0f);
      g_shadows[b->ptr] = b->shadow;
    }
    return b;
  }
  CG_CUDA(cudaMalloc(&b->ptr, (n ? n : 1) * sizeof(float)));
  return b;
}

// An end-to-end `minimize(&args, lr, 1)` call copies its inputs on a stream of its own while the iteration graph runs.
// The graph's branches and that stream are mapped onto the device's hardware work queues — 8 by default — and a copy that
// lands in a queue behind kernels of the graph waits for them: at config 5 (33 inputs, 4.4 GB) the call took 293 ms
// against 250 ms for the iteration itself, 307 ms with 2 queues and 257 ms with 32 (profiles/r02f_e2e_probe_T32_conn*.json).
// So the library asks for 32 queues unless the host has chosen; this must happen before the CUDA context exists, hence
// a load-time initialiser (the Python host layer and bench.py set the same default before importing torch).
static const int g_connections_default = [] {
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", /*overwrite=*/0);
  return 0;
}();

static Runtime g_rt;
Runtime& Runtime::get() {
  if (!g_rt.stream) g_rt.init();
  return g_rt;
}
bool Runtime::get_if_initialised() { return g_rt.stream != nullptr; }
void Runtime::init() {
  if (g_plan_only) {
    stream = reinterpret_cast<cudaStream_t>((uintptr_t)1);  // placeholder: marks the runtime as initialised
    ew_jit_min = -1;                                        // specialised kernels are loaded on a device
  } else {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw Error("libcg_b200: no CUDA device available — this backend has no CPU fallback");
    CG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    CG_CUDA(cudaMalloc(&flag, 16 * sizeof(int)));  // [0] comparison flag, [4] reduce_sum result
    CG_CUDA(cudaMalloc(&red_ws, reduce_workspace_floats() * sizeof(float)));
    CG_CUDA(cudaMemset(red_ws, 0, reduce_workspace_floats() * sizeof(float)));
  }
  if (const char* s = getenv("CG_MODE")) mode = (strcmp(s, "dag") == 0) ? 1 : 0;
  if (const char* s = getenv("CG_TF32")) tf32 = atoi(s);
  if (const char* s = getenv("CG_FIX_ACT_GRAD")) fix_act_grad = atoi(s);
  if (const char* s = getenv("CG_FUSE")) fuse = atoi(s);
  if (const char* s = getenv("CG_CUDA_GRAPH")) use_graph = atoi(s);
  if (const char* s = getenv("CG_GRAPH_STREAMS")) graph_streams = atoi(s);
  if (const char* s = getenv("CG_IO_GRAPH")) io_graph = atoi(s);
  if (const char* s = getenv("CG_HOIST")) hoist = atoi(s);
  if (const char* s = getenv("CG_VM_MAX_DIM")) vm_max_dim = atoi(s);
  if (const char* s = getenv("CG_TAPE_JIT")) tape_jit = atoi(s);
  if (const char* s = getenv("CG_ELIDE_ROOT")) elide_root = atoi(s);
  if (const char* s = getenv("CG_EW_JIT_MIN")) ew_jit_min = atoll(s);
  if (g_plan_only) ew_jit_min = -1;
}

static void fill_device(float* ptr, size_t n, float v) {
  if (g_plan_only) return;  // (the caller fills the shadow)
  Runtime& rt = Runtime::get();
  EwProgram p{};
  p.n = n;
  p.n_in = 0;
  p.n_out = 1;
  p.out[0] = ptr;
  p.imm[0] = v;
  p.code[0] = ew_encode(OP_LDI, 0);
  p.code[1] = ew_encode(OP_STG, 0);
  p.n_ins = 2;
  launch_ew_program(p, rt.stream);
}

// ---- leaves -------------------------------------------------------------------------------------------
static Function* new_leaf(Kind kind, const char* name, Value v) {
  auto* f = new Function();
  f->value = std::move(v);
  f->has_params = (kind == Kind::Param || kind == Kind::Input);
  f->body = std::make_shared<Expr>();
  f->body->kind = kind;
  if (name) f->body->name = name;
  return f;
}

Function* make_scalar_leaf(Kind kind, const char* name, float x) {
  Runtime& rt = Runtime::get();
  Value v;
  v.is_matrix = false;
  v.uniform = true;
  v.uval = x;
  v.buf = DevBuf::alloc(1);
  if (!g_plan_only) {
    CG_CUDA(cudaMemcpyAsync(v.buf->ptr, &x, sizeof(float), cudaMemcpyHostToDevice, rt.stream));
    CG_CUDA(cudaStreamSynchronize(rt.stream));
  } else if (v.buf->shadow) {
    (*v.buf->shadow)[0] = x;
  }
  return new_leaf(kind, name, std::move(v));
}

void value_upload_rowmajor(Value& v, const float* src, bool sync) {
  if (g_plan_only) {
    // the layout change the upload kernel would do: row-major host values -> column-major storage
    if (!v.buf->shadow) return;
    std::vector<float>& sh = *v.buf->shadow;
    if (!v.is_matrix) sh[0] = src[0];
    else
      for (unsigned i = 0; i < v.h; i++)
        for (unsigned j = 0; j < v.w; j++) sh[(size_t)j * v.h + i] = src[(size_t)i * v.w + j];
    return;
  }
  // Matrix::new -> set_array(values, transpose = true) (matrix.rs:59-66, cpu/matrix.cpp:26-36)
  Runtime& rt = Runtime::get();
  const size_t n = v.size();
  if (!v.is_matrix || v.h == 1 || v.w == 1) {  // a scalar or a vector: both layouts coincide
    CG_CUDA(cudaMemcpyAsync(v.buf->ptr, src, n * sizeof(float), cudaMemcpyHostToDevice, rt.stream));
  } else {
    // stream-ordered: the staging buffer is reused by the next upload only after this transpose has run
    if (!rt.staging || rt.staging->n < n) {
      CG_CUDA(cudaStreamSynchronize(rt.stream));
      rt.staging = DevBuf::alloc(n);
    }
    CG_CUDA(cudaMemcpyAsync(rt.staging->ptr, src, n * sizeof(float), cudaMemcpyHostToDevice, rt.stream));
    launch_rowmajor_to_colmajor(rt.staging->ptr, v.buf->ptr, v.h, v.w, rt.stream);
    rt.launches++;
  }
  if (sync) CG_CUDA(cudaStreamSynchronize(rt.stream));  // the caller may free `src` as soon as we return
}

Function* make_matrix_leaf(Kind kind, const char* name, unsigned h, unsigned w, const float* vals) {
  Value v;
  v.is_matrix = true;
  v.h = h;
  v.w = w;
  const size_t n = (size_t)h * w;
  v.buf = DevBuf::alloc(n);
  // host summary for the identity elimination: all elements equal?
  v.uniform = n > 0;
  v.uval = n ? vals[0] : 0.0f;
  for (size_t i = 1; i < n && v.uniform; i++)
    if (!(vals[i] == v.uval)) v.uniform = false;
  if (n && vals[0] != vals[0]) v.uniform = false;  // NaN never equals anything
  value_upload_rowmajor(v, vals);
  return new_leaf(kind, name, std::move(v));
}

Function* make_fill_leaf(Kind kind, const char* name, unsigned h, unsigned w, float x) {
  // Constant::single_val (constant.rs:30-43): alloc + fill_matrix
  Runtime& rt = Runtime::get();
  Value v;
  v.is_matrix = true;
  v.h = h;
  v.w = w;
  v.uniform = (size_t)h * w > 0 && x == x;
  v.uval = x;
  v.buf = DevBuf::alloc((size_t)h * w);
  fill_device(v.buf->ptr, (size_t)h * w, x);
  if (g_plan_only) {
    if (v.buf->shadow) std::fill(v.buf->shadow->begin(), v.buf->shadow->end(), x);
    return new_leaf(kind, name, std::move(v));
  }
  rt.launches++;
  CG_CUDA(cudaStreamSynchronize(rt.stream));
  return new_leaf(kind, name, std::move(v));
}

Function* make_input(const char* name, bool is_matrix, unsigned h, unsigned w) {
  // Function::input (function.rs:105-108): value = Constant::empty(dims) — uninitialised => undefined
  Value v;
  v.is_matrix = is_matrix;
  v.h = is_matrix ? h : 0;
  v.w = is_matrix ? w : 0;
  v.defined = false;
  v.buf = DevBuf::alloc(v.size());
  if (!g_plan_only) CG_CUDA(cudaMemset(v.buf->ptr, 0, v.size() * sizeof(float)));
  return new_leaf(Kind::Input, name, std::move(v));
}

// ---- clone: #[derive(Clone)] (function.rs:10-16) + Matrix::clone (matrix.rs:100-106) ---------------------
static Value clone_value(const Value& v) {
  Value c = v;
  if (v.buf) {
    Runtime& rt = Runtime::get();
    c.buf = DevBuf::alloc(v.size());
    if (!g_plan_only) {
      CG_CUDA(cudaMemcpyAsync(c.buf->ptr, v.buf->ptr, v.size() * sizeof(float), cudaMemcpyDeviceToDevice, rt.stream));
      CG_CUDA(cudaStreamSynchronize(rt.stream));
    } else if (v.buf->shadow && c.buf->shadow) {
      *c.buf->shadow = *v.buf->shadow;
    }
  }
  return c;
}
// Non-leaf nodes only need the *summary* of their (stale) initial value: it is never read by the loop
// (forward overwrites it before any use), only by later identity checks.
static Value summary_of(const Value& v) {
  Value c = v;
  c.buf.reset();
  return c;
}

Function* clone_function(const Function& f) {
  auto* g = new Function();
  // Constants and Inputs are never written by the loop (only Params are, optimize.rs:149-152), so their clones can
  // share one device buffer: the reference's per-clone copy (matrix.rs:100-106) is unobservable for them, and an
  // Input named once in `args` is then uploaded once, not once per use site. Params keep a private copy (Q1).
  const bool read_only_leaf = f.body && (f.body->kind == Kind::Constant || f.body->kind == Kind::Input);
  g->value = read_only_leaf ? f.value : clone_value(f.value);
  g->has_params = f.has_params;
  g->body = f.body;
  return g;
}

// ---- value predicates ------------------------------------------------------------------------------------
static bool device_cmp(bool less, const Value& v, float x) {
  if (g_plan_only) no_device("comparing a computed value");
  Runtime& rt = Runtime::get();
  if (!v.buf) throw Error("value has not been evaluated yet");
  if (less) launch_all_less_than(v.buf->ptr, v.size(), x, rt.flag, rt.stream);
  else launch_all_equal(v.buf->ptr, v.size(), x, rt.flag, rt.stream);
  rt.launches++;
  int violation = 0;
  CG_CUDA(cudaMemcpyAsync(&violation, rt.flag, sizeof(int), cudaMemcpyDeviceToHost, rt.stream));
  CG_CUDA(cudaStreamSynchronize(rt.stream));
  return violation == 0;
}

bool value_all_equal(const Value& v, float x) {
  if (!v.defined) return false;
  if (v.summary_valid) return v.uniform && v.uval == x;
  return device_cmp(false, v, x);
}
bool value_all_less_than(const Value& v, float x) {
  if (v.summary_valid && v.uniform) return v.uval < x;
  return device_cmp(true, v, x);
}
void value_download(const Value& v, float* dst) {
  if (g_plan_only) no_device("reading a value back");
  Runtime& rt = Runtime::get();
  if (!v.buf) throw Error("value has not been evaluated yet");
  CG_CUDA(cudaMemcpyAsync(dst, v.buf->ptr, v.size() * sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
  CG_CUDA(cudaStreamSynchronize(rt.stream));
}

// ---- op constructors ---------------------------------------------------------------------------------------
static float unary_identity(Kind k) { return k == Kind::Sigmoid ? 0.5f : 0.0f; }  // ops.rs:331-336,366

Function* make_unary(Kind kind, const Function& f) {
  // Function1! with identity (ops.rs:45-53): `if f.all_equal(identity) { f.clone() }`
  if (value_all_equal(f.value, unary_identity(kind))) return clone_function(f);
  auto* g = new Function();
  g->value = summary_of(f.value);  // `$f.value().clone()` (ops.rs:39)
  g->has_params = f.has_params;
  g->body = std::make_shared<Expr>();
  g->body->kind = kind;
  g->body->f1.reset(clone_function(f));
  return g;
}

static Function* fold_constant(Function* g);

Function* make_binary(Kind kind, const Function& a, const Function& b) {
  const float identity = kind == Kind::Mul ? 1.0f : 0.0f;  // ops.rs:338-343 (Sub shares Add's rule, Q4)
  // corrected semantics (cg_set_fix_sub_sign): 0 - x keeps its sign
  if (kind == Kind::Sub && Runtime::get().fix_sub_sign && value_all_equal(a.value, identity)) return make_unary(Kind::Neg, b);
  if (value_all_equal(a.value, identity)) return clone_function(b);  // ops.rs:255-259
  if (value_all_equal(b.value, identity)) return clone_function(a);
  auto* g = new Function();
  // value-shape rule, ops.rs:218-228
  const Value &va = a.value, &vb = b.value;
  if (va.is_matrix && vb.is_matrix && (va.h != vb.h || va.w != vb.w)) {
    delete g;
    throw Error("assertion failed: m1.dims() == m2.dims()");
  }
  g->value = summary_of((!va.is_matrix && vb.is_matrix) ? vb : va);
  g->has_params = a.has_params || b.has_params;
  g->body = std::make_shared<Expr>();
  g->body->kind = kind;
  g->body->f1.reset(clone_function(a));
  g->body->f2.reset(clone_function(b));
  if (a.body->kind == Kind::Constant && b.body->kind == Kind::Constant) return fold_constant(g);
  return g;
}

Function* make_dot(const Function& a, bool t1, const Function& b, bool t2) {
  if (!a.value.is_matrix || !b.value.is_matrix) throw Error("dot should not be used with scalars");  // ops.rs:420
  const unsigned inner1 = t1 ? a.value.h : a.value.w, inner2 = t2 ? b.value.w : b.value.h;
  if (inner1 != inner2) throw Error("inner_dim must equal height2");  // cpu/ops.cpp:98
  auto* g = new Function();
  g->value.is_matrix = true;  // Matrix::empty_for_dot (matrix.rs:68-83)
  g->value.h = t1 ? a.value.w : a.value.h;
  g->value.w = t2 ? b.value.h : b.value.w;
  g->value.defined = false;  // uninitialised in the reference
  g->has_params = a.has_params || b.has_params;
  g->body = std::make_shared<Expr>();
  g->body->kind = Kind::Dot;
  g->body->t1 = t1;
  g->body->t2 = t2;
  g->body->f1.reset(clone_function(a));
  g->body->f2.reset(clone_function(b));
  if (a.body->kind == Kind::Constant && b.body->kind == Kind::Constant) return fold_constant(g);  // ops.rs:398-400
  return g;
}

// const (+) const folding (ops.rs:209-213). The reference folds through `eval`, whose
// `Constant (+) Constant` computes `empty (+) rhs` (Q9, a bug); we fold the mathematics, on the device.
static Function* fold_constant(Function* g) {
  std::unique_ptr<Function> node(g);
  minimize(*node, {}, 0.0f, 1, 2, nullptr, true);  // one forward pass; no params => backward is empty
  auto* c = new Function();
  c->value = clone_value(node->value);
  c->value.defined = true;
  c->value.summary_valid = false;
  c->has_params = false;
  c->body = std::make_shared<Expr>();
  c->body->kind = Kind::Constant;
  return c;
}

// ---- symbolic gradient (optimize.rs:34-56): the README's "naive" way of getting gradients ------------------------
// `f.grad(param)` builds a new Function for df/dparam out of the ordinary constructors (so the reference's
// construction-time rules — identity elimination, constant folding — apply to it as well). The reference's Sq rule,
// `&f.grad(param) * f`, drops the factor 2 (optimize.rs:38); this is the mathematics: d(f^2) = 2 f f'.
static bool mentions(const Function& f, const std::string& name, std::unordered_set<const Expr*>& seen) {
  const Expr& e = *f.body;
  if (e.kind == Kind::Param || e.kind == Kind::Input) return e.name == name;  // function.rs:105-112: both name a "param"
  if (e.kind == Kind::Constant) return false;
  if (!seen.insert(&e).second) return false;  // already searched through another path
  return (e.f1 && mentions(*e.f1, name, seen)) || (e.f2 && mentions(*e.f2, name, seen));
}

Function* function_grad(const Function& f, const std::string& param) {
  typedef std::unique_ptr<Function> P;
  {
    std::unordered_set<const Expr*> seen;
    if (!mentions(f, param, seen)) return make_scalar_leaf(Kind::Constant, nullptr, 0.0f);  // optimize.rs:53-55
  }
  const Expr& e = *f.body;
  auto G = [&](const Function& c) { return P(function_grad(c, param)); };
  auto un = [](Kind k, const Function& a) { return P(make_unary(k, a)); };
  auto bin = [](Kind k, const Function& a, const Function& b) { return P(make_binary(k, a, b)); };
  switch (e.kind) {
    case Kind::Neg: return make_unary(Kind::Neg, *G(*e.f1));
    case Kind::Sq: {
      P two(make_scalar_leaf(Kind::Constant, nullptr, 2.0f));
      return make_binary(Kind::Mul, *bin(Kind::Mul, *two, *G(*e.f1)), *e.f1);
    }
    case Kind::Abs: return make_binary(Kind::Mul, *un(Kind::Signum, *e.f1), *G(*e.f1));  // :39
    case Kind::Signum: throw Error("signum is nondifferentiable");                        // :40
    case Kind::Sigmoid: {                                                                 // :41-42: f' * (s * (1 - s))
      P one(make_scalar_leaf(Kind::Constant, nullptr, 1.0f));
      return make_binary(Kind::Mul, *G(*e.f1), *bin(Kind::Mul, f, *bin(Kind::Sub, *one, f)));
    }
    case Kind::Tanh: {                                                                    // :43-44: f' * (1 - t^2)
      P one(make_scalar_leaf(Kind::Constant, nullptr, 1.0f));
      return make_binary(Kind::Mul, *G(*e.f1), *bin(Kind::Sub, *one, *un(Kind::Sq, f)));
    }
    case Kind::Add: return make_binary(Kind::Add, *G(*e.f1), *G(*e.f2));
    case Kind::Sub: return make_binary(Kind::Sub, *G(*e.f1), *G(*e.f2));
    case Kind::Mul:                                                                       // :47-48
      return make_binary(Kind::Add, *bin(Kind::Mul, *G(*e.f1), *e.f2), *bin(Kind::Mul, *G(*e.f2), *e.f1));
    case Kind::Dot: throw Error("not implemented");                                       // :49
    case Kind::Param:                                                                     // :50-51: ones like the value
      return f.value.is_matrix ? make_fill_leaf(Kind::Constant, nullptr, f.value.h, f.value.w, 1.0f)
                               : make_scalar_leaf(Kind::Constant, nullptr, 1.0f);
    default: throw Error("should never reach here");                                      // :52 (an Input named `param`)
  }
}

// Param leaves in memoised post-order (children f1 then f2); every use site is its own leaf (Q1).
static void collect(Function& f, std::unordered_set<const Expr*>& seen, std::vector<Function*>& out) {
  Expr& e = *f.body;
  if (e.kind == Kind::Param) {
    out.push_back(&f);
    return;
  }
  if (is_leaf(e.kind)) return;
  if (!seen.insert(&e).second) return;
  if (e.f1) collect(*e.f1, seen, out);
  if (e.f2) collect(*e.f2, seen, out);
}
std::vector<Function*> collect_param_leaves(Function& root) {
  std::unordered_set<const Expr*> seen;
  std::vector<Function*> out;
  collect(root, seen, out);
  return out;
}

}  // namespace cg
