// plan.cpp — tape compiler: Function tree -> SSA tape -> fused steps -> device arena -> CUDA graph.
//
// The walker restates src/optimize.rs:84-204 (assign_values / backprop) and the Constant dispatch
// tables of src/ops.rs:61-158 as *emitters*: instead of computing, each rule appends micro-ops.
//   exact mode: forward memoised per shared body (pure => bit-identical to the reference's
//               re-evaluation), backward = the literal pre-order tree walk, one tape segment per
//               root->node path, leaf SGD emitted where the walk reaches the leaf (Q2).
//   dag mode  : backward visits every unique node once in reverse topological order, summing the
//               contributions of its consumers in consumer-visit order (SURVEY §8 mode table).
#include "tape.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <map>
#include <unordered_set>

namespace cg {

Plan::~Plan() {
  tape_jit_free(tape_kernel);
  if (vm_steps) cudaFree(vm_steps);
  if (vm_progs) cudaFree(vm_progs);
  release_graphs();
  for (cudaEvent_t e : events) cudaEventDestroy(e);
  if (io_done) cudaEventDestroy(io_done);
}

void Plan::release_graphs() {
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (graph) cudaGraphDestroy(graph);
  graph_exec = nullptr;
  graph = nullptr;
  for (IoGraph& g : io) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    for (cudaEvent_t e : g.in_ev) cudaEventDestroy(e);  // after the graph whose wait nodes name them
    g = IoGraph();
  }
}

static std::vector<std::weak_ptr<Plan>>& plan_registry() {
  static std::vector<std::weak_ptr<Plan>> r;
  return r;
}

void release_all_plan_graphs() {
  auto& reg = plan_registry();
  std::vector<std::weak_ptr<Plan>> alive;
  for (auto& w : reg)
    if (auto p = w.lock()) {
      p->release_graphs();
      alive.push_back(w);
    }
  reg.swap(alive);
}

namespace {

enum UOp { U_NEG, U_SQ, U_ABS, U_SIGNUM, U_SIGMOID, U_TANH, U_ONE_MINUS };
enum BOp { B_ADD, B_SUB, B_MUL };

[[noreturn]] void panic(const char* msg) { throw Error(std::string("panic: ") + msg); }
[[noreturn]] void check_fail(const char* msg) { throw Error(std::string("ERROR: ") + msg); }  // cpu/util.cpp:5-10

bool unary_nonlinear(UOp op) { return op == U_SQ || op == U_ABS || op == U_SIGNUM || op == U_SIGMOID || op == U_TANH; }

struct Builder {
  Plan& p;
  float lr;
  int fix_act_grad;
  std::unordered_map<const Function*, int> leaf_cur;  // leaf use site -> current SSA version
  std::unordered_map<const Expr*, int> node_val;      // shared body -> forward value
  std::unordered_set<const Function*> seen_params;
  std::unordered_set<const float*> seen_input_bufs;

  bool dp = false;
  bool in_backward = false;      // set around the backward walk: only there does fix_broadcast_sum apply
  int fix_broadcast_sum = 0;
  // Corrected semantics (cg_set_fix_tied_params): every use of a parameter NAME is one parameter. The
  // per-use-site copies stay (Q1's storage), but the backward walk only accumulates what each copy receives into one
  // gradient per name — so every forward value it reads is the un-updated one — and after the walk every copy of the
  // name takes the same update, p -= lr * (sum of contributions): the copies stay bit-identical.
  int fix_tied_params = 0;
  struct Tie {
    int grad = -1;
    std::vector<Function*> leaves;
  };
  std::vector<std::pair<std::string, Tie>> ties;  // in first-contribution order (deterministic tape)

  Builder(Plan& plan, float lr_, int fix) : p(plan), lr(lr_), fix_act_grad(fix), dp(dp_enabled()) {}

  static Shape shape_of(const Function& f) { return Shape{f.value.is_matrix, f.value.h, f.value.w}; }

  int new_val(const Shape& sh) {
    Val v;
    v.sh = sh;
    p.vals.push_back(std::move(v));
    return (int)p.vals.size() - 1;
  }
  int new_fixed(const Shape& sh, const BufRef& buf) {
    int id = new_val(sh);
    p.vals[id].fixed = buf->ptr;
    p.vals[id].keep = buf;
    p.vals[id].persistent = true;
    return id;
  }
  Shape sh(int v) const { return p.vals[v].sh; }  // by value: emitting grows p.vals

  int emit(IOp op, const Shape& dsh, int a = -1, int b = -1, float imm = 0.0f) {
    Instr i;
    i.op = op;
    i.dst = new_val(dsh);
    i.a = a;
    i.b = b;
    i.imm = imm;
    p.tape.push_back(i);
    return i.dst;
  }

  // ---- leaves ------------------------------------------------------------------------------------------
  int leaf_val(Function& f) {
    auto it = leaf_cur.find(&f);
    if (it != leaf_cur.end()) return it->second;
    if (!f.value.buf) throw Error("leaf without device storage");
    int id = new_fixed(shape_of(f), f.value.buf);
    leaf_cur[&f] = id;
    // clones of an Input share its buffer (graph.cpp clone_function): one upload per distinct buffer
    if (f.body->kind == Kind::Input && seen_input_bufs.insert(f.value.buf->ptr).second) p.input_leaves.push_back(&f);
    if (f.body->kind == Kind::Param && seen_params.insert(&f).second) p.param_leaves.push_back(&f);
    return id;
  }
  // `value!(f)`: the cached forward value of a child (macros.rs:43-49)
  int read(Function& f) {
    if (is_leaf(f.body->kind)) return leaf_val(f);
    auto it = node_val.find(f.body.get());
    if (it == node_val.end()) throw Error("internal: node read before evaluation");
    return it->second;
  }

  // ---- Constant dispatch tables as emitters ---------------------------------------------------------------
  // ops.rs:430-442. Corrected semantics (cg_set_fix_broadcast_sum): in the backward pass a scalar absorbing a matrix
  // error takes the SUM (imm = the divisor: 1) — the gradient of a broadcast scalar — instead of the reference's mean.
  int avg(int a) {
    if (!sh(a).m) return a;
    const int v = emit(IOp::Avg, Shape{}, a);
    if (in_backward && fix_broadcast_sum) p.tape.back().imm = 1.0f;
    return v;
  }

  static IOp unary_iop(UOp op, bool matrix_semantics) {
    switch (op) {
      case U_NEG: return IOp::Neg;
      case U_SQ: return IOp::Sq;
      case U_ABS: return matrix_semantics ? IOp::AbsM : IOp::AbsS;  // Q8
      case U_SIGNUM: return matrix_semantics ? IOp::SgnM : IOp::SgnS;
      case U_SIGMOID: return IOp::Sigmoid;
      case U_TANH: return IOp::Tanh;
      case U_ONE_MINUS: return IOp::OneMinus;
    }
    return IOp::Neg;
  }
  static IOp binary_iop(BOp op) { return op == B_ADD ? IOp::Add : op == B_SUB ? IOp::Sub : IOp::Mul; }

  void check_dims(const Shape& a, const Shape& b, const char* msg) {
    if (a.m && b.m && (a.h != b.h || a.w != b.w)) check_fail(msg);
  }

  // Constant1!(linear|nonlinear: op, result, arg) — ops.rs:70-99
  int equals_unary(UOp op, const Shape& r, int a) {
    const Shape as = sh(a);
    if (unary_nonlinear(op) && !r.m && as.m)
      panic("The result of this operation yields a matrix. Can't place a matrix result in a scalar");
    if (!r.m && as.m) return emit(unary_iop(op, false), r, avg(a));                 // f(avg(m))
    if (r.m && as.m) check_dims(r, as, "m->height must equal result->height");
    return emit(unary_iop(op, r.m && as.m), r, a);  // (S,S) closure | (M,S) fill(f(x)) | (M,M) map_*
  }
  // Constant1!(op, c) in place — ops.rs:62-68
  int assign_unary(UOp op, int c) { return emit(unary_iop(op, sh(c).m), sh(c), c); }
  // Constant2!(op, result, other): result op= other — ops.rs:102-116
  int op_assign(BOp op, int r, int o, int dst = -1) {
    const Shape rs = sh(r);
    int rhs = (!rs.m && sh(o).m) ? avg(o) : o;
    if (rs.m && sh(o).m) check_dims(rs, sh(o), "m1->height must equal m2->height");
    if (dst < 0) return emit(binary_iop(op), rs, r, rhs);
    Instr i;
    i.op = binary_iop(op);
    i.dst = dst;
    i.a = r;
    i.b = rhs;
    p.tape.push_back(i);
    return dst;
  }
  int mul_imm(int r, float s) { return emit(IOp::MulI, sh(r), r, -1, s); }  // `r *= Constant::Scalar(s)`
  // Constant2!(op, result, c1, c2) — ops.rs:118-148
  int equals_bin(BOp op, const Shape& r, int a, int b) {
    if (!r.m) return emit(binary_iop(op), r, avg(a), avg(b));
    if (sh(a).m) check_dims(r, sh(a), "m1->height must equal result->height");
    if (sh(b).m) check_dims(r, sh(b), "m2->height must equal result->height");
    return emit(binary_iop(op), r, a, b);  // fill | broadcast | broadcast_rev | elemwise
  }
  // Constant::absorb — constant.rs:106-113. Values are immutable here, so a same-kind copy is an alias.
  int absorb(const Shape& r, int o) {
    const Shape os = sh(o);
    if (!r.m && os.m) return avg(o);
    if (r.m && !os.m) return emit(IOp::Copy, r, o);  // fill_matrix(scalar)
    if (r.m) check_dims(r, os, "copy_matrix: dims differ");
    return o;
  }
  // Constant::equals_dot — ops.rs:411-422 + cpu/ops.cpp:88-116 shape checks (generalised, Q7)
  int equals_dot(const Shape& r, int a, bool t1, int b, bool t2) {
    const Shape as = sh(a), bs = sh(b);
    if (!(r.m && as.m && bs.m)) panic("dot should not be used with scalars");
    const unsigned h1 = t1 ? as.w : as.h, inner = t1 ? as.h : as.w;
    const unsigned h2 = t2 ? bs.w : bs.h, w2 = t2 ? bs.h : bs.w;
    if (h1 != r.h) check_fail("height1 must equal result->height");
    if (inner != h2) check_fail("inner_dim must equal height2");
    if (w2 != r.w) check_fail("width2 must equal result->width");
    Instr i;
    i.op = IOp::Gemm;
    i.dst = new_val(r);
    i.a = a;
    i.b = b;
    i.ta = t1;
    i.tb = t2;
    p.tape.push_back(i);
    return i.dst;
  }

  // ---- forward: assign_values, optimize.rs:84-138, memoised per shared body ----------------------------------
  std::vector<Function*> order;  // first-evaluation (post-)order of the non-leaf nodes

  void forward(Function& f) {
    Expr& e = *f.body;
    if (is_leaf(e.kind)) {
      leaf_val(f);
      return;
    }
    if (node_val.count(&e)) return;
    if (e.f1) forward(*e.f1);
    if (e.f2) forward(*e.f2);
    const Shape r = shape_of(f);
    int v = -1;
    switch (e.kind) {
      case Kind::Neg: v = equals_unary(U_NEG, r, read(*e.f1)); break;
      case Kind::Sq: v = equals_unary(U_SQ, r, read(*e.f1)); break;
      case Kind::Abs: v = equals_unary(U_ABS, r, read(*e.f1)); break;
      case Kind::Signum: v = equals_unary(U_SIGNUM, r, read(*e.f1)); break;
      case Kind::Sigmoid: v = equals_unary(U_SIGMOID, r, read(*e.f1)); break;
      case Kind::Tanh: v = equals_unary(U_TANH, r, read(*e.f1)); break;
      case Kind::Add: v = equals_bin(B_ADD, r, read(*e.f1), read(*e.f2)); break;
      case Kind::Sub: v = equals_bin(B_SUB, r, read(*e.f1), read(*e.f2)); break;
      case Kind::Mul: v = equals_bin(B_MUL, r, read(*e.f1), read(*e.f2)); break;
      case Kind::Dot: v = equals_dot(r, read(*e.f1), e.t1, read(*e.f2), e.t2); break;
      default: throw Error("internal: unexpected node kind");
    }
    node_val[&e] = v;
    order.push_back(&f);
  }

  // ---- Param arm of backprop, optimize.rs:149-152 ----------------------------------------------------------------
  void sgd_leaf(Function& f, int error) {
    if (fix_tied_params) {
      Tie* t = nullptr;
      for (auto& kv : ties)
        if (kv.first == f.body->name) t = &kv.second;
      if (!t) {
        ties.emplace_back(f.body->name, Tie{});
        t = &ties.back().second;
      }
      if (std::find(t->leaves.begin(), t->leaves.end(), &f) == t->leaves.end()) t->leaves.push_back(&f);
      const int part = absorb(shape_of(f), error);  // in the parameter's shape (S <- M: mean, or sum when that is fixed too)
      t->grad = t->grad < 0 ? part : emit(IOp::Add, sh(t->grad), t->grad, part);
      return;
    }
    int e2 = mul_imm(error, lr);                  // *error *= Constant::Scalar(learn_rate)
    int cur = leaf_val(f);
    int nv = new_fixed(sh(cur), f.value.buf);     // next version lives in the same leaf buffer
    op_assign(B_SUB, cur, e2, nv);                // self.value -= error   (scalar -= matrix uses avg, Q5)
    leaf_cur[&f] = nv;
  }

  // One node's backward rule (optimize.rs:153-200). Returns the errors handed to f1 / f2.
  void node_backward(Function& f, int error, int& to_f1, int& to_f2) {
    Expr& e = *f.body;
    const Shape r = shape_of(f);  // placeholders are shaped like the node's value (function.rs:87-98)
    to_f1 = to_f2 = -1;
    switch (e.kind) {
      case Kind::Neg: to_f1 = assign_unary(U_NEG, error); break;
      case Kind::Sq: {
        int t = mul_imm(error, 2.0f);
        to_f1 = op_assign(B_MUL, t, read(*e.f1));
      } break;
      case Kind::Abs: {
        int ph = equals_unary(U_SIGNUM, r, read(*e.f1));
        to_f1 = op_assign(B_MUL, ph, error);
      } break;
      case Kind::Signum: panic("sign is not differentiable");
      case Kind::Sigmoid: {
        int v = read(f);
        int ph = equals_unary(U_ONE_MINUS, r, v);
        ph = op_assign(B_MUL, ph, v);
        // Q3: the child receives the local derivative; `error *= ph` is dead (optimize.rs:168-173)
        to_f1 = fix_act_grad ? op_assign(B_MUL, error, ph) : ph;
      } break;
      case Kind::Tanh: {
        int v = read(f);
        int ph = equals_unary(U_SQ, r, v);
        ph = assign_unary(U_ONE_MINUS, ph);
        to_f1 = fix_act_grad ? op_assign(B_MUL, error, ph) : ph;
      } break;
      case Kind::Add:
        to_f2 = absorb(r, error);
        to_f1 = error;
        break;
      case Kind::Sub:
        to_f2 = equals_unary(U_NEG, r, error);
        to_f1 = error;
        break;
      case Kind::Mul:
        to_f2 = equals_bin(B_MUL, r, read(*e.f1), error);
        to_f1 = op_assign(B_MUL, error, read(*e.f2));
        break;
      case Kind::Dot:
        // placeholders sized for the gradients they hold (dims of f1 / f2); the reference sizes both like
        // the result (ops.rs:405), which coincides for its square-only products (Q7)
        to_f1 = equals_dot(shape_of(*e.f1), error, false, read(*e.f2), !e.t2);
        to_f2 = equals_dot(shape_of(*e.f2), read(*e.f1), !e.t1, error, false);
        if (dp && p.mode == 1 && e.f2->has_params) {
          // Data parallel (dag mode only — exact mode and C1-C3 run as replicas, its in-line SGD is
          // interleaved with the walk): the rows of every left operand are the sharded batch dimension, so the
          // right-operand (weight) gradient X^T.E is a per-rank partial sum: sum-allreduce it, in place.
          Instr ar;
          ar.op = IOp::AllReduce;
          ar.a = to_f2;
          ar.dst = new_val(sh(to_f2));
          p.vals[ar.dst].alias_of = to_f2;
          p.tape.push_back(ar);
          to_f2 = ar.dst;
        }
        break;
      default: break;
    }
  }

  // fix_tied_params: the one update per name, applied to every copy, after the walk
  void apply_tied_updates() {
    for (auto& kv : ties) {
      Tie& t = kv.second;
      if (t.grad < 0) continue;
      const int e2 = mul_imm(t.grad, lr);
      for (Function* f : t.leaves) {
        const int cur = leaf_val(*f);
        const int nv = new_fixed(sh(cur), f->value.buf);
        op_assign(B_SUB, cur, e2, nv);
        leaf_cur[f] = nv;
      }
    }
  }

  // exact mode: the literal pre-order recursion of optimize.rs:140-204
  void backprop_tree(Function& f, int error) {
    if (!f.has_params) return;  // :147
    Expr& e = *f.body;
    if (e.kind == Kind::Param) return sgd_leaf(f, error);
    if (is_leaf(e.kind)) return;
    int e1, e2;
    node_backward(f, error, e1, e2);
    if (e.f1) backprop_tree(*e.f1, e1);
    if (e.f2) backprop_tree(*e.f2, e2);
  }

  // dag mode
  std::unordered_map<const Expr*, int> node_err;
  void dag_deliver(Function& child, int g) {
    if (!child.has_params) return;
    Expr& ce = *child.body;
    if (ce.kind == Kind::Param) return sgd_leaf(child, g);
    if (is_leaf(ce.kind)) return;
    auto it = node_err.find(&ce);
    if (it == node_err.end()) node_err[&ce] = g;           // first contribution
    else it->second = op_assign(B_ADD, it->second, g);     // f32 sum in consumer-visit order
  }
  void backprop_dag(Function& root, int error) {
    if (!root.has_params) return;
    Expr& re = *root.body;
    if (re.kind == Kind::Param) return sgd_leaf(root, error);
    if (is_leaf(re.kind)) return;
    node_err[&re] = error;
    for (int idx = (int)order.size() - 1; idx >= 0; idx--) {
      Function& f = *order[idx];
      auto it = node_err.find(f.body.get());
      if (it == node_err.end()) continue;
      if (f.has_params) {
        int e1, e2;
        node_backward(f, it->second, e1, e2);
        Expr& e = *f.body;
        if (e.f1) dag_deliver(*e.f1, e1);
        if (e.f2) dag_deliver(*e.f2, e2);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------
// passes
// ------------------------------------------------------------------------------------------------------
void for_each_use(const Instr& i, const std::function<void(int)>& fn) {
  if (i.a >= 0) fn(i.a);
  if (i.b >= 0) fn(i.b);
  if (i.c_in >= 0) fn(i.c_in);
}

void dead_code_elimination(Plan& p) {
  std::vector<char> live(p.vals.size(), 0);
  for (size_t v = 0; v < p.vals.size(); v++) live[v] = p.vals[v].persistent;
  for (int k = (int)p.tape.size() - 1; k >= 0; k--) {
    Instr& i = p.tape[k];
    if (!live[i.dst]) {
      i.dead = true;
      continue;
    }
    for_each_use(i, [&](int v) { live[v] = 1; });
  }
}

void count_uses(Plan& p) {
  for (auto& v : p.vals) v.nuses = 0;
  for (auto& i : p.tape)
    if (!i.dead) for_each_use(i, [&](int v) { p.vals[v].nuses++; });
}

// dW product + SGD update:  g = A^T.E ; e = g * lr ; W' = W - e   ==>   W' = W - (A^T.E) * lr in the gemm
// epilogue (two roundings, as the reference: optimize.rs:149-152). The update is hoisted to the gemm's
// position, which is legal when nothing in between reads the old version of W.
void fuse_gemm_sgd(Plan& p) {
  count_uses(p);
  std::vector<int> user(p.vals.size(), -1);  // the single user of a value (valid when nuses == 1)
  for (size_t k = 0; k < p.tape.size(); k++)
    if (!p.tape[k].dead) for_each_use(p.tape[k], [&](int v) { user[v] = (int)k; });
  for (size_t k = 0; k < p.tape.size(); k++) {
    Instr& g = p.tape[k];
    if (g.dead || g.op != IOp::Gemm || g.epi != 0) continue;
    Val& gv = p.vals[g.dst];
    if (gv.persistent || gv.nuses != 1) continue;
    Instr& mu = p.tape[user[g.dst]];
    if (mu.op != IOp::MulI || mu.a != g.dst || p.vals[mu.dst].nuses != 1 || p.vals[mu.dst].persistent) continue;
    const int sub_idx = user[mu.dst];
    Instr& sb = p.tape[sub_idx];
    if (sb.op != IOp::Sub || sb.b != mu.dst || sb.a == mu.dst) continue;
    const Val &oldv = p.vals[sb.a], &newv = p.vals[sb.dst];
    if (!oldv.fixed || oldv.fixed != newv.fixed || !oldv.sh.same(gv.sh) || !gv.sh.m) continue;
    bool blocked = false;
    for (int q = (int)k + 1; q < sub_idx && !blocked; q++) {
      const Instr& o = p.tape[q];
      if (o.dead) continue;
      for_each_use(o, [&](int v) {
        if (p.vals[v].fixed == oldv.fixed) blocked = true;  // someone still reads (a version of) this leaf
      });
      if (p.vals[o.dst].fixed == oldv.fixed) blocked = true;
    }
    if (blocked) continue;
    g.epi = 1;
    g.c_in = sb.a;
    g.imm = mu.imm;
    g.dst = sb.dst;
    mu.dead = true;
    sb.dead = true;
    p.stats.gemms_fused_sgd++;
  }
  count_uses(p);
}

// Data parallel: g = allreduce(X^T.E) ; e = g * lr ; W' = W - e  ==>  one DpUpdate (dp_kernels.cu): the kernel sums the
// per-rank partial gradients over NVLink peer memory, applies the update once per element and stores the new weight
// into every rank's copy. The update is hoisted to the all-reduce's position (same legality rule as fuse_gemm_sgd).
void fuse_dp_update(Plan& p) {
  if (!dp_fused_available()) return;
  count_uses(p);
  std::vector<int> user(p.vals.size(), -1);
  for (size_t k = 0; k < p.tape.size(); k++)
    if (!p.tape[k].dead) for_each_use(p.tape[k], [&](int v) { user[v] = (int)k; });
  for (size_t k = 0; k < p.tape.size(); k++) {
    Instr& ar = p.tape[k];
    if (ar.dead || ar.op != IOp::AllReduce) continue;
    Val& gv = p.vals[ar.dst];
    if (gv.persistent || gv.nuses != 1 || !gv.sh.m || gv.sh.n() % 4 || gv.sh.n() < dp_fused_min_elems()) continue;
    Instr& mu = p.tape[user[ar.dst]];
    if (mu.op != IOp::MulI || mu.a != ar.dst || p.vals[mu.dst].nuses != 1 || p.vals[mu.dst].persistent) continue;
    const int sub_idx = user[mu.dst];
    Instr& sb = p.tape[sub_idx];
    if (sb.op != IOp::Sub || sb.b != mu.dst || sb.a == mu.dst) continue;
    const Val &oldv = p.vals[sb.a], &newv = p.vals[sb.dst];
    if (!oldv.fixed || oldv.fixed != newv.fixed || !oldv.sh.same(gv.sh)) continue;
    bool blocked = false;
    for (int q = (int)k + 1; q < sub_idx && !blocked; q++) {
      const Instr& o = p.tape[q];
      if (o.dead) continue;
      for_each_use(o, [&](int v) {
        if (p.vals[v].fixed == oldv.fixed) blocked = true;
      });
      if (p.vals[o.dst].fixed == oldv.fixed) blocked = true;
    }
    if (blocked) continue;
    ar.op = IOp::DpUpdate;
    ar.c_in = sb.a;
    ar.imm = mu.imm;
    p.vals[ar.dst].alias_of = -1;  // the summed gradient is never materialised
    ar.dst = sb.dst;
    mu.dead = true;
    sb.dead = true;
    p.stats.dp_fused_updates++;
  }
  count_uses(p);
}

// dX accumulation in dag mode:  g = E.W^T ; err' = err + g  ==>  err' = err + E.W^T in the gemm epilogue.
void fuse_gemm_accumulate(Plan& p) {
  count_uses(p);
  std::vector<int> user(p.vals.size(), -1), def_idx(p.vals.size(), -1);
  for (size_t k = 0; k < p.tape.size(); k++)
    if (!p.tape[k].dead) {
      for_each_use(p.tape[k], [&](int v) { user[v] = (int)k; });
      def_idx[p.tape[k].dst] = (int)k;
    }
  for (size_t k = 0; k < p.tape.size(); k++) {
    Instr& g = p.tape[k];
    if (g.dead || g.op != IOp::Gemm || g.epi != 0) continue;
    Val& gv = p.vals[g.dst];
    if (gv.persistent || gv.nuses != 1) continue;
    const int add_idx = user[g.dst];
    Instr& ad = p.tape[add_idx];
    if (ad.op != IOp::Add || ad.b != g.dst || ad.a == g.dst) continue;
    Val& acc = p.vals[ad.a];
    if (acc.fixed || acc.persistent || acc.nuses != 1 || !acc.sh.same(gv.sh) || p.vals[ad.dst].persistent) continue;
    // the accumulator must already exist when the gemm runs: its definition precedes the gemm
    if (def_idx[ad.a] < 0 || def_idx[ad.a] >= (int)k || ad.a == g.a || ad.a == g.b) continue;
    g.epi = 2;
    g.c_in = ad.a;
    g.dst = ad.dst;
    p.vals[ad.dst].alias_of = ad.a;
    ad.dead = true;
  }
  count_uses(p);
}

// Hoisting. The reference's backward pass hands the child of a Sigmoid / Tanh only the local derivative (Q3,
// optimize.rs:168-179), so most of "backward" — every gate's derivative, the weight-gradient products X^T.ph and
// their SGD updates — depends on forward values alone, not on the loss. This pass moves every backward instruction
// to just behind the last forward instruction it (transitively) depends on; the loss-carrying chain stays where it
// was. Pure dataflow reordering: every instruction computes the same bits. What it buys: derivative ops land next
// to the forward ops of the same shape (one elementwise kernel instead of two, one HBM pass less), activations die
// early (smaller arena), and under data parallelism the weight exchanges are spread over the whole iteration
// instead of queueing up behind the forward pass, where the wire then sits idle.
// Dependencies are RAW / WAR / WAW on memory identity: SSA value, or the leaf buffer its versions share, or the
// value an in-place result aliases.
void hoist_backward(Plan& p, size_t n_forward) {
  const size_t N = p.tape.size();
  auto mem = [&](int v) -> long long {
    while (p.vals[v].alias_of >= 0) v = p.vals[v].alias_of;
    return p.vals[v].fixed ? (long long)(intptr_t)p.vals[v].fixed : -(long long)(v + 1);
  };
  struct Loc {
    int last_writer = -1;
    std::vector<int> readers;
  };
  std::unordered_map<long long, Loc> loc;
  std::vector<long long> key(N, 0);
  std::vector<std::vector<int>> deps(N);
  std::vector<int> live;
  for (size_t i = 0; i < N; i++) {
    const Instr& in = p.tape[i];
    if (in.dead) continue;
    live.push_back((int)i);
    auto dep = [&](int d) {
      if (d >= 0 && d != (int)i) deps[i].push_back(d);
    };
    std::vector<long long> reads;
    for_each_use(in, [&](int v) { reads.push_back(mem(v)); });
    const long long w = mem(in.dst);
    for (long long r : reads) dep(loc[r].last_writer);
    {
      Loc& l = loc[w];
      dep(l.last_writer);
      for (int rd : l.readers) dep(rd);
    }
    if (i < n_forward) {
      key[i] = (long long)i;  // the forward pass is the anchor sequence
    } else {
      key[i] = (long long)n_forward - 1;  // no dependency at all (the root error, a Fill): start of backward
      bool first = true;
      for (int d : deps[i]) {
        key[i] = first ? key[d] : std::max(key[i], key[d]);
        first = false;
      }
    }
    for (long long r : reads) loc[r].readers.push_back((int)i);
    Loc& l = loc[w];
    l.last_writer = (int)i;
    l.readers.clear();
  }
  std::stable_sort(live.begin(), live.end(), [&](int a, int b) { return key[a] < key[b]; });
  // the new order must be a topological order of the same dependencies
  std::vector<int> pos(N, -1);
  for (size_t k = 0; k < live.size(); k++) pos[live[k]] = (int)k;
  for (int i : live)
    for (int d : deps[i])
      if (pos[d] < 0 || pos[d] >= pos[i]) throw Error("internal: hoisting broke a dependency");
  std::vector<Instr> out;
  out.reserve(live.size());
  for (int i : live) out.push_back(p.tape[i]);
  p.tape.swap(out);
}

// ------------------------------------------------------------------------------------------------------
// Elementwise fusion: compile tape[s, e) into one register-bytecode program, or fail.
// ------------------------------------------------------------------------------------------------------
struct GroupCompiler {
  const Plan& p;
  const std::vector<int>& last_use_idx;  // tape index of the last live use of each value (-1: none)
  GroupCompiler(const Plan& plan, const std::vector<int>& lu) : p(plan), last_use_idx(lu) {}

  static EwOpcode opcode(IOp op) {
    switch (op) {
      case IOp::Neg: return OP_NEG;
      case IOp::Sq: return OP_SQ;
      case IOp::AbsM: return OP_ABS_M;
      case IOp::SgnM: return OP_SGN_M;
      case IOp::AbsS: return OP_ABS_S;
      case IOp::SgnS: return OP_SGN_S;
      case IOp::Sigmoid: return OP_SIGMOID;
      case IOp::Tanh: return OP_TANH;
      case IOp::OneMinus: return OP_ONE_MINUS;
      case IOp::Add: return OP_ADD;
      case IOp::Sub: return OP_SUB;
      case IOp::Mul: return OP_MUL;
      default: return OP_END;
    }
  }

  bool compile(const std::vector<int>& idx, EwGroup& g) const {
    // idx: tape indices (live, elementwise, same domain) in order
    g = EwGroup();
    const Instr& first = p.tape[idx[0]];
    g.n = p.vals[first.dst].sh.n();
    std::unordered_map<int, int> reg;       // value -> register
    std::unordered_map<int, int> uses_left; // in-group uses remaining
    std::unordered_set<int> defined;
    const int group_end = idx.back();
    for (int k : idx) {
      const Instr& i = p.tape[k];
      for_each_use(i, [&](int v) { uses_left[v]++; });
    }
    // inputs: used before defined in the group, in order of first use
    {
      std::unordered_set<int> def;
      for (int k : idx) {
        const Instr& i = p.tape[k];
        for_each_use(i, [&](int v) {
          if (!def.count(v) && !reg.count(v)) {
            reg[v] = (int)g.in.size();
            g.in.push_back(v);
          }
        });
        def.insert(i.dst);
      }
    }
    if (g.in.size() > (size_t)EW_MAX_IN) return false;
    std::vector<char> reg_busy(EW_REGS, 0);
    for (size_t r = 0; r < g.in.size(); r++) reg_busy[r] = 1;
    auto alloc_reg = [&]() {
      for (int r = 0; r < EW_REGS; r++)
        if (!reg_busy[r]) {
          reg_busy[r] = 1;
          return r;
        }
      return -1;
    };
    auto imm_index = [&](float x) {
      for (size_t q = 0; q < g.imm.size(); q++)
        if (memcmp(&g.imm[q], &x, sizeof(float)) == 0) return (int)q;
      g.imm.push_back(x);
      return (int)g.imm.size() - 1;
    };
    auto release_use = [&](int v) {
      if (--uses_left[v] == 0) {
        auto it = reg.find(v);
        if (it != reg.end()) {
          reg_busy[it->second] = 0;
          reg.erase(it);
        }
      }
    };
    int acc = -1;  // value currently held by the accumulator
    auto load_acc = [&](int v) -> bool {
      if (acc == v) return true;
      auto it = reg.find(v);
      if (it == reg.end()) return false;
      g.code.push_back(ew_encode(OP_LDR, it->second));
      acc = v;
      return true;
    };
    for (size_t q = 0; q < idx.size(); q++) {
      const Instr& i = p.tape[idx[q]];
      switch (i.op) {
        case IOp::Fill: g.code.push_back(ew_encode(OP_LDI, imm_index(i.imm))); break;
        case IOp::Copy:
          if (!load_acc(i.a)) return false;
          release_use(i.a);
          break;
        case IOp::MulI: {
          if (!load_acc(i.a)) return false;
          int k = imm_index(i.imm);
          g.code.push_back(ew_encode(OP_MULI, k));
          release_use(i.a);
        } break;
        case IOp::Add: case IOp::Sub: case IOp::Mul: {
          EwOpcode oc = opcode(i.op);
          if (i.a == i.b && i.op == IOp::Mul && (acc == i.a || reg.count(i.a))) {
            if (!load_acc(i.a)) return false;
            g.code.push_back(ew_encode(OP_SQ));
          } else if (reg.count(i.b)) {
            if (!load_acc(i.a)) return false;
            g.code.push_back(ew_encode(oc, reg[i.b]));
          } else if (acc == i.b && reg.count(i.a)) {
            // b lives only in the accumulator: commute, or reverse-subtract
            g.code.push_back(ew_encode(i.op == IOp::Sub ? OP_RSUB : oc, reg[i.a]));
          } else {
            return false;
          }
          release_use(i.a);
          release_use(i.b);
        } break;
        default:
          if (!load_acc(i.a)) return false;
          g.code.push_back(ew_encode(opcode(i.op)));
          release_use(i.a);
          break;
      }
      acc = i.dst;
      if (g.imm.size() > (size_t)EW_MAX_IMM) return false;
      const Val& dv = p.vals[i.dst];
      // memory: persistent, or read after the group ends
      if (dv.persistent || last_use_idx[i.dst] > group_end) {
        if (g.out.size() >= (size_t)EW_MAX_OUT) return false;
        g.code.push_back(ew_encode(OP_STG, (int)g.out.size()));
        g.out.push_back(i.dst);
      }
      // register: needed unless the only in-group use is the next instruction's accumulator operand
      int ul = uses_left.count(i.dst) ? uses_left[i.dst] : 0;
      if (ul > 0) {
        bool acc_only = false;
        if (ul == 1 && q + 1 < idx.size()) {
          const Instr& nx = p.tape[idx[q + 1]];
          int cnt = (nx.a == i.dst) + (nx.b == i.dst);
          if (cnt == 1) {
            if (nx.a == i.dst) {
              // first operand: fine when the second (if any) is in a register
              acc_only = (nx.b < 0) || reg.count(nx.b);
            } else {
              acc_only = reg.count(nx.a) > 0;  // commuted / reverse-subtract form
            }
          }
        }
        if (!acc_only) {
          int r = alloc_reg();
          if (r < 0) return false;
          reg[i.dst] = r;
          g.code.push_back(ew_encode(OP_STR, r));
        }
      }
      if (g.code.size() > (size_t)EW_MAX_INS) return false;
    }
    return !g.code.empty() || !g.out.empty();
  }
};

void build_steps(Plan& p) {
  // last live use of every value, as a tape index
  std::vector<int> last_use(p.vals.size(), -1);
  for (size_t k = 0; k < p.tape.size(); k++)
    if (!p.tape[k].dead) for_each_use(p.tape[k], [&](int v) { last_use[v] = (int)k; });
  GroupCompiler gc(p, last_use);
  const size_t max_group = p.fuse ? 96 : 1;
  size_t k = 0;
  const size_t N = p.tape.size();
  while (k < N) {
    const Instr& i = p.tape[k];
    if (i.dead) {
      k++;
      continue;
    }
    p.stats.live_instrs++;
    if (!is_ew(i.op)) {
      Step s;
      s.kind = i.op == IOp::Gemm ? Step::GEMM : i.op == IOp::Avg ? Step::AVG : i.op == IOp::DpUpdate ? Step::DPUPDATE : Step::ALLREDUCE;
      s.ins = i;
      p.steps.push_back(std::move(s));
      k++;
      continue;
    }
    // grow an elementwise group: consecutive live EW instrs with the same domain
    const Shape dom = p.vals[i.dst].sh;
    std::vector<int> idx{(int)k};
    EwGroup best;
    if (!gc.compile(idx, best)) throw Error("internal: single elementwise instruction does not compile");
    size_t next = k + 1;
    while (next < N && idx.size() < max_group) {
      const Instr& c = p.tape[next];
      if (c.dead) {
        next++;
        continue;
      }
      if (!is_ew(c.op) || !p.vals[c.dst].sh.same(dom)) break;
      idx.push_back((int)next);
      EwGroup trial;
      if (!gc.compile(idx, trial)) {
        idx.pop_back();
        break;
      }
      best = std::move(trial);
      next++;
    }
    p.stats.live_instrs += idx.size() - 1;
    Step s;
    s.kind = Step::EW;
    s.ew = std::move(best);
    p.steps.push_back(std::move(s));
    k = (size_t)idx.back() + 1;
  }
}

// ------------------------------------------------------------------------------------------------------
// Device arena: every value that crosses a kernel boundary gets a slot; slots are recycled by liveness.
// ------------------------------------------------------------------------------------------------------
void step_defs_uses(const Step& s, std::vector<int>& defs, std::vector<int>& uses) {
  defs.clear();
  uses.clear();
  if (s.kind == Step::EW) {
    defs = s.ew.out;
    uses = s.ew.in;
  } else {
    defs.push_back(s.ins.dst);
    for_each_use(s.ins, [&](int v) { uses.push_back(v); });
  }
}

void assign_memory(Plan& p) {
  std::vector<int> defs, uses;
  for (size_t si = 0; si < p.steps.size(); si++) {
    step_defs_uses(p.steps[si], defs, uses);
    for (int v : defs) {
      p.vals[v].def_step = (int)si;
      p.vals[v].needs_mem = true;
      if (p.vals[v].last_step < (int)si) p.vals[v].last_step = (int)si;
    }
    for (int v : uses) p.vals[v].last_step = (int)si;
  }
  // alias chains extend the life of the underlying slot
  auto root_of = [&](int v) {
    while (p.vals[v].alias_of >= 0) v = p.vals[v].alias_of;
    return v;
  };
  for (size_t v = 0; v < p.vals.size(); v++)
    if (p.vals[v].alias_of >= 0 && p.vals[v].needs_mem) {
      int r = root_of((int)v);
      p.vals[r].last_step = std::max(p.vals[r].last_step, p.vals[v].last_step);
    }
  // Slots that go through an all-reduce are held back for a while after their last use: recycling a gradient buffer
  // at once would make the next weight-gradient product wait (WAR) for the previous collective and its update, and
  // chain product -> all-reduce -> update -> product with no overlap at all.
  std::vector<char> reduced(p.vals.size(), 0);
  for (const Step& s : p.steps)
    if (s.kind == Step::ALLREDUCE) reduced[root_of(s.ins.dst)] = 1;
    else if (s.kind == Step::DPUPDATE) reduced[root_of(s.ins.a)] = 1;
  const int hold = 64;  // steps (a backward LSTM step is ~25 kernels: about 20 gradient buffers in rotation)
  constexpr size_t ALIGN = 256;
  // In-switch exchange (dp_multimem_update_kernel): the partial gradients it sums must sit in the symmetric multicast
  // heap, at the same offset on every rank — they get an arena of their own there (index 1).
  const bool mc = dp_multimem_active();
  std::vector<char> in_mc(p.vals.size(), 0);
  if (mc)
    for (const Step& s : p.steps)
      if (s.kind == Step::DPUPDATE) in_mc[root_of(s.ins.a)] = 1;
  std::map<size_t, std::vector<size_t>> free_slots_of[2];  // bytes -> offsets
  std::vector<std::vector<int>> release_at(p.steps.size());
  size_t top_of[2] = {0, 0};
  std::vector<size_t> off(p.vals.size(), SIZE_MAX);
  for (size_t si = 0; si < p.steps.size(); si++) {
    step_defs_uses(p.steps[si], defs, uses);
    for (int v : defs) {
      Val& val = p.vals[v];
      if (val.fixed || val.alias_of >= 0) continue;
      size_t bytes = (val.sh.n() * sizeof(float) + ALIGN - 1) / ALIGN * ALIGN;
      auto& fl = free_slots_of[(int)in_mc[v]][bytes];
      size_t& top = top_of[(int)in_mc[v]];
      if (!fl.empty()) {
        off[v] = fl.back();
        fl.pop_back();
      } else {
        off[v] = top;
        top += bytes;
      }
      int rel = val.last_step;
      if (reduced[v]) rel = std::min<int>(rel + hold, (int)p.steps.size() - 1);
      release_at[rel].push_back(v);
    }
    for (int v : release_at[si]) {
      size_t bytes = (p.vals[v].sh.n() * sizeof(float) + ALIGN - 1) / ALIGN * ALIGN;
      free_slots_of[(int)in_mc[v]][bytes].push_back(off[v]);
    }
  }
  p.stats.arena_bytes = top_of[0] + top_of[1];
  if (top_of[0]) p.arena = DevBuf::alloc(top_of[0] / sizeof(float));
  if (top_of[1]) p.dp_arena = DevBuf::alloc_symm(top_of[1] / sizeof(float));
  for (size_t v = 0; v < p.vals.size(); v++) {
    Val& val = p.vals[v];
    if (val.fixed) val.ptr = val.fixed;
    else if (off[v] != SIZE_MAX) val.ptr = (in_mc[v] ? p.dp_arena->ptr : p.arena->ptr) + off[v] / sizeof(float);
  }
  for (size_t v = 0; v < p.vals.size(); v++) {
    Val& val = p.vals[v];
    if (!val.fixed && val.alias_of >= 0) val.ptr = p.vals[root_of((int)v)].ptr;
  }
}

void finalize_steps(Plan& p) {
  for (Step& s : p.steps) {
    if (s.kind == Step::EW) {
      EwProgram& pr = s.prog;
      memset(&pr, 0, sizeof pr);
      pr.n = s.ew.n;
      pr.n_in = (uint32_t)s.ew.in.size();
      pr.n_out = (uint32_t)s.ew.out.size();
      pr.n_ins = (uint32_t)s.ew.code.size();
      for (size_t k = 0; k < s.ew.in.size(); k++) {
        const Val& v = p.vals[s.ew.in[k]];
        if (!v.ptr) throw Error("internal: elementwise input without storage");
        pr.in[k] = v.ptr;
        if (!v.sh.m) pr.in_scalar_mask |= 1u << k;  // scalar operand: broadcast
        if (!(pr.in_scalar_mask >> k & 1u)) p.stats.ew_bytes_per_iter += 4.0 * (double)pr.n;
      }
      for (size_t k = 0; k < s.ew.out.size(); k++) {
        const Val& v = p.vals[s.ew.out[k]];
        if (!v.ptr) throw Error("internal: elementwise output without storage");
        pr.out[k] = v.ptr;
        p.stats.ew_bytes_per_iter += 4.0 * (double)pr.n;
      }
      for (size_t k = 0; k < s.ew.imm.size(); k++) pr.imm[k] = s.ew.imm[k];
      for (size_t k = 0; k < s.ew.code.size(); k++) pr.code[k] = s.ew.code[k];
      p.stats.ew_groups++;
      if (p.ew_jit_min >= 0 && (long long)pr.n >= p.ew_jit_min) {
        s.jit = ew_jit_get(pr);
        if (s.jit) p.stats.ew_specialised++;
      }
    } else if (s.kind == Step::GEMM) {
      const Shape &as = p.vals[s.ins.a].sh, &cs = p.vals[s.ins.dst].sh;
      const double k = s.ins.ta ? as.h : as.w;
      p.stats.gemm_flop_per_iter += 2.0 * cs.h * cs.w * k;
      p.stats.gemms++;
    } else if (s.kind == Step::AVG) {
      p.stats.reduces++;
    } else if (s.kind == Step::ALLREDUCE) {
      p.stats.allreduces++;
    }
  }
  p.stats.kernels_per_iter = p.steps.size();
}

// Data parallel: every buffer a fused exchange touches (the partial gradient in the arena, the weight leaf) is opened
// by every peer, and the peer pointers go into the step. Collective: every rank builds the same plan at the same time.
// In-switch form: the gradient slot and the weight must be heap memory at the same (segment, offset) on every rank —
// checked here, collectively, so that every rank takes the same decision.
void bind_dp_multimem(Plan& p) {
  std::vector<unsigned long long> where;
  bool local_ok = true;
  for (const Step& s : p.steps) {
    if (s.kind != Step::DPUPDATE) continue;
    for (const float* ptr : {p.vals[s.ins.a].ptr, p.vals[s.ins.dst].ptr}) {
      unsigned long long seg = ~0ull, off = ~0ull;
      if (!symm_locate(ptr, &seg, &off, nullptr)) local_ok = false;
      where.push_back(seg);
      where.push_back(off);
    }
  }
  if (where.empty()) return;
  where.push_back(local_ok ? 1 : 0);
  std::vector<unsigned char> all = dp_allgather_bytes(where.data(), where.size() * sizeof(unsigned long long));
  const size_t stride = where.size() * sizeof(unsigned long long);
  for (size_t r = 0; r * stride < all.size(); r++)
    if (memcmp(all.data() + r * stride, where.data(), stride) != 0 || !local_ok)
      throw Error("data parallel: the ranks' weights and gradient slots are not at the same offsets of the multicast heap "
                  "(ranks must build the same graphs in the same order); cg_dp_set_multimem(0) selects the peer-memory exchange");
  for (Step& s : p.steps) {
    if (s.kind != Step::DPUPDATE) continue;
    if (p.vals[s.ins.c_in].ptr != p.vals[s.ins.dst].ptr) throw Error("internal: fused data-parallel update is not in place");
    void *g_mc = nullptr, *w_mc = nullptr;
    symm_locate(p.vals[s.ins.a].ptr, nullptr, nullptr, &g_mc);
    symm_locate(p.vals[s.ins.dst].ptr, nullptr, nullptr, &w_mc);
    memset(&s.mc, 0, sizeof s.mc);
    s.mc.g_mc = static_cast<const float*>(g_mc);
    s.mc.w_uc = p.vals[s.ins.dst].ptr;
    s.mc.w_mc = static_cast<float*>(w_mc);
    s.mc.lr = s.ins.imm;
    s.mc.n = p.vals[s.ins.dst].sh.n();
    dp_fill_mc_flags(s.mc);
    s.use_mc = true;
    p.stats.dp_multimem_updates++;
  }
}

void bind_dp_updates(Plan& p) {
  if (dp_multimem_active()) {
    if (!plan_only()) return bind_dp_multimem(p);
    for (Step& s : p.steps)
      if (s.kind == Step::DPUPDATE) s.use_mc = true, p.stats.dp_multimem_updates++;
    return;
  }
  std::vector<const void*> ptrs;
  for (const Step& s : p.steps)
    if (s.kind == Step::DPUPDATE) {
      ptrs.push_back(p.vals[s.ins.a].ptr);
      ptrs.push_back(p.vals[s.ins.dst].ptr);
    }
  if (ptrs.empty() || plan_only()) return;
  if (!dp_register_windows(ptrs))
    throw Error("data parallel: the ranks could not open each other's buffers (CUDA IPC); set CG_DP_FUSED=0 to use ncclAllReduce");
  for (Step& s : p.steps) {
    if (s.kind != Step::DPUPDATE) continue;
    if (p.vals[s.ins.c_in].ptr != p.vals[s.ins.dst].ptr) throw Error("internal: fused data-parallel update is not in place");
    void *g[DP_MAX_WORLD], *w[DP_MAX_WORLD];
    if (!dp_peer_pointers(p.vals[s.ins.a].ptr, g) || !dp_peer_pointers(p.vals[s.ins.dst].ptr, w))
      throw Error("internal: data-parallel buffer is not registered");
    memset(&s.dp, 0, sizeof s.dp);
    for (int r = 0; r < DP_MAX_WORLD; r++) {
      s.dp.g[r] = static_cast<const float*>(g[r]);
      s.dp.w[r] = static_cast<float*>(w[r]);
    }
    s.dp.lr = s.ins.imm;
    s.dp.n = p.vals[s.ins.dst].sh.n();
    dp_fill_flags(s.dp);
  }
}

// In-switch exchange: every weight it will update (a matrix parameter that is the right operand of a product — the
// pattern fuse_dp_update recognises) moves into the symmetric multicast heap before the tape takes its address. The
// walk is memoised per shared body and visits the leaves in the deterministic order every rank shares, so the heap
// hands out the same offsets everywhere. Returns whether anything moved.
bool migrate_dp_weights(Function& root) {
  bool moved = false;
  std::unordered_set<const Expr*> seen;
  std::function<void(Function&)> walk = [&](Function& f) {
    Expr& e = *f.body;
    if (is_leaf(e.kind) || !seen.insert(&e).second) return;
    if (e.f1) walk(*e.f1);
    if (e.f2) walk(*e.f2);
    if (e.kind != Kind::Dot || !e.f2 || e.f2->body->kind != Kind::Param) return;
    Value& v = e.f2->value;
    if (!v.is_matrix || !v.buf || v.buf->symm || v.size() % 4 || v.size() < dp_fused_min_elems()) return;
    if (plan_only()) return;
    BufRef nb = DevBuf::alloc_symm(v.size());
    Runtime& rt = Runtime::get();
    CG_CUDA(cudaMemcpyAsync(nb->ptr, v.buf->ptr, v.size() * sizeof(float), cudaMemcpyDeviceToDevice, rt.stream));
    CG_CUDA(cudaStreamSynchronize(rt.stream));  // the old buffer may be freed by the assignment below
    v.buf = nb;
    moved = true;
  };
  walk(root);
  return moved;
}

}  // namespace

void build_vm(Plan& p);

std::shared_ptr<Plan> build_plan(Function& root, float lr, bool keep_root, bool forward_only) {
  Runtime& rt = Runtime::get();
  auto plan = std::make_shared<Plan>();
  plan_registry().push_back(plan);
  Plan& p = *plan;
  p.lr = lr;
  p.mode = rt.mode;
  p.tf32 = rt.tf32;
  p.fix_act_grad = rt.fix_act_grad;
  p.fuse = rt.fuse;
  p.dp_epoch = rt.dp_epoch;
  p.ew_jit_min = rt.ew_jit_min;
  p.hoist = rt.hoist;

  if (p.mode == 1 && !forward_only && dp_multimem_active() && migrate_dp_weights(root)) root.eval_plan.reset();
  Builder b(p, lr, rt.fix_act_grad);
  b.fix_broadcast_sum = rt.fix_broadcast_sum;
  b.fix_tied_params = rt.fix_tied_params;
  b.forward(root);
  const Shape rsh = Builder::shape_of(root);
  p.root_n = rsh.n();
  int root_val = b.read(root);
  if (is_leaf(root.body->kind)) {
    p.root_is_leaf = true;
    p.root_ptr = root.value.buf->ptr;
  } else if (forward_only) {
    // `eval` returns a fresh Constant and leaves the cached value alone (optimize.rs:8-32): a buffer of the plan's own
    p.eval_root = DevBuf::alloc(rsh.n());
    p.vals[root_val].fixed = p.eval_root->ptr;
    p.vals[root_val].keep = p.eval_root;
    p.vals[root_val].persistent = true;
    p.root_ptr = p.eval_root->ptr;
  } else if (keep_root) {
    // the root's cached value must survive the iteration (Q12; print / all_less_than / value readback)
    if (!root.value.buf || root.value.buf->n != rsh.n()) root.value.buf = DevBuf::alloc(rsh.n());
    p.vals[root_val].fixed = root.value.buf->ptr;
    p.vals[root_val].keep = root.value.buf;
    p.vals[root_val].persistent = true;
    p.root_ptr = root.value.buf->ptr;
  }
  const size_t n_forward = p.tape.size();
  // optimize.rs:69 — `let mut error = self.value_mut().copy_and_fill(1.)`
  if (!forward_only) {
    int error = b.emit(IOp::Fill, rsh, -1, -1, 1.0f);
    b.in_backward = true;
    if (p.mode == 0) b.backprop_tree(root, error);
    else b.backprop_dag(root, error);
    if (b.fix_tied_params) b.apply_tied_updates();
  }

  p.stats.tape_instrs = p.tape.size();
  dead_code_elimination(p);
  if (p.fuse) {
    fuse_gemm_sgd(p);
    fuse_dp_update(p);
    fuse_gemm_accumulate(p);
    dead_code_elimination(p);
  }
  if (p.hoist && p.mode == 1) hoist_backward(p, n_forward);
  build_steps(p);
  assign_memory(p);
  finalize_steps(p);
  bind_dp_updates(p);
  if (!keep_root || forward_only) return plan;
  build_vm(p);
  // Elementwise-only graphs whose root store is a visible share of the traffic get a second, "quiet" plan for the
  // iterations whose root value nobody can observe (exec.cpp).
  if (rt.elide_root && !p.root_is_leaf && !p.tape_kernel && !p.vm_steps && p.stats.vm_resident == 0 && p.stats.gemms == 0 &&
      p.stats.allreduces == 0 && p.stats.dp_fused_updates == 0 && 4.0 * (double)p.root_n >= 0.05 * p.stats.ew_bytes_per_iter) {
    p.quiet = build_plan(root, lr, false);
    p.stats.quiet_variant = 1;
    p.stats.quiet_ew_bytes_per_iter = p.quiet->stats.ew_bytes_per_iter;
  }
  return plan;
}

// Latency-bound plans (every matrix a few hundred elements: BASELINE config 3) run on the resident tape executor:
// the whole step list goes to device memory and one CTA walks it, iteration loop included (vm_kernel.cu).
void build_vm(Plan& p) {
  Runtime& rt = Runtime::get();
  if (rt.vm_max_dim <= 0 || !rt.use_graph || p.steps.empty() || p.steps.size() > 8192) return;
  const int max_dim = std::min(rt.vm_max_dim, rt.strict_gemm_max);  // products must be the strict kernel's on both executors
  const size_t max_elems = (size_t)rt.vm_max_dim * rt.vm_max_dim;
  std::vector<VmStep> steps;
  std::vector<EwProgram> progs;
  int dim_seen = 0;
  size_t elems_seen = 0;
  for (const Step& s : p.steps) {
    VmStep v;
    memset(&v, 0, sizeof v);
    if (s.kind == Step::EW) {
      if (s.prog.n > max_elems) return;
      elems_seen = std::max<size_t>(elems_seen, s.prog.n);
      v.kind = VmStep::EW;
      v.prog = (int)progs.size();
      progs.push_back(s.prog);
    } else if (s.kind == Step::AVG) {
      const Val& a = p.vals[s.ins.a];
      if (a.sh.n() > kStrictReduceMax || a.sh.n() > max_elems) return;
      if (s.ins.imm > 0.0f) return;  // a sum instead of a mean (corrected semantics): the CUDA-graph executor has it
      elems_seen = std::max(elems_seen, a.sh.n());
      v.kind = VmStep::AVG;
      v.n = (int)a.sh.n();
      v.a = a.ptr;
      v.c = p.vals[s.ins.dst].ptr;
    } else if (s.kind == Step::GEMM) {
      const Val &a = p.vals[s.ins.a], &b = p.vals[s.ins.b], &c = p.vals[s.ins.dst];
      v.kind = VmStep::GEMM;
      v.m = (int)c.sh.h, v.n = (int)c.sh.w, v.k = (int)(s.ins.ta ? a.sh.h : a.sh.w);
      if (std::max(v.m, std::max(v.n, v.k)) > max_dim) return;
      dim_seen = std::max(dim_seen, std::max(v.m, std::max(v.n, v.k)));
      if (s.ins.epi && p.vals[s.ins.c_in].ptr != c.ptr) throw Error("internal: gemm epilogue operand is not in place");
      v.lda = (int)a.sh.h, v.ldb = (int)b.sh.h, v.ldc = (int)c.sh.h;
      v.ta = s.ins.ta, v.tb = s.ins.tb, v.epi = s.ins.epi;
      v.lr = s.ins.imm;
      v.a = a.ptr, v.b = b.ptr, v.c = c.ptr;
    } else {
      return;  // data-parallel exchanges are not for one CTA
    }
    steps.push_back(v);
  }
  // 1. compiled, shared-memory resident (ew_jit.cpp): whenever the plan's values fit in one SM's shared memory
  if (rt.tape_jit) {
    p.tape_kernel = tape_jit_build(p, !plan_only());
    if (p.tape_kernel) {
      p.stats.vm_resident = 2;
      p.stats.tape_smem_bytes = tape_jit_smem_bytes(p.tape_kernel);
      return;
    }
  }
  // 2. interpreted, values in L1 / L2 (vm_kernel.cu): measured faster than the CUDA graph up to d = 20 on the README
  //    LSTM (155 vs 203 us per iteration), slower from d = 25 on
  if (dim_seen > 20 || elems_seen > 400) return;
  p.stats.vm_resident = 1;
  if (plan_only()) return;
  p.vm_nsteps = (int)steps.size();
  CG_CUDA(cudaMalloc(&p.vm_steps, steps.size() * sizeof(VmStep)));
  CG_CUDA(cudaMemcpy(p.vm_steps, steps.data(), steps.size() * sizeof(VmStep), cudaMemcpyHostToDevice));
  if (!progs.empty()) {
    CG_CUDA(cudaMalloc(&p.vm_progs, progs.size() * sizeof(EwProgram)));
    CG_CUDA(cudaMemcpy(p.vm_progs, progs.data(), progs.size() * sizeof(EwProgram), cudaMemcpyHostToDevice));
  }
}

std::string plan_describe(const Plan& p) {
  std::string out;
  char line[256];
  for (size_t i = 0; i < p.steps.size(); i++) {
    const Step& s = p.steps[i];
    switch (s.kind) {
      case Step::EW:
        snprintf(line, sizeof line, "%zu EW n=%zu in=%zu out=%zu ins=%zu\n", i, s.ew.n, s.ew.in.size(), s.ew.out.size(), s.ew.code.size());
        break;
      case Step::AVG: snprintf(line, sizeof line, "%zu AVG n=%zu\n", i, p.vals[s.ins.a].sh.n()); break;
      case Step::GEMM: {
        const Shape &as = p.vals[s.ins.a].sh, &cs = p.vals[s.ins.dst].sh;
        snprintf(line, sizeof line, "%zu GEMM m=%u n=%u k=%u ta=%d tb=%d epi=%d\n", i, cs.h, cs.w, s.ins.ta ? as.h : as.w, (int)s.ins.ta,
                 (int)s.ins.tb, s.ins.epi);
      } break;
      case Step::ALLREDUCE: snprintf(line, sizeof line, "%zu ALLREDUCE n=%zu\n", i, p.vals[s.ins.dst].sh.n()); break;
      case Step::DPUPDATE: snprintf(line, sizeof line, "%zu DPUPDATE n=%zu%s\n", i, p.vals[s.ins.dst].sh.n(), s.use_mc ? " multimem" : ""); break;
    }
    out += line;
  }
  return out;
}

double plan_step_flop(const Plan& p, size_t i) {
  const Step& s = p.steps[i];
  if (s.kind != Step::GEMM) return 0.0;
  const Shape &as = p.vals[s.ins.a].sh, &cs = p.vals[s.ins.dst].sh;
  return 2.0 * cs.h * cs.w * (double)(s.ins.ta ? as.h : as.w);
}

void plan_launch_iteration(Plan& p, cudaStream_t stream) {
  for (size_t i = 0; i < p.steps.size(); i++) plan_launch_step(p, i, stream);
}

void plan_launch_step(Plan& p, size_t index, cudaStream_t stream) {
  Runtime& rt = Runtime::get();
  Step& s = p.steps[index];
  {
    switch (s.kind) {
      case Step::EW:
        if (s.jit) launch_ew_jit(s.jit, s.prog, stream);
        else launch_ew_program(s.prog, stream);
        break;
      case Step::AVG: {
        const Val& a = p.vals[s.ins.a];
        if (s.ins.imm > 0.0f) launch_reduce_div(a.ptr, a.sh.n(), s.ins.imm, p.vals[s.ins.dst].ptr, rt.red_ws, stream);
        else launch_reduce_avg(a.ptr, a.sh.n(), p.vals[s.ins.dst].ptr, rt.red_ws, stream);
      } break;
      case Step::GEMM: {
        const Val &a = p.vals[s.ins.a], &b = p.vals[s.ins.b], &c = p.vals[s.ins.dst];
        const int m = (int)c.sh.h, n = (int)c.sh.w, k = (int)(s.ins.ta ? a.sh.h : a.sh.w);
        GemmEpilogue epi;
        epi.mode = s.ins.epi;
        epi.lr = s.ins.imm;
        if (s.ins.epi && p.vals[s.ins.c_in].ptr != c.ptr) throw Error("internal: gemm epilogue operand is not in place");
        launch_gemm(a.ptr, s.ins.ta, b.ptr, s.ins.tb, c.ptr, m, n, k, (int)a.sh.h, (int)b.sh.h, (int)c.sh.h, epi,
                    p.tf32, stream);
      } break;
      case Step::ALLREDUCE: dp_allreduce_sum(p.vals[s.ins.dst].ptr, p.vals[s.ins.dst].sh.n(), stream); break;
      case Step::DPUPDATE:
        if (s.use_mc) launch_dp_multimem_update(s.mc, stream);
        else launch_dp_reduce_update(s.dp, stream);
        break;
    }
  }
}

}  // namespace cg
