// exec.cpp — `minimize` (src/optimize.rs:59-72) over a compiled plan: one CUDA-graph replay per iteration.
#include <algorithm>
#include <cstring>
#include <functional>
#include <unordered_map>

#include "tape.h"

namespace cg {

static void upload_inputs(Plan& p, const std::vector<Arg>& args) {
  // optimize.rs:87-90: an Input node takes the value of args[name] (kind and dims must match).
  for (Function* f : p.input_leaves) {
    const Arg* a = nullptr;
    for (const Arg& c : args)
      if (c.name == f->body->name) a = &c;
    if (!a) throw Error("panic: missing arg `" + f->body->name + "`");
    const bool is_matrix = !(a->h == 0 && a->w == 0);
    if (is_matrix != f->value.is_matrix) throw Error("panic: Cannot copy mismatched Constants.");
    if (is_matrix && (a->h != f->value.h || a->w != f->value.w)) throw Error("ERROR: copy_matrix: dims differ");
    value_upload_rowmajor(f->value, a->vals, /*sync=*/false);  // minimize() synchronises before returning
  }
}

static void print_host_value(const Value& v, const float* host) {
  if (!v.is_matrix) {
    printf("%g\n", host[0]);
    return;
  }
  for (unsigned i = 0; i < v.h; i++) {
    for (unsigned j = 0; j < v.w; j++) printf("%s%g", j ? " " : "", host[(size_t)j * v.h + i]);
    printf("\n");
  }
}
static void print_value(const Value& v) {
  // optimize.rs:66-68 `self.value().print()`; the reference's Display formatting (print.rs) is out of scope
  std::vector<float> host(v.size());
  value_download(v, host.data());
  print_host_value(v, host.data());
}

// The `print_freq` loss log (optimize.rs:66-68; SURVEY §5 "per-iteration loss ring buffer on device, D2H every print_freq
// only" / §8f-3). The reference prints inside the loop, which on a GPU means a blocking read-back per print; here the
// printed root values are copied device-to-device into a small ring as the iterations are enqueued, and the ring is
// read back (one copy, one synchronisation) and printed when it is full and at the end of the call — same lines, same
// order, no synchronisation inside the loop.
struct LossRing {
  static constexpr size_t kSlots = 32, kMaxBytes = (size_t)64 << 20;
  const Value& v;
  size_t n, slots = 0, used = 0;
  float* dev = nullptr;
  std::vector<float> host;
  cudaStream_t stream;
  LossRing(const Value& value, size_t root_n, size_t prints, cudaStream_t st) : v(value), n(root_n), stream(st) {
    slots = std::min(prints, kSlots);
    while (slots > 1 && slots * n * sizeof(float) > kMaxBytes) slots /= 2;
    if (slots && slots * n * sizeof(float) <= kMaxBytes) CG_CUDA(cudaMalloc(&dev, slots * n * sizeof(float)));
    host.resize(dev ? slots * n : 0);
  }
  bool active() const { return dev != nullptr; }
  void push(const float* root_ptr) {  // enqueued right behind the iteration whose value is printed
    CG_CUDA(cudaMemcpyAsync(dev + used * n, root_ptr, n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    if (++used == slots) flush();
  }
  void flush() {
    if (!used) return;
    CG_CUDA(cudaMemcpyAsync(host.data(), dev, used * n * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CG_CUDA(cudaStreamSynchronize(stream));
    for (size_t k = 0; k < used; k++) print_host_value(v, host.data() + k * n);
    fflush(stdout);
    used = 0;
  }
  ~LossRing() {
    if (dev) cudaFree(dev);
  }
};

// ---------------------------------------------------------------------------------------------------------
// CUDA-graph capture of one iteration as a DAG: steps are spread over several streams and joined with events,
// so independent kernels (the 4 + 4 gate products of an LSTM step, the elementwise chains next to the products,
// the weight-gradient all-reduces) run concurrently inside the graph. Dependencies are derived from the memory
// each step reads and writes (RAW, WAR and WAW on arena slots and leaf buffers), which also covers slot reuse.
// ---------------------------------------------------------------------------------------------------------
namespace {

// One node of the captured graph: the memory it reads / writes and how to enqueue it.
struct CapNode {
  std::vector<const float*> reads, writes;
  std::function<void(cudaStream_t)> launch;
  // >= 0: pinned to a dedicated stream (0 = copy-in, 1 = copy-out, 2 = layout kernels of the inputs, 3 = collectives: every
  // rank must issue the all-reduces of one communicator in the same order, so they share one in-order lane); -1: any compute
  // stream. A node that waits for a late transfer must not sit on a compute stream: everything captured behind it on
  // that stream would inherit the wait.
  int lane = -1;
};

CapNode step_node(Plan& p, size_t index, const Runtime& rt) {
  const Step& s = p.steps[index];
  CapNode io;
  if (s.kind == Step::EW) {
    for (uint32_t k = 0; k < s.prog.n_in; k++) io.reads.push_back(s.prog.in[k]);
    for (uint32_t k = 0; k < s.prog.n_out; k++) io.writes.push_back(s.prog.out[k]);
  } else {
    if (s.ins.a >= 0) io.reads.push_back(p.vals[s.ins.a].ptr);
    if (s.ins.b >= 0) io.reads.push_back(p.vals[s.ins.b].ptr);
    if (s.ins.c_in >= 0) io.reads.push_back(p.vals[s.ins.c_in].ptr);
    io.writes.push_back(p.vals[s.ins.dst].ptr);
    if (s.kind == Step::AVG) io.writes.push_back(rt.red_ws);  // the reduction workspace is a shared resource
  }
  if (s.kind == Step::ALLREDUCE || s.kind == Step::DPUPDATE) io.lane = 3;
  Plan* pp = &p;
  io.launch = [pp, index](cudaStream_t st) { plan_launch_step(*pp, index, st); };
  return io;
}

struct Loc {
  int last_writer = -1;
  std::vector<int> readers;
};

constexpr int kIoLanes = 4;

void capture_nodes(Plan& p, Runtime& rt, std::vector<CapNode>& nodes, cudaGraph_t* graph_out, cudaGraphExec_t* exec_out) {
  const int nstreams = rt.graph_streams > 0 ? rt.graph_streams : 0;  // (only the copy lane of launch_io_graph is a real stream)
  const size_t n = nodes.size();
  if ((int)rt.side_streams.size() < nstreams + kIoLanes) {
    rt.side_streams.resize(nstreams + kIoLanes, nullptr);
    for (auto& st : rt.side_streams)
      if (!st) CG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  }
  // dependencies from memory locations (RAW, WAR, WAW), in program order
  std::unordered_map<const float*, Loc> loc;
  std::vector<std::vector<int>> deps(n);
  int lane_prev[kIoLanes];
  for (int& l : lane_prev) l = -1;
  for (size_t i = 0; i < n; i++) {
    const CapNode& io = nodes[i];
    auto add = [&](int d) {
      if (d >= 0 && d != (int)i && std::find(deps[i].begin(), deps[i].end(), d) == deps[i].end()) deps[i].push_back(d);
    };
    for (const float* r : io.reads) add(loc[r].last_writer);
    for (const float* w : io.writes) {
      Loc& l = loc[w];
      add(l.last_writer);
      for (int rd : l.readers) add(rd);
    }
    for (const float* r : io.reads) loc[r].readers.push_back((int)i);
    for (const float* w : io.writes) {
      Loc& l = loc[w];
      l.last_writer = (int)i;
      l.readers.clear();
    }
    // lanes stay in order: every rank must issue the exchanges of one communicator in the same order, and the inputs
    // are copied in the order the iteration needs them
    if (io.lane >= 0 && io.lane < kIoLanes) {
      add(lane_prev[io.lane]);
      lane_prev[io.lane] = (int)i;
    }
  }
  // Transitive reduction: a dependency that is already an ancestor of another dependency of the same node adds an edge
  // and nothing else. The memory-derived sets above are full of those (every reader of a leaf depends on its last writer
  // AND on whatever that writer's other readers depended on), and what a graph launch costs the host grows with its edges:
  // for config 5 (~1000 nodes) an unreduced launch took ~45 ms of CPU time — invisible while launches queue up behind one
  // another, fully exposed in a `minimize(&args, lr, 1)` call, which waits for its result.
  {
    const size_t words = (n + 63) / 64;
    std::vector<uint64_t> anc(n * words, 0);  // anc[i]: every node i transitively depends on
    for (size_t i = 0; i < n; i++) {
      std::vector<int>& d = deps[i];
      std::sort(d.begin(), d.end(), std::greater<int>());  // nearest first: the likeliest to cover the others
      std::vector<int> kept;
      uint64_t* mine = &anc[i * words];
      for (int dep : d) {
        if (mine[(size_t)dep / 64] >> ((size_t)dep % 64) & 1) continue;  // reachable through a dependency already kept
        kept.push_back(dep);
        const uint64_t* theirs = &anc[(size_t)dep * words];
        for (size_t w = 0; w < words; w++) mine[w] |= theirs[w];
        mine[(size_t)dep / 64] |= (uint64_t)1 << ((size_t)dep % 64);
      }
      d.swap(kept);
    }
  }
  {
    // Exact DAG: everything is captured on ONE stream whose capture dependency set is replaced, before every node,
    // by exactly the nodes that node depends on (cudaStreamUpdateCaptureDependencies). No stream-order edges exist, so
    // a kernel that waits for a late input can never hold back an unrelated kernel that happened to be captured
    // behind it on the same stream — with side streams that cost 2.0 ms of a C4 iteration fed from the host.
    std::vector<std::vector<cudaGraphNode_t>> after(n);  // the dependency set of the stream after node i
    std::vector<char> has_succ(n, 0);
    CG_CUDA(cudaStreamBeginCapture(rt.stream, cudaStreamCaptureModeThreadLocal));
    try {
      for (size_t i = 0; i < n; i++) {
        std::vector<cudaGraphNode_t> d;
        auto add_after = [&](int dep) {
          has_succ[dep] = 1;
          for (cudaGraphNode_t g : after[dep])
            if (std::find(d.begin(), d.end(), g) == d.end()) d.push_back(g);
        };
        for (int dep : deps[i]) add_after(dep);
        CG_CUDA(cudaStreamUpdateCaptureDependencies(rt.stream, d.empty() ? nullptr : d.data(), d.size(), cudaStreamSetCaptureDependencies));
        nodes[i].launch(rt.stream);
        cudaStreamCaptureStatus status;
        const cudaGraphNode_t* cur = nullptr;
        size_t ncur = 0;
        CG_CUDA(cudaStreamGetCaptureInfo(rt.stream, &status, nullptr, nullptr, &cur, &ncur));
        if (status != cudaStreamCaptureStatusActive) throw Error("internal: stream capture was invalidated");
        after[i].assign(cur, cur + ncur);
      }
      // end on the join of every sink
      std::vector<cudaGraphNode_t> sinks;
      for (size_t i = 0; i < n; i++)
        if (!has_succ[i])
          for (cudaGraphNode_t g : after[i])
            if (std::find(sinks.begin(), sinks.end(), g) == sinks.end()) sinks.push_back(g);
      CG_CUDA(cudaStreamUpdateCaptureDependencies(rt.stream, sinks.empty() ? nullptr : sinks.data(), sinks.size(), cudaStreamSetCaptureDependencies));
    } catch (...) {
      cudaGraph_t g = nullptr;
      cudaStreamEndCapture(rt.stream, &g);
      if (g) cudaGraphDestroy(g);
      throw;
    }
    CG_CUDA(cudaStreamEndCapture(rt.stream, graph_out));
    CG_CUDA(cudaGraphInstantiate(exec_out, *graph_out, 0));
  }
}

void capture_dag(Plan& p, Runtime& rt) {
  std::vector<CapNode> nodes;
  for (size_t i = 0; i < p.steps.size(); i++) nodes.push_back(step_node(p, i, rt));
  capture_nodes(p, rt, nodes, &p.graph, &p.graph_exec);
}

// The node a stream capture has just appended (the only dependency of whatever is enqueued next on `st`).
cudaGraphNode_t last_captured_node(cudaStream_t st) {
  cudaStreamCaptureStatus status;
  const cudaGraphNode_t* deps = nullptr;
  size_t ndeps = 0;
  CG_CUDA(cudaStreamGetCaptureInfo(st, &status, nullptr, nullptr, &deps, &ndeps));
  if (status != cudaStreamCaptureStatusActive || ndeps != 1) throw Error("internal: cannot identify the captured copy node");
  return deps[0];
}

bool is_pinned_host(const void* ptr) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

const Arg* find_arg(const std::vector<Arg>& args, const Function* f) {
  const Arg* a = nullptr;
  for (const Arg& c : args)
    if (c.name == f->body->name) a = &c;
  if (!a) throw Error("panic: missing arg `" + f->body->name + "`");
  const bool is_matrix = !(a->h == 0 && a->w == 0);
  if (is_matrix != f->value.is_matrix) throw Error("panic: Cannot copy mismatched Constants.");
  if (is_matrix && (a->h != f->value.h || a->w != f->value.w)) throw Error("ERROR: copy_matrix: dims differ");
  return a;
}

// The transfers of one call can live inside the graph when every host buffer is page-locked (a copy node from
// pageable memory would be staged synchronously by the driver and could not overlap anything).
bool io_graph_eligible(const Plan& p, const Runtime& rt, const std::vector<Arg>& args, const float* trace) {
  if (!rt.use_graph || !rt.io_graph || rt.graph_streams <= 0 || p.steps.empty() || p.input_leaves.empty()) return false;
  for (Function* f : p.input_leaves)
    if (!is_pinned_host(find_arg(args, f)->vals)) return false;
  return !trace || is_pinned_host(trace);
}

// The order in which the inputs of one call should arrive. The walker discovers leaves depth first, which for a
// recurrence is LAST time-step first (the root is C_T; x_T-1 is met before the walk descends to step 0), but the
// step that needs x_0 has the whole recurrence behind it. So inputs are ranked by the weight (flop) of the longest
// dependency chain that starts at one of their consumers: with x_0 copied last, 2.0 of the 2.8 ms of H2D time of a
// C4 iteration were exposed (11.05 vs 9.02 ms, tools/probes/e2e_timing.py).
std::vector<size_t> input_urgency_order(Plan& p, const Runtime& rt) {
  const size_t n = p.steps.size();
  std::vector<CapNode> nodes;
  for (size_t i = 0; i < n; i++) nodes.push_back(step_node(p, i, rt));
  std::unordered_map<const float*, Loc> loc;
  std::vector<std::vector<int>> succ(n);
  for (size_t i = 0; i < n; i++) {
    auto dep = [&](int d) {
      if (d >= 0 && d != (int)i) succ[d].push_back((int)i);
    };
    for (const float* r : nodes[i].reads) dep(loc[r].last_writer);
    for (const float* w : nodes[i].writes) {
      Loc& l = loc[w];
      dep(l.last_writer);
      for (int rd : l.readers) dep(rd);
    }
    for (const float* r : nodes[i].reads) loc[r].readers.push_back((int)i);
    for (const float* w : nodes[i].writes) {
      Loc& l = loc[w];
      l.last_writer = (int)i;
      l.readers.clear();
    }
  }
  std::vector<double> tail(n, 0.0);
  for (size_t k = n; k-- > 0;) {
    double best = 0.0;
    for (int s : succ[k]) best = std::max(best, tail[s]);
    tail[k] = best + plan_step_flop(p, k) + 1e6;  // every kernel costs something, products their flop
  }
  const size_t nin = p.input_leaves.size();
  std::vector<double> urgency(nin, 0.0);
  for (size_t j = 0; j < nin; j++) {
    const float* buf = p.input_leaves[j]->value.buf->ptr;
    for (size_t i = 0; i < n; i++)
      if (std::find(nodes[i].reads.begin(), nodes[i].reads.end(), buf) != nodes[i].reads.end()) urgency[j] = std::max(urgency[j], tail[i]);
  }
  std::vector<size_t> order(nin);
  for (size_t j = 0; j < nin; j++) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return urgency[a] > urgency[b]; });
  return order;
}

void launch_io_graph(Plan& p, Runtime& rt, const std::vector<Arg>& args, float* trace) {
  IoGraph& g = p.io[trace ? 1 : 0];
  const size_t nin = p.input_leaves.size();
  std::vector<const float*> src(nin);
  for (size_t i = 0; i < nin; i++) src[i] = find_arg(args, p.input_leaves[i])->vals;
  if (!g.exec) {
    if (p.io_staging.size() < nin) p.io_staging.resize(nin);
    g.h2d_nodes.assign(nin, nullptr);
    g.h2d_src = src;
    g.h2d_dst.assign(nin, nullptr);
    g.h2d_bytes.assign(nin, 0);
    g.d2h_dst = trace;
    // H2D memcpy nodes inside a multi-branch graph cost ~1.3 ms per C4 iteration over the same copies overlapped by
    // hand (11.1 vs 9.85 ms on B200; tools/e2e_probe.py): by default the copies run on a copy stream of their own and
    // the graph holds an external-event wait per input. CG_IO_EXTERNAL=0 keeps them as memcpy nodes.
    static const bool external = [] {
      const char* e = getenv("CG_IO_EXTERNAL");
      return !e || atoi(e) != 0;
    }();
    g.external_h2d = external;
    if (external) {
      g.in_ev.resize(nin);
      for (auto& e : g.in_ev) CG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    std::vector<CapNode> nodes;
    IoGraph* gp = &g;
    Plan* pp = &p;
    auto d2h = [&]() {
      CapNode c;
      c.lane = 1;
      c.reads.push_back(p.root_ptr);
      c.launch = [gp, pp](cudaStream_t st) {
        CG_CUDA(cudaMemcpyAsync(gp->d2h_dst, pp->root_ptr, pp->root_n * sizeof(float), cudaMemcpyDeviceToHost, st));
        gp->d2h_node = last_captured_node(st);
      };
      return c;
    };
    // a leaf root is read BEFORE this iteration's update overwrites it (Q12)
    if (trace && p.root_is_leaf) nodes.push_back(d2h());
    g.order = input_urgency_order(p, rt);
    for (size_t i : g.order) {
      Value& v = p.input_leaves[i]->value;
      const size_t n = v.size();
      const bool direct = !v.is_matrix || v.h == 1 || v.w == 1;  // both layouts coincide
      if (!direct && (!p.io_staging[i] || p.io_staging[i]->n < n)) p.io_staging[i] = DevBuf::alloc(n);
      g.h2d_dst[i] = direct ? v.buf->ptr : p.io_staging[i]->ptr;
      g.h2d_bytes[i] = n * sizeof(float);
      CapNode c;
      c.lane = 0;  // one copy-in lane: the inputs arrive in the order the forward pass consumes them
      c.writes.push_back(g.h2d_dst[i]);
      if (g.external_h2d) {
        // the copy itself is enqueued on the copy stream before every launch; the graph only waits for it
        c.launch = [gp, i](cudaStream_t st) { CG_CUDA(cudaStreamWaitEvent(st, gp->in_ev[i], cudaEventWaitExternal)); };
      } else {
        c.launch = [gp, i](cudaStream_t st) {
          CG_CUDA(cudaMemcpyAsync(gp->h2d_dst[i], gp->h2d_src[i], gp->h2d_bytes[i], cudaMemcpyHostToDevice, st));
          gp->h2d_nodes[i] = last_captured_node(st);
        };
      }
      nodes.push_back(std::move(c));
      if (!direct) {
        CapNode t;
        t.lane = 2;
        t.reads.push_back(g.h2d_dst[i]);
        t.writes.push_back(v.buf->ptr);
        const float* from = g.h2d_dst[i];
        float* to = v.buf->ptr;
        const unsigned h = v.h, w = v.w;
        t.launch = [from, to, h, w](cudaStream_t st) { launch_rowmajor_to_colmajor(from, to, h, w, st); };
        nodes.push_back(std::move(t));
        g.layout_kernels++;
      }
    }
    for (size_t i = 0; i < p.steps.size(); i++) nodes.push_back(step_node(p, i, rt));
    if (trace && !p.root_is_leaf) nodes.push_back(d2h());
    capture_nodes(p, rt, nodes, &g.graph, &g.exec);
  } else if (!g.external_h2d) {
    for (size_t i = 0; i < nin; i++)
      if (g.h2d_src[i] != src[i]) {
        CG_CUDA(cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.h2d_nodes[i], g.h2d_dst[i], src[i], g.h2d_bytes[i],
                                                   cudaMemcpyHostToDevice));
        g.h2d_src[i] = src[i];
      }
  }
  if (g.exec && trace && g.d2h_dst != trace) {
    CG_CUDA(cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.d2h_node, trace, p.root_ptr, p.root_n * sizeof(float),
                                               cudaMemcpyDeviceToHost));
    g.d2h_dst = trace;
  }
  if (g.external_h2d) {
    // this call's copies, in consumption order, behind everything that still reads the landing buffers
    cudaStream_t copy = rt.side_streams[(size_t)(rt.graph_streams > 0 ? rt.graph_streams : 0) + 0];
    if (!p.io_done) CG_CUDA(cudaEventCreateWithFlags(&p.io_done, cudaEventDisableTiming));
    CG_CUDA(cudaEventRecord(p.io_done, rt.stream));
    CG_CUDA(cudaStreamWaitEvent(copy, p.io_done, 0));
    for (size_t i : g.order) {
      CG_CUDA(cudaMemcpyAsync(g.h2d_dst[i], src[i], g.h2d_bytes[i], cudaMemcpyHostToDevice, copy));
      CG_CUDA(cudaEventRecord(g.in_ev[i], copy));
    }
  }
  CG_CUDA(cudaGraphLaunch(g.exec, rt.stream));
  rt.launches += p.steps.size() + g.layout_kernels;
}

}  // namespace

// `f.eval(&args)` (optimize.rs:8-32): one forward pass, no parameter is touched and the cached value of `f` stays what
// it was. The forward-only tape is compiled once per function and launched eagerly (eval is not a loop).
void function_eval(Function& f, const std::vector<Arg>& args, float* dst) {
  if (plan_only()) throw Error("plan-only mode: eval needs a CUDA device (this backend has no CPU path)");
  Runtime& rt = Runtime::get();
  if (!f.eval_plan || f.eval_plan->mode != rt.mode || f.eval_plan->tf32 != rt.tf32 || f.eval_plan->fuse != rt.fuse ||
      f.eval_plan->dp_epoch != rt.dp_epoch || f.eval_plan->ew_jit_min != rt.ew_jit_min) {
    f.eval_plan.reset();
    f.eval_plan = build_plan(f, 0.0f, /*keep_root=*/true, /*forward_only=*/true);
  }
  Plan& p = *f.eval_plan;
  upload_inputs(p, args);
  plan_launch_iteration(p, rt.stream);
  rt.launches += p.steps.size();
  CG_CUDA(cudaMemcpyAsync(dst, p.root_ptr, p.root_n * sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
  CG_CUDA(cudaStreamSynchronize(rt.stream));
}

// Captures the iteration of `p` as a CUDA graph (a DAG over side streams, or one linear stream) if it is not yet.
static void ensure_graph(Plan& p, Runtime& rt) {
  if (!rt.use_graph || p.graph_exec || p.steps.empty()) return;
  if (rt.graph_streams > 0) {
    capture_dag(p, rt);
    return;
  }
  CG_CUDA(cudaStreamBeginCapture(rt.stream, cudaStreamCaptureModeThreadLocal));
  try {
    plan_launch_iteration(p, rt.stream);
  } catch (...) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(rt.stream, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  CG_CUDA(cudaStreamEndCapture(rt.stream, &p.graph));
  CG_CUDA(cudaGraphInstantiate(&p.graph_exec, p.graph, 0));
}

void minimize(Function& f, const std::vector<Arg>& args, float lr, int iters, int print_freq, float* trace,
              bool synchronize, bool reuse_inputs) {
  Runtime& rt = Runtime::get();
  if (!f.plan || f.plan->lr != lr || f.plan->mode != rt.mode || f.plan->tf32 != rt.tf32 ||
      f.plan->fix_act_grad != rt.fix_act_grad || f.plan->fuse != rt.fuse || f.plan->dp_epoch != rt.dp_epoch || f.plan->ew_jit_min != rt.ew_jit_min) {
    f.plan.reset();
    f.plan = build_plan(f, lr);
  }
  Plan& p = *f.plan;
  if (plan_only()) {
    if (iters > 0) throw Error("plan-only mode: minimize needs a CUDA device (this backend has no CPU path)");
    return;  // the plan is compiled; nothing is captured or launched
  }
  // One iteration whose host transfers ride inside the graph (pinned buffers only); otherwise: upload, then loop.
  const bool vm = p.vm_steps != nullptr || p.tape_kernel != nullptr;
  const bool io = !vm && !reuse_inputs && iters >= 1 && (iters == 1 || !trace) && io_graph_eligible(p, rt, args, trace);
  if (!reuse_inputs && !io) upload_inputs(p, args);
  for (Function* leaf : p.param_leaves) leaf->value.summary_valid = false;  // device values diverge from here on
  f.value.summary_valid = false;
  f.value.defined = true;

  if (vm) {
    // Resident tape executor: one kernel runs every iteration up to the next print (vm_kernel.cu).
    float* trace_dev = nullptr;
    if (trace && iters > 0) CG_CUDA(cudaMalloc(&trace_dev, (size_t)iters * p.root_n * sizeof(float)));
    int done = 0;
    while (done < iters) {
      int chunk = iters - done;
      if (print_freq > 0) chunk = std::min(chunk, print_freq - done % print_freq);
      float* tr = trace_dev ? trace_dev + (size_t)done * p.root_n : nullptr;
      if (p.tape_kernel) launch_tape_jit(p.tape_kernel, chunk, tr, rt.stream);
      else launch_tape_vm(p.vm_steps, p.vm_progs, p.vm_nsteps, chunk, p.root_ptr, p.root_n, p.root_is_leaf, tr, rt.stream);
      rt.launches++;
      done += chunk;
      if (print_freq > 0 && done % print_freq == 0) print_value(f.value);
    }
    if (trace_dev) {
      CG_CUDA(cudaMemcpyAsync(trace, trace_dev, (size_t)iters * p.root_n * sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
      CG_CUDA(cudaStreamSynchronize(rt.stream));
      CG_CUDA(cudaFree(trace_dev));
    }
    if (synchronize) CG_CUDA(cudaStreamSynchronize(rt.stream));
    return;
  }
  ensure_graph(p, rt);
  if (p.quiet) ensure_graph(*p.quiet, rt);

  const size_t prints = print_freq > 0 ? (size_t)(iters / print_freq) : 0;
  LossRing ring(f.value, p.root_n, prints, rt.stream);
  auto print_now = [&] {
    if (ring.active()) ring.push(p.root_ptr);
    else print_value(f.value);
  };
  float* trace_dev = nullptr;
  const bool direct_trace = trace && iters == 1;  // one iteration: read the root value straight back
  if (trace && iters > 1) CG_CUDA(cudaMalloc(&trace_dev, (size_t)iters * p.root_n * sizeof(float)));
  if (direct_trace && p.root_is_leaf && !io)
    CG_CUDA(cudaMemcpyAsync(trace, p.root_ptr, p.root_n * sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
  for (int i = 0; i < iters; i++) {
    if (i == 0 && io) {
      launch_io_graph(p, rt, args, trace);
      if (print_freq > 0 && (i + 1) % print_freq == 0) print_now();
      continue;
    }
    // the traced value is the forward value of iteration i, i.e. before its update (Q12)
    if (trace_dev && p.root_is_leaf)
      CG_CUDA(cudaMemcpyAsync(trace_dev + (size_t)i * p.root_n, p.root_ptr, p.root_n * sizeof(float),
                              cudaMemcpyDeviceToDevice, rt.stream));
    // Only the LAST iteration's root value is observable after the call (Q12) unless it is traced or printed: the
    // others run the variant of the plan that does not materialise it (build_plan, `quiet`).
    const bool observed = i == iters - 1 || trace_dev || (print_freq > 0 && (i + 1) % print_freq == 0);
    Plan& q = (p.quiet && !observed) ? *p.quiet : p;
    if (q.graph_exec) CG_CUDA(cudaGraphLaunch(q.graph_exec, rt.stream));
    else plan_launch_iteration(q, rt.stream);
    rt.launches += q.steps.size();
    if (trace_dev && !p.root_is_leaf)
      CG_CUDA(cudaMemcpyAsync(trace_dev + (size_t)i * p.root_n, p.root_ptr, p.root_n * sizeof(float),
                              cudaMemcpyDeviceToDevice, rt.stream));
    if (print_freq > 0 && (i + 1) % print_freq == 0) print_now();
  }
  ring.flush();
  if (direct_trace) {
    if (!p.root_is_leaf && !io)
      CG_CUDA(cudaMemcpyAsync(trace, p.root_ptr, p.root_n * sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
    CG_CUDA(cudaStreamSynchronize(rt.stream));
  }
  if (trace_dev) {
    CG_CUDA(cudaMemcpyAsync(trace, trace_dev, (size_t)iters * p.root_n * sizeof(float), cudaMemcpyDeviceToHost,
                            rt.stream));
    CG_CUDA(cudaStreamSynchronize(rt.stream));
    CG_CUDA(cudaFree(trace_dev));
  }
  if (synchronize) CG_CUDA(cudaStreamSynchronize(rt.stream));
  // a fused data-parallel exchange that lost a peer records it instead of hanging the GPU: raise it here
  if ((synchronize || trace) && p.stats.dp_fused_updates) dp_check_errors();
}

}  // namespace cg
