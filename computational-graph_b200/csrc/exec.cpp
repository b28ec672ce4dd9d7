// exec.cpp — `minimize` (src/optimize.rs:59-72) over a compiled plan: one CUDA-graph replay per iteration.
#include <algorithm>
#include <cstring>
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

static void print_value(const Value& v) {
  // optimize.rs:66-68 `self.value().print()`; the reference's Display formatting (print.rs) is out of scope
  std::vector<float> host(v.size());
  value_download(v, host.data());
  if (!v.is_matrix) {
    printf("%g\n", host[0]);
    return;
  }
  for (unsigned i = 0; i < v.h; i++) {
    for (unsigned j = 0; j < v.w; j++) printf("%s%g", j ? " " : "", host[(size_t)j * v.h + i]);
    printf("\n");
  }
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-graph capture of one iteration as a DAG: steps are spread over several streams and joined with events,
// so independent kernels (the 4 + 4 gate products of an LSTM step, the elementwise chains next to the products,
// the weight-gradient all-reduces) run concurrently inside the graph. Dependencies are derived from the memory
// each step reads and writes (RAW, WAR and WAW on arena slots and leaf buffers), which also covers slot reuse.
// ---------------------------------------------------------------------------------------------------------
namespace {

struct StepIO {
  std::vector<const float*> reads, writes;
};

StepIO step_io(const Plan& p, const Step& s, const Runtime& rt) {
  StepIO io;
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
  return io;
}

struct Loc {
  int last_writer = -1;
  std::vector<int> readers;
};

void capture_dag(Plan& p, Runtime& rt) {
  const int nstreams = rt.graph_streams;
  const size_t n = p.steps.size();
  if ((int)rt.side_streams.size() < nstreams) {
    rt.side_streams.resize(nstreams, nullptr);
    for (auto& st : rt.side_streams)
      if (!st) CG_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  }
  while (p.events.size() < n + 1) {
    cudaEvent_t e;
    CG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p.events.push_back(e);
  }
  // dependencies from memory locations
  std::unordered_map<const float*, Loc> loc;
  std::vector<std::vector<int>> deps(n);
  for (size_t i = 0; i < n; i++) {
    StepIO io = step_io(p, p.steps[i], rt);
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
  }
  // stream 0 is the capturing (origin) stream
  std::vector<cudaStream_t> streams{rt.stream};
  for (int k = 0; k < nstreams; k++) streams.push_back(rt.side_streams[k]);
  std::vector<int> stream_last(streams.size(), -1), step_stream(n, 0);
  std::vector<char> forked(streams.size(), 0);
  forked[0] = 1;
  cudaEvent_t fork_ev = p.events[n];
  CG_CUDA(cudaStreamBeginCapture(rt.stream, cudaStreamCaptureModeThreadLocal));
  try {
    CG_CUDA(cudaEventRecord(fork_ev, rt.stream));
    for (size_t i = 0; i < n; i++) {
      int chosen = -1;
      // 1. a stream whose most recent node is one of our dependencies: no extra edge, no false dependency
      for (size_t s = 0; s < streams.size(); s++)
        if (stream_last[s] >= 0 && std::find(deps[i].begin(), deps[i].end(), stream_last[s]) != deps[i].end())
          if (chosen < 0 || stream_last[s] > stream_last[chosen]) chosen = (int)s;
      // 2. an unused stream; 3. the stream that has been idle longest
      if (chosen < 0)
        for (size_t s = 0; s < streams.size() && chosen < 0; s++)
          if (stream_last[s] < 0) chosen = (int)s;
      if (chosen < 0) {
        chosen = 0;
        for (size_t s = 1; s < streams.size(); s++)
          if (stream_last[s] < stream_last[chosen]) chosen = (int)s;
      }
      cudaStream_t st = streams[chosen];
      if (!forked[chosen]) {
        CG_CUDA(cudaStreamWaitEvent(st, fork_ev, 0));
        forked[chosen] = 1;
      }
      for (int d : deps[i])
        if (step_stream[d] != chosen) CG_CUDA(cudaStreamWaitEvent(st, p.events[d], 0));
      plan_launch_step(p, i, st);
      CG_CUDA(cudaEventRecord(p.events[i], st));
      stream_last[chosen] = (int)i;
      step_stream[i] = chosen;
    }
    for (size_t s = 1; s < streams.size(); s++)
      if (stream_last[s] >= 0) CG_CUDA(cudaStreamWaitEvent(rt.stream, p.events[stream_last[s]], 0));
  } catch (...) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(rt.stream, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  CG_CUDA(cudaStreamEndCapture(rt.stream, &p.graph));
  CG_CUDA(cudaGraphInstantiate(&p.graph_exec, p.graph, 0));
}

}  // namespace

void minimize(Function& f, const std::vector<Arg>& args, float lr, int iters, int print_freq, float* trace,
              bool synchronize, bool reuse_inputs) {
  Runtime& rt = Runtime::get();
  if (!f.plan || f.plan->lr != lr || f.plan->mode != rt.mode || f.plan->tf32 != rt.tf32 ||
      f.plan->fix_act_grad != rt.fix_act_grad || f.plan->fuse != rt.fuse || f.plan->dp_epoch != rt.dp_epoch) {
    f.plan.reset();
    f.plan = build_plan(f, lr);
  }
  Plan& p = *f.plan;
  if (!reuse_inputs) upload_inputs(p, args);
  for (Function* leaf : p.param_leaves) leaf->value.summary_valid = false;  // device values diverge from here on
  f.value.summary_valid = false;
  f.value.defined = true;

  if (rt.use_graph && !p.graph_exec && !p.steps.empty() && rt.graph_streams > 0) {
    capture_dag(p, rt);
  } else if (rt.use_graph && !p.graph_exec && !p.steps.empty()) {
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

  float* trace_dev = nullptr;
  const bool direct_trace = trace && iters == 1;  // one iteration: read the root value straight back
  if (trace && iters > 1) CG_CUDA(cudaMalloc(&trace_dev, (size_t)iters * p.root_n * sizeof(float)));
  if (direct_trace && p.root_is_leaf)
    CG_CUDA(cudaMemcpyAsync(trace, p.root_ptr, p.root_n * sizeof(float), cudaMemcpyDeviceToHost, rt.stream));
  for (int i = 0; i < iters; i++) {
    // the traced value is the forward value of iteration i, i.e. before its update (Q12)
    if (trace_dev && p.root_is_leaf)
      CG_CUDA(cudaMemcpyAsync(trace_dev + (size_t)i * p.root_n, p.root_ptr, p.root_n * sizeof(float),
                              cudaMemcpyDeviceToDevice, rt.stream));
    if (p.graph_exec) CG_CUDA(cudaGraphLaunch(p.graph_exec, rt.stream));
    else plan_launch_iteration(p, rt.stream);
    rt.launches += p.steps.size();
    if (trace_dev && !p.root_is_leaf)
      CG_CUDA(cudaMemcpyAsync(trace_dev + (size_t)i * p.root_n, p.root_ptr, p.root_n * sizeof(float),
                              cudaMemcpyDeviceToDevice, rt.stream));
    if (print_freq > 0 && (i + 1) % print_freq == 0) print_value(f.value);
  }
  if (direct_trace) {
    if (!p.root_is_leaf)
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
}

}  // namespace cg
