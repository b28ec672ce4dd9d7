// exec.cpp — `minimize` (src/optimize.rs:59-72) over a compiled plan: one CUDA-graph replay per iteration.
#include <cstring>

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

  if (rt.use_graph && !p.graph_exec && !p.steps.empty()) {
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
