/* cg_graph.h — graph-level C ABI of libcg_b200.so: the reference's L4 surface
 * (Function / ops / minimize) as plain `extern "C"` entry points.
 *
 * Why a second boundary besides cg_matrix.h: the reference crosses its FFI once per node per
 * iteration (one eager kernel per call, src/ops.rs:61-158). Fusing sub-trees into one kernel, a
 * device-resident pool with zero allocations in the loop, CUDA-graph replay and data-parallel
 * gradient all-reduce cannot be expressed through per-op calls, so the host layer a Rust shim
 * (function.rs / ops.rs / optimize.rs re-done as thin wrappers) binds is THIS header.
 * Each entry point cites the reference interface it replaces. Handles are opaque; every function
 * that can fail returns NULL / non-zero and leaves a message in cg_last_error() — the message is the
 * reference's own `panic!` / `check()` text ("panic: …" / "ERROR: …").
 *
 * Values: matrices are passed in as ROW-major host floats (what `Constant::matrix(h, w, vals)` takes,
 * src/matrix.rs:59-66) and stored column-major on the device; read-backs return raw storage order
 * (column-major), exactly like `get_array`.
 */
#ifndef CG_GRAPH_H
#define CG_GRAPH_H
#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cg_function cg_function; /* `struct Function` (src/function.rs:10-16) */

/* ---- leaves: src/function.rs:100-155 ----------------------------------------------------------- */
cg_function *cg_scalar_param(const char *name, float value);                               /* :115-118 */
cg_function *cg_matrix_param(const char *name, unsigned h, unsigned w, const float *vals); /* :120-123 */
cg_function *cg_single_val_param(const char *name, unsigned h, unsigned w, float value);   /* :125-128 */
cg_function *cg_scalar(float value);                                                       /* :141-144 */
cg_function *cg_matrix(unsigned h, unsigned w, const float *vals);                         /* :146-149 */
cg_function *cg_single_val_matrix(unsigned h, unsigned w, float value);                    /* :151-154 */
cg_function *cg_input(const char *name, int is_matrix, unsigned h, unsigned w);            /* :104-108 */
cg_function *cg_clone(const cg_function *f);   /* #[derive(Clone)]: deep value copy, shared body (Q1) */
void cg_release(cg_function *f);               /* Drop */

/* ---- op constructors: src/ops.rs:36-54,199-264,331-343,363-376,395-407 -------------------------
 * Value-based identity elimination and const folding happen here, as in the reference (Q4). */
cg_function *cg_neg(const cg_function *f);
cg_function *cg_sq(const cg_function *f);
cg_function *cg_abs(const cg_function *f);
cg_function *cg_signum(const cg_function *f);
cg_function *cg_sigmoid(const cg_function *f);
cg_function *cg_tanh(const cg_function *f);
cg_function *cg_add(const cg_function *a, const cg_function *b);
cg_function *cg_sub(const cg_function *a, const cg_function *b);
cg_function *cg_mul(const cg_function *a, const cg_function *b);
cg_function *cg_dot(const cg_function *a, int trans1, const cg_function *b, int trans2); /* dot!, macros.rs:51-64 */

/* ---- optimisation: src/optimize.rs:59-81 ---------------------------------------------------------
 * args: `&HashMap<&str, Constant>` flattened — for entry i, arg_vals[i] points at arg_h[i]*arg_w[i]
 * row-major floats, or one float with arg_h[i] == arg_w[i] == 0 for a scalar.
 * trace (may be NULL): receives the root value of every iteration (iters * size floats, storage
 * order) — the value the reference would print, i.e. the forward value BEFORE that iteration's
 * update (Q12). print_freq <= 0 or > iters: never print. Returns 0 on success. Blocks until done. */
int cg_minimize(cg_function *f, int nargs, const char **names, const float **arg_vals, const unsigned *arg_h,
                const unsigned *arg_w, float learn_rate, int iters, int print_freq, float *trace);
int cg_maximize(cg_function *f, int nargs, const char **names, const float **arg_vals, const unsigned *arg_h,
                const unsigned *arg_w, float learn_rate, int iters, int print_freq, float *trace);
/* Page-locked host memory for `args` values and `trace` buffers. When every host buffer of a
 * `cg_minimize(..., iters = 1, ...)` call is page-locked (these, cudaHostAlloc/cudaHostRegister, torch pin_memory),
 * the H2D copies of the inputs and the D2H read of the root value become nodes of the iteration's CUDA graph and
 * overlap the compute; pageable buffers are copied before / after it as the reference does. */
void *cg_host_alloc(size_t bytes);
void cg_host_free(void *ptr);
/* Same loop, enqueued on the library stream without waiting (device-resident timing: bench.py). */
int cg_minimize_async(cg_function *f, float learn_rate, int iters);
int cg_synchronize(void);

/* ---- read-back: Function::value / all_less_than / all_equal (src/ops.rs:309-329) ----------------- */
/* `f.eval(&args)` (src/optimize.rs:8-32): the forward value of f at the current parameters, written to dst in storage
 * (column-major) order — cg_value_dims gives the shape. No parameter and no cached value changes. */
int cg_eval(cg_function *f, int nargs, const char **names, const float **arg_vals, const unsigned *arg_h,
            const unsigned *arg_w, float *dst);
/* `f.grad(param)` (src/optimize.rs:34-56): a NEW function, d f / d param, built symbolically (the README's "naive"
 * gradient; products are "not implemented", as in the reference). d(f^2) = 2 f f' — the reference drops the 2. */
cg_function *cg_grad(const cg_function *f, const char *param);
/* The device address of leaf i's storage (in plan-only mode: its placeholder) — identifies the leaf in a dumped
 * resident-kernel pointer table (CG_TAPE_JIT_DUMP). 0 when out of range. */
unsigned long long cg_leaf_address(cg_function *f, int i);
int cg_value_dims(const cg_function *f, unsigned *h, unsigned *w); /* returns 1 for a matrix, 0 for a scalar */
void cg_value_get(const cg_function *f, float *dst);
bool cg_all_less_than(const cg_function *f, float x);
bool cg_all_equal(const cg_function *f, float x);
/* Parameter leaves in memoised post-order; every use site of a param is its own leaf (Q1). */
int cg_num_leaves(cg_function *f);
int cg_leaf_info(cg_function *f, int i, char *name, int name_cap, unsigned *h, unsigned *w);
int cg_leaf_get(cg_function *f, int i, float *dst);
int cg_leaf_set(cg_function *f, int i, const float *src_rowmajor); /* checkpoint restore */
int cg_kind(const cg_function *f);

/* ---- execution switches (defaults reproduce the reference; also CG_MODE / CG_TF32 / … env) ------- */
void cg_set_mode(int mode);           /* 0 = exact (the reference's tree walk), 1 = dag (memoised, visit-once) */
void cg_set_tf32(int on);             /* 1: tcgen05 TF32 tensor-core products for eligible shapes; 2: FP32-accurate products on the
                                       * tensor core for eligible shapes (every operand split into the part TF32 keeps and the
                                       * part it drops, three MMAs per k-step: ~1e-6 per term, 1/3 of the TF32 rate) */
void cg_set_fix_act_grad(int on);     /* 1: sigmoid/tanh pass error*derivative (repairs Q3) */
/* Corrected-semantics switches, each off by default so that the reference is reproduced:
 * broadcast_sum: in the backward pass a scalar that absorbs a matrix-shaped error takes its SUM (the gradient of a
 *   broadcast scalar; the root error is ones, so the loss is the sum of the root's elements) instead of the reference's
 *   mean (Q5: src/ops.rs:79-80,103-104,126-134); sub_sign: `0 - x` builds `-x` instead of `x` (Q4: src/ops.rs:255-263). */
void cg_set_fix_broadcast_sum(int on);
void cg_set_fix_sub_sign(int on);
/* tied_params: every use of a parameter NAME is one parameter (the reference gives each use site a private copy that
 * drifts apart, Q1: src/function.rs:10-16). The backward pass accumulates one gradient per name from un-updated values
 * and every copy takes the same update — true gradient descent on shared weights (an LSTM whose steps share W). */
void cg_set_fix_tied_params(int on);
void cg_set_fuse(int on);             /* 0: one kernel per tape instruction (ablation) */
void cg_set_cuda_graph(int on);       /* 0: launch kernels eagerly instead of replaying a CUDA graph */
void cg_set_strict_gemm_max(int dim); /* products with max(m,n,k) <= dim use the bit-exact kernel */
void cg_set_graph_streams(int n);     /* side streams of the per-iteration CUDA graph (0: one linear stream) */
/* Elementwise programs over >= min_elems elements are unrolled into straight-line sm_100a code at plan time
 * (NVRTC; bit-identical to the interpreter kernel, which keeps the smaller ones). < 0: never. Default 32768. */
void cg_set_ew_jit_min(long long min_elems);
int cg_ew_jit_available(void);
long long cg_ew_jit_selftest(void);   /* compiles a program using every opcode; cubin bytes or -1 (no GPU needed) */
void cg_set_tf32_pair(int on);        /* 1 (default): 2-SM (cta_group::2) TF32 product kernel where the shape allows (m % 256 == 0); 0: single-CTA kernel */
void cg_set_io_graph(int on);         /* 0: never put host transfers inside the iteration graph */
/* dag mode. 1 (default): backward work that does not depend on the loss (Q3: a Sigmoid / Tanh hands its child the local
 * derivative only, src/optimize.rs:168-179 — every gate derivative, weight-gradient product and weight update of the
 * LSTM) is scheduled right behind the forward op it depends on. Same arithmetic, same bits; 0 keeps the walk order. */
void cg_set_hoist(int on);
/* Latency-bound plans — every elementwise domain <= dim*dim elements and every product dimension <= min(dim, the strict
 * limit) — are executed by ONE resident CTA that walks the whole tape, iteration loop included, instead of one CUDA-graph
 * node per step (vm_kernel.cu; same arithmetic, same bits). Default 48 (covers the README LSTM sweep, d <= 30); 0: never. */
void cg_set_vm_max_dim(int dim);
/* 1 (default): a resident plan whose values all fit in one SM's shared memory (<= 224 KB) is compiled at plan time
 * (NVRTC) into ONE straight-line kernel — shapes as constants, every leaf and temporary in shared memory for the
 * whole minimize call, leaves loaded once and written back once. 0: the interpreting executor / the CUDA graph. */
void cg_set_tape_jit(int on);
/* 1 (default): after minimize(n) only the LAST iteration's root value is observable (src/optimize.rs:64-71) unless it is
 * traced or printed, so for elementwise-only graphs iterations 1..n-1 run a variant of the fused kernel that does not
 * store it (C2: 24 instead of 28 bytes per element). Same parameter updates, same final value. 0: always store. */
void cg_set_elide_root(int on);

/* ---- introspection -------------------------------------------------------------------------------- */
typedef struct {
  unsigned long long tape_instrs, live_instrs, kernels_per_iter, ew_kernels, gemms, gemms_fused_sgd, reduces;
  unsigned long long arena_bytes;
  double gemm_flop_per_iter, ew_bytes_per_iter;
  unsigned long long ew_specialised; /* elementwise kernels compiled to straight-line code (cg_set_ew_jit_min) */
  unsigned long long allreduces;       /* data parallel: weight gradients summed with ncclAllReduce */
  unsigned long long dp_fused_updates; /* data parallel: all-reduce + SGD update as one peer-memory kernel */
  unsigned long long vm_resident;      /* resident executor: 0 none (CUDA graph), 1 interpreting (vm_kernel.cu), 2 compiled */
  unsigned long long tape_smem_bytes;  /* compiled resident kernel: shared memory holding every value of the plan */
  unsigned long long quiet_variant;    /* 1: iterations with an unobservable root value skip its store (cg_set_elide_root) */
  double quiet_ew_bytes_per_iter;      /* elementwise bytes of such an iteration */
  unsigned long long dp_multimem_updates; /* of dp_fused_updates: summed and broadcast inside the NVSwitch (multimem) */
} cg_plan_stats;
int cg_plan_stats_get(cg_function *f, float learn_rate, cg_plan_stats *out); /* compiles the plan if needed */
/* One text line per kernel of the compiled iteration, in launch order ("<i> GEMM m= n= k= ta= tb= epi=", "<i> EW n= in= out=
 * ins=", "<i> AVG", "<i> ALLREDUCE", "<i> DPUPDATE"). Returns the bytes needed (incl. NUL) or -1; buf may be NULL. */
long long cg_plan_describe(cg_function *f, float learn_rate, char *buf, unsigned long long cap);
/* Plan-only mode, for machines without a GPU: graphs and plans are built over placeholder addresses so that the tape
 * compiler can be inspected (cg_plan_stats_get, cg_plan_describe); every call that would compute raises. Must be the
 * first call into the library. */
int cg_set_plan_only(int on);
unsigned long long cg_launch_count(void); /* kernels launched by this library so far */
const char *cg_last_error(void);
/* Times `reps` replays of f's iteration with CUDA events on the library stream; returns ms per iteration. */
double cg_time_iterations(cg_function *f, float learn_rate, int warmup, int reps);

/* One iteration launched eagerly with a CUDA event between every kernel: per-kind device time
 * (index 0 elementwise, 1 reduce, 2 product, 3 all-reduce) and the average duration of the largest product. */
int cg_profile_iteration(cg_function *f, float learn_rate, double kind_ms[4], unsigned long long kind_count[4],
                         double *max_gemm_flop, double *max_gemm_ms_avg, unsigned long long *max_gemm_count);

/* ---- data parallel: one process per GPU, NCCL sum-allreduce of weight gradients ----- */
#define CG_DP_ID_BYTES 128
int cg_dp_unique_id(unsigned char id[CG_DP_ID_BYTES]);                      /* rank 0: ncclGetUniqueId */
int cg_dp_init(int rank, int world, const unsigned char id[CG_DP_ID_BYTES]); /* every rank: ncclCommInitRank */
int cg_dp_shutdown(void);
/* Plan-only mode: pretend to be one of `world` ranks (fused: with the peer-memory exchange available) so that the
 * data-parallel tape can be inspected on a machine without GPUs. */
int cg_dp_plan_only(int world, int fused); /* fused: 0 ncclAllReduce, 1 peer-memory exchange, 2 in-switch exchange */
/* 1 when the fused all-reduce + update kernel over NVLink peer memory (CUDA IPC) is in use for weights of at least
 * CG_DP_FUSED_MIN elements (default 2^18); 0 when the ranks fell back to ncclAllReduce (or CG_DP_FUSED=0). */
int cg_dp_fused(void);
int cg_dp_set_fused_min(long long min_elems); /* weights below this stay on ncclAllReduce; < 0: the default (2^18) */
/* In-switch exchange (NVLS): the weights and gradient slots of a data-parallel plan live in a symmetric multicast heap
 * (VMM memory bound to a CUDA multicast object on every rank) and the fused kernel sums with multimem.ld_reduce and
 * broadcasts with multimem.st — (1 + 1/world) n bytes per GPU and direction instead of 2 (world-1)/world n.
 * cg_dp_multimem_supported: the heap could be created at cg_dp_init (else cg_dp_multimem_reason says why not);
 * cg_dp_multimem: plans built now use it. cg_dp_set_multimem: 0 never, 1 (default) from three ranks up — below that the
 * peer-memory kernel moves fewer bytes — 2 always. Also CG_DP_MULTIMEM=0 (environment, read at cg_dp_init). */
int cg_dp_multimem_supported(void);
int cg_dp_multimem(void);
const char *cg_dp_multimem_reason(void);
int cg_dp_set_multimem(int mode);
int cg_set_device(int device);

#ifdef __cplusplus
}
#endif
#endif /* CG_GRAPH_H */
