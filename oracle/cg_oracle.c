/*
 * oracle/cg_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the hot path of ethanabrooks/computational-graph:
 *   Function tree  ->  assign_values (forward)  ->  backprop (+ in-line SGD)  ->  minimize loop.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under computational-graph_b200/ links, imports or calls it.
 *
 * Parity pinning: (1) every matrix op can be routed through the reference's own CPU backend
 * (src/cpu/{matrix,ops,util}.cpp compiled unmodified into oracle/_ref/libmatrix_ref_*.so, see
 * oracle/Makefile) with orc_use_ref(); (2) the built-in "port" arithmetic below is checked
 * symbol-by-symbol against that library; (3) the walker is pinned by the reference's 15 threshold
 * tests (src/main.rs:85-166) and the C1 golden trajectory.
 *
 * Each function cites the reference file:line it restates (paths relative to the reference root).
 * Compile with -ffp-contract=off: the reference is built without FMA contraction (build.rs:54-61).
 */
#include <dlfcn.h>
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * L2: the 26-symbol matrix backend.  Struct layout == src/cpu/matrix.h:4-8 == src/matrix.rs:13-18.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  unsigned height;
  unsigned width;
  float *array; /* column-major storage (src/cpu/matrix.cpp:26-36) */
} Matrix;

typedef struct {
  void (*alloc_matrix)(Matrix *);
  void (*free_matrix)(Matrix *);
  void (*fill_matrix)(Matrix *, float);
  void (*copy_matrix)(const Matrix *, Matrix *);
  void (*get_array)(const Matrix *, float *);
  void (*set_array)(Matrix *, const float *, bool);
  void (*map[7])(const Matrix *, Matrix *);                   /* neg sq abs signum sigmoid tanh one_minus */
  void (*elemwise[3])(const Matrix *, const Matrix *, Matrix *); /* add sub mul */
  void (*broadcast[3])(const Matrix *, float, Matrix *);
  void (*broadcast_rev[3])(float, const Matrix *, Matrix *);
  void (*gemm)(const Matrix *, bool, const Matrix *, bool, Matrix *);
  bool (*all_equal)(const Matrix *, float);
  bool (*all_less_than)(const Matrix *, float);
  float (*reduce_sum)(const Matrix *);
} MatOps;

enum { U_NEG, U_SQ, U_ABS, U_SIGNUM, U_SIGMOID, U_TANH, U_ONE_MINUS, U_COUNT };
enum { B_ADD, B_SUB, B_MUL, B_COUNT };

static void die(const char *msg) { /* src/cpu/util.cpp:5-10 */
  fprintf(stderr, "ERROR: %s\n", msg);
  exit(EXIT_FAILURE);
}

static int msize(const Matrix *m) { return (int)(m->width * m->height); } /* matrix.cpp:15 */

/* --- port of src/cpu/matrix.cpp ---------------------------------------------------------------- */
static void p_alloc_matrix(Matrix *m) { /* matrix.cpp:18-20, util.h:7-12 */
  size_t n = (size_t)m->width * m->height;
  m->array = (float *)malloc((n ? n : 1) * sizeof(float));
  if (!m->array) die("safe_malloc failed");
}
static void p_free_matrix(Matrix *m) { free(m->array); } /* matrix.cpp:69-71 */
static void p_fill_matrix(Matrix *m, float v) {          /* matrix.cpp:38-43 */
  size_t n = (size_t)m->width * m->height;
  for (size_t i = 0; i < n; i++) m->array[i] = v;
}
static void p_get_array(const Matrix *m, float *dst) { /* matrix.cpp:22-24 */
  memcpy(dst, m->array, (size_t)m->width * m->height * sizeof(float));
}
static void p_set_array(Matrix *m, const float *src, bool transpose) { /* matrix.cpp:26-36 */
  size_t n = (size_t)m->width * m->height;
  if (transpose) { /* src is row-major; storage becomes column-major */
    for (size_t i = 0; i < n; i++) {
      size_t iT = (i % m->height) * m->width + i / m->height;
      m->array[i] = src[iT];
    }
  } else {
    memcpy(m->array, src, n * sizeof(float));
  }
}
static void p_copy_matrix(const Matrix *src, Matrix *dst) { /* matrix.cpp:53-57 */
  if (dst->height != src->height || dst->width != src->width) die("copy_matrix: dims differ");
  p_set_array(dst, src->array, false);
}

/* --- port of src/cpu/ops.cpp ------------------------------------------------------------------- */
#define CHECK_EQ(a, b, msg) \
  if ((a) != (b)) die(msg)

/* ops.cpp:68-74 (UN_MAP bodies) */
static float f_neg(float x) { return -x; }
static float f_sq(float x) { return x * x; }
static float f_abs(float x) { return x < 0 ? -x : x; }        /* abs(-0) stays -0 (Q8) */
static float f_signum(float x) { return x < 0 ? -1 : 1; }     /* signum(+-0)=+1, signum(NaN)=+1 (Q8) */
static float f_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
static float f_tanh(float x) { return tanhf(x); }             /* `tanh(x)` on float -> tanhf */
static float f_one_minus(float x) { return 1.0f - x; }
static float (*const UNARY_F[U_COUNT])(float) = {f_neg, f_sq, f_abs, f_signum, f_sigmoid, f_tanh, f_one_minus};

#define DEF_MAP(name, idx)                                                   \
  static void p_map_##name(const Matrix *m, Matrix *r) { /* ops.cpp:11-20 */ \
    CHECK_EQ(m->height, r->height, "m->height must equal result->height");   \
    CHECK_EQ(m->width, r->width, "m->width must equal result->width");       \
    int n = msize(m);                                                        \
    for (int i = 0; i < n; i++) r->array[i] = UNARY_F[idx](m->array[i]);     \
  }
DEF_MAP(neg, U_NEG)
DEF_MAP(sq, U_SQ)
DEF_MAP(abs, U_ABS)
DEF_MAP(signum, U_SIGNUM)
DEF_MAP(sigmoid, U_SIGMOID)
DEF_MAP(tanh, U_TANH)
DEF_MAP(one_minus, U_ONE_MINUS)

static float b_add(float a, float b) { return a + b; } /* ops.cpp:76-86 */
static float b_sub(float a, float b) { return a - b; }
static float b_mul(float a, float b) { return a * b; }
static float (*const BIN_F[B_COUNT])(float, float) = {b_add, b_sub, b_mul};

#define DEF_BIN(name, idx)                                                                            \
  static void p_elemwise_##name(const Matrix *a, const Matrix *b, Matrix *r) { /* ops.cpp:42-55 */    \
    CHECK_EQ(a->height, b->height, "m1->height must equal m2->height");                               \
    CHECK_EQ(a->width, b->width, "m1->width must equal m2->width");                                   \
    CHECK_EQ(a->height, r->height, "m1->height must equal result->height");                           \
    CHECK_EQ(a->width, r->width, "m1->width must equal result->width");                               \
    int n = msize(a);                                                                                 \
    for (int i = 0; i < n; i++) r->array[i] = BIN_F[idx](a->array[i], b->array[i]);                   \
  }                                                                                                   \
  static void p_broadcast_##name(const Matrix *m, float v, Matrix *r) { /* ops.cpp:22-32 */           \
    CHECK_EQ(m->height, r->height, "m->height must equal result->height");                            \
    CHECK_EQ(m->width, r->width, "m->width must equal result->width");                                \
    int n = msize(m);                                                                                 \
    for (int i = 0; i < n; i++) r->array[i] = BIN_F[idx](m->array[i], v);                             \
  }                                                                                                   \
  static void p_broadcast_##name##_rev(float v, const Matrix *m, Matrix *r) { /* ops.cpp:34-43 */     \
    CHECK_EQ(m->height, r->height, "m->height must equal result->height");                            \
    CHECK_EQ(m->width, r->width, "m->width must equal result->width");                                \
    int n = msize(m);                                                                                 \
    for (int i = 0; i < n; i++) r->array[i] = BIN_F[idx](v, m->array[i]);                             \
  }
DEF_BIN(add, B_ADD)
DEF_BIN(sub, B_SUB)
DEF_BIN(mul, B_MUL)

/* ops.cpp:88-116, generalised to non-square operands (the reference reads m1's dims where it means
 * m2's and uses result->width as the leading dimension, so it only works when every operand is the
 * same square size — Q7).  For equal square operands this is the reference's arithmetic exactly:
 * result zeroed, then r[i,j] += a[i,k]*b[k,j] for k ascending, one f32 rounding per multiply and one
 * per add, column-major everywhere. */
static void p_gemm(const Matrix *m1, bool t1, const Matrix *m2, bool t2, Matrix *r) {
  int h1 = t1 ? m1->width : m1->height;
  int inner = t1 ? m1->height : m1->width;
  int h2 = t2 ? m2->width : m2->height;
  int w2 = t2 ? m2->height : m2->width;
  CHECK_EQ(h1, (int)r->height, "height1 must equal result->height");
  CHECK_EQ(inner, h2, "inner_dim must equal height2");
  CHECK_EQ(w2, (int)r->width, "width2 must equal result->width");
  for (int j = 0; j < w2; j++) {
    for (int i = 0; i < h1; i++) {
      float acc = 0.0f;
      for (int k = 0; k < inner; k++) {
        float a = t1 ? m1->array[(size_t)i * m1->height + k] : m1->array[(size_t)k * m1->height + i];
        float b = t2 ? m2->array[(size_t)k * m2->height + j] : m2->array[(size_t)j * m2->height + k];
        float p = a * b;
        acc = acc + p;
      }
      r->array[(size_t)j * r->height + i] = acc;
    }
  }
}
static bool p_all_equal(const Matrix *m, float x) { /* ops.cpp:118-126 */
  int n = msize(m);
  for (int i = 0; i < n; i++)
    if (m->array[i] != x) return false;
  return true;
}
static bool p_all_less_than(const Matrix *m, float x) { /* ops.cpp:128-136 */
  int n = msize(m);
  for (int i = 0; i < n; i++)
    if (m->array[i] >= x) return false;
  return true;
}
static float p_reduce_sum(const Matrix *m) { /* ops.cpp:138-145: sequential f32 sum */
  float sum = 0;
  int n = msize(m);
  for (int i = 0; i < n; i++) sum += m->array[i];
  return sum;
}

static const MatOps PORT_OPS = {
    p_alloc_matrix, p_free_matrix, p_fill_matrix, p_copy_matrix, p_get_array, p_set_array,
    {p_map_neg, p_map_sq, p_map_abs, p_map_signum, p_map_sigmoid, p_map_tanh, p_map_one_minus},
    {p_elemwise_add, p_elemwise_sub, p_elemwise_mul},
    {p_broadcast_add, p_broadcast_sub, p_broadcast_mul},
    {p_broadcast_add_rev, p_broadcast_sub_rev, p_broadcast_mul_rev},
    p_gemm, p_all_equal, p_all_less_than, p_reduce_sum};

static MatOps OPS;
static int OPS_INIT = 0;
static int USING_REF = 0;
static void *REF_HANDLE = NULL;

static void ensure_ops(void) {
  if (!OPS_INIT) {
    OPS = PORT_OPS;
    OPS_INIT = 1;
  }
}

void orc_use_port(void) {
  OPS = PORT_OPS;
  OPS_INIT = 1;
  USING_REF = 0;
}

/* Route every matrix op through the reference-compiled library (oracle/_ref/libmatrix_ref_*.so). */
int orc_use_ref(const char *path) {
  void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "orc_use_ref: %s\n", dlerror());
    return -1;
  }
  MatOps o;
#define SYM(field, name)                        \
  *(void **)(&o.field) = dlsym(h, name);        \
  if (!o.field) {                               \
    fprintf(stderr, "missing symbol %s\n", name); \
    return -2;                                  \
  }
  SYM(alloc_matrix, "alloc_matrix") SYM(free_matrix, "free_matrix") SYM(fill_matrix, "fill_matrix")
  SYM(copy_matrix, "copy_matrix") SYM(get_array, "get_array") SYM(set_array, "set_array")
  SYM(map[U_NEG], "map_neg") SYM(map[U_SQ], "map_sq") SYM(map[U_ABS], "map_abs")
  SYM(map[U_SIGNUM], "map_signum") SYM(map[U_SIGMOID], "map_sigmoid") SYM(map[U_TANH], "map_tanh")
  SYM(map[U_ONE_MINUS], "map_one_minus")
  SYM(elemwise[B_ADD], "elemwise_add") SYM(elemwise[B_SUB], "elemwise_sub") SYM(elemwise[B_MUL], "elemwise_mul")
  SYM(broadcast[B_ADD], "broadcast_add") SYM(broadcast[B_SUB], "broadcast_sub") SYM(broadcast[B_MUL], "broadcast_mul")
  SYM(broadcast_rev[B_ADD], "broadcast_add_rev") SYM(broadcast_rev[B_SUB], "broadcast_sub_rev")
  SYM(broadcast_rev[B_MUL], "broadcast_mul_rev")
  SYM(gemm, "gemm") SYM(all_equal, "all_equal") SYM(all_less_than, "all_less_than") SYM(reduce_sum, "reduce_sum")
#undef SYM
  OPS = o;
  OPS_INIT = 1;
  USING_REF = 1;
  REF_HANDLE = h;
  return 0;
}

int orc_using_ref(void) { return USING_REF; }

/* gemm dispatch: the reference gemm is only valid for equal square operands (Q7); anything else goes
 * to the corrected port (the reference would exit(1) or index out of bounds). */
static void o_gemm(const Matrix *m1, bool t1, const Matrix *m2, bool t2, Matrix *r) {
  bool square = m1->height == m1->width && m2->height == m2->width && r->height == r->width &&
                m1->height == m2->height && m1->height == r->height;
  if (USING_REF && square)
    OPS.gemm(m1, t1, m2, t2, r);
  else
    p_gemm(m1, t1, m2, t2, r);
}

/* Raw access to the active backend for the symbol-level tests (tests/test_oracle_symbols.py). */
void orc_sym_map(int op, const Matrix *m, Matrix *r) { ensure_ops(); OPS.map[op](m, r); }
void orc_sym_elemwise(int op, const Matrix *a, const Matrix *b, Matrix *r) { ensure_ops(); OPS.elemwise[op](a, b, r); }
void orc_sym_broadcast(int op, const Matrix *m, float v, Matrix *r) { ensure_ops(); OPS.broadcast[op](m, v, r); }
void orc_sym_broadcast_rev(int op, float v, const Matrix *m, Matrix *r) { ensure_ops(); OPS.broadcast_rev[op](v, m, r); }
void orc_sym_gemm(const Matrix *a, bool t1, const Matrix *b, bool t2, Matrix *r) { ensure_ops(); o_gemm(a, t1, b, t2, r); }
void orc_sym_set_array(Matrix *m, const float *src, bool tr) { ensure_ops(); OPS.set_array(m, src, tr); }
void orc_sym_get_array(const Matrix *m, float *dst) { ensure_ops(); OPS.get_array(m, dst); }
void orc_sym_fill(Matrix *m, float v) { ensure_ops(); OPS.fill_matrix(m, v); }
void orc_sym_copy(const Matrix *s, Matrix *d) { ensure_ops(); OPS.copy_matrix(s, d); }
float orc_sym_reduce_sum(const Matrix *m) { ensure_ops(); return OPS.reduce_sum(m); }
bool orc_sym_all_equal(const Matrix *m, float x) { ensure_ops(); return OPS.all_equal(m, x); }
bool orc_sym_all_less_than(const Matrix *m, float x) { ensure_ops(); return OPS.all_less_than(m, x); }

/* ------------------------------------------------------------------------------------------------
 * L3: Constant (src/constant.rs:5-8) and Matrix RAII helpers (src/matrix.rs:49-106)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int is_matrix;
  int defined; /* oracle rule: Dot/Input-derived values are "undefined": never match an identity */
  float s;
  Matrix m;
} Constant;

static Constant c_scalar(float x) {
  Constant c;
  memset(&c, 0, sizeof c);
  c.defined = 1;
  c.s = x;
  return c;
}
static Constant c_empty(int is_matrix, unsigned h, unsigned w) { /* constant.rs:115-121, matrix.rs:49-57 */
  Constant c;
  memset(&c, 0, sizeof c);
  c.is_matrix = is_matrix;
  c.defined = 1;
  if (is_matrix) {
    c.m.height = h;
    c.m.width = w;
    OPS.alloc_matrix(&c.m);
    /* the reference leaves the contents uninitialised; zero them so runs are reproducible */
    memset(c.m.array, 0, (size_t)h * w * sizeof(float));
  }
  return c;
}
static Constant c_empty_like(const Constant *o) { return c_empty(o->is_matrix, o->m.height, o->m.width); }
static Constant c_clone(const Constant *o) { /* #[derive(Clone)] + matrix.rs:100-106 */
  Constant c = c_empty_like(o);
  c.defined = o->defined;
  c.s = o->s;
  if (o->is_matrix) OPS.copy_matrix(&o->m, &c.m);
  return c;
}
static void c_free(Constant *c) {
  if (c->is_matrix && c->m.array) OPS.free_matrix(&c->m);
  c->m.array = NULL;
}
/* Corrected-semantics switches (defaults reproduce the reference):
 *   FIX_BROADCAST_SUM: in the BACKWARD pass a scalar that absorbs a matrix-shaped error takes its SUM — the gradient of
 *     the loss (root error == ones, i.e. the sum of the root's elements) with respect to a broadcast scalar — instead
 *     of the mean the reference's S<-M rule gives it (Q5; ops.rs:79-80,103-104,126-134). The forward pass keeps the mean:
 *     there it is the value-shape rule, not a gradient.
 *   FIX_SUB_SIGN: `0 - x` is `-x`; the reference shares Add's identity rule and returns `x` (Q4; ops.rs:255-263).
 *   FIX_TIED_PARAMS: every use of a parameter NAME is the same parameter. The reference gives each use site a private
 *     copy that drifts apart (Q1; function.rs:10-16). Here the copies stay, but the backward walk only ACCUMULATES the
 *     error each copy receives (reduced to the parameter's shape) into one gradient per name — every forward value it
 *     reads is therefore the un-updated one — and after the walk every copy of the name takes the same update
 *     p -= lr * (sum of the contributions), so the copies remain bit-identical: one parameter, true gradient descent. */
static int FIX_BROADCAST_SUM = 0, FIX_SUB_SIGN = 0, IN_BACKWARD = 0, FIX_TIED_PARAMS = 0;
static float m_avg(const Matrix *m) { /* ops.rs:438-442 */
  if (FIX_BROADCAST_SUM && IN_BACKWARD) return OPS.reduce_sum(m);
  return OPS.reduce_sum(m) / (float)((size_t)m->height * m->width);
}
static float c_avg(const Constant *c) { return c->is_matrix ? m_avg(&c->m) : c->s; } /* ops.rs:430-435 */

static void panic_msg(const char *msg) {
  fprintf(stderr, "panic: %s\n", msg);
  exit(101); /* Rust panic exit status */
}

/* Rust scalar closures (ops.rs:167, 331-336): f32::abs / f32::signum / x*x / f32::tanh / 1/(1+exp(-x)) */
static float rs_signum(float x) {
  if (isnan(x)) return NAN;
  return signbit(x) ? -1.0f : 1.0f;
}
static float scalar_unary(int op, float x) {
  switch (op) {
    case U_NEG: return -x;
    case U_SQ: return x * x;
    case U_ABS: return fabsf(x);
    case U_SIGNUM: return rs_signum(x);
    case U_SIGMOID: return 1.0f / (1.0f + expf(-x));
    case U_TANH: return tanhf(x);
    case U_ONE_MINUS: return 1.0f - x;
  }
  return x;
}
static int unary_nonlinear(int op) { return op == U_SQ || op == U_ABS || op == U_SIGNUM || op == U_SIGMOID || op == U_TANH; }

/* Constant1!(linear|nonlinear: op, result, arg, f)  — ops.rs:70-99 */
static void c_equals_unary(int op, Constant *r, const Constant *a) {
  if (unary_nonlinear(op) && !r->is_matrix && a->is_matrix)
    panic_msg("The result of this operation yields a matrix. Can't place a matrix result in a scalar");
  if (!r->is_matrix && !a->is_matrix) r->s = scalar_unary(op, a->s);
  else if (r->is_matrix && !a->is_matrix) OPS.fill_matrix(&r->m, scalar_unary(op, a->s));
  else if (r->is_matrix && a->is_matrix) OPS.map[op](&a->m, &r->m);
  else r->s = scalar_unary(op, m_avg(&a->m));
}
/* Constant1!(op, c, f) in place — ops.rs:62-68 */
static void c_assign_unary(int op, Constant *c) {
  if (c->is_matrix) OPS.map[op](&c->m, &c->m);
  else c->s = scalar_unary(op, c->s);
}
/* Constant2!(linear: op, result, other, f)  (op_assign) — ops.rs:102-116 */
static void c_op_assign(int op, Constant *r, const Constant *o) {
  if (!r->is_matrix && !o->is_matrix) r->s = BIN_F[op](r->s, o->s);
  else if (r->is_matrix && !o->is_matrix) OPS.broadcast[op](&r->m, o->s, &r->m);
  else if (!r->is_matrix && o->is_matrix) r->s = BIN_F[op](r->s, m_avg(&o->m));
  else OPS.elemwise[op](&r->m, &o->m, &r->m);
}
/* Constant2!(linear: op, result, c1, c2, f)  (equals_op) — ops.rs:118-148 */
static void c_equals_bin(int op, Constant *r, const Constant *a, const Constant *b) {
  if (!r->is_matrix) {
    float x1 = a->is_matrix ? m_avg(&a->m) : a->s;
    float x2 = b->is_matrix ? m_avg(&b->m) : b->s;
    r->s = BIN_F[op](x1, x2);
  } else if (!a->is_matrix && !b->is_matrix) {
    OPS.fill_matrix(&r->m, BIN_F[op](a->s, b->s));
  } else if (a->is_matrix && !b->is_matrix) {
    OPS.broadcast[op](&a->m, b->s, &r->m);
  } else if (!a->is_matrix && b->is_matrix) {
    OPS.broadcast_rev[op](a->s, &b->m, &r->m);
  } else {
    OPS.elemwise[op](&a->m, &b->m, &r->m);
  }
}
/* constant.rs:106-113 */
static void c_absorb(Constant *r, const Constant *o) {
  if (!r->is_matrix && !o->is_matrix) r->s = o->s;
  else if (r->is_matrix && o->is_matrix) OPS.copy_matrix(&o->m, &r->m);
  else if (!r->is_matrix && o->is_matrix) r->s = c_avg(o);
  else OPS.fill_matrix(&r->m, o->s);
}
/* constant.rs:85-91 */
static Constant c_copy_and_fill(const Constant *c, float v) {
  if (!c->is_matrix) return c_scalar(v);
  Constant r = c_empty(1, c->m.height, c->m.width);
  OPS.fill_matrix(&r.m, v);
  return r;
}
/* ops.rs:411-422 */
static void c_equals_dot(Constant *r, const Constant *a, int t1, const Constant *b, int t2) {
  if (!(r->is_matrix && a->is_matrix && b->is_matrix)) panic_msg("dot should not be used with scalars");
  o_gemm(&a->m, t1 != 0, &b->m, t2 != 0, &r->m);
}
static bool c_all_equal(const Constant *c, float v) { /* ops.rs:309-326 */
  if (!c->defined) return false;
  return c->is_matrix ? OPS.all_equal(&c->m, v) : c->s == v;
}
static bool c_all_less_than(const Constant *c, float v) {
  return c->is_matrix ? OPS.all_less_than(&c->m, v) : c->s < v;
}

/* ------------------------------------------------------------------------------------------------
 * L4: Function / Expr / Pool  (src/function.rs:7-44, 58-98)
 * ---------------------------------------------------------------------------------------------- */
enum { K_CONSTANT, K_INPUT, K_PARAM, K_NEG, K_SQ, K_ABS, K_SIGNUM, K_SIGMOID, K_TANH, K_ADD, K_SUB, K_MUL, K_DOT };

typedef struct Expr Expr;
typedef struct Function {
  Constant value;   /* RefCell<Constant>: deep-copied on clone */
  int has_params;   /* !params.is_empty() */
  Expr *body;       /* Rc<Expr>: shared on clone */
  int pool_n;       /* Pool: deep-copied on clone */
  Constant pool[2];
} Function;

struct Expr {
  int refcnt;
  int kind;
  char name[64];
  Function *f1, *f2; /* children are owned by the Expr (by-value fields in the reference) */
  int t1, t2;
  /* dag-mode scratch */
  long mark;
  Function *canon;
  Constant err;
  int has_err;
};

static long EPOCH = 1;
static int MODE = 0;            /* 0 = exact (tree walk, the reference), 1 = dag */
static int FIX_ACT_GRAD = 0;    /* 0 = reproduce Q3 */
static long COUNT_GEMM = 0, COUNT_OPS = 0;

void orc_set_mode(int m) { MODE = m; }
void orc_set_fix_act_grad(int v) { FIX_ACT_GRAD = v; }
void orc_set_fix_broadcast_sum(int v) { FIX_BROADCAST_SUM = v; }
void orc_set_fix_sub_sign(int v) { FIX_SUB_SIGN = v; }
void orc_set_fix_tied_params(int v) { FIX_TIED_PARAMS = v; }
long orc_count_gemm(void) { return COUNT_GEMM; }
void orc_reset_counts(void) { COUNT_GEMM = COUNT_OPS = 0; }

static int unary_kind_to_op(int k) {
  switch (k) {
    case K_NEG: return U_NEG;
    case K_SQ: return U_SQ;
    case K_ABS: return U_ABS;
    case K_SIGNUM: return U_SIGNUM;
    case K_SIGMOID: return U_SIGMOID;
    case K_TANH: return U_TANH;
  }
  return -1;
}
static int binary_kind_to_op(int k) { return k == K_ADD ? B_ADD : k == K_SUB ? B_SUB : B_MUL; }

/* Function::new (function.rs:87-98): pool of `pool_n` Constants shaped like `value` */
static Function *fn_new(Constant value, int has_params, Expr *body, int pool_n) {
  Function *f = (Function *)calloc(1, sizeof(Function));
  f->value = value;
  f->has_params = has_params;
  f->body = body;
  f->pool_n = pool_n;
  for (int i = 0; i < pool_n; i++) f->pool[i] = c_empty_like(&value);
  return f;
}
static Expr *expr_new(int kind, const char *name, Function *f1, Function *f2, int t1, int t2) {
  Expr *e = (Expr *)calloc(1, sizeof(Expr));
  e->refcnt = 1;
  e->kind = kind;
  if (name) snprintf(e->name, sizeof e->name, "%s", name);
  e->f1 = f1;
  e->f2 = f2;
  e->t1 = t1;
  e->t2 = t2;
  return e;
}
/* #[derive(Clone)] for Function (function.rs:10-16): value and pool deep-copied, body shared (Q1) */
Function *orc_clone(const Function *f) {
  ensure_ops();
  Function *g = (Function *)calloc(1, sizeof(Function));
  g->value = c_clone(&f->value);
  g->has_params = f->has_params;
  g->body = f->body;
  g->body->refcnt++;
  g->pool_n = f->pool_n;
  for (int i = 0; i < f->pool_n; i++) g->pool[i] = c_clone(&f->pool[i]);
  return g;
}
void orc_release(Function *f);
static void expr_release(Expr *e) {
  if (--e->refcnt > 0) return;
  if (e->f1) orc_release(e->f1);
  if (e->f2) orc_release(e->f2);
  if (e->has_err) c_free(&e->err);
  free(e);
}
void orc_release(Function *f) {
  if (!f) return;
  c_free(&f->value);
  for (int i = 0; i < f->pool_n; i++) c_free(&f->pool[i]);
  expr_release(f->body);
  free(f);
}

/* --- leaf constructors (function.rs:100-155) --- */
static Constant c_matrix_from_rows(unsigned h, unsigned w, const float *vals) { /* matrix.rs:59-66 */
  Constant c = c_empty(1, h, w);
  OPS.set_array(&c.m, vals, true);
  return c;
}
Function *orc_scalar_param(const char *name, float v) {
  ensure_ops();
  return fn_new(c_scalar(v), 1, expr_new(K_PARAM, name, NULL, NULL, 0, 0), 0);
}
Function *orc_matrix_param(const char *name, unsigned h, unsigned w, const float *vals) {
  ensure_ops();
  return fn_new(c_matrix_from_rows(h, w, vals), 1, expr_new(K_PARAM, name, NULL, NULL, 0, 0), 0);
}
Function *orc_single_val_param(const char *name, unsigned h, unsigned w, float v) {
  ensure_ops();
  Constant c = c_empty(1, h, w);
  OPS.fill_matrix(&c.m, v);
  return fn_new(c, 1, expr_new(K_PARAM, name, NULL, NULL, 0, 0), 0);
}
Function *orc_scalar(float v) {
  ensure_ops();
  return fn_new(c_scalar(v), 0, expr_new(K_CONSTANT, NULL, NULL, NULL, 0, 0), 0);
}
Function *orc_matrix(unsigned h, unsigned w, const float *vals) {
  ensure_ops();
  return fn_new(c_matrix_from_rows(h, w, vals), 0, expr_new(K_CONSTANT, NULL, NULL, NULL, 0, 0), 0);
}
Function *orc_single_val_matrix(unsigned h, unsigned w, float v) {
  ensure_ops();
  Constant c = c_empty(1, h, w);
  OPS.fill_matrix(&c.m, v);
  return fn_new(c, 0, expr_new(K_CONSTANT, NULL, NULL, NULL, 0, 0), 0);
}
/* Function::input (function.rs:105-108): params = {name}; value = empty(dims) (undefined) */
Function *orc_input(const char *name, int is_matrix, unsigned h, unsigned w) {
  ensure_ops();
  Constant c = c_empty(is_matrix, h, w);
  c.defined = 0;
  return fn_new(c, 1, expr_new(K_INPUT, name, NULL, NULL, 0, 0), 0);
}

/* --- op constructors (ops.rs:36-54, 199-264, 331-343, 363-376, 395-407) --- */
static const float UNARY_IDENTITY[] = {/*neg*/ 0.f, /*sq*/ 0.f, /*abs*/ 0.f, /*signum*/ 0.f, /*sigmoid*/ 0.5f, /*tanh*/ 0.f};
static const int UNARY_POOL[] = {0, 0, 1, 0, 1, 1};

static Function *make_unary(int kind, const Function *f) {
  ensure_ops();
  int op = unary_kind_to_op(kind);
  if (c_all_equal(&f->value, UNARY_IDENTITY[op])) return orc_clone(f); /* Function1! identity arm, ops.rs:45-53 */
  Expr *e = expr_new(kind, NULL, orc_clone(f), NULL, 0, 0);
  return fn_new(c_clone(&f->value), f->has_params, e, UNARY_POOL[op]); /* ops.rs:38-43 */
}
Function *orc_neg(const Function *f) { return make_unary(K_NEG, f); }
Function *orc_sq(const Function *f) { return make_unary(K_SQ, f); }
Function *orc_abs(const Function *f) { return make_unary(K_ABS, f); }
Function *orc_signum(const Function *f) { return make_unary(K_SIGNUM, f); }
Function *orc_sigmoid(const Function *f) { return make_unary(K_SIGMOID, f); }
Function *orc_tanh(const Function *f) { return make_unary(K_TANH, f); }

static void forward(Function *f, int nargs, const char **names, const Constant *args);

static Function *make_binary(int kind, const Function *a, const Function *b) {
  ensure_ops();
  float identity = kind == K_MUL ? 1.0f : 0.0f; /* ops.rs:338-343; Sub shares Add's rule: 0 - x -> x (Q4) */
  if (FIX_SUB_SIGN && kind == K_SUB && c_all_equal(&a->value, identity)) return orc_neg(b);
  if (c_all_equal(&a->value, identity)) return orc_clone(b); /* ops.rs:255-259 */
  if (c_all_equal(&b->value, identity)) return orc_clone(a);
  /* value-shape rule, ops.rs:218-228 */
  Constant value;
  if (!a->value.is_matrix && !b->value.is_matrix) value = c_scalar(a->value.s), value.defined = a->value.defined;
  else if (a->value.is_matrix && !b->value.is_matrix) value = c_clone(&a->value);
  else if (!a->value.is_matrix && b->value.is_matrix) value = c_clone(&b->value);
  else {
    if (a->value.m.height != b->value.m.height || a->value.m.width != b->value.m.width)
      panic_msg("assertion failed: m1.dims() == m2.dims()");
    value = c_clone(&a->value);
  }
  Expr *e = expr_new(kind, NULL, orc_clone(a), orc_clone(b), 0, 0);
  Function *f = fn_new(value, a->has_params || b->has_params, e, 1);
  if (a->body->kind == K_CONSTANT && b->body->kind == K_CONSTANT) {
    /* const (+) const folding, ops.rs:209-213.  The reference folds through `eval`, whose
     * Constant (+) Constant computes `empty (+) rhs` (Q9, a bug); the oracle folds the mathematics. */
    forward(f, 0, NULL, NULL);
    Function *folded = fn_new(c_clone(&f->value), 0, expr_new(K_CONSTANT, NULL, NULL, NULL, 0, 0), 0);
    folded->value.defined = 1;
    orc_release(f);
    return folded;
  }
  return f;
}
Function *orc_add(const Function *a, const Function *b) { return make_binary(K_ADD, a, b); }
Function *orc_sub(const Function *a, const Function *b) { return make_binary(K_SUB, a, b); }
Function *orc_mul(const Function *a, const Function *b) { return make_binary(K_MUL, a, b); }

/* ops.rs:395-407 + matrix.rs:68-83 (empty_for_dot) */
Function *orc_dot(const Function *a, int t1, const Function *b, int t2) {
  ensure_ops();
  if (!a->value.is_matrix || !b->value.is_matrix) panic_msg("dot should not be used with scalars");
  unsigned h = t1 ? a->value.m.width : a->value.m.height;
  unsigned w = t2 ? b->value.m.height : b->value.m.width;
  Constant value = c_empty(1, h, w);
  Expr *e = expr_new(K_DOT, NULL, orc_clone(a), orc_clone(b), t1, t2);
  Function *f = fn_new(value, a->has_params || b->has_params, e, 2);
  if (a->body->kind == K_CONSTANT && b->body->kind == K_CONSTANT) {
    forward(f, 0, NULL, NULL);
    Function *folded = fn_new(c_clone(&f->value), 0, expr_new(K_CONSTANT, NULL, NULL, NULL, 0, 0), 0);
    orc_release(f);
    return folded;
  }
  f->value.defined = 0; /* uninitialised in the reference */
  /* Dot pool: the reference sizes both placeholders like the result (ops.rs:405, Q7); the oracle
   * sizes them for the gradients they hold (dims of f1 / f2), which coincides for square operands. */
  c_free(&f->pool[0]);
  c_free(&f->pool[1]);
  f->pool[0] = c_empty_like(&a->value);
  f->pool[1] = c_empty_like(&b->value);
  return f;
}

/* ------------------------------------------------------------------------------------------------
 * optimize.rs: assign_values (forward), backprop, minimize
 * ---------------------------------------------------------------------------------------------- */
static const Constant *find_arg(const char *name, int nargs, const char **names, const Constant *args) {
  for (int i = 0; i < nargs; i++)
    if (strcmp(names[i], name) == 0) return &args[i];
  panic_msg("missing arg");
  return NULL;
}

/* the node's own op applied to already-evaluated children; optimize.rs:86-137 */
static void forward_node(Function *f, int nargs, const char **names, const Constant *args) {
  Expr *e = f->body;
  switch (e->kind) {
    case K_CONSTANT:
    case K_PARAM: return;
    case K_INPUT: { /* optimize.rs:87-90 + constant.rs:94-103 */
      const Constant *a = find_arg(e->name, nargs, names, args);
      if (a->is_matrix != f->value.is_matrix) panic_msg("Cannot copy mismatched Constants.");
      c_absorb(&f->value, a);
      return;
    }
    case K_NEG: case K_SQ: case K_ABS: case K_SIGNUM: case K_SIGMOID: case K_TANH:
      c_equals_unary(unary_kind_to_op(e->kind), &f->value, &e->f1->value);
      COUNT_OPS++;
      return;
    case K_ADD: case K_SUB: case K_MUL:
      c_equals_bin(binary_kind_to_op(e->kind), &f->value, &e->f1->value, &e->f2->value);
      COUNT_OPS++;
      return;
    case K_DOT:
      c_equals_dot(&f->value, &e->f1->value, e->t1, &e->f2->value, e->t2);
      COUNT_GEMM++;
      return;
  }
}

/* exact mode: the literal un-memoised post-order recursion of optimize.rs:84-138 */
static void forward_tree(Function *f, int nargs, const char **names, const Constant *args) {
  Expr *e = f->body;
  if (e->f1) forward_tree(e->f1, nargs, names, args);
  if (e->f2) forward_tree(e->f2, nargs, names, args);
  forward_node(f, nargs, names, args);
}

/* dag mode: each shared body is evaluated once; further clones copy the cached value */
typedef struct { Function **v; int n, cap; } FnList;
static void fl_push(FnList *l, Function *f) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 64;
    l->v = (Function **)realloc(l->v, (size_t)l->cap * sizeof(Function *));
  }
  l->v[l->n++] = f;
}
static void forward_dag(Function *f, int nargs, const char **names, const Constant *args, FnList *order) {
  Expr *e = f->body;
  if (e->kind == K_CONSTANT || e->kind == K_PARAM) return;
  if (e->mark == EPOCH) {
    if (e->canon != f) c_absorb(&f->value, &e->canon->value);
    return;
  }
  if (e->f1) forward_dag(e->f1, nargs, names, args, order);
  if (e->f2) forward_dag(e->f2, nargs, names, args, order);
  forward_node(f, nargs, names, args);
  e->mark = EPOCH;
  e->canon = f;
  if (order) fl_push(order, f);
}
static void forward(Function *f, int nargs, const char **names, const Constant *args) {
  forward_tree(f, nargs, names, args);
}

/* FIX_TIED_PARAMS: one accumulated gradient per parameter name, and the leaves (use sites) that carry that name */
typedef struct {
  char name[64];
  Constant grad;
  int has_grad;
  Function **leaves;
  int n, cap;
} TieGroup;
static TieGroup *TIES = NULL;
static int N_TIES = 0, CAP_TIES = 0;
static TieGroup *tie_group(const char *name) {
  for (int i = 0; i < N_TIES; i++)
    if (strcmp(TIES[i].name, name) == 0) return &TIES[i];
  if (N_TIES == CAP_TIES) {
    CAP_TIES = CAP_TIES ? CAP_TIES * 2 : 16;
    TIES = (TieGroup *)realloc(TIES, (size_t)CAP_TIES * sizeof(TieGroup));
  }
  TieGroup *g = &TIES[N_TIES++];
  memset(g, 0, sizeof *g);
  snprintf(g->name, sizeof g->name, "%s", name);
  return g;
}
static void tie_accumulate(Function *f, const Constant *error) {
  TieGroup *g = tie_group(f->body->name);
  int seen = 0;
  for (int i = 0; i < g->n; i++) seen |= g->leaves[i] == f;
  if (!seen) {
    if (g->n == g->cap) {
      g->cap = g->cap ? g->cap * 2 : 8;
      g->leaves = (Function **)realloc(g->leaves, (size_t)g->cap * sizeof(Function *));
    }
    g->leaves[g->n++] = f;
  }
  Constant part = c_empty_like(&f->value); /* the contribution in the parameter's shape (S<-M: mean, or sum when fixed) */
  c_absorb(&part, error);
  if (!g->has_grad) {
    g->grad = part;
    g->has_grad = 1;
  } else {
    c_op_assign(B_ADD, &g->grad, &part);
    c_free(&part);
  }
}
static void tie_apply(float lr) {
  Constant clr = c_scalar(lr);
  for (int i = 0; i < N_TIES; i++) {
    TieGroup *g = &TIES[i];
    if (g->has_grad) {
      c_op_assign(B_MUL, &g->grad, &clr);
      for (int k = 0; k < g->n; k++) c_op_assign(B_SUB, &g->leaves[k]->value, &g->grad);
      c_free(&g->grad);
    }
    free(g->leaves);
  }
  free(TIES);
  TIES = NULL;
  N_TIES = CAP_TIES = 0;
}

/* Param arm of backprop, optimize.rs:149-152 */
static void sgd_leaf(Function *f, Constant *error, float lr) {
  if (FIX_TIED_PARAMS) {
    tie_accumulate(f, error);
    return;
  }
  Constant clr = c_scalar(lr);
  c_op_assign(B_MUL, error, &clr);      /* *error *= Scalar(lr)   (possibly matrix * scalar) */
  c_op_assign(B_SUB, &f->value, error); /* value -= error         (possibly scalar -= matrix) */
}

/* exact mode: the literal pre-order recursion of optimize.rs:140-204 */
static void backprop_tree(Function *f, Constant *error, float lr) {
  if (!f->has_params) return; /* :147 */
  Expr *e = f->body;
  Constant two = c_scalar(2.0f);
  switch (e->kind) {
    case K_PARAM: sgd_leaf(f, error, lr); return;
    case K_NEG:
      c_assign_unary(U_NEG, error);
      backprop_tree(e->f1, error, lr);
      return;
    case K_SQ:
      c_op_assign(B_MUL, error, &two);
      c_op_assign(B_MUL, error, &e->f1->value);
      backprop_tree(e->f1, error, lr);
      return;
    case K_ABS:
      c_equals_unary(U_SIGNUM, &f->pool[0], &e->f1->value);
      c_op_assign(B_MUL, &f->pool[0], error);
      backprop_tree(e->f1, &f->pool[0], lr);
      return;
    case K_SIGNUM: panic_msg("sign is not differentiable"); return;
    case K_SIGMOID:
      c_equals_unary(U_ONE_MINUS, &f->pool[0], &f->value);
      c_op_assign(B_MUL, &f->pool[0], &f->value);
      c_op_assign(B_MUL, error, &f->pool[0]);
      /* Q3: the child receives the local derivative, not error*derivative (optimize.rs:168-173) */
      backprop_tree(e->f1, FIX_ACT_GRAD ? error : &f->pool[0], lr);
      return;
    case K_TANH:
      c_equals_unary(U_SQ, &f->pool[0], &f->value);
      c_assign_unary(U_ONE_MINUS, &f->pool[0]);
      c_op_assign(B_MUL, error, &f->pool[0]);
      backprop_tree(e->f1, FIX_ACT_GRAD ? error : &f->pool[0], lr);
      return;
    case K_ADD:
      c_absorb(&f->pool[0], error);
      backprop_tree(e->f1, error, lr);
      backprop_tree(e->f2, &f->pool[0], lr);
      return;
    case K_SUB:
      c_equals_unary(U_NEG, &f->pool[0], error);
      backprop_tree(e->f1, error, lr);
      backprop_tree(e->f2, &f->pool[0], lr);
      return;
    case K_MUL:
      c_equals_bin(B_MUL, &f->pool[0], &e->f1->value, error);
      c_op_assign(B_MUL, error, &e->f2->value);
      backprop_tree(e->f1, error, lr);
      backprop_tree(e->f2, &f->pool[0], lr);
      return;
    case K_DOT:
      /* optimize.rs:196-200; only (t1,t2)=(false,false) yields correctly shaped gradients (§8c-3) */
      c_equals_dot(&f->pool[0], error, 0, &e->f2->value, !e->t2);
      c_equals_dot(&f->pool[1], &e->f1->value, !e->t1, error, 0);
      COUNT_GEMM += 2;
      backprop_tree(e->f1, &f->pool[0], lr);
      backprop_tree(e->f2, &f->pool[1], lr);
      return;
    default: return;
  }
}

/* dag mode: deliver a gradient contribution to a child (SURVEY §8 "Execution-mode definitions") */
static void dag_deliver(Function *child, Constant *g, float lr) {
  if (!child->has_params) return;
  Expr *ce = child->body;
  if (ce->kind == K_PARAM) {
    sgd_leaf(child, g, lr);
    return;
  }
  if (ce->kind == K_CONSTANT || ce->kind == K_INPUT) return;
  if (!ce->has_err) {
    ce->err = c_clone(g);
    ce->has_err = 1;
  } else {
    c_op_assign(B_ADD, &ce->err, g); /* f32 sum in consumer-visit order */
  }
}

static void backprop_dag(Function *root, float lr, FnList *order) {
  Constant two = c_scalar(2.0f);
  for (int idx = order->n - 1; idx >= 0; idx--) {
    Function *f = order->v[idx];
    Expr *e = f->body;
    if (!e->has_err) continue;
    Constant *error = &e->err;
    if (f->has_params) {
      switch (e->kind) {
        case K_NEG:
          c_assign_unary(U_NEG, error);
          dag_deliver(e->f1, error, lr);
          break;
        case K_SQ:
          c_op_assign(B_MUL, error, &two);
          c_op_assign(B_MUL, error, &e->f1->value);
          dag_deliver(e->f1, error, lr);
          break;
        case K_ABS:
          c_equals_unary(U_SIGNUM, &f->pool[0], &e->f1->value);
          c_op_assign(B_MUL, &f->pool[0], error);
          dag_deliver(e->f1, &f->pool[0], lr);
          break;
        case K_SIGNUM: panic_msg("sign is not differentiable"); break;
        case K_SIGMOID:
          c_equals_unary(U_ONE_MINUS, &f->pool[0], &f->value);
          c_op_assign(B_MUL, &f->pool[0], &f->value);
          c_op_assign(B_MUL, error, &f->pool[0]);
          dag_deliver(e->f1, FIX_ACT_GRAD ? error : &f->pool[0], lr);
          break;
        case K_TANH:
          c_equals_unary(U_SQ, &f->pool[0], &f->value);
          c_assign_unary(U_ONE_MINUS, &f->pool[0]);
          c_op_assign(B_MUL, error, &f->pool[0]);
          dag_deliver(e->f1, FIX_ACT_GRAD ? error : &f->pool[0], lr);
          break;
        case K_ADD:
          c_absorb(&f->pool[0], error);
          dag_deliver(e->f1, error, lr);
          dag_deliver(e->f2, &f->pool[0], lr);
          break;
        case K_SUB:
          c_equals_unary(U_NEG, &f->pool[0], error);
          dag_deliver(e->f1, error, lr);
          dag_deliver(e->f2, &f->pool[0], lr);
          break;
        case K_MUL:
          c_equals_bin(B_MUL, &f->pool[0], &e->f1->value, error);
          c_op_assign(B_MUL, error, &e->f2->value);
          dag_deliver(e->f1, error, lr);
          dag_deliver(e->f2, &f->pool[0], lr);
          break;
        case K_DOT:
          if (e->f1->has_params) { /* dX is dead when f1 has no params (constant input x) */
            c_equals_dot(&f->pool[0], error, 0, &e->f2->value, !e->t2);
            COUNT_GEMM++;
          }
          if (e->f2->has_params) {
            c_equals_dot(&f->pool[1], &e->f1->value, !e->t1, error, 0);
            COUNT_GEMM++;
          }
          dag_deliver(e->f1, &f->pool[0], lr);
          dag_deliver(e->f2, &f->pool[1], lr);
          break;
        default: break;
      }
    }
    c_free(&e->err);
    e->has_err = 0;
  }
  (void)root;
}

/* minimize, optimize.rs:59-72.  `trace` (may be NULL) receives the root value of every iteration
 * (storage order, size elements each): the forward value *before* that iteration's update (Q12). */
static int run_minimize(Function *f, int nargs, const char **names, const Constant *args, float lr, int iters,
                        float *trace) {
  size_t vsize = f->value.is_matrix ? (size_t)f->value.m.height * f->value.m.width : 1;
  FnList order = {0};
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) {
      forward_tree(f, nargs, names, args);
    } else {
      EPOCH++;
      order.n = 0;
      forward_dag(f, nargs, names, args, &order);
    }
    if (trace) {
      if (f->value.is_matrix) OPS.get_array(&f->value.m, trace + (size_t)i * vsize);
      else trace[i] = f->value.s;
    }
    Constant error = c_copy_and_fill(&f->value, 1.0f); /* :69 — allocated every iteration (Q11) */
    IN_BACKWARD = 1;
    if (MODE == 0) {
      backprop_tree(f, &error, lr);
      c_free(&error);
    } else if (f->has_params) {
      Expr *e = f->body;
      if (e->kind == K_PARAM) {
        sgd_leaf(f, &error, lr);
        c_free(&error);
      } else if (e->kind != K_CONSTANT && e->kind != K_INPUT) {
        e->err = error;
        e->has_err = 1;
        backprop_dag(f, lr, &order);
      } else {
        c_free(&error);
      }
    } else {
      c_free(&error);
    }
    if (FIX_TIED_PARAMS) tie_apply(lr);
    IN_BACKWARD = 0;
  }
  free(order.v);
  return 0;
}

/* args: for name i, arg_vals[i] points at row-major h*w floats (matrix) or one float (scalar, h=w=0) */
int orc_minimize(Function *f, int nargs, const char **names, const float **arg_vals, const unsigned *arg_h,
                 const unsigned *arg_w, float lr, int iters, int print_freq, float *trace) {
  ensure_ops();
  (void)print_freq; /* printing (optimize.rs:66-68, print.rs) is out of scope; `trace` replaces it */
  Constant *args = nargs ? (Constant *)calloc((size_t)nargs, sizeof(Constant)) : NULL;
  for (int i = 0; i < nargs; i++) {
    if (arg_h[i] == 0 && arg_w[i] == 0) args[i] = c_scalar(arg_vals[i][0]);
    else args[i] = c_matrix_from_rows(arg_h[i], arg_w[i], arg_vals[i]);
  }
  int rc = run_minimize(f, nargs, names, args, lr, iters, trace);
  for (int i = 0; i < nargs; i++) c_free(&args[i]);
  free(args);
  return rc;
}
/* maximize, optimize.rs:74-81: (-self).minimize(...) */
int orc_maximize(Function *f, int nargs, const char **names, const float **arg_vals, const unsigned *arg_h,
                 const unsigned *arg_w, float lr, int iters, int print_freq, float *trace) {
  Function *n = orc_neg(f);
  int rc = orc_minimize(n, nargs, names, arg_vals, arg_h, arg_w, lr, iters, print_freq, trace);
  orc_release(n);
  return rc;
}

/* --- read-back -------------------------------------------------------------------------------- */
int orc_value_dims(const Function *f, unsigned *h, unsigned *w) {
  *h = f->value.is_matrix ? f->value.m.height : 0;
  *w = f->value.is_matrix ? f->value.m.width : 0;
  return f->value.is_matrix;
}
void orc_value_get(const Function *f, float *dst) { /* storage (column-major) order, as get_array */
  ensure_ops();
  if (f->value.is_matrix) OPS.get_array(&f->value.m, dst);
  else dst[0] = f->value.s;
}
bool orc_all_less_than(const Function *f, float x) { ensure_ops(); return c_all_less_than(&f->value, x); }
bool orc_all_equal(const Function *f, float x) { ensure_ops(); return c_all_equal(&f->value, x); }
int orc_kind(const Function *f) { return f->body->kind; }

/* Param leaves in memoised post-order (children f1 then f2); each use site is its own leaf (Q1) */
static void collect_leaves(Function *f, FnList *out) {
  Expr *e = f->body;
  if (e->kind == K_PARAM) {
    fl_push(out, f);
    return;
  }
  if (e->kind == K_CONSTANT || e->kind == K_INPUT) return;
  if (e->mark == EPOCH) return;
  e->mark = EPOCH;
  if (e->f1) collect_leaves(e->f1, out);
  if (e->f2) collect_leaves(e->f2, out);
}
static FnList leaves_of(Function *f) {
  FnList l = {0};
  EPOCH++;
  collect_leaves(f, &l);
  return l;
}
int orc_num_leaves(Function *f) {
  FnList l = leaves_of(f);
  int n = l.n;
  free(l.v);
  return n;
}
int orc_leaf_info(Function *f, int i, char *name, int name_cap, unsigned *h, unsigned *w) {
  FnList l = leaves_of(f);
  if (i < 0 || i >= l.n) {
    free(l.v);
    return -1;
  }
  Function *p = l.v[i];
  snprintf(name, (size_t)name_cap, "%s", p->body->name);
  int ism = orc_value_dims(p, h, w);
  free(l.v);
  return ism;
}
int orc_leaf_get(Function *f, int i, float *dst) {
  FnList l = leaves_of(f);
  if (i < 0 || i >= l.n) {
    free(l.v);
    return -1;
  }
  orc_value_get(l.v[i], dst);
  free(l.v);
  return 0;
}
/* number of non-leaf evaluations one exact-mode forward performs (tree size, Q2) */
static long count_tree(const Function *f, long *dots) {
  const Expr *e = f->body;
  if (e->kind == K_CONSTANT || e->kind == K_PARAM || e->kind == K_INPUT) return 0;
  long n = 1;
  if (e->kind == K_DOT) (*dots)++;
  if (e->f1) n += count_tree(e->f1, dots);
  if (e->f2) n += count_tree(e->f2, dots);
  return n;
}
long orc_tree_size(const Function *f, long *dots) {
  *dots = 0;
  return count_tree(f, dots);
}
