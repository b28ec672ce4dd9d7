/* cg_matrix.h — the reference's 26-symbol matrix ABI, served by hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary that exists in the reference today: its Rust host declares exactly
 * these symbols (src/matrix.rs:4-11, src/ops.rs:8-33; C declarations src/cpu/matrix.h:10-21 and
 * src/cpu/ops.h:6-33) and `build.rs` links whichever `libmatrix.a` it built (build.rs:27-41,76-92).
 * libcg_b200.so / libmatrix.a export them unchanged, with `array` a DEVICE pointer (as in the
 * reference's own src/gpu backend): an unmodified Rust host links against it, and the oracle can be
 * diffed symbol by symbol (tests/test_matrix_abi.py).
 *
 * Contract (same as the reference): the caller owns every Matrix; storage is column-major f32
 * (src/cpu/matrix.cpp:26-36); results are caller-allocated; there are no return codes — any failure
 * prints "ERROR: <msg>" to stderr and exits with EXIT_FAILURE (src/cpu/util.cpp:5-10). Each op is one
 * kernel launch on the library's stream; get_array, set_array, reduce_sum, all_equal and
 * all_less_than block until the result is on the host. There is no CPU fallback: without a CUDA
 * device every symbol fails loudly.
 */
#ifndef CG_MATRIX_H
#define CG_MATRIX_H
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/cpu/matrix.h:4-8 == #[repr(C)] struct Matrix (src/matrix.rs:13-18); 16 bytes. */
typedef struct {
  unsigned height;
  unsigned width;
  float *array; /* device pointer, column-major */
} Matrix;

/* ---- memory: src/cpu/matrix.h:10-21 / src/matrix.rs:4-11 ------------------------------------- */
void alloc_matrix(Matrix *m);                 /* fills m->array for the preset height/width; contents uninitialised */
void free_matrix(Matrix *m);
void fill_matrix(Matrix *m, float value);
void copy_matrix(const Matrix *src, Matrix *dst);
void get_array(const Matrix *m, float *dst);  /* host dst, raw storage (column-major) order, h*w floats */
void set_array(Matrix *m, const float *src, bool transpose); /* host src; transpose=true: src is row-major */

/* ---- elementwise maps: src/cpu/ops.h:6-12 / src/ops.rs:9-15 (m may alias result) -------------- */
void map_neg(const Matrix *m, Matrix *result);
void map_sq(const Matrix *m, Matrix *result);
void map_abs(const Matrix *m, Matrix *result);       /* x < 0 ? -x : x   (abs(-0) = -0)   */
void map_signum(const Matrix *m, Matrix *result);    /* x < 0 ? -1 : 1   (signum(0) = +1) */
void map_sigmoid(const Matrix *m, Matrix *result);   /* 1 / (1 + expf(-x)) */
void map_tanh(const Matrix *m, Matrix *result);
void map_one_minus(const Matrix *m, Matrix *result);

/* ---- binary: src/cpu/ops.h:14-27 / src/ops.rs:16-27 ------------------------------------------- */
void elemwise_add(const Matrix *m1, const Matrix *m2, Matrix *result);
void elemwise_sub(const Matrix *m1, const Matrix *m2, Matrix *result);
void elemwise_mul(const Matrix *m1, const Matrix *m2, Matrix *result);
void broadcast_add(const Matrix *m, float val, Matrix *result);
void broadcast_sub(const Matrix *m, float val, Matrix *result);      /* m - val */
void broadcast_mul(const Matrix *m, float val, Matrix *result);
void broadcast_add_rev(float val, const Matrix *m, Matrix *result);
void broadcast_sub_rev(float val, const Matrix *m, Matrix *result);  /* val - m */
void broadcast_mul_rev(float val, const Matrix *m, Matrix *result);

/* ---- product and reductions: src/cpu/ops.h:29-33 / src/ops.rs:28-33 --------------------------- */
/* result = op(m1) . op(m2), column-major. Unlike the reference's CPU gemm (src/cpu/ops.cpp:92-99,
 * valid only for equal square operands) any conforming shapes are accepted. */
void gemm(const Matrix *m1, bool trans1, const Matrix *m2, bool trans2, Matrix *result);
bool all_equal(const Matrix *m, float x);
bool all_less_than(const Matrix *m, float x);
float reduce_sum(const Matrix *m);

#ifdef __cplusplus
}
#endif
#endif /* CG_MATRIX_H */
