//! Raw bindings of include/cg_graph.h (the graph-level C ABI of libcg_b200.so). Every function that returns a
//! handle returns NULL on failure and every `c_int` status is non-zero on failure; `cg_last_error()` then holds the
//! reference's own panic text ("dot should not be used with scalars", "inner_dim must equal height2", ...).
use libc::{c_char, c_float, c_int, c_uint};

#[repr(C)]
pub struct CgFunction {
    _private: [u8; 0],
}

extern "C" {
    // leaves (function.rs:100-155)
    pub fn cg_scalar_param(name: *const c_char, value: c_float) -> *mut CgFunction;
    pub fn cg_matrix_param(name: *const c_char, h: c_uint, w: c_uint, rowmajor: *const c_float) -> *mut CgFunction;
    pub fn cg_single_val_param(name: *const c_char, h: c_uint, w: c_uint, value: c_float) -> *mut CgFunction;
    pub fn cg_scalar(value: c_float) -> *mut CgFunction;
    pub fn cg_matrix(h: c_uint, w: c_uint, rowmajor: *const c_float) -> *mut CgFunction;
    pub fn cg_single_val_matrix(h: c_uint, w: c_uint, value: c_float) -> *mut CgFunction;
    pub fn cg_input(name: *const c_char, is_matrix: c_int, h: c_uint, w: c_uint) -> *mut CgFunction;
    pub fn cg_clone(f: *const CgFunction) -> *mut CgFunction;
    pub fn cg_release(f: *mut CgFunction);
    // operators (ops.rs:331-343, 363-407)
    pub fn cg_neg(f: *const CgFunction) -> *mut CgFunction;
    pub fn cg_sq(f: *const CgFunction) -> *mut CgFunction;
    pub fn cg_abs(f: *const CgFunction) -> *mut CgFunction;
    pub fn cg_signum(f: *const CgFunction) -> *mut CgFunction;
    pub fn cg_sigmoid(f: *const CgFunction) -> *mut CgFunction;
    pub fn cg_tanh(f: *const CgFunction) -> *mut CgFunction;
    pub fn cg_add(a: *const CgFunction, b: *const CgFunction) -> *mut CgFunction;
    pub fn cg_sub(a: *const CgFunction, b: *const CgFunction) -> *mut CgFunction;
    pub fn cg_mul(a: *const CgFunction, b: *const CgFunction) -> *mut CgFunction;
    pub fn cg_dot(a: *const CgFunction, t1: c_int, b: *const CgFunction, t2: c_int) -> *mut CgFunction;
    // optimize.rs:59-81
    pub fn cg_minimize(f: *mut CgFunction, nargs: c_int, names: *const *const c_char, vals: *const *const c_float,
                       hs: *const c_uint, ws: *const c_uint, learn_rate: c_float, iters: c_int, print_freq: c_int,
                       trace: *mut c_float) -> c_int;
    pub fn cg_maximize(f: *mut CgFunction, nargs: c_int, names: *const *const c_char, vals: *const *const c_float,
                       hs: *const c_uint, ws: *const c_uint, learn_rate: c_float, iters: c_int, print_freq: c_int,
                       trace: *mut c_float) -> c_int;
    // optimize.rs:8-56
    pub fn cg_eval(f: *mut CgFunction, nargs: c_int, names: *const *const c_char, vals: *const *const c_float,
                   hs: *const c_uint, ws: *const c_uint, dst_colmajor: *mut c_float) -> c_int;
    pub fn cg_grad(f: *const CgFunction, param: *const c_char) -> *mut CgFunction;
    // read-back (ops.rs:309-329, print.rs)
    pub fn cg_value_dims(f: *const CgFunction, h: *mut c_uint, w: *mut c_uint) -> c_int;
    pub fn cg_value_get(f: *const CgFunction, dst_colmajor: *mut c_float);
    pub fn cg_all_less_than(f: *const CgFunction, x: c_float) -> bool;
    pub fn cg_all_equal(f: *const CgFunction, x: c_float) -> bool;
    pub fn cg_last_error() -> *const c_char;
    // execution switches
    pub fn cg_set_mode(mode: c_int);
    pub fn cg_set_tf32(on: c_int);
}
