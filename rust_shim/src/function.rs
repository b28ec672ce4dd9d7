//! `Function` (reference src/function.rs:10-16) as an owning handle on a node of the graph library.
//!
//! The reference's `Function` is {value, params, body: Rc<Expr>, pool}; cloning deep-copies the value and shares
//! the body, so every *use site* of a parameter trains its own copy (function.rs:10-16, pinned by
//! main.rs:144-154). `cg_clone` has exactly those semantics, so `#[derive(Clone)]` becomes a call to it.
use constant::Constant;
use ffi::*;
use libc::c_uint;
use std::ffi::{CStr, CString};
use std::fmt;

pub struct Function {
    pub(crate) h: *mut CgFunction,
}

/// NULL handle / non-zero status -> the reference's `panic!` with the reference's message.
pub(crate) fn check(h: *mut CgFunction) -> *mut CgFunction {
    if h.is_null() {
        panic!("{}", last_error());
    }
    h
}

pub(crate) fn last_error() -> String {
    unsafe { CStr::from_ptr(cg_last_error()).to_string_lossy().into_owned() }
}

impl Clone for Function {
    fn clone(&self) -> Function {
        Function { h: check(unsafe { cg_clone(self.h) }) }
    }
}

impl Drop for Function {
    fn drop(&mut self) {
        unsafe { cg_release(self.h) }
    }
}

impl Function {
    // ---- constructors (function.rs:100-155) ---------------------------------------------------------------
    pub fn constant(value: Constant) -> Function {
        match value {
            Constant::Scalar(x) => Function::scalar(x),
            Constant::Matrix(ref m) => Function {
                h: check(unsafe { cg_matrix(m.height() as c_uint, m.width() as c_uint, m.as_slice().as_ptr()) }),
            },
        }
    }

    pub fn input(s: &str, dims: Vec<usize>) -> Function {
        let name = CString::new(s).unwrap();
        let h = match dims.len() {
            0 => unsafe { cg_input(name.as_ptr(), 0, 0, 0) },
            2 => unsafe { cg_input(name.as_ptr(), 1, dims[0] as c_uint, dims[1] as c_uint) },
            _ => panic!("not supported"),
        };
        Function { h: check(h) }
    }

    pub fn param(s: &str, value: Constant) -> Function {
        let name = CString::new(s).unwrap();
        let h = match value {
            Constant::Scalar(x) => unsafe { cg_scalar_param(name.as_ptr(), x) },
            Constant::Matrix(ref m) => unsafe {
                cg_matrix_param(name.as_ptr(), m.height() as c_uint, m.width() as c_uint, m.as_slice().as_ptr())
            },
        };
        Function { h: check(h) }
    }

    pub fn scalar_param(s: &str, value: f32) -> Function {
        Function::param(s, Constant::Scalar(value))
    }

    pub fn matrix_param(s: &str, height: usize, width: usize, vals: Vec<f32>) -> Function {
        Function::param(s, Constant::matrix(height, width, vals))
    }

    pub fn single_val_param(s: &str, dims: Vec<usize>, val: f32) -> Function {
        match dims.len() {
            2 => {
                let name = CString::new(s).unwrap();
                Function { h: check(unsafe { cg_single_val_param(name.as_ptr(), dims[0] as c_uint, dims[1] as c_uint, val) }) }
            }
            _ => Function::param(s, Constant::single_val(dims, val)),
        }
    }

    pub fn random_param(s: &str, dims: Vec<usize>, lo: f32, hi: f32) -> Function {
        Function::param(s, Constant::random(dims, lo, hi))
    }

    pub fn random_matrix(dims: Vec<usize>, lo: f32, hi: f32) -> Function {
        Function::constant(Constant::random(dims, lo, hi))
    }

    pub fn scalar(x: f32) -> Function {
        Function { h: check(unsafe { cg_scalar(x) }) }
    }

    pub fn matrix(height: usize, width: usize, values: Vec<f32>) -> Function {
        Function::constant(Constant::matrix(height, width, values))
    }

    pub fn single_val_matrix(height: usize, width: usize, value: f32) -> Function {
        Function { h: check(unsafe { cg_single_val_matrix(height as c_uint, width as c_uint, value) }) }
    }

    /// function.rs:176-183 — in the reference this asserts that every argument is a `Param`; leaf kinds are not
    /// observable through the handle, and lstm.rs only ever passes `random_param`s, so this is a no-op.
    pub fn check_params(_functions: Vec<&Function>) {}

    /// Shape of the value ([] for a scalar), known from construction on.
    pub fn dims(&self) -> Vec<usize> {
        let (mut h, mut w) = (0 as c_uint, 0 as c_uint);
        if unsafe { cg_value_dims(self.h, &mut h, &mut w) } != 0 { vec![h as usize, w as usize] } else { vec![] }
    }

    // ---- read-back ----------------------------------------------------------------------------------------
    /// The cached forward value (row-major for a matrix). After `minimize(n)` it is the forward value of the
    /// last iteration, i.e. at the parameters after n-1 updates (optimize.rs:64-71).
    pub fn value(&self) -> Constant {
        let (mut h, mut w) = (0 as c_uint, 0 as c_uint);
        let is_matrix = unsafe { cg_value_dims(self.h, &mut h, &mut w) } != 0;
        if !is_matrix {
            let mut x = 0f32;
            unsafe { cg_value_get(self.h, &mut x) };
            return Constant::Scalar(x);
        }
        let (h, w) = (h as usize, w as usize);
        let mut colmajor = vec![0f32; h * w];
        unsafe { cg_value_get(self.h, colmajor.as_mut_ptr()) };
        let mut rowmajor = vec![0f32; h * w];
        for j in 0..w {
            for i in 0..h {
                rowmajor[i * w + j] = colmajor[j * h + i];
            }
        }
        Constant::matrix(h, w, rowmajor)
    }

    /// ops.rs:309-329
    pub fn all_less_than(&self, val: f32) -> bool {
        unsafe { cg_all_less_than(self.h, val) }
    }

    pub fn all_equal(&self, val: f32) -> bool {
        unsafe { cg_all_equal(self.h, val) }
    }
}

/// print.rs renders the whole expression tree; the handle exposes the value, which is what the tests'
/// `assert!(.., "F: {}", f)` needs on failure.
impl fmt::Display for Function {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        match self.value() {
            Constant::Scalar(x) => write!(f, "{}", x),
            Constant::Matrix(ref m) => write!(f, "\n{}", m),
        }
    }
}
