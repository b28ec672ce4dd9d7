//! `minimize` / `maximize` (reference src/optimize.rs:59-81). The reference re-walks the tree every iteration,
//! one FFI call per node; here the call hands the whole loop to the library, which compiles the walk once into a
//! fused tape and replays it (CUDA graph, or one resident CTA for latency-bound graphs) `iters` times.
use constant::Constant;
use ffi::*;
use function::{check, last_error, Function};
use libc::{c_char, c_float, c_int, c_uint};
use std::collections::HashMap;
use std::ffi::CString;
use std::ptr;

struct FlatArgs {
    _names: Vec<CString>,
    name_ptrs: Vec<*const c_char>,
    scalars: Vec<Box<f32>>,
    val_ptrs: Vec<*const c_float>,
    hs: Vec<c_uint>,
    ws: Vec<c_uint>,
}

/// `args` maps an `Input`'s name to its value (optimize.rs:87-90): row-major floats, (0, 0) dims for a scalar.
fn flatten(args: &HashMap<&str, Constant>) -> FlatArgs {
    let mut a = FlatArgs { _names: vec![], name_ptrs: vec![], scalars: vec![], val_ptrs: vec![], hs: vec![], ws: vec![] };
    for (name, value) in args.iter() {
        let c = CString::new(*name).unwrap();
        a.name_ptrs.push(c.as_ptr());
        a._names.push(c);
        match *value {
            Constant::Scalar(x) => {
                let b = Box::new(x);
                a.val_ptrs.push(&*b as *const f32);
                a.scalars.push(b);
                a.hs.push(0);
                a.ws.push(0);
            }
            Constant::Matrix(ref m) => {
                a.val_ptrs.push(m.as_slice().as_ptr());
                a.hs.push(m.height() as c_uint);
                a.ws.push(m.width() as c_uint);
            }
        }
    }
    a
}

impl Function {
    /// optimize.rs:8-32 — the forward value at the current parameters; nothing cached changes.
    pub fn eval(&self, args: &HashMap<&str, Constant>) -> Constant {
        let a = flatten(args);
        let dims = self.dims();
        let n = if dims.is_empty() { 1 } else { dims[0] * dims[1] };
        let mut colmajor = vec![0f32; n];
        let rc = unsafe {
            cg_eval(self.h, a.name_ptrs.len() as c_int, a.name_ptrs.as_ptr(), a.val_ptrs.as_ptr(), a.hs.as_ptr(),
                    a.ws.as_ptr(), colmajor.as_mut_ptr())
        };
        if rc != 0 {
            panic!("{}", last_error());
        }
        if dims.is_empty() {
            return Constant::Scalar(colmajor[0]);
        }
        let (h, w) = (dims[0], dims[1]);
        let mut rowmajor = vec![0f32; n];
        for j in 0..w {
            for i in 0..h {
                rowmajor[i * w + j] = colmajor[j * h + i];
            }
        }
        Constant::matrix(h, w, rowmajor)
    }

    /// optimize.rs:34-56 — d self / d param as a new Function (the README's "naive" gradient).
    pub fn grad(&self, param: &str) -> Function {
        let name = CString::new(param).unwrap();
        Function { h: check(unsafe { cg_grad(self.h, name.as_ptr()) }) }
    }

    pub fn minimize(&self, args: &HashMap<&str, Constant>, learn_rate: f32, iters: i32, print_freq: i32) {
        let a = flatten(args);
        let rc = unsafe {
            cg_minimize(self.h, a.name_ptrs.len() as c_int, a.name_ptrs.as_ptr(), a.val_ptrs.as_ptr(), a.hs.as_ptr(),
                        a.ws.as_ptr(), learn_rate, iters as c_int, print_freq as c_int, ptr::null_mut())
        };
        if rc != 0 {
            panic!("{}", last_error());
        }
    }

    pub fn maximize(&self, args: &HashMap<&str, Constant>, learn_rate: f32, iters: i32, print_freq: i32) {
        let a = flatten(args);
        let rc = unsafe {
            cg_maximize(self.h, a.name_ptrs.len() as c_int, a.name_ptrs.as_ptr(), a.val_ptrs.as_ptr(), a.hs.as_ptr(),
                        a.ws.as_ptr(), learn_rate, iters as c_int, print_freq as c_int, ptr::null_mut())
        };
        if rc != 0 {
            panic!("{}", last_error());
        }
    }
}
