//! Operators (reference src/ops.rs:36-54, 199-264, 331-407): every constructor is one call into the graph library,
//! which applies the reference's construction-time rules itself — value-based identity elimination (`x + 0 -> x`,
//! `0 - x -> x`, `sigmoid(x)` with x == 0.5 -> x, ...), constant folding and the value-shape rule for mixed
//! scalar/matrix operands.
use ffi::*;
use function::{check, Function};
use libc::c_int;
use std::ops::{Add, Mul, Neg, Sub};

macro_rules! unary {
    ($name:ident, $sym:ident) => {
        impl Function {
            pub fn $name(&self) -> Function {
                Function { h: check(unsafe { $sym(self.h) }) }
            }
        }
    };
}
unary!(sq, cg_sq);
unary!(abs, cg_abs);
unary!(signum, cg_signum);
unary!(sigmoid, cg_sigmoid);
unary!(tanh, cg_tanh);

impl Neg for Function {
    type Output = Function;
    fn neg(self) -> Function { Function { h: check(unsafe { cg_neg(self.h) }) } }
}
impl<'a> Neg for &'a Function {
    type Output = Function;
    fn neg(self) -> Function { Function { h: check(unsafe { cg_neg(self.h) }) } }
}

// all four owned / borrowed combinations, as ops.rs:199-236 provides them
macro_rules! binary {
    ($Trait:ident, $method:ident, $sym:ident) => {
        impl $Trait for Function {
            type Output = Function;
            fn $method(self, other: Function) -> Function { Function { h: check(unsafe { $sym(self.h, other.h) }) } }
        }
        impl<'a> $Trait<&'a Function> for Function {
            type Output = Function;
            fn $method(self, other: &Function) -> Function { Function { h: check(unsafe { $sym(self.h, other.h) }) } }
        }
        impl<'a> $Trait<Function> for &'a Function {
            type Output = Function;
            fn $method(self, other: Function) -> Function { Function { h: check(unsafe { $sym(self.h, other.h) }) } }
        }
        impl<'a, 'b> $Trait<&'b Function> for &'a Function {
            type Output = Function;
            fn $method(self, other: &Function) -> Function { Function { h: check(unsafe { $sym(self.h, other.h) }) } }
        }
    };
}
binary!(Add, add, cg_add);
binary!(Sub, sub, cg_sub);
binary!(Mul, mul, cg_mul);

/// ops.rs:395-407; `dot!(&x, &w)`, `dot!((&x)^T, &w)` ... expand to this (macros.rs).
pub fn dot(f1: &Function, trans1: bool, f2: &Function, trans2: bool) -> Function {
    Function { h: check(unsafe { cg_dot(f1.h, trans1 as c_int, f2.h, trans2 as c_int) }) }
}

/// The README's free-function spelling, `sq(&(x + a))` (README.md:60).
pub fn sq(f: &Function) -> Function { f.sq() }
