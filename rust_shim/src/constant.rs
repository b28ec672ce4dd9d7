//! `Constant` (reference src/constant.rs:5-8): a host-side value used to *construct* leaves and to pass `args`.
//! In the reference a `Constant::Matrix` owns device memory; here device memory belongs to the graph library and a
//! `Constant` is plain host data (row-major, the order `Constant::matrix` takes its values in, constant.rs:62-76).
use matrix::Matrix;
use rand;
use rand::distributions::{IndependentSample, Range};

#[derive(Clone, Debug)]
pub enum Constant {
    Scalar(f32),
    Matrix(Matrix),
}

impl Constant {
    /// 1 for a scalar, else the matrix width (constant.rs:12-17).
    pub fn width(&self) -> usize {
        match *self {
            Constant::Scalar(_) => 1,
            Constant::Matrix(ref m) => m.width(),
        }
    }

    /// 1 for a scalar, else the matrix height (constant.rs:20-25).
    pub fn height(&self) -> usize {
        match *self {
            Constant::Scalar(_) => 1,
            Constant::Matrix(ref m) => m.height(),
        }
    }

    pub fn dims(&self) -> Vec<usize> {
        match *self {
            Constant::Scalar(_) => vec![],
            Constant::Matrix(ref m) => vec![m.height(), m.width()],
        }
    }

    /// constant.rs:36-42
    pub fn single_val(dims: Vec<usize>, val: f32) -> Constant {
        match dims.len() {
            0 => Constant::Scalar(val),
            2 => Constant::Matrix(Matrix::single_val(dims[0], dims[1], val)),
            _ => panic!("not supported"),
        }
    }

    /// constant.rs:45-60 — unseeded uniform draws, as in the reference.
    pub fn random(dims: Vec<usize>, lo: f32, hi: f32) -> Constant {
        let between = Range::new(lo, hi);
        let mut rng = rand::thread_rng();
        match dims.len() {
            0 => Constant::Scalar(between.ind_sample(&mut rng)),
            2 => {
                let vals = (0..dims[0] * dims[1]).map(|_| between.ind_sample(&mut rng)).collect();
                Constant::matrix(dims[0], dims[1], vals)
            }
            _ => panic!("not supported"),
        }
    }

    /// Row-major values (constant.rs:62-76).
    pub fn matrix(height: usize, width: usize, vals: Vec<f32>) -> Constant {
        Constant::Matrix(Matrix::new(height, width, vals))
    }

    pub fn avg(&self) -> f32 {
        match *self {
            Constant::Scalar(x) => x,
            Constant::Matrix(ref m) => m.avg(),
        }
    }
}
