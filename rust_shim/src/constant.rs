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
    fn shape(&self) -> (usize, usize) {
        if let Constant::Matrix(ref m) = *self { (m.height(), m.width()) } else { (1, 1) }
    }

    /// Matrix height, 1 for a scalar (same contract as the reference's accessor, constant.rs:20-25).
    pub fn height(&self) -> usize { self.shape().0 }

    /// Matrix width, 1 for a scalar (constant.rs:12-17).
    pub fn width(&self) -> usize { self.shape().1 }

    /// `[]` for a scalar, `[height, width]` for a matrix.
    pub fn dims(&self) -> Vec<usize> {
        if let Constant::Matrix(ref m) = *self { vec![m.height(), m.width()] } else { Vec::new() }
    }

    /// A value whose every element is `val`; `dims` is `[]` or `[height, width]` (constant.rs:36-42).
    pub fn single_val(dims: Vec<usize>, val: f32) -> Constant {
        if dims.is_empty() {
            return Constant::Scalar(val);
        }
        assert!(dims.len() == 2, "not supported");
        Constant::Matrix(Matrix::single_val(dims[0], dims[1], val))
    }

    /// Uniform draws from [lo, hi) with the thread-local generator — unseeded, like the reference (constant.rs:45-60).
    pub fn random(dims: Vec<usize>, lo: f32, hi: f32) -> Constant {
        let dist = Range::new(lo, hi);
        let mut rng = rand::thread_rng();
        let count = match dims.len() {
            0 => 1,
            2 => dims[0] * dims[1],
            _ => panic!("not supported"),
        };
        let mut draws: Vec<f32> = Vec::with_capacity(count);
        while draws.len() < count {
            draws.push(dist.ind_sample(&mut rng));
        }
        if dims.is_empty() { Constant::Scalar(draws[0]) } else { Constant::matrix(dims[0], dims[1], draws) }
    }

    /// Values in ROW-major order, the order the reference's constructor takes them in (constant.rs:62-76).
    pub fn matrix(height: usize, width: usize, vals: Vec<f32>) -> Constant {
        Constant::Matrix(Matrix::new(height, width, vals))
    }

    /// The scalar itself, or the mean of a matrix (ops.rs:430-442).
    pub fn avg(&self) -> f32 {
        if let Constant::Matrix(ref m) = *self { m.avg() } else if let Constant::Scalar(x) = *self { x } else { unreachable!() }
    }
}
