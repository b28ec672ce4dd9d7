//! `Matrix` (reference src/matrix.rs:13-112) as host data. The reference's struct is the FFI handle of one device
//! allocation; with the graph-level ABI the library owns every device buffer (one arena per compiled plan, zero
//! allocations inside the loop), so what user code holds is the row-major host copy it built the matrix from.
//! Programs that link the 26-symbol ABI directly (include/cg_matrix.h = matrix.rs:4-11 + ops.rs:8-33) keep the
//! reference's own matrix.rs instead of this file.
use std::fmt;

#[derive(Clone, Debug)]
pub struct Matrix {
    height: usize,
    width: usize,
    vals: Vec<f32>, // row-major
}

impl Matrix {
    /// matrix.rs:59-66
    pub fn new(height: usize, width: usize, values: Vec<f32>) -> Matrix {
        assert!(values.len() == height * width, "wrong number of values");
        Matrix { height: height, width: width, vals: values }
    }

    /// matrix.rs:68-72
    pub fn single_val(height: usize, width: usize, value: f32) -> Matrix {
        Matrix { height: height, width: width, vals: vec![value; height * width] }
    }

    pub fn height(&self) -> usize { self.height }
    pub fn width(&self) -> usize { self.width }
    pub fn size(&self) -> usize { self.height * self.width }
    pub fn as_slice(&self) -> &[f32] { &self.vals }

    /// ops.rs:438-442 — sequential f32 sum over the storage (column-major in the reference) divided by the size.
    pub fn avg(&self) -> f32 {
        let mut sum = 0f32;
        for j in 0..self.width {
            for i in 0..self.height {
                sum += self.vals[i * self.width + j];
            }
        }
        sum / self.size() as f32
    }

    pub fn print(&self) {
        println!("{}", self);
    }
}

impl fmt::Display for Matrix {
    fn fmt(&self, f: &mut fmt::Formatter) -> fmt::Result {
        for i in 0..self.height {
            let row: Vec<String> = (0..self.width).map(|j| format!("{:^10.3}", self.vals[i * self.width + j])).collect();
            writeln!(f, "{}", row.join(" "))?;
        }
        Ok(())
    }
}
