// cholesky.rs -- `CholeskyDecomposition<T>` (reference src/decomp/cholesky.rs:52-144) on the B200 path.
// Same public surface: `new(&Matrix<T>) -> Option<Self>` (None unless square, exactly symmetric and positive definite),
// `get_l(&self) -> &Matrix<T>`, `solve(&self, &Matrix<T>) -> Matrix<T>` (panics on a row-count mismatch before any FFI call).
// Source only: there is no Rust toolchain in the build image.
use std::os::raw::c_int;

use crate::ffi;
use crate::matrix::{DeviceScalar, Matrix};

/// Element types with a Cholesky kernel (`la_chol_*_f64` / `_f32`).
pub trait CholScalar: DeviceScalar {
    unsafe fn chol_factor_host(a: *const Self, l_out: *mut Self, n: usize, ok: *mut c_int) -> c_int;
    unsafe fn chol_solve_host(l: *const Self, n: usize, b: *const Self, nx: usize, x: *mut Self) -> c_int;
}
impl CholScalar for f64 {
    unsafe fn chol_factor_host(a: *const f64, l: *mut f64, n: usize, ok: *mut c_int) -> c_int {
        ffi::la_chol_factor_f64_host(a, l, n, ok)
    }
    unsafe fn chol_solve_host(l: *const f64, n: usize, b: *const f64, nx: usize, x: *mut f64) -> c_int {
        ffi::la_chol_solve_f64_host(l, n, b, nx, x)
    }
}
impl CholScalar for f32 {
    unsafe fn chol_factor_host(a: *const f32, l: *mut f32, n: usize, ok: *mut c_int) -> c_int {
        ffi::la_chol_factor_f32_host(a, l, n, ok)
    }
    unsafe fn chol_solve_host(l: *const f32, n: usize, b: *const f32, nx: usize, x: *mut f32) -> c_int {
        ffi::la_chol_solve_f32_host(l, n, b, nx, x)
    }
}

pub struct CholeskyDecomposition<T> {
    l: Matrix<T>,
}

impl<T: CholScalar> CholeskyDecomposition<T> {
    pub fn new(m: &Matrix<T>) -> Option<CholeskyDecomposition<T>> {
        if m.rows() != m.cols() {
            return None; // cholesky.rs:57-59
        }
        let n = m.rows();
        let mut data = Matrix::<T>::dirty_vec(n * n); // alloc_dirty_vec
        let mut ok: c_int = 0;
        ffi::check(unsafe { T::chol_factor_host(m.get_data().as_ptr(), data.as_mut_ptr(), n, &mut ok) });
        if ok == 0 {
            return None; // not symmetric (:91-93) or not positive definite (:99-102)
        }
        Some(CholeskyDecomposition { l: Matrix::new(n, n, data) })
    }

    #[inline]
    pub fn get_l(&self) -> &Matrix<T> {
        &self.l
    }

    pub fn solve(&self, b: &Matrix<T>) -> Matrix<T> {
        assert!(self.l.rows() == b.rows()); // cholesky.rs:118 -- panics before the FFI call
        let (n, nx) = (self.l.rows(), b.cols());
        let mut x = Matrix::<T>::dirty_vec(n * nx);
        ffi::check(unsafe {
            T::chol_solve_host(self.l.get_data().as_ptr(), n, b.get_data().as_ptr(), nx, x.as_mut_ptr())
        });
        Matrix::new(n, nx, x)
    }
}
