//! `Matrix<T>`: `{ no_rows, data: Vec<T> }`, row-major (reference src/matrix/mod.rs:26-30), plus the `Mul` impls
//! (mod.rs:957-998) and `mmul` (mmatrix.rs:82-98) routed to the CUDA GEMM through `DeviceScalar`.
//!
//! H1: the reference's `Mul` is generic over `T: Add + Mul + Zero + Copy`.  Stable Rust has no
//! specialisation, so the GPU dispatch is a sealed trait `DeviceScalar` implemented for f64, f32 and i64 -- the three
//! element types the CUDA library has a kernel for.  This NARROWS the bound of `Mul`; there is no CPU loop left.
use std::ops::Mul;
use std::os::raw::c_int;
use std::sync::Mutex;

use crate::ffi;

/// Devices the products are spread over.  Empty or one entry: the single-device entry points.
static DEVICES: Mutex<Vec<c_int>> = Mutex::new(Vec::new());

/// Configure the device list used by `Mul` / `mmul` (and available to `LUDecomposition::new_on_devices`): with more than
/// one device the rows of A and C are sharded and the column blocks of B travel over NVLink inside the library
/// (`la_gemm_f64_mg` / `la_gemm_f32_mg`).
pub fn set_devices(devices: &[c_int]) {
    let mut d = DEVICES.lock().unwrap();
    d.clear();
    d.extend_from_slice(devices);
}
pub fn devices() -> Vec<c_int> { DEVICES.lock().unwrap().clone() }

mod sealed {
    pub trait Sealed {}
    impl Sealed for f64 {}
    impl Sealed for f32 {}
    impl Sealed for i64 {}
}

/// Element types with a CUDA GEMM kernel behind the C ABI.
pub trait DeviceScalar: Copy + PartialEq + sealed::Sealed {
    unsafe fn gemm_host(a: *const Self, b: *const Self, c: *mut Self, m: usize, k: usize, n: usize) -> c_int;
    /// Product over several devices; the default (element types without a multi-device kernel) runs on one.
    unsafe fn gemm_mg(_devices: &[c_int], a: *const Self, b: *const Self, c: *mut Self, m: usize, k: usize, n: usize)
                      -> c_int {
        Self::gemm_host(a, b, c, m, k, n)
    }
    /// Dispatch on the configured device list.
    unsafe fn gemm(a: *const Self, b: *const Self, c: *mut Self, m: usize, k: usize, n: usize) -> c_int {
        let devs = devices();
        if devs.len() > 1 { Self::gemm_mg(&devs, a, b, c, m, k, n) } else { Self::gemm_host(a, b, c, m, k, n) }
    }
}
impl DeviceScalar for f64 {
    unsafe fn gemm_host(a: *const f64, b: *const f64, c: *mut f64, m: usize, k: usize, n: usize) -> c_int {
        ffi::la_gemm_f64_host(a, b, c, m, k, n)
    }
    unsafe fn gemm_mg(devices: &[c_int], a: *const f64, b: *const f64, c: *mut f64, m: usize, k: usize, n: usize) -> c_int {
        ffi::la_gemm_f64_mg(devices.len() as c_int, devices.as_ptr(), a, b, c, m, k, n)
    }
}
impl DeviceScalar for f32 {
    unsafe fn gemm_host(a: *const f32, b: *const f32, c: *mut f32, m: usize, k: usize, n: usize) -> c_int {
        ffi::la_gemm_f32_host(a, b, c, m, k, n)
    }
    unsafe fn gemm_mg(devices: &[c_int], a: *const f32, b: *const f32, c: *mut f32, m: usize, k: usize, n: usize) -> c_int {
        ffi::la_gemm_f32_mg(devices.len() as c_int, devices.as_ptr(), a, b, c, m, k, n)
    }
}
impl DeviceScalar for i64 {
    unsafe fn gemm_host(a: *const i64, b: *const i64, c: *mut i64, m: usize, k: usize, n: usize) -> c_int {
        ffi::la_gemm_i64_host(a, b, c, m, k, n)
    }
}

#[derive(PartialEq, Clone, Debug)]
pub struct Matrix<T> {
    no_rows: usize,
    data: Vec<T>,
}

impl<T: Copy> Matrix<T> {
    /// src/matrix/mod.rs:207-211
    pub fn new(no_rows: usize, no_cols: usize, data: Vec<T>) -> Matrix<T> {
        assert!(no_rows * no_cols == data.len());
        assert!(no_rows > 0 && no_cols > 0);
        Matrix { no_rows: no_rows, data: data }
    }
    #[inline]
    pub fn rows(&self) -> usize { self.no_rows }
    #[inline]
    pub fn cols(&self) -> usize { self.data.len() / self.no_rows }
    #[inline]
    pub fn get_data<'a>(&'a self) -> &'a Vec<T> { &self.data }
    #[inline]
    pub fn get_mut_data<'a>(&'a mut self) -> &'a mut Vec<T> { &mut self.data }
    #[inline]
    pub fn get(&self, row: usize, col: usize) -> T {
        assert!(row < self.no_rows && col < self.cols());
        self.data[row * self.cols() + col]
    }
}

impl<T: DeviceScalar> Matrix<T> {
    /// Output allocation of Mul: the reference's `alloc_dirty_vec` (src/internalutil.rs:7-13).
    pub(crate) fn dirty_vec(len: usize) -> Vec<T> {
        let mut v = Vec::with_capacity(len);
        unsafe { v.set_len(len) };
        v
    }

    /// src/matrix/mmatrix.rs:82-98
    pub fn mmul<'a>(&self, m: &Matrix<T>, dst: &'a mut Matrix<T>) -> &'a mut Matrix<T> {
        assert!(self.cols() == m.no_rows);
        assert!(dst.rows() == self.no_rows);
        assert!(dst.cols() == m.cols());
        ffi::check(unsafe {
            T::gemm(self.data.as_ptr(), m.data.as_ptr(), dst.data.as_mut_ptr(), self.no_rows, self.cols(), m.cols())
        });
        dst
    }
}

/// src/matrix/mod.rs:957-980: `C = A * B`; panics on a shape mismatch BEFORE any FFI call (:961).
impl<'a, 'b, T: DeviceScalar> Mul<&'a Matrix<T>> for &'b Matrix<T> {
    type Output = Matrix<T>;
    fn mul(self, m: &'a Matrix<T>) -> Matrix<T> {
        assert!(self.cols() == m.no_rows);
        let elems = self.no_rows * m.cols();
        let mut d = Matrix::<T>::dirty_vec(elems);
        ffi::check(unsafe {
            T::gemm(self.data.as_ptr(), m.data.as_ptr(), d.as_mut_ptr(), self.no_rows, self.cols(), m.cols())
        });
        Matrix { no_rows: self.no_rows, data: d }
    }
}
/// Forwarders, src/matrix/mod.rs:982-998.
impl<'a, T: DeviceScalar> Mul<Matrix<T>> for &'a Matrix<T> {
    type Output = Matrix<T>;
    fn mul(self, m: Matrix<T>) -> Matrix<T> { self * &m }
}
impl<T: DeviceScalar> Mul<Matrix<T>> for Matrix<T> {
    type Output = Matrix<T>;
    fn mul(self, m: Matrix<T>) -> Matrix<T> { (&self) * &m }
}
impl<'a, T: DeviceScalar> Mul<&'a Matrix<T>> for Matrix<T> {
    type Output = Matrix<T>;
    fn mul(self, m: &'a Matrix<T>) -> Matrix<T> { (&self) * m }
}
