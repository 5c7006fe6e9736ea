//! `LUDecomposition<T>` (reference src/decomp/lu.rs:95-278) on the CUDA blocked LU.
//! The packed factors stay resident in HBM (`la_buf`); the host `Matrix` is materialised on demand for get_l / get_u.
use std::os::raw::c_int;
use std::ptr;

use crate::ffi;
use crate::matrix::Matrix;

/// f32 / f64 dispatch to the typed C-ABI entry points.
pub trait LuScalar: Copy + PartialEq {
    fn zero() -> Self;
    fn one() -> Self;
    unsafe fn factor(lu: *mut ffi::la_buf, m: usize, n: usize, piv: *mut u64, sign: *mut c_int) -> c_int;
    unsafe fn nonsingular(lu: *const ffi::la_buf, n: usize, out: *mut c_int) -> c_int;
    unsafe fn det(lu: *const ffi::la_buf, n: usize, pos: c_int, out: *mut Self) -> c_int;
    unsafe fn solve(lu: *const ffi::la_buf, m: usize, n: usize, piv: *const u64, b: *const ffi::la_buf, nx: usize,
                    x: *mut ffi::la_buf) -> c_int;
}
macro_rules! lu_scalar {
    ($t:ty, $factor:ident, $ns:ident, $det:ident, $solve:ident) => {
        impl LuScalar for $t {
            fn zero() -> $t { 0.0 }
            fn one() -> $t { 1.0 }
            unsafe fn factor(lu: *mut ffi::la_buf, m: usize, n: usize, piv: *mut u64, sign: *mut c_int) -> c_int {
                ffi::$factor(lu, m, n, piv, sign)
            }
            unsafe fn nonsingular(lu: *const ffi::la_buf, n: usize, out: *mut c_int) -> c_int { ffi::$ns(lu, n, out) }
            unsafe fn det(lu: *const ffi::la_buf, n: usize, pos: c_int, out: *mut $t) -> c_int {
                ffi::$det(lu, n, pos, out)
            }
            unsafe fn solve(lu: *const ffi::la_buf, m: usize, n: usize, piv: *const u64, b: *const ffi::la_buf,
                            nx: usize, x: *mut ffi::la_buf) -> c_int {
                ffi::$solve(lu, m, n, piv, b, nx, x)
            }
        }
    };
}
lu_scalar!(f64, la_lu_factor_f64, la_lu_is_nonsingular_f64, la_lu_det_f64, la_lu_solve_f64);
lu_scalar!(f32, la_lu_factor_f32, la_lu_is_nonsingular_f32, la_lu_det_f32, la_lu_solve_f32);

pub(crate) struct DevBuf(pub(crate) *mut ffi::la_buf);
// The C ABI is thread-safe (include/la_cabi.h: per-thread streams and scratch, no unguarded global state) and a la_buf is
// only a handle to device memory, so the decompositions stay Send + Sync like the reference's plain-Vec structs.
unsafe impl Send for DevBuf {}
unsafe impl Sync for DevBuf {}
impl DevBuf {
    pub(crate) fn new(bytes: usize) -> DevBuf {
        let mut h: *mut ffi::la_buf = ptr::null_mut();
        ffi::check(unsafe { ffi::la_buf_alloc(bytes, 0, &mut h) });
        DevBuf(h)
    }
}
impl Drop for DevBuf {
    fn drop(&mut self) { unsafe { ffi::la_buf_free(self.0); } }
}

pub struct LUDecomposition<T> {
    m: usize,
    n: usize,
    lu_dev: DevBuf,
    pospivsign: bool,
    piv: Vec<usize>,
    _marker: ::std::marker::PhantomData<T>,
}

impl LUDecomposition<f64> {
    /// The same factorisation spread over several GPUs of the node (`la_lu_factor_f64_mg`: 128-column blocks dealt
    /// round-robin, the panel owner's block column copied to every device).  The packed factors come back to device 0,
    /// so `solve`, `det`, `get_l` ... work as after `new`.
    pub fn new_on_devices(a: &Matrix<f64>, devices: &[c_int]) -> LUDecomposition<f64> {
        let (m, n) = (a.rows(), a.cols());
        assert!(m == n);
        assert!(!devices.is_empty());
        let mut lu = vec![0f64; n * n];
        let mut piv64 = vec![0u64; n];
        let mut sign: c_int = 1;
        ffi::check(unsafe {
            ffi::la_lu_factor_f64_mg(devices.len() as c_int, devices.as_ptr(), a.get_data().as_ptr(), lu.as_mut_ptr(), n,
                                     piv64.as_mut_ptr(), &mut sign)
        });
        let bytes = n * n * ::std::mem::size_of::<f64>();
        let buf = DevBuf::new(bytes);
        ffi::check(unsafe { ffi::la_buf_upload(buf.0, 0, lu.as_ptr() as *const _, bytes) });
        LUDecomposition {
            m: n, n: n, lu_dev: buf, pospivsign: sign != 0,
            piv: piv64.into_iter().map(|p| p as usize).collect(),
            _marker: ::std::marker::PhantomData,
        }
    }
}

impl<T: LuScalar> LUDecomposition<T> {
    /// src/decomp/lu.rs:104-168.  Factorises a copy of `a` (:105) on the device.
    pub fn new(a: &Matrix<T>) -> LUDecomposition<T> {
        let (m, n) = (a.rows(), a.cols());
        let bytes = m * n * ::std::mem::size_of::<T>();
        let buf = DevBuf::new(bytes);
        ffi::check(unsafe { ffi::la_buf_upload(buf.0, 0, a.get_data().as_ptr() as *const _, bytes) });
        let mut piv64 = vec![0u64; m];
        let mut sign: c_int = 1;
        ffi::check(unsafe { T::factor(buf.0, m, n, piv64.as_mut_ptr(), &mut sign) });
        LUDecomposition {
            m: m, n: n, lu_dev: buf, pospivsign: sign != 0,
            piv: piv64.into_iter().map(|p| p as usize).collect(),
            _marker: ::std::marker::PhantomData,
        }
    }

    fn lu_host(&self) -> Vec<T> {
        let len = self.m * self.n;
        let mut v: Vec<T> = Vec::with_capacity(len);
        unsafe { v.set_len(len) };
        ffi::check(unsafe {
            ffi::la_buf_download(self.lu_dev.0, 0, v.as_mut_ptr() as *mut _, len * ::std::mem::size_of::<T>())
        });
        v
    }

    pub fn is_singular(&self) -> bool { !self.is_non_singular() }

    /// src/decomp/lu.rs:174-182
    pub fn is_non_singular(&self) -> bool {
        assert!(self.m >= self.n); // the reference indexes lu[j*n+j] for j < n and panics out of bounds otherwise
        let mut out: c_int = 0;
        ffi::check(unsafe { T::nonsingular(self.lu_dev.0, self.n, &mut out) });
        out != 0
    }

    /// src/decomp/lu.rs:184-202
    pub fn get_l(&self) -> Matrix<T> {
        let lu = self.lu_host();
        let (m, n) = (self.m, if self.m >= self.n { self.n } else { self.m });
        let mut l = Vec::with_capacity(m * n);
        for i in 0..m {
            for j in 0..n {
                l.push(if i > j { lu[i * self.n + j] } else if i == j { T::one() } else { T::zero() });
            }
        }
        Matrix::new(m, n, l)
    }

    /// src/decomp/lu.rs:204-215
    pub fn get_u(&self) -> Matrix<T> {
        let lu = self.lu_host();
        let (m, n) = (if self.m >= self.n { self.n } else { self.m }, self.n);
        let mut u = Vec::with_capacity(m * n);
        for i in 0..m {
            for j in 0..n {
                u.push(if i <= j { lu[i * n + j] } else { T::zero() });
            }
        }
        Matrix::new(m, n, u)
    }

    /// src/decomp/lu.rs:217-220: id(len, len).permute_rows(piv)
    pub fn get_p(&self) -> Matrix<T> {
        let len = self.piv.len();
        let mut p = vec![T::zero(); len * len];
        for i in 0..len { p[i * len + self.piv[i]] = T::one(); }
        Matrix::new(len, len, p)
    }

    pub fn get_piv<'lt>(&'lt self) -> &'lt Vec<usize> { &self.piv }

    /// src/decomp/lu.rs:224-232
    pub fn det(&self) -> T {
        assert!(self.m == self.n);
        let mut out = T::zero();
        ffi::check(unsafe { T::det(self.lu_dev.0, self.n, self.pospivsign as c_int, &mut out) });
        out
    }

    /// src/decomp/lu.rs:237-278: Some(X) with L*U*X = B(piv,:), None when singular.
    pub fn solve(&self, b: &Matrix<T>) -> Option<Matrix<T>> {
        assert!(b.rows() == self.m);
        if !self.is_non_singular() {
            return None;
        }
        let nx = b.cols();
        let bytes = self.m * nx * ::std::mem::size_of::<T>();
        let (bbuf, xbuf) = (DevBuf::new(bytes), DevBuf::new(bytes));
        ffi::check(unsafe { ffi::la_buf_upload(bbuf.0, 0, b.get_data().as_ptr() as *const _, bytes) });
        let piv64: Vec<u64> = self.piv.iter().map(|&p| p as u64).collect();
        ffi::check(unsafe { T::solve(self.lu_dev.0, self.m, self.n, piv64.as_ptr(), bbuf.0, nx, xbuf.0) });
        let mut x: Vec<T> = Vec::with_capacity(self.m * nx);
        unsafe { x.set_len(self.m * nx) };
        ffi::check(unsafe { ffi::la_buf_download(xbuf.0, 0, x.as_mut_ptr() as *mut _, bytes) });
        Some(Matrix::new(self.m, nx, x))
    }
}

/// The LU callers on Matrix (src/matrix/mod.rs:1025-1047); each call re-factorises, like the reference.
impl<T: LuScalar> Matrix<T> {
    pub fn det(&self) -> T {
        assert!(self.cols() == self.rows());
        LUDecomposition::new(self).det()
    }
    pub fn solve(&self, b: &Matrix<T>) -> Option<Matrix<T>> { LUDecomposition::new(self).solve(b) }
    pub fn inverse(&self) -> Option<Matrix<T>> {
        assert!(self.rows() == self.cols());
        let n = self.rows();
        let mut id = vec![T::zero(); n * n];
        for i in 0..n { id[i * n + i] = T::one(); }
        LUDecomposition::new(self).solve(&Matrix::new(n, n, id))
    }
    pub fn is_singular(&self) -> bool { !self.is_non_singular() }
    pub fn is_non_singular(&self) -> bool {
        assert!(self.rows() == self.cols());
        LUDecomposition::new(self).is_non_singular()
    }
}
