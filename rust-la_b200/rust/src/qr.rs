//! `QRDecomposition<T>` (reference src/decomp/qr.rs:20-238) on the CUDA blocked Householder QR (`la_qr_*`).
//! Same public surface: `new`, `is_full_rank`, `get_h`, `get_r`, `get_q`, `solve`; the packed factors, `rdiag` and the
//! block factors of the compact-WY form stay resident in HBM.  Quirks of the reference are kept: `is_full_rank` indexes
//! `rdiag[0..cols)` (out of bounds for m < n, :112) and `solve` builds `Matrix::new(cols, nx, <m * nx values>)` (:237),
//! which panics unless m == n; its first phase applies I - u u'/u_k (:214-224), not the reflection.
//! Source only: there is no Rust toolchain in the build image.
use std::os::raw::c_int;

use crate::ffi;
use crate::lu::DevBuf;
use crate::matrix::Matrix;

/// f32 / f64 dispatch to the typed C-ABI entry points.
pub trait QrScalar: Copy + PartialEq + Default {
    unsafe fn factor(qr: *mut ffi::la_buf, m: usize, n: usize, rd: *mut ffi::la_buf, tm: *mut ffi::la_buf) -> c_int;
    unsafe fn get_r(qr: *const ffi::la_buf, m: usize, n: usize, rd: *const ffi::la_buf, r: *mut ffi::la_buf) -> c_int;
    unsafe fn get_q(qr: *const ffi::la_buf, m: usize, n: usize, tm: *const ffi::la_buf, q: *mut ffi::la_buf) -> c_int;
    unsafe fn solve(qr: *const ffi::la_buf, m: usize, n: usize, rd: *const ffi::la_buf, b: *const ffi::la_buf, nx: usize,
                    x: *mut ffi::la_buf) -> c_int;
}
macro_rules! qr_scalar {
    ($t:ty, $factor:ident, $get_r:ident, $get_q:ident, $solve:ident) => {
        impl QrScalar for $t {
            unsafe fn factor(qr: *mut ffi::la_buf, m: usize, n: usize, rd: *mut ffi::la_buf, tm: *mut ffi::la_buf) -> c_int {
                ffi::$factor(qr, m, n, rd, tm)
            }
            unsafe fn get_r(qr: *const ffi::la_buf, m: usize, n: usize, rd: *const ffi::la_buf, r: *mut ffi::la_buf) -> c_int {
                ffi::$get_r(qr, m, n, rd, r)
            }
            unsafe fn get_q(qr: *const ffi::la_buf, m: usize, n: usize, tm: *const ffi::la_buf, q: *mut ffi::la_buf) -> c_int {
                ffi::$get_q(qr, m, n, tm, q)
            }
            unsafe fn solve(qr: *const ffi::la_buf, m: usize, n: usize, rd: *const ffi::la_buf, b: *const ffi::la_buf,
                            nx: usize, x: *mut ffi::la_buf) -> c_int {
                ffi::$solve(qr, m, n, rd, b, nx, x)
            }
        }
    };
}
qr_scalar!(f64, la_qr_factor_f64, la_qr_get_r_f64, la_qr_get_q_f64, la_qr_solve_f64);
qr_scalar!(f32, la_qr_factor_f32, la_qr_get_r_f32, la_qr_get_q_f32, la_qr_solve_f32);

pub struct QRDecomposition<T> {
    m: usize,
    n: usize,
    qr_dev: DevBuf,
    rdiag_dev: DevBuf,
    tmat_dev: DevBuf,
    rdiag: Vec<T>,
}

impl<T: QrScalar> QRDecomposition<T> {
    /// qr.rs:26-43
    pub fn new(a: &Matrix<T>) -> QRDecomposition<T> {
        let (m, n) = (a.rows(), a.cols());
        let es = ::std::mem::size_of::<T>();
        let dc = if m < n { m } else { n };
        let qr_dev = DevBuf::new(m * n * es);
        ffi::check(unsafe { ffi::la_buf_upload(qr_dev.0, 0, a.get_data().as_ptr() as *const _, m * n * es) });
        let mut te: usize = 0;
        ffi::check(unsafe { ffi::la_qr_tmat_elems(m, n, 0, es, &mut te) });
        let (rdiag_dev, tmat_dev) = (DevBuf::new((dc + 1) * es), DevBuf::new(te * es));
        ffi::check(unsafe { T::factor(qr_dev.0, m, n, rdiag_dev.0, tmat_dev.0) });
        let mut rdiag = vec![T::default(); dc];
        ffi::check(unsafe { ffi::la_buf_download(rdiag_dev.0, 0, rdiag.as_mut_ptr() as *mut _, dc * es) });
        QRDecomposition { m: m, n: n, qr_dev: qr_dev, rdiag_dev: rdiag_dev, tmat_dev: tmat_dev, rdiag: rdiag }
    }

    /// qr.rs:110-117 (indexes rdiag[j] for every column j: panics out of bounds when m < n, like the reference)
    pub fn is_full_rank(&self) -> bool {
        for j in 0..self.n {
            if self.rdiag[j] == T::default() {
                return false;
            }
        }
        true
    }

    fn download(&self, buf: &DevBuf, rows: usize, cols: usize) -> Matrix<T> {
        let mut d = vec![T::default(); rows * cols];
        ffi::check(unsafe { ffi::la_buf_download(buf.0, 0, d.as_mut_ptr() as *mut _, rows * cols * ::std::mem::size_of::<T>()) });
        Matrix::new(rows, cols, d)
    }

    /// qr.rs:121-135
    pub fn get_h(&self) -> Matrix<T> {
        let qr = self.download(&self.qr_dev, self.m, self.n);
        let mut d = qr.get_data().clone();
        for i in 0..self.m {
            for j in (i + 1)..self.n {
                d[i * self.n + j] = T::default();
            }
        }
        Matrix::new(self.m, self.n, d)
    }

    /// qr.rs:138-152
    pub fn get_r(&self) -> Matrix<T> {
        let r = DevBuf::new(self.m * self.n * ::std::mem::size_of::<T>());
        ffi::check(unsafe { T::get_r(self.qr_dev.0, self.m, self.n, self.rdiag_dev.0, r.0) });
        self.download(&r, self.m, self.n)
    }

    /// qr.rs:155-194
    pub fn get_q(&self) -> Matrix<T> {
        let q = DevBuf::new(self.m * self.m * ::std::mem::size_of::<T>());
        ffi::check(unsafe { T::get_q(self.qr_dev.0, self.m, self.n, self.tmat_dev.0, q.0) });
        self.download(&q, self.m, self.m)
    }

    /// qr.rs:199-238
    pub fn solve(&self, b: &Matrix<T>) -> Option<Matrix<T>> {
        assert!(b.rows() == self.m);
        if !self.is_full_rank() {
            return None;
        }
        let nx = b.cols();
        let bytes = self.m * nx * ::std::mem::size_of::<T>();
        let (bbuf, xbuf) = (DevBuf::new(bytes), DevBuf::new(bytes));
        ffi::check(unsafe { ffi::la_buf_upload(bbuf.0, 0, b.get_data().as_ptr() as *const _, bytes) });
        assert!(self.n * nx == self.m * nx); // Matrix::new(cols, nx, xdata), src/matrix/mod.rs:208
        ffi::check(unsafe { T::solve(self.qr_dev.0, self.m, self.n, self.rdiag_dev.0, bbuf.0, nx, xbuf.0) });
        Some(self.download(&xbuf, self.n, nx))
    }
}
