//! `la` -- the rust-la public surface for the dense hot path, with the arithmetic on B200 (sm_100a) CUDA kernels.
//!
//! Kept from the reference (src/lib.rs:6-28): `Matrix`, `LUDecomposition`, `ApproxEq`, the `m!` macro, operator `*`.
//! Widened since: `CholeskyDecomposition` and `QRDecomposition` run on the device as well.  The
//! reference's remaining modules (SVD, Eigen, iterators, CSV) are host code that is unchanged by this work and is not
//! duplicated here; they keep compiling against this `Matrix` because its field layout and accessors are the reference's.
extern crate num;

#[macro_use]
mod macros;
mod approxeq;
mod cholesky;
mod ffi;
mod lu;
mod matrix;
mod qr;

pub use approxeq::ApproxEq;
pub use cholesky::{CholScalar, CholeskyDecomposition};
pub use lu::LUDecomposition;
pub use matrix::{devices, set_devices, DeviceScalar, Matrix};
pub use qr::{QRDecomposition, QrScalar};
