//! Raw bindings of include/la_cabi.h.  One `extern "C"` item per header declaration, same order.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct la_buf {
    _private: [u8; 0],
}
#[repr(C)]
pub struct la_mg {
    _private: [u8; 0],
}
#[repr(C)]
pub struct la_lu_mg {
    _private: [u8; 0],
}
pub const LA_MG_HANDLE_BYTES: usize = 256;
pub const LA_EW_ADD: c_int = 0;
pub const LA_EW_SUB: c_int = 1;
pub const LA_EW_MUL: c_int = 2;
pub const LA_EW_DIV: c_int = 3;
pub const LA_EW_SCALE: c_int = 4;
pub const LA_EW_NEG: c_int = 5;
pub const LA_RED_SUMSQ: c_int = 0;
pub const LA_RED_ABS_SUM: c_int = 1;
pub const LA_RED_ABS_MAX: c_int = 2;
pub const LA_RED_DOT: c_int = 3;

pub const LA_OK: c_int = 0;
pub const LA_ERR_INVALID: c_int = 1;
pub const LA_ERR_CUDA: c_int = 2;
pub const LA_ERR_NOMEM: c_int = 3;
pub const LA_ERR_NO_DEVICE: c_int = 4;
pub const LA_ERR_UNSUPPORTED: c_int = 5;

pub const LA_GEMM_ASSIGN: c_int = 0;
pub const LA_GEMM_SUB: c_int = 1;
pub const LA_GEMM_ADD: c_int = 2;

extern "C" {
    pub fn la_version() -> c_int;
    pub fn la_last_error() -> *const c_char;
    pub fn la_device_count(out: *mut c_int) -> c_int;
    pub fn la_device_sm_count(device: c_int, out: *mut c_int) -> c_int;
    pub fn la_sync(device: c_int) -> c_int;

    pub fn la_buf_alloc(bytes: usize, device: c_int, out: *mut *mut la_buf) -> c_int;
    pub fn la_buf_free(buf: *mut la_buf) -> c_int;
    pub fn la_buf_upload(dst: *mut la_buf, dst_offset_bytes: usize, host: *const c_void, bytes: usize) -> c_int;
    pub fn la_buf_download(src: *const la_buf, src_offset_bytes: usize, host: *mut c_void, bytes: usize) -> c_int;
    pub fn la_buf_copy(dst: *mut la_buf, src: *const la_buf, bytes: usize) -> c_int;
    pub fn la_buf_device_ptr(buf: *const la_buf) -> *mut c_void;
    pub fn la_buf_bytes(buf: *const la_buf) -> usize;
    pub fn la_buf_device(buf: *const la_buf) -> c_int;
    pub fn la_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn la_host_free(ptr: *mut c_void) -> c_int;

    pub fn la_gemm_f64(a: *const la_buf, b: *const la_buf, c: *mut la_buf, m: usize, k: usize, n: usize) -> c_int;
    pub fn la_gemm_f32(a: *const la_buf, b: *const la_buf, c: *mut la_buf, m: usize, k: usize, n: usize) -> c_int;
    pub fn la_gemm_f64_host(a: *const f64, b: *const f64, c: *mut f64, m: usize, k: usize, n: usize) -> c_int;
    pub fn la_gemm_f32_host(a: *const f32, b: *const f32, c: *mut f32, m: usize, k: usize, n: usize) -> c_int;
    pub fn la_gemm_i64_host(a: *const i64, b: *const i64, c: *mut i64, m: usize, k: usize, n: usize) -> c_int;
    pub fn la_gemm_f64_dev(a: *const f64, lda: usize, b: *const f64, ldb: usize, c: *mut f64, ldc: usize,
                           m: usize, k: usize, n: usize, mode: c_int, cuda_stream: *mut c_void) -> c_int;
    pub fn la_gemm_f32_dev(a: *const f32, lda: usize, b: *const f32, ldb: usize, c: *mut f32, ldc: usize,
                           m: usize, k: usize, n: usize, mode: c_int, cuda_stream: *mut c_void) -> c_int;

    pub fn la_lu_factor_f64(lu: *mut la_buf, m: usize, n: usize, piv_out: *mut u64, pospivsign_out: *mut c_int) -> c_int;
    pub fn la_lu_factor_f32(lu: *mut la_buf, m: usize, n: usize, piv_out: *mut u64, pospivsign_out: *mut c_int) -> c_int;
    pub fn la_lu_factor_f64_host(a: *const f64, lu_out: *mut f64, m: usize, n: usize, piv_out: *mut u64,
                                 pospivsign_out: *mut c_int) -> c_int;
    pub fn la_lu_factor_f32_host(a: *const f32, lu_out: *mut f32, m: usize, n: usize, piv_out: *mut u64,
                                 pospivsign_out: *mut c_int) -> c_int;
    pub fn la_lu_factor_f64_dev(lu: *mut f64, m: usize, n: usize, piv_dev: *mut u64, sign_dev: *mut c_int,
                                cuda_stream: *mut c_void) -> c_int;
    pub fn la_lu_factor_f32_dev(lu: *mut f32, m: usize, n: usize, piv_dev: *mut u64, sign_dev: *mut c_int,
                                cuda_stream: *mut c_void) -> c_int;

    pub fn la_lu_is_nonsingular_f64(lu: *const la_buf, n: usize, out: *mut c_int) -> c_int;
    pub fn la_lu_is_nonsingular_f32(lu: *const la_buf, n: usize, out: *mut c_int) -> c_int;
    pub fn la_lu_det_f64(lu: *const la_buf, n: usize, pospivsign: c_int, out: *mut f64) -> c_int;
    pub fn la_lu_det_f32(lu: *const la_buf, n: usize, pospivsign: c_int, out: *mut f32) -> c_int;
    pub fn la_lu_solve_f64(lu: *const la_buf, m: usize, n: usize, piv: *const u64, b: *const la_buf, nx: usize,
                           x: *mut la_buf) -> c_int;
    pub fn la_lu_solve_f32(lu: *const la_buf, m: usize, n: usize, piv: *const u64, b: *const la_buf, nx: usize,
                           x: *mut la_buf) -> c_int;
    pub fn la_lu_solve_f64_host(lu: *const f64, m: usize, n: usize, piv: *const u64, b: *const f64, nx: usize,
                                x: *mut f64) -> c_int;
    pub fn la_lu_solve_f32_host(lu: *const f32, m: usize, n: usize, piv: *const u64, b: *const f32, nx: usize,
                                x: *mut f32) -> c_int;
    pub fn la_lu_solve_f64_dev(lu: *const f64, n: usize, piv_dev: *const u64, b: *const f64, nx: usize, x: *mut f64,
                               cuda_stream: *mut c_void) -> c_int;
    pub fn la_lu_solve_f32_dev(lu: *const f32, n: usize, piv_dev: *const u64, b: *const f32, nx: usize, x: *mut f32,
                               cuda_stream: *mut c_void) -> c_int;

    pub fn la_chol_factor_f64(a_inout: *mut la_buf, n: usize, ok_out: *mut c_int) -> c_int;
    pub fn la_chol_factor_f32(a_inout: *mut la_buf, n: usize, ok_out: *mut c_int) -> c_int;
    pub fn la_chol_factor_f64_host(a: *const f64, l_out: *mut f64, n: usize, ok_out: *mut c_int) -> c_int;
    pub fn la_chol_factor_f32_host(a: *const f32, l_out: *mut f32, n: usize, ok_out: *mut c_int) -> c_int;
    pub fn la_chol_factor_f64_dev(a_inout: *mut f64, n: usize, flags_dev: *mut c_int, stream: *mut c_void) -> c_int;
    pub fn la_chol_factor_f32_dev(a_inout: *mut f32, n: usize, flags_dev: *mut c_int, stream: *mut c_void) -> c_int;
    pub fn la_chol_solve_f64_dev(l: *const f64, n: usize, b: *const f64, nx: usize, x: *mut f64, stream: *mut c_void) -> c_int;
    pub fn la_chol_solve_f32_dev(l: *const f32, n: usize, b: *const f32, nx: usize, x: *mut f32, stream: *mut c_void) -> c_int;
    pub fn la_chol_solve_f64(l: *const la_buf, n: usize, b: *const la_buf, nx: usize, x: *mut la_buf) -> c_int;
    pub fn la_chol_solve_f32(l: *const la_buf, n: usize, b: *const la_buf, nx: usize, x: *mut la_buf) -> c_int;
    pub fn la_chol_solve_f64_host(l: *const f64, n: usize, b: *const f64, nx: usize, x: *mut f64) -> c_int;
    pub fn la_chol_solve_f32_host(l: *const f32, n: usize, b: *const f32, nx: usize, x: *mut f32) -> c_int;
    pub fn la_identity_f64(dst: *mut la_buf, n: usize) -> c_int;
    pub fn la_identity_f32(dst: *mut la_buf, n: usize) -> c_int;
    pub fn la_transpose_f64(src: *const la_buf, dst: *mut la_buf, rows: usize, cols: usize) -> c_int;
    pub fn la_transpose_f32(src: *const la_buf, dst: *mut la_buf, rows: usize, cols: usize) -> c_int;
    pub fn la_permute_rows_f64(src: *const la_buf, rows: usize, cols: usize, idx: *const u64, out_rows: usize,
                               dst: *mut la_buf) -> c_int;
    pub fn la_permute_rows_f32(src: *const la_buf, rows: usize, cols: usize, idx: *const u64, out_rows: usize,
                               dst: *mut la_buf) -> c_int;
    // ---- multi-GPU Mul, QR, elementwise operators / norms, fp32 accuracy mode (added in round 2) ----
    pub fn la_elementwise_f32(op: c_int, a: *const la_buf, b: *const la_buf, scalar: f32, c: *mut la_buf, count: usize) -> c_int;
    pub fn la_elementwise_f32_dev(op: c_int, a: *const f32, b: *const f32, scalar: f32, c: *mut f32, count: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn la_elementwise_f64(op: c_int, a: *const la_buf, b: *const la_buf, scalar: f64, c: *mut la_buf, count: usize) -> c_int;
    pub fn la_elementwise_f64_dev(op: c_int, a: *const f64, b: *const f64, scalar: f64, c: *mut f64, count: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn la_gemm_f32_mg(ngpus: c_int, devices: *const c_int, a: *const f32, b: *const f32, c: *mut f32, m: usize, k: usize, n: usize) -> c_int;
    pub fn la_gemm_f32_mg_rank(ctx: *mut la_mg, a_shard: *const f32, lda: usize, c_shard: *mut f32, ldc: usize, m_local: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn la_gemm_f32_mg_rank_host(ctx: *mut la_mg, a_shard: *const f32, b_block: *const f32, ldb: usize, c_shard: *mut f32, m_local: usize) -> c_int;
    pub fn la_gemm_f64_mg(ngpus: c_int, devices: *const c_int, a: *const f64, b: *const f64, c: *mut f64, m: usize, k: usize, n: usize) -> c_int;
    pub fn la_gemm_f64_mg_rank(ctx: *mut la_mg, a_shard: *const f64, lda: usize, c_shard: *mut f64, ldc: usize, m_local: usize, cuda_stream: *mut c_void) -> c_int;
    pub fn la_gemm_f64_mg_rank_host(ctx: *mut la_mg, a_shard: *const f64, b_block: *const f64, ldb: usize, c_shard: *mut f64, m_local: usize) -> c_int;
    pub fn la_get_gemm_f32_mode(out: *mut c_int) -> c_int;
    pub fn la_lu_factor_f64_mg(ngpus: c_int, devices: *const c_int, a: *const f64, lu_out: *mut f64, n: usize, piv_out: *mut u64, pospivsign_out: *mut c_int) -> c_int;
    pub fn la_lu_mg_create(ngpus: c_int, devices: *const c_int, n: usize, out: *mut *mut la_lu_mg) -> c_int;
    pub fn la_lu_mg_destroy(ctx: *mut la_lu_mg) -> c_int;
    pub fn la_lu_mg_devices(ctx: *const la_lu_mg, ndev_out: *mut c_int) -> c_int;
    pub fn la_lu_mg_download_f64(ctx: *mut la_lu_mg, lu_out: *mut f64, piv_out: *mut u64, pospivsign_out: *mut c_int) -> c_int;
    pub fn la_lu_mg_factor_f64(ctx: *mut la_lu_mg) -> c_int;
    pub fn la_lu_mg_fill_hash_f64(ctx: *mut la_lu_mg, seed: u64) -> c_int;
    pub fn la_lu_mg_last_ms(ctx: *mut la_lu_mg, ms_out: *mut f32) -> c_int;
    pub fn la_lu_mg_plan(n: usize, ngpus: c_int, sm_count: c_int, block_width_out: *mut c_int, nblocks_out: *mut c_int, ndev_out: *mut c_int, ncols_out: *mut usize) -> c_int;
    pub fn la_lu_mg_sync(ctx: *mut la_lu_mg) -> c_int;
    pub fn la_lu_mg_upload_f64(ctx: *mut la_lu_mg, a: *const f64) -> c_int;
    pub fn la_mg_b_block(ctx: *const la_mg, block_dev: *mut *mut c_void, ldb: *mut usize, col0: *mut usize, col1: *mut usize) -> c_int;
    pub fn la_mg_connect(ctx: *mut la_mg, handles: *const c_void) -> c_int;
    pub fn la_mg_create(rank: c_int, nranks: c_int, device: c_int, elem_bytes: usize, k: usize, n: usize, out: *mut *mut la_mg) -> c_int;
    pub fn la_mg_destroy(ctx: *mut la_mg) -> c_int;
    pub fn la_mg_handle(ctx: *const la_mg, handle_out: *mut c_void) -> c_int;
    pub fn la_mg_reserve(ctx: *mut la_mg, m_local: usize) -> c_int;
    pub fn la_mg_quiesce(ctx: *mut la_mg, cuda_stream: *mut c_void) -> c_int;
    pub fn la_mg_shard(nranks: c_int, rank: c_int, m: usize, n: usize, elem_bytes: usize, row0: *mut usize, row1: *mut usize, col0: *mut usize, col1: *mut usize) -> c_int;
    pub fn la_qr_factor_f32(qr_inout: *mut la_buf, m: usize, n: usize, rdiag: *mut la_buf, tmat: *mut la_buf) -> c_int;
    pub fn la_qr_factor_f32_dev(qr_inout: *mut f32, m: usize, n: usize, rdiag: *mut f32, tmat: *mut f32, cuda_stream: *mut c_void) -> c_int;
    pub fn la_qr_factor_f32_host(a: *const f32, qr_out: *mut f32, rdiag_out: *mut f32, m: usize, n: usize) -> c_int;
    pub fn la_qr_factor_f64(qr_inout: *mut la_buf, m: usize, n: usize, rdiag: *mut la_buf, tmat: *mut la_buf) -> c_int;
    pub fn la_qr_factor_f64_dev(qr_inout: *mut f64, m: usize, n: usize, rdiag: *mut f64, tmat: *mut f64, cuda_stream: *mut c_void) -> c_int;
    pub fn la_qr_factor_f64_host(a: *const f64, qr_out: *mut f64, rdiag_out: *mut f64, m: usize, n: usize) -> c_int;
    pub fn la_qr_get_q_f32(qr: *const la_buf, m: usize, n: usize, tmat: *const la_buf, q: *mut la_buf) -> c_int;
    pub fn la_qr_get_q_f64(qr: *const la_buf, m: usize, n: usize, tmat: *const la_buf, q: *mut la_buf) -> c_int;
    pub fn la_qr_get_r_f32(qr: *const la_buf, m: usize, n: usize, rdiag: *const la_buf, r: *mut la_buf) -> c_int;
    pub fn la_qr_get_r_f64(qr: *const la_buf, m: usize, n: usize, rdiag: *const la_buf, r: *mut la_buf) -> c_int;
    pub fn la_qr_solve_f32(qr: *const la_buf, m: usize, n: usize, rdiag: *const la_buf, b: *const la_buf, nx: usize, x: *mut la_buf) -> c_int;
    pub fn la_qr_solve_f64(qr: *const la_buf, m: usize, n: usize, rdiag: *const la_buf, b: *const la_buf, nx: usize, x: *mut la_buf) -> c_int;
    pub fn la_qr_tmat_elems(m: usize, n: usize, device: c_int, elem_bytes: usize, elems_out: *mut usize) -> c_int;
    pub fn la_reduce_f32(kind: c_int, a: *const la_buf, b: *const la_buf, count: usize, out: *mut f32) -> c_int;
    pub fn la_reduce_f32_dev(kind: c_int, a: *const f32, b: *const f32, count: usize, out_host: *mut f32, cuda_stream: *mut c_void) -> c_int;
    pub fn la_reduce_f64(kind: c_int, a: *const la_buf, b: *const la_buf, count: usize, out: *mut f64) -> c_int;
    pub fn la_reduce_f64_dev(kind: c_int, a: *const f64, b: *const f64, count: usize, out_host: *mut f64, cuda_stream: *mut c_void) -> c_int;
    pub fn la_set_gemm_f32_mode(mode: c_int) -> c_int;
    pub fn la_fill_hash_f64_dev(dst: *mut f64, count: usize, seed: u64, first_idx: u64, cuda_stream: *mut c_void) -> c_int;
    pub fn la_fill_hash_f32_dev(dst: *mut f32, count: usize, seed: u64, first_idx: u64, cuda_stream: *mut c_void) -> c_int;
    pub fn la_debug_set_gemm_path(path: c_int) -> c_int;
    pub fn la_debug_set_gemm_f32_path(path: c_int) -> c_int;
}

/// Panics with the library's message when `status != LA_OK` -- contract violations panic in the reference too
/// (`assert!`, src/matrix/mod.rs:961; src/decomp/lu.rs:225,240).
pub fn check(status: c_int) {
    if status != LA_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(la_last_error()) }.to_string_lossy().into_owned();
        panic!("la_b200 status {}: {}", status, msg);
    }
}
