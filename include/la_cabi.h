/*
 * la_cabi.h -- C ABI of the B200-native dense hot path behind rust-la's public API.
 *
 * The reference (xasmx/rust-la, crate `la`) has NO FFI/plugin interface: its boundary is the Rust API
 * (`Matrix<T>`, operator `*`, `LUDecomposition<T>`).  This header is the seam a maintainer binds from a new
 * `src/ffi.rs`: every entry point below names the reference item it replaces
 * (file:line relative to the reference repository root).
 *
 * Conventions
 *   - all matrices are dense ROW-MAJOR with tight leading dimension unless an explicit `ld*` is given
 *     (reference layout: data[r * cols + c], src/matrix/mod.rs:26-30, :557-560);
 *   - every function returns an `int` status, LA_OK == 0.  Shape/contract violations are the CALLER's job
 *     (the Rust side keeps the reference's `assert!`s so the panic happens before any FFI call); the library
 *     still validates and returns LA_ERR_INVALID instead of computing garbage;
 *   - numerical singularity is NOT an error: `la_lu_is_nonsingular_*` reports it and the Rust side maps it to
 *     `Option::None` exactly like src/decomp/lu.rs:241-243;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with LA_ERR_NO_DEVICE;
 *   - thread safety: all entry points may be called concurrently from many host threads (Matrix<T> is
 *     Send + Sync in the reference).  `la_last_error` is thread-local.  `*_host` entry points run on a
 *     per-host-thread CUDA stream; `*_dev` entry points run on the stream the caller passes.
 *   - `piv` has the reference's semantics (src/decomp/lu.rs:108-111,147-149): piv[i] = index of the ORIGINAL
 *     row that ends up in row i, i.e. A(piv,:) = L*U.  Rust `usize` == uint64_t on the supported target.
 */
#ifndef LA_CABI_H
#define LA_CABI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LA_API __attribute__((visibility("default")))
#else
#define LA_API
#endif

enum {
  LA_OK = 0,
  LA_ERR_INVALID = 1,   /* bad argument: null pointer, zero dimension, size overflow, misaligned pointer        */
  LA_ERR_CUDA = 2,      /* a CUDA runtime/driver call failed; la_last_error() has the text                      */
  LA_ERR_NOMEM = 3,     /* device or pinned-host allocation failed                                              */
  LA_ERR_NO_DEVICE = 4, /* no usable sm_100 device (the library never falls back to the CPU)                    */
  LA_ERR_UNSUPPORTED = 5
};

/* Epilogue of the device GEMM: what is done with the product P = A*B. */
enum {
  LA_GEMM_ASSIGN = 0, /* C  = P   (operator Mul / mmul)                                   */
  LA_GEMM_SUB = 1,    /* C -= P   (LU trailing update, src/decomp/lu.rs:122-129, i > j)   */
  LA_GEMM_ADD = 2     /* C += P   (K-panel pipelined multi-GPU Mul)                       */
};

/* Arithmetic of the fp32 tensor-core Mul (la_set_gemm_f32_mode).  The reference's fp32 Mul is exact fp32 arithmetic
 * (src/matrix/mod.rs:965-973 with T = f32), so the DEFAULT is the split-compensated form. */
enum {
  LA_F32_3XTF32 = 0, /* default: every operand split into TF32 big + small parts, three tcgen05 passes, fp32 accumulate:
                        relative error ~1e-6 of |A||B| (fp32-grade), about a third of the TF32 rate               */
  LA_F32_TF32 = 1    /* opt-in: inputs rounded to TF32 (10 mantissa bits, ~1e-3 of |A||B|), full tensor rate; used
                        only for k >= 32 where BASELINE's 1e-4*k bar covers it                                     */
};

typedef struct la_buf la_buf; /* opaque device buffer handle */

/* ---- library / device ---------------------------------------------------------------------------- */
LA_API int la_version(void);                      /* 100 * major + minor */
LA_API const char* la_last_error(void);           /* thread-local, never NULL */
LA_API int la_device_count(int* out);             /* number of visible CUDA devices */
LA_API int la_device_sm_count(int device, int* out);
LA_API int la_sync(int device);                   /* waits for all work this library queued on `device` */

/* ---- buffers: replace `alloc_dirty_vec` (src/internalutil.rs:7-13) and the `Vec<T>` backing of
 *      `Matrix<T>` (src/matrix/mod.rs:26-30).  Contents of a fresh buffer are unspecified ("dirty"). ---- */
LA_API int la_buf_alloc(size_t bytes, int device, la_buf** out);
LA_API int la_buf_free(la_buf* buf);
LA_API int la_buf_upload(la_buf* dst, size_t dst_offset_bytes, const void* host, size_t bytes);
LA_API int la_buf_download(const la_buf* src, size_t src_offset_bytes, void* host, size_t bytes);
LA_API int la_buf_copy(la_buf* dst, const la_buf* src, size_t bytes); /* `ludata = a.get_data().clone()`, lu.rs:105 */
LA_API void* la_buf_device_ptr(const la_buf* buf);
LA_API size_t la_buf_bytes(const la_buf* buf);
LA_API int la_buf_device(const la_buf* buf);
/* pinned host memory for the host mirror of a device-backed Matrix (fast H2D/D2H) */
LA_API int la_host_alloc(size_t bytes, void** out);
LA_API int la_host_free(void* ptr);

/* ---- GEMM: replaces `impl Mul<&Matrix<T>> for &Matrix<T>` (src/matrix/mod.rs:957-980, loop nest :965-973)
 *      and `Matrix::mmul` (src/matrix/mmatrix.rs:82-98).  C[m x n] = A[m x k] * B[k x n]. ---- */
LA_API int la_gemm_f64(const la_buf* A, const la_buf* B, la_buf* C, size_t m, size_t k, size_t n);
LA_API int la_gemm_f32(const la_buf* A, const la_buf* B, la_buf* C, size_t m, size_t k, size_t n);
/* host-pointer form (what `&a * &b` on host-resident matrices binds): H2D, kernel, D2H, synchronous */
LA_API int la_gemm_f64_host(const double* A, const double* B, double* C, size_t m, size_t k, size_t n);
LA_API int la_gemm_f32_host(const float* A, const float* B, float* C, size_t m, size_t k, size_t n);
/* integer instance: the reference's Mul is generic over T and is unit-tested on integer matrices
 * (src/matrix/mod.rs:1479-1484, src/matrix/mmatrix.rs:234-241).  Exact, two's-complement wrapping. */
LA_API int la_gemm_i64_host(const int64_t* A, const int64_t* B, int64_t* C, size_t m, size_t k, size_t n);
/* raw device pointers, explicit leading dimensions (in elements), epilogue mode, caller's stream
 * (`cuda_stream` is a cudaStream_t; NULL = the per-thread default stream).  Asynchronous. */
LA_API int la_gemm_f64_dev(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc,
                           size_t m, size_t k, size_t n, int mode, void* cuda_stream);
LA_API int la_gemm_f32_dev(const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc,
                           size_t m, size_t k, size_t n, int mode, void* cuda_stream);

/* ---- multi-GPU Mul: the same product, src/matrix/mod.rs:957-980, at N GPUs of one node.  GPU g owns
 *      a row block of A and C; rank q owns a COLUMN block of B (it uploads / produces only that) inside its full-size
 *      replica, and every rank pulls the other blocks over NVLink with the library's own kernel (peer-mapped loads,
 *      device-side ready/ack flags) while it already multiplies by the blocks it has.  No reduction: K is not split. ---- */
/* one process, `ngpus` distinct devices, host operands; synchronous.  Small products run on devices[0] alone. */
LA_API int la_gemm_f64_mg(int ngpus, const int* devices, const double* A, const double* B, double* C, size_t m, size_t k,
                          size_t n);
LA_API int la_gemm_f32_mg(int ngpus, const int* devices, const float* A, const float* B, float* C, size_t m, size_t k,
                          size_t n);
/* One process per GPU (torchrun / MPI style).  Each rank creates a context (it allocates the rank's replica of B,
 * k x n, row-major), publishes LA_MG_HANDLE_BYTES of handle, and connects with all ranks' handles in rank order (the
 * launcher moves the bytes: MPI_Allgather, torch.distributed.all_gather, a file ...).  la_mg_shard is the partition
 * every rank must agree on -- pure arithmetic, callable without a device: rows [row0,row1) of A and C, columns
 * [col0,col1) of B. */
#define LA_MG_HANDLE_BYTES 256
typedef struct la_mg la_mg;
LA_API int la_mg_shard(int nranks, int rank, size_t m, size_t n, size_t elem_bytes, size_t* row0, size_t* row1,
                       size_t* col0, size_t* col1);
LA_API int la_mg_create(int rank, int nranks, int device, size_t elem_bytes, size_t k, size_t n, la_mg** out);
LA_API int la_mg_handle(const la_mg* ctx, void* handle_out /* LA_MG_HANDLE_BYTES */);
LA_API int la_mg_connect(la_mg* ctx, const void* handles /* nranks * LA_MG_HANDLE_BYTES, rank order */);
LA_API int la_mg_destroy(la_mg* ctx);
/* where this rank's column block lives: *block_dev = &replica[0][col0], leading dimension *ldb (= n) elements */
LA_API int la_mg_b_block(const la_mg* ctx, void** block_dev, size_t* ldb, size_t* col0, size_t* col1);
/* C_shard[m_local x n] = A_shard[m_local x k] * B, all device-resident; the rank's own column block must already be in
 * its replica (written earlier on `cuda_stream`).  Collective: every rank calls it once per product.  Asynchronous. */
LA_API int la_gemm_f64_mg_rank(la_mg* ctx, const double* A_shard, size_t lda, double* C_shard, size_t ldc, size_t m_local,
                               void* cuda_stream);
LA_API int la_gemm_f32_mg_rank(la_mg* ctx, const float* A_shard, size_t lda, float* C_shard, size_t ldc, size_t m_local,
                               void* cuda_stream);
/* host shards: A rows (tight), the rank's column block of B (k rows, leading dimension ldb elements), C rows out (tight).
 * Uploads, pulls, multiplies and downloads in one pipeline; synchronous. */
LA_API int la_gemm_f64_mg_rank_host(la_mg* ctx, const double* A_shard, const double* B_block, size_t ldb, double* C_shard,
                                    size_t m_local);
LA_API int la_gemm_f32_mg_rank_host(la_mg* ctx, const float* A_shard, const float* B_block, size_t ldb, float* C_shard,
                                    size_t m_local);
/* Sizes the context's device copies of a host shard of `m_local` rows (la_gemm_*_mg_rank_host allocates on demand
 * otherwise).  Call it on every rank after la_mg_connect and before the ranks start multiplying when several ranks share
 * a device or a process: cudaMalloc may wait for the device, and a rank that allocates before it has published its block
 * would then wait for peers whose kernels are waiting for that block. */
LA_API int la_mg_reserve(la_mg* ctx, size_t m_local);
/* Before a rank rewrites its column block IN PLACE for the next product: makes `cuda_stream` wait until every peer has
 * finished pulling the block of the previous product. */
LA_API int la_mg_quiesce(la_mg* ctx, void* cuda_stream);

/* ---- LU: replaces `LUDecomposition::new` (src/decomp/lu.rs:104-168).  Factorises IN PLACE the m x n
 *      row-major matrix in `LU` (the caller clones A first, lu.rs:105).  Outputs: packed L\U, the permutation
 *      vector `piv_out[m]` and `*pospivsign_out` (1 = even number of swaps, lu.rs:113,151).
 *      Pivot rule: largest |x| at or below the diagonal, lowest row index wins ties, NaN never displaces the
 *      incumbent (lu.rs:132-137); a zero pivot skips the division and the factorisation continues (:156-160). */
LA_API int la_lu_factor_f64(la_buf* LU, size_t m, size_t n, uint64_t* piv_out, int* pospivsign_out);
LA_API int la_lu_factor_f32(la_buf* LU, size_t m, size_t n, uint64_t* piv_out, int* pospivsign_out);
LA_API int la_lu_factor_f64_host(const double* A, double* LU_out, size_t m, size_t n, uint64_t* piv_out,
                                 int* pospivsign_out);
LA_API int la_lu_factor_f32_host(const float* A, float* LU_out, size_t m, size_t n, uint64_t* piv_out,
                                 int* pospivsign_out);
/* device-pointer form: `piv_dev` is uint64_t[m] and `sign_dev` is int[1] in DEVICE memory.  Asynchronous. */
LA_API int la_lu_factor_f64_dev(double* LU, size_t m, size_t n, uint64_t* piv_dev, int* sign_dev, void* cuda_stream);
LA_API int la_lu_factor_f32_dev(float* LU, size_t m, size_t n, uint64_t* piv_dev, int* sign_dev, void* cuda_stream);

/* ---- LU of one n x n fp64 matrix across several GPUs of one node, driven by one host thread (SURVEY.md 8(f) rank 4;
 * same result contract as la_lu_factor_f64: `LUDecomposition::new`, src/decomp/lu.rs:104-168).  128-column blocks are dealt
 * round-robin to the devices; the panel owner's factored block column travels to every device by peer copy; streams,
 * events and copies only -- no collective library, no host round trip.  A device may be listed more than once (its
 * entries then share it: the 1-GPU test configuration).  Contexts are not thread-safe; one factorisation at a time.
 *   create -> upload (or fill_hash: element (i, j) = hash(seed, i * n + j), the bench / test generator) -> factor (queues
 *   the work and returns) -> download (waits; any output may be NULL) -> ... -> destroy.
 * la_lu_mg_last_ms: device time of the last factorisation (events on the first device, which waits for all others). */
typedef struct la_lu_mg la_lu_mg;
LA_API int la_lu_mg_create(int ngpus, const int* devices, size_t n, la_lu_mg** out);
LA_API int la_lu_mg_destroy(la_lu_mg* ctx);
LA_API int la_lu_mg_devices(const la_lu_mg* ctx, int* ndev_out); /* devices in use: min(ngpus, number of block columns) */
/* The layout as pure arithmetic (no device needed): block-column width (128, narrower when a panel of n rows would not fit
 * the shared memory of `sm_count` SMs), number of block columns, devices in use and the local column count of each
 * (ncols_out[ngpus]).  Block column b lives on device b % ndev at local column (b / ndev) * width. */
LA_API int la_lu_mg_plan(size_t n, int ngpus, int sm_count, int* block_width_out, int* nblocks_out, int* ndev_out,
                         size_t* ncols_out);
LA_API int la_lu_mg_upload_f64(la_lu_mg* ctx, const double* A /* host, row-major n x n */);
LA_API int la_lu_mg_fill_hash_f64(la_lu_mg* ctx, uint64_t seed);
LA_API int la_lu_mg_factor_f64(la_lu_mg* ctx);
LA_API int la_lu_mg_sync(la_lu_mg* ctx);
LA_API int la_lu_mg_download_f64(la_lu_mg* ctx, double* LU_out, uint64_t* piv_out, int* pospivsign_out);
LA_API int la_lu_mg_last_ms(la_lu_mg* ctx, float* ms_out);
/* one call from host memory (create + upload + factor + download + destroy) */
LA_API int la_lu_factor_f64_mg(int ngpus, const int* devices, const double* A, double* LU_out, size_t n, uint64_t* piv_out,
                               int* pospivsign_out);

/* `is_non_singular` (src/decomp/lu.rs:174-182): *out = 0 iff some LU[j*n+j] == 0 exactly, j < n. */
LA_API int la_lu_is_nonsingular_f64(const la_buf* LU, size_t n, int* out);
LA_API int la_lu_is_nonsingular_f32(const la_buf* LU, size_t n, int* out);
/* `det` (src/decomp/lu.rs:224-232): (+1|-1) * LU[0][0] * LU[1][1] * ... multiplied in index order. */
LA_API int la_lu_det_f64(const la_buf* LU, size_t n, int pospivsign, double* out);
LA_API int la_lu_det_f32(const la_buf* LU, size_t n, int pospivsign, float* out);
/* `solve` (src/decomp/lu.rs:237-278) for an n x n factorisation: X[n x nx] = A^-1 * B[n x nx].  `piv` is a HOST
 * array of n entries.  The caller has already checked non-singularity (lu.rs:241-243 -> None). */
LA_API int la_lu_solve_f64(const la_buf* LU, size_t m, size_t n, const uint64_t* piv, const la_buf* B, size_t nx,
                           la_buf* X);
LA_API int la_lu_solve_f32(const la_buf* LU, size_t m, size_t n, const uint64_t* piv, const la_buf* B, size_t nx,
                           la_buf* X);
LA_API int la_lu_solve_f64_host(const double* LU, size_t m, size_t n, const uint64_t* piv, const double* B, size_t nx,
                                double* X);
LA_API int la_lu_solve_f32_host(const float* LU, size_t m, size_t n, const uint64_t* piv, const float* B, size_t nx,
                                float* X);
LA_API int la_lu_solve_f64_dev(const double* LU, size_t n, const uint64_t* piv_dev, const double* B, size_t nx,
                               double* X, void* cuda_stream);
LA_API int la_lu_solve_f32_dev(const float* LU, size_t n, const uint64_t* piv_dev, const float* B, size_t nx,
                               float* X, void* cuda_stream);

/* ---- Cholesky (widening step) ------------------------------------------------ */
/* `CholeskyDecomposition::new` (src/decomp/cholesky.rs:56-110), in place on a device buffer holding the n x n matrix:
 * on return *ok_out = 1 and the buffer holds L (lower triangle, zeros above), or *ok_out = 0 -- the reference's `None`:
 * not symmetric (exact `!=` on every pair, :91-93) or not positive definite (`a[j][j] - sum <= 0`, :99-102); the buffer
 * content is then unspecified.  "Not square" (:57-59) is the caller's check, as are all shape panics. */
LA_API int la_chol_factor_f64(la_buf* A_inout, size_t n, int* ok_out);
LA_API int la_chol_factor_f32(la_buf* A_inout, size_t n, int* ok_out);
LA_API int la_chol_factor_f64_host(const double* A, double* L_out, size_t n, int* ok_out);
LA_API int la_chol_factor_f32_host(const float* A, float* L_out, size_t n, int* ok_out);
/* Device-pointer forms, asynchronous on `cuda_stream` (NULL = the calling thread's stream).  flags_dev points to two
 * device ints that are cleared and then set: [0] != 0 = not symmetric, [1] != 0 = not positive definite. */
LA_API int la_chol_factor_f64_dev(double* A_inout, size_t n, int* flags_dev, void* cuda_stream);
LA_API int la_chol_factor_f32_dev(float* A_inout, size_t n, int* flags_dev, void* cuda_stream);
LA_API int la_chol_solve_f64_dev(const double* L, size_t n, const double* B, size_t nx, double* X, void* cuda_stream);
LA_API int la_chol_solve_f32_dev(const float* L, size_t n, const float* B, size_t nx, float* X, void* cuda_stream);
/* `CholeskyDecomposition::solve` (cholesky.rs:116-144): X (n x nx) = A^-1 B from L; B and X distinct buffers. */
LA_API int la_chol_solve_f64(const la_buf* L, size_t n, const la_buf* B, size_t nx, la_buf* X);
LA_API int la_chol_solve_f32(const la_buf* L, size_t n, const la_buf* B, size_t nx, la_buf* X);
LA_API int la_chol_solve_f64_host(const double* L, size_t n, const double* B, size_t nx, double* X);
LA_API int la_chol_solve_f32_host(const float* L, size_t n, const float* B, size_t nx, float* X);

/* ---- QR (widening step) ----------------------------------------------------- */
/* `QRDecomposition::new` (src/decomp/qr.rs:26-106), in place on a device buffer holding the m x n matrix: on return the
 * buffer is the reference's packed `qr` (R strictly above the diagonal, the UNNORMALISED Householder vectors u_k = x - a e_k
 * from the diagonal down) and `rdiag` holds min(m, n) values a_k = -+|x| (:58).  Blocked compact-WY on the device; `tmat`
 * (la_qr_tmat_elems values) receives the triangular factor T' of every 128-column block, which la_qr_get_q needs. */
LA_API int la_qr_tmat_elems(size_t m, size_t n, int device, size_t elem_bytes, size_t* elems_out);
LA_API int la_qr_factor_f64(la_buf* QR_inout, size_t m, size_t n, la_buf* rdiag, la_buf* tmat);
LA_API int la_qr_factor_f32(la_buf* QR_inout, size_t m, size_t n, la_buf* rdiag, la_buf* tmat);
LA_API int la_qr_factor_f64_host(const double* A, double* QR_out, double* rdiag_out, size_t m, size_t n);
LA_API int la_qr_factor_f32_host(const float* A, float* QR_out, float* rdiag_out, size_t m, size_t n);
LA_API int la_qr_factor_f64_dev(double* QR_inout, size_t m, size_t n, double* rdiag, double* tmat, void* cuda_stream);
LA_API int la_qr_factor_f32_dev(float* QR_inout, size_t m, size_t n, float* rdiag, float* tmat, void* cuda_stream);
/* `get_r` (qr.rs:138-152): R (m x n) = strict upper part of qr with rdiag on the diagonal, device-resident so that
 * pinverse's `(r.t() * &r).inverse() * &a.t()` (src/matrix/mod.rs:1049-1057) stays in HBM. */
LA_API int la_qr_get_r_f64(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, la_buf* R);
LA_API int la_qr_get_r_f32(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, la_buf* R);
/* `get_q` (qr.rs:155-194): Q (m x m) = H_1 (H_2 (... I')), I' the identity on the first min(m, n) diagonal entries. */
LA_API int la_qr_get_q_f64(const la_buf* QR, size_t m, size_t n, const la_buf* tmat, la_buf* Q);
LA_API int la_qr_get_q_f32(const la_buf* QR, size_t m, size_t n, const la_buf* tmat, la_buf* Q);
/* `solve` (qr.rs:199-238), the arithmetic only: X is the reference's full m x nx work array after both phases (the
 * caller returns its first n rows when m == n and reproduces the panic of `Matrix::new(cols, nx, ..)` (:237) otherwise;
 * `is_full_rank` (:110-117) is the caller's check on rdiag).  Requires n <= m.  The first phase applies I - u u'/u_k as
 * the reference does (not a reflection for these unnormalised vectors): parity, not least squares. */
LA_API int la_qr_solve_f64(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, const la_buf* B, size_t nx, la_buf* X);
LA_API int la_qr_solve_f32(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, const la_buf* B, size_t nx, la_buf* X);

/* ---- elementwise operators and norms on device-resident data ------------------ */
enum la_elementwise_op {
  LA_EW_ADD = 0,   /* `Add`      src/matrix/mod.rs:874-890:  c[i] = a[i] + b[i]       */
  LA_EW_SUB = 1,   /* `Sub`      src/matrix/mod.rs:913-929:  c[i] = a[i] - b[i]       */
  LA_EW_MUL = 2,   /* `elem_mul` src/matrix/mod.rs:499-512:  c[i] = a[i] * b[i]       */
  LA_EW_DIV = 3,   /* `elem_div` src/matrix/mod.rs:514-527:  c[i] = a[i] / b[i]       */
  LA_EW_SCALE = 4, /* `scale`    src/matrix/mod.rs:487-497:  c[i] = scalar * a[i]     */
  LA_EW_NEG = 5    /* `Neg`      src/matrix/mod.rs:853-866:  c[i] = -a[i]             */
};
enum la_reduce_kind {
  LA_RED_SUMSQ = 0,   /* `frobenius_norm` / `vector_euclidean_norm` src/matrix/mod.rs:1059-1068, :1094-1101: sqrt(sum a[i]^2) */
  LA_RED_ABS_SUM = 1, /* `vector_1_norm` src/matrix/mod.rs:1075-1084: sum |a[i]|                                           */
  LA_RED_ABS_MAX = 2, /* `vector_inf_norm` src/matrix/mod.rs:1103-1115: max |a[i]| (strict `>`: a NaN never replaces)      */
  LA_RED_DOT = 3      /* `dot` src/matrix/mod.rs:529-: sum a[i] * b[i]                                                      */
};
/* One IEEE operation per element, `count` elements (shape checks are the caller's panics): bit-identical to the reference.
 * B may be NULL for LA_EW_SCALE / LA_EW_NEG; C may alias A or B. */
LA_API int la_elementwise_f64(int op, const la_buf* A, const la_buf* B, double scalar, la_buf* C, size_t count);
LA_API int la_elementwise_f32(int op, const la_buf* A, const la_buf* B, float scalar, la_buf* C, size_t count);
LA_API int la_elementwise_f64_dev(int op, const double* A, const double* B, double scalar, double* C, size_t count,
                                  void* cuda_stream);
LA_API int la_elementwise_f32_dev(int op, const float* A, const float* B, float scalar, float* C, size_t count,
                                  void* cuda_stream);
/* Reductions to one host scalar (synchronous).  Fixed-shape tree instead of the reference's sequential sum: equal up to
 * rounding (<= count * eps relative for the sums of non-negative terms), exact for LA_RED_ABS_MAX. */
LA_API int la_reduce_f64(int kind, const la_buf* A, const la_buf* B, size_t count, double* out);
LA_API int la_reduce_f32(int kind, const la_buf* A, const la_buf* B, size_t count, float* out);
LA_API int la_reduce_f64_dev(int kind, const double* A, const double* B, size_t count, double* out_host, void* cuda_stream);
LA_API int la_reduce_f32_dev(int kind, const float* A, const float* B, size_t count, float* out_host, void* cuda_stream);

/* ---- aux ------------------------------------------------------------------------------------------ */
/* `Matrix::id` (src/matrix/mod.rs:416-426), the RHS of `inverse` (mod.rs:1034-1037) */
LA_API int la_identity_f64(la_buf* dst, size_t n);
LA_API int la_identity_f32(la_buf* dst, size_t n);
/* `Matrix::t` (src/matrix/mod.rs:653-669): dst (cols x rows) = transpose of src (rows x cols), both device-resident, so
 * that chains like pinverse's `(r.t() * &r).inverse() * &a.t()` (mod.rs:1049-1057) never leave HBM.  No aliasing. */
LA_API int la_transpose_f64(const la_buf* src, la_buf* dst, size_t rows, size_t cols);
LA_API int la_transpose_f32(const la_buf* src, la_buf* dst, size_t rows, size_t cols);
/* `Matrix::permute_rows` (src/matrix/mod.rs:757-759): dst row i = src row idx[i], i < out_rows; idx is HOST memory
 * (a `&[usize]`), every entry < rows or LA_ERR_INVALID (the reference panics on the out-of-range index). */
LA_API int la_permute_rows_f64(const la_buf* src, size_t rows, size_t cols, const uint64_t* idx, size_t out_rows,
                               la_buf* dst);
LA_API int la_permute_rows_f32(const la_buf* src, size_t rows, size_t cols, const uint64_t* idx, size_t out_rows,
                               la_buf* dst);
/* Counter-based synthetic inputs, uniform [0,1) like `Matrix::random` (src/matrix/mod.rs:842-851):
 * element i of dst gets hash(seed, first_idx + i).  Device pointers; asynchronous on `cuda_stream`. */
LA_API int la_fill_hash_f64_dev(double* dst, size_t count, uint64_t seed, uint64_t first_idx, void* cuda_stream);
LA_API int la_fill_hash_f32_dev(float* dst, size_t count, uint64_t seed, uint64_t first_idx, void* cuda_stream);

/* Process-wide accuracy mode of the fp32 tensor-core product (LA_F32_*; initial value from the environment variable
 * LA_GEMM_F32_MODE = "3xtf32" | "tf32").  Products that are small (m*n*k <= 128^3), unaligned or LA_GEMM_SUB always run
 * the exact reference-order CUDA-core kernel, whatever the mode. */
LA_API int la_set_gemm_f32_mode(int mode);
LA_API int la_get_gemm_f32_mode(int* out);

/* Test hooks (not part of the drop-in surface).  fp64: 0 = automatic kernel choice, 1 = force the generic CUDA-core GEMM,
 * 2 = force the TMA/DMMA GEMM (fails with LA_ERR_INVALID when the operands are not TMA-addressable), 3 / 4 = force its
 * 128x64 / 128x128 tile configuration. */
LA_API int la_debug_set_gemm_path(int path);
/* fp32: 0 = automatic, 1 = force the CUDA-core kernel, 2 = force the tcgen05 TF32 kernel */
LA_API int la_debug_set_gemm_f32_path(int path);

#ifdef __cplusplus
}
#endif
#endif /* LA_CABI_H */
