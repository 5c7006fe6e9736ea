// capi.cu -- the extern "C" surface declared in include/la_cabi.h.  Thin glue: argument validation, device selection,
// host<->device staging for the *_host forms.  All arithmetic lives in the kernels (gemm_*.cu, lu*.cu).
#include <string.h>

#include "la_common.cuh"

using namespace la;

struct la_buf {
  void* ptr;
  size_t bytes;
  int device;
};

namespace la {
void debug_set_gemm_path(int p);
void debug_set_gemm_f32_path(int p);
int set_gemm_f32_mode(int mode);
int get_gemm_f32_mode();
}

// *_dev entry points: the caller's stream on the CURRENT device (see CallScope)
#define DEV_SCOPE(stream_arg)                  \
  int _scope_dev = -1;                         \
  if (cudaGetDevice(&_scope_dev) != cudaSuccess) { cudaGetLastError(); _scope_dev = -1; } \
  la::CallScope _scope(_scope_dev, la::resolve_stream(stream_arg))

namespace {

struct DeviceGuard {
  int prev = -1;
  bool active = false;
  int enter(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(LA_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                  cudaGetErrorString(e));
    }
    if (prev != device) {
      LA_CUDA_TRY(cudaSetDevice(device));
      active = true;
    }
    return LA_OK;
  }
  ~DeviceGuard() {
    if (active) cudaSetDevice(prev);
  }
};

bool mul_overflows(size_t a, size_t b, size_t elem, size_t* out) {
  if (a != 0 && b > SIZE_MAX / a) return true;
  size_t ab = a * b;
  if (elem != 0 && ab > SIZE_MAX / elem) return true;
  *out = ab * elem;
  return false;
}

template <typename T>
int gemm_bufs(const la_buf* A, const la_buf* B, la_buf* C, size_t m, size_t k, size_t n) {
  LA_REQUIRE(A && B && C, "la_gemm: null buffer handle");
  LA_REQUIRE(m > 0 && k > 0 && n > 0, "la_gemm: zero dimension (m=%zu k=%zu n=%zu)", m, k, n);
  size_t ba, bb, bc;
  LA_REQUIRE(!mul_overflows(m, k, sizeof(T), &ba) && !mul_overflows(k, n, sizeof(T), &bb) &&
                 !mul_overflows(m, n, sizeof(T), &bc),
             "la_gemm: size overflow");
  LA_REQUIRE(A->bytes >= ba && B->bytes >= bb && C->bytes >= bc, "la_gemm: buffer smaller than the matrix it must hold");
  LA_REQUIRE(A->device == B->device && A->device == C->device, "la_gemm: buffers live on different devices");
  LA_REQUIRE(C->ptr != A->ptr && C->ptr != B->ptr, "la_gemm: output aliases an input");
  DeviceGuard g;
  LA_TRY(g.enter(A->device));
  CallScope scope(A->device, cudaStreamPerThread);
  return gemm_dev<T>((const T*)A->ptr, k, (const T*)B->ptr, n, (T*)C->ptr, n, m, k, n, LA_GEMM_ASSIGN,
                     cudaStreamPerThread);
}

// Host-pointer Mul: H2D, kernel, D2H on the calling thread's stream.  A is uploaded and multiplied in row blocks so
// that the copy of block i+1 and the download of block i-1 overlap the kernel of block i (B is needed by every block
// and goes first).
template <typename T>
int gemm_host(const T* A, const T* B, T* C, size_t m, size_t k, size_t n) {
  LA_REQUIRE(A && B && C, "la_gemm_host: null pointer");
  LA_REQUIRE(m > 0 && k > 0 && n > 0, "la_gemm_host: zero dimension (m=%zu k=%zu n=%zu)", m, k, n);
  size_t ba, bb, bc;
  LA_REQUIRE(!mul_overflows(m, k, sizeof(T), &ba) && !mul_overflows(k, n, sizeof(T), &bb) &&
                 !mul_overflows(m, n, sizeof(T), &bc),
             "la_gemm_host: size overflow");
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  CallScope scope(ctx->device, cudaStreamPerThread);
  void *dA, *dB, *dC;
  LA_TRY(scratch_get(ctx->device, 0, ba, &dA));
  LA_TRY(scratch_get(ctx->device, 1, bb, &dB));
  LA_TRY(scratch_get(ctx->device, 2, bc, &dC));
  cudaStream_t st = cudaStreamPerThread;

  // small problems: one shot
  const size_t flops2 = m * n;  // proxy; block pipelining only pays for big outputs
  size_t blocks = 1;
  if (flops2 >= (size_t)4096 * 4096 && m >= 1024) blocks = 8;
  if (blocks == 1) {
    LA_CUDA_TRY(cudaMemcpyAsync(dB, B, bb, cudaMemcpyHostToDevice, st));
    LA_CUDA_TRY(cudaMemcpyAsync(dA, A, ba, cudaMemcpyHostToDevice, st));
    LA_TRY(gemm_dev<T>((const T*)dA, k, (const T*)dB, n, (T*)dC, n, m, k, n, LA_GEMM_ASSIGN, st));
    LA_CUDA_TRY(cudaMemcpyAsync(C, dC, bc, cudaMemcpyDeviceToHost, st));
    LA_CUDA_TRY(cudaStreamSynchronize(st));
    return LA_OK;
  }
  // pipelined: copy stream(s) + compute stream, events between them
  static thread_local cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  static thread_local cudaEvent_t ev_up[16], ev_done[16];
  static thread_local bool ev_init = false;
  if (!s_h2d) {
    LA_CUDA_TRY(cudaStreamCreateWithFlags(&s_h2d, cudaStreamNonBlocking));
    LA_CUDA_TRY(cudaStreamCreateWithFlags(&s_d2h, cudaStreamNonBlocking));
  }
  if (!ev_init) {
    for (int i = 0; i < 16; ++i) {
      LA_CUDA_TRY(cudaEventCreateWithFlags(&ev_up[i], cudaEventDisableTiming));
      LA_CUDA_TRY(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    }
    ev_init = true;
  }
  // rows per block: multiple of 128 (the GEMM tile)
  size_t rows_blk = ((m + blocks - 1) / blocks + 127) / 128 * 128;
  blocks = (m + rows_blk - 1) / rows_blk;
  static thread_local cudaEvent_t ev_b2 = nullptr;
  if (!ev_b2) LA_CUDA_TRY(cudaEventCreateWithFlags(&ev_b2, cudaEventDisableTiming));

  // Deep products: two phases so that neither the upload of B nor the download of C is exposed.
  //   phase 1 (columns [0, kh) of A / rows [0, kh) of B): K-panels -- panel p of A (strided) and of B go up, C (+)= A_p * B_p
  //            over ALL rows starts as soon as the first pair has landed and hides the rest of the upload;
  //   phase 2 (the other half of K): by row blocks, each block's rows of C are final after its GEMM and go down while
  //            the next block multiplies.
  // The accumulation order over k differs from a single launch only by where the partial sums are rounded into C.
  size_t kh = 0;
  if (k >= 2048) {
    const size_t kp = 1024;
    kh = (k * 5 / 8) / kp * kp;  // phase 2 keeps enough of K to cover the download of C
    // the first panel is narrow (256): the first GEMM starts after 1/4 of a full panel's upload
    for (size_t k0 = 0, p = 0; k0 < kh; ++p) {
      const size_t w = (p == 0) ? 256 : ((p == 1) ? kp - 256 : kp);
      LA_CUDA_TRY(cudaMemcpyAsync((T*)dB + k0 * n, B + k0 * n, w * n * sizeof(T), cudaMemcpyHostToDevice, s_h2d));
      LA_CUDA_TRY(cudaMemcpy2DAsync((T*)dA + k0, k * sizeof(T), A + k0, k * sizeof(T), w * sizeof(T), m,
                                    cudaMemcpyHostToDevice, s_h2d));
      LA_CUDA_TRY(cudaEventRecord(ev_b2, s_h2d));
      LA_CUDA_TRY(cudaStreamWaitEvent(st, ev_b2, 0));
      LA_TRY(gemm_dev<T>((const T*)dA + k0, k, (const T*)dB + k0 * n, n, (T*)dC, n, m, w, n,
                         p == 0 ? LA_GEMM_ASSIGN : LA_GEMM_ADD, st));
      k0 += w;
    }
  }
  LA_CUDA_TRY(cudaMemcpyAsync((T*)dB + kh * n, B + kh * n, (k - kh) * n * sizeof(T), cudaMemcpyHostToDevice, s_h2d));
  // Row blocks shrink towards the end (1/4, 1/4, 1/4, 1/8, 1/16, 1/16 of the rows): large blocks fill whole waves of
  // tiles, the small last ones leave little of C to download after the final GEMM.
  size_t blk_rows[16];
  if (kh != 0 && m >= 16 * 128) {
    const size_t q = (m / 4 + 127) / 128 * 128, e = (m / 8 + 127) / 128 * 128, x = (m / 16 + 127) / 128 * 128;
    const size_t want[6] = {q, q, q, e, x, x};
    size_t left = m;
    blocks = 0;
    for (int i = 0; i < 6 && left > 0; ++i) {
      const size_t take = (i == 5 || want[i] > left) ? left : want[i];
      blk_rows[blocks++] = take;
      left -= take;
    }
    if (left > 0) blk_rows[blocks - 1] += left;
  } else {
    for (size_t b = 0; b < blocks; ++b) {
      const size_t r0 = b * rows_blk;
      blk_rows[b] = (m - r0 < rows_blk) ? (m - r0) : rows_blk;
    }
  }
  size_t r0 = 0;
  for (size_t b = 0; b < blocks; r0 += blk_rows[b], ++b) {
    const size_t nr = blk_rows[b];
    if (kh == 0) {
      LA_CUDA_TRY(cudaMemcpyAsync((T*)dA + r0 * k, A + r0 * k, nr * k * sizeof(T), cudaMemcpyHostToDevice, s_h2d));
    } else {
      LA_CUDA_TRY(cudaMemcpy2DAsync((T*)dA + r0 * k + kh, k * sizeof(T), A + r0 * k + kh, k * sizeof(T),
                                    (k - kh) * sizeof(T), nr, cudaMemcpyHostToDevice, s_h2d));
    }
    LA_CUDA_TRY(cudaEventRecord(ev_up[b], s_h2d));
    LA_CUDA_TRY(cudaStreamWaitEvent(st, ev_up[b], 0));
    LA_TRY(gemm_dev<T>((const T*)dA + r0 * k + kh, k, (const T*)dB + kh * n, n, (T*)dC + r0 * n, n, nr, k - kh, n,
                       kh == 0 ? LA_GEMM_ASSIGN : LA_GEMM_ADD, st));
    LA_CUDA_TRY(cudaEventRecord(ev_done[b], st));
    LA_CUDA_TRY(cudaStreamWaitEvent(s_d2h, ev_done[b], 0));
    LA_CUDA_TRY(cudaMemcpyAsync(C + r0 * n, (T*)dC + r0 * n, nr * n * sizeof(T), cudaMemcpyDeviceToHost, s_d2h));
  }
  LA_CUDA_TRY(cudaStreamSynchronize(s_d2h));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}

template <typename T>
int lu_factor_buf(la_buf* LU, size_t m, size_t n, uint64_t* piv_out, int* pospivsign_out) {
  LA_REQUIRE(LU && piv_out && pospivsign_out, "la_lu_factor: null pointer");
  LA_REQUIRE(m > 0 && n > 0, "la_lu_factor: zero dimension");
  size_t bytes;
  LA_REQUIRE(!mul_overflows(m, n, sizeof(T), &bytes) && LU->bytes >= bytes, "la_lu_factor: buffer too small");
  DeviceGuard g;
  LA_TRY(g.enter(LU->device));
  void* meta;
  LA_TRY(scratch_get(LU->device, 3, sizeof(uint64_t) * m + 64, &meta));
  uint64_t* piv_dev = (uint64_t*)meta;
  int* sign_dev = (int*)(piv_dev + m);
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(LU->device, st);
  LA_TRY(lu_factor_dev<T>((T*)LU->ptr, m, n, piv_dev, sign_dev, st));
  LA_CUDA_TRY(cudaMemcpyAsync(piv_out, piv_dev, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaMemcpyAsync(pospivsign_out, sign_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}

template <typename T>
int lu_factor_host(const T* A, T* LU_out, size_t m, size_t n, uint64_t* piv_out, int* pospivsign_out) {
  LA_REQUIRE(A && LU_out && piv_out && pospivsign_out, "la_lu_factor_host: null pointer");
  LA_REQUIRE(m > 0 && n > 0, "la_lu_factor_host: zero dimension");
  size_t bytes;
  LA_REQUIRE(!mul_overflows(m, n, sizeof(T), &bytes), "la_lu_factor_host: size overflow");
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  void *dLU, *meta;
  LA_TRY(scratch_get(ctx->device, 0, bytes, &dLU));
  LA_TRY(scratch_get(ctx->device, 3, sizeof(uint64_t) * m + 64, &meta));
  uint64_t* piv_dev = (uint64_t*)meta;
  int* sign_dev = (int*)(piv_dev + m);
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(ctx->device, st);
  LA_CUDA_TRY(cudaMemcpyAsync(dLU, A, bytes, cudaMemcpyHostToDevice, st));  // == ludata = a.get_data().clone()
  LA_TRY(lu_factor_dev<T>((T*)dLU, m, n, piv_dev, sign_dev, st));
  LA_CUDA_TRY(cudaMemcpyAsync(LU_out, dLU, bytes, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaMemcpyAsync(piv_out, piv_dev, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaMemcpyAsync(pospivsign_out, sign_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}

template <typename T>
int lu_solve_buf(const la_buf* LU, size_t m, size_t n, const uint64_t* piv, const la_buf* B, size_t nx, la_buf* X) {
  LA_REQUIRE(LU && piv && B && X, "la_lu_solve: null pointer");
  LA_REQUIRE(m == n, "la_lu_solve: the factorisation must be square (m=%zu n=%zu); see src/decomp/lu.rs:237-278", m, n);
  LA_REQUIRE(n > 0 && nx > 0, "la_lu_solve: zero dimension");
  size_t bl, bx;
  LA_REQUIRE(!mul_overflows(n, n, sizeof(T), &bl) && !mul_overflows(n, nx, sizeof(T), &bx), "la_lu_solve: size overflow");
  LA_REQUIRE(LU->bytes >= bl && B->bytes >= bx && X->bytes >= bx, "la_lu_solve: buffer too small");
  LA_REQUIRE(LU->device == B->device && LU->device == X->device, "la_lu_solve: buffers live on different devices");
  DeviceGuard g;
  LA_TRY(g.enter(LU->device));
  void* meta;
  LA_TRY(scratch_get(LU->device, 3, sizeof(uint64_t) * n + 64, &meta));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(LU->device, st);
  LA_CUDA_TRY(cudaMemcpyAsync(meta, piv, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
  LA_TRY(lu_solve_dev<T>((const T*)LU->ptr, n, (const uint64_t*)meta, (const T*)B->ptr, nx, (T*)X->ptr, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}

template <typename T>
int lu_solve_host(const T* LU, size_t m, size_t n, const uint64_t* piv, const T* B, size_t nx, T* X) {
  LA_REQUIRE(LU && piv && B && X, "la_lu_solve_host: null pointer");
  LA_REQUIRE(m == n, "la_lu_solve_host: the factorisation must be square (m=%zu n=%zu)", m, n);
  LA_REQUIRE(n > 0 && nx > 0, "la_lu_solve_host: zero dimension");
  size_t bl, bx;
  LA_REQUIRE(!mul_overflows(n, n, sizeof(T), &bl) && !mul_overflows(n, nx, sizeof(T), &bx),
             "la_lu_solve_host: size overflow");
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  void *dLU, *dB, *dX, *meta;
  LA_TRY(scratch_get(ctx->device, 0, bl, &dLU));
  LA_TRY(scratch_get(ctx->device, 1, bx, &dB));
  LA_TRY(scratch_get(ctx->device, 2, bx, &dX));
  LA_TRY(scratch_get(ctx->device, 3, sizeof(uint64_t) * n + 64, &meta));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(ctx->device, st);
  LA_CUDA_TRY(cudaMemcpyAsync(dLU, LU, bl, cudaMemcpyHostToDevice, st));
  LA_CUDA_TRY(cudaMemcpyAsync(dB, B, bx, cudaMemcpyHostToDevice, st));
  LA_CUDA_TRY(cudaMemcpyAsync(meta, piv, sizeof(uint64_t) * n, cudaMemcpyHostToDevice, st));
  LA_TRY(lu_solve_dev<T>((const T*)dLU, n, (const uint64_t*)meta, (const T*)dB, nx, (T*)dX, st));
  LA_CUDA_TRY(cudaMemcpyAsync(X, dX, bx, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}

}  // namespace

namespace {
template <typename T>
int transpose_buf(const la_buf* src, la_buf* dst, size_t rows, size_t cols) {
  size_t bytes;
  LA_REQUIRE(src && dst && rows > 0 && cols > 0, "la_transpose: bad arguments");
  LA_REQUIRE(!mul_overflows(rows, cols, sizeof(T), &bytes) && src->bytes >= bytes && dst->bytes >= bytes,
             "la_transpose: buffer too small for %zu x %zu", rows, cols);
  LA_REQUIRE(src->device == dst->device, "la_transpose: buffers live on different devices");
  DeviceGuard g;
  LA_TRY(g.enter(src->device));
  return transpose_dev<T>((const T*)src->ptr, (T*)dst->ptr, rows, cols, cudaStreamPerThread);
}
template <typename T>
int permute_rows_buf(const la_buf* src, size_t rows, size_t cols, const uint64_t* idx, size_t out_rows, la_buf* dst) {
  size_t sbytes, dbytes;
  LA_REQUIRE(src && dst && idx && rows > 0 && cols > 0 && out_rows > 0, "la_permute_rows: bad arguments");
  LA_REQUIRE(!mul_overflows(rows, cols, sizeof(T), &sbytes) && src->bytes >= sbytes &&
                 !mul_overflows(out_rows, cols, sizeof(T), &dbytes) && dst->bytes >= dbytes,
             "la_permute_rows: buffer too small");
  LA_REQUIRE(src->device == dst->device && src->ptr != dst->ptr, "la_permute_rows: buffers must be distinct, same device");
  for (size_t i = 0; i < out_rows; ++i)
    LA_REQUIRE(idx[i] < rows, "la_permute_rows: row index %llu out of range (rows = %zu)", (unsigned long long)idx[i], rows);
  DeviceGuard g;
  LA_TRY(g.enter(src->device));
  void* meta;
  LA_TRY(scratch_get(src->device, 3, sizeof(uint64_t) * out_rows + 64, &meta));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(src->device, st);
  LA_CUDA_TRY(cudaMemcpyAsync(meta, idx, sizeof(uint64_t) * out_rows, cudaMemcpyHostToDevice, st));
  LA_TRY(permute_rows_dev<T>((const T*)src->ptr, (T*)dst->ptr, (const uint64_t*)meta, out_rows, cols, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));  // `idx` is the caller's (pageable) memory
  return LA_OK;
}
}  // namespace


namespace {
template <typename T>
int chol_factor_buf(la_buf* A, size_t n, int* ok_out) {
  size_t bytes;
  LA_REQUIRE(A && ok_out && n > 0, "la_chol_factor: bad arguments");
  LA_REQUIRE(!mul_overflows(n, n, sizeof(T), &bytes) && A->bytes >= bytes, "la_chol_factor: buffer too small");
  DeviceGuard g;
  LA_TRY(g.enter(A->device));
  void* meta;
  LA_TRY(scratch_get(A->device, 3, 64, &meta));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(A->device, st);
  LA_TRY(chol_factor_dev<T>((T*)A->ptr, n, (int*)meta, st));
  int flags[2] = {0, 0};
  LA_CUDA_TRY(cudaMemcpyAsync(flags, meta, sizeof(flags), cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  *ok_out = (flags[0] == 0 && flags[1] == 0) ? 1 : 0;
  return LA_OK;
}
template <typename T>
int chol_factor_host(const T* A, T* L_out, size_t n, int* ok_out) {
  size_t bytes;
  LA_REQUIRE(A && L_out && ok_out && n > 0, "la_chol_factor_host: bad arguments");
  LA_REQUIRE(!mul_overflows(n, n, sizeof(T), &bytes), "la_chol_factor_host: size overflow");
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  void *dA, *meta;
  LA_TRY(scratch_get(ctx->device, 0, bytes, &dA));
  LA_TRY(scratch_get(ctx->device, 3, 64, &meta));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(ctx->device, st);
  LA_CUDA_TRY(cudaMemcpyAsync(dA, A, bytes, cudaMemcpyHostToDevice, st));
  LA_TRY(chol_factor_dev<T>((T*)dA, n, (int*)meta, st));
  int flags[2] = {0, 0};
  LA_CUDA_TRY(cudaMemcpyAsync(flags, meta, sizeof(flags), cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaMemcpyAsync(L_out, dA, bytes, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  *ok_out = (flags[0] == 0 && flags[1] == 0) ? 1 : 0;
  return LA_OK;
}
template <typename T>
int chol_solve_buf(const la_buf* L, size_t n, const la_buf* B, size_t nx, la_buf* X) {
  size_t bl, bx;
  LA_REQUIRE(L && B && X && n > 0 && nx > 0, "la_chol_solve: bad arguments");
  LA_REQUIRE(!mul_overflows(n, n, sizeof(T), &bl) && L->bytes >= bl && !mul_overflows(n, nx, sizeof(T), &bx) &&
                 B->bytes >= bx && X->bytes >= bx,
             "la_chol_solve: buffer too small");
  LA_REQUIRE(L->device == B->device && L->device == X->device, "la_chol_solve: buffers live on different devices");
  DeviceGuard g;
  LA_TRY(g.enter(L->device));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(L->device, st);
  LA_TRY(chol_solve_dev<T>((const T*)L->ptr, n, (const T*)B->ptr, nx, (T*)X->ptr, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
template <typename T>
int chol_solve_host(const T* L, size_t n, const T* B, size_t nx, T* X) {
  size_t bl, bx;
  LA_REQUIRE(L && B && X && n > 0 && nx > 0, "la_chol_solve_host: bad arguments");
  LA_REQUIRE(!mul_overflows(n, n, sizeof(T), &bl) && !mul_overflows(n, nx, sizeof(T), &bx), "la_chol_solve_host: overflow");
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  void *dL, *dB, *dX;
  LA_TRY(scratch_get(ctx->device, 0, bl, &dL));
  LA_TRY(scratch_get(ctx->device, 1, bx, &dB));
  LA_TRY(scratch_get(ctx->device, 2, bx, &dX));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(ctx->device, st);
  LA_CUDA_TRY(cudaMemcpyAsync(dL, L, bl, cudaMemcpyHostToDevice, st));
  LA_CUDA_TRY(cudaMemcpyAsync(dB, B, bx, cudaMemcpyHostToDevice, st));
  LA_TRY(chol_solve_dev<T>((const T*)dL, n, (const T*)dB, nx, (T*)dX, st));
  LA_CUDA_TRY(cudaMemcpyAsync(X, dX, bx, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
}  // namespace


namespace {
template <typename T>
int qr_sizes(size_t m, size_t n, size_t* bytes_mn, size_t* bytes_rd) {
  LA_REQUIRE(m > 0 && n > 0, "la_qr: zero dimension (m=%zu n=%zu)", m, n);
  LA_REQUIRE(!mul_overflows(m, n, sizeof(T), bytes_mn), "la_qr: size overflow");
  *bytes_rd = (m < n ? m : n) * sizeof(T);
  return LA_OK;
}
template <typename T>
int qr_factor_buf(la_buf* QR, size_t m, size_t n, la_buf* rdiag, la_buf* tmat) {
  size_t bmn = 0, brd = 0, te = 0;
  LA_REQUIRE(QR && rdiag && tmat, "la_qr_factor: null buffer handle");
  LA_TRY(qr_sizes<T>(m, n, &bmn, &brd));
  LA_REQUIRE(QR->device == rdiag->device && QR->device == tmat->device, "la_qr_factor: buffers live on different devices");
  DeviceGuard g;
  LA_TRY(g.enter(QR->device));
  LA_TRY(qr_tmat_elems<T>(m, n, &te));
  LA_REQUIRE(QR->bytes >= bmn && rdiag->bytes >= brd && tmat->bytes >= te * sizeof(T), "la_qr_factor: buffer too small");
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(QR->device, st);
  LA_TRY(qr_factor_dev<T>((T*)QR->ptr, m, n, (T*)rdiag->ptr, (T*)tmat->ptr, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
template <typename T>
int qr_factor_host(const T* A, T* QR_out, T* rdiag_out, size_t m, size_t n) {
  size_t bmn = 0, brd = 0, te = 0;
  LA_REQUIRE(A && QR_out && rdiag_out, "la_qr_factor_host: null pointer");
  LA_TRY(qr_sizes<T>(m, n, &bmn, &brd));
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_TRY(qr_tmat_elems<T>(m, n, &te));
  void *dA, *dR, *dT;
  LA_TRY(scratch_get(ctx->device, 0, bmn, &dA));
  LA_TRY(scratch_get(ctx->device, 3, brd + 64, &dR));
  LA_TRY(scratch_get(ctx->device, 1, te * sizeof(T), &dT));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(ctx->device, st);
  LA_CUDA_TRY(cudaMemcpyAsync(dA, A, bmn, cudaMemcpyHostToDevice, st));
  LA_TRY(qr_factor_dev<T>((T*)dA, m, n, (T*)dR, (T*)dT, st));
  LA_CUDA_TRY(cudaMemcpyAsync(QR_out, dA, bmn, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaMemcpyAsync(rdiag_out, dR, brd, cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
template <typename T>
int qr_get_r_buf(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, la_buf* R) {
  size_t bmn = 0, brd = 0;
  LA_REQUIRE(QR && rdiag && R, "la_qr_get_r: null buffer handle");
  LA_TRY(qr_sizes<T>(m, n, &bmn, &brd));
  LA_REQUIRE(QR->bytes >= bmn && rdiag->bytes >= brd && R->bytes >= bmn, "la_qr_get_r: buffer too small");
  LA_REQUIRE(QR->device == rdiag->device && QR->device == R->device && QR->ptr != R->ptr,
             "la_qr_get_r: buffers must be distinct and on one device");
  DeviceGuard g;
  LA_TRY(g.enter(QR->device));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(QR->device, st);
  return qr_get_r_dev<T>((const T*)QR->ptr, m, n, (const T*)rdiag->ptr, (T*)R->ptr, st);
}
template <typename T>
int qr_get_q_buf(const la_buf* QR, size_t m, size_t n, const la_buf* tmat, la_buf* Q) {
  size_t bmn = 0, brd = 0, bq = 0, te = 0;
  LA_REQUIRE(QR && tmat && Q, "la_qr_get_q: null buffer handle");
  LA_TRY(qr_sizes<T>(m, n, &bmn, &brd));
  LA_REQUIRE(!mul_overflows(m, m, sizeof(T), &bq), "la_qr_get_q: size overflow");
  LA_REQUIRE(QR->device == tmat->device && QR->device == Q->device && QR->ptr != Q->ptr,
             "la_qr_get_q: buffers must be distinct and on one device");
  DeviceGuard g;
  LA_TRY(g.enter(QR->device));
  LA_TRY(qr_tmat_elems<T>(m, n, &te));
  LA_REQUIRE(QR->bytes >= bmn && tmat->bytes >= te * sizeof(T) && Q->bytes >= bq, "la_qr_get_q: buffer too small");
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(QR->device, st);
  LA_TRY(qr_get_q_dev<T>((const T*)QR->ptr, m, n, (const T*)tmat->ptr, (T*)Q->ptr, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
template <typename T>
int qr_solve_buf(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, const la_buf* B, size_t nx, la_buf* X) {
  size_t bmn = 0, brd = 0, bx = 0;
  LA_REQUIRE(QR && rdiag && B && X && nx > 0, "la_qr_solve: bad arguments");
  LA_TRY(qr_sizes<T>(m, n, &bmn, &brd));
  LA_REQUIRE(!mul_overflows(m, nx, sizeof(T), &bx), "la_qr_solve: size overflow");
  LA_REQUIRE(QR->bytes >= bmn && rdiag->bytes >= brd && B->bytes >= bx && X->bytes >= bx, "la_qr_solve: buffer too small");
  LA_REQUIRE(QR->device == rdiag->device && QR->device == B->device && QR->device == X->device,
             "la_qr_solve: buffers live on different devices");
  DeviceGuard g;
  LA_TRY(g.enter(QR->device));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(QR->device, st);
  LA_TRY(qr_solve_dev<T>((const T*)QR->ptr, m, n, (const T*)rdiag->ptr, (const T*)B->ptr, nx, (T*)X->ptr, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
}  // namespace


namespace {
template <typename T>
int elementwise_buf(int op, const la_buf* A, const la_buf* B, T scalar, la_buf* C, size_t count) {
  LA_REQUIRE(A && C && count > 0, "la_elementwise: bad arguments");
  LA_REQUIRE(count <= SIZE_MAX / sizeof(T), "la_elementwise: size overflow");
  const size_t bytes = count * sizeof(T);
  LA_REQUIRE(A->bytes >= bytes && C->bytes >= bytes && (!B || B->bytes >= bytes), "la_elementwise: buffer too small");
  LA_REQUIRE(A->device == C->device && (!B || B->device == A->device), "la_elementwise: buffers live on different devices");
  DeviceGuard g;
  LA_TRY(g.enter(A->device));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(A->device, st);
  return elementwise_dev<T>(op, (const T*)A->ptr, B ? (const T*)B->ptr : nullptr, scalar, (T*)C->ptr, count, st);
}
template <typename T>
int reduce_buf(int kind, const la_buf* A, const la_buf* B, size_t count, T* out) {
  LA_REQUIRE(A && out && count > 0, "la_reduce: bad arguments");
  LA_REQUIRE(count <= SIZE_MAX / sizeof(T), "la_reduce: size overflow");
  const size_t bytes = count * sizeof(T);
  LA_REQUIRE(A->bytes >= bytes && (!B || B->bytes >= bytes), "la_reduce: buffer too small");
  LA_REQUIRE(!B || B->device == A->device, "la_reduce: buffers live on different devices");
  DeviceGuard g;
  LA_TRY(g.enter(A->device));
  cudaStream_t st = cudaStreamPerThread;
  CallScope scope(A->device, st);
  return reduce_dev<T>(kind, (const T*)A->ptr, B ? (const T*)B->ptr : nullptr, count, out, st);
}
}  // namespace

extern "C" {

int la_version(void) { return 1; }
const char* la_last_error(void) { return la::error_text(); }

int la_device_count(int* out) {
  LA_REQUIRE(out, "la_device_count: null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    *out = 0;
    return fail(LA_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
  }
  *out = n;
  return LA_OK;
}
int la_device_sm_count(int device, int* out) {
  LA_REQUIRE(out, "la_device_sm_count: null output");
  const DeviceCtx* ctx;
  LA_TRY(device_ctx(device, &ctx));
  *out = ctx->sm_count;
  return LA_OK;
}
int la_sync(int device) {
  DeviceGuard g;
  LA_TRY(g.enter(device));
  LA_CUDA_TRY(cudaDeviceSynchronize());
  return LA_OK;
}

int la_buf_alloc(size_t bytes, int device, la_buf** out) {
  LA_REQUIRE(out, "la_buf_alloc: null output");
  *out = nullptr;
  LA_REQUIRE(bytes > 0, "la_buf_alloc: zero bytes");
  const DeviceCtx* ctx;
  LA_TRY(device_ctx(device, &ctx));
  DeviceGuard g;
  LA_TRY(g.enter(device));
  void* p = nullptr;
  LA_CUDA_TRY(cudaMalloc(&p, bytes));
  la_buf* b = new la_buf{p, bytes, device};
  *out = b;
  return LA_OK;
}
int la_buf_free(la_buf* buf) {
  if (!buf) return LA_OK;
  DeviceGuard g;
  LA_TRY(g.enter(buf->device));
  cudaError_t e = cudaFree(buf->ptr);
  delete buf;
  if (e != cudaSuccess) return fail(LA_ERR_CUDA, "cudaFree failed: %s", cudaGetErrorString(e));
  return LA_OK;
}
int la_buf_upload(la_buf* dst, size_t off, const void* host, size_t bytes) {
  LA_REQUIRE(dst && host, "la_buf_upload: null pointer");
  LA_REQUIRE(off <= dst->bytes && bytes <= dst->bytes - off, "la_buf_upload: range exceeds the buffer");
  DeviceGuard g;
  LA_TRY(g.enter(dst->device));
  LA_CUDA_TRY(cudaMemcpyAsync((char*)dst->ptr + off, host, bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
  LA_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
  return LA_OK;
}
int la_buf_download(const la_buf* src, size_t off, void* host, size_t bytes) {
  LA_REQUIRE(src && host, "la_buf_download: null pointer");
  LA_REQUIRE(off <= src->bytes && bytes <= src->bytes - off, "la_buf_download: range exceeds the buffer");
  DeviceGuard g;
  LA_TRY(g.enter(src->device));
  LA_CUDA_TRY(cudaMemcpyAsync(host, (const char*)src->ptr + off, bytes, cudaMemcpyDeviceToHost, cudaStreamPerThread));
  LA_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
  return LA_OK;
}
int la_buf_copy(la_buf* dst, const la_buf* src, size_t bytes) {
  LA_REQUIRE(dst && src, "la_buf_copy: null pointer");
  LA_REQUIRE(bytes <= dst->bytes && bytes <= src->bytes, "la_buf_copy: range exceeds a buffer");
  DeviceGuard g;
  LA_TRY(g.enter(dst->device));
  LA_CUDA_TRY(cudaMemcpyAsync(dst->ptr, src->ptr, bytes, cudaMemcpyDefault, cudaStreamPerThread));
  return LA_OK;
}
void* la_buf_device_ptr(const la_buf* buf) { return buf ? buf->ptr : nullptr; }
size_t la_buf_bytes(const la_buf* buf) { return buf ? buf->bytes : 0; }
int la_buf_device(const la_buf* buf) { return buf ? buf->device : -1; }

int la_host_alloc(size_t bytes, void** out) {
  LA_REQUIRE(out && bytes > 0, "la_host_alloc: null output or zero bytes");
  *out = nullptr;
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return LA_OK;
}
int la_host_free(void* ptr) {
  if (!ptr) return LA_OK;
  LA_CUDA_TRY(cudaFreeHost(ptr));
  return LA_OK;
}

int la_gemm_f64(const la_buf* A, const la_buf* B, la_buf* C, size_t m, size_t k, size_t n) {
  return gemm_bufs<double>(A, B, C, m, k, n);
}
int la_gemm_f32(const la_buf* A, const la_buf* B, la_buf* C, size_t m, size_t k, size_t n) {
  return gemm_bufs<float>(A, B, C, m, k, n);
}
int la_gemm_f64_host(const double* A, const double* B, double* C, size_t m, size_t k, size_t n) {
  return gemm_host<double>(A, B, C, m, k, n);
}
int la_gemm_f32_host(const float* A, const float* B, float* C, size_t m, size_t k, size_t n) {
  return gemm_host<float>(A, B, C, m, k, n);
}
int la_gemm_i64_host(const int64_t* A, const int64_t* B, int64_t* C, size_t m, size_t k, size_t n) {
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t must be long long sized");
  return gemm_host<long long>((const long long*)A, (const long long*)B, (long long*)C, m, k, n);
}
int la_gemm_f64_dev(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, void* stream) {
  DEV_SCOPE(stream);
  return gemm_f64_dev(A, lda, B, ldb, C, ldc, m, k, n, mode, resolve_stream(stream));
}
int la_gemm_f32_dev(const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, void* stream) {
  DEV_SCOPE(stream);
  return gemm_f32_dev(A, lda, B, ldb, C, ldc, m, k, n, mode, resolve_stream(stream));
}

int la_lu_factor_f64(la_buf* LU, size_t m, size_t n, uint64_t* piv_out, int* sign_out) {
  return lu_factor_buf<double>(LU, m, n, piv_out, sign_out);
}
int la_lu_factor_f32(la_buf* LU, size_t m, size_t n, uint64_t* piv_out, int* sign_out) {
  return lu_factor_buf<float>(LU, m, n, piv_out, sign_out);
}
int la_lu_factor_f64_host(const double* A, double* LU_out, size_t m, size_t n, uint64_t* piv_out, int* sign_out) {
  return lu_factor_host<double>(A, LU_out, m, n, piv_out, sign_out);
}
int la_lu_factor_f32_host(const float* A, float* LU_out, size_t m, size_t n, uint64_t* piv_out, int* sign_out) {
  return lu_factor_host<float>(A, LU_out, m, n, piv_out, sign_out);
}
int la_lu_factor_f64_dev(double* LU, size_t m, size_t n, uint64_t* piv_dev, int* sign_dev, void* stream) {
  DEV_SCOPE(stream);
  return lu_factor_dev<double>(LU, m, n, piv_dev, sign_dev, resolve_stream(stream));
}
int la_lu_factor_f32_dev(float* LU, size_t m, size_t n, uint64_t* piv_dev, int* sign_dev, void* stream) {
  DEV_SCOPE(stream);
  return lu_factor_dev<float>(LU, m, n, piv_dev, sign_dev, resolve_stream(stream));
}

int la_lu_is_nonsingular_f64(const la_buf* LU, size_t n, int* out) {
  LA_REQUIRE(LU && out && n > 0 && LU->bytes / sizeof(double) / n >= n, "la_lu_is_nonsingular: bad arguments");
  DeviceGuard g;
  LA_TRY(g.enter(LU->device));
  return lu_is_nonsingular_dev<double>((const double*)LU->ptr, n, out, cudaStreamPerThread);
}
int la_lu_is_nonsingular_f32(const la_buf* LU, size_t n, int* out) {
  LA_REQUIRE(LU && out && n > 0 && LU->bytes / sizeof(float) / n >= n, "la_lu_is_nonsingular: bad arguments");
  DeviceGuard g;
  LA_TRY(g.enter(LU->device));
  return lu_is_nonsingular_dev<float>((const float*)LU->ptr, n, out, cudaStreamPerThread);
}
int la_lu_det_f64(const la_buf* LU, size_t n, int pospivsign, double* out) {
  LA_REQUIRE(LU && out && n > 0 && LU->bytes / sizeof(double) / n >= n, "la_lu_det: bad arguments");
  DeviceGuard g;
  LA_TRY(g.enter(LU->device));
  return lu_det_dev<double>((const double*)LU->ptr, n, pospivsign, out, cudaStreamPerThread);
}
int la_lu_det_f32(const la_buf* LU, size_t n, int pospivsign, float* out) {
  LA_REQUIRE(LU && out && n > 0 && LU->bytes / sizeof(float) / n >= n, "la_lu_det: bad arguments");
  DeviceGuard g;
  LA_TRY(g.enter(LU->device));
  return lu_det_dev<float>((const float*)LU->ptr, n, pospivsign, out, cudaStreamPerThread);
}
int la_lu_solve_f64(const la_buf* LU, size_t m, size_t n, const uint64_t* piv, const la_buf* B, size_t nx, la_buf* X) {
  return lu_solve_buf<double>(LU, m, n, piv, B, nx, X);
}
int la_lu_solve_f32(const la_buf* LU, size_t m, size_t n, const uint64_t* piv, const la_buf* B, size_t nx, la_buf* X) {
  return lu_solve_buf<float>(LU, m, n, piv, B, nx, X);
}
int la_lu_solve_f64_host(const double* LU, size_t m, size_t n, const uint64_t* piv, const double* B, size_t nx,
                         double* X) {
  return lu_solve_host<double>(LU, m, n, piv, B, nx, X);
}
int la_lu_solve_f32_host(const float* LU, size_t m, size_t n, const uint64_t* piv, const float* B, size_t nx, float* X) {
  return lu_solve_host<float>(LU, m, n, piv, B, nx, X);
}
int la_lu_solve_f64_dev(const double* LU, size_t n, const uint64_t* piv_dev, const double* B, size_t nx, double* X,
                        void* stream) {
  DEV_SCOPE(stream);
  return lu_solve_dev<double>(LU, n, piv_dev, B, nx, X, resolve_stream(stream));
}
int la_lu_solve_f32_dev(const float* LU, size_t n, const uint64_t* piv_dev, const float* B, size_t nx, float* X,
                        void* stream) {
  DEV_SCOPE(stream);
  return lu_solve_dev<float>(LU, n, piv_dev, B, nx, X, resolve_stream(stream));
}

int la_identity_f64(la_buf* dst, size_t n) {
  LA_REQUIRE(dst && n > 0 && dst->bytes / sizeof(double) / n >= n, "la_identity: bad arguments");
  DeviceGuard g;
  LA_TRY(g.enter(dst->device));
  return identity_dev<double>((double*)dst->ptr, n, cudaStreamPerThread);
}
int la_identity_f32(la_buf* dst, size_t n) {
  LA_REQUIRE(dst && n > 0 && dst->bytes / sizeof(float) / n >= n, "la_identity: bad arguments");
  DeviceGuard g;
  LA_TRY(g.enter(dst->device));
  return identity_dev<float>((float*)dst->ptr, n, cudaStreamPerThread);
}
int la_transpose_f64(const la_buf* src, la_buf* dst, size_t rows, size_t cols) {
  return transpose_buf<double>(src, dst, rows, cols);
}
int la_transpose_f32(const la_buf* src, la_buf* dst, size_t rows, size_t cols) {
  return transpose_buf<float>(src, dst, rows, cols);
}
int la_permute_rows_f64(const la_buf* src, size_t rows, size_t cols, const uint64_t* idx, size_t out_rows, la_buf* dst) {
  return permute_rows_buf<double>(src, rows, cols, idx, out_rows, dst);
}
int la_permute_rows_f32(const la_buf* src, size_t rows, size_t cols, const uint64_t* idx, size_t out_rows, la_buf* dst) {
  return permute_rows_buf<float>(src, rows, cols, idx, out_rows, dst);
}
int la_chol_factor_f64(la_buf* A, size_t n, int* ok_out) { return chol_factor_buf<double>(A, n, ok_out); }
int la_chol_factor_f32(la_buf* A, size_t n, int* ok_out) { return chol_factor_buf<float>(A, n, ok_out); }
int la_chol_factor_f64_host(const double* A, double* L_out, size_t n, int* ok_out) {
  return chol_factor_host<double>(A, L_out, n, ok_out);
}
int la_chol_factor_f32_host(const float* A, float* L_out, size_t n, int* ok_out) {
  return chol_factor_host<float>(A, L_out, n, ok_out);
}
int la_chol_factor_f64_dev(double* A, size_t n, int* flags_dev, void* stream) {
  DEV_SCOPE(stream);
  return chol_factor_dev<double>(A, n, flags_dev, resolve_stream(stream));
}
int la_chol_factor_f32_dev(float* A, size_t n, int* flags_dev, void* stream) {
  DEV_SCOPE(stream);
  return chol_factor_dev<float>(A, n, flags_dev, resolve_stream(stream));
}
int la_chol_solve_f64_dev(const double* L, size_t n, const double* B, size_t nx, double* X, void* stream) {
  DEV_SCOPE(stream);
  return chol_solve_dev<double>(L, n, B, nx, X, resolve_stream(stream));
}
int la_chol_solve_f32_dev(const float* L, size_t n, const float* B, size_t nx, float* X, void* stream) {
  DEV_SCOPE(stream);
  return chol_solve_dev<float>(L, n, B, nx, X, resolve_stream(stream));
}
int la_chol_solve_f64(const la_buf* L, size_t n, const la_buf* B, size_t nx, la_buf* X) {
  return chol_solve_buf<double>(L, n, B, nx, X);
}
int la_chol_solve_f32(const la_buf* L, size_t n, const la_buf* B, size_t nx, la_buf* X) {
  return chol_solve_buf<float>(L, n, B, nx, X);
}
int la_chol_solve_f64_host(const double* L, size_t n, const double* B, size_t nx, double* X) {
  return chol_solve_host<double>(L, n, B, nx, X);
}
int la_chol_solve_f32_host(const float* L, size_t n, const float* B, size_t nx, float* X) {
  return chol_solve_host<float>(L, n, B, nx, X);
}
int la_qr_tmat_elems(size_t m, size_t n, int device, size_t elem_bytes, size_t* elems_out) {
  LA_REQUIRE(elems_out && (elem_bytes == 4 || elem_bytes == 8), "la_qr_tmat_elems: bad arguments");
  DeviceGuard g;
  LA_TRY(g.enter(device));
  return elem_bytes == 8 ? qr_tmat_elems<double>(m, n, elems_out) : qr_tmat_elems<float>(m, n, elems_out);
}
int la_qr_factor_f64(la_buf* QR, size_t m, size_t n, la_buf* rdiag, la_buf* tmat) { return qr_factor_buf<double>(QR, m, n, rdiag, tmat); }
int la_qr_factor_f32(la_buf* QR, size_t m, size_t n, la_buf* rdiag, la_buf* tmat) { return qr_factor_buf<float>(QR, m, n, rdiag, tmat); }
int la_qr_factor_f64_host(const double* A, double* QR_out, double* rdiag_out, size_t m, size_t n) {
  return qr_factor_host<double>(A, QR_out, rdiag_out, m, n);
}
int la_qr_factor_f32_host(const float* A, float* QR_out, float* rdiag_out, size_t m, size_t n) {
  return qr_factor_host<float>(A, QR_out, rdiag_out, m, n);
}
int la_qr_factor_f64_dev(double* QR, size_t m, size_t n, double* rdiag, double* tmat, void* stream) {
  DEV_SCOPE(stream);
  return qr_factor_dev<double>(QR, m, n, rdiag, tmat, resolve_stream(stream));
}
int la_qr_factor_f32_dev(float* QR, size_t m, size_t n, float* rdiag, float* tmat, void* stream) {
  DEV_SCOPE(stream);
  return qr_factor_dev<float>(QR, m, n, rdiag, tmat, resolve_stream(stream));
}
int la_qr_get_r_f64(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, la_buf* R) { return qr_get_r_buf<double>(QR, m, n, rdiag, R); }
int la_qr_get_r_f32(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, la_buf* R) { return qr_get_r_buf<float>(QR, m, n, rdiag, R); }
int la_qr_get_q_f64(const la_buf* QR, size_t m, size_t n, const la_buf* tmat, la_buf* Q) { return qr_get_q_buf<double>(QR, m, n, tmat, Q); }
int la_qr_get_q_f32(const la_buf* QR, size_t m, size_t n, const la_buf* tmat, la_buf* Q) { return qr_get_q_buf<float>(QR, m, n, tmat, Q); }
int la_qr_solve_f64(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, const la_buf* B, size_t nx, la_buf* X) {
  return qr_solve_buf<double>(QR, m, n, rdiag, B, nx, X);
}
int la_qr_solve_f32(const la_buf* QR, size_t m, size_t n, const la_buf* rdiag, const la_buf* B, size_t nx, la_buf* X) {
  return qr_solve_buf<float>(QR, m, n, rdiag, B, nx, X);
}
int la_elementwise_f64(int op, const la_buf* A, const la_buf* B, double scalar, la_buf* C, size_t count) {
  return elementwise_buf<double>(op, A, B, scalar, C, count);
}
int la_elementwise_f32(int op, const la_buf* A, const la_buf* B, float scalar, la_buf* C, size_t count) {
  return elementwise_buf<float>(op, A, B, scalar, C, count);
}
int la_elementwise_f64_dev(int op, const double* A, const double* B, double scalar, double* C, size_t count, void* stream) {
  DEV_SCOPE(stream);
  return elementwise_dev<double>(op, A, B, scalar, C, count, resolve_stream(stream));
}
int la_elementwise_f32_dev(int op, const float* A, const float* B, float scalar, float* C, size_t count, void* stream) {
  DEV_SCOPE(stream);
  return elementwise_dev<float>(op, A, B, scalar, C, count, resolve_stream(stream));
}
int la_reduce_f64(int kind, const la_buf* A, const la_buf* B, size_t count, double* out) { return reduce_buf<double>(kind, A, B, count, out); }
int la_reduce_f32(int kind, const la_buf* A, const la_buf* B, size_t count, float* out) { return reduce_buf<float>(kind, A, B, count, out); }
int la_reduce_f64_dev(int kind, const double* A, const double* B, size_t count, double* out_host, void* stream) {
  DEV_SCOPE(stream);
  return reduce_dev<double>(kind, A, B, count, out_host, resolve_stream(stream));
}
int la_reduce_f32_dev(int kind, const float* A, const float* B, size_t count, float* out_host, void* stream) {
  DEV_SCOPE(stream);
  return reduce_dev<float>(kind, A, B, count, out_host, resolve_stream(stream));
}
int la_fill_hash_f64_dev(double* dst, size_t count, uint64_t seed, uint64_t first_idx, void* stream) {
  return fill_hash_dev<double>(dst, count, seed, first_idx, resolve_stream(stream));
}
int la_fill_hash_f32_dev(float* dst, size_t count, uint64_t seed, uint64_t first_idx, void* stream) {
  return fill_hash_dev<float>(dst, count, seed, first_idx, resolve_stream(stream));
}

int la_set_gemm_f32_mode(int mode) { return la::set_gemm_f32_mode(mode); }
int la_get_gemm_f32_mode(int* out) {
  LA_REQUIRE(out, "la_get_gemm_f32_mode: null output");
  *out = la::get_gemm_f32_mode();
  return LA_OK;
}

/* test hook: 0 = automatic kernel choice, 1 = force the CUDA-core kernel, 2 = force the TMA/DMMA kernel */
int la_debug_set_gemm_path(int path) {
  LA_REQUIRE(path >= 0 && path <= 4, "la_debug_set_gemm_path: bad value %d", path);
  la::debug_set_gemm_path(path);
  return LA_OK;
}

/* test hook: 0 = automatic, 1 = force the CUDA-core fp32 kernel, 2 = force the tcgen05 TF32 kernel */
int la_debug_set_gemm_f32_path(int path) {
  LA_REQUIRE(path >= 0 && path <= 2, "la_debug_set_gemm_f32_path: bad value %d", path);
  la::debug_set_gemm_f32_path(path);
  return LA_OK;
}

}  // extern "C"
