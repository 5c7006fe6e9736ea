// cholesky.cu -- blocked right-looking Cholesky factorisation A = L L' and its solve, row-major, fp64/fp32, sm_100a.
//
// Widening step SURVEY.md section 8(f) rank 2.  Replaces `CholeskyDecomposition::new` (reference
// src/decomp/cholesky.rs:56-110) and `solve` (:116-144).  Semantics kept from the reference:
//   * `None` for a matrix that is not symmetric -- exact `!=` on every (k, j) / (j, k) pair, so a NaN pair counts as
//     asymmetric (:91-93) -- or not positive definite: `a[j][j] - sum L[j][k]^2 <= 0` (:99-102; a NaN there is NOT
//     rejected and propagates, exactly as in the reference);
//   * L is lower triangular with explicit zeros above the diagonal (:107-109);
//   * every L[j][k] = (a[j][k] - sum_{i<k} L[k][i] L[j][i]) / L[k][k] with the sum taken i-ascending.
// Matrices of up to 128 rows are factored by ONE kernel that evaluates those expressions in the reference's order with
// separately rounded multiply and add (right-looking with deferred sums, like the exact LU panel): bit-identical L, so
// the reference's `==` tests hold.  Larger matrices: per 128-column block
//   1. chol_diag_kernel  : the same kernel on the (already updated) diagonal block, in shared memory;
//   2. W' = inv(L11)'    : lu.cu's batched triangular inversion (mode 2, transposed output);
//   3. L21 = A21 * W'    : GEMM (DMMA for fp64, CUDA-core fp32 -- TF32 would move the positive-definiteness test);
//   4. A22 -= L21 * L21' : the transposed panel is materialised once (tile transpose), then one GEMM per 2048-column
//                          strip restricted to the rows at or below the strip (the upper triangle is never read);
//                          the strips are independent and alternate over three streams.
// The solve builds L' once, inverts the diagonal blocks of L and L', and runs both sweeps with the LU solve's persistent
// kernels (nx <= 16) or as GEMMs (fp64, n >= 512, even sizes);
// everything else takes a reference-order kernel that is bit-identical to the reference given the same L.
#include <stdlib.h>

#include <type_traits>

#include "la_common.cuh"

namespace la {
int gemm_f64_tensor(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, cudaStream_t st);  // gemm_f64.cu
template <typename T>
int tri_block_inverses(const T* M, size_t n, int mode, int first_block, int nblocks, T* W, int trans_out,
                       cudaStream_t st);  // lu.cu
int tri_sweeps_dev(const double* Lmat, const double* Umat, size_t n, const uint64_t* piv_dev, const double* B, size_t nx,
                   double* X, const double* wl, const double* wu, cudaStream_t st, int wu_mode);  // lu_solve.cu

namespace {
constexpr int CB = 128;  // block size
constexpr int CHOL_THREADS = 256;

// flags[0] = asymmetric pair found, flags[1] = non-positive pivot found
template <typename T>
__global__ void __launch_bounds__(256) chol_sym_kernel(const T* __restrict__ A, size_t n, int* __restrict__ flags) {
  // tile (bi, bj), bj >= bi, against its mirror image: both are read as coalesced rows through shared memory
  __shared__ T t1[32][33];
  __shared__ T t2[32][33];
  const size_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const size_t i = bi * 32 + r, j = bj * 32 + tx;
    t1[r][tx] = (i < n && j < n) ? A[i * n + j] : (T)0;
    const size_t i2 = bj * 32 + r, j2 = bi * 32 + tx;
    t2[r][tx] = (i2 < n && j2 < n) ? A[i2 * n + j2] : (T)0;
  }
  __syncthreads();
  bool bad = false;
  for (int r = ty; r < 32; r += 8) {
    const size_t i = bi * 32 + r, j = bj * 32 + tx;
    if (i < n && j < n && j > i && t1[r][tx] != t2[tx][r]) bad = true;  // cholesky.rs:92 (NaN != NaN is true)
  }
  if (bad) flags[0] = 1;
}

__device__ __forceinline__ int tri(int j, int k) { return j * (j + 1) / 2 + k; }  // packed lower triangle, k <= j

// Factor the jb x jb diagonal block at (j0, j0) in place.  Right-looking with deferred sums: after column i is final,
// s[j][k] += L[k][i] * L[j][i] for every i < k < j -- for a fixed (j, k) the products arrive i-ascending and are rounded
// separately, which is the reference's `s = s + data[k*n+i] * data[j*n+i]` (cholesky.rs:80-82) bit for bit, but the adds
// of different (j, k) are independent (the literal loop is one dependent chain of k additions per element).  The sum of
// squares d[j] likewise grows by L[j][i]^2 as soon as column i exists (cholesky.rs:89, same order).
template <typename T>
__global__ void __launch_bounds__(CHOL_THREADS, 1)
chol_diag_kernel(T* __restrict__ A, size_t ld, int j0, int jb, int* __restrict__ flags) {
  extern __shared__ __align__(16) unsigned char chol_smem[];
  T* Lp = reinterpret_cast<T*>(chol_smem);  // packed lower triangle: a[j][k] on entry, L[j][k] on exit
  T* Sp = Lp + CB * (CB + 1) / 2;           // packed deferred sums
  __shared__ T col[CB];                     // column i of L (dense)
  __shared__ T dsum[CB];
  __shared__ T ldiag[CB];
  const int tid = threadIdx.x;
  for (int e = tid; e < jb * (jb + 1) / 2; e += CHOL_THREADS) Sp[e] = (T)0;
  for (int e = tid; e < jb * jb; e += CHOL_THREADS) {
    const int j = e / jb, k = e - j * jb;
    if (k <= j) Lp[tri(j, k)] = __ldcg(&A[(size_t)(j0 + j) * ld + j0 + k]);
  }
  if (tid < CB) dsum[tid] = (T)0;
  __syncthreads();
  for (int i = 0; i < jb; ++i) {
    // diagonal: L[i][i] = sqrt(a[i][i] - d), None when a[i][i] - d <= 0 (cholesky.rs:99-104) -- evaluated redundantly by
    // every thread of the column phase (same inputs, same result) instead of by one thread between two barriers;
    // column i: L[j][i] = (a[j][i] - s[j][i]) / L[i][i] (cholesky.rs:85), d[j] += L[j][i]^2 (:89)
    const T dii = sub_rn(Lp[tri(i, i)], dsum[i]);
    const T lii = sqrt(dii);
    if (tid == 0) {
      if (dii <= (T)0) flags[1] = 1;
      ldiag[i] = lii;  // kept apart from the packed triangle: a[i][i] there is still being read by the other threads
    }
    for (int j = i + 1 + tid; j < jb; j += CHOL_THREADS) {
      const T v = sub_rn(Lp[tri(j, i)], Sp[tri(j, i)]) / lii;
      Lp[tri(j, i)] = v;
      col[j] = v;
      dsum[j] = add_rn(dsum[j], mul_rn(v, v));
    }
    __syncthreads();
    // deferred sums of the columns still to come: i < k < j.  Threads form a 16 x 16 grid over (k, j); four rows j per
    // trip with loads, arithmetic and stores grouped, so the ~64-cycle FP64 latencies of independent elements overlap
    // (one element at a time, with an integer division for its index, made this kernel 376 us per block).
    {
      const int tx = tid & 15, ty = tid >> 4;
      for (int jb0 = i + 2 + ty; jb0 < jb; jb0 += 64) {  // j = jb0, jb0 + 16, jb0 + 32, jb0 + 48
        T cj[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cj[u] = (jb0 + 16 * u < jb) ? col[jb0 + 16 * u] : (T)0;
        const int kmax = min(jb0 + 48, jb - 1);  // k < j for the largest j of the trip
        for (int k = i + 1 + tx; k < kmax; k += 16) {
          const T ck = col[k];
          T sv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = jb0 + 16 * u;
            sv[u] = (j < jb && k < j) ? Sp[tri(j, k)] : (T)0;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) sv[u] = add_rn(sv[u], mul_rn(ck, cj[u]));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = jb0 + 16 * u;
            if (j < jb && k < j) Sp[tri(j, k)] = sv[u];
          }
        }
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < jb * jb; e += CHOL_THREADS) {
    const int j = e / jb, k = e - j * jb;
    A[(size_t)(j0 + j) * ld + j0 + k] = (k < j) ? Lp[tri(j, k)] : (k == j ? ldiag[j] : (T)0);  // zeros above (:107-109)
  }
}

template <typename T>
__global__ void chol_zero_upper_kernel(T* __restrict__ A, size_t n) {
  const size_t total = n * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / n, j = e - i * n;
    if (j > i) A[e] = (T)0;
  }
}

// Reference-order solve (cholesky.rs:116-144), one thread per right-hand-side column, X holds B on entry: every element
// receives its updates in the reference's order with separately rounded operations -> bit-identical given the same L.
template <typename T>
__global__ void chol_solve_exact_kernel(const T* __restrict__ L, size_t n, T* __restrict__ X, size_t nx) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nx) return;
  for (size_t k = 0; k < n; ++k) {
    T acc = X[k * nx + j];
    for (size_t i = 0; i < k; ++i) acc = sub_rn(acc, mul_rn(X[i * nx + j], L[k * n + i]));
    X[k * nx + j] = acc / L[k * n + k];
  }
  for (size_t k = n; k-- > 0;) {
    T acc = X[k * nx + j];
    for (size_t i = k + 1; i < n; ++i) acc = sub_rn(acc, mul_rn(X[i * nx + j], L[i * n + k]));
    X[k * nx + j] = acc / L[k * n + k];
  }
}

}  // namespace
int transpose_lower_f64_dev(const double* src, double* dst, size_t n, cudaStream_t st);  // la_runtime.cu
int gemm_f64_sub_lower(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                       cudaStream_t st);  // gemm_f64.cu
namespace {
template <typename T>
int gemm_exact(const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc, size_t m, size_t k, size_t n, int mode,
               cudaStream_t st) {
  if constexpr (std::is_same<T, double>::value) {
    return gemm_dev<double>(A, lda, B, ldb, C, ldc, m, k, n, mode, st);  // DMMA when TMA-addressable, CUDA cores otherwise
  } else {
    return gemm_simt<T>(A, lda, B, ldb, C, ldc, m, k, n, mode, st);  // exact fp32: no TF32 in a factorisation
  }
}
}  // namespace

// In place: on success the lower triangle of A holds L and the upper triangle zeros; flags_dev[0] / [1] != 0 afterwards
// mean "not symmetric" / "not positive definite" (the reference returns None for either).
template <typename T>
int chol_factor_dev(T* A, size_t n, int* flags_dev, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(A && flags_dev, "la_chol_factor: null pointer");
  LA_REQUIRE(n > 0 && n < (1u << 30), "la_chol_factor: bad dimension %zu", n);
  const int N = (int)n;
  LA_CUDA_TRY(cudaMemsetAsync(flags_dev, 0, 2 * sizeof(int), st));
  {
    const unsigned tiles = (unsigned)((n + 31) / 32);
    LA_REQUIRE(tiles <= 65535, "la_chol_factor: dimension too large for the symmetry check grid");
    chol_sym_kernel<T><<<dim3(tiles, tiles), 256, 0, st>>>(A, n, flags_dev);
    LA_CUDA_TRY(cudaGetLastError());
  }
  const int DIAG_SMEM = (int)(sizeof(T) * 2 * (CB * (CB + 1) / 2));
  LA_CUDA_TRY(cudaFuncSetAttribute(chol_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
  if (N <= CB) {
    chol_diag_kernel<T><<<1, CHOL_THREADS, DIAG_SMEM, st>>>(A, n, 0, N, flags_dev);
    LA_CUDA_TRY(cudaGetLastError());
    return LA_OK;
  }
  void *wbuf = nullptr, *sbuf = nullptr, *tbuf = nullptr;
  LA_TRY(scratch_get(ctx->device, 16, sizeof(T) * 2 * CB * CB, &wbuf));
  LA_TRY(scratch_get(ctx->device, 17, sizeof(T) * 2 * n * CB, &sbuf));
  LA_TRY(scratch_get(ctx->device, 18, sizeof(T) * 2 * n * CB, &tbuf));
  // double-buffered by block parity: the bulk update of block i reads S/ST[i & 1] while the chain builds block i+1's
  T* WT[2] = {(T*)wbuf, (T*)wbuf + CB * CB};       // inv(L11)', zero-padded to 128 x 128
  T* S[2] = {(T*)sbuf, (T*)sbuf + n * CB};         // L21, R x 128 (leading dimension 128)
  T* ST[2] = {(T*)tbuf, (T*)tbuf + n * CB};        // L21', 128 x R (leading dimension R)
  constexpr int STRIP = 2048;
  // Look-ahead, as in lu.cu: a high-priority CHAIN stream factors the next block column (its trailing update, diagonal
  // block, inv(L11)', panel GEMM, copy back, transpose) while the BULK streams run the rest of the current trailing
  // update.  The strips of one trailing update are independent and alternate over three streams (the half-empty last wave
  // of one GEMM overlaps the next).
  struct Side {  // per host thread and device
    cudaStream_t chain = nullptr, side[2] = {nullptr, nullptr};
    cudaEvent_t ev_in = nullptr, ev_fork = nullptr, ev_join[2] = {nullptr, nullptr}, ev_chain = nullptr, ev_bulk = nullptr;
  };
  static thread_local Side sides[64];
  LA_REQUIRE(ctx->device >= 0 && ctx->device < 64, "device ordinal out of range");
  Side& sd = sides[ctx->device];
  if (!sd.chain) {
    int lo = 0, hi = 0;
    LA_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&sd.chain, cudaStreamNonBlocking, hi));
    for (int i = 0; i < 2; ++i) {
      LA_CUDA_TRY(cudaStreamCreateWithFlags(&sd.side[i], cudaStreamNonBlocking));
      LA_CUDA_TRY(cudaEventCreateWithFlags(&sd.ev_join[i], cudaEventDisableTiming));
    }
    LA_CUDA_TRY(cudaEventCreateWithFlags(&sd.ev_in, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&sd.ev_fork, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&sd.ev_chain, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&sd.ev_bulk, cudaEventDisableTiming));
  }
  cudaStream_t chain = sd.chain;
  cudaStream_t* side = sd.side;
  cudaEvent_t ev_in = sd.ev_in, ev_fork = sd.ev_fork, ev_chain = sd.ev_chain, ev_bulk = sd.ev_bulk;
  cudaEvent_t* ev_join = sd.ev_join;
  // factor block column `blk` on stream s: diagonal block, then (if rows remain) L21 into S/ST[blk & 1] and back into A
  auto factor_block = [&](int blk, cudaStream_t s) -> int {
    const int j0 = blk * CB;
    const int jb = (N - j0 < CB) ? (N - j0) : CB;
    chol_diag_kernel<T><<<1, CHOL_THREADS, DIAG_SMEM, s>>>(A, n, j0, jb, flags_dev);
    LA_CUDA_TRY(cudaGetLastError());
    const int c1 = j0 + jb;
    if (c1 >= N) return LA_OK;
    const size_t R = (size_t)(N - c1);
    const int p = blk & 1;
    LA_TRY(tri_block_inverses<T>(A, n, 2, blk, 1, WT[p], 1, s));
    // L21 = A21 * inv(L11)'  (out of place: every CTA of a tile row reads the whole row of A21)
    LA_TRY(gemm_exact<T>(A + (size_t)c1 * n + j0, n, WT[p], CB, S[p], CB, R, (size_t)jb, (size_t)CB, LA_GEMM_ASSIGN, s));
    LA_CUDA_TRY(cudaMemcpy2DAsync(A + (size_t)c1 * n + j0, n * sizeof(T), S[p], CB * sizeof(T), jb * sizeof(T), R,
                                  cudaMemcpyDeviceToDevice, s));
    LA_TRY(transpose_dev<T>(S[p], ST[p], R, CB, s));
    return LA_OK;
  };
  const int nblk = (N + CB - 1) / CB;
  static const int lower_knob = getenv("LA_CHOL_LOWER") ? atoi(getenv("LA_CHOL_LOWER")) : 1;  // 0: 2048-column strips
  const bool lower_gemm = std::is_same<T, double>::value && lower_knob != 0 && N % 2 == 0 && (uintptr_t)A % 16 == 0;
  LA_CUDA_TRY(cudaEventRecord(ev_in, st));
  LA_CUDA_TRY(cudaStreamWaitEvent(chain, ev_in, 0));
  LA_TRY(factor_block(0, chain));
  for (int blk = 0; blk + 1 < nblk; ++blk) {
    const int p = blk & 1;
    const int c1 = (blk + 1) * CB;          // first row/column of the trailing matrix
    const size_t R = (size_t)(N - c1);
    const size_t hw = (R < (size_t)CB) ? R : (size_t)CB;  // width of the next block column
    LA_CUDA_TRY(cudaEventRecord(ev_chain, chain));  // S/ST[p] are ready
    // ---- chain: the next block column first, then its factorisation ----
    if (blk > 0) LA_CUDA_TRY(cudaStreamWaitEvent(chain, ev_bulk, 0));  // bulk(blk-1) updated it and is done with S/ST[1-p]
    LA_TRY(gemm_exact<T>(S[p], CB, ST[p], R, A + (size_t)c1 * n + c1, n, R, (size_t)CB, hw, LA_GEMM_SUB, chain));
    LA_TRY(factor_block(blk + 1, chain));
    // ---- bulk: the trailing update right of the next block column, strip by strip (rows at or below the strip) ----
    LA_CUDA_TRY(cudaStreamWaitEvent(st, ev_chain, 0));
    if (lower_gemm && R > hw) {
      // one launch: only the tiles that touch or lie below the diagonal are computed (the strips below compute a
      // 2048-wide staircase: 2048 * 1.5 / n = 19 % more flops than the triangle at n = 16384)
      LA_TRY(gemm_f64_sub_lower((const double*)(S[p] + hw * CB), CB, (const double*)(ST[p] + hw), R,
                                (double*)(A + (c1 + hw) * n + (c1 + hw)), n, R - hw, (size_t)CB, st));
      LA_CUDA_TRY(cudaEventRecord(ev_bulk, st));
      continue;
    }
    LA_CUDA_TRY(cudaEventRecord(ev_fork, st));
    int used = 0, si = 0;
    for (size_t s0 = hw; s0 < R; s0 += STRIP, ++si) {
      const size_t w = (R - s0 < STRIP) ? (R - s0) : STRIP;
      cudaStream_t s = (si % 3 == 0) ? st : side[si % 3 - 1];
      if (s != st && !(used & (1 << (si % 3)))) {
        LA_CUDA_TRY(cudaStreamWaitEvent(s, ev_fork, 0));
        used |= 1 << (si % 3);
      }
      LA_TRY(gemm_exact<T>(S[p] + s0 * CB, CB, ST[p] + s0, R, A + (c1 + s0) * n + (c1 + s0), n, R - s0, (size_t)CB, w,
                           LA_GEMM_SUB, s));
    }
    for (int i = 0; i < 2; ++i)
      if (used & (1 << (i + 1))) {
        LA_CUDA_TRY(cudaEventRecord(ev_join[i], side[i]));
        LA_CUDA_TRY(cudaStreamWaitEvent(st, ev_join[i], 0));
      }
    LA_CUDA_TRY(cudaEventRecord(ev_bulk, st));
  }
  LA_CUDA_TRY(cudaEventRecord(ev_chain, chain));
  LA_CUDA_TRY(cudaStreamWaitEvent(st, ev_chain, 0));
  {
    size_t blocks = (n * n + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    chol_zero_upper_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(A, n);  // the strips left stale values above the diagonal
    LA_CUDA_TRY(cudaGetLastError());
  }
  return LA_OK;
}
template int chol_factor_dev<double>(double*, size_t, int*, cudaStream_t);
template int chol_factor_dev<float>(float*, size_t, int*, cudaStream_t);

// X = A^-1 B given L (A = L L'); B and X must not alias.
template <typename T>
int chol_solve_dev(const T* L, size_t n, const T* B, size_t nx, T* X, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(L && B && X, "la_chol_solve: null pointer");
  LA_REQUIRE(n > 0 && nx > 0 && n < (1u << 30) && nx < (1u << 30), "la_chol_solve: bad dimension");
  LA_REQUIRE((const void*)B != (const void*)X, "la_chol_solve: B and X must not alias");
  LA_CUDA_TRY(cudaMemcpyAsync(X, B, n * nx * sizeof(T), cudaMemcpyDeviceToDevice, st));  // xdata = b.get_data().clone()
  if constexpr (std::is_same<T, double>::value) {
    if (n >= 4 * CB && n % 2 == 0 && nx % 2 == 0 && (uintptr_t)L % 16 == 0 && (uintptr_t)X % 16 == 0) {
      // L' once, inverted diagonal blocks of L (lower, non-unit) and of L' (upper), then GEMM sweeps as in lu_solve.cu
      const int G = (int)((n + CB - 1) / CB);
      void *ltbuf = nullptr, *wbuf = nullptr;
      LA_TRY(scratch_get(ctx->device, 19, sizeof(double) * n * n, &ltbuf));
      LA_TRY(scratch_get(ctx->device, 15, sizeof(double) * 2 * (size_t)G * CB * CB, &wbuf));
      double* LT = (double*)ltbuf;
      double* WL = (double*)wbuf;
      double* WU = WL + (size_t)G * CB * CB;
      // only the upper triangle of L' is ever read (off-diagonal blocks above the diagonal, upper-triangular diagonal blocks)
      LA_TRY(transpose_lower_f64_dev(L, LT, n, st));
      LA_TRY(tri_block_inverses<double>(L, n, 2, 0, G, WL, 0, st));
      if (nx <= 16 && ctx->coop)  // few right-hand sides: the LU solve's persistent sweep kernels (they invert the blocks
        return tri_sweeps_dev(L, LT, n, nullptr, B, nx, X, WL, WU, st, 1);  // of L' beside the forward sweep)
      LA_TRY(tri_block_inverses<double>(LT, n, 1, 0, G, WU, 0, st));
      for (int b = 0; b < G; ++b) {  // L Y = B
        const size_t r0 = (size_t)b * CB, nr = (n - r0 < (size_t)CB) ? (n - r0) : (size_t)CB;
        double* Xb = X + r0 * nx;
        LA_TRY(gemm_f64_tensor(WL + (size_t)b * CB * CB, CB, Xb, nx, Xb, nx, nr, nr, nx, LA_GEMM_ASSIGN, st));
        if (r0 + nr < n)
          LA_TRY(gemm_f64_tensor(L + (r0 + nr) * n + r0, n, Xb, nx, X + (r0 + nr) * nx, nx, n - r0 - nr, nr, nx,
                                 LA_GEMM_SUB, st));
      }
      for (int b = G - 1; b >= 0; --b) {  // L' X = Y
        const size_t r0 = (size_t)b * CB, nr = (n - r0 < (size_t)CB) ? (n - r0) : (size_t)CB;
        double* Xb = X + r0 * nx;
        LA_TRY(gemm_f64_tensor(WU + (size_t)b * CB * CB, CB, Xb, nx, Xb, nx, nr, nr, nx, LA_GEMM_ASSIGN, st));
        if (r0 > 0) LA_TRY(gemm_f64_tensor(LT + r0, n, Xb, nx, X, nx, r0, nr, nx, LA_GEMM_SUB, st));
      }
      return LA_OK;
    }
  }
  chol_solve_exact_kernel<T><<<(unsigned)((nx + 63) / 64), 64, 0, st>>>(L, n, X, nx);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template int chol_solve_dev<double>(const double*, size_t, const double*, size_t, double*, cudaStream_t);
template int chol_solve_dev<float>(const float*, size_t, const float*, size_t, float*, cudaStream_t);

}  // namespace la
