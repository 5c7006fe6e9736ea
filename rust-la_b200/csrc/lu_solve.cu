// lu_solve.cu -- is_non_singular / det / solve on a packed row-major LU factorisation, fp64/fp32, sm_100a.
//
// Replaces (reference, src/decomp/lu.rs):
//   is_non_singular :174-182  exact `== 0` scan of the diagonal
//   det             :224-232  (+1|-1) * prod lu[j][j], multiplied sequentially in index order (overflow / underflow /
//                             -0.0 are part of the result, so the product is NOT tree-reduced)
//   solve           :237-278  X = B(piv,:); forward substitution with unit L (:257-263); backward with U, dividing the
//                             row by the diagonal first (:266-275)
// The solve is HBM-bound for few right-hand sides (8*n^2 bytes of LU for 2*n^2*nx flops): it is blocked so that LU is
// streamed exactly once per sweep in coalesced row segments: for each diagonal block (SB rows) a single-CTA triangular
// solve in shared memory, then a all-SM rank-SB update of the remaining rows.  Per element the updates arrive in the
// reference's order (k ascending in the forward sweep, descending in the backward sweep) with separately rounded
// multiply and subtract, so given the same packed LU and piv the result is BIT-IDENTICAL to the reference's solve.
#include "la_common.cuh"

namespace la {
namespace {

constexpr int SB = 64;        // diagonal block
constexpr int SOLVE_NXT = 16; // RHS columns handled per CTA pass (nx is tiled by this)
constexpr int SOLVE_RG = 256 / SOLVE_NXT;  // row groups per CTA
constexpr int UPD_ROWS = 64;  // rows per CTA in the update kernel

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ B, T* __restrict__ X, const uint64_t* __restrict__ piv, size_t n,
                                   size_t nx) {
  // X[i][:] = B[piv[i]][:]   (lu.rs:246-254)
  const size_t total = n * nx;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t i = idx / nx, j = idx - i * nx;
    X[idx] = B[(size_t)piv[i] * nx + j];
  }
}

// Solve with the diagonal block [r0, r0+sb) for RHS columns [c0, c0+nxt): forward (unit lower) or backward (upper).
template <typename T, bool FORWARD>
__global__ void __launch_bounds__(256) solve_diag_kernel(const T* __restrict__ LU, size_t n, int r0, int sb,
                                                         T* __restrict__ X, size_t nx) {
  __shared__ T D[SB][SB + 1];
  __shared__ T Xs[SB][SOLVE_NXT];
  const int c0 = blockIdx.x * SOLVE_NXT;
  const int nxt = min((int)SOLVE_NXT, (int)nx - c0);
  for (int idx = threadIdx.x; idx < sb * sb; idx += blockDim.x) {
    const int i = idx / sb, k = idx - i * sb;
    D[i][k] = LU[(size_t)(r0 + i) * n + r0 + k];
  }
  for (int idx = threadIdx.x; idx < sb * SOLVE_NXT; idx += blockDim.x) {
    const int i = idx / SOLVE_NXT, j = idx - i * SOLVE_NXT;
    Xs[i][j] = (j < nxt) ? X[(size_t)(r0 + i) * nx + c0 + j] : (T)0;
  }
  __syncthreads();
  const int j = threadIdx.x % SOLVE_NXT;   // RHS column
  const int rg = threadIdx.x / SOLVE_NXT;  // row group
  if (FORWARD) {
    for (int k = 0; k < sb; ++k) {
      const T xk = Xs[k][j];
      for (int i = k + 1 + rg; i < sb; i += SOLVE_RG) Xs[i][j] = sub_rn(Xs[i][j], mul_rn(xk, D[i][k]));
      __syncthreads();
    }
  } else {
    for (int k = sb - 1; k >= 0; --k) {
      if (rg == 0) Xs[k][j] = Xs[k][j] / D[k][k];  // true division by the diagonal (lu.rs:268)
      __syncthreads();
      const T xk = Xs[k][j];
      for (int i = rg; i < k; i += SOLVE_RG) Xs[i][j] = sub_rn(Xs[i][j], mul_rn(xk, D[i][k]));
      __syncthreads();
    }
  }
  for (int idx = threadIdx.x; idx < sb * SOLVE_NXT; idx += blockDim.x) {
    const int i = idx / SOLVE_NXT, jj = idx - i * SOLVE_NXT;
    if (jj < nxt) X[(size_t)(r0 + i) * nx + c0 + jj] = Xs[i][jj];
  }
}

// X[rows][:] -= LU[rows][k0:k0+kb] * X[k0:k0+kb][:] for rows in [row0, row1); streams LU row segments once.
template <typename T, bool FORWARD>
__global__ void __launch_bounds__(256) solve_update_kernel(const T* __restrict__ LU, size_t n, int k0, int kb, int row0,
                                                           int row1, T* __restrict__ X, size_t nx) {
  __shared__ T Ls[UPD_ROWS][SB + 1];
  __shared__ T Xk[SB][SOLVE_NXT];
  const int rbase = row0 + blockIdx.x * UPD_ROWS;
  const int nrows = min(UPD_ROWS, row1 - rbase);
  for (int idx = threadIdx.x; idx < nrows * kb; idx += blockDim.x) {
    const int i = idx / kb, k = idx - i * kb;
    Ls[i][k] = LU[(size_t)(rbase + i) * n + k0 + k];
  }
  const int j = threadIdx.x % SOLVE_NXT;
  const int rg = threadIdx.x / SOLVE_NXT;
  for (int c0 = 0; c0 < (int)nx; c0 += SOLVE_NXT) {
    const int nxt = min((int)SOLVE_NXT, (int)nx - c0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < kb * SOLVE_NXT; idx += blockDim.x) {
      const int k = idx / SOLVE_NXT, jj = idx - k * SOLVE_NXT;
      Xk[k][jj] = (jj < nxt) ? X[(size_t)(k0 + k) * nx + c0 + jj] : (T)0;
    }
    __syncthreads();
    if (j < nxt) {
      for (int i = rg; i < nrows; i += SOLVE_RG) {
        T acc = X[(size_t)(rbase + i) * nx + c0 + j];
        if (FORWARD) {
          for (int k = 0; k < kb; ++k) acc = sub_rn(acc, mul_rn(Xk[k][j], Ls[i][k]));
        } else {
          for (int k = kb - 1; k >= 0; --k) acc = sub_rn(acc, mul_rn(Xk[k][j], Ls[i][k]));
        }
        X[(size_t)(rbase + i) * nx + c0 + j] = acc;
      }
    }
  }
}

template <typename T>
__global__ void nonsingular_kernel(const T* __restrict__ LU, size_t n, int* __restrict__ flag) {
  // flag starts at 1; any exact zero on the diagonal clears it (lu.rs:176-180)
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    if (LU[j * n + j] == (T)0) *flag = 0;
}

template <typename T>
__global__ void diag_gather_kernel(const T* __restrict__ LU, size_t n, T* __restrict__ diag) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x)
    diag[j] = LU[j * n + j];
}
// One warp: lanes prefetch 32 diagonal entries at a time, lane 0's running product is multiplied strictly in index
// order (lu.rs:226-231), so the result -- including inf/0/-0.0 -- is the reference's.
template <typename T>
__global__ void det_kernel(const T* __restrict__ diag, size_t n, int pospivsign, T* __restrict__ out) {
  const int lane = threadIdx.x;
  T d = pospivsign ? (T)1 : -(T)1;
  for (size_t base = 0; base < n; base += 32) {
    const T v = (base + lane < n) ? diag[base + lane] : (T)1;
    const int cnt = (int)min((size_t)32, n - base);
    for (int l = 0; l < cnt; ++l) d = d * __shfl_sync(0xffffffffu, v, l);
  }
  if (lane == 0) *out = d;
}

}  // namespace

template <typename T>
int lu_solve_dev(const T* LU, size_t n, const uint64_t* piv_dev, const T* B, size_t nx, T* X, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(LU && piv_dev && B && X, "la_lu_solve: null pointer");
  LA_REQUIRE(n > 0 && nx > 0, "la_lu_solve: zero dimension");
  LA_REQUIRE(n < (1u << 30) && nx < (1u << 30), "la_lu_solve: dimension too large");
  LA_REQUIRE((const void*)B != (const void*)X, "la_lu_solve: B and X must not alias");
  const int N = (int)n;
  {
    size_t blocks = (n * nx + 255) / 256;
    size_t cap = (size_t)ctx->sm_count * 8;
    gather_rows_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(B, X, piv_dev, n, nx);
    LA_CUDA_TRY(cudaGetLastError());
  }
  const unsigned nxtiles = (unsigned)((nx + SOLVE_NXT - 1) / SOLVE_NXT);
  // forward sweep: L * Y = B(piv,:)
  for (int r0 = 0; r0 < N; r0 += SB) {
    const int sb = (N - r0 < SB) ? (N - r0) : SB;
    solve_diag_kernel<T, true><<<nxtiles, 256, 0, st>>>(LU, n, r0, sb, X, nx);
    const int row0 = r0 + sb;
    if (row0 < N)
      solve_update_kernel<T, true><<<(N - row0 + UPD_ROWS - 1) / UPD_ROWS, 256, 0, st>>>(LU, n, r0, sb, row0, N, X, nx);
  }
  LA_CUDA_TRY(cudaGetLastError());
  // backward sweep: U * X = Y
  for (int r1 = N; r1 > 0; r1 -= SB) {
    const int r0 = (r1 - SB > 0) ? (r1 - SB) : 0;
    const int sb = r1 - r0;
    solve_diag_kernel<T, false><<<nxtiles, 256, 0, st>>>(LU, n, r0, sb, X, nx);
    if (r0 > 0)
      solve_update_kernel<T, false><<<(r0 + UPD_ROWS - 1) / UPD_ROWS, 256, 0, st>>>(LU, n, r0, sb, 0, r0, X, nx);
  }
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template int lu_solve_dev<double>(const double*, size_t, const uint64_t*, const double*, size_t, double*, cudaStream_t);
template int lu_solve_dev<float>(const float*, size_t, const uint64_t*, const float*, size_t, float*, cudaStream_t);

template <typename T>
int lu_is_nonsingular_dev(const T* LU, size_t n, int* out_host, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(LU && out_host && n > 0, "la_lu_is_nonsingular: null pointer or n == 0");
  void* flag = nullptr;
  LA_TRY(scratch_get(ctx->device, 9, 256, &flag));
  const int one = 1;
  LA_CUDA_TRY(cudaMemcpyAsync(flag, &one, sizeof(int), cudaMemcpyHostToDevice, st));
  size_t blocks = (n + 255) / 256;
  nonsingular_kernel<T><<<(unsigned)(blocks < 1024 ? blocks : 1024), 256, 0, st>>>(LU, n, (int*)flag);
  LA_CUDA_TRY(cudaGetLastError());
  LA_CUDA_TRY(cudaMemcpyAsync(out_host, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
template int lu_is_nonsingular_dev<double>(const double*, size_t, int*, cudaStream_t);
template int lu_is_nonsingular_dev<float>(const float*, size_t, int*, cudaStream_t);

template <typename T>
int lu_det_dev(const T* LU, size_t n, int pospivsign, T* out_host, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(LU && out_host && n > 0, "la_lu_det: null pointer or n == 0");
  void* buf = nullptr;
  LA_TRY(scratch_get(ctx->device, 10, sizeof(T) * (n + 1), &buf));
  T* diag = (T*)buf;
  size_t blocks = (n + 255) / 256;
  diag_gather_kernel<T><<<(unsigned)(blocks < 1024 ? blocks : 1024), 256, 0, st>>>(LU, n, diag);
  det_kernel<T><<<1, 32, 0, st>>>(diag, n, pospivsign, diag + n);
  LA_CUDA_TRY(cudaGetLastError());
  LA_CUDA_TRY(cudaMemcpyAsync(out_host, diag + n, sizeof(T), cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}
template int lu_det_dev<double>(const double*, size_t, int, double*, cudaStream_t);
template int lu_det_dev<float>(const float*, size_t, int, float*, cudaStream_t);

}  // namespace la
