// qr.cu -- QRDecomposition::new / get_q / get_r / solve on a row-major matrix, fp64/fp32, sm_100a.
//
// Replaces (reference, src/decomp/qr.rs):
//   new   :26-106   Householder reflections, column by column.  The k-th vector u = x - a e_k stays UNNORMALISED in
//                   column k from the diagonal down, a = -+|x| (sign opposite to the diagonal entry, :58) goes to
//                   rdiag[k]; H_k = I + u u' / (a u_k) (:94-105); a zero column is skipped (:62).
//   get_r :138-152  strict upper part of qr + rdiag on the diagonal
//   get_q :155-194  the reflections applied in reverse order to the m x m (partial) identity
//   solve :199-238  Y = "Q'" B column by column, then R X = Y by back substitution.  NOTE (parity, not a fix): the
//                   reference applies I - u u'/u_k in the first phase -- JAMA's formula for vectors normalised to
//                   v_k = 1 + x_k/|x|, which these are not -- so its result is not the least-squares solution; and
//                   it builds Matrix::new(cols, nx, <m*nx values>), which panics unless m == n.  The kernels here
//                   reproduce the reference's arithmetic; the mirrors reproduce the panic.
//
// B200 design (blocked Householder, compact WY):
//   panel (nb = 128 columns, all rows below the block row) lives in SHARED MEMORY across a cooperative grid, rows split
//   over the CTAs.  Per column ONE grid-wide reduction of a 128-vector g_j = sum_{r>=k} P[r][k] P[r][j]:
//     j = k  -> |x|^2 (the norm),  j > k -> <x, a_j> from which <u, a_j> = g_j - a P[k][j] (the update's dot products),
//     j < k  -> <u_j, x> from which <u_j, u_k> = g_j - a P[k][j] (column k of the Gram matrix V'V, needed for T).
//   The reduction is a two-hop exchange of flagged ("LL") words through L2 -- partials to the component's owner CTA,
//   totals (plus the diagonal row, which only one CTA holds) back to every CTA -- no grid barrier, no atomics, and a
//   fixed summation order (results are reproducible run to run).
//   T of the compact-WY form H_1 ... H_nb = I - V T V' is inv(S), S upper triangular with S_kk = 1/beta_k = -a u_k and
//   S_ik = <u_i, u_k> (T^-1 + T^-T = V'V): one call of the blocked triangular inversion the LU already uses.
//   Trailing update A <- (I - V T' V') A as three GEMMs on the DMMA kernel: W = V' A (K = rows), W = T' W, A -= V W.
// ---------------------------------------------------------------------------------------------------------------
#include <stdlib.h>

#include <type_traits>

#include "la_common.cuh"
#include "ll_exchange.cuh"

namespace la {
int gemm_f64_splitk(const double* A, size_t lda, const double* B, size_t ldb, double* Cslabs, size_t ldc, size_t slab_elems,
                    size_t m, size_t k, size_t n, int slices, int* slices_out, cudaStream_t st);  // gemm_f64.cu
int gemm_f64_sum_slabs(const double* slabs, size_t slab_elems, int nslabs, double* out, size_t count, cudaStream_t st);
template <typename T>
int tri_block_inverses(const T* M, size_t n, int mode, int first_block, int nblocks, T* W, int trans_out, cudaStream_t st);

namespace {

constexpr int QR_THREADS = 256;
constexpr int QR_NB = 128;                      // block width (also the stride of the stored T' blocks)
constexpr size_t QR_SMEM_BUDGET = 200 * 1024;   // panel rows per CTA * (nb | 1) * sizeof(T)
constexpr int QR_SMEM_EXTRA = 5 * QR_NB * 8 + 64;
constexpr int QR_MAX_PER_LANE = 6;             // exchange words per lane: supports up to 6 * 32 - 1 = 191 CTAs

template <typename T>
using LLW = typename LL<T>::word;

template <typename T>
__host__ __device__ inline size_t qr_ws_words(int G) {
  return (size_t)2 * QR_NB * (G + 1) + (size_t)2 * G * QR_NB * 2;
}

// ---------------------------------------------------------------------------------------------------------------
// panel factorisation of columns [j0, j0 + jb), rows [j0, m)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(QR_THREADS, 1)
qr_panel_kernel(T* __restrict__ A, size_t ld, int m, int j0, int jb, int rpc, LLW<T>* __restrict__ ws, unsigned tag_base,
                T* __restrict__ rdiag, T* __restrict__ S /* [QR_NB][QR_NB]: T^-1 of this block (upper triangle) */) {
  extern __shared__ __align__(16) unsigned char qr_smem[];
  const int G = gridDim.x, c = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stride = jb | 1;  // odd: row walks and column walks are both conflict-free
  T* P = reinterpret_cast<T*>(qr_smem);            // [rpc][stride]
  T* gp = P + (size_t)rpc * stride;                // [2][QR_NB] partial dots of the two row halves
  T* g = gp + 2 * QR_NB;                           // [QR_NB] reduced dots
  T* rowk = g + QR_NB;                             // [QR_NB] row k of the panel (the diagonal row)
  LLW<T>* part = ws;                               // [2][QR_NB][G + 1]
  LLW<T>* tot = ws + (size_t)2 * QR_NB * (G + 1);  // [2][G][QR_NB][2]

  const int row0 = j0 + c * rpc;
  const int nrows = max(0, min(rpc, m - row0));
  for (int idx = tid; idx < nrows * jb; idx += QR_THREADS) {
    const int r = idx / jb, cc = idx - r * jb;
    P[(size_t)r * stride + cc] = A[(size_t)(row0 + r) * ld + j0 + cc];
  }
  // S beyond the block (the inversion works on all QR_NB rows): identity, so that no stale value or 0/0 can leak in
  if (c == 0)
    for (int idx = tid; idx < QR_NB * QR_NB; idx += QR_THREADS) {
      const int i = idx / QR_NB, jj = idx - i * QR_NB;
      if (i >= jb || jj >= jb) S[idx] = (i == jj) ? (T)1 : (T)0;
    }
  __syncthreads();

  const int j = tid & (QR_NB - 1), h = tid >> 7;  // column slot and row half of this thread
  // partial dots of column 0 with every column of the panel; later columns get theirs from the fused update pass
  {
    const int rbeg = max(0, j0 - row0);
    T a0 = (T)0, a1 = (T)0, a2 = (T)0, a3 = (T)0;
    if (j < jb) {
      int r = rbeg + h;
      for (; r + 6 < nrows; r += 8) {
        a0 = fma(P[(size_t)r * stride], P[(size_t)r * stride + j], a0);
        a1 = fma(P[(size_t)(r + 2) * stride], P[(size_t)(r + 2) * stride + j], a1);
        a2 = fma(P[(size_t)(r + 4) * stride], P[(size_t)(r + 4) * stride + j], a2);
        a3 = fma(P[(size_t)(r + 6) * stride], P[(size_t)(r + 6) * stride + j], a3);
      }
      for (; r < nrows; r += 2) a0 = fma(P[(size_t)r * stride], P[(size_t)r * stride + j], a0);
    }
    gp[h * QR_NB + j] = (a0 + a1) + (a2 + a3);
  }
  __syncthreads();
  for (int k = 0; k < jb; ++k) {
    const int lk = j0 + k - row0;   // local index of the diagonal row (negative: above this CTA, >= nrows: below)
    const int rbeg = max(0, lk);    // rows >= the diagonal take part
    // ---- grid-wide sums of the partial dots (gp) + the diagonal row, to every CTA ----
    if (G == 1) {
      if (tid < jb) {
        g[tid] = gp[tid] + gp[QR_NB + tid];
        rowk[tid] = P[(size_t)lk * stride + tid];
      }
    } else {
      const unsigned tag = tag_base + (unsigned)k;
      const int par = k & 1;
      if (tid < jb) {
        LL<T>::store(&part[((size_t)par * QR_NB + tid) * (G + 1) + c], gp[tid] + gp[QR_NB + tid], tag);
        if (lk >= 0 && lk < nrows) LL<T>::store(&part[((size_t)par * QR_NB + tid) * (G + 1) + G], P[(size_t)lk * stride + tid], tag);
      }
      // owners: component jj belongs to CTA jj mod G; one warp per owned component.  All of a lane's words are requested
      // before any is examined (a poll per word would serialise the L2 round trips); fixed summation order.
      for (int jj = c + warp * G; jj < jb; jj += (QR_THREADS / 32) * G) {
        const LLW<T>* src = part + ((size_t)par * QR_NB + jj) * (G + 1);
        T v[QR_MAX_PER_LANE];
        bool ok[QR_MAX_PER_LANE];
#pragma unroll
        for (int u = 0; u < QR_MAX_PER_LANE; ++u) {
          v[u] = (T)0;
          ok[u] = lane + 32 * u > G;  // entries 0..G-1 are the partials, entry G is the diagonal row's value
        }
        bool all;
        do {
          all = true;
#pragma unroll
          for (int u = 0; u < QR_MAX_PER_LANE; ++u)
            if (!ok[u]) {
              ok[u] = LL<T>::load(&src[lane + 32 * u], tag, v[u]);
              all = all && ok[u];
            }
        } while (!all);
        T sum = (T)0, rk = (T)0;
#pragma unroll
        for (int u = 0; u < QR_MAX_PER_LANE; ++u) {
          if (lane + 32 * u < G) sum += v[u];
          if (lane + 32 * u == G) rk = v[u];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, off);
          rk += __shfl_xor_sync(0xffffffffu, rk, off);  // exactly one lane holds a non-zero contribution (or all zero)
        }
        for (int cc = lane; cc < G; cc += 32) {
          LLW<T>* dst = tot + (((size_t)par * G + cc) * QR_NB + jj) * 2;
          LL<T>::store(dst, sum, tag);
          LL<T>::store(dst + 1, rk, tag);
        }
      }
      if (tid < jb) {
        const LLW<T>* src = tot + (((size_t)par * G + c) * QR_NB + tid) * 2;
        T v = (T)0, w = (T)0;
        bool okv = false, okw = false;
        do {
          if (!okv) okv = LL<T>::load(src, tag, v);
          if (!okw) okw = LL<T>::load(src + 1, tag, w);
        } while (!(okv && okw));
        g[tid] = v;
        rowk[tid] = w;
      }
    }
    __syncthreads();
    // ---- the reflection (qr.rs:58-105) ----
    const T xkk = rowk[k];
    const T nrm = sqrt(g[k]);
    const T a = xkk > (T)0 ? -nrm : nrm;
    const T ukk = xkk - a;
    const T den = a * ukk;
    const bool reflect = a != (T)0;  // :62
    if (c == 0) {  // bookkeeping by the first CTA: rdiag, column k of S = T^-1
      if (tid == 0) {
        rdiag[j0 + k] = a;
        S[(size_t)k * QR_NB + k] = reflect ? -den : (T)1;
      }
      if (tid < k) S[(size_t)tid * QR_NB + k] = reflect ? g[tid] - a * rowk[tid] : (T)0;
    }
    if (reflect && tid == 0 && lk >= 0 && lk < nrows) P[(size_t)lk * stride + k] = ukk;  // :77
    const bool has_next = k + 1 < jb;
    // factors <a_j, u> / (a u_k) of this thread's column and of column k + 1 (every thread needs the latter)
    const T f = (reflect && j > k && j < jb) ? (g[j] - a * rowk[j]) / den : (T)0;
    const T fn = (reflect && has_next) ? (g[k + 1] - a * rowk[k + 1]) / den : (T)0;
    __syncthreads();
    // ---- column k+1 first (one row per thread): every thread of the fused pass below reads its UPDATED values ----
    if (reflect && has_next)
      for (int r = rbeg + tid; r < nrows; r += QR_THREADS)
        P[(size_t)r * stride + k + 1] = fma(fn, P[(size_t)r * stride + k], P[(size_t)r * stride + k + 1]);
    __syncthreads();
    // ---- fused pass over the rows >= k: apply the reflection to column j (> k+1) and accumulate column k+1's dot
    //      products with the updated values of every column ----
    {
      T a0 = (T)0, a1 = (T)0;
      if (j < jb) {
        const int kn = has_next ? k + 1 : k;
        const bool upd = reflect && j > k + 1;
        for (int r = rbeg + h; r < nrows; r += 4) {
          const int r2 = r + 2;
          T x0 = P[(size_t)r * stride + j];
          if (upd) {
            x0 = fma(f, P[(size_t)r * stride + k], x0);
            P[(size_t)r * stride + j] = x0;
          }
          if (r > lk) a0 = fma(P[(size_t)r * stride + kn], x0, a0);  // rows below this diagonal = rows >= the next one
          if (r2 < nrows) {
            T x1 = P[(size_t)r2 * stride + j];
            if (upd) {
              x1 = fma(f, P[(size_t)r2 * stride + k], x1);
              P[(size_t)r2 * stride + j] = x1;
            }
            if (r2 > lk) a1 = fma(P[(size_t)r2 * stride + kn], x1, a1);
          }
        }
      }
      gp[h * QR_NB + j] = a0 + a1;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < nrows * jb; idx += QR_THREADS) {
    const int r = idx / jb, cc = idx - r * jb;
    A[(size_t)(row0 + r) * ld + j0 + cc] = P[(size_t)r * stride + cc];
  }
}

// Clean copies of the block's Householder vectors: Vc[(r - j0) * QR_NB + k] and Vt[k * ldvt + (r - j0)] with zeros above
// the diagonal (those positions of the packed matrix hold R).
template <typename T>
__global__ void __launch_bounds__(256) qr_extract_v_kernel(const T* __restrict__ A, size_t ld, int m, int j0, int jb,
                                                           T* __restrict__ Vc, T* __restrict__ Vt, size_t ldvt) {
  __shared__ T tile[32][33];
  const int R = m - j0;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int rb = blockIdx.x * 32, cb = blockIdx.y * 32;
  for (int i = ty; i < 32; i += 8) {
    const int r = rb + i, k = cb + tx;
    T v = (T)0;
    if (r < R && k < jb && r >= k) v = A[(size_t)(j0 + r) * ld + j0 + k];
    tile[i][tx] = v;
    if (r < R && k < QR_NB) Vc[(size_t)r * QR_NB + k] = v;  // columns >= jb are zero padding
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int k = cb + i, r = rb + tx;
    if (k < jb && r < R) Vt[(size_t)k * ldvt + r] = tile[tx][i];
  }
}

template <typename T>
__global__ void qr_get_r_kernel(const T* __restrict__ QR, const T* __restrict__ rdiag, size_t m, size_t n, T* __restrict__ R) {
  const size_t total = m * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / n, jj = e - i * n;
    R[e] = i < jj ? QR[e] : (i == jj ? rdiag[i] : (T)0);  // qr.rs:145-149
  }
}

template <typename T>
__global__ void qr_init_q_kernel(T* __restrict__ Q, size_t m, size_t dc) {
  const size_t total = m * m;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / m, jj = e - i * m;
    Q[e] = (i == jj && i < dc) ? (T)1 : (T)0;  // qr.rs:158-161
  }
}

// The reference's solve (qr.rs:199-238), one CTA per right-hand side: the column of X lives in shared memory, column k of
// the packed matrix is row k of its transpose QRt (contiguous).
template <typename T>
__global__ void __launch_bounds__(256) qr_solve_kernel(const T* __restrict__ QRt /* n x m */, const T* __restrict__ rdiag,
                                                       int m, int n, const T* __restrict__ B, int nx, T* __restrict__ X) {
  extern __shared__ __align__(16) unsigned char qs_smem[];
  T* xs = reinterpret_cast<T*>(qs_smem);  // [m]
  __shared__ T red[8];
  __shared__ T sval;
  const int col = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < m; i += 256) xs[i] = B[(size_t)i * nx + col];
  __syncthreads();
  for (int k = 0; k < n; ++k) {  // :214-224
    const T* u = QRt + (size_t)k * m;
    T s = (T)0;
    for (int i = k + tid; i < m; i += 256) s = fma(u[i], xs[i], s);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      T t = (T)0;
      for (int w = 0; w < 8; ++w) t += red[w];
      sval = -t / u[k];
    }
    __syncthreads();
    const T sv = sval;
    for (int i = k + tid; i < m; i += 256) xs[i] = fma(sv, u[i], xs[i]);
    __syncthreads();
  }
  for (int k = n - 1; k >= 0; --k) {  // :227-236
    const T* u = QRt + (size_t)k * m;
    if (tid == 0) xs[k] = xs[k] / rdiag[k];
    __syncthreads();
    const T xk = xs[k];
    for (int i = tid; i < k; i += 256) xs[i] = xs[i] - xk * u[i];
    __syncthreads();
  }
  for (int i = tid; i < m; i += 256) X[(size_t)i * nx + col] = xs[i];
}

template <typename T>
int qr_block_width(int M, int sms) {
  int rpc = (M + sms - 1) / sms;
  if (rpc < 8) rpc = 8;
  int nb = (int)(QR_SMEM_BUDGET / ((size_t)rpc * sizeof(T))) - 1;
  nb = nb < 0 ? 0 : nb / 16 * 16;
  return nb > QR_NB ? QR_NB : nb;
}

// Scratch shared by the factorisation and get_q: clean V copies, the two W panels, S, the transposed T block.
template <typename T>
struct QrScratch {
  T *Vc, *Vt, *W, *W2, *S, *Tn;
  size_t ldvt, ldw;
  int slab_slot = 42;  // scratch slot of the split-K partial products (the chain stream of the look-ahead has its own)
};
// Per host thread and device: the chain stream (highest priority) of the look-ahead pipeline and its fences.
struct QrSide {
  cudaStream_t sp = nullptr;
  cudaEvent_t e_in = nullptr, e_panel = nullptr, e_bulk = nullptr, e_tail = nullptr;
};
int qr_side(int device, QrSide** out) {
  static thread_local QrSide side[64];
  LA_REQUIRE(device >= 0 && device < 64, "device ordinal out of range");
  QrSide& s = side[device];
  if (!s.sp) {
    int lo = 0, hi = 0;
    LA_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&s.sp, cudaStreamNonBlocking, hi));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_in, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_panel, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_bulk, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_tail, cudaEventDisableTiming));
  }
  *out = &s;
  return LA_OK;
}
template <typename T>
int qr_scratch(int device, size_t rows, size_t cols, QrScratch<T>* out) {
  out->ldvt = (rows + 3) & ~(size_t)3;
  out->ldw = (cols + 3) & ~(size_t)3;
  void* p;
  LA_TRY(scratch_get(device, 31, sizeof(T) * rows * QR_NB, &p));
  out->Vc = (T*)p;
  LA_TRY(scratch_get(device, 32, sizeof(T) * out->ldvt * QR_NB, &p));
  out->Vt = (T*)p;
  LA_TRY(scratch_get(device, 33, sizeof(T) * out->ldw * QR_NB, &p));
  out->W = (T*)p;
  LA_TRY(scratch_get(device, 34, sizeof(T) * out->ldw * QR_NB, &p));
  out->W2 = (T*)p;
  LA_TRY(scratch_get(device, 35, sizeof(T) * 2 * QR_NB * QR_NB, &p));
  out->S = (T*)p;
  out->Tn = out->S + QR_NB * QR_NB;
  {  // split-K partial products: slices * 128 * ldw with slices * ceil(nc / 128) <= SM count, at most 32 slices
    const DeviceCtx* ctx;
    LA_TRY(current_device_ctx(&ctx));
    LA_TRY(scratch_get(device, 42, sizeof(T) * (size_t)(ctx->sm_count + 32) * QR_NB * (QR_NB + 4), &p));
    LA_TRY(scratch_get(device, 43, sizeof(T) * 32 * QR_NB * (QR_NB + 4), &p));
  }
  return LA_OK;
}
// The look-ahead pipeline keeps the clean V copies of two consecutive blocks alive (the bulk stream applies block i while
// the chain stream already builds block i + 1) and gives the chain its own 128-column W panels.
template <typename T>
int qr_scratch_lookahead(int device, size_t rows, QrScratch<T>* odd, QrScratch<T>* chain, const QrScratch<T>& even) {
  *odd = even;
  *chain = even;
  void* p;
  LA_TRY(scratch_get(device, 39, sizeof(T) * rows * QR_NB, &p));
  odd->Vc = (T*)p;
  LA_TRY(scratch_get(device, 40, sizeof(T) * even.ldvt * QR_NB, &p));
  odd->Vt = (T*)p;
  LA_TRY(scratch_get(device, 41, sizeof(T) * 2 * QR_NB * QR_NB, &p));
  chain->W = (T*)p;
  chain->W2 = chain->W + QR_NB * QR_NB;
  chain->ldw = QR_NB;
  chain->slab_slot = 43;
  return LA_OK;
}

// C[rows j0.., cols c0..c0+nc) <- (I - V op(T) V') C for the block whose clean copies sit in sc; Tm is op(T), ld QR_NB.
template <typename T>
int qr_apply_block(const QrScratch<T>& sc, const T* Tm, T* C, size_t ldc, size_t R, size_t jb, size_t nc, cudaStream_t st) {
  const size_t ldw = (nc + 3) & ~(size_t)3;  // compact leading dimension of the two W panels for this call (<= sc.ldw)
  // W = V' C has one tile row (jb <= 128) and a deep contraction over the rows: too few output tiles to fill the machine
  // (nc / 128 CTAs), so the contraction is split over the SMs and the partial products are summed in a fixed order.
  bool done = false;
  if constexpr (std::is_same<T, double>::value) {
    const DeviceCtx* ctx;
    LA_TRY(current_device_ctx(&ctx));
    const size_t tiles = (nc + 127) / 128;
    int slices = (int)((size_t)ctx->sm_count / tiles);
    const int max_by_depth = (int)(R / 512);  // at least 32 k-tiles per slice
    if (slices > max_by_depth) slices = max_by_depth;
    if (slices > 32) slices = 32;
    const bool aligned = ((uintptr_t)C % 16 == 0) && ldc % 2 == 0 && ((uintptr_t)sc.Vt % 16 == 0) && sc.ldvt % 2 == 0 &&
                         ((uintptr_t)sc.W % 16 == 0);
    if (slices >= 2 && aligned) {
      const size_t slab = (size_t)QR_NB * ldw;
      void* p;
      LA_TRY(scratch_get(ctx->device, sc.slab_slot, sizeof(double) * slab * (size_t)slices, &p));  // pre-sized: no growth here
      int used = 1;
      LA_TRY(gemm_f64_splitk(sc.Vt, sc.ldvt, C, ldc, (double*)p, ldw, slab, jb, R, nc, slices, &used, st));
      LA_TRY(gemm_f64_sum_slabs((const double*)p, slab, used, sc.W, jb * ldw, st));
      done = true;
    }
  }
  if (!done) LA_TRY(gemm_dev<T>(sc.Vt, sc.ldvt, C, ldc, sc.W, ldw, jb, R, nc, LA_GEMM_ASSIGN, st));
  LA_TRY(gemm_dev<T>(Tm, QR_NB, sc.W, ldw, sc.W2, ldw, jb, jb, nc, LA_GEMM_ASSIGN, st));
  LA_TRY(gemm_dev<T>(sc.Vc, QR_NB, sc.W2, ldw, C, ldc, R, jb, nc, LA_GEMM_SUB, st));
  return LA_OK;
}

template <typename T>
int qr_extract(const T* QR, size_t ld, int M, int j0, int jb, const QrScratch<T>& sc, cudaStream_t st) {
  const int R = M - j0;
  dim3 grid((unsigned)((R + 31) / 32), QR_NB / 32);
  qr_extract_v_kernel<T><<<grid, 256, 0, st>>>(QR, ld, M, j0, jb, sc.Vc, sc.Vt, sc.ldvt);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}

}  // namespace

template <typename T>
int qr_tmat_elems(size_t m, size_t n, size_t* out) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(m > 0 && n > 0 && m < (1u << 30) && n < (1u << 30), "la_qr: bad dimensions");
  const int nb = qr_block_width<T>((int)m, ctx->sm_count);
  if (nb < 16)
    return fail(LA_ERR_UNSUPPORTED, "la_qr: %zu rows exceed the shared-memory panel capacity of %d SMs", m, ctx->sm_count);
  const size_t dc = m < n ? m : n;
  *out = ((dc + nb - 1) / nb) * (size_t)QR_NB * QR_NB;
  return LA_OK;
}

// In place on QR (m x n, tight rows); rdiag gets min(m, n) values; tmat (la_qr_tmat_elems values) keeps T' of every
// block for get_q.
template <typename T>
int qr_factor_dev(T* QR, size_t m, size_t n, T* rdiag, T* tmat, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(QR && rdiag && tmat, "la_qr_factor: null pointer");
  LA_REQUIRE(m > 0 && n > 0 && m < (1u << 30) && n < (1u << 30), "la_qr_factor: bad dimensions (m=%zu n=%zu)", m, n);
  if (!ctx->coop) return fail(LA_ERR_UNSUPPORTED, "la_qr_factor: device lacks cooperative launch");
  const int M = (int)m, N = (int)n, sms = ctx->sm_count;
  const int dc = M < N ? M : N;
  const int nb = qr_block_width<T>(M, sms);
  if (nb < 16) return fail(LA_ERR_UNSUPPORTED, "la_qr_factor: %d rows exceed the shared-memory panel capacity of %d SMs", M, sms);

  void* wsp;
  const size_t ws_bytes = sizeof(LLW<T>) * qr_ws_words<T>(sms);
  LA_TRY(scratch_get(ctx->device, 30, ws_bytes, &wsp));
  LA_CUDA_TRY(cudaMemsetAsync(wsp, 0, ws_bytes, st));  // all tags invalid; tags then count up within this call
  QrScratch<T> sc;
  LA_TRY(qr_scratch<T>(ctx->device, m, n, &sc));
  LA_CUDA_TRY(cudaFuncSetAttribute(qr_panel_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)QR_SMEM_BUDGET + QR_SMEM_EXTRA));
  unsigned tag = 1;
  // panel of block `blk` on stream s: reflections, S -> T' into tmat; few full CTAs while a bulk update runs beside it
  auto run_panel = [&](int j0, int jb, int blk, bool beside_bulk, cudaStream_t s) -> int {
    const int R = M - j0;
    int rpc = (R + sms - 1) / sms;
    if (beside_bulk) {  // a panel CTA owns its SM (shared memory): fewer, fuller CTAs leave the rest to the DMMA GEMMs
      static const int rpc_div = getenv("LA_QR_RPC_DIV") ? atoi(getenv("LA_QR_RPC_DIV")) : 48;  // tuning knob
      const int fit = (int)(QR_SMEM_BUDGET / ((size_t)(jb | 1) * sizeof(T)));
      int want = R / rpc_div;
      if (want > fit) want = fit;
      if (rpc < want) rpc = want;
    }
    if (rpc < 16) rpc = 16;  // two row halves, a few rows each: short panels run on fewer CTAs
    const int G = (R + rpc - 1) / rpc;
    const size_t smem = (size_t)rpc * (jb | 1) * sizeof(T) + 4 * QR_NB * sizeof(T);
    T* a = QR;
    size_t ld = n;
    int mm = M, jj0 = j0, jjb = jb, rr = rpc;
    LLW<T>* wsw = (LLW<T>*)wsp;
    unsigned tg = tag;
    T* rd = rdiag;
    T* Sp = sc.S;
    void* args[] = {&a, &ld, &mm, &jj0, &jjb, &rr, &wsw, &tg, &rd, &Sp};
    LA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)qr_panel_kernel<T>, dim3(G), dim3(QR_THREADS), args, smem, s));
    tag += QR_NB;
    return tri_block_inverses<T>(sc.S, QR_NB, 1, 0, 1, tmat + (size_t)blk * QR_NB * QR_NB, 1, s);  // T' = inv(S)'
  };
  static const int lookahead = getenv("LA_QR_LOOKAHEAD") ? atoi(getenv("LA_QR_LOOKAHEAD")) : 1;
  const int nblk = (dc + nb - 1) / nb;
  // fp32 block reflectors run on the tcgen05 product, whose operand scratch is per host thread, not per stream: the two
  // streams of the look-ahead would share it -- fp32 stays on the single-stream schedule
  if (!lookahead || nblk < 3 || !std::is_same<T, double>::value) {
    for (int blk = 0; blk < nblk; ++blk) {
      const int j0 = blk * nb, jb = dc - j0 < nb ? dc - j0 : nb, c1 = j0 + jb;
      LA_TRY(run_panel(j0, jb, blk, false, st));
      if (c1 < N) {
        LA_TRY(qr_extract<T>(QR, n, M, j0, jb, sc, st));
        LA_TRY(qr_apply_block<T>(sc, tmat + (size_t)blk * QR_NB * QR_NB, QR + (size_t)j0 * n + c1, n, (size_t)(M - j0),
                                 (size_t)jb, (size_t)(N - c1), st));
      }
    }
    return LA_OK;
  }
  // ---- look-ahead: chain stream sp (high priority): reflector i on the NEXT block's columns -> panel(i+1) -> T', V copies;
  //      bulk stream st: reflector i on everything right of the next block.  Chain step i needs bulk(i-1) (which
  //      updated the next block's columns); bulk(i) needs chain step i-1 (V_i, T_i).  V copies alternate by block parity.
  QrSide* side;
  LA_TRY(qr_side(ctx->device, &side));
  cudaStream_t sp = side->sp;
  QrScratch<T> vbuf[2], chain;
  vbuf[0] = sc;
  LA_TRY(qr_scratch_lookahead<T>(ctx->device, m, &vbuf[1], &chain, sc));
  LA_CUDA_TRY(cudaEventRecord(side->e_in, st));
  LA_CUDA_TRY(cudaStreamWaitEvent(sp, side->e_in, 0));
  LA_TRY(run_panel(0, dc < nb ? dc : nb, 0, false, sp));
  LA_TRY(qr_extract<T>(QR, n, M, 0, dc < nb ? dc : nb, vbuf[0], sp));
  LA_CUDA_TRY(cudaEventRecord(side->e_panel, sp));
  for (int blk = 0; blk < nblk; ++blk) {
    const int j0 = blk * nb, jb = dc - j0 < nb ? dc - j0 : nb, c1 = j0 + jb;
    const bool has_next = blk + 1 < nblk;
    const int jb2 = has_next ? (dc - c1 < nb ? dc - c1 : nb) : 0, c2 = c1 + jb2;
    const QrScratch<T>& vb = vbuf[blk & 1];
    const T* Tt = tmat + (size_t)blk * QR_NB * QR_NB;
    // bulk(i) is queued first: it only needs what chain step i-1 produced
    LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_panel, 0));
    if (c2 < N)
      LA_TRY(qr_apply_block<T>(vb, Tt, QR + (size_t)j0 * n + c2, n, (size_t)(M - j0), (size_t)jb, (size_t)(N - c2), st));
    // chain step i
    if (has_next) {
      if (blk > 0) LA_CUDA_TRY(cudaStreamWaitEvent(sp, side->e_bulk, 0));  // bulk(i-1) updated the next block's columns
      QrScratch<T> cv = vb;  // V_i with the chain's own W panels
      cv.W = chain.W;
      cv.W2 = chain.W2;
      cv.ldw = chain.ldw;
      cv.slab_slot = chain.slab_slot;
      LA_TRY(qr_apply_block<T>(cv, Tt, QR + (size_t)j0 * n + c1, n, (size_t)(M - j0), (size_t)jb, (size_t)jb2, sp));
      LA_TRY(run_panel(c1, jb2, blk + 1, c2 < N, sp));
      LA_TRY(qr_extract<T>(QR, n, M, c1, jb2, vbuf[(blk + 1) & 1], sp));
      LA_CUDA_TRY(cudaEventRecord(side->e_panel, sp));
    }
    LA_CUDA_TRY(cudaEventRecord(side->e_bulk, st));
  }
  LA_CUDA_TRY(cudaEventRecord(side->e_tail, sp));
  LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_tail, 0));
  return LA_OK;
}

template <typename T>
int qr_get_r_dev(const T* QR, size_t m, size_t n, const T* rdiag, T* R, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(QR && rdiag && R && m > 0 && n > 0, "la_qr_get_r: bad arguments");
  size_t blocks = (m * n + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 16;
  qr_get_r_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(QR, rdiag, m, n, R);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}

template <typename T>
int qr_get_q_dev(const T* QR, size_t m, size_t n, const T* tmat, T* Q, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(QR && tmat && Q && m > 0 && n > 0 && m < (1u << 30) && n < (1u << 30), "la_qr_get_q: bad arguments");
  const int M = (int)m, N = (int)n, sms = ctx->sm_count;
  const int dc = M < N ? M : N;
  const int nb = qr_block_width<T>(M, sms);
  if (nb < 16) return fail(LA_ERR_UNSUPPORTED, "la_qr_get_q: %d rows exceed the panel capacity", M);
  QrScratch<T> sc;
  LA_TRY(qr_scratch<T>(ctx->device, m, m, &sc));
  {
    size_t blocks = (m * m + 255) / 256;
    const size_t cap = (size_t)sms * 16;
    qr_init_q_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(Q, m, (size_t)dc);
    LA_CUDA_TRY(cudaGetLastError());
  }
  const int nblk = (dc + nb - 1) / nb;
  for (int blk = nblk - 1; blk >= 0; --blk) {  // Q = H_1 (H_2 (... (H_dc I))), a block of reflections at a time
    const int j0 = blk * nb;
    const int jb = dc - j0 < nb ? dc - j0 : nb;
    LA_TRY(qr_extract<T>(QR, n, M, j0, jb, sc, st));
    LA_TRY(transpose_dev<T>(tmat + (size_t)blk * QR_NB * QR_NB, sc.Tn, QR_NB, QR_NB, st));  // T from the stored T'
    LA_TRY(qr_apply_block<T>(sc, sc.Tn, Q + (size_t)j0 * m + j0, m, (size_t)(M - j0), (size_t)jb, (size_t)(M - j0), st));
  }
  return LA_OK;
}

// X is the reference's full m x nx work array (the caller takes the first n rows when m == n, qr.rs:237).
template <typename T>
int qr_solve_dev(const T* QR, size_t m, size_t n, const T* rdiag, const T* B, size_t nx, T* X, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(QR && rdiag && B && X && m > 0 && n > 0 && nx > 0, "la_qr_solve: bad arguments");
  LA_REQUIRE(n <= m, "la_qr_solve: more columns than rows (the reference's is_full_rank indexes out of bounds, qr.rs:112)");
  LA_REQUIRE(m < (1u << 30) && nx < (1u << 30), "la_qr_solve: dimension too large");
  const size_t smem = sizeof(T) * m;
  if (smem > QR_SMEM_BUDGET)
    return fail(LA_ERR_UNSUPPORTED, "la_qr_solve: %zu rows exceed the shared-memory column capacity", m);
  void* p;
  LA_TRY(scratch_get(ctx->device, 36, sizeof(T) * m * n, &p));
  T* QRt = (T*)p;
  LA_TRY(transpose_dev<T>(QR, QRt, m, n, st));
  LA_CUDA_TRY(cudaFuncSetAttribute(qr_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)QR_SMEM_BUDGET));
  qr_solve_kernel<T><<<(unsigned)nx, 256, smem, st>>>(QRt, rdiag, (int)m, (int)n, B, (int)nx, X);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}

#define LA_QR_INSTANTIATE(T)                                                                            \
  template int qr_tmat_elems<T>(size_t, size_t, size_t*);                                               \
  template int qr_factor_dev<T>(T*, size_t, size_t, T*, T*, cudaStream_t);                              \
  template int qr_get_r_dev<T>(const T*, size_t, size_t, const T*, T*, cudaStream_t);                   \
  template int qr_get_q_dev<T>(const T*, size_t, size_t, const T*, T*, cudaStream_t);                   \
  template int qr_solve_dev<T>(const T*, size_t, size_t, const T*, const T*, size_t, T*, cudaStream_t);
LA_QR_INSTANTIATE(double)
LA_QR_INSTANTIATE(float)

}  // namespace la
