// lu.cu -- blocked right-looking LU with partial (row) pivoting, row-major, fp64/fp32, for sm_100a.
//
// Replaces `LUDecomposition::new` (reference src/decomp/lu.rs:104-168), `is_non_singular` (:174-182), `det` (:224-232)
// and `solve` (:237-278).  Semantics kept from the reference:
//   * A(piv,:) = L*U, packed storage, unit diagonal of L implicit (:96-98); piv[i] = original row now in row i;
//   * pivot = largest |x| at/below the diagonal, strict '>' so the LOWEST row index wins ties and a NaN candidate never
//     displaces the incumbent (a NaN incumbent is never displaced either) (:132-137);
//   * the whole row is swapped, including the already-final L columns (:140-146);
//   * L entries by true division; a pivot that is exactly 0 skips the division and the factorisation continues (:156-160);
//   * all n columns are processed also when m < n (:116): columns >= m only receive the U update (TRSM below).
//
// Algorithm per panel of `jb` columns starting at j0 (everything on one stream, no host round trip):
//   1. lu_panel_kernel  (cooperative, one CTA per SM): the (m-j0) x jb panel is distributed row-wise over the CTAs and
//      kept in SHARED MEMORY for the whole panel; per column: local arg-max -> candidates published to global ->
//      one grid barrier -> every CTA redundantly picks the winner, swaps, scales by the pivot and rank-1 updates its
//      rows while tracking the next column's arg-max.  Row-major makes the pivot search a strided column walk in the
//      reference; here it is a register/shared-memory reduction.
//   2. lu_perm_kernel   (one warp): folds the jb sequential interchanges into a net permutation (<= 2*jb moved rows),
//      and applies them to `piv` / the sign.
//   3. lu_swap_kernel   : applies the net permutation to the columns left of the panel (coalesced row segments).
//   4. lu_swap_trsm_kernel : applies it to the columns right of the panel and solves U12 = L11^-1 * A12 in shared memory.
//   5. trailing update A22 -= L21 * U12 on the DMMA GEMM (gemm_f64.cu, mode LA_GEMM_SUB).
#include <cooperative_groups.h>
#include <float.h>
#include <limits.h>
#include <stdlib.h>

#include <type_traits>

#include "la_common.cuh"

namespace la {
namespace {

constexpr int PANEL_THREADS = 256;
constexpr int PANEL_WARPS = PANEL_THREADS / 32;
constexpr int MAX_NB = 128;
constexpr int MAX_MOVES = 2 * MAX_NB;
constexpr size_t PANEL_SMEM_BUDGET = 200 * 1024;

// Global workspace of one factorisation (per stream use; lives in the scratch pool).
// Exchange rows carry 2*MAX_NB values: the row itself and (EXACT mode) its deferred-subtraction sums.
constexpr int XROW = 2 * MAX_NB;
template <typename T>
struct PanelWs {
  unsigned int barrier;  // grid barrier counter, zeroed before every panel launch
  int n_moves;
  int pad[2];
  int ipiv[MAX_NB];             // absolute pivot row chosen for column j0 + c
  int move_dst[MAX_MOVES];      // net permutation of the panel: row move_dst[i] receives old row move_src[i]
  int move_src[MAX_MOVES];
  // double-buffered per-step exchange area, laid out after the struct:
  //   double cand_key[2][G]; int cand_idx[2][G]; T cand_row[2][G][XROW]; T top_row[2][XROW];
};

template <typename T>
__host__ __device__ inline size_t ws_bytes(int G) {
  size_t b = sizeof(PanelWs<T>);
  b += sizeof(double) * 2 * G;
  b += sizeof(int) * 2 * G;
  b = (b + 15) & ~(size_t)15;
  b += sizeof(T) * 2 * (size_t)G * XROW;
  b += sizeof(T) * 2 * XROW;
  return b;
}
template <typename T>
struct WsView {
  PanelWs<T>* hdr;
  double* cand_key;  // [2][G]
  int* cand_idx;     // [2][G]
  T* cand_row;       // [2][G][XROW]
  T* top_row;        // [2][XROW]
};
template <typename T>
__host__ __device__ inline WsView<T> ws_view(void* base, int G) {
  WsView<T> v;
  char* p = (char*)base;
  v.hdr = (PanelWs<T>*)p;
  p += sizeof(PanelWs<T>);
  v.cand_key = (double*)p;
  p += sizeof(double) * 2 * G;
  v.cand_idx = (int*)p;
  p += sizeof(int) * 2 * G;
  p = (char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  v.cand_row = (T*)p;
  p += sizeof(T) * 2 * (size_t)G * XROW;
  v.top_row = (T*)p;
  return v;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Pivot key: |x|, with the reference's NaN behaviour folded in (lu.rs:132-137).
//   NaN in a candidate row  -> -1 (never wins: `abs(x) > abs(cur)` is false)
//   NaN in the incumbent (the diagonal row) -> +inf with the lowest index (never displaced)
template <typename T>
__device__ __forceinline__ double pivot_key(T v, bool is_diag_row) {
  double a = fabs((double)v);
  if (a != a) return is_diag_row ? (double)INFINITY : -1.0;
  return a;
}
// strict '>' with lowest-index tie break == first maximum of a sequential scan
__device__ __forceinline__ void key_merge(double& k, int& i, double k2, int i2) {
  if (k2 > k || (k2 == k && i2 < i)) {
    k = k2;
    i = i2;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 1. panel factorisation
//
// EXACT = true (used when the whole factorisation is one panel, i.e. min(m,n) <= nb): every element keeps its original
// value a and a separate running sum s = sum_k l[i][k]*u[k][j] (k ascending, product and sum rounded separately); the
// value a - s is formed once.  That is the reference's left-looking expression (lu.rs:122-129) evaluated in
// right-looking order, so the packed LU is BIT-IDENTICAL to the reference's -- in particular exact zeros (singular
// matrices) are reproduced.  EXACT = false subtracts each separately rounded product immediately (half the shared
// memory); later panels carry DMMA-rounded trailing updates anyway.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool EXACT>
__global__ void __launch_bounds__(PANEL_THREADS, 1)
lu_panel_kernel(T* __restrict__ A, size_t ld, int m, int j0, int jb, int rows_per_cta, void* ws_base) {
  extern __shared__ __align__(16) unsigned char panel_smem[];
  T* rows = reinterpret_cast<T*>(panel_smem);            // [rows_per_cta][jb]: original values, then final L / U
  T* sums = rows + (size_t)rows_per_cta * jb;            // [rows_per_cta][jb]: deferred sums (EXACT only)
  __shared__ double wkey[PANEL_WARPS];
  __shared__ int widx[PANEL_WARPS];

  const int G = gridDim.x;
  const WsView<T> ws = ws_view<T>(ws_base, G);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_base = j0 + blockIdx.x * rows_per_cta;          // first absolute row of this CTA
  const int nloc = max(0, min(rows_per_cta, m - row_base));     // rows held by this CTA
  const int nq = (jb + 31) >> 5;                                // column groups of 32 per lane

  auto value = [&](int lr, int col) -> T {
    const T a = rows[(size_t)lr * jb + col];
    return EXACT ? sub_rn(a, sums[(size_t)lr * jb + col]) : a;
  };

  // ---- load the CTA's rows of the panel into shared memory (row segments of jb contiguous elements) ----
  for (int idx = threadIdx.x; idx < nloc * jb; idx += PANEL_THREADS) {
    const int lr = idx / jb, c = idx - lr * jb;
    rows[idx] = __ldcg(&A[(size_t)(row_base + lr) * ld + j0 + c]);
    if (EXACT) sums[idx] = (T)0;
  }
  __syncthreads();

  // ---- initial per-warp arg-max of column 0 ----
  {
    double k = -2.0;
    int ki = INT_MAX;
    for (int lr = warp; lr < nloc; lr += PANEL_WARPS) {
      const int gr = row_base + lr;
      if (lane == 0) key_merge(k, ki, pivot_key(rows[(size_t)lr * jb], gr == j0), gr);
    }
    if (lane == 0) {
      wkey[warp] = k;
      widx[warp] = ki;
    }
  }

  unsigned int bar_target = 0;
  for (int c = 0; c < jb; ++c) {
    const int par = c & 1;
    const int diag = j0 + c;  // absolute row/col index of this step's diagonal
    __syncthreads();          // (A) all rows updated, wkey/widx written
    if (warp == 0) {
      double k = (lane < PANEL_WARPS) ? wkey[lane] : -2.0;
      int ki = (lane < PANEL_WARPS) ? widx[lane] : INT_MAX;
#pragma unroll
      for (int off = 4; off > 0; off >>= 1) {
        double k2 = __shfl_down_sync(0xffffffffu, k, off);
        int i2 = __shfl_down_sync(0xffffffffu, ki, off);
        key_merge(k, ki, k2, i2);
      }
      k = __shfl_sync(0xffffffffu, k, 0);
      ki = __shfl_sync(0xffffffffu, ki, 0);
      if (lane == 0) {
        ws.cand_key[par * G + blockIdx.x] = k;
        ws.cand_idx[par * G + blockIdx.x] = ki;
      }
      if (ki != INT_MAX) {
        const size_t off = (size_t)(ki - row_base) * jb;
        T* dst = ws.cand_row + ((size_t)par * G + blockIdx.x) * XROW;
        for (int q = 0; q < nq; ++q)
          if (lane + 32 * q < jb) {
            dst[lane + 32 * q] = rows[off + lane + 32 * q];
            if (EXACT) dst[MAX_NB + lane + 32 * q] = sums[off + lane + 32 * q];
          }
      }
    } else if (warp == 1) {
      if (diag >= row_base && diag < row_base + nloc) {  // this CTA holds the diagonal row: publish it for the swap
        const size_t off = (size_t)(diag - row_base) * jb;
        T* dst = ws.top_row + (size_t)par * XROW;
        for (int q = 0; q < nq; ++q)
          if (lane + 32 * q < jb) {
            dst[lane + 32 * q] = rows[off + lane + 32 * q];
            if (EXACT) dst[MAX_NB + lane + 32 * q] = sums[off + lane + 32 * q];
          }
      }
    }
    __syncthreads();  // (B) publication complete within the CTA
    bar_target += G;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&ws.hdr->barrier, 1u);
      while (ld_acquire_u32(&ws.hdr->barrier) < bar_target) {
      }
    }
    __syncthreads();  // (C) every CTA's candidates are visible

    // ---- every warp redundantly reduces the G candidates (L1 is not coherent: read through L2) ----
    double k = -2.0;
    int p = INT_MAX, pcta = 0;
    for (int b = lane; b < G; b += 32) {
      const double k2 = __ldcg(&ws.cand_key[par * G + b]);
      const int i2 = __ldcg(&ws.cand_idx[par * G + b]);
      if (k2 > k || (k2 == k && i2 < p)) {
        k = k2;
        p = i2;
        pcta = b;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double k2 = __shfl_xor_sync(0xffffffffu, k, off);
      const int i2 = __shfl_xor_sync(0xffffffffu, p, off);
      const int c2 = __shfl_xor_sync(0xffffffffu, pcta, off);
      if (k2 > k || (k2 == k && i2 < p)) {
        k = k2;
        p = i2;
        pcta = c2;
      }
    }
    // p = absolute pivot row (>= diag), held by CTA pcta
    const T* prow_g = ws.cand_row + ((size_t)par * G + pcta) * XROW;
    T prow[MAX_NB / 32];  // the pivot row as stored (cols < c: final L entries; cols >= c: original values in EXACT)
    T u[MAX_NB / 32];     // its current values = row c of U for cols >= c
#pragma unroll
    for (int q = 0; q < MAX_NB / 32; ++q) {
      const int col = lane + 32 * q;
      prow[q] = (col < jb) ? __ldcg(&prow_g[col]) : (T)0;
      u[q] = prow[q];
      if (EXACT && col < jb) u[q] = sub_rn(prow[q], __ldcg(&prow_g[MAX_NB + col]));
    }
    T pv = __ldcg(&prow_g[c]);
    if (EXACT) pv = sub_rn(pv, __ldcg(&prow_g[MAX_NB + c]));

    if (blockIdx.x == 0 && threadIdx.x == 0) ws.hdr->ipiv[c] = p;

    // ---- interchange (whole panel row; the rest of the row is swapped by lu_swap*_kernel) ----
    if (p != diag && p >= row_base && p < row_base + nloc && ((p - row_base) % PANEL_WARPS) == warp) {
      const T* trow_g = ws.top_row + (size_t)par * XROW;
      const size_t off = (size_t)(p - row_base) * jb;
#pragma unroll
      for (int q = 0; q < MAX_NB / 32; ++q)
        if (lane + 32 * q < jb) {
          rows[off + lane + 32 * q] = __ldcg(&trow_g[lane + 32 * q]);
          if (EXACT) sums[off + lane + 32 * q] = __ldcg(&trow_g[MAX_NB + lane + 32 * q]);
        }
    }
    if ((p != diag || EXACT) && diag >= row_base && diag < row_base + nloc &&
        ((diag - row_base) % PANEL_WARPS) == warp) {
      const size_t off = (size_t)(diag - row_base) * jb;  // the diagonal row becomes final: L for cols < c, U for >= c
#pragma unroll
      for (int q = 0; q < MAX_NB / 32; ++q) {
        const int col = lane + 32 * q;
        if (col < jb) rows[off + col] = (col >= c) ? u[q] : prow[q];
      }
    }
    __syncwarp();

    // ---- scale column c and rank-1 update of the warp's rows below the diagonal; track arg-max of column c+1 ----
    double nk = -2.0;
    int nki = INT_MAX;
    const int cn = c + 1;                     // next column
    const bool track = (cn < jb) && (lane == (cn & 31));
    const int qn = cn >> 5;
    int lr0 = warp;
    if (row_base <= diag) {                   // skip rows at or above the diagonal
      const int first = diag + 1 - row_base;  // first local row strictly below the diagonal
      lr0 = first + ((warp - first) % PANEL_WARPS + PANEL_WARPS) % PANEL_WARPS;
    }
    for (int lr = lr0; lr < nloc; lr += PANEL_WARPS) {
      T* r = rows + (size_t)lr * jb;
      T* sacc = sums + (size_t)lr * jb;
      T l = value(lr, c);
      if (pv != (T)0) l = l / pv;             // true division, skipped for an exactly-zero pivot (lu.rs:156-160)
      __syncwarp();
      if (lane == (c & 31)) r[c] = l;         // final L entry
      T nv = (T)0;
#pragma unroll
      for (int q = 0; q < MAX_NB / 32; ++q) {
        const int col = lane + 32 * q;
        if (col > c && col < jb) {
          T v;
          if (EXACT) {
            const T sn = add_rn(sacc[col], mul_rn(l, u[q]));  // s = s + l*u, k ascending (lu.rs:125)
            sacc[col] = sn;
            v = sub_rn(r[col], sn);
          } else {
            v = sub_rn(r[col], mul_rn(l, u[q]));
            r[col] = v;
          }
          if (q == qn) nv = v;
        }
      }
      if (track) key_merge(nk, nki, pivot_key(nv, false), row_base + lr);
    }
    // the row that becomes the next diagonal row (absolute row diag+1) takes part with its own value as the incumbent
    if (cn < jb) {
      const int nd = diag + 1;
      if (nd >= row_base && nd < row_base + nloc && ((nd - row_base) % PANEL_WARPS) == warp && nd < m) {
        if (track) {
          const double kk = pivot_key(value(nd - row_base, cn), true);
          if (kk == (double)INFINITY) {  // NaN incumbent: never displaced
            nk = kk;
            nki = nd;
          }
        }
      }
      nk = __shfl_sync(0xffffffffu, nk, cn & 31);
      nki = __shfl_sync(0xffffffffu, nki, cn & 31);
      if (lane == 0) {
        wkey[warp] = nk;
        widx[warp] = nki;
      }
    }
  }
  __syncthreads();

  // ---- write the factored panel back ----
  for (int idx = threadIdx.x; idx < nloc * jb; idx += PANEL_THREADS) {
    const int lr = idx / jb, c = idx - lr * jb;
    A[(size_t)(row_base + lr) * ld + j0 + c] = rows[idx];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. net permutation of the panel's interchanges + piv / sign bookkeeping (one warp)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void lu_perm_kernel(void* ws_base, int G, int j0, int jb, uint64_t* __restrict__ piv, int* __restrict__ sign) {
  __shared__ int pos[MAX_MOVES];
  __shared__ int src[MAX_MOVES];
  __shared__ int ipiv_s[MAX_NB];
  __shared__ uint64_t pold[MAX_MOVES];
  const WsView<T> ws = ws_view<T>(ws_base, G);
  const int lane = threadIdx.x;
  for (int i = lane; i < jb; i += 32) {
    pos[i] = j0 + i;
    src[i] = j0 + i;
    ipiv_s[i] = ws.hdr->ipiv[i];
  }
  __syncwarp();
  int count = jb;
  int flips = 0;
  for (int c = 0; c < jb; ++c) {
    const int p = ipiv_s[c];
    if (p == j0 + c) continue;  // warp-uniform
    ++flips;
    int k;
    if (p < j0 + jb) {
      k = p - j0;
    } else {
      int found = -1;
      for (int base = jb; base < count; base += 32) {
        const int i = base + lane;
        const unsigned hit = __ballot_sync(0xffffffffu, i < count && pos[i] == p);
        if (hit) {
          found = base + __ffs(hit) - 1;
          break;
        }
      }
      if (found < 0) {
        if (lane == 0) {
          pos[count] = p;
          src[count] = p;
        }
        found = count++;
      }
      k = found;
    }
    __syncwarp();
    if (lane == 0) {
      const int t = src[c];
      src[c] = src[k];
      src[k] = t;
    }
    __syncwarp();
  }
  // compact the rows that actually move; `piv` is permuted exactly like a matrix column (lu.rs:147-149)
  int nm = 0;
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    const bool mv = i < count && pos[i] != src[i];
    const unsigned mask = __ballot_sync(0xffffffffu, mv);
    if (mv) {
      const int slot = nm + __popc(mask & ((1u << lane) - 1));
      ws.hdr->move_dst[slot] = pos[i];
      ws.hdr->move_src[slot] = src[i];
      pold[slot] = piv[src[i]];
    }
    nm += __popc(mask);
    __syncwarp();
  }
  __syncwarp();
  for (int i = lane; i < nm; i += 32) piv[ws.hdr->move_dst[i]] = pold[i];
  if (lane == 0) {
    ws.hdr->n_moves = nm;
    if (flips & 1) *sign = !*sign;  // pospivsign flips once per interchange (lu.rs:151)
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. apply the net permutation to a column range [col0, col1) (left of the panel)
// ---------------------------------------------------------------------------------------------------------------
constexpr int SWAP_W = 32;  // columns per CTA strip
// Columns [col0, col1) EXCLUDING the panel's own columns [skip0, skip1) (already interchanged inside the panel kernel).
template <typename T>
__global__ void __launch_bounds__(256) lu_swap_kernel(T* __restrict__ A, size_t ld, int col0, int col1, int skip0,
                                                      int skip1, const void* ws_base, int G) {
  extern __shared__ __align__(16) unsigned char swap_smem[];
  T(*stage)[SWAP_W] = reinterpret_cast<T(*)[SWAP_W]>(swap_smem);  // [MAX_MOVES][SWAP_W]
  __shared__ int mdst[MAX_MOVES];
  __shared__ int msrc[MAX_MOVES];
  const WsView<T> ws = ws_view<T>(const_cast<void*>(ws_base), G);
  const int nm = ws.hdr->n_moves;
  if (nm == 0) return;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) {
    mdst[i] = ws.hdr->move_dst[i];
    msrc[i] = ws.hdr->move_src[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int col = col0 + blockIdx.x * SWAP_W + lane;
  if (col >= skip0) col += skip1 - skip0;
  const bool ok = col < col1;
  for (int i = warp; i < nm; i += 8)
    if (ok) stage[i][lane] = __ldcg(&A[(size_t)msrc[i] * ld + col]);
  __syncthreads();
  for (int i = warp; i < nm; i += 8)
    if (ok) A[(size_t)mdst[i] * ld + col] = stage[i][lane];
}

// ---------------------------------------------------------------------------------------------------------------
// 3b. W = L11^-1 for the jb x jb unit-lower-triangular diagonal block (one CTA).  Multi-panel factorisations compute
//     U12 = L11^-1 * A12 as W * A12 on the DMMA GEMM instead of a latency-bound substitution.
//     Blocked: invert the 32 x 32 diagonal blocks by substitution (one thread per column), then fill the blocks below
//     the diagonal, distance by distance: W_ij = -W_ii * sum_{k=j}^{i-1} L_ik * W_kj.
// ---------------------------------------------------------------------------------------------------------------
constexpr int IB = 32;
constexpr int INVL_LD = MAX_NB + 1;
template <typename T>
__global__ void __launch_bounds__(256) lu_invl_kernel(const T* __restrict__ A, size_t ld, int j0, int jb,
                                                      T* __restrict__ W /* [MAX_NB][MAX_NB] */) {
  // One padded square in shared memory holds both operands: the lower triangle (with diagonal) is W, the strictly
  // upper triangle is L11 transposed (L[i][k], k < i, lives at SQ[k][i]).  Products in flight use a separate scratch.
  extern __shared__ __align__(16) unsigned char invl_smem[];
  T(*SQ)[INVL_LD] = reinterpret_cast<T(*)[INVL_LD]>(invl_smem);
  T* scratch = reinterpret_cast<T*>(invl_smem) + (size_t)MAX_NB * INVL_LD;  // [<= 3][IB][IB]
  const int tid = threadIdx.x;
  for (int idx = tid; idx < MAX_NB * MAX_NB; idx += blockDim.x) {
    const int i = idx / MAX_NB, k = idx - i * MAX_NB;
    if (k < i) {
      SQ[k][i] = (i < jb) ? __ldcg(&A[(size_t)(j0 + i) * ld + j0 + k]) : (T)0;  // L^T into the upper triangle
      SQ[i][k] = (T)0;
    } else if (k == i) {
      SQ[i][i] = (T)0;
    }
  }
  __syncthreads();
  auto Lat = [&](int i, int k) -> T { return SQ[k][i]; };  // L[i][k], k < i
  const int nblk = (jb + IB - 1) / IB;
  // diagonal blocks: thread (b, j) solves L_bb * w = e_j by forward substitution
  if (tid < MAX_NB) {
    const int b = tid / IB, j = tid % IB, o = b * IB;
    if (o + j < jb) {
      SQ[o + j][o + j] = (T)1;
      for (int i = j + 1; i < IB && o + i < jb; ++i) {
        T acc = (T)0;
        for (int k = j; k < i; ++k) acc += Lat(o + i, o + k) * SQ[o + k][o + j];
        SQ[o + i][o + j] = -acc;
      }
    }
  }
  __syncthreads();
  for (int d = 1; d < nblk; ++d) {
    const int pairs = nblk - d;  // blocks (i, j) = (d + pr, pr)
    // phase 1: T_ij = sum_{k=j}^{i-1} L_ik * W_kj
    for (int e = tid; e < pairs * IB * IB; e += blockDim.x) {
      const int pr = e / (IB * IB), r = (e / IB) % IB, c = e % IB;
      const int i = d + pr, j = pr;
      T acc = (T)0;
      // W_jj is lower triangular: its entries above the diagonal are zero (that part of SQ holds L^T), so start at c
      for (int kk = j * IB + c; kk < i * IB; ++kk) acc += Lat(i * IB + r, kk) * SQ[kk][j * IB + c];
      scratch[e] = acc;
    }
    __syncthreads();
    // phase 2: W_ij = -W_ii * T_ij
    for (int e = tid; e < pairs * IB * IB; e += blockDim.x) {
      const int pr = e / (IB * IB), r = (e / IB) % IB, c = e % IB;
      const int i = d + pr, j = pr;
      T acc = (T)0;
      for (int kk = 0; kk <= r; ++kk) acc += SQ[i * IB + r][i * IB + kk] * scratch[pr * IB * IB + kk * IB + c];
      SQ[i * IB + r][j * IB + c] = -acc;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < MAX_NB * MAX_NB; idx += blockDim.x) {
    const int i = idx / MAX_NB, k = idx - i * MAX_NB;
    W[idx] = (k <= i && i < jb) ? SQ[i][k] : (T)0;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 4. right of the panel: permutation + U12 = L11^-1 * A12 (the i <= j part of lu.rs:122-129).
//    One CTA per strip of TRSM_W columns.  Each element keeps its original value and a running sum of separately
//    rounded products (k ascending); U[i][j] = a - s is formed once -- the reference's expression, bit for bit.
// ---------------------------------------------------------------------------------------------------------------
constexpr int TRSM_W = 32;
template <typename T>
__global__ void __launch_bounds__(256)
lu_swap_trsm_kernel(T* __restrict__ A, size_t ld, int j0, int jb, int col0, int col1, const void* ws_base, int G) {
  extern __shared__ __align__(16) unsigned char trsm_smem[];
  T(*X)[TRSM_W] = reinterpret_cast<T(*)[TRSM_W]>(trsm_smem);            // [MAX_NB]: top jb rows of the strip (a)
  T(*S)[TRSM_W] = X + MAX_NB;                                           // [MAX_NB]: running sums
  T(*stage)[TRSM_W] = S + MAX_NB;                                       // [MAX_NB]: rows leaving the top block
  __shared__ int top_src[MAX_NB];
  __shared__ int out_dst[MAX_NB];
  __shared__ int out_src[MAX_NB];
  __shared__ int n_out;
  const WsView<T> ws = ws_view<T>(const_cast<void*>(ws_base), G);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int col = col0 + blockIdx.x * TRSM_W + lane;
  const bool ok = col < col1;

  for (int i = threadIdx.x; i < jb; i += blockDim.x) top_src[i] = j0 + i;
  if (threadIdx.x == 0) n_out = 0;
  __syncthreads();
  const int nm = ws.hdr->n_moves;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) {
    const int d = ws.hdr->move_dst[i], s = ws.hdr->move_src[i];
    if (d < j0 + jb) {
      top_src[d - j0] = s;
    } else {
      const int slot = atomicAdd(&n_out, 1);
      out_dst[slot] = d;
      out_src[slot] = s;
    }
  }
  __syncthreads();
  const int no = n_out;
  // gather
  for (int i = warp; i < jb; i += 8) {
    X[i][lane] = ok ? __ldcg(&A[(size_t)top_src[i] * ld + col]) : (T)0;
    S[i][lane] = (T)0;
  }
  for (int i = warp; i < no; i += 8) stage[i][lane] = ok ? __ldcg(&A[(size_t)out_src[i] * ld + col]) : (T)0;
  __syncthreads();
  // rows that left the top block
  for (int i = warp; i < no; i += 8)
    if (ok) A[(size_t)out_dst[i] * ld + col] = stage[i][lane];

  // forward substitution, right-looking over k with deferred subtraction
  const T* L = A + (size_t)j0 * ld + j0;  // L11, unit lower, written by the panel kernel
  for (int k = 0; k < jb; ++k) {
    const T xk = sub_rn(X[k][lane], S[k][lane]);  // U[k][col], final
    if (warp == (k & 7) && ok) A[(size_t)(j0 + k) * ld + col] = xk;
    for (int i = k + 1 + warp; i < jb; i += 8) S[i][lane] = add_rn(S[i][lane], mul_rn(__ldcg(&L[(size_t)i * ld + k]), xk));
    __syncthreads();
  }
}

template <typename T>
__global__ void lu_init_piv_kernel(uint64_t* __restrict__ piv, int m, int* __restrict__ sign) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) piv[i] = (uint64_t)i;  // lu.rs:108-111
  if (i == 0) *sign = 1;            // pospivsign = true, lu.rs:113
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
namespace {
// Per host thread and device: the look-ahead stream (highest priority) and the two events that fence it.
struct LuSide {
  cudaStream_t sp = nullptr;
  cudaEvent_t e1 = nullptr, e2 = nullptr;
};
int lu_side(int device, LuSide** out) {
  static thread_local LuSide side[64];
  LA_REQUIRE(device >= 0 && device < 64, "device ordinal out of range");
  LuSide& s = side[device];
  if (!s.sp) {
    int lo = 0, hi = 0;
    LA_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&s.sp, cudaStreamNonBlocking, hi));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e1, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e2, cudaEventDisableTiming));
  }
  *out = &s;
  return LA_OK;
}
}  // namespace

int gemm_f64_tensor(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, cudaStream_t st);  // gemm_f64.cu: TMA/DMMA kernel regardless of size

template <typename T>
int lu_factor_dev(T* LU, size_t m, size_t n, uint64_t* piv_dev, int* sign_dev, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(LU && piv_dev && sign_dev, "la_lu_factor: null pointer");
  LA_REQUIRE(m > 0 && n > 0, "la_lu_factor: zero dimension (m=%zu n=%zu)", m, n);
  LA_REQUIRE(m < (1u << 30) && n < (1u << 30), "la_lu_factor: dimension too large");
  if (!ctx->coop) return fail(LA_ERR_UNSUPPORTED, "la_lu_factor: device lacks cooperative launch");

  const int M = (int)m, N = (int)n;
  const int kmin = M < N ? M : N;
  const int sms = ctx->sm_count;

  // panel width: as wide as shared memory allows for the tallest (first) panel, multiple of 16, <= MAX_NB
  int rpc_first = (M + sms - 1) / sms;
  if (rpc_first < 8) rpc_first = 8;
  int nb = (int)(PANEL_SMEM_BUDGET / ((size_t)rpc_first * sizeof(T)));
  nb = nb / 16 * 16;
  if (nb > MAX_NB) nb = MAX_NB;
  if (nb < 16)
    return fail(LA_ERR_UNSUPPORTED, "la_lu_factor: %d rows exceed the shared-memory panel capacity of %d SMs", M, sms);
  // single-panel factorisations run the bit-exact (deferred subtraction) panel when twice the panel fits
  const bool exact = kmin <= nb && (size_t)2 * rpc_first * kmin * sizeof(T) <= PANEL_SMEM_BUDGET;
  // multi-panel fp64 on TMA-addressable storage: look-ahead pipeline with U12 = inv(L11) * A12 on the DMMA GEMM
  static const int dbg = getenv("LA_LU_DEBUG") ? atoi(getenv("LA_LU_DEBUG")) : 0;  // 1: no look-ahead, 2: plain loop
  const bool fast = std::is_same<T, double>::value && kmin > nb && (N % 2 == 0) && ((uintptr_t)LU % 16 == 0) &&
                    (kmin % 2 == 0 || kmin == N) && dbg != 2;

  void* ws_base = nullptr;
  LA_TRY(scratch_get(ctx->device, 8, ws_bytes<T>(sms), &ws_base));
  void* w_base = nullptr;
  LA_TRY(scratch_get(ctx->device, 11, sizeof(T) * MAX_NB * MAX_NB, &w_base));
  T* W = (T*)w_base;
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_panel_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)PANEL_SMEM_BUDGET + 2048));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_panel_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)PANEL_SMEM_BUDGET + 2048));

  const int SWAP_SMEM = (int)(sizeof(T) * MAX_MOVES * SWAP_W);
  const int TRSM_SMEM = (int)(sizeof(T) * 3 * MAX_NB * TRSM_W);
  const int INVL_SMEM = (int)(sizeof(T) * ((size_t)MAX_NB * INVL_LD + 3 * IB * IB));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_swap_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWAP_SMEM));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_swap_trsm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_invl_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, INVL_SMEM));

  {
    // every kernel of the pipeline prefers the maximum shared-memory carve-out, so co-resident kernels never ask an SM
    // for a different L1/shared split
    const void* fns[] = {(const void*)lu_panel_kernel<T, false>, (const void*)lu_panel_kernel<T, true>,
                         (const void*)lu_perm_kernel<T>,         (const void*)lu_swap_kernel<T>,
                         (const void*)lu_swap_trsm_kernel<T>,    (const void*)lu_invl_kernel<T>,
                         (const void*)lu_init_piv_kernel<T>};
    for (const void* f : fns)
      LA_CUDA_TRY(cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  lu_init_piv_kernel<T><<<(M + 255) / 256, 256, 0, st>>>(piv_dev, M, sign_dev);
  LA_CUDA_TRY(cudaGetLastError());

  int G_cur = 1;
  // panel factorisation + net permutation / piv bookkeeping of columns [j0, j0+jb) on stream s
  auto launch_panel = [&](int j0, int jb, cudaStream_t s) -> int {
    const int R = M - j0;
    int rpc = (R + sms - 1) / sms;
    if (rpc < 8) rpc = 8;  // at least one row per warp; fewer, fuller CTAs make the barrier cheaper
    const int G = (R + rpc - 1) / rpc;
    size_t smem = (size_t)rpc * jb * sizeof(T) * (exact ? 2 : 1);
    if ((dbg == 6 || dbg == 9) && smem < 150 * 1024) smem = 150 * 1024;  // debug: too big to co-reside with a GEMM CTA
    LA_CUDA_TRY(cudaMemsetAsync(ws_base, 0, sizeof(unsigned int) * 4, s));
    T* a = LU;
    size_t ld = n;
    int mm = M, jj0 = j0, jjb = jb, rr = rpc;
    void* wsb = ws_base;
    void* args[] = {&a, &ld, &mm, &jj0, &jjb, &rr, &wsb};
    const void* fn = exact ? (const void*)lu_panel_kernel<T, true> : (const void*)lu_panel_kernel<T, false>;
    if (dbg == 8)
      LA_CUDA_TRY(cudaLaunchKernel(fn, dim3(G), dim3(PANEL_THREADS), args, smem, s));  // debug: plain launch
    else
      LA_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(PANEL_THREADS), args, smem, s));
    G_cur = G;
    if ((dbg == 7 || dbg == 9) && s != st) return LA_OK;  // debug: perm deferred to the main stream
    lu_perm_kernel<T><<<1, 32, 0, s>>>(ws_base, G, j0, jb, piv_dev, sign_dev);
    LA_CUDA_TRY(cudaGetLastError());
    return LA_OK;
  };

  if (!fast) {
    // ---- plain right-looking loop (single panel / fp32 / odd leading dimension): substitution TRSM kernel ----
    for (int j0 = 0; j0 < kmin; j0 += nb) {
      const int jb = (kmin - j0 < nb) ? (kmin - j0) : nb;
      LA_TRY(launch_panel(j0, jb, st));
      if (j0 > 0) {
        lu_swap_kernel<T><<<(j0 + SWAP_W - 1) / SWAP_W, 256, SWAP_SMEM, st>>>(LU, n, 0, j0, j0, j0, ws_base, G_cur);
        LA_CUDA_TRY(cudaGetLastError());
      }
      const int c1 = j0 + jb;
      if (c1 < N) {
        lu_swap_trsm_kernel<T><<<(N - c1 + TRSM_W - 1) / TRSM_W, 256, TRSM_SMEM, st>>>(LU, n, j0, jb, c1, N, ws_base,
                                                                                     G_cur);
        LA_CUDA_TRY(cudaGetLastError());
        if (c1 < M)
          LA_TRY(gemm_dev<T>(LU + (size_t)c1 * n + j0, n, LU + (size_t)j0 * n + c1, n, LU + (size_t)c1 * n + c1, n,
                             (size_t)(M - c1), (size_t)jb, (size_t)(N - c1), LA_GEMM_SUB, st));
      }
    }
    return LA_OK;
  }

  // ---- look-ahead pipeline (fp64) ----
  // Stream st: row interchanges, inv(L11), U12 = inv(L11)*A12 and the trailing update, next panel's columns FIRST.
  // Stream sp (highest priority): the next panel's factorisation, overlapping the rest of the trailing update.  The
  // cooperative panel CTAs (latency-bound, shared-memory resident) co-reside with the DMMA GEMM CTAs on the SMs.
  if constexpr (std::is_same<T, double>::value) {
    LuSide* side;
    LA_TRY(lu_side(ctx->device, &side));
    cudaStream_t sp = dbg == 1 ? st : side->sp;
    double* A = LU;
    const size_t ld = n;
    LA_TRY(launch_panel(0, nb, st));
    const int stop_iters = getenv("LA_LU_STOP") ? atoi(getenv("LA_LU_STOP")) : (1 << 30);  // debug: truncate the loop
    for (int j0 = 0; j0 < kmin; j0 += nb) {
      if (j0 / nb >= stop_iters) break;
      const int jb = (kmin - j0 < nb) ? (kmin - j0) : nb;
      const int c1 = j0 + jb;
      if (N - jb > 0) {
        lu_swap_kernel<T><<<(N - jb + SWAP_W - 1) / SWAP_W, 256, SWAP_SMEM, st>>>(LU, n, 0, N, j0, c1, ws_base, G_cur);
        LA_CUDA_TRY(cudaGetLastError());
      }
      if (c1 >= N) break;
      lu_invl_kernel<T><<<1, 256, INVL_SMEM, st>>>(LU, n, j0, jb, W);
      LA_CUDA_TRY(cudaGetLastError());
      const double* L21 = A + (size_t)c1 * ld + j0;
      auto trsm_update = [&](int cb, int ce) -> int {  // columns [cb, ce)
        double* U12 = A + (size_t)j0 * ld + cb;
        LA_TRY(gemm_f64_tensor(W, MAX_NB, U12, ld, U12, ld, (size_t)jb, (size_t)jb, (size_t)(ce - cb), LA_GEMM_ASSIGN,
                               st));  // in place: one tile row, every CTA reads its whole column block first
        if (c1 < M)
          LA_TRY(gemm_f64_tensor(L21, ld, U12, ld, A + (size_t)c1 * ld + cb, ld, (size_t)(M - c1), (size_t)jb,
                                 (size_t)(ce - cb), LA_GEMM_SUB, st));
        return LA_OK;
      };
      if (c1 < kmin) {
        const int nb2 = (kmin - c1 < nb) ? (kmin - c1) : nb;
        const int c2 = c1 + nb2;
        LA_TRY(trsm_update(c1, c2));  // the next panel's columns first
        if (dbg == 5 && c2 < N) {  // debug: TRSM of the rest before the panel starts, only the update overlaps
          double* U12r = A + (size_t)j0 * ld + c2;
          LA_TRY(gemm_f64_tensor(W, MAX_NB, U12r, ld, U12r, ld, (size_t)jb, (size_t)jb, (size_t)(N - c2), LA_GEMM_ASSIGN, st));
        }
        LA_CUDA_TRY(cudaEventRecord(side->e1, st));
        LA_CUDA_TRY(cudaStreamWaitEvent(sp, side->e1, 0));
        LA_TRY(launch_panel(c1, nb2, sp));
        LA_CUDA_TRY(cudaEventRecord(side->e2, sp));
        if (dbg == 3) LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e2, 0));  // debug: two streams, no overlap
        if (c2 < N) {
          if (dbg == 4 || dbg == 5) {
            double* U12r = A + (size_t)j0 * ld + c2;
            if (dbg == 4) {  // debug: only the TRSM overlaps
              LA_TRY(gemm_f64_tensor(W, MAX_NB, U12r, ld, U12r, ld, (size_t)jb, (size_t)jb, (size_t)(N - c2), LA_GEMM_ASSIGN, st));
              LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e2, 0));
            }
            LA_TRY(gemm_f64_tensor(L21, ld, U12r, ld, A + (size_t)c1 * ld + c2, ld, (size_t)(M - c1), (size_t)jb,
                                   (size_t)(N - c2), LA_GEMM_SUB, st));
          } else {
            LA_TRY(trsm_update(c2, N));  // overlaps the panel on sp
          }
        }
        LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e2, 0));
        if (dbg == 7 || dbg == 9) {
          lu_perm_kernel<T><<<1, 32, 0, st>>>(ws_base, G_cur, c1, nb2, piv_dev, sign_dev);
          LA_CUDA_TRY(cudaGetLastError());
        }
      } else {
        LA_TRY(trsm_update(c1, N));
      }
    }
  }
  return LA_OK;
}
template int lu_factor_dev<double>(double*, size_t, size_t, uint64_t*, int*, cudaStream_t);
template int lu_factor_dev<float>(float*, size_t, size_t, uint64_t*, int*, cudaStream_t);

}  // namespace la
