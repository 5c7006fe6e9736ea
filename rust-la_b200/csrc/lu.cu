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
#include <time.h>

#include <string>
#include <type_traits>
#include <vector>
#include <stdio.h>

#include "la_common.cuh"
#include "ll_exchange.cuh"

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
constexpr int ROW_REPLICAS = 4;  // copies of every candidate row (spreads the readers of the winner's row)
// Net permutation of one panel's interchanges: row dst[i] receives old row src[i].  Two lists (panel parity): the
// bulk stream may still be applying panel i's list to the far columns while the chain stream builds panel i+1's.
struct MoveList {
  int n_moves;
  int pad[3];
  int dst[MAX_MOVES];
  int src[MAX_MOVES];
};
// One candidate record per CTA and step: |pivot candidate| as a double and its absolute row, packed into two
// self-validating 8-byte units (16 bytes, one load per record): {key.hi32 | row.hi16 | tag16}, {key.lo32 | row.lo16 |
// tag16}.  A 16-bit tag suffices here: a slot is rewritten at the same step of every panel, so the only stale values
// a reader can meet are one panel (or two steps) old.
struct Rec {
  unsigned long long a, b;
};
__device__ __forceinline__ unsigned long long rec_pack(unsigned key32, unsigned row16, unsigned tag16) {
  return ((unsigned long long)key32 << 32) | ((unsigned long long)(row16 & 0xffffu) << 16) | (tag16 & 0xffffu);
}

template <typename T>
struct PanelWs {
  int ipiv[MAX_NB];  // absolute pivot row chosen for column j0 + c
  MoveList moves[4];  // [0], [1]: whole panels by parity; [2], [3]: the two 64-wide halves of a split panel
  // double-buffered (step parity) exchange area, laid out after the struct:
  //   Rec inbox[2][G reader][G writer]; LL<T>::word cand_row[2][G][XROW]; LL<T>::word top_row[2][XROW];
  // Every CTA pushes its record into each reader's PRIVATE inbox and polls only its own: an all-to-all in which no
  // cache line has more than one poller (148 CTAs spinning on the same 37 lines cost ~9000 cycles per column).
};

template <typename T>
__host__ __device__ inline size_t ws_hdr_bytes() {
  return (sizeof(PanelWs<T>) + 31) & ~(size_t)31;
}
template <typename T>
__host__ __device__ inline size_t ws_bytes(int G) {
  return ws_hdr_bytes<T>() + sizeof(Rec) * 2 * (size_t)G * G +
         sizeof(typename LL<T>::word) * (2 * (size_t)ROW_REPLICAS * G * XROW + 2 * XROW);
}
template <typename T>
struct WsView {
  PanelWs<T>* hdr;
  Rec* rec;                         // [2][G][G]
  typename LL<T>::word* cand_row;   // [2][ROW_REPLICAS][G][XROW]
  typename LL<T>::word* top_row;    // [2][XROW]
};
template <typename T>
__host__ __device__ inline WsView<T> ws_view(void* base, int G) {
  WsView<T> v;
  char* p = (char*)base;
  v.hdr = (PanelWs<T>*)p;
  p += ws_hdr_bytes<T>();
  v.rec = (Rec*)p;
  p += sizeof(Rec) * 2 * (size_t)G * G;
  v.cand_row = (typename LL<T>::word*)p;
  p += sizeof(typename LL<T>::word) * 2 * (size_t)ROW_REPLICAS * G * XROW;
  v.top_row = (typename LL<T>::word*)p;
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
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Per column c the CTAs must agree on the pivot (a grid-wide arg-max) and every CTA needs the pivot row: two dependent
// trips through L2.  The rank-1 update of column c is therefore split (look-ahead INSIDE the panel):
//   urgent : multipliers of column c, update of column c+1 only, arg-max of column c+1, and the full update of the two
//            rows that are published for step c+1 (the CTA's candidate row and, if it lives here, the next diagonal row);
//   bulk   : columns c+2.. of all other rows -- done by warps 1..7 WHILE warp 0 runs the exchange of step c+1.
// so the column time is max(exchange latency, update time) instead of their sum.
template <typename T, bool EXACT>
__global__ void __launch_bounds__(PANEL_THREADS, 1)
lu_panel_kernel(T* __restrict__ A, size_t ld, int m, int j0, int jb, int rows_per_cta, void* ws_base, unsigned epoch,
                int ipiv_off) {
  typedef typename LL<T>::word llw;
  constexpr int Q = MAX_NB / 32;
  extern __shared__ __align__(16) unsigned char panel_smem[];
  const int ldr = jb | 1;  // odd row stride: column walks (lanes <-> rows) and row walks are both conflict-free
  T* rows = reinterpret_cast<T*>(panel_smem);            // [rows_per_cta][ldr]: original values, then final L / U
  T* sums = rows + (size_t)rows_per_cta * ldr;           // [rows_per_cta][ldr]: deferred sums (EXACT only)
  __shared__ double wkey[PANEL_WARPS];
  __shared__ int widx[PANEL_WARPS];
  // per step parity: the pivot row as stored / its current values (row c of U) / the diagonal row it trades places with
  __shared__ T s_prow[2][MAX_NB];
  __shared__ T s_u[2][MAX_NB];
  __shared__ T s_trow[2][XROW];
  __shared__ int s_p[2];

  const int G = gridDim.x;
  const WsView<T> ws = ws_view<T>(ws_base, G);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row_base = j0 + blockIdx.x * rows_per_cta;          // first absolute row of this CTA
  const int nloc = max(0, min(rows_per_cta, m - row_base));     // rows held by this CTA
  const int WB = (nloc + PANEL_WARPS - 1) / PANEL_WARPS;        // urgent phase: warp w owns rows [w*WB, (w+1)*WB)

  auto value = [&](int lr, int col) -> T {
    const T a = rows[(size_t)lr * ldr + col];
    return EXACT ? sub_rn(a, sums[(size_t)lr * ldr + col]) : a;
  };

  // ---- load the CTA's rows of the panel into shared memory (row segments of jb contiguous elements) ----
  for (int idx = threadIdx.x; idx < nloc * jb; idx += PANEL_THREADS) {
    const int lr = idx / jb, c = idx - lr * jb;
    rows[(size_t)lr * ldr + c] = __ldcg(&A[(size_t)(row_base + lr) * ld + j0 + c]);
    if (EXACT) sums[(size_t)lr * ldr + c] = (T)0;
  }
  __syncthreads();

  // ---- per-warp arg-max of column 0 ----
  {
    double k = -2.0;
    int ki = INT_MAX;
    for (int l = lane; l < WB; l += 32) {
      const int lr = warp * WB + l;
      if (lr < nloc) key_merge(k, ki, pivot_key(rows[(size_t)lr * ldr], row_base + lr == j0), row_base + lr);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double k2 = __shfl_xor_sync(0xffffffffu, k, off);
      const int i2 = __shfl_xor_sync(0xffffffffu, ki, off);
      key_merge(k, ki, k2, i2);
    }
    if (lane == 0) {
      wkey[warp] = k;
      widx[warp] = ki;
    }
  }
  __syncthreads();

  // Publishes this CTA's contribution to step cn and (warp 0) runs the exchange; everything crosses CTAs as flagged
  // words -- no grid barrier, no atomics.  `c` = cn - 1 is the step whose rank-1 update is still pending on the two
  // published rows (c < 0: none).  Warps 0..3 each publish one REPLICA of the candidate row (readers spread over the
  // replicas, so the winner's row is not one 148-reader hot spot in L2); warp 1 also publishes the next diagonal row
  // and writes the two updated rows back once the other replica warps have read them (named barrier 1).
  // Returns (to every warp) the local indices of the rows the bulk update must skip.
  // The exchange area is double-buffered by step parity: a CTA publishes step cn+1 only after it has read every step
  // cn record, and those exist only once every CTA has finished reading step cn-1.
  auto publish_and_exchange = [&](int cn, int& skip_a, int& skip_b) {
    const int c = cn - 1;
    const int par = cn & 1;
    const int diag = j0 + cn;  // absolute row/col index of step cn's diagonal
    const unsigned tag = (epoch << 8) | (unsigned)(cn + 1);
    const unsigned tag16 = tag & 0xffffu;
    // CTA-wide candidate (every warp)
    double k = (lane < PANEL_WARPS) ? wkey[lane] : -2.0;
    int ki = (lane < PANEL_WARPS) ? widx[lane] : INT_MAX;
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) {
      const double k2 = __shfl_xor_sync(0xffffffffu, k, off);
      const int i2 = __shfl_xor_sync(0xffffffffu, ki, off);
      key_merge(k, ki, k2, i2);
    }
    k = __shfl_sync(0xffffffffu, k, 0);
    ki = __shfl_sync(0xffffffffu, ki, 0);
    const bool own_diag = diag >= row_base && diag < row_base + nloc;
    skip_a = (ki != INT_MAX) ? ki - row_base : -1;
    skip_b = own_diag ? diag - row_base : -1;
    if (warp >= ROW_REPLICAS) return;

    if (warp == 0) {  // the records first: they are what every other CTA is waiting for
      const unsigned long long kb = (unsigned long long)__double_as_longlong(k);
      for (int rd = lane; rd < G; rd += 32)  // one copy into every reader's inbox
        st_relaxed_2x64(&ws.rec[((size_t)par * G + rd) * G + blockIdx.x],
                        rec_pack((unsigned)(kb >> 32), (unsigned)ki >> 16, tag), rec_pack((unsigned)kb, (unsigned)ki, tag));
    }
    // a row with the pending update of step c applied (columns > cn; column cn was updated in the urgent phase)
    const T* uc = s_u[c & 1];
    auto updated_row = [&](int lr, T (&va)[Q], T (&vs)[Q]) {
      const size_t off = (size_t)lr * ldr;
      const T lk = (c >= 0) ? rows[off + c] : (T)0;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int col = lane + 32 * q;
        va[q] = vs[q] = (T)0;
        if (col < jb) {
          va[q] = rows[off + col];
          if (EXACT) vs[q] = sums[off + col];
          if (c >= 0 && col > cn) {
            if (EXACT) vs[q] = add_rn(vs[q], mul_rn(lk, uc[col]));  // s = s + l*u, k ascending (lu.rs:125)
            else va[q] = sub_rn(va[q], mul_rn(lk, uc[col]));
          }
        }
      }
    };
    auto publish_row = [&](llw* dst, const T (&va)[Q], const T (&vs)[Q]) {
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (lane + 32 * q < jb) {
          LL<T>::store(&dst[lane + 32 * q], va[q], tag);
          if (EXACT) LL<T>::store(&dst[MAX_NB + lane + 32 * q], vs[q], tag);
        }
    };
    auto write_back = [&](int lr, const T (&va)[Q], const T (&vs)[Q]) {
      const size_t off = (size_t)lr * ldr;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int col = lane + 32 * q;
        if (col < jb && col > cn) {
          if (EXACT) sums[off + col] = vs[q];
          else rows[off + col] = va[q];
        }
      }
    };
    T ca[Q], cs[Q];
    if (skip_a >= 0) {
      updated_row(skip_a, ca, cs);
      if (c >= 0 && warp != 1) named_bar_arrive(1, 32 * ROW_REPLICAS);  // our read of the row is done
      publish_row(ws.cand_row + (((size_t)par * ROW_REPLICAS + warp) * G + blockIdx.x) * XROW, ca, cs);
    }
    if (warp == 1) {
      T da[Q], ds[Q];
      if (skip_b >= 0) {
        updated_row(skip_b, da, ds);  // == the candidate row when skip_b == skip_a (not yet written back)
        publish_row(ws.top_row + (size_t)par * XROW, da, ds);
      }
      if (c >= 0) {
        if (skip_a >= 0) named_bar_sync(1, 32 * ROW_REPLICAS);  // warps 0, 2, 3 have read the candidate row
        if (skip_a >= 0) write_back(skip_a, ca, cs);
        if (skip_b >= 0 && skip_b != skip_a) write_back(skip_b, da, ds);
      }
      return;
    }
    if (warp != 0) return;

    // ---- warp 0: poll the own inbox (one poller per CTA: polling traffic delays the very stores it is waiting for);
    //      all loads of a pass are in flight together ----
    double bk = -2.0;
    int bp = INT_MAX, bcta = 0;
    constexpr int RPL = 5;
    for (int base = 0; base < G; base += 32 * RPL) {
      unsigned long long ra[RPL], rb[RPL];
      bool ok[RPL];
#pragma unroll
      for (int i = 0; i < RPL; ++i) ok[i] = base + lane + 32 * i >= G;
      bool all;
      do {
#pragma unroll
        for (int i = 0; i < RPL; ++i)
          if (!ok[i]) ld_relaxed_2x64(&ws.rec[((size_t)par * G + blockIdx.x) * G + base + lane + 32 * i], ra[i], rb[i]);
        all = true;
#pragma unroll
        for (int i = 0; i < RPL; ++i)
          if (!ok[i]) {
            ok[i] = (unsigned)(ra[i] & 0xffffu) == tag16 && (unsigned)(rb[i] & 0xffffu) == tag16;
            all &= ok[i];
          }
      } while (!all);
#pragma unroll
      for (int i = 0; i < RPL; ++i) {
        const int b = base + lane + 32 * i;
        if (b < G) {
          const double k2 = __longlong_as_double((long long)((ra[i] & 0xffffffff00000000ull) | (rb[i] >> 32)));
          const int i2 = (int)((((unsigned)(ra[i] >> 16) & 0xffffu) << 16) | ((unsigned)(rb[i] >> 16) & 0xffffu));
          if (k2 > bk || (k2 == bk && i2 < bp)) {
            bk = k2;
            bp = i2;
            bcta = b;
          }
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double k2 = __shfl_xor_sync(0xffffffffu, bk, off);
      const int i2 = __shfl_xor_sync(0xffffffffu, bp, off);
      const int c2 = __shfl_xor_sync(0xffffffffu, bcta, off);
      if (k2 > bk || (k2 == bk && i2 < bp)) {
        bk = k2;
        bp = i2;
        bcta = c2;
      }
    }
    // bp = absolute pivot row (>= diag), held by CTA bcta: fetch its row from "our" replica.  If the pivot row is one
    // of OURS (and an interchange is due) the diagonal row it trades places with is fetched in the same round trip --
    // every other CTA waits for this one in the next step.
    const llw* prow_g = ws.cand_row + (((size_t)par * ROW_REPLICAS + (blockIdx.x % ROW_REPLICAS)) * G + bcta) * XROW;
    const llw* trow_g = ws.top_row + (size_t)par * XROW;
    const bool own_p = bp != diag && bp >= row_base && bp < row_base + nloc;
    T pa[Q], ps[Q], ta[Q], ts[Q];
    bool ok[Q], okt[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      ok[q] = lane + 32 * q >= jb;
      okt[q] = ok[q] || !own_p;
      pa[q] = ps[q] = ta[q] = ts[q] = (T)0;
    }
    bool all;
    do {
      all = true;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        if (!ok[q]) {
          bool v = LL<T>::load(&prow_g[lane + 32 * q], tag, pa[q]);
          if (EXACT) v &= LL<T>::load(&prow_g[MAX_NB + lane + 32 * q], tag, ps[q]);
          ok[q] = v;
          all &= v;
        }
        if (!okt[q]) {
          bool v = LL<T>::load(&trow_g[lane + 32 * q], tag, ta[q]);
          if (EXACT) v &= LL<T>::load(&trow_g[MAX_NB + lane + 32 * q], tag, ts[q]);
          okt[q] = v;
          all &= v;
        }
      }
    } while (!all);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int col = lane + 32 * q;
      if (col < jb) {
        s_prow[par][col] = pa[q];
        s_u[par][col] = EXACT ? sub_rn(pa[q], ps[q]) : pa[q];
        if (own_p) {
          s_trow[par][col] = ta[q];
          if (EXACT) s_trow[par][MAX_NB + col] = ts[q];
        }
      }
    }
    if (lane == 0) s_p[par] = bp;
  };

  {
    int sa, sb;
    publish_and_exchange(0, sa, sb);
  }

  for (int c = 0; c < jb; ++c) {
    const int par = c & 1;
    const int diag = j0 + c;
    const int cn = c + 1;  // next column
    __syncthreads();       // (A) step c's pivot row is in shared memory; the bulk update of step c-1 is complete
    const int p = s_p[par];
    const T pv = s_u[par][c];
    if (blockIdx.x == 0 && threadIdx.x == 0) ws.hdr->ipiv[ipiv_off + c] = p;

    // ---- interchange (whole panel row; the rest of the row is swapped by lu_swap/lu_head kernels), by the warp that
    //      owns the row in the urgent phase below ----
    if (p != diag && p >= row_base && p < row_base + nloc && (p - row_base) / WB == warp) {
      const size_t off = (size_t)(p - row_base) * ldr;  // the old diagonal row (fetched by warp 0) moves down here
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (lane + 32 * q < jb) {
          rows[off + lane + 32 * q] = s_trow[par][lane + 32 * q];
          if (EXACT) sums[off + lane + 32 * q] = s_trow[par][MAX_NB + lane + 32 * q];
        }
    }
    const bool own_diag = diag >= row_base && diag < row_base + nloc;
    if ((p != diag || EXACT) && own_diag && (diag - row_base) / WB == warp) {
      const size_t off = (size_t)(diag - row_base) * ldr;  // the diagonal row becomes final: L for cols < c, U for >= c
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int col = lane + 32 * q;
        if (col < jb) rows[off + col] = (col >= c) ? s_u[par][col] : s_prow[par][col];
      }
    }
    __syncwarp();

    // ---- urgent (lanes <-> rows): multipliers of column c, update of column c+1, arg-max of column c+1 ----
    // EXACT keeps the reference's true division (lu.rs:158); the multi-panel path multiplies by the reciprocal of the
    // pivot (<= 1.5 ulp from the quotient, far inside the 1e-12*n element bar): a double division is a ~30 instruction
    // dependent sequence on the column's critical path.
    {
      const T rcp = EXACT ? (T)0 : (T)1 / pv;
      const T ucn = (cn < jb) ? s_u[par][cn] : (T)0;
      double nk = -2.0;
      int nki = INT_MAX;
      for (int l = lane; l < WB; l += 32) {
        const int lr = warp * WB + l;
        if (lr < nloc && row_base + lr > diag) {
          const size_t off = (size_t)lr * ldr;
          T lv = value(lr, c);
          if (pv != (T)0) lv = EXACT ? lv / pv : lv * rcp;  // skipped for an exactly-zero pivot (lu.rs:156-160)
          rows[off + c] = lv;                                // final L entry
          if (cn < jb) {
            T v;
            if (EXACT) {
              const T sv = add_rn(sums[off + cn], mul_rn(lv, ucn));
              sums[off + cn] = sv;
              v = sub_rn(rows[off + cn], sv);
            } else {
              v = sub_rn(rows[off + cn], mul_rn(lv, ucn));
              rows[off + cn] = v;
            }
            // strict '>' + lowest row == the reference's first maximum
            key_merge(nk, nki, pivot_key(v, row_base + lr == diag + 1), row_base + lr);
          }
        }
      }
      if (cn < jb) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double k2 = __shfl_xor_sync(0xffffffffu, nk, off);
          const int i2 = __shfl_xor_sync(0xffffffffu, nki, off);
          key_merge(nk, nki, k2, i2);
        }
        if (lane == 0) {
          wkey[warp] = nk;
          widx[warp] = nki;
        }
      }
    }
    if (cn >= jb) break;
    __syncthreads();  // (B) column c of L, column c+1 and its per-warp arg-max are in shared memory

    int skip_a, skip_b;
    publish_and_exchange(cn, skip_a, skip_b);

    // ---- bulk (lanes <-> columns), warps 1..7, under the exchange: columns > c+1 of the rows below the diagonal ----
    if (warp > 0 && cn + 1 < jb) {
      T u[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int col = lane + 32 * q;
        u[q] = (col > cn && col < jb) ? s_u[par][col] : (T)0;
      }
      const int first = max(0, diag + 1 - row_base);
      const int q0 = (cn + 1) >> 5;  // first column group with work left
      // four rows per trip, loads / arithmetic / stores grouped: the compiler cannot reorder the shared-memory stores
      // of one row behind the loads of the next (possible aliasing), and one row at a time is a ~200-cycle dependent
      // chain (FP64 latency) -- too slow for CTAs that hold 160+ rows
      constexpr int UR = 4;
      for (int lrb = first + warp - 1; lrb < nloc; lrb += UR * (PANEL_WARPS - 1)) {
        T li[UR];
        T v[UR][Q];
        bool live[UR];
#pragma unroll
        for (int i = 0; i < UR; ++i) {
          const int lr = lrb + i * (PANEL_WARPS - 1);
          live[i] = lr < nloc && lr != skip_a && lr != skip_b;
          const size_t off = (size_t)(live[i] ? lr : 0) * ldr;
          li[i] = rows[off + c];
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const int col = lane + 32 * q;
            // A row that is not live (the published candidate / diagonal row, which warp 1 may be writing back) is replaced
            // by row 0 as a dummy: its values are loaded, updated and DISCARDED (never stored).  compute-sanitizer's
            // racecheck reports that dummy read against warp 1's write-back when row 0 is the candidate row; building with
            // -DLA_RACECHECK_CLEAN predicates the load instead (0 hazards, profiles/r2_sanitizer_summary.txt) at the price of
            // 4 % of the factorisation time (135 vs 130 ms at n = 16384: the predicate lengthens the hot loop).
#ifdef LA_RACECHECK_CLEAN
            v[i][q] = (live[i] && q >= q0 && col > cn && col < jb) ? (EXACT ? sums[off + col] : rows[off + col]) : (T)0;
#else
            v[i][q] = (q >= q0 && col > cn && col < jb) ? (EXACT ? sums[off + col] : rows[off + col]) : (T)0;
#endif
          }
        }
#pragma unroll
        for (int i = 0; i < UR; ++i)
#pragma unroll
          for (int q = 0; q < Q; ++q)
            v[i][q] = EXACT ? add_rn(v[i][q], mul_rn(li[i], u[q])) : sub_rn(v[i][q], mul_rn(li[i], u[q]));
#pragma unroll
        for (int i = 0; i < UR; ++i) {
          if (!live[i]) continue;
          const size_t off = (size_t)(lrb + i * (PANEL_WARPS - 1)) * ldr;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const int col = lane + 32 * q;
            if (q >= q0 && col > cn && col < jb) {
              if (EXACT) sums[off + col] = v[i][q];
              else rows[off + col] = v[i][q];
            }
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- write the factored panel back ----
  for (int idx = threadIdx.x; idx < nloc * jb; idx += PANEL_THREADS) {
    const int lr = idx / jb, c = idx - lr * jb;
    A[(size_t)(row_base + lr) * ld + j0 + c] = rows[(size_t)lr * ldr + c];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. net permutation of the panel's interchanges + piv / sign bookkeeping (one warp)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(2 * MAX_NB)
lu_perm_kernel(void* ws_base, int G, int parity, int j0, int jb, uint64_t* __restrict__ piv, int* __restrict__ sign,
               int ipiv_off, int update_piv, const int* __restrict__ ipiv_src /* null: the workspace header's */) {
  // One thread per position that can change: the jb top rows and the (distinct) pivot rows below them.  The content
  // that ends up at position q is the original row reached by tracing q BACKWARDS through the jb transpositions.
  __shared__ int ipiv_s[MAX_NB];
  __shared__ int nm;
  __shared__ int flips;
  const WsView<T> ws = ws_view<T>(ws_base, G);
  MoveList* ml = &ws.hdr->moves[parity];
  const int tid = threadIdx.x;
  if (tid < jb) ipiv_s[tid] = ipiv_src ? ipiv_src[tid] : ws.hdr->ipiv[ipiv_off + tid];
  if (tid == 0) {
    nm = 0;
    flips = 0;
  }
  __syncthreads();
  int q = -1;
  if (tid < jb) {
    q = j0 + tid;
    if (ipiv_s[tid] != q) atomicAdd(&flips, 1);  // one sign flip per actual interchange (lu.rs:151)
  } else if (tid < 2 * jb) {
    const int c = tid - jb, p = ipiv_s[c];
    bool first = p >= j0 + jb;  // pivot rows inside the top block are already covered
    for (int c2 = 0; c2 < c && first; ++c2) first = ipiv_s[c2] != p;
    if (first) q = p;
  }
  int r = q;
  if (q >= 0) {
    for (int c = jb - 1; c >= 0; --c) {
      const int top = j0 + c, p = ipiv_s[c];
      r = (r == top) ? p : ((r == p) ? top : r);
    }
  }
  uint64_t pold = 0;
  int slot = -1;
  if (q >= 0 && r != q) {
    slot = atomicAdd(&nm, 1);
    ml->dst[slot] = q;
    ml->src[slot] = r;
    if (update_piv) pold = piv[r];  // `piv` is permuted exactly like a matrix column (lu.rs:147-149)
  }
  __syncthreads();
  if (slot >= 0 && update_piv) piv[q] = pold;
  if (tid == 0) {
    ml->n_moves = nm;
    if (update_piv && (flips & 1)) *sign = !*sign;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. apply the net permutation to a column range [col0, col1) (left of the panel)
// ---------------------------------------------------------------------------------------------------------------
constexpr int SWAP_W = 32;  // columns per CTA strip
// Columns [col0, col1) EXCLUDING the panel's own columns [skip0, skip1) (already interchanged inside the panel kernel).
template <typename T>
__global__ void __launch_bounds__(256) lu_swap_kernel(T* __restrict__ A, size_t ld, int col0, int col1, int skip0,
                                                      int skip1, const void* ws_base, int G, int parity) {
  extern __shared__ __align__(16) unsigned char swap_smem[];
  T(*stage)[SWAP_W] = reinterpret_cast<T(*)[SWAP_W]>(swap_smem);  // [MAX_MOVES][SWAP_W]
  __shared__ int mdst[MAX_MOVES];
  __shared__ int msrc[MAX_MOVES];
  const WsView<T> ws = ws_view<T>(const_cast<void*>(ws_base), G);
  const MoveList* ml = &ws.hdr->moves[parity];
  const int nm = ml->n_moves;
  if (nm == 0) return;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) {
    mdst[i] = ml->dst[i];
    msrc[i] = ml->src[i];
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
constexpr int INVL_THREADS = 512;
// UPPER = true inverts the upper-triangular, non-unit block U instead (used by the many-right-hand-side solve): with
// D = diag(U), U = D (I + N), N strictly upper; mirrored (i -> jb-1-i) I + N is a unit lower triangle, so the same code
// inverts it; inv(U) = inv(I + N) D^-1 is un-mirrored and column-scaled on the way out.
// Batched: block b of the grid handles the diagonal block at j0 + 128 b (clamped to `total`) and writes W + b * 128 * 128.
// MODE 2 inverts a lower-triangular NON-unit block (Cholesky's L11): L = D (I + N) with N = D^-1 * strict(L), so
// inv(L) = inv(I + N) D^-1 -- the same unit-lower inversion on the row-scaled block, columns rescaled on the way out.
// trans_out writes the transpose of the result (the B operand of `X = A21 * inv(L11)'`).
template <typename T, int MODE>
__global__ void __launch_bounds__(INVL_THREADS) lu_invl_kernel(const T* __restrict__ A, size_t ld, int j0, int jb, int total,
                                                               T* __restrict__ W /* [grid][MAX_NB][MAX_NB] */, int trans_out) {
  constexpr bool UPPER = MODE == 1;
  j0 += blockIdx.x * MAX_NB;
  jb = min(jb, total - j0);
  W += (size_t)blockIdx.x * MAX_NB * MAX_NB;
  // One padded square in shared memory holds both operands: the lower triangle (with diagonal) is W, the strictly
  // upper triangle is L11 transposed (L[i][k], k < i, lives at SQ[k][i]).  Products in flight use a separate scratch.
  extern __shared__ __align__(16) unsigned char invl_smem[];
  T(*SQ)[INVL_LD] = reinterpret_cast<T(*)[INVL_LD]>(invl_smem);
  T* scratch = reinterpret_cast<T*>(invl_smem) + (size_t)MAX_NB * INVL_LD;  // [<= 3][IB][IB]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < MAX_NB * MAX_NB; idx += INVL_THREADS) {
    const int i = idx / MAX_NB, k = idx - i * MAX_NB;
    if (k < i) {
      T v = (T)0;
      if (i < jb) {
        if (UPPER) {  // mirrored, row-scaled: (I + N)[r][c] = U[r][c] / U[r][r], r = jb-1-i < c = jb-1-k
          const int r = jb - 1 - i, c = jb - 1 - k;
          v = __ldcg(&A[(size_t)(j0 + r) * ld + j0 + c]) / __ldcg(&A[(size_t)(j0 + r) * ld + j0 + r]);
        } else {
          v = __ldcg(&A[(size_t)(j0 + i) * ld + j0 + k]);
          if (MODE == 2) v = v / __ldcg(&A[(size_t)(j0 + i) * ld + j0 + i]);
        }
      }
      SQ[k][i] = v;  // L^T into the upper triangle
      SQ[i][k] = (T)0;
    } else if (k == i) {
      SQ[i][i] = (T)0;
    }
  }
  __syncthreads();
  auto Lat = [&](int i, int k) -> T { return SQ[k][i]; };  // L[i][k], k < i
  const int nblk = (jb + IB - 1) / IB;
  // diagonal blocks: warp b, lane j solves L_bb * w = e_j by right-looking substitution held in registers
  if (warp < nblk) {
    const int o = warp * IB, j = lane;
    T w[IB];
#pragma unroll
    for (int i = 0; i < IB; ++i) w[i] = (i == j) ? (T)1 : (T)0;
#pragma unroll
    for (int k = 0; k < IB - 1; ++k) {
      const T wk = w[k];
#pragma unroll
      for (int i = k + 1; i < IB; ++i) w[i] -= Lat(o + i, o + k) * wk;  // Lat is a warp-wide broadcast
    }
#pragma unroll
    for (int i = 0; i < IB; ++i)
      if (i >= j) SQ[o + i][o + j] = w[i];  // only the lower part: the upper triangle of SQ is L^T
  }
  __syncthreads();
  for (int d = 1; d < nblk; ++d) {
    const int pairs = nblk - d;  // blocks (i, j) = (d + pr, pr)
    // phase 1: T_ij = sum_{k=j}^{i-1} L_ik * W_kj
    for (int e = tid; e < pairs * IB * IB; e += INVL_THREADS) {
      const int pr = e / (IB * IB), r = (e / IB) % IB, c = e % IB;
      const int i = d + pr, j = pr;
      T acc0 = (T)0, acc1 = (T)0;
      // W_jj is lower triangular: its entries above the diagonal are zero (that part of SQ holds L^T), so start at c
      int kk = j * IB + c;
      for (; kk + 1 < i * IB; kk += 2) {
        acc0 += Lat(i * IB + r, kk) * SQ[kk][j * IB + c];
        acc1 += Lat(i * IB + r, kk + 1) * SQ[kk + 1][j * IB + c];
      }
      if (kk < i * IB) acc0 += Lat(i * IB + r, kk) * SQ[kk][j * IB + c];
      scratch[e] = acc0 + acc1;
    }
    __syncthreads();
    // phase 2: W_ij = -W_ii * T_ij
    for (int e = tid; e < pairs * IB * IB; e += INVL_THREADS) {
      const int pr = e / (IB * IB), r = (e / IB) % IB, c = e % IB;
      const int i = d + pr, j = pr;
      T acc = (T)0;
      for (int kk = 0; kk <= r; ++kk) acc += SQ[i * IB + r][i * IB + kk] * scratch[pr * IB * IB + kk * IB + c];
      SQ[i * IB + r][j * IB + c] = -acc;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < MAX_NB * MAX_NB; idx += INVL_THREADS) {
    const int i = idx / MAX_NB, k = idx - i * MAX_NB;
    T v = (T)0;
    if (UPPER) {  // W[r][c] = inv(I + N)[r][c] / U[c][c], c >= r
      if (k >= i && k < jb) v = SQ[jb - 1 - i][jb - 1 - k] / __ldcg(&A[(size_t)(j0 + k) * ld + j0 + k]);
    } else if (k <= i && i < jb) {
      v = (MODE == 2) ? ((k == i ? (T)1 : SQ[i][k]) / __ldcg(&A[(size_t)(j0 + k) * ld + j0 + k])) : SQ[i][k];
    }
    W[trans_out ? k * MAX_NB + i : idx] = v;
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
lu_swap_trsm_kernel(T* __restrict__ A, size_t ld, int j0, int jb, int col0, int col1, const void* ws_base, int G,
                    int parity) {
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
  const MoveList* ml = &ws.hdr->moves[parity];
  const int nm = ml->n_moves;
  for (int i = threadIdx.x; i < nm; i += blockDim.x) {
    const int d = ml->dst[i], s = ml->src[i];
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

// ---------------------------------------------------------------------------------------------------------------
// 4b. look-ahead head: row interchanges + U12 = L11^-1 * A12 for the NEXT panel's columns only (<= MAX_NB of them), by
//     substitution with L11 resident in shared memory.  It sits on the critical chain (panel -> head -> panel), so it
//     must not wait for inv(L11) (which only the bulk stream needs): one warp owns two columns, keeps them in registers
//     (lane <-> rows lane, lane+32, ...), broadcasts x_k by shuffle and streams row k of L^T from shared memory --
//     128 dependent steps of ~40 cycles instead of a 60 us triangular inversion plus a GEMM launch.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HEAD_COLS = 16;  // columns per CTA (8 warps x 2)
constexpr int HEAD_LD = MAX_NB + 1;
template <typename T>
__global__ void __launch_bounds__(256)
lu_head_kernel(T* __restrict__ A, size_t ld, int j0, int jb, int col0, int col1, const void* ws_base, int G, int parity,
               const T* __restrict__ L11 /* the panel's diagonal block (the multi-device driver keeps a copy) */, size_t ldl) {
  extern __shared__ __align__(16) unsigned char head_smem[];
  T(*LT)[HEAD_LD] = reinterpret_cast<T(*)[HEAD_LD]>(head_smem);                  // [MAX_NB]: LT[k][i] = L11[i][k]
  T(*X)[HEAD_COLS + 1] = reinterpret_cast<T(*)[HEAD_COLS + 1]>(LT + MAX_NB);     // [MAX_NB]: top jb rows of the strip
  T(*stage)[HEAD_COLS] = reinterpret_cast<T(*)[HEAD_COLS]>(X + MAX_NB);          // [MAX_NB]: rows leaving the top block
  __shared__ int top_src[MAX_NB];
  __shared__ int out_dst[MAX_NB];
  __shared__ int out_src[MAX_NB];
  __shared__ int n_out;
  const WsView<T> ws = ws_view<T>(const_cast<void*>(ws_base), G);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tx = tid & (HEAD_COLS - 1), ty = tid / HEAD_COLS;  // 16 columns x 16 rows per pass
  const int col = col0 + blockIdx.x * HEAD_COLS + tx;
  const bool ok = col < col1;

  for (int i = tid; i < MAX_NB; i += blockDim.x) top_src[i] = j0 + i;
  if (tid == 0) n_out = 0;
  __syncthreads();
  const MoveList* ml = &ws.hdr->moves[parity];
  const int nm = ml->n_moves;
  for (int i = tid; i < nm; i += blockDim.x) {
    const int d = ml->dst[i], s = ml->src[i];
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
  for (int i = ty; i < MAX_NB; i += 256 / HEAD_COLS)
    X[i][tx] = (ok && i < jb) ? __ldcg(&A[(size_t)top_src[i] * ld + col]) : (T)0;
  for (int i = ty; i < no; i += 256 / HEAD_COLS) stage[i][tx] = ok ? __ldcg(&A[(size_t)out_src[i] * ld + col]) : (T)0;
  // L11 (strictly lower part), transposed so that step k reads one contiguous shared-memory row
  for (int idx = tid; idx < MAX_NB * MAX_NB; idx += 256) {
    const int i = idx / MAX_NB, k = idx - i * MAX_NB;
    LT[k][i] = (i < jb && k < i) ? __ldcg(&L11[(size_t)i * ldl + k]) : (T)0;
  }
  __syncthreads();
  for (int i = ty; i < no; i += 256 / HEAD_COLS)
    if (ok) A[(size_t)out_dst[i] * ld + col] = stage[i][tx];

  // forward substitution (unit diagonal), k ascending
  constexpr int Q = MAX_NB / 32;
  T x[Q][2];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    x[q][0] = X[lane + 32 * q][2 * warp];
    x[q][1] = X[lane + 32 * q][2 * warp + 1];
  }
#pragma unroll
  for (int kq = 0; kq < Q; ++kq) {
    if (kq * 32 < jb) {
#pragma unroll 8
      for (int kl = 0; kl < 32; ++kl) {
        const int k = kq * 32 + kl;
        const T xk0 = __shfl_sync(0xffffffffu, x[kq][0], kl);
        const T xk1 = __shfl_sync(0xffffffffu, x[kq][1], kl);
#pragma unroll
        for (int q = kq; q < Q; ++q) {
          const T l = LT[k][lane + 32 * q];  // zero on and above the diagonal and beyond jb
          if (q > kq || lane > kl) {         // rows strictly below k (predicated: 0 * inf must not make a NaN)
            x[q][0] = sub_rn(x[q][0], mul_rn(l, xk0));
            x[q][1] = sub_rn(x[q][1], mul_rn(l, xk1));
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    X[lane + 32 * q][2 * warp] = x[q][0];
    X[lane + 32 * q][2 * warp + 1] = x[q][1];
  }
  __syncthreads();
  for (int i = ty; i < jb; i += 256 / HEAD_COLS)
    if (ok) A[(size_t)(j0 + i) * ld + col] = X[i][tx];
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
// Per host thread and device: the chain stream (highest priority) and the events that fence it against the bulk stream.
struct LuSide {
  cudaStream_t sp = nullptr, sw = nullptr;
  cudaEvent_t e_in = nullptr, e_head = nullptr, e_bulk = nullptr, e_panel = nullptr, e_w = nullptr, e_u12 = nullptr;
};
// Rows per panel CTA.  Fewer, fuller CTAs leave more SMs wholly to the bulk GEMMs (a panel CTA takes half the register
// file, so a GEMM runs at half occupancy next to it) and the in-panel look-ahead hides their longer update; near the end
// the bulk is negligible and many small CTAs give the shortest column time.
struct PanelShape {
  int rpc, G;
  size_t smem;
};
template <typename T>
PanelShape panel_shape(int R, int jb, int sms, bool exact) {
  static const int rpc_min = getenv("LA_LU_RPC_MIN") ? atoi(getenv("LA_LU_RPC_MIN")) : 200;  // tuning knobs
  static const int rpc_div = getenv("LA_LU_RPC_DIV") ? atoi(getenv("LA_LU_RPC_DIV")) : 64;
  int rpc = (R + sms - 1) / sms;
  int want = R / rpc_div;
  int rpc_cap = rpc_min;
  if (jb <= MAX_NB / 2 && !exact) {  // a 64-wide half panel holds twice the rows in the same shared memory
    const int fit = (int)(PANEL_SMEM_BUDGET / ((size_t)(jb | 1) * sizeof(T))) / 8 * 8;
    rpc_cap = 2 * rpc_min < fit ? 2 * rpc_min : fit;
    want = 2 * want;
  }
  if (want > rpc_cap) want = rpc_cap;
  // never more rows than the shared-memory budget holds (the bit-exact panel keeps a second array of deferred sums:
  // a tall single-panel matrix such as 8000 x 128 would otherwise ask for 125 rows x 129 x 16 B = 258 KB)
  const int fit_rows = (int)(PANEL_SMEM_BUDGET / ((size_t)(jb | 1) * sizeof(T) * (exact ? 2 : 1)));
  if (want > fit_rows) want = fit_rows;
  if (want < 8) want = 8;  // at least one row per warp
  if (rpc < want) rpc = want;
  PanelShape ps;
  ps.rpc = rpc;
  ps.G = (R + rpc - 1) / rpc;
  ps.smem = (size_t)rpc * (jb | 1) * sizeof(T) * (exact ? 2 : 1);
  return ps;
}

int lu_side(int device, LuSide** out) {
  static thread_local LuSide side[64];
  LA_REQUIRE(device >= 0 && device < 64, "device ordinal out of range");
  LuSide& s = side[device];
  if (!s.sp) {
    int lo = 0, hi = 0;
    LA_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&s.sp, cudaStreamNonBlocking, hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&s.sw, cudaStreamNonBlocking, hi));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_panel, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_w, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_in, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_head, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_bulk, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_u12, cudaEventDisableTiming));
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
  int nb = (int)(PANEL_SMEM_BUDGET / ((size_t)rpc_first * sizeof(T))) - 1;  // rows are padded to an odd stride
  nb = nb < 0 ? 0 : nb / 16 * 16;
  if (nb > MAX_NB) nb = MAX_NB;
  if (nb < 16)
    return fail(LA_ERR_UNSUPPORTED, "la_lu_factor: %d rows exceed the shared-memory panel capacity of %d SMs", M, sms);
  // single-panel factorisations run the bit-exact (deferred subtraction) panel when twice the panel fits
  const bool exact = kmin <= nb && (size_t)2 * rpc_first * (kmin | 1) * sizeof(T) <= PANEL_SMEM_BUDGET;
  // multi-panel fp64 on TMA-addressable storage: look-ahead pipeline with U12 = inv(L11) * A12 on the DMMA GEMM
  static const int dbg = getenv("LA_LU_DEBUG") ? atoi(getenv("LA_LU_DEBUG")) : 0;  // 1: one stream, 2: plain loop
  const bool fast = std::is_same<T, double>::value && kmin > nb && (N % 2 == 0) && ((uintptr_t)LU % 16 == 0) &&
                    (kmin % 2 == 0 || kmin == N) && dbg != 2;
  // multi-panel fp32: the same chain / bulk look-ahead, but exact fp32 arithmetic throughout (the parity bar is an
  // IDENTICAL pivot sequence to the fp32 reference, which TF32 tensor tiles would not give): U12 by substitution for
  // all columns, trailing update on the CUDA-core GEMM
  const bool fast32 = std::is_same<T, float>::value && kmin > nb && dbg != 2;

  void* ws_base = nullptr;
  LA_TRY(scratch_get(ctx->device, 8, ws_bytes<T>(sms), &ws_base));
  void* w_base = nullptr;
  LA_TRY(scratch_get(ctx->device, 11, sizeof(T) * 2 * MAX_NB * MAX_NB, &w_base));
  T* Wbuf[2] = {(T*)w_base, (T*)w_base + MAX_NB * MAX_NB};

  const int SWAP_SMEM = (int)(sizeof(T) * MAX_MOVES * SWAP_W);
  const int TRSM_SMEM = (int)(sizeof(T) * 3 * MAX_NB * TRSM_W);
  const int INVL_SMEM = (int)(sizeof(T) * ((size_t)MAX_NB * INVL_LD + 3 * IB * IB));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_panel_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)PANEL_SMEM_BUDGET + 2048));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_panel_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)PANEL_SMEM_BUDGET + 2048));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_swap_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWAP_SMEM));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_swap_trsm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM));
  LA_CUDA_TRY(cudaFuncSetAttribute(lu_invl_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, INVL_SMEM));

  lu_init_piv_kernel<T><<<(M + 255) / 256, 256, 0, st>>>(piv_dev, M, sign_dev);
  LA_CUDA_TRY(cudaGetLastError());
  // all tags invalid (0) before the first panel; epochs then count up within this factorisation (< 2^24 panels)
  LA_CUDA_TRY(cudaMemsetAsync((char*)ws_base + ws_hdr_bytes<T>(), 0, ws_bytes<T>(sms) - ws_hdr_bytes<T>(), st));
  unsigned epoch = 0;

  int G_cur = 1;
  // panel factorisation of columns [j0, j0+jb) on stream s (ipiv lands in the workspace header)
  auto launch_panel = [&](int j0, int jb, cudaStream_t s, int ipiv_off = 0) -> int {
    const PanelShape ps = panel_shape<T>(M - j0, jb, sms, exact);
    const int rpc = ps.rpc, G = ps.G;
    const size_t smem = ps.smem;
    T* a = LU;
    size_t ld = n;
    int mm = M, jj0 = j0, jjb = jb, rr = rpc;
    void* wsb = ws_base;
    unsigned ep = ++epoch;  // distinguishes this panel's flagged words from every earlier panel's
    int ioff = ipiv_off;
    void* args[] = {&a, &ld, &mm, &jj0, &jjb, &rr, &wsb, &ep, &ioff};
    const void* fn = exact ? (const void*)lu_panel_kernel<T, true> : (const void*)lu_panel_kernel<T, false>;
    LA_CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(PANEL_THREADS), args, smem, s));
    G_cur = G;
    return LA_OK;
  };
  // net permutation + piv / sign bookkeeping of the panel just factored
  auto launch_perm = [&](int j0, int jb, int parity, cudaStream_t s, int ipiv_off = 0, int update_piv = 1) -> int {
    lu_perm_kernel<T><<<1, 2 * MAX_NB, 0, s>>>(ws_base, G_cur, parity, j0, jb, piv_dev, sign_dev, ipiv_off, update_piv,
                                               nullptr);
    LA_CUDA_TRY(cudaGetLastError());
    return LA_OK;
  };
  // row interchanges of columns [col0, col1) minus [skip0, skip1)
  auto launch_swap = [&](int col0, int col1, int skip0, int skip1, int parity, cudaStream_t s) -> int {
    const int ncols = (col1 - col0) - (skip1 - skip0);
    if (ncols <= 0) return LA_OK;
    lu_swap_kernel<T><<<(ncols + SWAP_W - 1) / SWAP_W, 256, SWAP_SMEM, s>>>(LU, n, col0, col1, skip0, skip1, ws_base, G_cur,
                                                                          parity);
    LA_CUDA_TRY(cudaGetLastError());
    return LA_OK;
  };

  if (!fast && !fast32) {
    // ---- plain right-looking loop (single panel / odd leading dimension): substitution TRSM kernel ----
    for (int j0 = 0; j0 < kmin; j0 += nb) {
      const int jb = (kmin - j0 < nb) ? (kmin - j0) : nb;
      LA_TRY(launch_panel(j0, jb, st));
      LA_TRY(launch_perm(j0, jb, 0, st));
      LA_TRY(launch_swap(0, j0, j0, j0, 0, st));
      const int c1 = j0 + jb;
      if (c1 < N) {
        lu_swap_trsm_kernel<T><<<(N - c1 + TRSM_W - 1) / TRSM_W, 256, TRSM_SMEM, st>>>(LU, n, j0, jb, c1, N, ws_base,
                                                                                     G_cur, 0);
        LA_CUDA_TRY(cudaGetLastError());
        if (c1 < M)
          LA_TRY(gemm_dev<T>(LU + (size_t)c1 * n + j0, n, LU + (size_t)j0 * n + c1, n, LU + (size_t)c1 * n + c1, n,
                             (size_t)(M - c1), (size_t)jb, (size_t)(N - c1), LA_GEMM_SUB, st));
      }
    }
    return LA_OK;
  }

  // ---- look-ahead pipeline (fp64) -------------------------------------------------------------------------------
  // chain stream sp (highest priority), per panel i:  perm(i) -> [wait bulk(i-1)] -> head: row interchanges + U12 by
  //   substitution for the NEXT panel's columns -> their trailing update (DMMA) -> panel(i+1)
  // side stream sw, per panel i:                      W(i) = inv(L11(i))  (only the bulk needs it)
  // bulk stream st (the caller's), per panel i:       [wait perm(i), W(i)] -> row interchanges of all other columns ->
  //   U12 = W * A12 and A22 -= L21 * U12 for the columns right of the next panel
  // Early on the bulk GEMMs are the critical path and run back to back (nothing they wait for depends on bulk(i-1));
  // late the chain is, and it carries nothing the bulk could do.  The cooperative panel CTAs are latency-bound and
  // co-reside with the DMMA CTAs.  W and the move lists are double-buffered by panel parity.
  if constexpr (std::is_same<T, double>::value) {
    LuSide* side;
    LA_TRY(lu_side(ctx->device, &side));
    cudaStream_t sp = dbg == 1 ? st : side->sp;
    cudaStream_t sw = dbg == 1 ? st : side->sw;
    double* A = LU;
    const size_t ld = n;
    const int HEAD_SMEM = (int)(sizeof(T) * ((size_t)MAX_NB * HEAD_LD + (size_t)MAX_NB * (HEAD_COLS + 1) +
                                             (size_t)MAX_NB * HEAD_COLS));
    LA_CUDA_TRY(cudaFuncSetAttribute(lu_head_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM));
    // LA_LU_TRACE=<file>: per-panel timeline from timing events (diagnostic; perturbs the run by the event records)
    static const char* trace_path = getenv("LA_LU_TRACE");
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t s) {
      if (!trace_path) return;
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, s);
      tev.push_back(e);
    };
    LA_CUDA_TRY(cudaEventRecord(side->e_in, st));
    LA_CUDA_TRY(cudaStreamWaitEvent(sp, side->e_in, 0));
    mark(sp);  // t0
    LA_TRY(launch_panel(0, nb, sp));
    // Grouped bulk updates: while the trailing matrix is tall, `group` consecutive panels share ONE trailing GEMM of depth
    // K = group * nb (half the C read-modify-write traffic per flop at group = 2; the K = 128 update is bound by it).
    // Steps that defer ("S") do only the interchanges and their U12 rows in the bulk -- U12_s = W_s (A12_s - L[s][g..s)
    // U[g..s)) -- and the step that closes the group adds A22 -= L[:, g..c1) U[g..c1, :] for all columns except the
    // next-next panel's, which the chain brings up to date itself (same K range, 128 columns wide, high priority) as
    // soon as the U12 rows exist: the chain never waits for a bulk GEMM.
    static const int group_max = getenv("LA_LU_GROUP") ? atoi(getenv("LA_LU_GROUP")) : 2;            // 1 = no grouping
    static const int group_rows = getenv("LA_LU_GROUP_ROWS") ? atoi(getenv("LA_LU_GROUP_ROWS")) : 6144;
    int gstart = 0;        // first column of the open group (panels whose bulk trailing update is deferred)
    int gcount = 0;        // panels deferred so far in the open group
    int strip_k0 = 0;      // the next panel's columns lack the updates of panels [strip_k0, j0) (== j0: none)
    int prev_mode = 0;     // 0: previous step updated everything (or first step), 1: it deferred, 2: it closed a group
    int it = 0;
    for (int j0 = 0; j0 < kmin; j0 += nb, ++it) {
      const int jb = (kmin - j0 < nb) ? (kmin - j0) : nb;
      const int c1 = j0 + jb;
      const int parity = it & 1;
      const bool has_next = c1 < kmin;
      const int nb2 = has_next ? ((kmin - c1 < nb) ? (kmin - c1) : nb) : 0;
      const int c2 = c1 + nb2;
      if (gcount == 0) gstart = j0;
      if (prev_mode == 0) strip_k0 = j0;
      // defer this panel's trailing update?  Needs a full next panel and a full panel after it (whose chain step does the
      // catch-up), a tall trailing matrix, and room in the group.
      const bool defer = group_max > 1 && gcount + 1 < group_max && jb == nb && nb2 == nb && c2 + nb <= kmin &&
                         M - c1 >= group_rows;
      const int kdef = j0 - gstart;     // depth of the updates the columns right of c2 still lack (0: none)
      const int kstrip = j0 - strip_k0;  // depth of the updates the next panel's columns still lack
      // width of the next-next panel: after a group closes, its columns are left to the chain
      const int w_strip = (kdef > 0 && !defer && c2 < kmin) ? ((kmin - c2 < nb) ? (kmin - c2) : nb) : 0;
      double* W = Wbuf[parity];
      const double* L21 = A + (size_t)c1 * ld + j0;
      auto trailing = [&](int cb, int ce, cudaStream_t s) -> int {  // A22 -= L21 * U12 for columns [cb, ce)
        if (ce <= cb || c1 >= M) return LA_OK;
        return gemm_f64_tensor(L21, ld, A + (size_t)j0 * ld + cb, ld, A + (size_t)c1 * ld + cb, ld, (size_t)(M - c1),
                               (size_t)jb, (size_t)(ce - cb), LA_GEMM_SUB, s);
      };
      // ---- side: W = inv(L11) as soon as the panel is done ----
      mark(sp);  // [0] panel(i) done
      const bool need_w = c2 < N;
      if (need_w) {
        LA_CUDA_TRY(cudaEventRecord(side->e_panel, sp));
        LA_CUDA_TRY(cudaStreamWaitEvent(sw, side->e_panel, 0));
        lu_invl_kernel<T, 0><<<1, INVL_THREADS, INVL_SMEM, sw>>>(LU, n, j0, jb, j0 + jb, W, 0);
        LA_CUDA_TRY(cudaGetLastError());
        LA_CUDA_TRY(cudaEventRecord(side->e_w, sw));
      }
      // ---- chain ----
      if (has_next && kstrip > 0) {
        // catch-up: the next panel's columns lack the updates of the panels [strip_k0, j0); their U rows were
        // finished by the previous bulk step (e_bulk after a deferring step, e_u12 after a closing one).  This GEMM reads
        // L columns left of j0, which bulk(i) interchanges: it must precede e_head.
        LA_CUDA_TRY(cudaStreamWaitEvent(sp, prev_mode == 1 ? side->e_bulk : side->e_u12, 0));
        LA_TRY(gemm_f64_tensor(A + (size_t)j0 * ld + strip_k0, ld, A + (size_t)strip_k0 * ld + c1, ld,
                               A + (size_t)j0 * ld + c1, ld, (size_t)(M - j0), (size_t)kstrip, (size_t)(c2 - c1), LA_GEMM_SUB,
                               sp));
      }
      LA_TRY(launch_perm(j0, jb, parity, sp));
      LA_CUDA_TRY(cudaEventRecord(side->e_head, sp));
      if (has_next && it > 0 && prev_mode == 0)
        LA_CUDA_TRY(cudaStreamWaitEvent(sp, side->e_bulk, 0));  // bulk(i-1) updated columns >= c1
      mark(sp);  // [1] perm done and bulk(i-1) done
      if (has_next) {
        lu_head_kernel<T><<<(nb2 + HEAD_COLS - 1) / HEAD_COLS, 256, HEAD_SMEM, sp>>>(LU, n, j0, jb, c1, c2, ws_base, G_cur,
                                                                                    parity, LU + (size_t)j0 * n + j0, n);
        LA_CUDA_TRY(cudaGetLastError());
        mark(sp);  // [2] next panel's columns interchanged, U12 solved
        LA_TRY(trailing(c1, c2, sp));
        mark(sp);  // [3] next panel's columns updated
        // While the trailing matrix is tall the bulk GEMM is the critical path and the panel costs it SMs (its CTAs are
        // exclusive on their SMs because of their shared memory): factor the panel as two 64-wide halves -- half the
        // shared memory per row, half the SMs -- with the half-panel head/update in between.  The net effect on the 128
        // columns is that of one panel; the 128 pivots land in ipiv[0..128) for the usual perm/head/bulk of the next step.
        static const int split_rows = getenv("LA_LU_SPLIT_ROWS") ? atoi(getenv("LA_LU_SPLIT_ROWS")) : 12288;  // 0 = never
        if (split_rows > 0 && nb2 == MAX_NB && M - c1 >= split_rows) {
          const int h = MAX_NB / 2, cm = c1 + h;
          LA_TRY(launch_panel(c1, h, sp, 0));
          LA_TRY(launch_perm(c1, h, 2, sp, 0, 0));
          lu_head_kernel<T><<<(h + HEAD_COLS - 1) / HEAD_COLS, 256, HEAD_SMEM, sp>>>(LU, n, c1, h, cm, c2, ws_base, G_cur, 2,
                                                                                   LU + (size_t)c1 * n + c1, n);
          LA_CUDA_TRY(cudaGetLastError());
          if (cm < M)
            LA_TRY(gemm_f64_tensor(A + (size_t)cm * ld + c1, ld, A + (size_t)c1 * ld + cm, ld, A + (size_t)cm * ld + cm, ld,
                                   (size_t)(M - cm), (size_t)h, (size_t)h, LA_GEMM_SUB, sp));
          LA_TRY(launch_panel(cm, h, sp, h));
          LA_TRY(launch_perm(cm, h, 3, sp, h, 0));
          LA_TRY(launch_swap(c1, cm, cm, cm, 3, sp));  // the second half's interchanges on the first half's columns
        } else {
          LA_TRY(launch_panel(c1, nb2, sp));
        }
      } else {
        mark(sp);
        mark(sp);
      }
      // ---- bulk ----
      LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_head, 0));
      if (need_w) LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_w, 0));
      mark(st);  // [4] bulk(i) start
      LA_TRY(launch_swap(0, N, j0, c2, parity, st));  // everything but this panel's and the next panel's columns
      mark(st);  // [5] bulk swaps done
      if (need_w) {
        double* U12 = A + (size_t)j0 * ld + c2;
        if (kdef > 0)  // this panel's rows of A12 first receive the deferred panels' updates
          LA_TRY(gemm_f64_tensor(A + (size_t)j0 * ld + gstart, ld, A + (size_t)gstart * ld + c2, ld, U12, ld, (size_t)jb,
                                 (size_t)kdef, (size_t)(N - c2), LA_GEMM_SUB, st));
        LA_TRY(gemm_f64_tensor(W, MAX_NB, U12, ld, U12, ld, (size_t)jb, (size_t)jb, (size_t)(N - c2), LA_GEMM_ASSIGN,
                               st));  // in place: one tile row, every CTA reads its whole column block first
        if (!defer) {
          if (kdef > 0) {  // close the group: one GEMM of depth c1 - gstart, the chain's strip excluded
            LA_CUDA_TRY(cudaEventRecord(side->e_u12, st));
            const int cb = c2 + w_strip;
            if (cb < N && c1 < M)
              LA_TRY(gemm_f64_tensor(A + (size_t)c1 * ld + gstart, ld, A + (size_t)gstart * ld + cb, ld,
                                     A + (size_t)c1 * ld + cb, ld, (size_t)(M - c1), (size_t)(c1 - gstart), (size_t)(N - cb),
                                     LA_GEMM_SUB, st));
          } else {
            LA_TRY(trailing(c2, N, st));
          }
        }
      }
      LA_CUDA_TRY(cudaEventRecord(side->e_bulk, st));
      mark(st);  // [6] bulk(i) done
      if (defer) {  // the next panel's columns received this panel's update from the chain, the next-next panel's did not
        ++gcount;
        prev_mode = 1;
        strip_k0 = gstart;
      } else {
        // after a closed group the strip [c2, c2 + w_strip) still lacks [gstart, c1): the next chain step catches up
        prev_mode = (kdef > 0 && w_strip > 0) ? 2 : 0;
        strip_k0 = gstart;
        gcount = 0;
      }
    }
    if (trace_path) {
      LA_CUDA_TRY(cudaStreamSynchronize(sp));
      LA_CUDA_TRY(cudaStreamSynchronize(st));
      if (FILE* f = fopen(trace_path, "w")) {
        fprintf(f, "panel,panel_done,perm_and_bulk_prev,head_solved,head_updated,bulk_start,bulk_swapped,bulk_done\n");
        for (size_t p = 0; p * 7 + 7 < tev.size(); ++p) {
          fprintf(f, "%zu", p);
          for (int q = 0; q < 7; ++q) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, tev[0], tev[1 + p * 7 + q]);
            fprintf(f, ",%.4f", ms);
          }
          fprintf(f, "\n");
        }
        fclose(f);
      }
      for (cudaEvent_t e : tev) cudaEventDestroy(e);
    }
    // the caller's stream must also cover the tail of the chain stream
    LA_CUDA_TRY(cudaEventRecord(side->e_head, sp));
    LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_head, 0));
  } else {
    // ---- look-ahead pipeline (fp32): chain = perm -> [wait bulk(i-1)] -> head (next panel's columns) -> their trailing
    //      update -> panel(i+1); bulk = interchanges left, head kernel (interchanges + U12) and trailing update right ----
    LuSide* side;
    LA_TRY(lu_side(ctx->device, &side));
    cudaStream_t sp = dbg == 1 ? st : side->sp;
    const size_t ld = n;
    const int HEAD_SMEM = (int)(sizeof(T) * ((size_t)MAX_NB * HEAD_LD + (size_t)MAX_NB * (HEAD_COLS + 1) +
                                             (size_t)MAX_NB * HEAD_COLS));
    LA_CUDA_TRY(cudaFuncSetAttribute(lu_head_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM));
    LA_CUDA_TRY(cudaEventRecord(side->e_in, st));
    LA_CUDA_TRY(cudaStreamWaitEvent(sp, side->e_in, 0));
    LA_TRY(launch_panel(0, nb, sp));
    int it = 0;
    for (int j0 = 0; j0 < kmin; j0 += nb, ++it) {
      const int jb = (kmin - j0 < nb) ? (kmin - j0) : nb;
      const int c1 = j0 + jb;
      const int parity = it & 1;
      const bool has_next = c1 < kmin;
      const int nb2 = has_next ? ((kmin - c1 < nb) ? (kmin - c1) : nb) : 0;
      const int c2 = c1 + nb2;
      auto head_and_update = [&](int cb, int ce, cudaStream_t s) -> int {  // columns [cb, ce)
        if (ce <= cb) return LA_OK;
        lu_head_kernel<T><<<(ce - cb + HEAD_COLS - 1) / HEAD_COLS, 256, HEAD_SMEM, s>>>(LU, n, j0, jb, cb, ce, ws_base, G_cur,
                                                                                    parity, LU + (size_t)j0 * n + j0, n);
        LA_CUDA_TRY(cudaGetLastError());
        if (c1 < M)
          LA_TRY(gemm_dev<T>(LU + (size_t)c1 * ld + j0, ld, LU + (size_t)j0 * ld + cb, ld, LU + (size_t)c1 * ld + cb, ld,
                             (size_t)(M - c1), (size_t)jb, (size_t)(ce - cb), LA_GEMM_SUB, s));
        return LA_OK;
      };
      // ---- chain ----
      LA_TRY(launch_perm(j0, jb, parity, sp));
      LA_CUDA_TRY(cudaEventRecord(side->e_head, sp));
      if (has_next) {
        if (it > 0) LA_CUDA_TRY(cudaStreamWaitEvent(sp, side->e_bulk, 0));  // bulk(i-1) updated columns >= c1
        LA_TRY(head_and_update(c1, c2, sp));
        LA_TRY(launch_panel(c1, nb2, sp));
      }
      // ---- bulk ----
      LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_head, 0));
      LA_TRY(launch_swap(0, j0, j0, j0, parity, st));
      LA_TRY(head_and_update(c2, N, st));
      LA_CUDA_TRY(cudaEventRecord(side->e_bulk, st));
    }
    LA_CUDA_TRY(cudaEventRecord(side->e_head, sp));
    LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_head, 0));
  }
  return LA_OK;
}
// Batched inverses of the 128 x 128 diagonal blocks [first_block, first_block + nblocks) of an order-n triangular factor:
// mode 0 = unit lower (L of a packed LU), 1 = upper non-unit (U of a packed LU, or L' of a Cholesky factor),
// 2 = lower non-unit (Cholesky's L).  One launch; W[b] is zero-padded; trans_out stores the transposes.
template <typename T>
int tri_block_inverses(const T* M, size_t n, int mode, int first_block, int nblocks, T* W, int trans_out, cudaStream_t st) {
  const int INVL_SMEM = (int)(sizeof(T) * ((size_t)MAX_NB * INVL_LD + 3 * IB * IB));
  LA_REQUIRE(mode >= 0 && mode <= 2 && nblocks > 0, "tri_block_inverses: bad arguments");
  const void* fn = mode == 0 ? (const void*)lu_invl_kernel<T, 0>
                             : (mode == 1 ? (const void*)lu_invl_kernel<T, 1> : (const void*)lu_invl_kernel<T, 2>);
  LA_CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, INVL_SMEM));
  const int j0 = first_block * MAX_NB;
  if (mode == 0) lu_invl_kernel<T, 0><<<nblocks, INVL_THREADS, INVL_SMEM, st>>>(M, n, j0, MAX_NB, (int)n, W, trans_out);
  if (mode == 1) lu_invl_kernel<T, 1><<<nblocks, INVL_THREADS, INVL_SMEM, st>>>(M, n, j0, MAX_NB, (int)n, W, trans_out);
  if (mode == 2) lu_invl_kernel<T, 2><<<nblocks, INVL_THREADS, INVL_SMEM, st>>>(M, n, j0, MAX_NB, (int)n, W, trans_out);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template int tri_block_inverses<double>(const double*, size_t, int, int, int, double*, int, cudaStream_t);
template int tri_block_inverses<float>(const float*, size_t, int, int, int, float*, int, cudaStream_t);

// WL[b] = inv(L_bb) (unit lower), WU[b] = inv(U_bb) for all diagonal blocks of a packed LU: the many-right-hand-side
// solve and the sweep kernels (lu_solve.cu) then need products only.
int lu_diag_block_inverses(const double* LU, size_t n, double* WL, double* WU, cudaStream_t st) {
  const int G = (int)((n + MAX_NB - 1) / MAX_NB);
  LA_TRY(tri_block_inverses<double>(LU, n, 0, 0, G, WL, 0, st));
  LA_TRY(tri_block_inverses<double>(LU, n, 1, 0, G, WU, 0, st));
  return LA_OK;
}

template int lu_factor_dev<double>(double*, size_t, size_t, uint64_t*, int*, cudaStream_t);
template int lu_factor_dev<float>(float*, size_t, size_t, uint64_t*, int*, cudaStream_t);

// ---------------------------------------------------------------------------------------------------------------
// Multi-device LU: LUDecomposition::new (lu.rs:104-168) of one n x n fp64 matrix across several
// GPUs driven by one host thread.
//
// Layout: 128-column blocks (narrower when a 128-wide panel of n rows does not fit one GPU's shared memory) dealt round-robin (block b lives on device b mod G, all n rows of it), so every device keeps
// a share of the trailing matrix until the end.  Panel k is factored by its owner with the single-device panel kernel;
// the factored block column (L11 over L21), and its 128 pivot rows, are copied into a ring slot on every device (peer
// copies queued on the owner's chain stream, the next owner first), after which each device is on its own: it folds the
// pivots into its own copy of `piv`, interchanges the rows of its columns, inverts L11 on a side stream, forms its U12
// rows (W * A12, DMMA GEMM) and updates its part of the trailing matrix (DMMA GEMM, L21 from the ring slot).  No
// collective and no host round trip: streams, events and peer copies only.
// Look-ahead as on one device: the owner of panel k+1 brings that panel's columns up to date first (substitution head +
// a 128-column GEMM on its high-priority chain stream) and factors it while every device's bulk update of step k runs.
// The chain (panel -> copy -> head -> update -> panel) hops from device to device and bounds the factorisation from
// below at ~0.6 ms per panel; the bulk work is what the devices share.
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr int MGLU_RING = 4;  // panels in flight: ring slots of L / W / pivots, move lists (PanelWs::moves has four)
constexpr int MGLU_MAX_DEV = 16;

struct LuMgDev {
  int device = -1, sms = 0;
  int ncols = 0;   // local columns
  size_t ld = 0;   // local leading dimension (even: TMA rows are 16-byte aligned)
  double* A = nullptr;
  double* Lring = nullptr;  // [RING][n][128]
  double* Wring = nullptr;  // [RING][128][128]
  int* ipiv_ring = nullptr; // [RING][128]
  void* ws = nullptr;       // panel workspace (exchange area + move lists)
  uint64_t* piv = nullptr;  // every device keeps the whole permutation
  int* sign = nullptr;
  cudaStream_t sp = nullptr, sw = nullptr, st = nullptr;  // chain (high priority), side, bulk
  cudaEvent_t e_arr[MGLU_MAX_DEV][MGLU_RING] = {};  // as owner: slot s has landed on device q
  cudaEvent_t e_sent[MGLU_RING] = {};               // as owner: every copy of slot s is done
  cudaEvent_t e_perm[MGLU_RING] = {}, e_w[MGLU_RING] = {}, e_bulk[MGLU_RING] = {};
  cudaEvent_t e_in = nullptr, e_tail = nullptr;
  unsigned epoch = 0;
};

struct DevSwitch {  // restores the caller's current device
  int prev = -1;
  DevSwitch() { cudaGetDevice(&prev); }
  ~DevSwitch() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

__global__ void lu_mg_fill_kernel(double* __restrict__ A, size_t ld, int n, int ncols, int d, int G, int nb, uint64_t seed) {
  const size_t total = (size_t)n * ncols;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / ncols;
    const int lc = (int)(e - r * ncols);
    const size_t gc = ((size_t)(lc / nb) * G + d) * nb + lc % nb;
    A[r * ld + lc] = (double)(hash64(seed, r * (size_t)n + gc) >> 11) * 0x1.0p-53;
  }
}
}  // namespace
}  // namespace la

struct la_lu_mg {
  int ndev = 0;
  size_t n = 0;
  int nb = 0;    // block-column (= panel) width: 128 unless a 128-wide panel of n rows exceeds one GPU's shared memory
  int nblk = 0;
  la::LuMgDev dev[la::MGLU_MAX_DEV];
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // on dev[0]: the factorisation as the devices saw it
  float last_ms = -1.f;
};

namespace la {
namespace {

// Width of a block column (= panel) for n rows on devices with `sms` SMs: 128 unless the tallest panel would not fit the
// shared memory of one device (cf. lu_factor_dev); < 16 means "does not fit at all".
int mglu_block_width(size_t n, int sms) {
  int rpc_first = (int)((n + sms - 1) / sms);
  if (rpc_first < 8) rpc_first = 8;
  int nb = (int)(PANEL_SMEM_BUDGET / ((size_t)rpc_first * sizeof(double))) - 1;  // rows are padded to an odd stride
  nb = nb < 0 ? 0 : nb / 16 * 16;
  if (nb > MAX_NB) nb = MAX_NB;
  if (const char* e = getenv("LA_LU_MG_NB")) {  // test hook: narrower block columns on small matrices
    const int want = atoi(e) / 16 * 16;
    if (want >= 16 && want < nb) nb = want;
  }
  return nb;
}

inline int mglu_width(const la_lu_mg* c, int b) {
  const size_t left = c->n - (size_t)b * c->nb;
  return left < (size_t)c->nb ? (int)left : c->nb;
}
// first local column (on device q of G) that lies right of block k
inline int mglu_cols_through(const la_lu_mg* c, int k, int q) {
  const int blocks = k >= q ? (k - q) / c->ndev + 1 : 0;
  const int cols = blocks * c->nb;
  return cols < c->dev[q].ncols ? cols : c->dev[q].ncols;
}

int lu_mg_destroy(la_lu_mg* c) {
  if (!c) return LA_OK;
  DevSwitch keep;
  for (int q = 0; q < c->ndev; ++q) {
    LuMgDev& D = c->dev[q];
    if (D.device < 0 || cudaSetDevice(D.device) != cudaSuccess) continue;
    if (D.st) cudaStreamSynchronize(D.st);
    if (D.sp) cudaStreamSynchronize(D.sp);
    if (D.sw) cudaStreamSynchronize(D.sw);
    for (int s = 0; s < MGLU_RING; ++s) {
      for (int r = 0; r < MGLU_MAX_DEV; ++r)
        if (D.e_arr[r][s]) cudaEventDestroy(D.e_arr[r][s]);
      if (D.e_sent[s]) cudaEventDestroy(D.e_sent[s]);
      if (D.e_perm[s]) cudaEventDestroy(D.e_perm[s]);
      if (D.e_w[s]) cudaEventDestroy(D.e_w[s]);
      if (D.e_bulk[s]) cudaEventDestroy(D.e_bulk[s]);
    }
    if (D.e_in) cudaEventDestroy(D.e_in);
    if (D.e_tail) cudaEventDestroy(D.e_tail);
    if (q == 0) {
      if (c->t0) cudaEventDestroy(c->t0);
      if (c->t1) cudaEventDestroy(c->t1);
    }
    if (D.sp) cudaStreamDestroy(D.sp);
    if (D.sw) cudaStreamDestroy(D.sw);
    if (D.st) cudaStreamDestroy(D.st);
    cudaFree(D.A);
    cudaFree(D.Lring);
    cudaFree(D.Wring);
    cudaFree(D.ipiv_ring);
    cudaFree(D.ws);
    cudaFree(D.piv);
    cudaFree(D.sign);
  }
  cudaGetLastError();
  delete c;
  return LA_OK;
}

int lu_mg_create_impl(la_lu_mg* c, int ngpus, const int* devices, size_t n) {
  c->n = n;
  // panel width: as wide as the shared memory of the smallest device allows for the tallest (first) panel (cf. lu_factor_dev)
  int min_sms = 1 << 30;
  for (int q = 0; q < ngpus; ++q) {
    const DeviceCtx* ctx;
    LA_TRY(device_ctx(devices[q], &ctx));
    if (ctx->sm_count < min_sms) min_sms = ctx->sm_count;
  }
  const int nb = mglu_block_width(n, min_sms);
  if (nb < 16)
    return fail(LA_ERR_UNSUPPORTED, "la_lu_mg: %zu rows exceed the shared-memory panel capacity of %d SMs", n, min_sms);
  c->nb = nb;
  c->nblk = (int)((n + nb - 1) / nb);
  c->ndev = ngpus < c->nblk ? ngpus : c->nblk;  // a device without a block column has nothing to do
  const int G = c->ndev;
  for (int q = 0; q < G; ++q) {
    LuMgDev& D = c->dev[q];
    const DeviceCtx* ctx;
    LA_TRY(device_ctx(devices[q], &ctx));
    if (!ctx->coop) return fail(LA_ERR_UNSUPPORTED, "la_lu_mg: device %d lacks cooperative launch", devices[q]);
    D.device = devices[q];
    D.sms = ctx->sm_count;
    LA_CUDA_TRY(cudaSetDevice(D.device));
    for (int r = 0; r < G; ++r) {  // peer copies go straight over NVLink where the devices can reach each other
      if (devices[r] == D.device) continue;
      int can = 0;
      LA_CUDA_TRY(cudaDeviceCanAccessPeer(&can, D.device, devices[r]));
      if (can) {
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(LA_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d) failed: %s", devices[r], cudaGetErrorString(e));
        cudaGetLastError();
      }
    }
    int nloc = 0;
    for (int b = q; b < c->nblk; b += G) nloc += mglu_width(c, b);
    D.ncols = nloc;
    D.ld = (size_t)((nloc + 1) / 2 * 2);
    LA_CUDA_TRY(cudaMalloc(&D.A, sizeof(double) * n * D.ld));
    LA_CUDA_TRY(cudaMalloc(&D.Lring, sizeof(double) * MGLU_RING * n * MAX_NB));
    LA_CUDA_TRY(cudaMalloc(&D.Wring, sizeof(double) * MGLU_RING * MAX_NB * MAX_NB));
    LA_CUDA_TRY(cudaMalloc(&D.ipiv_ring, sizeof(int) * MGLU_RING * MAX_NB));
    LA_CUDA_TRY(cudaMalloc(&D.ws, ws_bytes<double>(D.sms)));
    LA_CUDA_TRY(cudaMalloc(&D.piv, sizeof(uint64_t) * n));
    LA_CUDA_TRY(cudaMalloc(&D.sign, sizeof(int)));
    int lo = 0, hi = 0;
    LA_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&D.sp, cudaStreamNonBlocking, hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&D.sw, cudaStreamNonBlocking, hi));
    LA_CUDA_TRY(cudaStreamCreateWithPriority(&D.st, cudaStreamNonBlocking, lo));
    for (int s = 0; s < MGLU_RING; ++s) {
      for (int r = 0; r < G; ++r) LA_CUDA_TRY(cudaEventCreateWithFlags(&D.e_arr[r][s], cudaEventDisableTiming));
      LA_CUDA_TRY(cudaEventCreateWithFlags(&D.e_sent[s], cudaEventDisableTiming));
      LA_CUDA_TRY(cudaEventCreateWithFlags(&D.e_perm[s], cudaEventDisableTiming));
      LA_CUDA_TRY(cudaEventCreateWithFlags(&D.e_w[s], cudaEventDisableTiming));
      LA_CUDA_TRY(cudaEventCreateWithFlags(&D.e_bulk[s], cudaEventDisableTiming));
    }
    LA_CUDA_TRY(cudaEventCreateWithFlags(&D.e_in, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&D.e_tail, cudaEventDisableTiming));
    if (q == 0) {
      LA_CUDA_TRY(cudaEventCreate(&c->t0));
      LA_CUDA_TRY(cudaEventCreate(&c->t1));
    }
  }
  return LA_OK;
}

int lu_mg_create(int ngpus, const int* devices, size_t n, la_lu_mg** out) {
  LA_REQUIRE(devices && out, "la_lu_mg_create: null pointer");
  LA_REQUIRE(ngpus >= 1 && ngpus <= MGLU_MAX_DEV, "la_lu_mg_create: bad device count %d", ngpus);
  LA_REQUIRE(n > 0 && n < (1u << 30), "la_lu_mg_create: bad order %zu", n);
  *out = nullptr;
  DevSwitch keep;
  la_lu_mg* c = new la_lu_mg();
  const int s = lu_mg_create_impl(c, ngpus, devices, n);
  if (s != LA_OK) {
    const std::string why = error_text();
    lu_mg_destroy(c);
    set_error("%s", why.c_str());
    return s;
  }
  *out = c;
  return LA_OK;
}

// host matrix (row-major n x n) <-> the devices' block columns
int lu_mg_upload(la_lu_mg* c, const double* A) {
  LA_REQUIRE(c && A, "la_lu_mg_upload: null pointer");
  DevSwitch keep;
  const size_t n = c->n;
  for (int b = 0; b < c->nblk; ++b) {
    LuMgDev& D = c->dev[b % c->ndev];
    LA_CUDA_TRY(cudaSetDevice(D.device));
    LA_CUDA_TRY(cudaMemcpy2DAsync(D.A + (size_t)(b / c->ndev) * c->nb, D.ld * sizeof(double), A + (size_t)b * c->nb,
                                  n * sizeof(double), (size_t)mglu_width(c, b) * sizeof(double), n, cudaMemcpyHostToDevice,
                                  D.st));
  }
  for (int q = 0; q < c->ndev; ++q) {
    LA_CUDA_TRY(cudaSetDevice(c->dev[q].device));
    LA_CUDA_TRY(cudaStreamSynchronize(c->dev[q].st));
  }
  return LA_OK;
}

int lu_mg_fill_hash(la_lu_mg* c, uint64_t seed) {
  LA_REQUIRE(c, "la_lu_mg_fill_hash: null context");
  DevSwitch keep;
  for (int q = 0; q < c->ndev; ++q) {
    LuMgDev& D = c->dev[q];
    LA_CUDA_TRY(cudaSetDevice(D.device));
    lu_mg_fill_kernel<<<D.sms * 16, 256, 0, D.st>>>(D.A, D.ld, (int)c->n, D.ncols, q, c->ndev, c->nb, seed);
    LA_CUDA_TRY(cudaGetLastError());
  }
  return LA_OK;
}

int lu_mg_sync(la_lu_mg* c) {
  LA_REQUIRE(c, "la_lu_mg_sync: null context");
  DevSwitch keep;
  for (int q = 0; q < c->ndev; ++q) {
    LA_CUDA_TRY(cudaSetDevice(c->dev[q].device));
    LA_CUDA_TRY(cudaStreamSynchronize(c->dev[q].st));
  }
  if (c->last_ms == -2.f) {  // a factorisation was queued since the last read
    LA_CUDA_TRY(cudaSetDevice(c->dev[0].device));
    LA_CUDA_TRY(cudaEventElapsedTime(&c->last_ms, c->t0, c->t1));
  }
  return LA_OK;
}

int lu_mg_download(la_lu_mg* c, double* LU, uint64_t* piv, int* pospivsign) {
  LA_REQUIRE(c, "la_lu_mg_download: null context");
  LA_TRY(lu_mg_sync(c));
  DevSwitch keep;
  const size_t n = c->n;
  if (LU) {
    for (int b = 0; b < c->nblk; ++b) {
      LuMgDev& D = c->dev[b % c->ndev];
      LA_CUDA_TRY(cudaSetDevice(D.device));
      LA_CUDA_TRY(cudaMemcpy2DAsync(LU + (size_t)b * c->nb, n * sizeof(double), D.A + (size_t)(b / c->ndev) * c->nb,
                                    D.ld * sizeof(double), (size_t)mglu_width(c, b) * sizeof(double), n,
                                    cudaMemcpyDeviceToHost, D.st));
    }
  }
  LuMgDev& D0 = c->dev[0];
  LA_CUDA_TRY(cudaSetDevice(D0.device));
  if (piv) LA_CUDA_TRY(cudaMemcpyAsync(piv, D0.piv, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, D0.st));
  if (pospivsign) LA_CUDA_TRY(cudaMemcpyAsync(pospivsign, D0.sign, sizeof(int), cudaMemcpyDeviceToHost, D0.st));
  return lu_mg_sync(c);
}

int lu_mg_factor(la_lu_mg* c) {
  LA_REQUIRE(c, "la_lu_mg_factor: null context");
  DevSwitch keep;
  using T = double;
  const int G = c->ndev, M = (int)c->n, nblk = c->nblk;
  const int SWAP_SMEM = (int)(sizeof(T) * MAX_MOVES * SWAP_W);
  const int INVL_SMEM = (int)(sizeof(T) * ((size_t)MAX_NB * INVL_LD + 3 * IB * IB));
  const int HEAD_SMEM = (int)(sizeof(T) * ((size_t)MAX_NB * HEAD_LD + (size_t)MAX_NB * (HEAD_COLS + 1) +
                                           (size_t)MAX_NB * HEAD_COLS));
  auto use = [&](LuMgDev& D) -> int {
    LA_CUDA_TRY(cudaSetDevice(D.device));
    return LA_OK;
  };
  auto lslot = [&](LuMgDev& D, int slot) { return D.Lring + (size_t)slot * M * MAX_NB; };
  auto launch_panel = [&](LuMgDev& D, int j0, int jb, int lc) -> int {  // columns [lc, lc + jb) of D hold block column j0
    const PanelShape ps = panel_shape<T>(M - j0, jb, D.sms, false);
    // the kernel addresses column j0 + c of a full matrix: shift the base so that this lands on local column lc + c
    T* a = (T*)((uintptr_t)D.A + ((intptr_t)lc - (intptr_t)j0) * (intptr_t)sizeof(T));
    size_t ld = D.ld;
    int mm = M, jj0 = j0, jjb = jb, rr = ps.rpc, ioff = 0;
    void* wsb = D.ws;
    unsigned ep = ++D.epoch;
    void* args[] = {&a, &ld, &mm, &jj0, &jjb, &rr, &wsb, &ep, &ioff};
    LA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)lu_panel_kernel<T, false>, dim3(ps.G), dim3(PANEL_THREADS), args,
                                            ps.smem, D.sp));
    return LA_OK;
  };

  // ---- per device: kernel attributes, piv = identity, exchange tags invalid; the clock starts on device 0 ----
  LA_TRY(use(c->dev[0]));
  LA_CUDA_TRY(cudaEventRecord(c->t0, c->dev[0].st));
  for (int q = 0; q < G; ++q) {
    LuMgDev& D = c->dev[q];
    LA_TRY(use(D));
    LA_CUDA_TRY(cudaFuncSetAttribute(lu_panel_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)PANEL_SMEM_BUDGET + 2048));
    LA_CUDA_TRY(cudaFuncSetAttribute(lu_swap_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWAP_SMEM));
    LA_CUDA_TRY(cudaFuncSetAttribute(lu_invl_kernel<T, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, INVL_SMEM));
    LA_CUDA_TRY(cudaFuncSetAttribute(lu_head_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAD_SMEM));
    if (q > 0) LA_CUDA_TRY(cudaStreamWaitEvent(D.st, c->t0, 0));
    lu_init_piv_kernel<T><<<(M + 255) / 256, 256, 0, D.st>>>(D.piv, M, D.sign);
    LA_CUDA_TRY(cudaGetLastError());
    LA_CUDA_TRY(cudaMemsetAsync((char*)D.ws + ws_hdr_bytes<T>(), 0, ws_bytes<T>(D.sms) - ws_hdr_bytes<T>(), D.st));
    D.epoch = 0;
    LA_CUDA_TRY(cudaEventRecord(D.e_in, D.st));
    LA_CUDA_TRY(cudaStreamWaitEvent(D.sp, D.e_in, 0));
    LA_CUDA_TRY(cudaStreamWaitEvent(D.sw, D.e_in, 0));
  }
  LA_TRY(use(c->dev[0]));
  LA_TRY(launch_panel(c->dev[0], 0, mglu_width(c, 0), 0));
  static const bool trace = getenv("LA_LU_MG_TRACE") && atoi(getenv("LA_LU_MG_TRACE")) != 0;  // host issue time on stderr
  timespec ts0;
  clock_gettime(CLOCK_MONOTONIC, &ts0);

  for (int k = 0; k < nblk; ++k) {
    const int j0 = k * c->nb, jb = mglu_width(c, k), c1 = j0 + jb;
    const int o = k % G, slot = k % MGLU_RING;
    const bool has_next = k + 1 < nblk;
    const int nb2 = has_next ? mglu_width(c, k + 1) : 0;
    const int o2 = (k + 1) % G;
    const int lc0 = (k / G) * c->nb, lcn = ((k + 1) / G) * c->nb;
    LuMgDev& O = c->dev[o];
    // ---- owner: the factored block column and its pivots go to every device's ring slot (the next owner first) ----
    LA_TRY(use(O));
    for (int i = 0; i < G; ++i) {
      const int q = (o2 + i) % G;
      LuMgDev& D = c->dev[q];
      if (k >= MGLU_RING) LA_CUDA_TRY(cudaStreamWaitEvent(O.sp, D.e_bulk[slot], 0));  // step k - RING has left the slot
      LA_CUDA_TRY(cudaMemcpy2DAsync(lslot(D, slot), MAX_NB * sizeof(T), O.A + (size_t)j0 * O.ld + lc0, O.ld * sizeof(T),
                                    (size_t)jb * sizeof(T), (size_t)(M - j0), cudaMemcpyDeviceToDevice, O.sp));
      LA_CUDA_TRY(cudaMemcpyAsync(D.ipiv_ring + slot * MAX_NB, ((PanelWs<T>*)O.ws)->ipiv, sizeof(int) * MAX_NB,
                                  cudaMemcpyDeviceToDevice, O.sp));
      LA_CUDA_TRY(cudaEventRecord(O.e_arr[q][slot], O.sp));
    }
    LA_CUDA_TRY(cudaEventRecord(O.e_sent[slot], O.sp));
    // ---- every device, the next owner first ----
    for (int i = 0; i < G; ++i) {
      const int q = (o2 + i) % G;
      LuMgDev& D = c->dev[q];
      LA_TRY(use(D));
      T* Ls = lslot(D, slot);
      T* W = D.Wring + (size_t)slot * MAX_NB * MAX_NB;
      const bool next_owner = has_next && q == o2;
      const int lr = mglu_cols_through(c, k, q);       // local columns right of the panel start here
      const int cb = lr + (next_owner ? nb2 : 0);      // ... and the bulk's share of them here
      const int wcols = D.ncols - cb;
      // chain stream: pivots -> net permutation, this device's piv / sign
      LA_CUDA_TRY(cudaStreamWaitEvent(D.sp, O.e_arr[q][slot], 0));
      lu_perm_kernel<T><<<1, 2 * MAX_NB, 0, D.sp>>>(D.ws, 1, slot, j0, jb, D.piv, D.sign, 0, 1, D.ipiv_ring + slot * MAX_NB);
      LA_CUDA_TRY(cudaGetLastError());
      LA_CUDA_TRY(cudaEventRecord(D.e_perm[slot], D.sp));
      // side stream: W = inv(L11) from the ring slot (only the bulk needs it)
      if (wcols > 0) {
        LA_CUDA_TRY(cudaStreamWaitEvent(D.sw, O.e_arr[q][slot], 0));
        const T* Lshift = (const T*)((uintptr_t)Ls - ((size_t)j0 * MAX_NB + j0) * sizeof(T));  // (j0 + i, j0 + k) -> Ls[i][k]
        lu_invl_kernel<T, 0><<<1, INVL_THREADS, INVL_SMEM, D.sw>>>(Lshift, MAX_NB, j0, jb, c1, W, 0);
        LA_CUDA_TRY(cudaGetLastError());
        LA_CUDA_TRY(cudaEventRecord(D.e_w[slot], D.sw));
      }
      // chain stream of the next owner: its panel's columns first, then the panel
      if (next_owner) {
        if (k > 0) LA_CUDA_TRY(cudaStreamWaitEvent(D.sp, D.e_bulk[(k - 1) % MGLU_RING], 0));  // bulk(k-1) updated them
        lu_head_kernel<T><<<(nb2 + HEAD_COLS - 1) / HEAD_COLS, 256, HEAD_SMEM, D.sp>>>(D.A, D.ld, j0, jb, lcn, lcn + nb2,
                                                                                     D.ws, 1, slot, Ls, MAX_NB);
        LA_CUDA_TRY(cudaGetLastError());
        if (c1 < M)
          LA_TRY(gemm_f64_tensor(Ls + (size_t)(c1 - j0) * MAX_NB, MAX_NB, D.A + (size_t)j0 * D.ld + lcn, D.ld,
                                 D.A + (size_t)c1 * D.ld + lcn, D.ld, (size_t)(M - c1), (size_t)jb, (size_t)nb2, LA_GEMM_SUB,
                                 D.sp));
        LA_TRY(launch_panel(D, c1, nb2, lcn));
      }
      // bulk stream: interchanges of every other local column, U12 = W * A12, A22 -= L21 * U12
      LA_CUDA_TRY(cudaStreamWaitEvent(D.st, D.e_perm[slot], 0));
      if (q == o) LA_CUDA_TRY(cudaStreamWaitEvent(D.st, O.e_sent[slot], 0));  // later interchanges touch the panel's columns
      int s0 = 0, s1 = 0;  // columns interchanged elsewhere: the panel's own (inside the panel kernel), the next panel's (head)
      if (q == o) {
        s0 = lc0;
        s1 = lc0 + jb;
      }
      if (next_owner) {
        if (q == o) {
          s1 += nb2;  // one device: the two ranges are adjacent
        } else {
          s0 = lcn;
          s1 = lcn + nb2;
        }
      }
      const int nswap = D.ncols - (s1 - s0);
      if (nswap > 0) {
        lu_swap_kernel<T><<<(nswap + SWAP_W - 1) / SWAP_W, 256, SWAP_SMEM, D.st>>>(D.A, D.ld, 0, D.ncols, s0, s1, D.ws, 1, slot);
        LA_CUDA_TRY(cudaGetLastError());
      }
      if (wcols > 0) {
        LA_CUDA_TRY(cudaStreamWaitEvent(D.st, D.e_w[slot], 0));
        T* U12 = D.A + (size_t)j0 * D.ld + cb;
        LA_TRY(gemm_f64_tensor(W, MAX_NB, U12, D.ld, U12, D.ld, (size_t)jb, (size_t)jb, (size_t)wcols, LA_GEMM_ASSIGN, D.st));
        if (c1 < M)
          LA_TRY(gemm_f64_tensor(Ls + (size_t)(c1 - j0) * MAX_NB, MAX_NB, U12, D.ld, D.A + (size_t)c1 * D.ld + cb, D.ld,
                                 (size_t)(M - c1), (size_t)jb, (size_t)wcols, LA_GEMM_SUB, D.st));
      }
      LA_CUDA_TRY(cudaEventRecord(D.e_bulk[slot], D.st));
    }
  }
  // ---- every bulk stream covers its device's chain and side streams; device 0's covers every device; clock stops ----
  for (int q = 0; q < G; ++q) {
    LuMgDev& D = c->dev[q];
    LA_TRY(use(D));
    LA_CUDA_TRY(cudaEventRecord(D.e_tail, D.sp));
    LA_CUDA_TRY(cudaStreamWaitEvent(D.st, D.e_tail, 0));
    LA_CUDA_TRY(cudaEventRecord(D.e_in, D.sw));
    LA_CUDA_TRY(cudaStreamWaitEvent(D.st, D.e_in, 0));
    LA_CUDA_TRY(cudaEventRecord(D.e_tail, D.st));
  }
  LuMgDev& D0 = c->dev[0];
  LA_TRY(use(D0));
  for (int q = 1; q < G; ++q) LA_CUDA_TRY(cudaStreamWaitEvent(D0.st, c->dev[q].e_tail, 0));
  LA_CUDA_TRY(cudaEventRecord(c->t1, D0.st));
  c->last_ms = -2.f;
  if (trace) {
    timespec ts1;
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    fprintf(stderr, "[lu_mg] n=%d devices=%d: all work queued after %.2f ms of host time\n", M, G,
            (ts1.tv_sec - ts0.tv_sec) * 1e3 + (ts1.tv_nsec - ts0.tv_nsec) * 1e-6);
  }
  return LA_OK;
}

}  // namespace
}  // namespace la

extern "C" {
int la_lu_mg_create(int ngpus, const int* devices, size_t n, la_lu_mg** out) { return la::lu_mg_create(ngpus, devices, n, out); }
int la_lu_mg_destroy(la_lu_mg* ctx) { return la::lu_mg_destroy(ctx); }
int la_lu_mg_upload_f64(la_lu_mg* ctx, const double* A) { return la::lu_mg_upload(ctx, A); }
int la_lu_mg_fill_hash_f64(la_lu_mg* ctx, uint64_t seed) { return la::lu_mg_fill_hash(ctx, seed); }
int la_lu_mg_factor_f64(la_lu_mg* ctx) { return la::lu_mg_factor(ctx); }
int la_lu_mg_sync(la_lu_mg* ctx) { return la::lu_mg_sync(ctx); }
int la_lu_mg_download_f64(la_lu_mg* ctx, double* LU_out, uint64_t* piv_out, int* pospivsign_out) {
  return la::lu_mg_download(ctx, LU_out, piv_out, pospivsign_out);
}
int la_lu_mg_last_ms(la_lu_mg* ctx, float* ms_out) {
  if (!ctx || !ms_out) return la::fail(LA_ERR_INVALID, "la_lu_mg_last_ms: null pointer");
  int s = la::lu_mg_sync(ctx);
  if (s != LA_OK) return s;
  if (ctx->last_ms < 0.f) return la::fail(LA_ERR_INVALID, "la_lu_mg_last_ms: no factorisation has run in this context");
  *ms_out = ctx->last_ms;
  return LA_OK;
}
// The layout as pure arithmetic (no device needed): block width, number of block columns, devices in use, and the number
// of local columns of every device in use (ncols_out[ngpus]; unused entries are 0).  Block b lives on device b % ndev.
int la_lu_mg_plan(size_t n, int ngpus, int sm_count, int* block_width_out, int* nblocks_out, int* ndev_out,
                  size_t* ncols_out) {
  if (!block_width_out || !nblocks_out || !ndev_out || !ncols_out)
    return la::fail(LA_ERR_INVALID, "la_lu_mg_plan: null output");
  if (n == 0 || n >= (1u << 30) || ngpus < 1 || ngpus > la::MGLU_MAX_DEV || sm_count < 1)
    return la::fail(LA_ERR_INVALID, "la_lu_mg_plan: bad arguments (n=%zu ngpus=%d sm_count=%d)", n, ngpus, sm_count);
  const int nb = la::mglu_block_width(n, sm_count);
  if (nb < 16)
    return la::fail(LA_ERR_UNSUPPORTED, "la_lu_mg: %zu rows exceed the shared-memory panel capacity of %d SMs", n, sm_count);
  const int nblk = (int)((n + nb - 1) / nb);
  const int ndev = ngpus < nblk ? ngpus : nblk;
  for (int q = 0; q < ngpus; ++q) ncols_out[q] = 0;
  for (int b = 0; b < nblk; ++b) {
    const size_t left = n - (size_t)b * nb;
    ncols_out[b % ndev] += left < (size_t)nb ? left : (size_t)nb;
  }
  *block_width_out = nb;
  *nblocks_out = nblk;
  *ndev_out = ndev;
  return LA_OK;
}
int la_lu_mg_devices(const la_lu_mg* ctx, int* ndev_out) {
  if (!ctx || !ndev_out) return la::fail(LA_ERR_INVALID, "la_lu_mg_devices: null pointer");
  *ndev_out = ctx->ndev;
  return LA_OK;
}
// One call from host memory: A (row-major n x n) -> packed LU, piv, pospivsign, across the listed devices.
int la_lu_factor_f64_mg(int ngpus, const int* devices, const double* A, double* LU_out, size_t n, uint64_t* piv_out,
                        int* pospivsign_out) {
  if (!A || !LU_out || !piv_out || !pospivsign_out) return la::fail(LA_ERR_INVALID, "la_lu_factor_f64_mg: null pointer");
  la_lu_mg* c = nullptr;
  int s = la::lu_mg_create(ngpus, devices, n, &c);
  if (s == LA_OK) s = la::lu_mg_upload(c, A);
  if (s == LA_OK) s = la::lu_mg_factor(c);
  if (s == LA_OK) s = la::lu_mg_download(c, LU_out, piv_out, pospivsign_out);
  if (s != LA_OK) {
    const std::string why = la::error_text();
    la::lu_mg_destroy(c);
    la::set_error("%s", why.c_str());
    return s;
  }
  return la::lu_mg_destroy(c);
}
}  // extern "C"
