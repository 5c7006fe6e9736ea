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
#include <stdlib.h>

#include <stdio.h>

#include <type_traits>
#include <vector>

#include "la_common.cuh"
#include "ll_exchange.cuh"

namespace la {
int gemm_f64_tensor(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, cudaStream_t st);  // gemm_f64.cu: TMA/DMMA kernel regardless of size
int lu_diag_block_inverses(const double* LU, size_t n, double* WL, double* WU, cudaStream_t st);  // lu.cu
template <typename T>
int tri_block_inverses(const T* M, size_t n, int mode, int first_block, int nblocks, T* W, int trans_out, cudaStream_t st);  // lu.cu
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

// ---------------------------------------------------------------------------------------------------------------
// Few right-hand sides (nx <= 16), fp64: one persistent cooperative kernel per sweep instead of ~4*n/SB launches.
//   CTA g owns rows [128 g, 128 g + 128) of X for the whole sweep and keeps them in registers (thread <-> 2 rows x 8
//   columns).  Forward: for k = 0 .. g-1 it waits for block k of the solution (a flag per reader, set by CTA k after a
//   fence), subtracts LU[g-block][k-block] * X_k -- LU streamed through shared memory by cp.async, two 64-column
//   halves in flight -- then multiplies by the INVERTED diagonal block (lu_diag_block_inverses, one batched launch per
//   triangle up front; the same policy as the factorisation's U12 = inv(L11) * A12) and publishes.  Backward is the
//   mirror image with U.  A substitution inside the block would be one dependent chain of 128 x (FMA + shuffle
//   [+ divide]) per block -- measured 6.7 us forward / 13 us backward per block against 2 us for the product.
//   LU is read exactly once per sweep.  Systems below 512 rows take the launch-per-block path below, which keeps the
//   reference's order with separately rounded operations and is bit-identical to it.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PB = 128;  // rows per CTA = columns per step
constexpr int PH = 64;   // columns per shared-memory buffer
constexpr int SWEEP_THREADS = 256;
constexpr int SWEEP_SMEM = (2 * PB * PH + PB * 16) * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// Lbuf[half][row][64 columns]: 16-byte chunks XOR-swizzled by the row, so the inner loop's 128-bit reads (lane <-> row,
// fixed column pair) are conflict-free.

template <bool FORWARD>
__global__ void __launch_bounds__(SWEEP_THREADS, 1)
solve_sweep_kernel(const double* __restrict__ LU, size_t n, const uint64_t* __restrict__ piv, const double* __restrict__ B,
                   double* __restrict__ X, int nx, unsigned* __restrict__ flags /* [G readers][G blocks] */,
                   const double* __restrict__ Winv /* [G][128][128]: inverted diagonal blocks of L (FORWARD) or U */,
                   unsigned long long* __restrict__ dbg /* optional [G][8] phase timestamps (ns) of the last step */) {
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  double* Lbuf = reinterpret_cast<double*>(sweep_smem);      // [2][PB][PH]
  double* Ys = Lbuf + 2 * PB * PH;                           // [PB][16]: block k of the solution
  const int G = gridDim.x, g = blockIdx.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int N = (int)n;
  const int r0 = g * PB;
  const int nr = min(PB, N - r0);
  const uint32_t lbuf_s = (uint32_t)__cvta_generic_to_shared(Lbuf);

  // one 64-column half of LU[r0 .. r0+128)[128 kb + 64 hh ..) into buffer hh (zero-filled outside the matrix);
  // kb < 0: the CTA's own inverted diagonal block instead (dense 128 x 128, zero-padded)
  auto issue_half = [&](int kb, int hh) {
    const int cbase = kb * PB + hh * PH;
    const double* wsrc = Winv + (size_t)g * PB * PB + hh * PH;
#pragma unroll 4
    for (int i = 0; i < (PB * PH / 2) / SWEEP_THREADS; ++i) {
      const int id = t + SWEEP_THREADS * i;
      const int rr = id >> 5, cc = id & 31;
      const uint32_t dst = lbuf_s + (uint32_t)(((hh * PB + rr) * PH + ((cc ^ (rr & 7)) << 1)) * sizeof(double));
      if (kb < 0) {
        cp_async16(dst, wsrc + (size_t)rr * PB + 2 * cc, 16);
      } else {
        const int col = cbase + 2 * cc;
        const bool ok = r0 + rr < N && col < N;  // n is even: a pair of columns is inside or outside as a whole
        cp_async16(dst, ok ? LU + (size_t)(r0 + rr) * n + col : LU, ok ? 16 : 0);
      }
    }
    cp_async_commit();
  };

  // Warps 0..3 compute: thread <-> 2 rows x 8 columns of the right-hand sides (shared-memory reads, not FP64 issue
  // slots, bound the inner loop: 2x8 needs 6 128-bit loads per 32 FMAs).  Rows lane + 32 i keep the swizzled reads
  // conflict-free.  Even and odd k accumulate separately: a DFMA result takes ~64 cycles to come back, and one chain
  // of 128 dependent FMAs per element would be the whole step.
  const int urow = ((warp & 1) << 6) + lane;  // + 32 i
  const int ucol = (warp >> 1) << 3;          // 8 columns from here (warps 0..3 only)
  double acc[2][8];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rr = urow + 32 * i, col = ucol + j;
      acc[i][j] = 0.0;
      if (warp < 4 && rr < nr && col < nx)  // X = B(piv,:), lu.rs:246-254
        acc[i][j] = FORWARD ? B[(size_t)(piv ? piv[r0 + rr] : (uint64_t)(r0 + rr)) * nx + col]
                            : X[(size_t)(r0 + rr) * nx + col];
    }
  // a0/a1 -= (rows urow, urow+32 of buffer hh) * (rows [64 hh, 64 hh + 64) of Ys)
  auto mac_half = [&](int hh, double (&a0)[2][8], double (&a1)[2][8]) {
    const double* Lr0 = Lbuf + (hh * PB + urow) * PH;
    const double* Lr1 = Lr0 + 32 * PH;
    const int sw = urow & 7;  // == (urow + 32) & 7
    const double* Yh = Ys + (hh * PH) * 16 + ucol;
#pragma unroll 4
    for (int i = 0; i < PH / 2; ++i) {
      const int kk = 2 * i;
      const int lo = ((kk >> 1) ^ sw) << 1;
      const double2 l0 = *reinterpret_cast<const double2*>(Lr0 + lo);
      const double2 l1 = *reinterpret_cast<const double2*>(Lr1 + lo);
      double y[8];
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(&y[j]) = *reinterpret_cast<const double2*>(Yh + kk * 16 + j);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0[0][j] = fma(-y[j], l0.x, a0[0][j]);
        a0[1][j] = fma(-y[j], l1.x, a0[1][j]);
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2)
        *reinterpret_cast<double2*>(&y[j]) = *reinterpret_cast<const double2*>(Yh + (kk + 1) * 16 + j);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[0][j] = fma(-y[j], l0.y, a1[0][j]);
        a1[1][j] = fma(-y[j], l1.y, a1[1][j]);
      }
    }
  };

  auto stamp = [&](int slot) {
    if (dbg && t == 0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
      dbg[(size_t)g * 8 + slot] = ns;
    }
  };
  const int nsteps = FORWARD ? g : G - 1 - g;  // blocks of the solution this CTA consumes before its own
  auto step_block = [&](int s) { return FORWARD ? s : G - 1 - s; };
  {
    const int kb = nsteps > 0 ? step_block(0) : -1;
    issue_half(kb, 0);
    issue_half(kb, 1);
  }
  for (int s = 0; s < nsteps; ++s) {
    const int kb = step_block(s);
    const int next_kb = (s + 1 < nsteps) ? step_block(s + 1) : -1;  // the CTA's own inverted diagonal block comes last
    if (t == 0) {
      const volatile unsigned* f = flags + (size_t)g * G + kb;
      while (*f == 0u) {
      }
      __threadfence();
    }
    __syncthreads();
    if (s == nsteps - 1) stamp(0);  // flag of the last awaited block seen
    for (int idx = t; idx < PB * 16; idx += SWEEP_THREADS) {
      const int rr = idx >> 4, col = idx & 15;
      const int grow = kb * PB + rr;
      Ys[idx] = (grow < N && col < nx) ? __ldcg(&X[(size_t)grow * nx + col]) : 0.0;
    }
    double odd[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) odd[i][j] = 0.0;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      cp_async_wait<1>();
      __syncthreads();  // buffer hh (and, the first time round, Ys) is ready
      if (warp < 4) mac_half(hh, acc, odd);
      __syncthreads();  // everyone is done with buffer hh
      issue_half(next_kb, hh);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] += odd[i][j];
  }
  stamp(1);  // updates done

  // ---- X_g = W_g * (right-hand sides of this block): the same inner loop on the inverted diagonal block ----
  if (warp < 4) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) Ys[(urow + 32 * i) * 16 + ucol + j] = acc[i][j];
  }
  cp_async_wait<0>();
  __syncthreads();  // W_g is in the two buffers, the block's right-hand sides in Ys
  stamp(2);
  if (warp < 4) {
    double even[2][8], odd[2][8];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) even[i][j] = odd[i][j] = 0.0;
    mac_half(0, even, odd);
    mac_half(1, even, odd);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int rr = urow + 32 * i;
      if (rr < nr) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (ucol + j < nx) X[(size_t)(r0 + rr) * nx + ucol + j] = -(even[i][j] + odd[i][j]);
      }
    }
  }
  stamp(3);  // block solved, X stored
  __threadfence();
  __syncthreads();
  stamp(4);  // fenced
  // one flag per reader: nobody spins on a line somebody else spins on
  for (int rd = t; rd < G; rd += SWEEP_THREADS)
    if (FORWARD ? rd > g : rd < g) *((volatile unsigned*)(flags + (size_t)rd * G + g)) = 1u;
}

// ---------------------------------------------------------------------------------------------------------------
// Second-generation sweep (default): the same block-row ownership and cp.async streaming of LU, but
//   * the 128 x 128 x 16 block products run on DMMA.8x8x4 (mma.sync m8n8k4 f64) with ALL eight warps: warp w owns rows
//     [16 w, 16 w + 16) as 2 x 2 accumulator tiles; per 4-deep k-step it needs 4 shared-memory loads for 4 MMAs (the
//     CUDA-core version needed 10 128-bit loads per 32 FMAs and ran on four warps);
//   * the solved blocks travel as FLAGGED WORDS (value halves next to a per-call tag, ll_exchange.cuh): the producer
//     stores its 128 x 16 block once, every consumer polls the data itself -- no __threadfence, no flag array, no
//     separate fetch of X after the flag (one L2 round trip instead of three on the dependent chain).
// Shared memory: two 64-column halves of the LU block (swizzled 16-byte chunks, as above) + the X block as the MMA's
// B operand with a 24-double row stride (4 k-rows x 8 columns of a fragment load then cover all 32 banks twice).
// ---------------------------------------------------------------------------------------------------------------
// NXC = right-hand-side columns per chain: 16 (one chain, 64-column halves of LU, three 64 KB buffers, one CTA per SM) or
// 8 (two INDEPENDENT chains for 9..16 right-hand sides -- each column of X is its own triangular system -- or one chain for
// <= 8; 32-column quarters, three 32 KB buffers, two CTAs per SM).  Two chains halve the block products on the dependent
// path (each CTA multiplies 128 x 128 x 8), the critical CTAs of the two chains sit on different SMs, and the second reader
// of every LU block hits L2.
// BULK (experiment, off by default -- slower, see tri_sweeps_dev): every row segment of a part arrives by ONE cp.async.bulk
// (TMA engine, 256 / 512 contiguous bytes, completion on a per-warp mbarrier) instead of 16 / 32 sixteen-byte cp.async.
// Rows are padded by two doubles instead of swizzled (a bulk copy is contiguous); needs n a multiple of 128.
template <int NXC, bool BULK = false>
struct SweepCfg {
  static constexpr int PHC = NXC == 16 ? 64 : 32;      // LU columns per shared-memory buffer
  static constexpr int RS = BULK ? PHC + 2 : PHC;      // row stride of a buffer (doubles)
  static constexpr int BAR_BYTES = BULK ? 256 : 0;     // 8 warps x 3 buffers mbarriers
  static constexpr int PARTS = PB / PHC;               // buffers per 128-column block
  static constexpr int NT = NXC / 8;                   // 8-column MMA tiles
  // row stride of the X block: the four k-rows x 32 bytes a half warp reads must tile a 128-byte bank period (96 / 160 bytes)
  static constexpr int XLD = NXC == 16 ? 20 : 12;
  static constexpr int BUFS = 3;
  static constexpr int WORDS = PB * NXC;               // flagged words per solved block
  static constexpr int WPT = WORDS / SWEEP_THREADS;    // ... per thread
  static constexpr int SMEM = (BUFS * PB * RS + PB * XLD) * (int)sizeof(double) + BAR_BYTES;
  static constexpr int CTAS_PER_SM = NXC == 16 ? 1 : 2;
  // independent accumulator sets per output tile (k-steps are dealt round robin): a DMMA result takes ~100+ cycles to come
  // back, and one set per tile made every block product a chain of 32 dependent MMAs -- latency, not the FP64 pipe, set its time
  static constexpr int KSPLIT = NXC == 16 ? 2 : 4;
};
typedef LL<double>::word XWord;

// M_g = W_g * T[g][g -+ 1] (FORWARD: the block left of the diagonal block; backward: the block right of it), zero where
// the triangle has no such block.  With it the solved block of the neighbour enters a CTA's result by ONE product,
// x_g = W_g (b_g - sum_{others} ...) - M_g x_neighbour: the own-block product W_g (...) no longer waits for the neighbour.
template <bool FORWARD>
__global__ void __launch_bounds__(256) sweep_combine_kernel(const double* __restrict__ T, size_t ld, int n,
                                                             const double* __restrict__ Winv, double* __restrict__ Mout) {
  __shared__ double Ws[PB][17];   // W_g[:, k0 .. k0+16)
  __shared__ double Ts[16][PB + 1];  // T block rows k0 .. k0+16
  const int g = blockIdx.x, G = gridDim.x;
  const int nb = FORWARD ? g - 1 : g + 1;
  double* M = Mout + (size_t)g * PB * PB;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  if (nb < 0 || nb >= G) {
    for (int i = t; i < PB * PB; i += 256) M[i] = 0.0;
    return;
  }
  const double* W = Winv + (size_t)g * PB * PB;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < PB; k0 += 16) {
    for (int i = t; i < PB * 16; i += 256) {
      const int r = i >> 4, k = i & 15;
      Ws[r][k] = W[(size_t)r * PB + k0 + k];
    }
    for (int i = t; i < 16 * PB; i += 256) {
      const int k = i / PB, c = i - k * PB;
      const int row = g * PB + k0 + k, col = nb * PB + c;
      Ts[k][c] = (row < n && col < n) ? T[(size_t)row * ld + col] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = Ws[ty * 8 + i][k];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Ts[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) M[(size_t)(ty * 8 + i) * PB + tx + 16 * j] = acc[i][j];
}

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}

template <bool FORWARD, int NXC, bool BULK>
__global__ void __launch_bounds__(SWEEP_THREADS, SweepCfg<NXC, BULK>::CTAS_PER_SM)
solve_sweep_mma_kernel(const double* __restrict__ LU, size_t ld, size_t n, const uint64_t* __restrict__ piv,
                       const double* __restrict__ B, double* __restrict__ X, int nx, int G,
                       XWord* __restrict__ xbuf /* [chains][G][PB * NXC] flagged words */,
                       unsigned tag, const double* __restrict__ Winv /* [G][128][128] inverted diagonal blocks */,
                       const double* __restrict__ Mcomb /* [G][128][128] W_g * neighbour block, or null: plain last update */,
                       unsigned long long* __restrict__ dbg /* optional [chains * G][8] phase timestamps (ns) of the last step */) {
  using Cfg = SweepCfg<NXC, BULK>;
  constexpr int PHC = Cfg::PHC, PARTS = Cfg::PARTS, NT = Cfg::NT, XLD = Cfg::XLD, BUFS = Cfg::BUFS, WPT = Cfg::WPT;
  constexpr int KS = Cfg::KSPLIT, RS = Cfg::RS;
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  double* Lbuf = reinterpret_cast<double*>(sweep_smem);  // [BUFS][PB][RS]
  double* Xs = Lbuf + BUFS * PB * RS;                    // [PB][XLD]: MINUS block k of the solution / this block's rhs
  uint64_t* bars = reinterpret_cast<uint64_t*>(Xs + PB * XLD);  // BULK: [8 warps][BUFS] "part has landed"
  const int chain = blockIdx.x / G, g = blockIdx.x - chain * G;
  const int c0 = chain * NXC;                            // first right-hand side of this chain
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int gid = lane >> 2, tig = lane & 3;             // MMA fragment coordinates
  const int N = (int)n;
  const int r0 = g * PB;
  const int nr = min(PB, N - r0);
  const uint32_t lbuf_s = (uint32_t)__cvta_generic_to_shared(Lbuf);
  xbuf += (size_t)chain * G * Cfg::WORDS;

  const int nsteps = FORWARD ? g : G - 1 - g;  // blocks of the solution this CTA consumes before its own
  auto step_block = [&](int s) { return FORWARD ? s : G - 1 - s; };
  // The CTA consumes a sequence of PHC-column parts: PARTS per block of LU (steps 0 .. nsteps-1), then the parts of its own
  // inverted diagonal block.  Part i lives in buffer i % 3 and is requested three parts ahead, so the inverted block is
  // already (almost) in place when the last update finishes -- its load is off the dependent chain.
  // With the precombined neighbour blocks (Mcomb) the sequence is: plain updates for steps 0 .. nsteps-2, the own inverted
  // block (its product starts as soon as the second-to-last solved block has been applied), then M_g against the LAST
  // awaited block -- one product between that block's arrival and the publication of this CTA's block.
  const bool comb = Mcomb != nullptr && nsteps > 0;
  const int ublocks = comb ? nsteps - 1 : nsteps;       // blocks applied by plain updates
  const int nparts = PARTS * (ublocks + 1 + (comb ? 1 : 0));
  constexpr int CHUNKS = PHC / 2;  // 16-byte chunks per buffer row
  // Every warp streams ITS OWN 16 rows of each part into a private slice of the buffer (rows 16 warp .. 16 warp + 15) and
  // is the only reader of that slice: completion is a per-thread cp.async wait plus __syncwarp, no block-wide barrier per
  // part (eight warps rendezvousing sixteen times per step was most of a step's time).
  // The 16-byte chunks a thread copies keep their place from part to part (same rows of the warp's slice, same chunk
  // column): shared-memory offsets, row validity and the row-to-row pointer strides are computed once; per part only the
  // base pointers change (the address arithmetic of 8..16 cp.async per thread and part was as many issue slots as the MMAs).
  constexpr int NQ = (16 * CHUNKS) / 32;       // chunks per thread and part
  constexpr int ROWSTEP = 32 / CHUNKS;         // rows between a thread's consecutive chunks (CHUNKS = 16 or 32)
  const int cc = lane % CHUNKS;
  const int rr0 = 16 * warp + lane / CHUNKS;   // first row of this thread
  uint32_t dst_off[NQ];
  unsigned row_ok = 0;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int rr = rr0 + q * ROWSTEP;
    dst_off[q] = (uint32_t)((rr * PHC + ((cc ^ ((rr & 3) << 1)) << 1)) * sizeof(double));
    if (r0 + rr < N) row_ok |= 1u << q;
  }
  const char* lu_row0 = reinterpret_cast<const char*>(LU + (size_t)(r0 + rr0) * ld + 2 * cc);
  const size_t lu_stride = (size_t)ROWSTEP * ld * sizeof(double);
  const char* w_row0 = reinterpret_cast<const char*>(Winv + (size_t)g * PB * PB + (size_t)rr0 * PB + 2 * cc);
  const char* m_row0 = reinterpret_cast<const char*>((comb ? Mcomb : Winv) + (size_t)g * PB * PB + (size_t)rr0 * PB + 2 * cc);
  constexpr size_t W_STRIDE = (size_t)ROWSTEP * PB * sizeof(double);
  // BULK: lane q < 16 copies row 16 warp + q of the part (PHC * 8 contiguous bytes); the warp's mbarrier of that buffer
  // counts the bytes.  Every row and column is inside the matrix (n is a multiple of 128 on this path).
  const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bars + warp * BUFS);
  const char* lu_brow = reinterpret_cast<const char*>(LU + (size_t)(r0 + 16 * warp + (lane & 15)) * ld);
  const char* w_brow = reinterpret_cast<const char*>(Winv + (size_t)g * PB * PB + (size_t)(16 * warp + (lane & 15)) * PB);
  const char* m_brow = reinterpret_cast<const char*>((comb ? Mcomb : Winv) + (size_t)g * PB * PB +
                                                     (size_t)(16 * warp + (lane & 15)) * PB);
  const uint32_t brow_dst = (uint32_t)((16 * warp + (lane & 15)) * RS * sizeof(double));
  auto issue_part = [&](int i) {
    if constexpr (BULK) {
      if (i < nparts) {
        const int s = i / PARTS, hh = i - s * PARTS;
        const int buf = i % BUFS;
        const uint32_t bar = bar_s + (uint32_t)(buf * sizeof(uint64_t));
        if (lane == 0)
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16u * PHC * 8u) : "memory");
        __syncwarp();
        if (lane < 16) {
          const char* src = s < ublocks ? lu_brow + ((size_t)step_block(s) * PB + (size_t)hh * PHC) * sizeof(double)
                            : (s == ublocks ? w_brow : m_brow) + (size_t)hh * PHC * sizeof(double);
          bulk_copy_g2s(lbuf_s + (uint32_t)(buf * PB * RS * sizeof(double)) + brow_dst, src, PHC * 8u, bar);
        }
      }
      return;
    }
    if (i < nparts) {
      const int s = i / PARTS, hh = i - s * PARTS;
      const uint32_t dbase = lbuf_s + (uint32_t)((i % BUFS) * PB * PHC * sizeof(double));
      if (s < ublocks) {
        const int cbase = step_block(s) * PB + hh * PHC;
        const bool col_ok = cbase + 2 * cc < N;  // n is even: a pair of columns is inside or outside as a whole
        const char* src = lu_row0 + (size_t)cbase * sizeof(double);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const bool ok = col_ok && ((row_ok >> q) & 1u);
          cp_async16(dbase + dst_off[q], ok ? (const void*)src : (const void*)LU, ok ? 16 : 0);
          src += lu_stride;
        }
      } else {
        const char* src = (s == ublocks ? w_row0 : m_row0) + (size_t)hh * PHC * sizeof(double);
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          cp_async16(dbase + dst_off[q], src, 16);
          src += W_STRIDE;
        }
      }
    }
    cp_async_commit();  // always a group (possibly empty): the wait count below stays uniform
  };

  // acc[mt][nt][0..1] = element (row 16 warp + 8 mt + gid, columns c0 + 8 nt + 2 tig, + 1) of this block's right-hand sides
  double acc[KS][2][NT][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int rr = 16 * warp + 8 * mt + gid, col = c0 + 8 * nt + 2 * tig + e;
        double v = 0.0;
        if (rr < nr && col < nx)  // X = B(piv,:), lu.rs:246-254
          v = FORWARD ? B[(size_t)(piv ? piv[r0 + rr] : (uint64_t)(r0 + rr)) * nx + col] : X[(size_t)(r0 + rr) * nx + col];
        acc[0][mt][nt][e] = v;
#pragma unroll
        for (int q = 1; q < KS; ++q) acc[q][mt][nt][e] = 0.0;
      }
  // d += (rows of buffer buf) * (rows [PHC hh, PHC hh + PHC) of Xs)
  auto mma_part = [&](int buf, int hh, double (&d)[KS][2][NT][2]) {
    const int row_a0 = 16 * warp + gid;
    const double* La0 = Lbuf + (size_t)(buf * PB + row_a0) * RS;
    const double* La1 = La0 + 8 * RS;
    // 64-bit shared loads are served per HALF warp (gid 0..3 / 4..7, all tig): its four rows x two 16-byte chunks must land
    // in eight different chunk positions of a 128-byte bank period -> chunk index XOR 2 * (row & 3) (XOR with row & 7 made
    // rows r and r ^ 1 collide: ncu counted 4 wavefronts per fragment load instead of 2)
    const int sw = (row_a0 & 3) << 1;  // == ((row_a0 + 8) & 3) << 1
    const double* Xb = Xs + (size_t)(hh * PHC + tig) * XLD + gid;
#pragma unroll
    for (int ks = 0; ks < PHC / 4; ++ks) {
      const int col = 4 * ks + tig;
      const int off = BULK ? col : (((col >> 1) ^ sw) << 1) + (col & 1);  // padded rows (BULK) or swizzled chunks
      const double a0 = La0[off], a1 = La1[off];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const double bv = Xb[(size_t)(4 * ks) * XLD + 8 * nt];
        dmma884(d[ks % KS][0][nt][0], d[ks % KS][0][nt][1], a0, bv);
        dmma884(d[ks % KS][1][nt][0], d[ks % KS][1][nt][1], a1, bv);
      }
    }
  };

  auto stamp = [&](int slot) {
    if (dbg && t == 0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns));
      dbg[(size_t)blockIdx.x * 8 + slot] = ns;
    }
  };
  if constexpr (BULK) {
    if (lane == 0)
      for (int b = 0; b < BUFS; ++b) mbar_init(bars + warp * BUFS + b, 1);
    mbar_fence_init();
    __syncwarp();
  }
  issue_part(0);
  issue_part(1);
  issue_part(2);
  double res[KS][2][NT][2];
  for (int i = 0; i < nparts; ++i) {
    const int s = i / PARTS, hh = i - s * PARTS;
    const bool own = s == ublocks;   // the parts of the inverted diagonal block
    const bool last = s > ublocks;   // (comb) the parts of M_g against the last awaited block
    if (hh == 0) {
      __syncthreads();  // every warp has finished reading the previous block from Xs
      if (!own) {
        // ---- block kb of the solution: poll the producer's flagged words (WPT per thread), store MINUS the values ----
        const XWord* src = xbuf + (size_t)step_block(last ? nsteps - 1 : s) * Cfg::WORDS;
        double v[WPT];
        bool ok[WPT];
        {
          bool first = false;
          while (!first) first = LL<double>::load(src + t, tag, v[0]);  // one word per thread while nothing has arrived
          ok[0] = true;
        }
        bool all;
#pragma unroll
        for (int u = 1; u < WPT; ++u) ok[u] = false;
        do {
          all = true;
#pragma unroll
          for (int u = 1; u < WPT; ++u)
            if (!ok[u]) {
              ok[u] = LL<double>::load(src + t + SWEEP_THREADS * u, tag, v[u]);
              all = all && ok[u];
            }
        } while (!all);
#pragma unroll
        for (int u = 0; u < WPT; ++u) {
          const int idx = t + SWEEP_THREADS * u;
          Xs[(size_t)(idx / NXC) * XLD + (idx % NXC)] = -v[u];
        }
        if (last || (!comb && s == nsteps - 1)) stamp(0);  // the last awaited block has arrived
      } else {
        stamp(1);  // updates done
        // ---- X_g = W_g * (right-hand sides of this block): the same MMA loop on the inverted diagonal block ----
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int rr = 16 * warp + 8 * mt + gid, col = 8 * nt + 2 * tig;
            double r0v = acc[0][mt][nt][0], r1v = acc[0][mt][nt][1];
#pragma unroll
            for (int q = 1; q < KS; ++q) {
              r0v += acc[q][mt][nt][0];
              r1v += acc[q][mt][nt][1];
            }
            *reinterpret_cast<double2*>(&Xs[(size_t)rr * XLD + col]) = make_double2(r0v, r1v);
#pragma unroll
            for (int q = 0; q < KS; ++q) res[q][mt][nt][0] = res[q][mt][nt][1] = 0.0;
          }
      }
      __syncthreads();  // Xs is complete
      if (own) stamp(2);
    }
    if constexpr (BULK) {
      mbar_wait(bars + warp * BUFS + i % BUFS, (uint32_t)((i / BUFS) & 1));  // the warp's slice of part i has landed
    } else {
      cp_async_wait<BUFS - 1>();  // this thread's groups are committed in order: all but the two youngest have landed
      __syncwarp();               // ... and so have the other lanes' chunks of this warp's slice of part i
    }
    if (own || last) mma_part(i % BUFS, hh, res);  // (comb) res = W_g R - M_g x_last: Xs holds MINUS x_last
    else mma_part(i % BUFS, hh, acc);
    __syncwarp();               // the warp is done with its slice of buffer i % 3
    issue_part(i + BUFS);
  }
  // publish: flagged words for the CTAs that still need this block, plain values into X
  const bool has_readers = FORWARD ? g + 1 < G : g > 0;
  XWord* dst = xbuf + (size_t)g * Cfg::WORDS;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int rr = 16 * warp + 8 * mt + gid, cl = 8 * nt + 2 * tig + e, col = c0 + cl;
        double v = res[0][mt][nt][e];
#pragma unroll
        for (int q = 1; q < KS; ++q) v += res[q][mt][nt][e];
        if (has_readers) LL<double>::store(dst + rr * NXC + cl, (rr < nr && col < nx) ? v : 0.0, tag);
        if (rr < nr && col < nx) X[(size_t)(r0 + rr) * nx + col] = v;
      }
  stamp(3);  // block product done, solution published
  stamp(4);
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

// Two persistent sweeps (see solve_sweep_kernel): X = Umat^-1 (Lmat^-1 B(piv,:)) where only the blocks of Lmat strictly
// below and of Umat strictly above the block diagonal are read, the diagonal blocks arriving inverted in WL / WU
// ([G][128][128]).  Lmat == Umat == packed LU for the LU solve; L and L' for Cholesky.  piv_dev may be null (identity).
// Caller guarantees: nx <= 16, n even, 16-byte aligned matrices, ceil(n / 128) <= SM count, cooperative launch support.
namespace {
struct SweepSide {
  cudaStream_t s = nullptr;
  cudaEvent_t e_in = nullptr, e_out = nullptr;
};
int sweep_side(int device, SweepSide** out) {
  static thread_local SweepSide side[64];
  LA_REQUIRE(device >= 0 && device < 64, "device ordinal out of range");
  SweepSide& s = side[device];
  if (!s.s) {
    LA_CUDA_TRY(cudaStreamCreateWithFlags(&s.s, cudaStreamNonBlocking));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_in, cudaEventDisableTiming));
    LA_CUDA_TRY(cudaEventCreateWithFlags(&s.e_out, cudaEventDisableTiming));
  }
  *out = &s;
  return LA_OK;
}
}  // namespace

// wu_mode >= 0: the inverted diagonal blocks of Umat are NOT in wu yet -- they are computed here (tri_block_inverses mode
// wu_mode) on a side stream together with the U-side neighbour products, under the forward sweep.
int tri_sweeps_dev(const double* Lmat, const double* Umat, size_t n, const uint64_t* piv_dev, const double* B, size_t nx,
                   double* X, const double* wl, const double* wu, cudaStream_t st, int wu_mode) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  const int G = (int)((n + PB - 1) / PB);
  void* fl = nullptr;
  const size_t fbytes = sizeof(unsigned) * (size_t)G * G;
  LA_TRY(scratch_get(ctx->device, 13, 2 * fbytes, &fl));
  LA_CUDA_TRY(cudaMemsetAsync(fl, 0, 2 * fbytes, st));
  LA_CUDA_TRY(cudaFuncSetAttribute(solve_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM));
  LA_CUDA_TRY(cudaFuncSetAttribute(solve_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWEEP_SMEM));
  const double* lu_p = Lmat;
  const double* uu_p = Umat;
  const double* b_p = B;
  double* x_p = X;
  size_t nn = n;
  int nxi = (int)nx;
  unsigned* f0 = (unsigned*)fl;
  unsigned* f1 = f0 + (size_t)G * G;
  // LA_SOLVE_TRACE=1: phase timestamps of every CTA's last step, printed after a sync (diagnostic)
  static const int trace = getenv("LA_SOLVE_TRACE") ? atoi(getenv("LA_SOLVE_TRACE")) : 0;
  unsigned long long* d0 = nullptr;
  unsigned long long* d1 = nullptr;
  if (trace) {
    void* dp = nullptr;
    LA_TRY(scratch_get(ctx->device, 14, 4 * sizeof(unsigned long long) * 8 * G, &dp));  // [sweep][chain <= 2][G][8]
    LA_CUDA_TRY(cudaMemsetAsync(dp, 0, 4 * sizeof(unsigned long long) * 8 * G, st));
    d0 = (unsigned long long*)dp;
    d1 = d0 + 16 * G;
  }
  static const int old_sweep = getenv("LA_SOLVE_OLD_SWEEP") ? atoi(getenv("LA_SOLVE_OLD_SWEEP")) : 0;  // A/B knob
  if (!old_sweep) {
    // flagged-word exchange buffers of the two sweeps; tags are unique per launch pair of this host thread (never 0)
    void* xb = nullptr;
    const int max_g = ctx->sm_count;  // cooperative grid: one block row per SM
    const int Gp = G < max_g ? G : max_g;
    // 9..16 right-hand sides: two independent chains of 8 (two CTAs per SM); <= 8: one chain of 8; LA_SOLVE_CHAINS=1 keeps
    // the single 16-column chain (A/B knob)
    static const int chains_knob = getenv("LA_SOLVE_CHAINS") ? atoi(getenv("LA_SOLVE_CHAINS")) : 2;
    const bool narrow = nx <= 8 || chains_knob >= 2;
    const int chains = (narrow && nx > 8) ? 2 : 1;
    const size_t words = (size_t)Gp * PB * 16;  // per sweep, either layout
    LA_TRY(scratch_get(ctx->device, 38, 2 * words * sizeof(XWord), &xb));
    static thread_local unsigned call_tag = 0;
    XWord* xb0 = (XWord*)xb;
    XWord* xb1 = xb0 + words;
    // bulk (TMA engine) copies, one 256 / 512-byte row segment per instruction: OFF by default -- measured 2.45 ms against
    // 1.92 ms for the 16-byte cp.async form at n = 16384 (128 small bulk requests per part and CTA cost more than the issue
    // slots they save); LA_SOLVE_BULK=1 keeps the experiment reproducible (needs n a multiple of 128)
    static const int bulk_knob = getenv("LA_SOLVE_BULK") ? atoi(getenv("LA_SOLVE_BULK")) : 0;
    const bool bulk = bulk_knob && n % PB == 0 && ((uintptr_t)wl % 16 == 0) && ((uintptr_t)wu % 16 == 0);
    const int smem_bytes = narrow ? (bulk ? SweepCfg<8, true>::SMEM : SweepCfg<8>::SMEM)
                                  : (bulk ? SweepCfg<16, true>::SMEM : SweepCfg<16>::SMEM);
    const void* kf = narrow ? (bulk ? (const void*)solve_sweep_mma_kernel<true, 8, true> : (const void*)solve_sweep_mma_kernel<true, 8, false>)
                            : (bulk ? (const void*)solve_sweep_mma_kernel<true, 16, true> : (const void*)solve_sweep_mma_kernel<true, 16, false>);
    const void* kb = narrow ? (bulk ? (const void*)solve_sweep_mma_kernel<false, 8, true> : (const void*)solve_sweep_mma_kernel<false, 8, false>)
                            : (bulk ? (const void*)solve_sweep_mma_kernel<false, 16, true> : (const void*)solve_sweep_mma_kernel<false, 16, false>);
    LA_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    LA_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    LA_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    LA_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    size_t ldm = n;
    // precombined neighbour blocks M_g = W_g * T[g][g -+ 1] for both triangles (one launch each; LA_SOLVE_COMBINE=0: off)
    static const int combine_knob = getenv("LA_SOLVE_COMBINE") ? atoi(getenv("LA_SOLVE_COMBINE")) : 1;
    const double* mcl = nullptr;
    const double* mcu = nullptr;
    SweepSide* side = nullptr;
    LA_TRY(sweep_side(ctx->device, &side));
    const bool combine = combine_knob && G > 1;
    const bool use_side = combine || wu_mode >= 0;
    double* m1 = nullptr;
    if (combine) {
      void* mp = nullptr;
      LA_TRY(scratch_get(ctx->device, 44, sizeof(double) * 2 * (size_t)G * PB * PB, &mp));
      double* m0 = (double*)mp;
      m1 = m0 + (size_t)G * PB * PB;
      sweep_combine_kernel<true><<<G, 256, 0, st>>>(Lmat, n, (int)n, wl, m0);
      LA_CUDA_TRY(cudaGetLastError());
      mcl = m0;
      mcu = m1;
    }
    if (use_side) LA_CUDA_TRY(cudaEventRecord(side->e_in, st));  // everything the U side reads is ready here
    // U-side preparation: queued AFTER the first forward sweep has been launched, so that it fills the SMs that sweep
    // leaves free instead of delaying it; needed by the first backward sweep
    bool side_queued = false;
    auto queue_side = [&]() -> int {
      if (!use_side || side_queued) return LA_OK;
      side_queued = true;
      LA_CUDA_TRY(cudaStreamWaitEvent(side->s, side->e_in, 0));
      if (wu_mode >= 0) LA_TRY(tri_block_inverses<double>(Umat, n, wu_mode, 0, G, const_cast<double*>(wu), 0, side->s));
      if (combine) {
        sweep_combine_kernel<false><<<G, 256, 0, side->s>>>(Umat, n, (int)n, wu, m1);
        LA_CUDA_TRY(cudaGetLastError());
      }
      LA_CUDA_TRY(cudaEventRecord(side->e_out, side->s));
      return LA_OK;
    };
    auto join_side = [&]() -> int {  // before the first backward sweep
      LA_TRY(queue_side());
      if (use_side) LA_CUDA_TRY(cudaStreamWaitEvent(st, side->e_out, 0));
      return LA_OK;
    };
    auto sweep = [&](bool fwd, const double* M, size_t rows, const uint64_t* pv, const double* rhs, double* out,
                     const double* winv, const double* mcomb) -> int {
      call_tag += 1;
      if (call_tag == 0) call_tag = 1;
      unsigned tg = call_tag;
      XWord* xw = fwd ? xb0 : xb1;
      int gp = (int)((rows + PB - 1) / PB);
      unsigned long long* dg = trace ? (fwd ? d0 : d1) : nullptr;
      void* args[] = {&M, &ldm, &rows, &pv, &rhs, &out, &nxi, &gp, &xw, &tg, &winv, &mcomb, &dg};
      LA_CUDA_TRY(cudaLaunchCooperativeKernel(fwd ? kf : kb, dim3(gp * chains), dim3(SWEEP_THREADS), args, smem_bytes, st));
      return LA_OK;
    };
    if (G <= max_g) {
      LA_TRY(sweep(true, Lmat, n, piv_dev, B, X, wl, mcl));
      LA_TRY(queue_side());
      LA_TRY(join_side());
      LA_TRY(sweep(false, Umat, n, piv_dev, B, X, wu, mcu));
      if (trace) {  // chain 0 only: phase timestamps of every CTA's last step
        std::vector<unsigned long long> hbuf(32 * (size_t)G);
        LA_CUDA_TRY(cudaStreamSynchronize(st));
        LA_CUDA_TRY(cudaMemcpy(hbuf.data(), d0, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int ph = 0; ph < 2; ++ph) {
          const unsigned long long* hb = hbuf.data() + (size_t)ph * 16 * G;
          const unsigned long long t0 = ph == 0 ? hb[3] : hb[(size_t)(G - 1) * 8 + 3];
          for (int g = 0; g < G; g += (G > 32 ? G / 32 : 1)) {
            const unsigned long long* e = hb + (size_t)g * 8;
            auto us = [&](int k) { return e[k] ? ((double)e[k] - (double)t0) * 1e-3 : 0.0; };
            fprintf(stderr, "sweep2 %s cta %3d: last block arrived %8.2f us | own product started %8.2f | published %8.2f "
                            "(arrival -> publication %5.2f us)\n",
                    ph == 0 ? "fwd" : "bwd", g, us(0), us(2), us(3), e[0] ? us(3) - us(0) : 0.0);
          }
        }
      }
      return LA_OK;
    }
    // More block rows than SMs: leading parts of sm_count block rows each; between the parts the solved rows update the
    // remaining right-hand sides with one GEMM (the triangular system is block triangular in the parts as well).
    {
      size_t blocks = (n * nx + 255) / 256;
      const size_t cap = (size_t)ctx->sm_count * 8;
      if (piv_dev) gather_rows_kernel<double><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(B, X, piv_dev, n, nx);
      else LA_CUDA_TRY(cudaMemcpyAsync(X, B, n * nx * sizeof(double), cudaMemcpyDeviceToDevice, st));
      LA_CUDA_TRY(cudaGetLastError());
    }
    const size_t part = (size_t)max_g * PB;
    for (size_t p0 = 0; p0 < n; p0 += part) {  // forward
      const size_t rows = n - p0 < part ? n - p0 : part, p1 = p0 + rows;
      LA_TRY(sweep(true, Lmat + p0 * n + p0, rows, nullptr, X + p0 * nx, X + p0 * nx, wl + (p0 / PB) * PB * PB,
                   mcl ? mcl + (p0 / PB) * PB * PB : nullptr));
      LA_TRY(queue_side());
      if (p1 < n)
        LA_TRY(gemm_dev<double>(Lmat + p1 * n + p0, n, X + p0 * nx, nx, X + p1 * nx, nx, n - p1, rows, nx, LA_GEMM_SUB, st));
    }
    const size_t nparts = (n + part - 1) / part;
    LA_TRY(join_side());
    for (size_t ip = nparts; ip-- > 0;) {  // backward
      const size_t p0 = ip * part, rows = n - p0 < part ? n - p0 : part;
      LA_TRY(sweep(false, Umat + p0 * n + p0, rows, nullptr, X + p0 * nx, X + p0 * nx, wu + (p0 / PB) * PB * PB,
                   mcu ? mcu + (p0 / PB) * PB * PB : nullptr));
      if (p0 > 0) LA_TRY(gemm_dev<double>(Umat + p0, n, X + p0 * nx, nx, X, nx, p0, rows, nx, LA_GEMM_SUB, st));
    }
    return LA_OK;
  }
  LA_REQUIRE(G <= ctx->sm_count, "la_lu_solve: the first-generation sweep kernel needs ceil(n / 128) <= SM count");
  if (wu_mode >= 0) LA_TRY(tri_block_inverses<double>(Umat, n, wu_mode, 0, G, const_cast<double*>(wu), 0, st));
  void* a0[] = {&lu_p, &nn, &piv_dev, &b_p, &x_p, &nxi, &f0, &wl, &d0};
  void* a1[] = {&uu_p, &nn, &piv_dev, &b_p, &x_p, &nxi, &f1, &wu, &d1};
  LA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)solve_sweep_kernel<true>, dim3(G), dim3(SWEEP_THREADS), a0,
                                          SWEEP_SMEM, st));
  LA_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)solve_sweep_kernel<false>, dim3(G), dim3(SWEEP_THREADS), a1,
                                          SWEEP_SMEM, st));
  if (trace) {
    std::vector<unsigned long long> hbuf(32 * (size_t)G);
    LA_CUDA_TRY(cudaStreamSynchronize(st));
    LA_CUDA_TRY(cudaMemcpy(hbuf.data(), d0, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int ph = 0; ph < 2; ++ph) {
      const unsigned long long* hb = hbuf.data() + (size_t)ph * 16 * G;
      const unsigned long long t0 = ph == 0 ? hb[3] : hb[(size_t)(G - 1) * 8 + 3];
      for (int g = 0; g < G; g += (G > 16 ? G / 16 : 1)) {
        const unsigned long long* e = hb + (size_t)g * 8;
        fprintf(stderr, "solve %s cta %3d: flag %8.2f us | +Y/update %6.2f | +Lwait %5.2f | +block product %6.2f | +fence %5.2f\n",
                ph == 0 ? "fwd" : "bwd", g, e[0] ? (double)(e[0] - t0) * 1e-3 : 0.0,
                e[0] ? (double)(e[1] - e[0]) * 1e-3 : 0.0, (double)(e[2] - e[1]) * 1e-3, (double)(e[3] - e[2]) * 1e-3,
                (double)(e[4] - e[3]) * 1e-3);
      }
    }
  }
  return LA_OK;
}

template <typename T>
int lu_solve_dev(const T* LU, size_t n, const uint64_t* piv_dev, const T* B, size_t nx, T* X, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(LU && piv_dev && B && X, "la_lu_solve: null pointer");
  LA_REQUIRE(n > 0 && nx > 0, "la_lu_solve: zero dimension");
  LA_REQUIRE(n < (1u << 30) && nx < (1u << 30), "la_lu_solve: dimension too large");
  LA_REQUIRE((const void*)B != (const void*)X, "la_lu_solve: B and X must not alias");
  const int N = (int)n;
  // few right-hand sides, fp64, cp.async-addressable: one persistent kernel per sweep
  if constexpr (std::is_same<T, double>::value) {
    const int G = (N + PB - 1) / PB;
    static const int no_sweep = getenv("LA_SOLVE_NO_SWEEP") ? atoi(getenv("LA_SOLVE_NO_SWEEP")) : 0;  // debug knob
    if (!no_sweep && nx <= 16 && N >= 4 * PB && N % 2 == 0 && (uintptr_t)LU % 16 == 0 && ctx->coop) {
      void* wbuf = nullptr;
      LA_TRY(scratch_get(ctx->device, 15, sizeof(double) * 2 * (size_t)G * PB * PB, &wbuf));
      double* wl = (double*)wbuf;
      double* wu = wl + (size_t)G * PB * PB;
      LA_TRY(tri_block_inverses<double>(LU, n, 0, 0, G, wl, 0, st));  // inv(L_bb); inv(U_bb) is made under the forward sweep
      LA_TRY(tri_sweeps_dev(LU, LU, n, piv_dev, B, nx, X, wl, wu, st, 1));
      return LA_OK;
    }
  }
  // many right-hand sides (inverse = n of them), fp64, TMA-addressable: inverted 128 x 128 diagonal blocks turn both
  // sweeps into DMMA GEMMs -- per block row X_b = W_b * X_b (in place, one tile row) and X_rest -= LU[rest][b] * X_b
  if constexpr (std::is_same<T, double>::value) {
    static const int no_gemm = getenv("LA_SOLVE_NO_GEMM") ? atoi(getenv("LA_SOLVE_NO_GEMM")) : 0;  // debug knob
    if (!no_gemm && nx > 16 && N >= 4 * PB && N % 2 == 0 && nx % 2 == 0 && (uintptr_t)LU % 16 == 0 &&
        (uintptr_t)X % 16 == 0) {
      const int G = (N + PB - 1) / PB;
      void* wbuf = nullptr;
      LA_TRY(scratch_get(ctx->device, 15, sizeof(double) * 2 * (size_t)G * PB * PB, &wbuf));
      double* WL = (double*)wbuf;
      double* WU = WL + (size_t)G * PB * PB;
      LA_TRY(lu_diag_block_inverses(LU, n, WL, WU, st));
      {
        size_t blocks = (n * nx + 255) / 256;
        size_t cap = (size_t)ctx->sm_count * 8;
        gather_rows_kernel<T><<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(B, X, piv_dev, n, nx);
        LA_CUDA_TRY(cudaGetLastError());
      }
      for (int b = 0; b < G; ++b) {  // forward: L * Y = B(piv,:)
        const size_t r0 = (size_t)b * PB, nr = (n - r0 < (size_t)PB) ? (n - r0) : (size_t)PB;
        double* Xb = X + r0 * nx;
        LA_TRY(gemm_f64_tensor(WL + (size_t)b * PB * PB, PB, Xb, nx, Xb, nx, nr, nr, nx, LA_GEMM_ASSIGN, st));
        if (r0 + nr < n)
          LA_TRY(gemm_f64_tensor(LU + (r0 + nr) * n + r0, n, Xb, nx, X + (r0 + nr) * nx, nx, n - r0 - nr, nr, nx,
                                 LA_GEMM_SUB, st));
      }
      for (int b = G - 1; b >= 0; --b) {  // backward: U * X = Y
        const size_t r0 = (size_t)b * PB, nr = (n - r0 < (size_t)PB) ? (n - r0) : (size_t)PB;
        double* Xb = X + r0 * nx;
        LA_TRY(gemm_f64_tensor(WU + (size_t)b * PB * PB, PB, Xb, nx, Xb, nx, nr, nr, nx, LA_GEMM_ASSIGN, st));
        if (r0 > 0) LA_TRY(gemm_f64_tensor(LU + r0, n, Xb, nx, X, nx, r0, nr, nx, LA_GEMM_SUB, st));
      }
      return LA_OK;
    }
  }
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
