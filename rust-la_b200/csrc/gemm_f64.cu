// gemm_f64.cu -- C[m x n] (=, -=, +=) A[m x k] * B[k x n], fp64, row-major, for sm_100a.
//
// Replaces the i-j-k loop nest of `impl Mul<&Matrix<T>> for &Matrix<T>` (reference src/matrix/mod.rs:965-973) and
// `Matrix::mmul` (src/matrix/mmatrix.rs:87-95); with mode LA_GEMM_SUB it is also the trailing update of the blocked
// LU (the i > j part of src/decomp/lu.rs:122-129).
//
// Design (B200): tcgen05.mma has no fp64 kind, so fp64 runs on the DMMA pipe: mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4
// (every wider PTX f64 shape is split by ptxas into 8x8x4 on sm_100a).  Accumulators therefore live in registers:
// a 128x128 CTA tile = 128 KiB of accumulators = 128 regs/thread over 8 consumer warps (64x32 warp tiles).
// Operands are fed by TMA (cp.async.bulk.tensor, 128B swizzle) into a 4-stage shared-memory ring by a dedicated
// producer warp; consumers wait on `full` mbarriers and release stages through `empty` mbarriers.
//
// Shared-memory fragment trick.  The MMA contracts over 4 k-indices and produces 8 columns, but nothing forces those
// to be *adjacent* indices.  Each thread issues one LDS.128 for A at (row, k = 2t, 2t+1) and uses .x for the MMA that
// contracts over the even k of an 8-wide k chunk and .y for the MMA over the odd k; for B one LDS.128 at (k, n = 2g,
// 2g+1) feeds an "even columns" and an "odd columns" 8-wide tile.  With the TMA 128B swizzle both access patterns are
// bank-conflict free per quarter warp, every load is 16 bytes, and each thread ends up owning 4 CONSECUTIVE output
// columns (32 contiguous bytes per row in the epilogue).  Rows inside an 8-row MMA tile are permuted
// (g -> (g>>1)|((g&1)<<2)) for the same reason.
#include <stdlib.h>

#ifndef LA_GEMM_VARIANT
#define LA_GEMM_VARIANT 0
#endif
#include "la_common.cuh"

namespace la {
namespace {

constexpr int BM = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 8;  // 16 KiB: 128 rows x 128 B
constexpr int B_BOX_BYTES = BK * 16 * 8;    // 2 KiB: [16 k-rows x 128 B] holds 16 columns of B
// Two tile configurations of the same kernel:
//   BN = 128, 8 warps, 1 CTA/SM : least L2->smem traffic per flop; used for deep K (Mul at large n).
//   BN =  64, 4 warps, 2 CTA/SM : the epilogue of one CTA (C tile read-modify-write, HBM bound) overlaps the main loop
//                                 of the other; used for shallow K (the LU trailing update, K = nb).
// Every warp is a consumer (64 x 32 warp tile, 128 accumulator registers per thread).  The TMA producer role is folded
// into the consumers: the lane 0 of warp (p mod #warps) issues the loads of k-tile p, STAGES-1 tiles ahead of the
// math.  (An earlier version used a dedicated producer warpgroup with setmaxnreg register donation; CTAs of that
// kernel produced corrupted tiles whenever CTAs of ANOTHER kernel were co-resident on the SM -- reproducible with
// tools/overlap_stress.py -- which rules it out for the LU look-ahead, where the panel kernel shares the SMs.)
template <int BN_>
struct TileCfg {
  static constexpr int WARPS = BN_ / 16;                          // 2 (M) x BN/32 (N) warps of 64 x 32
  static constexpr int THREADS = WARPS * 32;
  static constexpr int CTAS_PER_SM = BN_ == 128 ? 1 : 2;
  static constexpr int B_STAGE_BYTES = BK * BN_ * 8;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;  // + barriers + alignment slack
};
constexpr double SMALL_GEMM_MNK = 128.0 * 128.0 * 128.0;  // <= this many multiply-adds: bit-exact CUDA-core kernel
constexpr size_t SHALLOW_K = 1024;
constexpr int GROUP_M = 16;  // tile rasterisation: GROUP_M tile-rows share each B tile-column while it is hot in L2

// Epilogue helpers.  A lane owns pairs of adjacent columns; `col` is even, so a pair is 16-byte aligned (ldc is even).
__device__ __forceinline__ double2 load_pair(const double* __restrict__ C, size_t ldc, int row, int col, int M, int N) {
  double2 v = make_double2(0.0, 0.0);
  if (row < M) {
    const double* p = C + (size_t)row * ldc + col;
    // L2-only loads (ld.global.cg): C was written by other kernels, possibly while a concurrent grid (the LU look-ahead
    // panel) kept this SM's L1 alive across the kernel boundary; the data is streamed once anyway.
    if (col + 1 < N)
      v = __ldcg(reinterpret_cast<const double2*>(p));
    else if (col < N)
      v.x = __ldcg(p);
  }
  return v;
}
__device__ __forceinline__ void store_pair(double* __restrict__ C, size_t ldc, int row, int col, int M, int N, double2 v) {
  if (row < M) {
    double* p = C + (size_t)row * ldc + col;
    if (col + 1 < N)
      *reinterpret_cast<double2*>(p) = v;
    else if (col < N)
      p[0] = v.x;
  }
}

template <int MODE, int BN>
__global__ void __launch_bounds__(TileCfg<BN>::THREADS, TileCfg<BN>::CTAS_PER_SM)
gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    double* __restrict__ C, size_t ldc, int M, int N, int K, int tiles_m, int tiles_n, int stagger_lo,
                    int stagger_hi, unsigned stagger_ns, int kt_per_slice, size_t c_slab, int lower_only) {
  using Cfg = TileCfg<BN>;
  // Shallow-K launches run two CTAs per SM so that one CTA's epilogue (HBM-bound C tile read-modify-write) overlaps the
  // other's main loop -- which only works if the two are out of phase.  CTAs launched together stay in lock step, so
  // the second resident slot of the first wave starts half a tile late; every later CTA inherits the phase of the CTA
  // whose slot it takes over.
  if ((int)blockIdx.x >= stagger_lo && (int)blockIdx.x < stagger_hi) __nanosleep(stagger_ns);
  constexpr int WARPS = Cfg::WARPS;
  constexpr int STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int WARPS_N = BN / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // grouped rasterisation of the 1-D grid onto (tile_m, tile_n)
  const int tile = blockIdx.x;
  const int tiles_per_group = GROUP_M * tiles_n;
  const int group = tile / tiles_per_group;
  const int first_m = group * GROUP_M;
  const int group_rows = min(GROUP_M, tiles_m - first_m);
  const int tile_m = first_m + (tile % tiles_per_group) % group_rows;
  const int tile_n = (tile % tiles_per_group) / group_rows;
  const int m0 = tile_m * BM;
  const int n0 = tile_n * BN;
  // symmetric updates (Cholesky's A22 -= L21 L21'): tiles that lie entirely above the diagonal of C are not computed
  if (lower_only && n0 > m0 + BM - 1) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // split-K (gridDim.y slices): slice y contracts k-tiles [y * kt_per_slice, ...) into its own copy of C at C + y * c_slab
  // (tall-skinny products such as the QR's W = V' A have too few output tiles to fill the machine otherwise)
  const int kt_first = (int)blockIdx.y * kt_per_slice;
  const int ktiles = min(kt_per_slice, (K + BK - 1) / BK - kt_first);
  C += (size_t)blockIdx.y * c_slab;
  const uint32_t smem_base = smem_u32(smem);

  // C -= P / C += P: pull the C tile into L2 now so the epilogue's loads do not pay an HBM round trip
  if (MODE != LA_GEMM_ASSIGN) {
    constexpr int LINES_PER_ROW = BN * 8 / 128;
    for (int idx = threadIdx.x; idx < BM * LINES_PER_ROW; idx += Cfg::THREADS) {
      const int r = m0 + idx / LINES_PER_ROW, c = n0 + (idx % LINES_PER_ROW) * 16;
      if (r < M && c < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (size_t)r * ldc + c));
    }
  }

  // TMA loads of k-tile p into stage p % STAGES (one elected lane).  A: box 16 k (inner) x 128 rows; B: BN/16 boxes
  // of 16 n (inner) x 16 k-rows.  Out-of-bounds elements are zero-filled and still counted in the transaction bytes.
  auto issue_tile = [&](int p) {
    const int s = p % STAGES;
    uint8_t* sA = smem + s * STAGE_BYTES;
    uint8_t* sB = sA + A_STAGE_BYTES;
    mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
    tma_load_2d(sA, &tmA, &full[s], (kt_first + p) * BK, m0);
#pragma unroll
    for (int j = 0; j < BN / 16; ++j) tma_load_2d(sB + j * B_BOX_BYTES, &tmB, &full[s], n0 + j * 16, (kt_first + p) * BK);
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int p = 0; p < STAGES - 1 && p < ktiles; ++p) issue_tile(p);  // fresh stages: nothing to wait for
  }

  // ===== 2 (M) x BN/32 (N) warps, warp tile 64 x 32 =====
  const int wm = warp / WARPS_N;
  const int wn = warp % WARPS_N;
  const int g = lane >> 2;
  const int t = lane & 3;
  const int x = (g >> 1) | ((g & 1) << 2);  // row of the 8-row MMA tile this lane's group supplies

  double acc[8][2][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][j][c] = 0.0;

  // byte offsets inside a stage (swizzle: 16B-chunk index ^= row & 7)
  const uint32_t a_row = (uint32_t)(wm * 64 + x) * 128u;
  uint32_t a_chunk[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) a_chunk[h] = (uint32_t)(((h * 4 + t) ^ x) << 4);
  uint32_t b_off[2][2];  // [k-set s][chunk h] for n-block j = 0; j adds B_BOX_BYTES
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h)
      b_off[s][h] = (uint32_t)A_STAGE_BYTES + (uint32_t)(wn * 2) * B_BOX_BYTES + (uint32_t)(h * 8 + 2 * t + s) * 128u +
                    (uint32_t)((g ^ (2 * t + s)) << 4);

  // Fragments are double-buffered in registers: while the DMMAs of one 8-wide k-chunk run, the 12 LDS.128 of the next
  // chunk are already in flight into the other buffer, so no fragment register is rewritten within ~1000 cycles of its
  // last use and every load has a whole chunk of math to land.
  double2 af[2][8];
  double2 bf[2][2][2];
  auto load_frags = [&](int buf, uint32_t st, int h) {
#pragma unroll
    for (int i = 0; i < 8; ++i) af[buf][i] = lds_f64x2(st + a_row + i * 1024 + a_chunk[h]);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int ss = 0; ss < 2; ++ss) bf[buf][j][ss] = lds_f64x2(st + b_off[ss][h] + j * B_BOX_BYTES);
  };
  auto mma_chunk = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        dmma884(acc[i][j][0], acc[i][j][2], af[buf][i].x, bf[buf][j][0].x);  // even k, even columns
        dmma884(acc[i][j][1], acc[i][j][3], af[buf][i].x, bf[buf][j][0].y);  // even k, odd columns
        dmma884(acc[i][j][0], acc[i][j][2], af[buf][i].y, bf[buf][j][1].x);  // odd k, even columns
        dmma884(acc[i][j][1], acc[i][j][3], af[buf][i].y, bf[buf][j][1].y);  // odd k, odd columns
      }
  };

  mbar_wait(&full[0], 0);
  load_frags(0, smem_base, 0);
  for (int kt = 0; kt < ktiles; ++kt) {
    // producer duty: k-tile p = kt + STAGES - 1 goes into the stage that k-tile kt - 1 occupied; its issuer first waits
    // until every warp has released that stage
    {
      const int p = kt + STAGES - 1;
      if (p < ktiles && (p % WARPS) == warp && lane == 0) {
        if (p >= STAGES) mbar_wait(&empty[p % STAGES], ((p / STAGES) - 1) & 1);
        issue_tile(p);
      }
      __syncwarp();
    }
    const int s = kt % STAGES;
    const uint32_t st = smem_base + (uint32_t)s * STAGE_BYTES;
    load_frags(1, st, 1);   // second chunk of this k-tile
    mma_chunk(0);
    if (kt + 1 < ktiles) {  // first chunk of the next k-tile
      const int s1 = (kt + 1) % STAGES;
      mbar_wait(&full[s1], ((kt + 1) / STAGES) & 1);
      load_frags(0, smem_base + (uint32_t)s1 * STAGE_BYTES, 0);
    }
    mma_chunk(1);
    // every LDS of stage s was issued a chunk ago and its result has been consumed by the DMMAs above
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ===== epilogue: each lane owns 4 consecutive columns (two 16-byte pairs) of 8 rows per 16-column block =====
  // For C -= P / C += P the old values are fetched in batches of 4 row-tiles (16 independent 16-byte loads in flight per
  // lane) BEFORE any store of the batch: a load-store-load chain would serialise 32 HBM round trips per tile.
  const int row0 = m0 + wm * 64 + x;
  const int col0 = n0 + wn * 32 + 4 * t;
  if (MODE == LA_GEMM_ASSIGN) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        store_pair(C, ldc, row0 + i * 8, col0 + j * 16, M, N, make_double2(acc[i][j][0], acc[i][j][1]));
        store_pair(C, ldc, row0 + i * 8, col0 + j * 16 + 2, M, N, make_double2(acc[i][j][2], acc[i][j][3]));
      }
  } else {
#pragma unroll
    for (int ib = 0; ib < 8; ib += 4) {
      double2 old[4][2][2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          old[i][j][0] = load_pair(C, ldc, row0 + (ib + i) * 8, col0 + j * 16, M, N);
          old[i][j][1] = load_pair(C, ldc, row0 + (ib + i) * 8, col0 + j * 16 + 2, M, N);
        }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double* a4 = acc[ib + i][j];
          double2 o0 = old[i][j][0], o1 = old[i][j][1];
          if (MODE == LA_GEMM_SUB) {
            o0 = make_double2(o0.x - a4[0], o0.y - a4[1]);
            o1 = make_double2(o1.x - a4[2], o1.y - a4[3]);
          } else {
            o0 = make_double2(o0.x + a4[0], o0.y + a4[1]);
            o1 = make_double2(o1.x + a4[2], o1.y + a4[3]);
          }
          store_pair(C, ldc, row0 + (ib + i) * 8, col0 + j * 16, M, N, o0);
          store_pair(C, ldc, row0 + (ib + i) * 8, col0 + j * 16 + 2, M, N, o1);
        }
    }
  }
}

int g_stagger = getenv("LA_GEMM_NO_STAGGER") ? 0 : 1;  // debug knob
int g_gemm_path = 0;  // test hook: 0 auto, 1 SIMT, 2 TMA/DMMA (auto tile), 3 TMA/DMMA BN=64, 4 TMA/DMMA BN=128

template <int MODE, int BN>
int launch_tma(const CUtensorMap& tmA, const CUtensorMap& tmB, double* C, size_t ldc, int M, int N, int K,
               cudaStream_t st, int slices = 1, size_t c_slab = 0, int lower_only = 0) {
  using Cfg = TileCfg<BN>;
  static const int extra_smem = getenv("LA_GEMM_EXTRA_SMEM") ? atoi(getenv("LA_GEMM_EXTRA_SMEM")) : 0;  // debug knob
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_f64_tma_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::SMEM_BYTES + extra_smem));
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_f64_tma_kernel<MODE, BN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  const int ktiles = (K + BK - 1) / BK;
  int stagger_lo = 0, stagger_hi = 0;
  unsigned stagger_ns = 0;
  if (Cfg::CTAS_PER_SM == 2 && ktiles <= 128 && g_stagger) {
    const DeviceCtx* ctx;
    LA_TRY(current_device_ctx(&ctx));
    stagger_lo = ctx->sm_count;
    stagger_hi = 2 * ctx->sm_count;
    stagger_ns = (unsigned)(ktiles * 520 + 2000);  // ~half of (main loop at full DMMA rate + prologue + epilogue)
  }
  const int kt_per_slice = (ktiles + slices - 1) / slices;
  const int nslices = (ktiles + kt_per_slice - 1) / kt_per_slice;  // every slice gets at least one k-tile
  gemm_f64_tma_kernel<MODE, BN><<<dim3(tiles_m * tiles_n, nslices), Cfg::THREADS, Cfg::SMEM_BYTES + extra_smem, st>>>(
      tmA, tmB, C, ldc, M, N, K, tiles_m, tiles_n, stagger_lo, stagger_hi, stagger_ns, kt_per_slice, c_slab, lower_only);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template <int BN>
int launch_tma_mode(int mode, const CUtensorMap& tmA, const CUtensorMap& tmB, double* C, size_t ldc, int M, int N,
                    int K, cudaStream_t st) {
  switch (mode) {
    case LA_GEMM_ASSIGN: return launch_tma<LA_GEMM_ASSIGN, BN>(tmA, tmB, C, ldc, M, N, K, st);
    case LA_GEMM_SUB: return launch_tma<LA_GEMM_SUB, BN>(tmA, tmB, C, ldc, M, N, K, st);
    default: return launch_tma<LA_GEMM_ADD, BN>(tmA, tmB, C, ldc, M, N, K, st);
  }
}

}  // namespace

void debug_set_gemm_path(int p) { g_gemm_path = p; }

int gemm_f64_dev(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                 size_t n, int mode, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(A && B && C, "la_gemm_f64: null matrix pointer");
  LA_REQUIRE(m > 0 && n > 0 && k > 0, "la_gemm_f64: zero dimension (m=%zu k=%zu n=%zu)", m, k, n);
  LA_REQUIRE(lda >= k && ldb >= n && ldc >= n, "la_gemm_f64: leading dimension smaller than row length");
  LA_REQUIRE(mode == LA_GEMM_ASSIGN || mode == LA_GEMM_SUB || mode == LA_GEMM_ADD, "la_gemm_f64: bad mode %d", mode);
  LA_REQUIRE(m < (1u << 30) && n < (1u << 30) && k < (1u << 30), "la_gemm_f64: dimension too large");

  const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
                       lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0;
  const bool tiles_ok = (m + BM - 1) / BM * ((n + 63) / 64) < (size_t)1 << 31;
  // Small products go to the CUDA-core kernel, which keeps the reference's exact per-element operation order (so the
  // reference's own `==` unit tests hold bit-for-bit); the tensor path takes over where throughput matters.
  const bool small = (double)m * (double)n * (double)k <= (double)SMALL_GEMM_MNK;
  bool use_tma = aligned && tiles_ok && !small;
  if (g_gemm_path >= 2) use_tma = aligned && tiles_ok;
  if (g_gemm_path == 1) use_tma = false;
  if (g_gemm_path >= 2 && !use_tma) return fail(LA_ERR_INVALID, "la_gemm_f64: TMA path forced but operands are unaligned");
  if (!use_tma) return gemm_simt<double>(A, lda, B, ldb, C, ldc, m, k, n, mode, st);

  CUtensorMap tmA, tmB;
  LA_TRY(encode_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, A, k, m, lda * 8, BK, BM,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, B, n, k, ldb * 8, 16, BK,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  // shallow K: the C tile traffic of the epilogue is a first-order cost -> two smaller CTAs per SM overlap it
  bool narrow = k <= SHALLOW_K;
  if (g_gemm_path == 3) narrow = true;
  if (g_gemm_path == 4) narrow = false;
  return narrow ? launch_tma_mode<64>(mode, tmA, tmB, C, ldc, (int)m, (int)n, (int)k, st)
                : launch_tma_mode<128>(mode, tmA, tmB, C, ldc, (int)m, (int)n, (int)k, st);
}

// Split-K product for short-and-wide outputs (m <= 128 rows, deep k): slice y of `slices` writes its partial product to
// Cslabs + y * slab_elems (ASSIGN); *slices_out receives the number of slabs actually written (every one holds at least
// one k-tile).  The caller sums the slabs (gemm_f64_sum_slabs) -- a fixed order, so results are reproducible.
int gemm_f64_splitk(const double* A, size_t lda, const double* B, size_t ldb, double* Cslabs, size_t ldc, size_t slab_elems,
                    size_t m, size_t k, size_t n, int slices, int* slices_out, cudaStream_t st) {
  const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)Cslabs % 16 == 0) &&
                       lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0 && slab_elems % 2 == 0;
  if (!aligned || m == 0 || n == 0 || k == 0 || slices < 1)
    return fail(LA_ERR_INVALID, "internal: split-K GEMM requested for operands that are not TMA-addressable");
  CUtensorMap tmA, tmB;
  LA_TRY(encode_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, A, k, m, lda * 8, BK, BM,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, B, n, k, ldb * 8, 16, BK,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  const int ktiles = (int)((k + BK - 1) / BK);
  const int per = (ktiles + slices - 1) / slices;
  *slices_out = (ktiles + per - 1) / per;
  return launch_tma<LA_GEMM_ASSIGN, 128>(tmA, tmB, Cslabs, ldc, (int)m, (int)n, (int)k, st, slices, slab_elems);
}

__global__ void __launch_bounds__(256) sum_slabs_kernel(const double* __restrict__ slabs, size_t slab_elems, int nslabs,
                                                        double* __restrict__ out, size_t count) {
  for (size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 2; i < count; i += (size_t)gridDim.x * 512) {
    double2 acc = *reinterpret_cast<const double2*>(slabs + i);
    for (int s = 1; s < nslabs; ++s) {
      const double2 v = *reinterpret_cast<const double2*>(slabs + (size_t)s * slab_elems + i);
      acc.x += v.x;
      acc.y += v.y;
    }
    *reinterpret_cast<double2*>(out + i) = acc;
  }
}
// out[0..count) = sum of the slabs (count even, 16-byte aligned); out may be slab 0 itself
int gemm_f64_sum_slabs(const double* slabs, size_t slab_elems, int nslabs, double* out, size_t count, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  size_t blocks = (count / 2 + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 8;
  sum_slabs_kernel<<<(unsigned)(blocks < cap ? (blocks ? blocks : 1) : cap), 256, 0, st>>>(slabs, slab_elems, nslabs, out, count);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}

int gemm_f64_preload() {
  cudaFuncAttributes fa;
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f64_tma_kernel<LA_GEMM_ASSIGN, 64>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f64_tma_kernel<LA_GEMM_SUB, 64>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f64_tma_kernel<LA_GEMM_ADD, 64>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f64_tma_kernel<LA_GEMM_ASSIGN, 128>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f64_tma_kernel<LA_GEMM_SUB, 128>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f64_tma_kernel<LA_GEMM_ADD, 128>));
  return LA_OK;
}

// C -= A * B restricted to the tiles of C that touch or lie below its diagonal (C square, m == n): the symmetric trailing
// update of the Cholesky factorisation in ONE launch; whole CTAs above the diagonal exit at once.
int gemm_f64_sub_lower(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                       cudaStream_t st) {
  const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
                       lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0;
  if (!aligned || m == 0 || k == 0)
    return fail(LA_ERR_INVALID, "internal: lower-triangle GEMM requested for operands that are not TMA-addressable");
  CUtensorMap tmA, tmB;
  LA_TRY(encode_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, A, k, m, lda * 8, BK, BM,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, B, m, k, ldb * 8, 16, BK,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  if (k <= SHALLOW_K) return launch_tma<LA_GEMM_SUB, 64>(tmA, tmB, C, ldc, (int)m, (int)m, (int)k, st, 1, 0, 1);
  return launch_tma<LA_GEMM_SUB, 128>(tmA, tmB, C, ldc, (int)m, (int)m, (int)k, st, 1, 0, 1);
}

// Tensor kernel regardless of problem size (the LU driver needs its in-place-safe tile structure: with m <= 128 there
// is one tile row and every CTA consumes its whole column block of B before the epilogue writes C == B).
int gemm_f64_tensor(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, cudaStream_t st) {
  const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
                       lda % 2 == 0 && ldb % 2 == 0 && ldc % 2 == 0;
  if (!aligned || m == 0 || n == 0 || k == 0)
    return fail(LA_ERR_INVALID, "internal: tensor GEMM requested for operands that are not TMA-addressable");
  CUtensorMap tmA, tmB;
  LA_TRY(encode_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, A, k, m, lda * 8, BK, BM,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, B, n, k, ldb * 8, 16, BK,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  const bool narrow = (k <= SHALLOW_K && g_gemm_path != 4) || g_gemm_path == 3;
  return narrow ? launch_tma_mode<64>(mode, tmA, tmB, C, ldc, (int)m, (int)n, (int)k, st)
                : launch_tma_mode<128>(mode, tmA, tmB, C, ldc, (int)m, (int)n, (int)k, st);
}

template <>
int gemm_dev<double>(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m,
                     size_t k, size_t n, int mode, cudaStream_t st) {
  return gemm_f64_dev(A, lda, B, ldb, C, ldc, m, k, n, mode, st);
}

}  // namespace la
