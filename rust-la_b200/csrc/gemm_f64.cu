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
#include "la_common.cuh"

namespace la {
namespace {

constexpr int BM = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 8;  // 16 KiB: 128 rows x 128 B
constexpr int B_BOX_BYTES = BK * 16 * 8;    // 2 KiB: [16 k-rows x 128 B] holds 16 columns of B
// Two tile configurations of the same kernel:
//   BN = 128, 8 consumer warps, 1 CTA/SM : least L2->smem traffic per flop; used for deep K (Mul at large n).
//   BN =  64, 4 consumer warps, 2 CTA/SM : the epilogue of one CTA (C tile read-modify-write, HBM bound) overlaps the
//                                           main loop of the other; used for shallow K (the LU trailing update, K = nb).
// Each configuration adds one producer warpgroup (only its first lane works).  The register file is split per SM
// sub-partition (16K regs each); the CTA launches at an even split and setmaxnreg moves registers from the producer
// warpgroup to the consumers (128 accumulator registers + fragments per consumer thread):
//   BN=128: 12 warps launch at 168 -> consumers 232, producer 40   (2*232 + 40 = 504 = 3*168)
//   BN= 64:  8 warps launch at 128 -> consumers 216, producer 40   (216 + 40 = 256 = 2*128), twice per SM
template <int BN_>
struct TileCfg {
  static constexpr int CONSUMER_WARPS = BN_ / 16;                 // 2 (M) x BN/32 (N) warps of 64 x 32
  static constexpr int THREADS = (CONSUMER_WARPS + 4) * 32;
  static constexpr int CTAS_PER_SM = BN_ == 128 ? 1 : 2;
  static constexpr int CONSUMER_REGS = BN_ == 128 ? 232 : 216;
  static constexpr int PRODUCER_REGS = 40;
  static constexpr int B_STAGE_BYTES = BK * BN_ * 8;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;  // + barriers + alignment slack
};
constexpr double SMALL_GEMM_MNK = 128.0 * 128.0 * 128.0;  // <= this many multiply-adds: bit-exact CUDA-core kernel
constexpr size_t SHALLOW_K = 1024;
constexpr int GROUP_M = 16;  // tile rasterisation: GROUP_M tile-rows share each B tile-column while it is hot in L2

template <int MODE>
__device__ __forceinline__ void store_pair(double* __restrict__ C, size_t ldc, int row, int col, int N, double v0,
                                           double v1) {
  double* p = C + (size_t)row * ldc + col;
  if (col + 1 < N) {
    double2 out;
    if (MODE == LA_GEMM_ASSIGN) {
      out = make_double2(v0, v1);
    } else {
      double2 old = *reinterpret_cast<const double2*>(p);
      out = (MODE == LA_GEMM_SUB) ? make_double2(old.x - v0, old.y - v1) : make_double2(old.x + v0, old.y + v1);
    }
    *reinterpret_cast<double2*>(p) = out;
  } else if (col < N) {
    if (MODE == LA_GEMM_ASSIGN)
      p[0] = v0;
    else if (MODE == LA_GEMM_SUB)
      p[0] = p[0] - v0;
    else
      p[0] = p[0] + v0;
  }
}

template <int MODE, int BN>
__global__ void __launch_bounds__(TileCfg<BN>::THREADS, TileCfg<BN>::CTAS_PER_SM)
gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    double* __restrict__ C, size_t ldc, int M, int N, int K, int tiles_m, int tiles_n) {
  using Cfg = TileCfg<BN>;
  constexpr int CONSUMER_WARPS = Cfg::CONSUMER_WARPS;
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

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  const int ktiles = (K + BK - 1) / BK;
  const uint32_t smem_base = smem_u32(smem);

  if (warp >= CONSUMER_WARPS) {
    // ===== TMA producer warpgroup: one elected lane works, the rest only donate registers =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::PRODUCER_REGS));
    if (warp == CONSUMER_WARPS && lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int kt = 0; kt < ktiles; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sA = smem + s * STAGE_BYTES;
        uint8_t* sB = sA + A_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        tma_load_2d(sA, &tmA, &full[s], kt * BK, m0);  // box: 16 k (inner) x 128 rows; OOB -> 0
#pragma unroll
        for (int j = 0; j < BN / 16; ++j)              // box: 16 n (inner) x 16 k-rows
          tma_load_2d(sB + j * B_BOX_BYTES, &tmB, &full[s], n0 + j * 16, kt * BK);
      }
    }
    return;
  }

  // ===== consumers: 2 (M) x BN/32 (N) warps, warp tile 64 x 32 =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::CONSUMER_REGS));
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

  for (int kt = 0; kt < ktiles; ++kt) {
    const int s = kt % STAGES;
    const uint32_t ph = (kt / STAGES) & 1;
    mbar_wait(&full[s], ph);
    const uint32_t st = smem_base + (uint32_t)s * STAGE_BYTES;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double2 af[8];
      double2 bf[2][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = lds_f64x2(st + a_row + i * 1024 + a_chunk[h]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int ss = 0; ss < 2; ++ss)
          bf[j][ss] = lds_f64x2(st + b_off[ss][h] + j * B_BOX_BYTES);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma884(acc[i][j][0], acc[i][j][2], af[i].x, bf[j][0].x);  // even k, even columns
          dmma884(acc[i][j][1], acc[i][j][3], af[i].x, bf[j][0].y);  // even k, odd columns
          dmma884(acc[i][j][0], acc[i][j][2], af[i].y, bf[j][1].x);  // odd k, even columns
          dmma884(acc[i][j][1], acc[i][j][3], af[i].y, bf[j][1].y);  // odd k, odd columns
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ===== epilogue: each lane owns 4 consecutive columns of 8 rows per n-block =====
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + wm * 64 + i * 8 + x;
    if (row < M) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = n0 + wn * 32 + j * 16 + 4 * t;
        store_pair<MODE>(C, ldc, row, col, N, acc[i][j][0], acc[i][j][1]);
        store_pair<MODE>(C, ldc, row, col + 2, N, acc[i][j][2], acc[i][j][3]);
      }
    }
  }
}

int g_gemm_path = 0;  // test hook: 0 auto, 1 SIMT, 2 TMA/DMMA (auto tile), 3 TMA/DMMA BN=64, 4 TMA/DMMA BN=128

template <int MODE, int BN>
int launch_tma(const CUtensorMap& tmA, const CUtensorMap& tmB, double* C, size_t ldc, int M, int N, int K,
               cudaStream_t st) {
  using Cfg = TileCfg<BN>;
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_f64_tma_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::SMEM_BYTES));
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  gemm_f64_tma_kernel<MODE, BN><<<tiles_m * tiles_n, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, C, ldc, M, N, K,
                                                                                        tiles_m, tiles_n);
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
