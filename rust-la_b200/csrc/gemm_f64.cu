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

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 8;  // 16 KiB: 128 rows x 128 B
constexpr int B_STAGE_BYTES = BK * BN * 8;  // 16 KiB: 8 boxes of [16 k-rows x 128 B]
constexpr int B_BOX_BYTES = BK * 16 * 8;    // 2 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int CONSUMER_WARPS = 8;
// 8 consumer warps (2 warpgroups) + 1 producer warpgroup (only its first lane works).  The register file is split
// per SM sub-partition (16K regs each), so a 9th warp would cap every thread at 168 registers and spill the 128
// accumulator registers; instead 12 warps launch at 168 and setmaxnreg moves registers from the producer warpgroup
// (40) to the consumers (232): 2*232*32 + 40*32 = 16128 <= 16384 per sub-partition.
constexpr int GEMM_THREADS = (CONSUMER_WARPS + 4) * 32;
constexpr int CONSUMER_REGS = 232;
constexpr int PRODUCER_REGS = 40;
constexpr int GEMM_SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;  // + barriers + alignment slack
constexpr double SMALL_GEMM_MNK = 128.0 * 128.0 * 128.0;  // <= this many multiply-adds: bit-exact CUDA-core kernel
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

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_f64_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    double* __restrict__ C, size_t ldc, int M, int N, int K, int tiles_m, int tiles_n) {
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
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
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

  // ===== consumers: 8 warps, 2 (M) x 4 (N), warp tile 64 x 32 =====
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
  const int wm = warp >> 2;
  const int wn = warp & 3;
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

int g_gemm_path = 0;  // 0 auto, 1 force SIMT, 2 force TMA/DMMA (test hook)

template <int MODE>
int launch_tma(const CUtensorMap& tmA, const CUtensorMap& tmB, double* C, size_t ldc, int M, int N, int K,
               cudaStream_t st) {
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_f64_tma_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   GEMM_SMEM_BYTES));
  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
  gemm_f64_tma_kernel<MODE><<<tiles_m * tiles_n, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(tmA, tmB, C, ldc, M, N, K, tiles_m,
                                                                                    tiles_n);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
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
  const bool tiles_ok = (m + BM - 1) / BM * ((n + BN - 1) / BN) < (size_t)1 << 31;
  // Small products go to the CUDA-core kernel, which keeps the reference's exact per-element operation order (so the
  // reference's own `==` unit tests hold bit-for-bit); the tensor path takes over where throughput matters.
  const bool small = (double)m * (double)n * (double)k <= (double)SMALL_GEMM_MNK;
  bool use_tma = aligned && tiles_ok && !small;
  if (g_gemm_path == 2) use_tma = aligned && tiles_ok;
  if (g_gemm_path == 1) use_tma = false;
  if (g_gemm_path == 2 && !use_tma) return fail(LA_ERR_INVALID, "la_gemm_f64: TMA path forced but operands are unaligned");
  if (!use_tma) return gemm_simt<double>(A, lda, B, ldb, C, ldc, m, k, n, mode, st);

  CUtensorMap tmA, tmB;
  LA_TRY(encode_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, A, k, m, lda * 8, BK, BM,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, B, n, k, ldb * 8, 16, BK,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  switch (mode) {
    case LA_GEMM_ASSIGN: return launch_tma<LA_GEMM_ASSIGN>(tmA, tmB, C, ldc, (int)m, (int)n, (int)k, st);
    case LA_GEMM_SUB: return launch_tma<LA_GEMM_SUB>(tmA, tmB, C, ldc, (int)m, (int)n, (int)k, st);
    default: return launch_tma<LA_GEMM_ADD>(tmA, tmB, C, ldc, (int)m, (int)n, (int)k, st);
  }
}

template <>
int gemm_dev<double>(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m,
                     size_t k, size_t n, int mode, cudaStream_t st) {
  return gemm_f64_dev(A, lda, B, ldb, C, ldc, m, k, n, mode, st);
}

}  // namespace la
