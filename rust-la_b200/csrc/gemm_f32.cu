// gemm_f32.cu -- fp32 GEMM: C[m x n] (=, +=) A[m x k] * B[k x n], row-major, for sm_100a.
//
// Replaces the reference loops src/matrix/mod.rs:965-973 with T = f32 (and src/matrix/simd.rs:189-219, which has the same
// per-element order).  Two paths:
//   * tcgen05 TF32 kernel (this file): 5th-generation tensor cores, `tcgen05.mma.cta_group::1.kind::tf32` (SASS UTCHMMA)
//     issued by ONE thread, accumulators in TENSOR MEMORY (128 lanes x 256 fp32 columns), operands fed by TMA into a
//     4-stage shared-memory ring, epilogue `tcgen05.ld` -> registers -> 128-byte row segments.  TF32 keeps 10 mantissa bits
//     of each operand (relative error <= 2^-10 per factor), inside the 1e-4*k parity bar for k >= 32; shallower or
//     unaligned problems and the fp32 LU trailing update (which needs fp32-grade accuracy: LA_GEMM_SUB) stay on
//   * the CUDA-core kernel (gemm_simt.cu): exact fp32, reference summation order.
//
// Tile: 128 (M) x 256 (N) x 32 (K, = 128 bytes of fp32 = one swizzle row).  Per stage: A box 32 k x 128 rows and B^T box
// 32 k x 256 rows, both K-major with SWIZZLE_128B.  Row-major B has N contiguous, i.e. it is an "MN-major" operand; the
// instruction descriptor has a bit for that, but kind::tf32 with an MN-major B returned all-zero accumulators on this
// toolchain/driver for every LBO/SBO choice (measured, round 1), so B is transposed once per call into a scratch buffer
// (B^T, K-major) by a bandwidth-bound tile kernel (<= 6 % of the GEMM time at 8192^3, 1 % at the 65536x1024x16384 config).
// One MMA = M128 N256 K8; four per stage.  Warp roles (192 threads): warp 0 lane 0 = TMA producer, warp 1 = TMEM
// allocator + MMA issuer (one elected lane), warps 2-5 = epilogue (TMEM lane quarter = warp % 4).  The kernel is
// persistent (one CTA per SM) with two accumulators in TMEM, and C leaves through swizzled shared memory and TMA
// stores (full 32-byte sectors; the first version's per-lane 16-byte stores made L2 read-fill every sector of C).
#include <stdlib.h>
#include <string.h>

#include "la_common.cuh"

namespace la {
namespace {

constexpr int TBM = 128, TBN = 256, TBK = 32;
constexpr int TSTAGES = 4;
constexpr int TA_BYTES = TBM * TBK * 4;        // 16 KiB
constexpr int TB_BYTES = TBK * TBN * 4;        // 32 KiB
constexpr int TSTAGE_BYTES = TA_BYTES + TB_BYTES;
constexpr int TF32_THREADS = 192;
constexpr int TOUT_BOX_BYTES = 32 * 32 * 4;              // one 32 x 32 fp32 store box
constexpr int TOUT_BYTES = 4 * 2 * TOUT_BOX_BYTES;      // 4 epilogue warps, double-buffered
constexpr int TF32_SMEM = TSTAGES * TSTAGE_BYTES + TOUT_BYTES + 256 + 1024;  // + barriers/tmem slot + alignment slack
constexpr int TMEM_COLS = 512;                           // two 128 x 256 fp32 accumulators
constexpr size_t TF32_MIN_K = 32;  // below this TF32's 2^-10 input rounding is not covered by the 1e-4*k parity bar

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading/stride byte offsets (all
// >> 4), version 1 (Blackwell), layout type SWIZZLE_128B = 2 in bits [61,64).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version_
  d |= (uint64_t)2 << 61;  // layout_type_ = SWIZZLE_128B
  return d;
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (2u << 7) | (2u << 10)        // a_format = b_format = TF32
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {  // arrives on `bar` when all prior MMAs of this thread finish
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// TMA store / reduce-add of a shared-memory box into a global tensor (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* smem_src, int c_inner, int c_outer) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Tile order: bands of group_m tile rows, m fastest inside a band.  The CTAs of one wave then share a handful of B^T
// tiles and a band of A that both stay in L2 while the band sweeps all tile columns.
__device__ __forceinline__ void tile_coords(int t, int tiles_m, int tiles_n, int group_m, int& tile_m, int& tile_n) {
  const int per_group = group_m * tiles_n;
  const int first_m = (t / per_group) * group_m;
  const int rows_here = min(group_m, tiles_m - first_m);
  tile_m = first_m + (t % per_group) % rows_here;
  tile_n = (t % per_group) / rows_here;
}

// PERSISTENT kernel: one CTA per SM walks the tile list.  The 512 TMEM columns hold TWO 128 x 256 fp32 accumulators, so the
// epilogue of tile i (TMEM -> registers -> swizzled shared memory -> TMA store / reduce-add) runs under the MMAs of tile
// i+1; the shared-memory ring keeps streaming across tile boundaries.
template <int MODE>
__global__ void __launch_bounds__(TF32_THREADS, 1)
gemm_f32_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                     const __grid_constant__ CUtensorMap tmC, int M, int N, int K, int tiles_m, int tiles_n, int group_m,
                     int kpasses) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* out_stage = smem + TSTAGES * TSTAGE_BYTES;                     // [4 warps][2][32 rows x 128 B], 1024-aligned
  uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + TOUT_BYTES);
  uint64_t* empty = full + TSTAGES;
  uint64_t* acc_full = empty + TSTAGES;   // [2]: MMAs of a tile done
  uint64_t* acc_empty = acc_full + 2;     // [2]: epilogue has drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ktiles1 = (K + TBK - 1) / TBK;
  // kpasses == 3 (split-compensated "3xTF32"): the K loop runs three times over operand pairs
  //   (A_small, B_big), (A_big, B_small), (A_big, B_big)   with tmA = A_big, tmA1 = A_small, tmB = B_big^T, tmB1 = B_small^T
  // into the same fp32 accumulator -- the small terms first, so they are summed at full relative precision.
  const int ktiles = ktiles1 * kpasses;
  const int ntiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // TMEM allocation is warp-wide; the base address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (kpasses > 1) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB1);
      }
      uint32_t g = 0;  // k-tiles issued so far (ring position)
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int tile_m, tile_n;
        tile_coords(t, tiles_m, tiles_n, group_m, tile_m, tile_n);
        for (int kt = 0; kt < ktiles; ++kt, ++g) {
          const int s = g % TSTAGES;
          mbar_wait(&empty[s], ((g / TSTAGES) & 1) ^ 1);
          uint8_t* sA = smem + s * TSTAGE_BYTES;
          uint8_t* sB = sA + TA_BYTES;
          mbar_arrive_expect_tx(&full[s], TSTAGE_BYTES);
          const int pass = (kpasses > 1) ? kt / ktiles1 : 2;  // 0: small x big, 1: big x small, 2: big x big
          const int kc = (kt - (kpasses > 1 ? pass * ktiles1 : 0)) * TBK;
          tma_load_2d(sA, pass == 0 ? &tmA1 : &tmA, &full[s], kc, tile_m * TBM);  // 32 k (inner, 128 B) x 128 rows
          tma_load_2d(sB, pass == 1 ? &tmB1 : &tmB, &full[s], kc, tile_n * TBN);  // B^T: 32 k (inner, 128 B) x 256 n-rows
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: a single thread drives the tensor core =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TBM, TBN, /*A K-major*/ 0, /*B^T K-major*/ 0);
      uint32_t g = 0;
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int b = it & 1;
        mbar_wait(&acc_empty[b], ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator (free at first use)
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * TBN);
        for (int kt = 0; kt < ktiles; ++kt, ++g) {
          const int s = g % TSTAGES;
          mbar_wait(&full[s], (g / TSTAGES) & 1);
          tcgen05_fence_after();
          const uint32_t sA = smem_u32(smem + s * TSTAGE_BYTES);
          const uint32_t sB = sA + TA_BYTES;
#pragma unroll
          for (int kk = 0; kk < TBK / 8; ++kk) {
            // K-major operands: 8 k = 32 bytes further along the 128-byte swizzled row; 8-row groups are 1024 B apart (SBO)
            const uint64_t adesc = umma_smem_desc(sA + kk * 32, 16, 1024);
            const uint64_t bdesc = umma_smem_desc(sB + kk * 32, 16, 1024);
            umma_tf32(tmem_d, adesc, bdesc, idesc, (kt | kk) ? 1u : 0u);
          }
          umma_commit(&empty[s]);  // frees the stage once these MMAs have read it
        }
        umma_commit(&acc_full[b]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes [32*(w%4), +32) == 32 tile rows =====
    const int quarter = warp & 3;
    uint8_t* my_stage = out_stage + (warp - 2) * (2 * TOUT_BOX_BYTES);
    int it = 0;
    uint32_t nbox = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      int tile_m, tile_n;
      tile_coords(t, tiles_m, tiles_n, group_m, tile_m, tile_n);
      const int b = it & 1;
      const int row0 = tile_m * TBM + quarter * 32;
      mbar_wait(&acc_full[b], (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TBN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * TBN + c0), r);
        const int col0 = tile_n * TBN + c0;
        if (row0 < M && col0 < N) {  // warp-uniform
          uint8_t* box = my_stage + (nbox & 1) * TOUT_BOX_BYTES;
          ++nbox;
          if (lane == 0) bulk_wait_read<1>();  // the store issued from this buffer two boxes ago has read it
          __syncwarp();
          // row `lane` of the box, 128 bytes, 16-byte chunks XOR-swizzled by the row (SWIZZLE_128B): conflict-free
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<uint4*>(box + lane * 128 + ((v ^ (lane & 7)) << 4)) =
                make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (MODE == LA_GEMM_ADD) tma_reduce_add_2d(&tmC, box, col0, row0);
            else tma_store_2d(&tmC, box, col0, row0);  // rows / columns outside C are clipped by the TMA unit
            bulk_commit();
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    if (lane == 0) bulk_wait_all();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-PAIR kernel (`tcgen05.mma.cta_group::2`): two CTAs of one cluster (the two SMs of a TPC) work on one 256 x 256 tile.
// Each CTA stages ITS 128 rows of A and ITS 128 of the tile's 256 B^T rows (32 KiB per stage instead of 48, six stages),
// the leader's elected thread issues M256 N256 K8 instructions that read A from both CTAs' shared memory and each half
// of B from its CTA, and every CTA ends up with its 128 x 256 block of the accumulator in its own TMEM.  The single-CTA
// kernel above is bound by the shared-memory pipe (TMA fills + operand reads: 73 % busy at 69 % tensor activity,
// profiles/r1_gemm_f32_tf32_persistent_ncu_summary.txt); the pair halves the B operand traffic per SM.
// Barriers: `full[s]` lives in the LEADER (both CTAs' TMA loads complete their bytes on it: `.cta_group::2` loads with the
// leader's barrier address), `empty[s]` and `acc_full[b]` exist in both CTAs and are signalled by one multicast
// `tcgen05.commit`, `acc_empty[b]` lives in the leader and counts the eight epilogue warps of both CTAs.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PSTAGES = 6;
constexpr int PBN_HALF = TBN / 2;                  // B^T rows staged per CTA
constexpr int PB_BYTES = TBK * PBN_HALF * 4;       // 16 KiB
constexpr int PSTAGE_BYTES = TA_BYTES + PB_BYTES;  // 32 KiB
constexpr int PAIR_SMEM = PSTAGES * PSTAGE_BYTES + TOUT_BYTES + 256 + 1024;
constexpr int PTM = 2 * TBM;                       // tile rows of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr,
                                                 int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` (same offset) in BOTH CTAs of the pair once all prior MMAs of this thread have finished
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TF32_THREADS, 1)
gemm_f32_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                          const __grid_constant__ CUtensorMap tmC, int M, int N, int K, int tiles_m, int tiles_n, int group_m,
                          int kpasses) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* out_stage = smem + PSTAGES * PSTAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_stage + TOUT_BYTES);
  uint64_t* empty = full + PSTAGES;
  uint64_t* acc_full = empty + PSTAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int ktiles1 = (K + TBK - 1) / TBK;
  const int ktiles = ktiles1 * kpasses;
  const int ntiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(&full[s], 1);   // the leader producer's arrive.expect_tx; bytes from both CTAs' loads
      mbar_init(&empty[s], 1);  // one multicast commit per use
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);  // four epilogue warps in each CTA (only the leader's copy is used)
    }
    mbar_fence_init();
  }
  cluster_sync_all();  // barriers of both CTAs exist before anybody signals across
  if (warp == 1) {     // the same warp of both CTAs: a pair-wide TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of A, own 128 rows of B^T; bytes land on the leader's barrier =====
    if (lane == 0) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      if (kpasses > 1) {
        tma_prefetch_desc(&tmA1);
        tma_prefetch_desc(&tmB1);
      }
      uint32_t g = 0;
      for (int t = pair; t < ntiles; t += npairs) {
        int tile_m, tile_n;
        tile_coords(t, tiles_m, tiles_n, group_m, tile_m, tile_n);
        for (int kt = 0; kt < ktiles; ++kt, ++g) {
          const int s = g % PSTAGES;
          mbar_wait(&empty[s], ((g / PSTAGES) & 1) ^ 1);
          uint8_t* sA = smem + s * PSTAGE_BYTES;
          uint8_t* sB = sA + TA_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * PSTAGE_BYTES);
          const uint32_t bar = mapa_shared(smem_u32(&full[s]), 0);
          const int pass = (kpasses > 1) ? kt / ktiles1 : 2;
          const int kc = (kt - (kpasses > 1 ? pass * ktiles1 : 0)) * TBK;
          tma_load_2d_pair(sA, pass == 0 ? &tmA1 : &tmA, bar, kc, tile_m * PTM + (int)rank * TBM);
          tma_load_2d_pair(sB, pass == 1 ? &tmB1 : &tmB, bar, kc, tile_n * TBN + (int)rank * PBN_HALF);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the LEADER drives both tensor cores =====
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(PTM, TBN, 0, 0);
      uint32_t g = 0;
      int it = 0;
      for (int t = pair; t < ntiles; t += npairs, ++it) {
        const int b = it & 1;
        mbar_wait(&acc_empty[b], ((it >> 1) & 1) ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(b * TBN);
        for (int kt = 0; kt < ktiles; ++kt, ++g) {
          const int s = g % PSTAGES;
          mbar_wait(&full[s], (g / PSTAGES) & 1);
          tcgen05_fence_after();
          const uint32_t sA = smem_u32(smem + s * PSTAGE_BYTES);
          const uint32_t sB = sA + TA_BYTES;
#pragma unroll
          for (int kk = 0; kk < TBK / 8; ++kk) {
            const uint64_t adesc = umma_smem_desc(sA + kk * 32, 16, 1024);
            const uint64_t bdesc = umma_smem_desc(sB + kk * 32, 16, 1024);
            umma_tf32_pair(tmem_d, adesc, bdesc, idesc, (kt | kk) ? 1u : 0u);
          }
          umma_commit_pair(&empty[s]);
        }
        umma_commit_pair(&acc_full[b]);
      }
    }
  } else {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the tile, TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    uint8_t* my_stage = out_stage + (warp - 2) * (2 * TOUT_BOX_BYTES);
    int it = 0;
    uint32_t nbox = 0;
    for (int t = pair; t < ntiles; t += npairs, ++it) {
      int tile_m, tile_n;
      tile_coords(t, tiles_m, tiles_n, group_m, tile_m, tile_n);
      const int b = it & 1;
      const int row0 = tile_m * PTM + (int)rank * TBM + quarter * 32;
      mbar_wait(&acc_full[b], (it >> 1) & 1);
      tcgen05_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < TBN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(b * TBN + c0), r);
        const int col0 = tile_n * TBN + c0;
        if (row0 < M && col0 < N) {  // warp-uniform
          uint8_t* box = my_stage + (nbox & 1) * TOUT_BOX_BYTES;
          ++nbox;
          if (lane == 0) bulk_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int v = 0; v < 8; ++v)
            *reinterpret_cast<uint4*>(box + lane * 128 + ((v ^ (lane & 7)) << 4)) =
                make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (MODE == LA_GEMM_ADD) tma_reduce_add_2d(&tmC, box, col0, row0);
            else tma_store_2d(&tmC, box, col0, row0);
            bulk_commit();
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[b]), 0));
    }
    if (lane == 0) bulk_wait_all();
  }
  // nobody leaves (or frees TMEM) while the partner may still read its shared memory or signal its barriers
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// Split of an fp32 value for the compensated ("3xTF32") product: big = x rounded to TF32 (10 explicit mantissa bits, so
// the tensor core's own input conversion is the identity on it), small = x - big (exact in fp32; the tensor core keeps
// its leading 11 bits).  x*y ~= big_x*big_y + big_x*small_y + small_x*big_y with relative error ~2^-21 per product.
__device__ __forceinline__ void split_tf32(float x, float& big, float& small) {
  const uint32_t u = __float_as_uint(x);
  if ((u & 0x7f800000u) == 0x7f800000u) {  // inf / nan travel in the big part
    big = x;
    small = 0.0f;
    return;
  }
  big = __uint_as_float((u + 0x1000u) & 0xffffe000u);  // round to nearest (ties away) at bit 13
  small = x - big;
}

// B^T[n][k] = B[k][n] through 32 x 33 shared-memory tiles: both the read and the write are coalesced 128-byte rows.
// SPLIT also writes the small parts (second destination).
template <bool SPLIT>
__global__ void __launch_bounds__(256) transpose_f32_kernel(const float* __restrict__ B, size_t ldb, float* __restrict__ BT,
                                                            float* __restrict__ BT_small, size_t ldt, int K, int N) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int kk = k0 + ty + r, nn = n0 + tx;
    tile[ty + r][tx] = (kk < K && nn < N) ? B[(size_t)kk * ldb + nn] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 32; r += 8) {
    const int nn = n0 + ty + r, kk = k0 + tx;
    if (nn < N && kk < K) {
      const float x = tile[tx][ty + r];
      if (SPLIT) {
        float b, s;
        split_tf32(x, b, s);
        BT[(size_t)nn * ldt + kk] = b;
        BT_small[(size_t)nn * ldt + kk] = s;
      } else {
        BT[(size_t)nn * ldt + kk] = x;
      }
    }
  }
}

// A -> (A_big, A_small), same row-major shape, leading dimension ldo
__global__ void __launch_bounds__(256) split_f32_kernel(const float* __restrict__ A, size_t lda, float* __restrict__ big,
                                                        float* __restrict__ small, size_t ldo, size_t M, size_t K) {
  const size_t k4 = (K + 3) / 4;  // groups of four columns
  const size_t total = M * k4;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / k4, c = (e - r * k4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c + 4 <= K) {
      v = *reinterpret_cast<const float4*>(A + r * lda + c);  // lda % 4 == 0 and A 16-byte aligned (tensor-path precondition)
    } else {
      v.x = A[r * lda + c];
      if (c + 1 < K) v.y = A[r * lda + c + 1];
      if (c + 2 < K) v.z = A[r * lda + c + 2];
    }
    float4 b, s;
    split_tf32(v.x, b.x, s.x);
    split_tf32(v.y, b.y, s.y);
    split_tf32(v.z, b.z, s.z);
    split_tf32(v.w, b.w, s.w);
    *reinterpret_cast<float4*>(big + r * ldo + c) = b;  // ldo is a multiple of 4: the padded tail is ours to write
    *reinterpret_cast<float4*>(small + r * ldo + c) = s;
  }
}

int g_f32_path = getenv("LA_GEMM_F32_PATH") ? atoi(getenv("LA_GEMM_F32_PATH")) : 0;  // 0 auto, 1 CUDA-core, 2 tcgen05
// Accuracy mode of the tensor path: LA_F32_3XTF32 (default, fp32-grade) or LA_F32_TF32 (opt-in, 10-bit inputs)
int initial_f32_mode() {
  const char* e = getenv("LA_GEMM_F32_MODE");
  if (e && (!strcmp(e, "tf32") || !strcmp(e, "1"))) return LA_F32_TF32;
  return LA_F32_3XTF32;
}
int g_f32_mode = initial_f32_mode();

template <int MODE>
int launch_tf32(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA1, const CUtensorMap& tmB1,
                const CUtensorMap& tmC, int M, int N, int K, int kpasses, int sms, cudaStream_t st) {
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_f32_tf32_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF32_SMEM));
  const int tiles_m = (M + TBM - 1) / TBM, tiles_n = (N + TBN - 1) / TBN;
  const int ntiles = tiles_m * tiles_n;
  static const int group_m = getenv("LA_TF32_GROUP_M") ? atoi(getenv("LA_TF32_GROUP_M")) : 16;  // tuning knob
  gemm_f32_tf32_kernel<MODE><<<ntiles < sms ? ntiles : sms, TF32_THREADS, TF32_SMEM, st>>>(
      tmA, tmB, tmA1, tmB1, tmC, M, N, K, tiles_m, tiles_n, group_m, kpasses);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}

template <int MODE>
int launch_tf32_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmA1, const CUtensorMap& tmB1,
                     const CUtensorMap& tmC, int M, int N, int K, int kpasses, int sms, cudaStream_t st) {
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_f32_tf32_pair_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
  const int tiles_m = (M + PTM - 1) / PTM, tiles_n = (N + TBN - 1) / TBN;
  const int ntiles = tiles_m * tiles_n;
  // bands of 16 pair-tile rows (4096 rows of A, 16 MiB at k = 1024, stay in L2 while the band sweeps B^T): measured
  // 722 / 752 / 770 / 765 / 679 TFLOP/s at 4 / 8 / 16 / 32 / 64 rows per band on the 65536 x 1024 x 16384 product
  static const int group_m = getenv("LA_TF32_PAIR_GROUP_M") ? atoi(getenv("LA_TF32_PAIR_GROUP_M")) : 16;
  const int pairs = ntiles < sms / 2 ? ntiles : sms / 2;
  gemm_f32_tf32_pair_kernel<MODE><<<2 * pairs, TF32_THREADS, PAIR_SMEM, st>>>(tmA, tmB, tmA1, tmB1, tmC, M, N, K, tiles_m,
                                                                             tiles_n, group_m < 1 ? 1 : group_m, kpasses);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}

}  // namespace

void debug_set_gemm_f32_path(int p) { g_f32_path = p; }
int set_gemm_f32_mode(int mode) {
  LA_REQUIRE(mode == LA_F32_3XTF32 || mode == LA_F32_TF32, "la_set_gemm_f32_mode: bad mode %d", mode);
  g_f32_mode = mode;
  return LA_OK;
}
int get_gemm_f32_mode() { return g_f32_mode; }

int gemm_f32_dev(const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, size_t m, size_t k,
                 size_t n, int mode, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(A && B && C, "la_gemm_f32: null matrix pointer");
  LA_REQUIRE(m > 0 && n > 0 && k > 0, "la_gemm_f32: zero dimension (m=%zu k=%zu n=%zu)", m, k, n);
  LA_REQUIRE(lda >= k && ldb >= n && ldc >= n, "la_gemm_f32: leading dimension smaller than row length");
  LA_REQUIRE(mode == LA_GEMM_ASSIGN || mode == LA_GEMM_SUB || mode == LA_GEMM_ADD, "la_gemm_f32: bad mode %d", mode);
  LA_REQUIRE(m < (1u << 30) && n < (1u << 30) && k < (1u << 30), "la_gemm_f32: dimension too large");

  const bool aligned = ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) &&
                       lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0;
  const bool small = (double)m * (double)n * (double)k <= 128.0 * 128.0 * 128.0;  // bit-exact reference-order kernel
  const int acc_mode = g_f32_mode;
  // plain TF32 rounds the inputs to 10 mantissa bits: only when the caller opted in, and only where the 1e-4*k parity bar
  // covers it (k >= 32).  The split-compensated mode is fp32-grade and has no such floor.
  bool use_tc = aligned && !small && mode != LA_GEMM_SUB && (acc_mode == LA_F32_3XTF32 || k >= TF32_MIN_K);
  if (g_f32_path == 1) use_tc = false;
  if (g_f32_path == 2) {
    if (!aligned || mode == LA_GEMM_SUB)
      return fail(LA_ERR_INVALID, "la_gemm_f32: tcgen05 path forced but operands are unaligned or mode is SUB");
    use_tc = true;
  }
  if (!use_tc) return gemm_simt<float>(A, lda, B, ldb, C, ldc, m, k, n, mode, st);

  // B^T into scratch (K-major operand for the tensor core), leading dimension padded to 16 bytes.  In the compensated
  // mode the same pass also writes the small parts, and A is split into (big, small) by a streaming kernel.
  const bool comp = acc_mode == LA_F32_3XTF32;
  const size_t ldt = (k + 3) & ~(size_t)3;
  void* bt = nullptr;
  LA_TRY(scratch_get(ctx->device, 12, (comp ? 2 : 1) * n * ldt * sizeof(float), &bt));
  float* bt_big = (float*)bt;
  float* bt_small = comp ? bt_big + n * ldt : nullptr;
  {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((k + 31) / 32));
    LA_REQUIRE(grid.y <= 65535, "la_gemm_f32: inner dimension too large for the transpose grid");
    if (comp) transpose_f32_kernel<true><<<grid, 256, 0, st>>>(B, ldb, bt_big, bt_small, ldt, (int)k, (int)n);
    else transpose_f32_kernel<false><<<grid, 256, 0, st>>>(B, ldb, bt_big, nullptr, ldt, (int)k, (int)n);
    LA_CUDA_TRY(cudaGetLastError());
  }
  const float* a_big = A;
  const float* a_small = A;
  size_t ld_a = lda;
  if (comp) {
    void* as = nullptr;
    LA_TRY(scratch_get(ctx->device, 20, 2 * m * ldt * sizeof(float), &as));
    float* ab = (float*)as;
    float* asml = ab + m * ldt;
    size_t blocks = (m * (ldt / 4) + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    split_f32_kernel<<<(unsigned)blocks, 256, 0, st>>>(A, lda, ab, asml, ldt, m, k);
    LA_CUDA_TRY(cudaGetLastError());
    a_big = ab;
    a_small = asml;
    ld_a = ldt;
  }
  // CTA pairs (cta_group::2, 256 x 256 tiles) when the product has more than one 128-row band
  static const int pair_knob = getenv("LA_TF32_PAIR") ? atoi(getenv("LA_TF32_PAIR")) : 1;  // 0: single-CTA kernel
  const bool pair = pair_knob != 0 && m > (size_t)TBM && ctx->sm_count >= 2;
  const int b_box = pair ? PBN_HALF : TBN;  // B^T rows per TMA box
  CUtensorMap tmA, tmB, tmA1, tmB1;
  LA_TRY(encode_tensor_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_big, k, m, ld_a * 4, TBK, TBM,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, bt_big, k, n, ldt * 4, TBK, b_box,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmA1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a_small, k, m, ld_a * 4, TBK, TBM,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  LA_TRY(encode_tensor_map_2d(&tmB1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, comp ? bt_small : bt_big, k, n, ldt * 4, TBK, b_box,
                              CU_TENSOR_MAP_SWIZZLE_128B));
  CUtensorMap tmC;  // store boxes: 32 columns (128 B) x 32 rows, same swizzle as the epilogue's shared-memory layout
  LA_TRY(encode_tensor_map_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, n, m, ldc * 4, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B));
  const int kpasses = comp ? 3 : 1;
  if (pair) {
    if (mode == LA_GEMM_ASSIGN)
      return launch_tf32_pair<LA_GEMM_ASSIGN>(tmA, tmB, tmA1, tmB1, tmC, (int)m, (int)n, (int)k, kpasses, ctx->sm_count, st);
    return launch_tf32_pair<LA_GEMM_ADD>(tmA, tmB, tmA1, tmB1, tmC, (int)m, (int)n, (int)k, kpasses, ctx->sm_count, st);
  }
  if (mode == LA_GEMM_ASSIGN)
    return launch_tf32<LA_GEMM_ASSIGN>(tmA, tmB, tmA1, tmB1, tmC, (int)m, (int)n, (int)k, kpasses, ctx->sm_count, st);
  return launch_tf32<LA_GEMM_ADD>(tmA, tmB, tmA1, tmB1, tmC, (int)m, (int)n, (int)k, kpasses, ctx->sm_count, st);
}

int gemm_f32_preload() {
  cudaFuncAttributes fa;
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f32_tf32_kernel<LA_GEMM_ASSIGN>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f32_tf32_kernel<LA_GEMM_ADD>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f32_tf32_pair_kernel<LA_GEMM_ASSIGN>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_f32_tf32_pair_kernel<LA_GEMM_ADD>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, transpose_f32_kernel<true>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, transpose_f32_kernel<false>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, split_f32_kernel));
  return LA_OK;
}

template <>
int gemm_dev<float>(const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, cudaStream_t st) {
  return gemm_f32_dev(A, lda, B, ldb, C, ldc, m, k, n, mode, st);
}

}  // namespace la
