// gemm_f32.cu -- fp32 GEMM entry point.  Replaces the reference loops src/matrix/mod.rs:965-973 with T = f32
// (and src/matrix/simd.rs:189-219, which has the same per-element order).
//
// Round-1 state: fp32 runs on the CUDA-core kernel (gemm_simt.cu, FFMA, full fp32 accuracy).  The tcgen05
// kind::tf32 kernel (TMEM accumulators, TMA-fed) is the next step; the fp32 LU trailing update must stay on an
// fp32-accurate path (FFMA or 3xTF32) to keep the backward error within 10x of the reference's.
#include "la_common.cuh"

namespace la {

int gemm_f32_dev(const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, size_t m, size_t k,
                 size_t n, int mode, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(A && B && C, "la_gemm_f32: null matrix pointer");
  LA_REQUIRE(m > 0 && n > 0 && k > 0, "la_gemm_f32: zero dimension (m=%zu k=%zu n=%zu)", m, k, n);
  LA_REQUIRE(lda >= k && ldb >= n && ldc >= n, "la_gemm_f32: leading dimension smaller than row length");
  LA_REQUIRE(mode == LA_GEMM_ASSIGN || mode == LA_GEMM_SUB || mode == LA_GEMM_ADD, "la_gemm_f32: bad mode %d", mode);
  LA_REQUIRE(m < (1u << 30) && n < (1u << 30) && k < (1u << 30), "la_gemm_f32: dimension too large");
  return gemm_simt<float>(A, lda, B, ldb, C, ldc, m, k, n, mode, st);
}

template <>
int gemm_dev<float>(const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, size_t m, size_t k,
                    size_t n, int mode, cudaStream_t st) {
  return gemm_f32_dev(A, lda, B, ldb, C, ldc, m, k, n, mode, st);
}

}  // namespace la
