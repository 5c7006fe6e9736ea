// gemm_simt.cu -- generic CUDA-core GEMM (fp64 / fp32 / i64) for operands the TMA kernels cannot take (odd leading
// dimensions, unaligned base pointers) and for small shapes, where it reproduces the reference bit-for-bit: every
// element is accumulated for k ascending from zero with separately rounded multiply and add.  Same contract as the tensor kernels (row-major, explicit
// leading dimensions, ASSIGN / SUB / ADD epilogue).  Replaces the same reference loops: src/matrix/mod.rs:965-973.
// This is still a GPU path -- the library has no CPU fallback.
#include "la_common.cuh"

namespace la {
namespace {

constexpr int TS = 64;   // CTA tile 64 x 64
constexpr int TK = 16;   // k-step
constexpr int TPB = 256; // 16 x 16 threads, 4 x 4 outputs each

// res = res + a * b with the product and the sum rounded separately, k ascending from a zero accumulator: exactly the
// reference's expression (src/matrix/mod.rs:969), so this kernel is BIT-IDENTICAL to the reference's Mul.
__device__ __forceinline__ double mul_add(double a, double b, double c) { return add_rn(c, mul_rn(a, b)); }
__device__ __forceinline__ float mul_add(float a, float b, float c) { return add_rn(c, mul_rn(a, b)); }
// integer instance (the reference's Mul is generic and is unit-tested on integer matrices, src/matrix/mod.rs:1479-1484);
// two's-complement wrapping like a Rust release build
__device__ __forceinline__ long long mul_add(long long a, long long b, long long c) {
  return (long long)((unsigned long long)a * (unsigned long long)b + (unsigned long long)c);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(TPB) gemm_simt_kernel(const T* __restrict__ A, size_t lda, const T* __restrict__ B,
                                                        size_t ldb, T* __restrict__ C, size_t ldc, int M, int N,
                                                        int K) {
  __shared__ T sA[TK][TS + 4];  // transposed: sA[k][m]
  __shared__ T sB[TK][TS + 4];  // sB[k][n]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = (T)0;

  for (int k0 = 0; k0 < K; k0 += TK) {
    // A tile: 64 rows x 16 k; thread loads 4 elements (consecutive k are contiguous in memory)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = threadIdx.x + r * TPB;  // 0..1023
      const int mm = idx >> 4, kk = idx & 15;
      const int gm = m0 + mm, gk = k0 + kk;
      sA[kk][mm] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : (T)0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = threadIdx.x + r * TPB;
      const int kk = idx >> 6, nn = idx & 63;
      const int gk = k0 + kk, gn = n0 + nn;
      sB[kk][nn] = (gk < K && gn < N) ? B[(size_t)gk * ldb + gn] : (T)0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = mul_add(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      T* p = C + (size_t)gm * ldc + gn;
      if (MODE == LA_GEMM_ASSIGN)
        *p = acc[i][j];
      else if (MODE == LA_GEMM_SUB)
        *p = *p - acc[i][j];
      else
        *p = *p + acc[i][j];
    }
  }
}

}  // namespace

template <typename T>
int gemm_simt(const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc, size_t m, size_t k, size_t n, int mode,
              cudaStream_t st) {
  dim3 grid((unsigned)((n + TS - 1) / TS), (unsigned)((m + TS - 1) / TS));
  LA_REQUIRE(grid.y <= 65535, "la_gemm (generic kernel): more than 65535 row tiles");
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_simt_kernel<T, LA_GEMM_ASSIGN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_simt_kernel<T, LA_GEMM_SUB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  LA_CUDA_TRY(cudaFuncSetAttribute(gemm_simt_kernel<T, LA_GEMM_ADD>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   cudaSharedmemCarveoutMaxShared));
  switch (mode) {
    case LA_GEMM_ASSIGN:
      gemm_simt_kernel<T, LA_GEMM_ASSIGN><<<grid, TPB, 0, st>>>(A, lda, B, ldb, C, ldc, (int)m, (int)n, (int)k);
      break;
    case LA_GEMM_SUB:
      gemm_simt_kernel<T, LA_GEMM_SUB><<<grid, TPB, 0, st>>>(A, lda, B, ldb, C, ldc, (int)m, (int)n, (int)k);
      break;
    default:
      gemm_simt_kernel<T, LA_GEMM_ADD><<<grid, TPB, 0, st>>>(A, lda, B, ldb, C, ldc, (int)m, (int)n, (int)k);
      break;
  }
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
// Forces the (lazily loaded) kernels of this file into the context: see gemm_preload in la_common.cuh.
namespace {
template <typename T>
int preload_simt() {
  cudaFuncAttributes fa;
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_simt_kernel<T, LA_GEMM_ASSIGN>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_simt_kernel<T, LA_GEMM_SUB>));
  LA_CUDA_TRY(cudaFuncGetAttributes(&fa, gemm_simt_kernel<T, LA_GEMM_ADD>));
  return LA_OK;
}
}  // namespace
int gemm_simt_preload() {
  LA_TRY(preload_simt<double>());
  LA_TRY(preload_simt<float>());
  return LA_OK;
}
template int gemm_simt<double>(const double*, size_t, const double*, size_t, double*, size_t, size_t, size_t, size_t,
                               int, cudaStream_t);
template int gemm_simt<float>(const float*, size_t, const float*, size_t, float*, size_t, size_t, size_t, size_t, int,
                              cudaStream_t);
template int gemm_simt<long long>(const long long*, size_t, const long long*, size_t, long long*, size_t, size_t, size_t,
                                  size_t, int, cudaStream_t);
template <>
int gemm_dev<long long>(const long long* A, size_t lda, const long long* B, size_t ldb, long long* C, size_t ldc,
                        size_t m, size_t k, size_t n, int mode, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  return gemm_simt<long long>(A, lda, B, ldb, C, ldc, m, k, n, mode, st);
}

}  // namespace la
