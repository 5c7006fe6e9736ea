// la_common.cuh -- shared host/device helpers for the sm_100a hot-path library.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/la_cabi.h"

namespace la {

// ---- error plumbing (thread-local message, int status across the C ABI) ---------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
const char* error_text();

#define LA_CUDA_TRY(expr)                                                                              \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      int _code = (_e == cudaErrorMemoryAllocation) ? LA_ERR_NOMEM                                     \
                  : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver) ? LA_ERR_NO_DEVICE  \
                                                                                   : LA_ERR_CUDA;      \
      return ::la::fail(_code, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                                  \
  } while (0)

#define LA_TRY(expr)                 \
  do {                               \
    int _s = (expr);                 \
    if (_s != LA_OK) return _s;      \
  } while (0)

#define LA_REQUIRE(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) return ::la::fail(LA_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// ---- runtime services (la_runtime.cu) --------------------------------------------------------------
struct DeviceCtx {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t smem_optin = 0;
  bool coop = false;
};
int device_ctx(int device, const DeviceCtx** out);  // validates sm_100, caches properties
int current_device_ctx(const DeviceCtx** out);
cudaStream_t resolve_stream(void* s);               // NULL -> cudaStreamPerThread
// Grow-only per-(thread,device) scratch used by the *_host entry points and the LU driver.
int scratch_get(int device, int slot, size_t bytes, void** out);

// The scratch pool, the LU / Cholesky side streams and their events are per (host thread, device), not per stream.  Two
// library calls from one thread on DIFFERENT streams would therefore share them while both are in flight.  Every entry
// point that queues work on a caller-visible stream opens a CallScope: when the stream differs from the one the thread's
// previous call used on this device, the new stream first waits for that call's tail (one event per thread and device).
struct CallScope {
  int device;
  cudaStream_t st;
  CallScope(int device, cudaStream_t st);
  ~CallScope();
};

// cuTensorMapEncodeTiled fetched through the runtime (no link-time libcuda dependency).
int encode_tensor_map_2d(CUtensorMap* map, CUtensorMapDataType dtype, size_t elem_bytes, const void* base,
                         uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner,
                         uint32_t box_outer, CUtensorMapSwizzle swizzle);

// ---- typed internal entry points shared between translation units ------------------------------------
int gemm_f64_dev(const double* A, size_t lda, const double* B, size_t ldb, double* C, size_t ldc, size_t m, size_t k,
                 size_t n, int mode, cudaStream_t st);
int gemm_f32_dev(const float* A, size_t lda, const float* B, size_t ldb, float* C, size_t ldc, size_t m, size_t k,
                 size_t n, int mode, cudaStream_t st);
template <typename T>
int gemm_dev(const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc, size_t m, size_t k, size_t n, int mode,
             cudaStream_t st);
template <typename T>
int gemm_simt(const T* A, size_t lda, const T* B, size_t ldb, T* C, size_t ldc, size_t m, size_t k, size_t n, int mode,
              cudaStream_t st);

// CUDA loads kernels lazily; the first launch of a kernel may have to wait for the device -- which never happens while a
// kernel that is itself waiting for that launch is spinning.  The multi-GPU contexts therefore load every kernel a product
// can launch up front (la_mg_connect).
int gemm_f64_preload();
int gemm_f32_preload();
int gemm_simt_preload();

template <typename T>
int lu_factor_dev(T* LU, size_t m, size_t n, uint64_t* piv_dev, int* sign_dev, cudaStream_t st);
template <typename T>
int lu_solve_dev(const T* LU, size_t n, const uint64_t* piv_dev, const T* B, size_t nx, T* X, cudaStream_t st);
template <typename T>
int lu_is_nonsingular_dev(const T* LU, size_t n, int* out_host, cudaStream_t st);
template <typename T>
int lu_det_dev(const T* LU, size_t n, int pospivsign, T* out_host, cudaStream_t st);
template <typename T>
int identity_dev(T* dst, size_t n, cudaStream_t st);
template <typename T>
int chol_factor_dev(T* A, size_t n, int* flags_dev, cudaStream_t st);  // cholesky.cu
template <typename T>
int chol_solve_dev(const T* L, size_t n, const T* B, size_t nx, T* X, cudaStream_t st);
template <typename T>
int elementwise_dev(int op, const T* A, const T* B, T scalar, T* C, size_t count, cudaStream_t st);  // elementwise.cu
template <typename T>
int reduce_dev(int kind, const T* A, const T* B, size_t count, T* out_host, cudaStream_t st);
template <typename T>
int qr_tmat_elems(size_t m, size_t n, size_t* out);  // qr.cu
template <typename T>
int qr_factor_dev(T* QR, size_t m, size_t n, T* rdiag, T* tmat, cudaStream_t st);
template <typename T>
int qr_get_r_dev(const T* QR, size_t m, size_t n, const T* rdiag, T* R, cudaStream_t st);
template <typename T>
int qr_get_q_dev(const T* QR, size_t m, size_t n, const T* tmat, T* Q, cudaStream_t st);
template <typename T>
int qr_solve_dev(const T* QR, size_t m, size_t n, const T* rdiag, const T* B, size_t nx, T* X, cudaStream_t st);
template <typename T>
int transpose_dev(const T* src, T* dst, size_t rows, size_t cols, cudaStream_t st);
template <typename T>
int permute_rows_dev(const T* src, T* dst, const uint64_t* idx_dev, size_t out_rows, size_t cols, cudaStream_t st);
template <typename T>
int fill_hash_dev(T* dst, size_t count, uint64_t seed, uint64_t first_idx, cudaStream_t st);

#ifdef __CUDACC__
// ---- device-side PTX wrappers (mbarrier / TMA / DMMA) ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 2-D tiled TMA load: coordinates are (inner, outer) in elements.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c_inner,
                                            int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 v;
#if defined(LA_GEMM_VARIANT) && LA_GEMM_VARIANT == 3
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
#else
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
#endif
  return v;
}
// D(8x8) += A(8x4, row) * B(4x8, col), fp64.  SASS: DMMA.8x8x4 (the only fp64 MMA shape sm_100a issues).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Separately rounded multiply / add / subtract: never contracted into FMA by the compiler.  The reference (Rust) rounds
// the product and the sum separately (src/matrix/mod.rs:969, src/decomp/lu.rs:125-128, :260, :272); the small-problem
// and panel paths use these so that results are bit-identical to the reference where the operation ORDER is also kept.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }

// counter-based hash shared by la_fill_hash_* and the oracle
__host__ __device__ __forceinline__ uint64_t hash64(uint64_t seed, uint64_t idx) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + idx;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
#endif  // __CUDACC__

}  // namespace la
