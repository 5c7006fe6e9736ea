// elementwise.cu -- device-resident elementwise operators and norms of Matrix<T>, fp64/fp32.
//
// Replaces (reference, src/matrix/mod.rs): scale :487-497, elem_mul :499-512, elem_div :514-527, Neg :853-866, Add :874-890,
// Sub :913-929 (one IEEE operation per element: results are bit-identical to the reference's) and dot :529-, the vector norms
// :1059-1109 (vector_euclidean_norm / frobenius_norm: sqrt of the sum of squares, vector_1_norm, vector_inf_norm).
// All HBM-bound streams: 16-byte accesses, grid = a multiple of the SM count, 8 independent loads in flight per thread.
// Reductions are two-stage with a fixed shape (per-thread strided partials -> warp tree -> block -> one final block), so
// a result is reproducible run to run on one device; the reference sums sequentially, the difference is rounding only.
#include "la_common.cuh"

namespace la {
namespace {

constexpr int EW_THREADS = 256;
constexpr int RED_BLOCKS_PER_SM = 4;

template <typename T, int OP>
__device__ __forceinline__ T ew_apply(T a, T b, T s) {
  if (OP == LA_EW_ADD) return add_rn(a, b);
  if (OP == LA_EW_SUB) return sub_rn(a, b);
  if (OP == LA_EW_MUL) return mul_rn(a, b);
  if (OP == LA_EW_DIV) return a / b;
  if (OP == LA_EW_SCALE) return mul_rn(s, a);  // factor * self.data[i]
  return -a;                                   // LA_EW_NEG
}

template <typename T>
struct Vec16;
template <>
struct Vec16<double> {
  typedef double2 type;
  static constexpr int N = 2;
};
template <>
struct Vec16<float> {
  typedef float4 type;
  static constexpr int N = 4;
};

template <typename T, int OP>
__global__ void __launch_bounds__(EW_THREADS) ew_kernel(const T* __restrict__ A, const T* __restrict__ B, T s,
                                                        T* __restrict__ C, size_t count, int vec_ok) {
  constexpr bool BINARY = OP == LA_EW_ADD || OP == LA_EW_SUB || OP == LA_EW_MUL || OP == LA_EW_DIV;
  typedef typename Vec16<T>::type V;
  constexpr int N = Vec16<T>::N;
  const size_t tid = (size_t)blockIdx.x * EW_THREADS + threadIdx.x, nthr = (size_t)gridDim.x * EW_THREADS;
  size_t done = 0;
  if (vec_ok) {
    const size_t nv = count / N;
    const V* Av = reinterpret_cast<const V*>(A);
    const V* Bv = reinterpret_cast<const V*>(B);
    V* Cv = reinterpret_cast<V*>(C);
    for (size_t i = tid; i < nv; i += nthr) {
      V a = Av[i], b = a, c;
      if (BINARY) b = Bv[i];
      T* ap = reinterpret_cast<T*>(&a);
      T* bp = reinterpret_cast<T*>(&b);
      T* cp = reinterpret_cast<T*>(&c);
#pragma unroll
      for (int e = 0; e < N; ++e) cp[e] = ew_apply<T, OP>(ap[e], bp[e], s);
      Cv[i] = c;
    }
    done = nv * N;
  }
  for (size_t i = done + tid; i < count; i += nthr) C[i] = ew_apply<T, OP>(A[i], BINARY ? B[i] : A[i], s);
}

template <typename T, int KIND>
__device__ __forceinline__ T red_term(T a, T b) {
  if (KIND == LA_RED_SUMSQ) return a * a;
  if (KIND == LA_RED_ABS_SUM) return fabs(a);
  if (KIND == LA_RED_ABS_MAX) return fabs(a);
  return a * b;  // LA_RED_DOT
}
template <typename T, int KIND>
__device__ __forceinline__ T red_join(T x, T y) {
  if (KIND == LA_RED_ABS_MAX) return (y > x) ? y : x;  // `if v > current_max` (mod.rs:1101): NaN never replaces
  return x + y;
}

template <typename T, int KIND>
__device__ __forceinline__ T block_reduce(T v, T* sh) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = red_join<T, KIND>(v, __shfl_xor_sync(0xffffffffu, v, off));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  T r = (T)0;
  if (warp == 0) {
    r = lane < EW_THREADS / 32 ? sh[lane] : (T)0;
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) r = red_join<T, KIND>(r, __shfl_xor_sync(0xffffffffu, r, off));
  }
  return r;  // valid in thread 0
}

template <typename T, int KIND>
__global__ void __launch_bounds__(EW_THREADS) red_stage1_kernel(const T* __restrict__ A, const T* __restrict__ B,
                                                                size_t count, T* __restrict__ partial) {
  __shared__ T sh[EW_THREADS / 32];
  const size_t tid = (size_t)blockIdx.x * EW_THREADS + threadIdx.x, nthr = (size_t)gridDim.x * EW_THREADS;
  T acc[4] = {(T)0, (T)0, (T)0, (T)0};
  size_t i = tid;
  for (; i + 3 * nthr < count; i += 4 * nthr) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t e = i + (size_t)u * nthr;
      acc[u] = red_join<T, KIND>(acc[u], red_term<T, KIND>(A[e], KIND == LA_RED_DOT ? B[e] : A[e]));
    }
  }
  for (; i < count; i += nthr) acc[0] = red_join<T, KIND>(acc[0], red_term<T, KIND>(A[i], KIND == LA_RED_DOT ? B[i] : A[i]));
  T v = red_join<T, KIND>(red_join<T, KIND>(acc[0], acc[1]), red_join<T, KIND>(acc[2], acc[3]));
  v = block_reduce<T, KIND>(v, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = v;
}
template <typename T, int KIND>
__global__ void __launch_bounds__(EW_THREADS) red_stage2_kernel(const T* __restrict__ partial, int n, T* __restrict__ out) {
  __shared__ T sh[EW_THREADS / 32];
  T v = (T)0;
  for (int i = threadIdx.x; i < n; i += EW_THREADS) v = red_join<T, KIND>(v, partial[i]);
  v = block_reduce<T, KIND>(v, sh);
  if (threadIdx.x == 0) *out = (KIND == LA_RED_SUMSQ) ? sqrt(v) : v;
}

template <typename T, int OP>
int ew_launch(const T* A, const T* B, T s, T* C, size_t count, int blocks, cudaStream_t st) {
  const int vec_ok = ((uintptr_t)A % 16 == 0) && ((uintptr_t)C % 16 == 0) && (!B || (uintptr_t)B % 16 == 0);
  ew_kernel<T, OP><<<blocks, EW_THREADS, 0, st>>>(A, B, s, C, count, vec_ok);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template <typename T, int KIND>
int red_launch(const T* A, const T* B, size_t count, T* partial, int blocks, cudaStream_t st) {
  red_stage1_kernel<T, KIND><<<blocks, EW_THREADS, 0, st>>>(A, B, count, partial);
  red_stage2_kernel<T, KIND><<<1, EW_THREADS, 0, st>>>(partial, blocks, partial + blocks);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}

}  // namespace

template <typename T>
int elementwise_dev(int op, const T* A, const T* B, T scalar, T* C, size_t count, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(A && C && count > 0, "la_elementwise: null pointer or zero count");
  const bool binary = op == LA_EW_ADD || op == LA_EW_SUB || op == LA_EW_MUL || op == LA_EW_DIV;
  LA_REQUIRE(!binary || B, "la_elementwise: operator %d needs a second operand", op);
  size_t want = (count / (16 / sizeof(T)) + EW_THREADS - 1) / EW_THREADS;
  const size_t cap = (size_t)ctx->sm_count * 16;
  const int blocks = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  switch (op) {
    case LA_EW_ADD: return ew_launch<T, LA_EW_ADD>(A, B, scalar, C, count, blocks, st);
    case LA_EW_SUB: return ew_launch<T, LA_EW_SUB>(A, B, scalar, C, count, blocks, st);
    case LA_EW_MUL: return ew_launch<T, LA_EW_MUL>(A, B, scalar, C, count, blocks, st);
    case LA_EW_DIV: return ew_launch<T, LA_EW_DIV>(A, B, scalar, C, count, blocks, st);
    case LA_EW_SCALE: return ew_launch<T, LA_EW_SCALE>(A, nullptr, scalar, C, count, blocks, st);
    case LA_EW_NEG: return ew_launch<T, LA_EW_NEG>(A, nullptr, scalar, C, count, blocks, st);
  }
  return fail(LA_ERR_INVALID, "la_elementwise: unknown operator %d", op);
}

template <typename T>
int reduce_dev(int kind, const T* A, const T* B, size_t count, T* out_host, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(A && out_host && count > 0, "la_reduce: null pointer or zero count");
  LA_REQUIRE(kind != LA_RED_DOT || B, "la_reduce: the dot product needs a second operand");
  size_t want = (count + EW_THREADS * 8 - 1) / (EW_THREADS * 8);
  const size_t cap = (size_t)ctx->sm_count * RED_BLOCKS_PER_SM;
  const int blocks = (int)(want < 1 ? 1 : (want < cap ? want : cap));
  void* p;
  LA_TRY(scratch_get(ctx->device, 37, sizeof(T) * (size_t)(blocks + 1), &p));
  T* partial = (T*)p;
  int s = LA_OK;
  switch (kind) {
    case LA_RED_SUMSQ: s = red_launch<T, LA_RED_SUMSQ>(A, B, count, partial, blocks, st); break;
    case LA_RED_ABS_SUM: s = red_launch<T, LA_RED_ABS_SUM>(A, B, count, partial, blocks, st); break;
    case LA_RED_ABS_MAX: s = red_launch<T, LA_RED_ABS_MAX>(A, B, count, partial, blocks, st); break;
    case LA_RED_DOT: s = red_launch<T, LA_RED_DOT>(A, B, count, partial, blocks, st); break;
    default: return fail(LA_ERR_INVALID, "la_reduce: unknown kind %d", kind);
  }
  LA_TRY(s);
  LA_CUDA_TRY(cudaMemcpyAsync(out_host, partial + blocks, sizeof(T), cudaMemcpyDeviceToHost, st));
  LA_CUDA_TRY(cudaStreamSynchronize(st));
  return LA_OK;
}

template int elementwise_dev<double>(int, const double*, const double*, double, double*, size_t, cudaStream_t);
template int elementwise_dev<float>(int, const float*, const float*, float, float*, size_t, cudaStream_t);
template int reduce_dev<double>(int, const double*, const double*, size_t, double*, cudaStream_t);
template int reduce_dev<float>(int, const float*, const float*, size_t, float*, cudaStream_t);

}  // namespace la
