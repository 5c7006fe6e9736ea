// la_runtime.cu -- device discovery, buffers, scratch pools, tensor-map encoding, small fill kernels.
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "la_common.cuh"

namespace la {

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char tl_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
}
const char* error_text() { return tl_error; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
  va_end(ap);
  return code;
}

// ---------------------------------------------------------------------------------------------------
// devices
// ---------------------------------------------------------------------------------------------------
static std::mutex g_dev_mutex;
static std::vector<DeviceCtx> g_devs;  // indexed by ordinal, device == -1 until initialised
static int g_dev_count = -1;

static int load_device_count() {
  if (g_dev_count >= 0) return LA_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(LA_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  g_dev_count = n;
  g_devs.assign(n, DeviceCtx{});
  return LA_OK;
}

int device_ctx(int device, const DeviceCtx** out) {
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  LA_TRY(load_device_count());
  LA_REQUIRE(device >= 0 && device < g_dev_count, "device ordinal %d out of range [0,%d)", device, g_dev_count);
  DeviceCtx& d = g_devs[device];
  if (d.device < 0) {
    cudaDeviceProp p;
    LA_CUDA_TRY(cudaGetDeviceProperties(&p, device));
    if (p.major != 10)
      return fail(LA_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library ships sm_100a code only", device, p.name,
                  p.major, p.minor);
    d.sm_count = p.multiProcessorCount;
    d.cc_major = p.major;
    d.cc_minor = p.minor;
    d.smem_optin = p.sharedMemPerBlockOptin;
    d.coop = p.cooperativeLaunch != 0;
    d.device = device;
  }
  *out = &d;
  return LA_OK;
}

int current_device_ctx(const DeviceCtx** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(LA_ERR_NO_DEVICE, "no CUDA device available (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  }
  return device_ctx(dev, out);
}

cudaStream_t resolve_stream(void* s) { return s ? (cudaStream_t)s : cudaStreamPerThread; }

// ---------------------------------------------------------------------------------------------------
// scratch: grow-only, per host thread, per device, per slot.  Freed when the thread exits.
// ---------------------------------------------------------------------------------------------------
struct ScratchSlot {
  void* ptr = nullptr;
  size_t bytes = 0;
};
struct ScratchPool {
  std::unordered_map<long long, ScratchSlot> slots;
  ~ScratchPool() {
    for (auto& kv : slots)
      if (kv.second.ptr) cudaFree(kv.second.ptr);  // best effort; the context may already be gone
  }
};
static thread_local ScratchPool tl_scratch;

int scratch_get(int device, int slot, size_t bytes, void** out) {
  ScratchSlot& s = tl_scratch.slots[(long long)device * 64 + slot];
  if (s.bytes < bytes) {
    if (s.ptr) {
      LA_CUDA_TRY(cudaStreamSynchronize(cudaStreamPerThread));
      LA_CUDA_TRY(cudaFree(s.ptr));
      s.ptr = nullptr;
      s.bytes = 0;
    }
    size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    LA_CUDA_TRY(cudaMalloc(&s.ptr, want));
    s.bytes = want;
  }
  *out = s.ptr;
  return LA_OK;
}

// ---------------------------------------------------------------------------------------------------
// serialisation of successive calls of one host thread that use different streams (see CallScope in la_common.cuh)
// ---------------------------------------------------------------------------------------------------
struct CallTail {
  cudaEvent_t ev = nullptr;
  cudaStream_t last = nullptr;
  bool valid = false;
};
static thread_local CallTail tl_tail[64];

CallScope::CallScope(int device_, cudaStream_t st_) : device(device_), st(st_) {
  if (device < 0 || device >= 64) return;
  CallTail& t = tl_tail[device];
  if (t.valid && t.last != st) cudaStreamWaitEvent(st, t.ev, 0);  // the previous call's tail, queued on another stream
}
CallScope::~CallScope() {
  if (device < 0 || device >= 64) return;
  CallTail& t = tl_tail[device];
  if (!t.ev && cudaEventCreateWithFlags(&t.ev, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    t.ev = nullptr;
    return;
  }
  if (cudaEventRecord(t.ev, st) == cudaSuccess) {
    t.last = st;
    t.valid = true;
  } else {
    cudaGetLastError();
  }
}

// ---------------------------------------------------------------------------------------------------
// TMA descriptor encoding through the runtime's driver entry point lookup
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_encode_once;

int encode_tensor_map_2d(CUtensorMap* map, CUtensorMapDataType dtype, size_t elem_bytes, const void* base,
                         uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner,
                         uint32_t box_outer, CUtensorMapSwizzle swizzle) {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = (EncodeTiledFn)fn;
    else
      cudaGetLastError();
  });
  if (!g_encode) return fail(LA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  (void)elem_bytes;
  CUresult r = g_encode(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(LA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (base=%p inner=%llu outer=%llu stride=%llu)",
                (int)r, base, (unsigned long long)inner, (unsigned long long)outer,
                (unsigned long long)row_stride_bytes);
  return LA_OK;
}

// ---------------------------------------------------------------------------------------------------
// small bandwidth kernels
// ---------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T hash_to_unit(uint64_t x);
template <>
__device__ __forceinline__ double hash_to_unit<double>(uint64_t x) {
  return (double)(x >> 11) * 0x1.0p-53;
}
template <>
__device__ __forceinline__ float hash_to_unit<float>(uint64_t x) {
  return (float)(x >> 40) * 0x1.0p-24f;
}

template <typename T>
__global__ void fill_hash_kernel(T* __restrict__ dst, size_t count, uint64_t seed, uint64_t first_idx) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    dst[i] = hash_to_unit<T>(hash64(seed, first_idx + i));
}

template <typename T>
int fill_hash_dev(T* dst, size_t count, uint64_t seed, uint64_t first_idx, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  if (count == 0) return LA_OK;
  LA_REQUIRE(dst != nullptr, "la_fill_hash: null destination");
  size_t blocks = (count + 255) / 256;
  size_t cap = (size_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  fill_hash_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(dst, count, seed, first_idx);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template int fill_hash_dev<double>(double*, size_t, uint64_t, uint64_t, cudaStream_t);
template int fill_hash_dev<float>(float*, size_t, uint64_t, uint64_t, cudaStream_t);

// Matrix::id, src/matrix/mod.rs:416-426
template <typename T>
__global__ void identity_kernel(T* __restrict__ dst, size_t n) {
  size_t total = n * n;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    dst[i] = (i / n == i % n) ? (T)1 : (T)0;
}
template <typename T>
int identity_dev(T* dst, size_t n, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(dst != nullptr && n > 0, "la_identity: null destination or n == 0");
  size_t blocks = (n * n + 255) / 256;
  size_t cap = (size_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  identity_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(dst, n);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template int identity_dev<double>(double*, size_t, cudaStream_t);
template int identity_dev<float>(float*, size_t, cudaStream_t);

// dst[c][r] = src[r][c] through 32 x 33 shared-memory tiles: reads and writes are both full 128/256-byte row segments
// (`Matrix::t`, src/matrix/mod.rs:653-669, walks the source with stride cols).
template <typename T, bool LOWER = false>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t rows, size_t cols) {
  __shared__ T tile[32][33];
  const size_t r0 = (size_t)blockIdx.y * 32, c0 = (size_t)blockIdx.x * 32;
  if (LOWER && c0 > r0 + 31) return;  // only the tiles that touch or lie below the diagonal of the source
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const size_t r = r0 + ty + i, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + i][tx] = src[r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const size_t c = c0 + ty + i, r = r0 + tx;
    if (c < cols && r < rows) dst[c * rows + r] = tile[tx][ty + i];
  }
}
template <typename T>
int transpose_dev(const T* src, T* dst, size_t rows, size_t cols, cudaStream_t st) {
  LA_REQUIRE(src && dst && rows > 0 && cols > 0, "la_transpose: null pointer or zero dimension");
  LA_REQUIRE((const void*)src != (const void*)dst, "la_transpose: source and destination must not alias");
  const size_t gx = (cols + 31) / 32, gy = (rows + 31) / 32;
  LA_REQUIRE(gy <= 65535 && gx < (1u << 31), "la_transpose: dimension too large");
  transpose_kernel<T><<<dim3((unsigned)gx, (unsigned)gy), 256, 0, st>>>(src, dst, rows, cols);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
// The same for a lower-triangular source (a Cholesky factor): only the tiles at or below the diagonal are moved, so the
// destination's upper triangle (with the diagonal) is the transpose and everything below its diagonal tiles is left untouched.
int transpose_lower_f64_dev(const double* src, double* dst, size_t n, cudaStream_t st) {
  LA_REQUIRE(src && dst && n > 0 && src != dst, "transpose_lower: bad arguments");
  const size_t g = (n + 31) / 32;
  LA_REQUIRE(g <= 65535, "transpose_lower: dimension too large");
  transpose_kernel<double, true><<<dim3((unsigned)g, (unsigned)g), 256, 0, st>>>(src, dst, n, n);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template int transpose_dev<double>(const double*, double*, size_t, size_t, cudaStream_t);
template int transpose_dev<float>(const float*, float*, size_t, size_t, cudaStream_t);

// dst row i = src row idx[i] (`Matrix::permute_rows`, src/matrix/mod.rs:757-759 -> sub_matrix(rows, ..)); one warp per
// destination row segment, coalesced along the row.
template <typename T>
__global__ void __launch_bounds__(256) permute_rows_kernel(const T* __restrict__ src, T* __restrict__ dst,
                                                           const uint64_t* __restrict__ idx, size_t out_rows, size_t cols) {
  const size_t total = out_rows * cols;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t i = e / cols, j = e - i * cols;
    dst[e] = src[(size_t)idx[i] * cols + j];
  }
}
template <typename T>
int permute_rows_dev(const T* src, T* dst, const uint64_t* idx_dev, size_t out_rows, size_t cols, cudaStream_t st) {
  const DeviceCtx* ctx;
  LA_TRY(current_device_ctx(&ctx));
  LA_REQUIRE(src && dst && idx_dev && out_rows > 0 && cols > 0, "la_permute_rows: null pointer or zero dimension");
  size_t blocks = (out_rows * cols + 255) / 256;
  const size_t cap = (size_t)ctx->sm_count * 16;
  if (blocks > cap) blocks = cap;
  permute_rows_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(src, dst, idx_dev, out_rows, cols);
  LA_CUDA_TRY(cudaGetLastError());
  return LA_OK;
}
template int permute_rows_dev<double>(const double*, double*, const uint64_t*, size_t, size_t, cudaStream_t);
template int permute_rows_dev<float>(const float*, float*, const uint64_t*, size_t, size_t, cudaStream_t);

}  // namespace la
