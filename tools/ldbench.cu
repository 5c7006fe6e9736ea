// Micro-benchmark: latency of a batch of N independent 16-byte loads per lane (one warp), for different load flavours
// and lane strides.  Answers: do strong (relaxed.gpu / volatile) loads overlap, and what does a poll round cost?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__device__ __forceinline__ void ld16(const void* p, unsigned long long& a, unsigned long long& b) {
  if (MODE == 0) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  if (MODE == 1) asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  if (MODE == 2) asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
  if (MODE == 3) asm volatile("ld.global.cv.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
template <int MODE, int N>
__global__ void k(const char* buf, int lane_stride, int inst_stride, long long* out, unsigned long long* sink) {
  const int lane = threadIdx.x;
  unsigned long long a[N], b[N];
  long long best = 1ll << 60;
  unsigned long long acc = 0;
  for (int rep = 0; rep < 20; ++rep) {
    const char* base = buf + (size_t)rep * 1048576 + blockIdx.x * 65536;
    __syncwarp();
    long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < N; ++i) ld16<MODE>(base + (size_t)lane * lane_stride + (size_t)i * inst_stride, a[i], b[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) acc += a[i] ^ b[i];
    long long t1 = clock64();
    if (acc == 0x1234567) t1 += 1;
    if (rep > 2 && t1 - t0 < best) best = t1 - t0;
  }
  if (lane == 0) out[blockIdx.x] = best;
  sink[threadIdx.x] = acc;
}
template <int MODE, int N>
void run(const char* name, const char* buf, int ls, int is, int ctas) {
  long long* out;
  unsigned long long* sink;
  cudaMalloc(&out, 8 * 256);
  cudaMalloc(&sink, 8 * 64);
  k<MODE, N><<<ctas, 32>>>(buf, ls, is, out, sink);
  long long h[256];
  cudaMemcpy(h, out, 8 * ctas, cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (int i = 0; i < ctas; ++i) { if (h[i] > mx) mx = h[i]; if (h[i] < mn) mn = h[i]; }
  printf("%-10s N=%2d lane_stride=%3d inst_stride=%5d ctas=%3d : %lld..%lld cycles\n", name, N, ls, is, ctas, mn, mx);
  cudaFree(out); cudaFree(sink);
}
int main() {
  char* buf;
  cudaMalloc(&buf, 64 << 20);
  cudaMemset(buf, 1, 64 << 20);
  for (int ctas : {1, 148}) {
    run<0, 1>("relaxed", buf, 32, 1024, ctas);
    run<0, 10>("relaxed", buf, 32, 1024, ctas);
    run<0, 10>("relaxed", buf, 16, 512, ctas);
    run<1, 10>("volatile", buf, 32, 1024, ctas);
    run<2, 10>("cg", buf, 32, 1024, ctas);
    run<2, 1>("cg", buf, 32, 1024, ctas);
    run<3, 10>("cv", buf, 32, 1024, ctas);
    run<0, 4>("relaxed", buf, 16, 512, ctas);
  }
  return 0;
}
