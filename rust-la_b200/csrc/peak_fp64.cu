// peak_fp64.cu -- measures the fp64 issue-rate ceilings of the device so roofline fractions are quoted against a
// MEASURED denominator (MEASURED_PEAKS.json has no fp64 entry):
//   * DMMA : register-resident mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) loop, all SMs
//   * DFMA : register-resident fma.rn.f64 loop, all SMs
// Usage: peak_fp64 [out.json]
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int ACCS>
__global__ void __launch_bounds__(256) dmma_loop(double* out, int iters) {
  double c[ACCS][2];
#pragma unroll
  for (int i = 0; i < ACCS; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACCS; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACCS; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACCS>
__global__ void __launch_bounds__(256) dfma_loop(double* out, int iters) {
  double c[ACCS];
#pragma unroll
  for (int i = 0; i < ACCS; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACCS; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACCS; ++i) s += c[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  double* out;
  CK(cudaMalloc(&out, sizeof(double) * 1024 * 1024));
  const int sms = p.multiProcessorCount;
  const int iters = 4096;
  double best_dmma = 0, best_dfma = 0;
  int best_dmma_w = 0, best_dfma_w = 0;
  // sweep resident warps per SM (256-thread CTAs, 1..4 CTAs per SM)
  for (int ctas = 1; ctas <= 4; ++ctas) {
    {
      constexpr int ACCS = 16;
      double ms = time_ms([&] { dmma_loop<ACCS><<<sms * ctas, 256>>>(out, iters); }, 5);
      double flops = 2.0 * 256 /*fma per mma*/ * ACCS * (double)iters * 8 /*warps*/ * sms * ctas;
      double tf = flops / (ms * 1e-3) / 1e12;
      printf("DMMA  warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", ctas * 8, ms, tf);
      if (tf > best_dmma) { best_dmma = tf; best_dmma_w = ctas * 8; }
    }
    {
      constexpr int ACCS = 16;
      double ms = time_ms([&] { dfma_loop<ACCS><<<sms * ctas, 256>>>(out, iters * 4); }, 5);
      double flops = 2.0 * ACCS * (double)iters * 4 * 256 * sms * ctas;
      double tf = flops / (ms * 1e-3) / 1e12;
      printf("DFMA  warps/SM=%2d  %.3f ms  %.2f TFLOP/s\n", ctas * 8, ms, tf);
      if (tf > best_dfma) { best_dfma = tf; best_dfma_w = ctas * 8; }
    }
  }
  // sustained: ~2 s of back-to-back DMMA launches (power-capped clocks)
  double sustained = 0;
  {
    constexpr int ACCS = 16;
    const int ctas = best_dmma_w / 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    int launches = 0;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 120; ++i) { dmma_loop<ACCS><<<sms * ctas, 256>>>(out, iters * 4); ++launches; }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double flops = 2.0 * 256 * ACCS * (double)iters * 4 * 8 * sms * ctas * launches;
    sustained = flops / (ms * 1e-3) / 1e12;
    printf("DMMA sustained over %.1f s: %.2f TFLOP/s\n", ms * 1e-3, sustained);
  }
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("device %s  SMs %d  max clock %d MHz\n", p.name, sms, clk_khz / 1000);
  const char* path = argc > 1 ? argv[1] : "peak_fp64.json";
  FILE* f = fopen(path, "w");
  if (f) {
    fprintf(f,
            "{\"gpu_name\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %d, \"dmma_tflops\": %.3f, \"dmma_warps_per_sm\": %d, "
            "\"dmma_tflops_sustained\": %.3f, \"dfma_tflops\": %.3f, \"dfma_warps_per_sm\": %d, "
            "\"how\": \"register-resident mma.sync.m8n8k4.f64 / fma.rn.f64 loops, 16 independent accumulators per thread, "
            "best of 5, CUDA events; sustained = 120 back-to-back launches (~2.5 s)\"}\n",
            p.name, sms, clk_khz / 1000, best_dmma, best_dmma_w, sustained, best_dfma, best_dfma_w);
    fclose(f);
  }
  return 0;
}
