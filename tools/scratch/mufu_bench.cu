// Throughput of ex2.approx.ftz.f32 / FFMA per SM on this GPU (one kernel, all SMs, many warps)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i])); }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  if (s == 1.2345f) out[0] = s;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* d; cudaMalloc(&d, 4);
  int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, warps_per_sm * 32>>>(d, 100);
  cudaEventRecord(e0);
  k<MODE><<<sms, warps_per_sm * 32>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  double ops = (double)sms * warps_per_sm * 32 * iters * 8;
  printf("%-22s warps/SM %2d: %.3f ms, %.1f Gop/s/SM -> %.1f op/clk/SM at %.2f GHz nominal\n", name, warps_per_sm, ms,
         ops / ms / 1e6 / sms, ops / ms / 1e6 / sms / (clk_khz / 1e6), clk_khz / 1e6);
}
int main() {
  for (int w : {4, 8, 16, 32}) run<0>("ex2.approx", w);
  for (int w : {8, 32}) run<1>("ffma", w);
  for (int w : {8, 32}) run<2>("ex2 + 4 ffma", w);
  return 0;
}
