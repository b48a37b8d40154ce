// micro-benchmark: MUFU throughput per SM for tanh.approx / ex2.approx / rcp.approx (sm_100a)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(float* out, int iters) {
  float v[8];
  for (int j = 0; j < 8; ++j) v[j] = threadIdx.x * 0.001f + j * 0.1f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[j]));
    }
  }
  float s = 0; for (int j = 0; j < 8; ++j) s += v[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, float* d) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000, threads = 1024;
  k<OP><<<148, threads>>>(d, 10);
  cudaEventRecord(e0); k<OP><<<148, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)iters * 8 * threads;   // per SM
  printf("%-6s %.3f ms -> %.2f ops/ns/SM (at ~1.9 GHz: %.1f per clk per SM)\n", name, ms, ops / (ms * 1e6), ops / (ms * 1e6) / 1.9);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  run<0>("tanh", d); run<1>("ex2", d); run<2>("rcp", d); run<3>("fma", d);
  return 0;
}
