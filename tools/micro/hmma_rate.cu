// micro-benchmark: legacy mma.sync m16n8k16 bf16 issue rate per SM on sm_100a (does a warp-level-MMA depthwise pay?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, int iters) {
  float c[4][4] = {};
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += c[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4 * 8);
  for (int warps = 4; warps <= 32; warps *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<<<148, warps * 32>>>(d, 100);
    cudaEventRecord(e0); k<<<148, warps * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)iters * 4 * warps;   // HMMAs per SM
    printf("warps/SM %2d: %.3f ms, %.2f ns per HMMA per SM, %.1f TFLOP/s dense-equivalent\n", warps, ms, ms * 1e6 / n, 148 * n * 4096 / (ms * 1e-3) / 1e12);
  }
  return 0;
}
