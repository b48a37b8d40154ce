// TEST INFRASTRUCTURE (oracle/): a minimal CUDA-on-CPU execution shim, just enough to run the plain SIMT kernels of
// csrc/train_ops.cu and csrc/loss_ops.cu on host threads so that their indexing, shared-memory reductions and warp shuffles
// can be checked against oracle/train_oracle.py in a container without a GPU (tests/test_emu_kernels.py).
// One OS thread per CUDA thread of a block, blocks run one after another; __syncthreads / __syncwarp / __shfl_xor_sync are
// barriers + exchange buffers; atomics take a global lock.  Nothing here is shipped or measured.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct emu_idx { unsigned x, y, z; };
typedef void* cudaStream_t;
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr int cudaFuncAttributeMaxDynamicSharedMemorySize = 0;
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline float2 make_float2(float a, float b) { return float2{a, b}; }

// bf16 with round-to-nearest-even conversion
struct __nv_bfloat16 { uint16_t v; };
inline float __bfloat162float(__nv_bfloat16 h) { uint32_t u = (uint32_t)h.v << 16; float f; memcpy(&f, &u, 4); return f; }
inline __nv_bfloat16 __float2bfloat16_rn(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{(uint16_t)0x7fc0};
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{(uint16_t)(u >> 16)};
}
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
inline float2 __bfloat1622float2(__nv_bfloat162 h) { return float2{__bfloat162float(h.x), __bfloat162float(h.y)}; }
inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }

template <typename A, typename B> inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
template <typename A, typename B> inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __expf(float x) { return expf(x); }
#ifndef INFINITY
#define INFINITY (__builtin_inff())
#endif

namespace emu {
struct Block {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<std::unique_ptr<std::barrier<>>> wbar;
  std::vector<uint64_t> xch;      // per thread exchange slot
  std::vector<uint32_t> frag;     // per thread: 4 A + 2 B fragment registers of an emulated mma.sync
};
extern thread_local emu_idx t_idx;
extern thread_local int t_lin;
extern emu_idx b_idx;
extern dim3 b_dim, g_dim;
extern Block* blk;
extern unsigned char* dyn_smem;
extern std::mutex atomic_lock;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace emu
#define threadIdx (emu::t_idx)
#define blockIdx (emu::b_idx)
#define blockDim (emu::b_dim)
#define gridDim (emu::g_dim)

inline void __syncthreads() { emu::blk->bar->arrive_and_wait(); }
inline void __syncwarp() { emu::blk->wbar[emu::t_lin >> 5]->arrive_and_wait(); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  emu::blk->xch[emu::t_lin] = raw;
  __syncwarp();
  uint64_t other = emu::blk->xch[(emu::t_lin & ~31) | ((emu::t_lin ^ o) & 31)];
  __syncwarp();
  T r; memcpy(&r, &other, sizeof(T));
  return r;
}
inline int atomicMax(int* p, int v) {
  std::lock_guard<std::mutex> g(emu::atomic_lock);
  int old = *p; if (v > old) *p = v; return old;
}
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
template <typename T> inline T atomicAdd(T* p, T v) {
  std::lock_guard<std::mutex> g(emu::atomic_lock);
  T old = *p; *p = old + v; return old;
}

// ---- warp-level tensor-core primitives, host-thread versions (fragment layouts as documented in the PTX ISA for
// ldmatrix .m8n8.x4(.trans).b16 and mma.sync.m16n8k16.row.col bf16) ----
inline void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  if (valid) memcpy(smem_dst, gmem_src, 16); else memset(smem_dst, 0, 16);
}
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
inline void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  // lane 8j + q supplies the address of row q of 8x8 matrix j; with .trans thread i receives, from each matrix,
  // the pair (M[2*(i%4)][i/4], M[2*(i%4)+1][i/4]) packed low / high
  const int lane = emu::t_lin & 31, base = emu::t_lin & ~31;
  emu::blk->xch[emu::t_lin] = (uint64_t)(uintptr_t)smem_row;
  __syncwarp();
  for (int j = 0; j < 4; ++j) {
    const uint16_t* row_lo = (const uint16_t*)(uintptr_t)emu::blk->xch[base + 8 * j + 2 * (lane % 4)];
    const uint16_t* row_hi = (const uint16_t*)(uintptr_t)emu::blk->xch[base + 8 * j + 2 * (lane % 4) + 1];
    r[j] = (uint32_t)row_lo[lane / 4] | ((uint32_t)row_hi[lane / 4] << 16);
  }
  __syncwarp();
}
inline float emu_bf16_bits(uint32_t reg, int half) {
  uint32_t u = ((reg >> (16 * half)) & 0xffffu) << 16; float f; memcpy(&f, &u, 4); return f;
}
inline void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  // A (16x16, row): a0 = (row g, k 2t..2t+1), a1 = (row g+8, same k), a2 = (row g, k 2t+8..), a3 = (row g+8, k 2t+8..)
  // B (16x8, col):  b0 = (k 2t..2t+1, n g), b1 = (k 2t+8.., n g);  C/D: c0,c1 = (row g, n 2t, 2t+1), c2,c3 = (row g+8, ...)
  const int lane = emu::t_lin & 31, base = emu::t_lin & ~31;
  uint32_t* f = emu::blk->frag.data();
  for (int i = 0; i < 4; ++i) f[(size_t)emu::t_lin * 6 + i] = a[i];
  f[(size_t)emu::t_lin * 6 + 4] = b0;
  f[(size_t)emu::t_lin * 6 + 5] = b1;
  __syncwarp();
  const int g = lane >> 2, t = lane & 3;
  for (int e = 0; e < 4; ++e) {
    const int row = g + 8 * (e >> 1), n = 2 * t + (e & 1);
    float sum = 0.f;
    for (int k = 0; k < 16; ++k) {
      const int la = (row % 8) * 4 + (k % 8) / 2, ra = (row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0);
      const int lb = n * 4 + (k % 8) / 2, rb = 4 + (k >= 8 ? 1 : 0);
      sum += emu_bf16_bits(f[(size_t)(base + la) * 6 + ra], k % 2) * emu_bf16_bits(f[(size_t)(base + lb) * 6 + rb], k % 2);
    }
    c[e] += sum;
  }
  __syncwarp();
}
