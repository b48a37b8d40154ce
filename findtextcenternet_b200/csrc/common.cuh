// Common device/host helpers for the findtextCenterNet B200 (sm_100a) hot path.
#pragma once
#ifdef FTC_EMU   // oracle/emu: the plain SIMT kernels compiled for host threads (CPU-side kernel checks, never shipped)
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace ftc {

typedef __nv_bfloat16 bf16;

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_SWIGLU = 3 };
enum DType : int { DT_F32 = 0, DT_BF16 = 1 };
enum OutLayout : int { OUT_NHWC = 0, OUT_NCHW_F32 = 1, OUT_NHWC_F32 = 2 };   // NHWC = row-major [M, out_stride]

// error plumbing: C-ABI entry points return negative codes, message kept per thread
void set_error(const std::string& msg);
#define FTC_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ftc::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                     std::to_string(__LINE__));                                                \
      return -2;                                                                               \
    }                                                                                          \
  } while (0)
// every kernel launch site counts itself (ftc_launch_count(), bench.py "gpu_launches") and checks the launch
void count_launch();
#define FTC_POST_LAUNCH()                  \
  do {                                     \
    ftc::count_launch();                   \
    FTC_CHECK_CUDA(cudaGetLastError());    \
  } while (0)
#define FTC_REQUIRE(cond, msg)                                                  \
  do {                                                                          \
    if (!(cond)) {                                                              \
      ftc::set_error(std::string("requirement failed: ") + #cond + " : " + msg); \
      return -1;                                                                \
    }                                                                           \
  } while (0)

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_precise(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

template <bool PRECISE> __device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_SILU) return PRECISE ? silu_precise(v) : silu_f(v);
  if (act == ACT_GELU) return gelu_erf(v);
  return v;
}

// 8 consecutive elements <-> 8 floats
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

#ifndef FTC_EMU
// packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per instruction; the issue-bound epilogues and
// the depthwise kernel use them to halve their FP instruction count)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// Programmatic dependent launch: a kernel launched through launch_pdl may become resident while its predecessor in the
// stream is still draining; it must execute pdl_wait() before it touches anything the predecessor wrote (and before it
// writes anything the predecessor may still read).  pdl_launch_dependents() lets the NEXT kernel's CTAs be scheduled as
// soon as this grid's CTAs have all started.  Hides the ~2-4 us prologue (barrier init, TMEM alloc, descriptor prefetch)
// of each of the ~370 launches of a forward.  FTC_NO_PDL=1 turns the launch attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
int pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#endif  // FTC_EMU

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ftc
