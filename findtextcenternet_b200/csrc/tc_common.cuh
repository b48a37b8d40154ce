// Shared device helpers of the tcgen05 convolution kernels (conv_gemm_tc.cu: cp.async im2col producers;
// conv_gemm_tma.cu: TMA tensor-tile producers): mbarrier / cp.async / TMA / tcgen05 PTX wrappers, the 128B-swizzle
// K-major shared-memory descriptor and the common epilogue (scale, bias table, activation, residuals, store).
#pragma once
#include "conv_gemm.cuh"

// Ablation bits of ConvTcPlan::flags (FTC_TMA_FLAGS / ftc_debug_set_gemm_tuning) are compiled in only with -DFTC_ABLATION
// (python -m findtextcenternet_b200.build --ablation): the runtime tests cost ~30 of the 230 instructions of an epilogue chunk.
#ifdef FTC_ABLATION
#define FTC_ABL(bit) (p.tc.flags & (bit))
#else
#define FTC_ABL(bit) (false)
#endif

namespace ftc {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// bounded wait: a broken pipeline traps (launch failure) instead of hanging the device.
// backoff_ns > 0: sleep between polls (warps off the critical path); hint_ns > 0: try_wait suspend-time hint
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t backoff_ns = 0, uint32_t hint_ns = 0) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t spins = 0;
  long long t0 = 0;
  for (;;) {
    if (hint_ns ? mbar_try_wait_hint(bar, parity, hint_ns) : mbar_try_wait(bar, parity)) return;
    if (backoff_ns) __nanosleep(backoff_ns);
    if ((++spins & 0xFFFu) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 8000000000LL) __trap();
    }
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  // the barrier receives this thread's arrival when all of its earlier cp.async copies have landed (non-blocking)
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {   // one lane of the (fully active) warp, the same one every time
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzle shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}


// K-major, 64B-swizzle variant (rows of 64 B, 8-row groups 512 B apart): the 32-channel k-blocks of the <= 32-channel 3x3 convs
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
}

// SiLU with ONE SFU op per element: x*sigmoid(x) = h + h*tanh(h), h = x/2 (tanh.approx.f32, rel. error ~2^-11:
// below bf16 output rounding).  The exp+rcp form costs two MUFU ops and made N=256 SiLU epilogues SFU-bound.
__device__ __forceinline__ float silu_tanh(float x) {
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// erf-GELU with ONE SFU op per element: x*Phi(x), Phi(x) ~ 0.5 + 0.5*tanh(x*(a + b*x^2 + c*x^4)) (minimax fit over
// [-8, 8]: max abs error 2.5e-5 plus tanh.approx's 2^-11, well below the bf16 rounding of the stored output).  erff()
// costs ~25 instructions per element and made the (non-overlapped) 256x192 head epilogues 30% of those kernels.
__device__ __forceinline__ float gelu_tanh3(float x) {
  const float x2 = fminf(x * x, 64.f);
  const float u = x * fmaf(x2, fmaf(x2, -3.51516792e-04f, 3.70056461e-02f), 7.97507884e-01f);
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return fmaf(h, t, h);
}

// one row x 16 accumulator columns: scale / bias / activation / residuals / store.
// sscale / sbias point at this tile's staged per-column vectors in shared memory (column c0 of the tile).
__device__ __forceinline__ void epilogue_store(const ConvGemmParams& p, const uint32_t (&raw16)[16], int g, int n0, int m,
                                               int b, int oy, int ox, int hw, int64_t r1row, int nvalid, int chb,
                                               const float* __restrict__ sscale, const float* __restrict__ sbias,
                                               const bf16* __restrict__ res1, const bf16* __restrict__ res2) {
  float v[16];
  const bool full16 = n0 + 16 <= p.N;
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const float4 sc = *reinterpret_cast<const float4*>(sscale + i);
    const float4 bi = *reinterpret_cast<const float4*>(sbias + i);
    v[i + 0] = fmaf(__uint_as_float(raw16[i + 0]), sc.x, bi.x);
    v[i + 1] = fmaf(__uint_as_float(raw16[i + 1]), sc.y, bi.y);
    v[i + 2] = fmaf(__uint_as_float(raw16[i + 2]), sc.z, bi.z);
    v[i + 3] = fmaf(__uint_as_float(raw16[i + 3]), sc.w, bi.w);
  }
  if (p.act == ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = silu_tanh(v[i]);
  } else if (p.act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_tanh3(v[i]);
  }
  const bool dbg_nostore = FTC_ABL(32) && v[0] != 12345.678f;   // ablation: keep the math, drop the stores
  if (res1 && !FTC_ABL(128)) {
    const bf16* rp = res1 + r1row * p.res1_stride + (int64_t)g * p.N + n0;
    if (full16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
      float a[8], c[8];
      load8(rp, a); load8(rp + 8, c);
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] += a[i]; v[8 + i] += c[i]; }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (n0 + i < p.N) v[i] += __bfloat162float(rp[i]);
    }
  }
  if (res2 && !FTC_ABL(128)) {
    const bf16* rp = res2 + (int64_t)m * p.res2_stride + (int64_t)g * p.N + n0;
    if (full16 && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
      float a[8], c[8];
      load8(rp, a); load8(rp + 8, c);
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] += a[i]; v[8 + i] += c[i]; }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (n0 + i < p.N) v[i] += __bfloat162float(rp[i]);
    }
  }
  if (dbg_nostore) return;
  if (p.act == ACT_SWIGLU) {
    // interleaved (x1, xg) column pairs -> x1 * silu(xg), 8 outputs per 16 columns
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = v[2 * i] * silu_f(v[2 * i + 1]);
    bf16* op = reinterpret_cast<bf16*>(p.out) + (int64_t)m * p.out_stride + chb + (n0 >> 1);
    if (n0 + 16 <= nvalid && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
      store8(op, o);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) if (n0 + 2 * i + 1 < nvalid) op[i] = __float2bfloat16_rn(o[i]);
    }
    return;
  }
  if (p.out_layout == OUT_NCHW_F32) {
    float* op = reinterpret_cast<float*>(p.out) + (((int64_t)b * p.out_stride + chb + n0) * p.Ho + oy) * p.Wo + ox;
#pragma unroll
    for (int i = 0; i < 16; ++i) if (n0 + i < nvalid) op[(int64_t)i * hw] = v[i];
  } else if (p.out_layout == OUT_NHWC_F32) {
    float* op = reinterpret_cast<float*>(p.out) + (int64_t)m * p.out_stride + chb + n0;
    if (n0 + 16 <= nvalid && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (n0 + i < nvalid) op[i] = v[i];
    }
  } else {
    bf16* op = reinterpret_cast<bf16*>(p.out) + (int64_t)m * p.out_stride + chb + n0;
    if (n0 + 16 <= nvalid && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
      float lo[8], hi[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { lo[i] = v[i]; hi[i] = v[8 + i]; }
      store8(op, lo); store8(op + 8, hi);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) if (n0 + i < nvalid) op[i] = __float2bfloat16_rn(v[i]);
    }
  }
}

}  // namespace
}  // namespace ftc
