// Bandwidth-bound detector kernels + device-side weight packing.  See detector_ops.cuh for the
// reference lines each op follows.  All activations are NHWC; fp32 (parity mode) or bf16 (product mode).
#include "detector_ops.cuh"
#include "tma_util.cuh"
#include <math.h>

namespace ftc {

// ------------------------------------------------------------------------------------------------
// stem
template <typename T, int COUT, bool NHWC255>
__global__ void __launch_bounds__(128) stem_conv_kernel(const float* __restrict__ img, T* __restrict__ out, int B,
                                                        int H, int W, const float* __restrict__ w,
                                                        const float* __restrict__ scale, const float* __restrict__ bias) {
  __shared__ __align__(16) float sw[27 * COUT];
  __shared__ float ss[COUT], sb[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) { ss[i] = scale[i]; sb[i] = bias[i]; }
  __syncthreads();
  const int Ho = H / 2, Wo = W / 2;
  int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (int64_t)B * Ho * Wo) return;
  int b = m / (Ho * Wo);
  int r = m - (int64_t)b * Ho * Wo;
  int oy = r / Wo, ox = r - oy * Wo;
  // packed fp32 pairs (FFMA2): the 27 x COUT FMAs per pixel made this kernel issue-bound at 3x its HBM time; same fma order
  f32x2 acc2[COUT / 2];
#pragma unroll
  for (int o = 0; o < COUT / 2; ++o) acc2[o] = pk2(0.f, 0.f);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + kx;
        float v = 0.f;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          if (NHWC255)   // backend-ABI tile: fp32 NHWC 0..255 (process_ocr_torch.py:44 divides by 255 first)
            v = (img[(((int64_t)b * H + iy) * W + ix) * 3 + c] / 255.f) * 2.f - 1.f;
          else
            v = img[(((int64_t)b * 3 + c) * H + iy) * W + ix] * 2.f - 1.f;
        }
        const f32x2 v2 = pk2(v, v);
        const ulonglong2* wk = reinterpret_cast<const ulonglong2*>(&sw[((c * 3 + ky) * 3 + kx) * COUT]);
#pragma unroll
        for (int o = 0; o < COUT / 4; ++o) {
          const ulonglong2 w4 = wk[o];
          acc2[2 * o] = ffma2(v2, w4.x, acc2[2 * o]);
          acc2[2 * o + 1] = ffma2(v2, w4.y, acc2[2 * o + 1]);
        }
      }
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT / 2; ++o) upk2(acc2[o], acc[2 * o], acc[2 * o + 1]);
  T* op = out + m * COUT;
#pragma unroll
  for (int o0 = 0; o0 < COUT; o0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = acc[o0 + j] * ss[o0 + j] + sb[o0 + j];
      v[j] = sizeof(T) == 4 ? silu_precise(t) : silu_f(t);
    }
    store8(op + o0, v);
  }
}

int stem_conv(const float* img, int nhwc255, void* out, int dtype, int B, int H, int W, int Cout, const float* w,
              const float* scale, const float* bias, cudaStream_t s) {
  FTC_REQUIRE(H % 2 == 0 && W % 2 == 0, "stem needs even H, W");
  int64_t M = (int64_t)B * (H / 2) * (W / 2);
  int grid = (int)((M + 127) / 128);
#define LAUNCH(TT, CC)                                                                            \
  do {                                                                                            \
    if (nhwc255) stem_conv_kernel<TT, CC, true><<<grid, 128, 0, s>>>(img, (TT*)out, B, H, W, w, scale, bias); \
    else stem_conv_kernel<TT, CC, false><<<grid, 128, 0, s>>>(img, (TT*)out, B, H, W, w, scale, bias);        \
  } while (0)
  if (Cout == 32) { if (dtype == DT_F32) LAUNCH(float, 32); else LAUNCH(bf16, 32); }
  else if (Cout == 24) { if (dtype == DT_F32) LAUNCH(float, 24); else LAUNCH(bf16, 24); }
  else FTC_REQUIRE(false, "stem Cout must be 24 or 32");
#undef LAUNCH
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 + BN + SiLU + SE partial sums.
// One CTA = a TH x TW output tile x 64 channels.  The halo tile is staged ONCE in shared memory with full-line
// (128 B per pixel) cp.async loads, so HBM/L2 sees each input ~1.2x instead of 9x.  Each thread owns one output column
// and 4 channels and walks down the tile with a 3-row register window: per output element 9 FMA + 3 conversions +
// 0.75 shared loads (the first version spent ~65 instructions per element, this one ~20: the op is instruction-bound
// on the CUDA cores, not HBM-bound, until that count is small).
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float silu_tanh_f(float x) {   // 1 SFU op: x*sigmoid(x) = h + h*tanh(h), h = x/2
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

template <typename T, int STRIDE>
__global__ void __launch_bounds__(256) dwconv3x3_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W,
                                                        int C, int Ho, int Wo, int TH, int TW, int tiles_x,
                                                        const float* __restrict__ w, const float* __restrict__ scale,
                                                        const float* __restrict__ bias, float* __restrict__ se_sum) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  T* tile = reinterpret_cast<T*>(dw_smem);                       // [IH][IW][64]
  const int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;
  float* red = reinterpret_cast<float*>(dw_smem + (((size_t)IH * IW * 64 * sizeof(T)) + 15) / 16 * 16);   // [TW][64]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.z;
  const int cb = blockIdx.y * 64;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int oy0 = ty * TH, ox0 = tx * TW;
  const int iy0 = oy0 * STRIDE - 1, ix0 = ox0 * STRIDE - 1;
  // ---- stage the halo tile (zeros outside the image): 16-byte cp.async per piece, all in flight at once ----
  {
    constexpr int PPP = (int)(64 * sizeof(T) / 16);   // 16-byte pieces per pixel: 8 (bf16) or 16 (fp32)
    constexpr int EPP = (int)(16 / sizeof(T));        // elements per piece
    for (int i = tid; i < IH * IW * PPP; i += nthr) {
      const int pix = i / PPP, piece = i - pix * PPP;
      const int py = pix / IW, px = pix - py * IW;
      const int iy = iy0 + py, ix = ix0 + px;
      const int c = cb + piece * EPP;
      const bool ok = c < C && iy >= 0 && iy < H && ix >= 0 && ix < W;
      const T* src = ok ? in + (((int64_t)b * H + iy) * W + ix) * C + c : in;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + (size_t)pix * 64 + piece * EPP);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16u : 0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int quad = tid & 15, col = tid >> 4;          // 4 channels, one output column
  const int c0 = cb + quad * 4;
  const bool cvalid = c0 < C;
  float wk[9][4], sc[4], bi[4], ssum[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { ssum[j] = 0.f; sc[j] = 0.f; bi[j] = 0.f; }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    if (cvalid) load4(w + t * C + c0, wk[t]);
    else {
#pragma unroll
      for (int j = 0; j < 4; ++j) wk[t][j] = 0.f;
    }
  }
  if (cvalid) { load4(scale + c0, sc); load4(bias + c0, bi); }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int ox = ox0 + col;
  if (cvalid && ox < Wo) {
    const T* tcol = tile + (size_t)(col * STRIDE) * 64 + quad * 4;   // window column 0 of this thread, tile row 0
    float r[3][3][4];                                                // [window row][window col][channel]
    auto load_row = [&](int slot, int trow) {
      const T* rp = tcol + (size_t)trow * IW * 64;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) load4(rp + kx * 64, r[slot][kx]);
    };
    if (STRIDE == 1) { load_row(0, 0); load_row(1, 1); }
    for (int py = 0; py < TH; ++py) {
      const int oy = oy0 + py;
      if (oy >= Ho) break;
      if (STRIDE == 1) load_row(2, py + 2);
      else { if (py == 0) load_row(0, 0); load_row(1, 2 * py + 1); load_row(2, 2 * py + 2); }
      float acc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(r[ky][kx][j], wk[ky * 3 + kx][j], acc[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float t = fmaf(acc[j], sc[j], bi[j]);
        acc[j] = sizeof(T) == 4 ? silu_precise(t) : silu_tanh_f(t);
        ssum[j] += acc[j];
      }
      store4(out + (((int64_t)b * Ho + oy) * Wo + ox) * C + c0, acc);
      // slide the window down
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (STRIDE == 1) { r[0][kx][j] = r[1][kx][j]; r[1][kx][j] = r[2][kx][j]; }
          else r[0][kx][j] = r[2][kx][j];
        }
    }
  }
  if (se_sum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) red[col * 64 + quad * 4 + j] = ssum[j];
  __syncthreads();
  if (tid < 64) {
    float t = 0.f;
    for (int q = 0; q < TW; ++q) t += red[q * 64 + tid];
    const int c = cb + tid;
    // per-tile partial (no atomics: the squeeze must not depend on the CTA schedule); se_fc1 adds the tiles in order
    if (c < C) se_sum[((int64_t)b * gridDim.x + blockIdx.x) * C + c] = t;
  }
}

static void dwconv3x3_tiling(int Ho, int Wo, int stride, size_t es, int* TWo, int* THo, size_t* smemo) {
  // one thread per (output column, channel quad): TW columns x 16 quads = block size; tall tiles amortise the window
  int TW = (Wo % 16 == 0) ? 16 : 8;
  int TH = (Ho % 24 == 0) ? 24 : ((Ho % 16 == 0) ? 16 : 8);
  size_t smem;
  for (;;) {
    const int IH = (TH - 1) * stride + 3, IW = (TW - 1) * stride + 3;
    smem = align_up((size_t)IH * IW * 64 * es, 16) + (size_t)TW * 64 * sizeof(float);
    if (smem <= 72 * 1024 || TH <= 4) break;                   // keep >= 3 CTAs per SM
    TH /= 2;
  }
  *TWo = TW; *THo = TH; *smemo = smem;
}

int dwconv3x3_tiles(int H, int W, int stride, int dtype) {
  int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1, TW, TH;
  size_t smem;
  dwconv3x3_tiling(Ho, Wo, stride, dtype == DT_F32 ? 4 : 2, &TW, &TH, &smem);
  return ceil_div(Wo, TW) * ceil_div(Ho, TH);
}

int dwconv3x3(const void* in, void* out, int dtype, int B, int H, int W, int C, int stride, const float* w,
              const float* scale, const float* bias, float* se_sum, cudaStream_t s) {
  FTC_REQUIRE(C % 8 == 0, "depthwise channels must be a multiple of 8");
  FTC_REQUIRE(stride == 1 || stride == 2, "depthwise stride must be 1 or 2");
  int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;   // k=3, pad=1
  const size_t es = dtype == DT_F32 ? 4 : 2;
  int TW, TH;
  size_t smem;
  dwconv3x3_tiling(Ho, Wo, stride, es, &TW, &TH, &smem);
  const int tiles_x = ceil_div(Wo, TW), tiles_y = ceil_div(Ho, TH);
  dim3 grid(tiles_x * tiles_y, ceil_div(C, 64), B);
  const int threads = TW * 16;
#define DW_LAUNCH(TT, SS)                                                                                             \
  do {                                                                                                                \
    static bool done = false;                                                                                         \
    if (!done) { FTC_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_kernel<TT, SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); done = true; } \
    dwconv3x3_kernel<TT, SS><<<grid, threads, smem, s>>>((const TT*)in, (TT*)out, H, W, C, Ho, Wo, TH, TW, tiles_x, w, scale, bias, se_sum); \
  } while (0)
  if (dtype == DT_F32) { if (stride == 1) DW_LAUNCH(float, 1); else DW_LAUNCH(float, 2); }
  else { if (stride == 1) DW_LAUNCH(bf16, 1); else DW_LAUNCH(bf16, 2); }
#undef DW_LAUNCH
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Stride-1 depthwise 3x3 + BN + SiLU with the SE squeeze AND the first SE layer folded in (the MBConv blocks after the
// first of each stage: 78 of the 80 depthwise convs of EfficientNetV2-XL).
// One CTA = one image x 32 channels, walking the image in strips of 8 output rows through a two-deep cp.async ring, so
// loads of strip s+1 overlap the FMAs of strip s inside the CTA.  Each thread owns one output column and 4 channels.
// Because a CTA sees ALL pixels of its channels, the squeeze needs no atomics: mean[b, c] is final inside the CTA, and by
// linearity the CTA adds its 32-channel share of fc1 -- sum_c w1[s, c] * mean[b, c] -- straight into hid_pre[b, s]
// (S atomics per CTA).  What is left of SE is one tiny kernel (se_fc2_hid) instead of two, and the se_sum round trip.
__device__ __forceinline__ void load4p(const float* p, f32x2 (&v)[2]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = pk2(a.x, a.y); v[1] = pk2(a.z, a.w);
}
__device__ __forceinline__ void load4p(const bf16* p, f32x2 (&v)[2]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = pk2(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u));
  v[1] = pk2(__uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
}

template <typename T, int TH, bool RAW>
__global__ void __launch_bounds__(384, 2) dwconv3x3_strip_kernel(const __grid_constant__ CUtensorMap tmIn, T* __restrict__ out, int H, int W, int C,
                                                              const float* __restrict__ w, const float* __restrict__ scale,
                                                              const float* __restrict__ bias, const float* __restrict__ w1,
                                                              int S, float inv_hw, float* __restrict__ hid_pre, int flip) {
  extern __shared__ __align__(16) unsigned char dw_smem_raw[];
  // TMA destinations need 128-byte alignment; the dynamic shared window only guarantees 16
  unsigned char* dw_smem = dw_smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(dw_smem_raw) & 127u)) & 127u);
  constexpr int CB = 32;                                // channels per CTA
  constexpr int PPP = (int)(CB * sizeof(T) / 16);       // 16-byte pieces per pixel
  constexpr int EPP = (int)(16 / sizeof(T));
  const int IW = W + 2;
  const int strip_elems = (TH + 2) * IW * CB;
  T* ring = reinterpret_cast<T*>(dw_smem);              // [2][(TH+2)][IW][CB]
  float* red = reinterpret_cast<float*>(dw_smem + (size_t)2 * strip_elems * sizeof(T));   // [W][CB] then mean[CB]
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int b = blockIdx.y, cb = blockIdx.x * CB;
  T* oimg = out + (int64_t)b * H * W * C + cb;
  const int nstrips = H / TH;
  // strip loader: ONE TMA tensor box per strip -- (TH+2) rows x (W+2) pixels x 32 channels of image b, issued by one thread;
  // out-of-bounds zero fill is the conv padding.  (The cp.async version spent ~20 % of the kernel's issue slots on its
  // address arithmetic.)
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(red + (size_t)(W + 1) * CB);   // two mbarriers behind red / mean
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmIn)) : "memory");
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t strip_bytes = (uint32_t)(strip_elems * sizeof(T));
  auto prefetch = [&](int s) {
    if (tid == 0) {
      const uint32_t bar = bar0 + 8u * (uint32_t)(s & 1);
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(ring + (size_t)(s & 1) * strip_elems);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(strip_bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tmIn)), "r"(cb), "r"(-1), "r"(s * TH - 1), "r"(b), "r"(bar) : "memory");
    }
  };
  auto wait_strip = [&](int s) {
    const uint32_t bar = bar0 + 8u * (uint32_t)(s & 1), parity = (uint32_t)(s >> 1) & 1u;
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
  };
  prefetch(0);
  const int quad = tid & 7, col = tid >> 3;            // 4 channels, one output column
  const int c0 = cb + quad * 4;
  // raw = plain depthwise convolution (train step: BatchNorm needs the batch statistics of the raw output first; the data
  // gradient of a stride-1 depthwise conv is the same kernel with the taps reversed): no BN / SiLU / SE
  // RAW is a template parameter: as a run-time flag it cost the inference instantiation 300-450 bytes of register spills
  // (7.4 -> 14.4 ms per B = 32 forward, measured)
  constexpr bool raw = RAW;
  f32x2 wk[9][2], sc[2], bi[2], ssum[2];
#pragma unroll
  for (int t = 0; t < 9; ++t) load4p(w + ((RAW && flip) ? 8 - t : t) * C + c0, wk[t]);      // flip: taps rotated 180 degrees (data gradient)
  if (!raw) { load4p(scale + c0, sc); load4p(bias + c0, bi); }
  else { sc[0] = sc[1] = pk2(1.f, 1.f); bi[0] = bi[1] = pk2(0.f, 0.f); }
  ssum[0] = ssum[1] = pk2(0.f, 0.f);
  const f32x2 half2 = pk2(0.5f, 0.5f);
  for (int s = 0; s < nstrips; ++s) {
    if (s + 1 < nstrips) prefetch(s + 1);
    wait_strip(s);
    const T* tcol = ring + (size_t)(s & 1) * strip_elems + (size_t)col * CB + quad * 4;
    f32x2 r[3][3][2];
    auto load_row = [&](int slot, int trow) {
      const T* rp = tcol + (size_t)trow * IW * CB;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) load4p(rp + kx * CB, r[slot][kx]);
    };
    load_row(0, 0); load_row(1, 1);
    T* op = oimg + ((int64_t)(s * TH) * W + col) * C + quad * 4;
#pragma unroll
    for (int py = 0; py < TH; ++py) {
      load_row((py + 2) % 3, py + 2);
      float o[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // same operation order as the scalar kernel: acc = fma(x, w, acc) over (ky, kx), then BN, then SiLU
        f32x2 acc = fmul2(r[py % 3][0][h], wk[0][h]);
#pragma unroll
        for (int t = 1; t < 9; ++t) acc = ffma2(r[(py + t / 3) % 3][t % 3][h], wk[t][h], acc);
        const f32x2 t2 = ffma2(acc, sc[h], bi[h]);
        float x0, x1;
        if (raw) {
          upk2(acc, x0, x1);
        } else if (sizeof(T) == 4) {
          upk2(t2, x0, x1);
          x0 = silu_precise(x0); x1 = silu_precise(x1);
        } else {
          const f32x2 hh = fmul2(t2, half2);
          float h0, h1, t0, t1;
          upk2(hh, h0, h1);
          asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
          asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
          upk2(ffma2(hh, pk2(t0, t1), hh), x0, x1);
        }
        ssum[h] = fadd2(ssum[h], pk2(x0, x1));
        o[2 * h] = x0; o[2 * h + 1] = x1;
      }
      store4(op + (int64_t)py * W * C, o);
    }
    __syncthreads();                                    // the strip buffer is refilled two iterations later
  }
  if (raw) return;
  // ---- squeeze (complete for these 32 channels) + this CTA's share of fc1 ----
  {
    float s0, s1, s2, s3;
    upk2(ssum[0], s0, s1); upk2(ssum[1], s2, s3);
    float* rp = red + col * CB + quad * 4;
    rp[0] = s0; rp[1] = s1; rp[2] = s2; rp[3] = s3;
  }
  __syncthreads();
  float* mean = red + (size_t)W * CB;
  if (tid < CB) {
    float t = 0.f;
    for (int q = 0; q < W; ++q) t += red[q * CB + tid];
    mean[tid] = t * inv_hw;
  }
  __syncthreads();
  for (int sidx = tid; sidx < S; sidx += nthr) {
    const float* wr = w1 + (int64_t)sidx * C + cb;
    float t = 0.f;
#pragma unroll
    for (int c = 0; c < CB; c += 4) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + c));
      t = fmaf(wv.x, mean[c], t); t = fmaf(wv.y, mean[c + 1], t); t = fmaf(wv.z, mean[c + 2], t); t = fmaf(wv.w, mean[c + 3], t);
    }
    hid_pre[((int64_t)b * gridDim.x + blockIdx.x) * S + sidx] = t;   // this CTA's 32-channel share (summed in order by se_fc2_hid)
  }
}

bool dwconv3x3_se_supported(int H, int W, int C, int stride) {
  return stride == 1 && W * 8 <= 384 && W >= 4 && W % 2 == 0 && H % 8 == 0 && C % 32 == 0;
}

int dwconv3x3_se(const void* in, void* out, int dtype, int B, int H, int W, int C, const float* w, const float* scale,
                 const float* bias, const float* w1, int S, float* hid_pre, cudaStream_t s, int flip) {
  FTC_REQUIRE(dwconv3x3_se_supported(H, W, C, 1), "dwconv3x3_se: unsupported geometry");
  FTC_REQUIRE(B <= 65535, "batch");
  constexpr int TH = 8;
  const size_t es = dtype == DT_F32 ? 4 : 2;
  const size_t smem = 2 * (size_t)(TH + 2) * (W + 2) * 32 * es + (size_t)(W + 1) * 32 * sizeof(float) + 16 + 128;   // + mbarriers + alignment slack
  FTC_REQUIRE(smem <= 200 * 1024, "dwconv3x3_se: strip does not fit shared memory");
  dim3 grid(C / 32, B);
  const int threads = W * 8;
  const float inv_hw = 1.0f / (float)(H * W);
  const bool raw_mode = scale == nullptr;
  alignas(64) CUtensorMap tmIn;
  {
    int rc = tma_encode_nhwc(&tmIn, in, dtype, C, C, W, H, B, 32, W + 2, TH + 2, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
#define DWS_LAUNCH(TT)                                                                                                 \
  do {                                                                                                                 \
    static bool done = false;                                                                                          \
    if (!done) {                                                                                                       \
      FTC_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_strip_kernel<TT, TH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      FTC_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_strip_kernel<TT, TH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      done = true;                                                                                                     \
    }                                                                                                                  \
    if (raw_mode) FTC_CHECK_CUDA(launch_pdl(dwconv3x3_strip_kernel<TT, TH, true>, grid, dim3(threads), smem, s, tmIn, (TT*)out, H, W, C, w, scale, bias, w1, S, inv_hw, hid_pre, flip)); \
    else FTC_CHECK_CUDA(launch_pdl(dwconv3x3_strip_kernel<TT, TH, false>, grid, dim3(threads), smem, s, tmIn, (TT*)out, H, W, C, w, scale, bias, w1, S, inv_hw, hid_pre, flip)); \
  } while (0)
  // (an mma.sync variant of this kernel was measured slower on B200 -- 0.26 vs 0.17 ms at 48x48x1536, B = 32: its BN / SiLU / SE
  // epilogue and fragment exchange cost as many issue slots as the FMAs they replace -- and was removed)
  static int env_th12 = -1;
  if (env_th12 < 0) { const char* e = getenv("FTC_DW_TH12"); env_th12 = e ? atoi(e) : 1; }   // default on (measured 6.6 vs 7.1 ms)
  if (dtype == DT_BF16 && env_th12 && H % 12 == 0) {
    // 12-row strips: 14/12 instead of 10/8 rows loaded per output row, fewer strip turn-arounds; 90 KB of shared memory
    constexpr int TH12 = 12;
    const size_t smem12 = 2 * (size_t)(TH12 + 2) * (W + 2) * 32 * es + (size_t)(W + 1) * 32 * sizeof(float) + 16 + 128;
    alignas(64) CUtensorMap tm12;
    int rc = tma_encode_nhwc(&tm12, in, dtype, C, C, W, H, B, 32, W + 2, TH12 + 2, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    static bool done12 = false;
    if (!done12) {
      FTC_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_strip_kernel<bf16, TH12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      FTC_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_strip_kernel<bf16, TH12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      done12 = true;
    }
    if (raw_mode) FTC_CHECK_CUDA(launch_pdl(dwconv3x3_strip_kernel<bf16, TH12, true>, grid, dim3(threads), smem12, s, tm12, (bf16*)out, H, W, C, w, scale, bias, w1, S, inv_hw, hid_pre, flip));
    else FTC_CHECK_CUDA(launch_pdl(dwconv3x3_strip_kernel<bf16, TH12, false>, grid, dim3(threads), smem12, s, tm12, (bf16*)out, H, W, C, w, scale, bias, w1, S, inv_hw, hid_pre, flip));
  }
  else if (dtype == DT_F32) DWS_LAUNCH(float);
  else DWS_LAUNCH(bf16);
#undef DWS_LAUNCH
  FTC_POST_LAUNCH();
  return 0;
}

// plain stride-1 depthwise 3x3 through the strip kernel (train step forward / data gradient)
int dwconv3x3_raw_strip(const void* in, void* out, int dtype, int B, int H, int W, int C, const float* w, int flip, cudaStream_t s) {
  FTC_REQUIRE(dwconv3x3_se_supported(H, W, C, 1), "dwconv3x3_raw_strip: unsupported geometry");
  return dwconv3x3_se(in, out, dtype, B, H, W, C, w, nullptr, nullptr, nullptr, 0, nullptr, s, flip);
}

// second half of SE for the folded path: scale[b,c] = sigmoid(b2[c] + w2t[:,c] . silu(sum_g hid_part[b,g,:] + b1)).
// hid_part [B][G][S] holds one fc1 share per depthwise CTA (G = C / 32 channel groups); the shares are added in a FIXED
// order (P interleaved partial sums over g, combined in order), so the excitation is bit-reproducible run to run and
// independent of the batch size -- near-tied peak scores downstream must not depend on the CTA schedule.
__global__ void __launch_bounds__(256) se_fc2_hid_kernel(const float* __restrict__ hid_part, int G,
                                                         float* __restrict__ scale_out, int C, int S,
                                                         const float* __restrict__ b1, const float* __restrict__ w2t,
                                                         const float* __restrict__ b2) {
  __shared__ float sh[256];
  __shared__ float part[256];
  const int b = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  const int P = 256 / S > 0 ? 256 / S : 1;              // partial sums per squeeze unit (S <= 256)
  {
    const int k = threadIdx.x % S, pi = threadIdx.x / S;
    if (pi < P) {
      const float* hp = hid_part + (int64_t)b * G * S + k;
      float t = 0.f;
#pragma unroll 4
      for (int g = pi; g < G; g += P) t += hp[(int64_t)g * S];
      part[pi * S + k] = t;
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < S; k += blockDim.x) {
    float t = part[k];
    for (int pi = 1; pi < P; ++pi) t += part[pi * S + k];
    sh[k] = silu_precise(t + b1[k]);
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = b2[c];
#pragma unroll 8
  for (int k = 0; k < S; ++k) t = fmaf(__ldg(w2t + (int64_t)k * C + c), sh[k], t);   // 8 loads in flight (same summation order)
  scale_out[(int64_t)b * C + c] = sigmoid_precise(t);
}

int se_fc2_hid(const float* hid_part, int G, float* scale_out, int B, int C, int S, const float* b1, const float* w2t,
               const float* b2, cudaStream_t s) {
  FTC_REQUIRE(S <= 256 && S > 0 && G > 0, "SE: squeeze <= 256");
  FTC_CHECK_CUDA(launch_pdl(se_fc2_hid_kernel, dim3(ceil_div(C, 256), B), dim3(256), 0, s, hid_part, G, scale_out, C, S, b1, w2t, b2));
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SE excitation in two small grid-filling kernels (one block per image was latency-bound: 0.22 ms per layer):
//   fc1: one warp per (image, squeeze unit)   hid[b,s] = silu(b1[s] + w1[s,:] . mean[b,:])
//   fc2: one thread per (image, channel)      scale[b,c] = sigmoid(b2[c] + w2t[:,c] . hid[b,:])
// sum: [B][nt][C] per-tile partial sums of the depthwise kernel, added tile by tile in order (deterministic)
__global__ void __launch_bounds__(256) se_fc1_kernel(const float* __restrict__ sum, int nt, float* __restrict__ hid, int C, int S,
                                                     float inv_hw, const float* __restrict__ w1,
                                                     const float* __restrict__ b1) {
  const int b = blockIdx.y;
  const int sidx = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (sidx >= S) return;
  const float* wr = w1 + (int64_t)sidx * C;
  const float* sp = sum + (int64_t)b * nt * C;
  float t = 0.f;
  for (int c = lane * 4; c < C; c += 128) {
    float4 w = *reinterpret_cast<const float4*>(wr + c);
    float4 m = *reinterpret_cast<const float4*>(sp + c);
    for (int q = 1; q < nt; ++q) {
      const float4 m2 = *reinterpret_cast<const float4*>(sp + (int64_t)q * C + c);
      m.x += m2.x; m.y += m2.y; m.z += m2.z; m.w += m2.w;
    }
    t = fmaf(w.x, m.x * inv_hw, t); t = fmaf(w.y, m.y * inv_hw, t);
    t = fmaf(w.z, m.z * inv_hw, t); t = fmaf(w.w, m.w * inv_hw, t);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (lane == 0) hid[(int64_t)b * S + sidx] = silu_precise(t + b1[sidx]);
}

__global__ void __launch_bounds__(256) se_fc2_kernel(const float* __restrict__ hid,
                                                     float* __restrict__ scale_out, int C, int S,
                                                     const float* __restrict__ w2t, const float* __restrict__ b2) {
  __shared__ float sh[256];
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < S; k += blockDim.x) sh[k] = hid[(int64_t)b * S + k];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = b2[c];
  for (int k = 0; k < S; ++k) t = fmaf(w2t[(int64_t)k * C + c], sh[k], t);
  scale_out[(int64_t)b * C + c] = sigmoid_precise(t);
}

int se_fc(const float* sum, int nt, float* scale_out, float* hid, int B, int C, int S, float inv_hw, const float* w1, const float* b1,
          const float* w2t, const float* b2, cudaStream_t s) {
  FTC_REQUIRE(S <= 256 && C % 4 == 0 && nt >= 1, "SE: squeeze <= 256 and channels % 4 == 0");
  se_fc1_kernel<<<dim3(ceil_div(S, 8), B), 256, 0, s>>>(sum, nt, hid, C, S, inv_hw, w1, b1);
  FTC_POST_LAUNCH();
  se_fc2_kernel<<<dim3(ceil_div(C, 256), B), 256, 0, s>>>(hid, scale_out, C, S, w2t, b2);
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// bilinear x2 align_corners=True
// One CTA per output row (b, oy): the row pair y0/y1 and its weights are CTA constants, warps stride over output pixels
// and lanes over 16-byte channel chunks -- no integer division anywhere, every load / store is a 512-byte-per-warp run.
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C,
                                                         float sy, float sx) {
  const int Ho = 2 * H, Wo = 2 * W, CH = C / 8;
  const int b = blockIdx.y, oy = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float fy = sy * oy;
  const int y0 = min((int)fy, H - 1), y1 = min(y0 + 1, H - 1);
  const float ly = fminf(fmaxf(fy - y0, 0.f), 1.f), hy = 1.f - ly;
  const T* row0 = in + ((int64_t)b * H + y0) * W * C;
  const T* row1 = in + ((int64_t)b * H + y1) * W * C;
  T* orow = out + ((int64_t)b * Ho + oy) * Wo * C;
  for (int ox = warp; ox < Wo; ox += 8) {
    const float fx = sx * ox;
    const int x0 = min((int)fx, W - 1), x1 = min(x0 + 1, W - 1);
    const float lx = fminf(fmaxf(fx - x0, 0.f), 1.f), hx = 1.f - lx;
    const T* p00 = row0 + (int64_t)x0 * C;
    const T* p01 = row0 + (int64_t)x1 * C;
    const T* p10 = row1 + (int64_t)x0 * C;
    const T* p11 = row1 + (int64_t)x1 * C;
    T* po = orow + (int64_t)ox * C;
#pragma unroll 2
    for (int ch = lane; ch < CH; ch += 32) {
      float a[8], bb[8], c[8], d[8], o[8];
      load8(p00 + ch * 8, a);
      load8(p01 + ch * 8, bb);
      load8(p10 + ch * 8, c);
      load8(p11 + ch * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = hy * (hx * a[j] + lx * bb[j]) + ly * (hx * c[j] + lx * d[j]);
      store8(po + ch * 8, o);
    }
  }
}

int upsample2x(const void* in, void* out, int dtype, int B, int H, int W, int C, cudaStream_t s) {
  FTC_REQUIRE(C % 8 == 0, "upsample channels must be a multiple of 8");
  FTC_REQUIRE(B <= 65535, "upsample batch");
  dim3 grid(2 * H, B);
  float sy = (float)(H - 1) / (float)(2 * H - 1), sx = (float)(W - 1) / (float)(2 * W - 1);
  if (dtype == DT_F32)
    upsample2x_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (float*)out, H, W, C, sy, sx);
  else
    upsample2x_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)in, (bf16*)out, H, W, C, sy, sx);
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Leafmap top_conv of the eight 1-/2-channel heads (models/detector.py:188-190, 3x3, 192 -> 1|2, +bias).
// As a GEMM this is N <= 2 per head with a private K = 1728 operand per head: pure operand bandwidth (it ran at
// 5 TFLOP/s on the tensor path).  Here each CTA stages one head's 192-channel halo tile in shared memory once and
// every lane owns a fixed set of (tap, 8-channel chunk) pairs whose weights live in registers.
struct HeadTopParams {
  const void* y; int pix_stride;          // NHWC source, channels of head h at [h*192, h*192+192)
  const float* w;                         // [rows][9*192] fp32, row = output channel (tap-major k)
  const float* bias;                      // [rows]
  float* out; int out_ch;                 // NCHW fp32 [B, out_ch, H, W]
  int H, W, tiles_x, n_heads;
  int row_base[8];                        // first output row (= channel) of head h
  int od[8];                              // 1 or 2
  int head0;                              // first head index in the source buffer
};

template <typename T>
__global__ void __launch_bounds__(256) head_top_conv_kernel(const HeadTopParams p) {
  extern __shared__ __align__(16) unsigned char ht_smem[];
  constexpr int CD = 192, CH = CD / 8, TH = 8, TW = 16, IH = TH + 2, IW = TW + 2;
  T* tile = reinterpret_cast<T*>(ht_smem);                                    // [IH*IW][192]
  float* outs = reinterpret_cast<float*>(ht_smem + (size_t)IH * IW * CD * sizeof(T));   // [2][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
  const int oy0 = ty * TH, ox0 = tx * TW;
  const T* src = reinterpret_cast<const T*>(p.y) + (size_t)(p.head0 + h) * CD;
  constexpr int PIECES = (int)(8 * sizeof(T) / 16);
  for (int i = tid; i < IH * IW * CH; i += 256) {
    const int pix = i / CH, ch = i - pix * CH;
    const int py = pix / IW, px = pix - py * IW;
    const int iy = oy0 - 1 + py, ix = ox0 - 1 + px;
    const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
    const T* g = ok ? src + (((int64_t)b * p.H + iy) * p.W + ix) * p.pix_stride + ch * 8 : src;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + (size_t)pix * CD + ch * 8);
#pragma unroll
    for (int q = 0; q < PIECES; ++q)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + 16u * q),
                   "l"(reinterpret_cast<const char*>(g) + 16 * q), "r"(ok ? 16u : 0u) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const int od = p.od[h];
  for (int o = 0; o < od; ++o) {
    // lane owns pairs q = lane + 32*it of the 9*24 = 216 (tap, chunk) pairs
    float wr[7][8];
    int toff[7];
    const float* wrow = p.w + (size_t)(p.row_base[h] + o) * (9 * CD);
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int q = lane + 32 * it;
      const bool ok = q < 9 * CH;
      const int tap = ok ? q / CH : 0, ch = ok ? q - tap * CH : 0;
      toff[it] = ((tap / 3) * IW + (tap % 3)) * CD + ch * 8;
      if (ok) load8(wrow + tap * CD + ch * 8, wr[it]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) wr[it][j] = 0.f;
      }
    }
    for (int pi = 0; pi < 16; ++pi) {
      const int pix = warp * 16 + pi;                 // 8 warps x 16 = 128 tile pixels
      const int py = pix / TW, px = pix - py * TW;
      const T* base = tile + (size_t)(py * IW + px) * CD;
      float acc = 0.f;
#pragma unroll
      for (int it = 0; it < 7; ++it) {
        float v[8];
        load8(base + toff[it], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(v[j], wr[it][j], acc);
      }
#pragma unroll
      for (int s2 = 16; s2 > 0; s2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s2);
      if (lane == 0) outs[o * 128 + pix] = acc + p.bias[p.row_base[h] + o];
    }
  }
  __syncthreads();
  for (int i = tid; i < od * 128; i += 256) {
    const int o = i >> 7, pix = i & 127;
    const int oy = oy0 + pix / TW, ox = ox0 + pix % TW;
    if (oy < p.H && ox < p.W)
      p.out[(((int64_t)b * p.out_ch + p.row_base[h] + o) * p.H + oy) * p.W + ox] = outs[i];
  }
}

// bf16 variant on the warp-level tensor cores (mma.sync m16n8k16, fp32 accumulate).  The op is pure operand bandwidth
// (N <= 2 per head), so the design minimises shared-memory reads per MAC: a CTA stages one head's 10x18x192 halo tile
// once (pixel stride padded to 200 channels: conflict-free ldmatrix), each of its 6 warps owns 32 channels and ALL 8x16
// output pixels, loads every (input row, column shift, 16-channel step) fragment once with ldmatrix.x4 and feeds it to
// the up-to-three output rows it contributes to (ky = 0..2).  The 18 weight fragments of a warp live in registers.
// Shared-memory reads: 2.7x the tile instead of 9x; the CUDA-core version above was FMA/LDS bound at 11 TFLOP/s.
__global__ void __launch_bounds__(192) head_top_mma_kernel(const HeadTopParams p, const __grid_constant__ CUtensorMap tmY) {
  extern __shared__ __align__(16) unsigned char ht_smem_raw[];
  constexpr int CD = 192, TH = 8, TW = 16, IH = TH + 2, IW = TW + 2, NW = 6;
  constexpr uint32_t SLAB = 23552;                                 // one 64-channel slab: 180 pixels x 128 B, padded to 1 KB
  // The halo tile arrives as THREE TMA boxes (64 channels x 18 px x 10 rows, zero-filled outside the image) in the 128B-swizzled
  // layout: pixel p of slab s is the 128-byte row p, its 16-byte chunk c sits at chunk (c ^ (p & 7)) -- conflict-free for ldmatrix
  // without padding, and no per-thread copy loop (the cp.async version spent more instructions loading than computing).
  unsigned char* ht_smem = ht_smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(ht_smem_raw) & 1023u)) & 1023u);
  float* part = reinterpret_cast<float*>(ht_smem);                 // [NW][128][2], aliases the tile after the MMAs
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int ty = blockIdx.x / p.tiles_x, tx = blockIdx.x - ty * p.tiles_x;
  const int oy0 = ty * TH, ox0 = tx * TW;
  const uint32_t tile_s0 = (uint32_t)__cvta_generic_to_shared(ht_smem);
  const uint32_t bar = tile_s0 + 3 * SLAB;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(3u * (uint32_t)(IH * IW * 128)) : "memory");
#pragma unroll
    for (int sl = 0; sl < 3; ++sl)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(tile_s0 + (uint32_t)sl * SLAB), "l"(reinterpret_cast<uint64_t>(&tmY)), "r"((p.head0 + h) * CD + sl * 64),
                     "r"(ox0 - 1), "r"(oy0 - 1), "r"(b), "r"(bar) : "memory");
  }
  // B fragments (k16 x n8, "col" layout): lane holds k = (lane%4)*2 + {0,1} (+8), n = lane/4; columns >= od are zero
  const int od = p.od[h];
  uint32_t wf[9][2][2];
  {
    const int n = lane >> 2, kk = (lane & 3) * 2;
    const float* wrow = p.w + (size_t)(p.row_base[h] + (n < od ? n : 0)) * (9 * CD) + warp * 32 + kk;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
        if (n < od) {
          lo = __ldg(reinterpret_cast<const float2*>(wrow + tap * CD + ks * 16));
          hi = __ldg(reinterpret_cast<const float2*>(wrow + tap * CD + ks * 16 + 8));
        }
        __nv_bfloat162 l2 = __floats2bfloat162_rn(lo.x, lo.y), h2 = __floats2bfloat162_rn(hi.x, hi.y);
        wf[tap][ks][0] = *reinterpret_cast<uint32_t*>(&l2);
        wf[tap][ks][1] = *reinterpret_cast<uint32_t*>(&h2);
      }
  }
  float acc[TH][4];
#pragma unroll
  for (int r = 0; r < TH; ++r)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
  __syncthreads();                                   // barrier initialised before anybody polls it
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                   : "=r"(ok) : "r"(bar) : "memory");
  }
  // ldmatrix.x4 row addresses: matrix (lane/8): pixels (lane%8) + 8*((lane/8)&1), channels + 8*((lane/8)>>1).
  // This warp's 32 channels are half of slab warp/2: 16-byte chunks (warp&1)*4 + ks*2 + (lane>>4) of the pixel's 128-byte row.
  const int lpix = (lane & 7) + 8 * ((lane >> 3) & 1);
  const uint32_t slab_s = tile_s0 + (uint32_t)(warp >> 1) * SLAB;
  const int cbase = (warp & 1) * 4 + (lane >> 4);
#pragma unroll
  for (int iy = 0; iy < IH; ++iy)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a0, a1, a2, a3;
        const int pix = iy * IW + kx + lpix;
        const uint32_t addr = slab_s + (uint32_t)pix * 128u + (uint32_t)(((cbase + ks * 2) ^ (pix & 7)) << 4);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr) : "memory");
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int oy = iy - ky;
          if (oy >= 0 && oy < TH)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                         : "+f"(acc[oy][0]), "+f"(acc[oy][1]), "+f"(acc[oy][2]), "+f"(acc[oy][3])
                         : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(wf[ky * 3 + kx][ks][0]), "r"(wf[ky * 3 + kx][ks][1]));
        }
      }
  __syncthreads();                                   // every warp is done with the tile: reuse it for the partial sums
  if ((lane & 3) == 0) {
    const int g = lane >> 2;
#pragma unroll
    for (int r = 0; r < TH; ++r) {
      float* pp = part + ((size_t)warp * 128 + r * TW + g) * 2;
      pp[0] = acc[r][0]; pp[1] = acc[r][1];
      pp[16] = acc[r][2]; pp[17] = acc[r][3];          // pixel g + 8
    }
  }
  __syncthreads();
  for (int i = tid; i < od * 128; i += NW * 32) {
    const int o = i >> 7, pix = i & 127;
    float t = p.bias[p.row_base[h] + o];
#pragma unroll
    for (int w = 0; w < NW; ++w) t += part[((size_t)w * 128 + pix) * 2 + o];
    const int oy = oy0 + pix / TW, ox = ox0 + pix % TW;
    if (oy < p.H && ox < p.W)
      p.out[(((int64_t)b * p.out_ch + p.row_base[h] + o) * p.H + oy) * p.W + ox] = t;
  }
}

int head_top_conv(const void* y, int dtype, int pix_stride, int head0, int n_heads, const int* od, const float* w,
                  const float* bias, float* out, int out_ch, int B, int H, int W, cudaStream_t s) {
  FTC_REQUIRE(n_heads >= 1 && n_heads <= 8, "head_top_conv handles up to 8 heads");
  HeadTopParams p;
  p.y = y; p.pix_stride = pix_stride; p.w = w; p.bias = bias; p.out = out; p.out_ch = out_ch; p.H = H; p.W = W;
  p.tiles_x = ceil_div(W, 16); p.n_heads = n_heads; p.head0 = head0;
  int row = 0;
  for (int i = 0; i < 8; ++i) {
    p.row_base[i] = row; p.od[i] = i < n_heads ? od[i] : 0;
    FTC_REQUIRE(p.od[i] <= 2, "head_top_conv: out_dim <= 2");
    row += p.od[i];
  }
  dim3 grid(p.tiles_x * ceil_div(H, 8), n_heads, B);
  const size_t es = dtype == DT_F32 ? 4 : 2;
  size_t smem = (size_t)10 * 18 * 192 * es + 2 * 128 * sizeof(float);
  static bool attr_done[2] = {false, false};
  if (dtype == DT_F32) {
    if (!attr_done[0]) { FTC_CHECK_CUDA(cudaFuncSetAttribute(head_top_conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_done[0] = true; }
    head_top_conv_kernel<float><<<grid, 256, smem, s>>>(p);
  } else {
    FTC_REQUIRE(pix_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "head_top_conv: bf16 source alignment");
    smem = 3 * 23552 + 16 + 1024;          // three swizzled slabs + mbarrier + 1 KB alignment slack
    alignas(64) CUtensorMap tmY;
    {
      int rc = tma_encode_nhwc(&tmY, y, DT_BF16, (uint64_t)pix_stride, (uint64_t)pix_stride, W, H, B, 64, 18, 10, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
    if (!attr_done[1]) { FTC_CHECK_CUDA(cudaFuncSetAttribute(head_top_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); attr_done[1] = true; }
    head_top_mma_kernel<<<grid, 192, smem, s>>>(p, tmY);
  }
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// peak pick
__device__ __forceinline__ float local_max3x3(const float* __restrict__ key, int H, int W, int y, int x) {
  float mx = -INFINITY;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      int yy = y + dy, xx = x + dx;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) mx = fmaxf(mx, key[yy * W + xx]);
    }
  return mx;
}

__global__ void __launch_bounds__(256) peak_pick_kernel(const float* __restrict__ heat9, float* __restrict__ heat10,
                                                        int B, int H, int W) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int hw = H * W;
  if (idx >= (int64_t)B * hw) return;
  int b = idx / hw, r = idx - (int64_t)b * hw;
  int y = r / W, x = r - y * W;
  const float* h9 = heat9 + (int64_t)b * 9 * hw;
  float* h10 = heat10 + (int64_t)b * 10 * hw;
  float k = h9[r];
  float mx = local_max3x3(h9, H, W, y, x);
  h10[r] = k;
  h10[hw + r] = (k < mx) ? -INFINITY : k;      // torch.where(keymap < local_peak, -inf, keymap)
#pragma unroll
  for (int c = 1; c < 9; ++c) h10[(c + 1) * hw + r] = h9[c * hw + r];
}

int peak_pick(const float* heat9, float* heat10, int B, int H, int W, cudaStream_t s) {
  int64_t total = (int64_t)B * H * W;
  peak_pick_kernel<<<(int)((total + 255) / 256), 256, 0, s>>>(heat9, heat10, B, H, W);
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// peak compaction + box decode
__device__ __forceinline__ float np_sigmoid(float x) { return (tanhf(x * 0.5f) + 1.0f) * 0.5f; }   // util_func.py:14

// candidates of tile b go to cand[b][0 .. total[b]) in arbitrary order (slot = atomic counter); the buffer holds one key per map
// pixel, so nothing is ever dropped here, and peak_emit sorts by the key -- the RESULT does not depend on the slot order.
__global__ void __launch_bounds__(256) peak_collect_kernel(const float* __restrict__ heat9, int H, int W,
                                                           const int* __restrict__ tile_meta, float cut_off,
                                                           float page_w, float page_h,
                                                           int* __restrict__ total, unsigned long long* __restrict__ cand) {
  const int b = blockIdx.y;
  const int hw = H * W;
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= hw) return;
  int y = r / W, x = r - y * W;
  const int* tm = tile_meta + b * 6;
  if (x < tm[2] || x >= tm[3] || y < tm[4] || y >= tm[5]) return;          // centre-crop mask
  const float* h9 = heat9 + (int64_t)b * 9 * hw;
  float k = h9[r];
  if (k < local_max3x3(h9, H, W, y, x)) return;                             // not a 3x3 local maximum
  float p = np_sigmoid(k);
  if (!(p >= cut_off)) return;
  float w = expf(h9[hw + r] - 3.f) * 1024.f;
  float h = expf(h9[2 * hw + r] - 3.f) * 1024.f;
  if (w <= 0.f || h <= 0.f) return;
  if (w > page_w || h > page_h) return;
  int slot = atomicAdd(&total[b], 1);
  cand[(int64_t)b * hw + slot] =
      ((unsigned long long)__float_as_uint(p) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)r);
}

template <int NMAX>
__global__ void __launch_bounds__(1024) peak_emit_kernel(const float* __restrict__ heat9, const float* __restrict__ feat,
                                                         int H, int W, int FC, const int* __restrict__ tile_meta,
                                                         int max_peaks, int* __restrict__ count, const int* __restrict__ total,
                                                         const unsigned long long* __restrict__ cand,
                                                         float* __restrict__ loc, float* __restrict__ gfeat) {
  __shared__ unsigned long long keys[NMAX];
  __shared__ int s_cnt;
  __shared__ unsigned long long s_thr;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int hw = H * W;
  const int ntot = total[b];
  const int n = min(ntot, max_peaks);
  const unsigned long long* cb = cand + (int64_t)b * hw;
  if (ntot <= NMAX) {
    for (int i = tid; i < NMAX; i += blockDim.x) keys[i] = (i < ntot) ? cb[i] : 0ull;
  } else {
    // overflow: keep the NMAX largest keys.  Keys are distinct (the pixel index is part of the key), so the NMAX-th largest
    // key T is found bit by bit from the top -- T = max { t : #(key >= t) >= NMAX } -- and exactly NMAX keys are >= T.
    unsigned long long thr = 0ull;
    for (int bit = 63; bit >= 0; --bit) {
      const unsigned long long trial = thr | (1ull << bit);
      if (tid == 0) s_cnt = 0;
      __syncthreads();
      int c = 0;
      for (int i = tid; i < ntot; i += blockDim.x) c += (cb[i] >= trial) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
      if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
      __syncthreads();
      if (s_cnt >= NMAX) thr = trial;
      __syncthreads();
    }
    if (tid == 0) { s_cnt = 0; s_thr = thr; }
    __syncthreads();
    for (int i = tid; i < ntot; i += blockDim.x) {
      const unsigned long long kk = cb[i];
      if (kk >= s_thr) keys[atomicAdd(&s_cnt, 1)] = kk;      // slot order is irrelevant: sorted below
    }
  }
  __syncthreads();
  // bitonic sort, descending
  for (int k = 2; k <= NMAX; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < NMAX; i += blockDim.x) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], c = keys[ixj];
          bool desc = (i & k) == 0;
          if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  const float* h9 = heat9 + (int64_t)b * 9 * hw;
  const int* tm = tile_meta + b * 6;
  for (int i = tid; i < n; i += blockDim.x) {
    unsigned long long kk = keys[i];
    int r = (int)(0xFFFFFFFFu - (unsigned)(kk & 0xFFFFFFFFull));
    int y = r / W, x = r - y * W;
    float* l = loc + ((int64_t)b * max_peaks + i) * 9;
    l[0] = __uint_as_float((unsigned)(kk >> 32));
    l[1] = (float)(x * 4 + tm[0]);
    l[2] = (float)(y * 4 + tm[1]);
    l[3] = expf(h9[hw + r] - 3.f) * 1024.f;
    l[4] = expf(h9[2 * hw + r] - 3.f) * 1024.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) l[5 + c] = np_sigmoid(h9[(5 + c) * hw + r]);
  }
  // feature gather: warp per peak
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < n; i += nw) {
    int r = (int)(0xFFFFFFFFu - (unsigned)(keys[i] & 0xFFFFFFFFull));
    for (int c = lane; c < FC; c += 32)
      gfeat[((int64_t)b * max_peaks + i) * FC + c] = feat[((int64_t)b * FC + c) * hw + r];
  }
  if (tid == 0) count[b] = n;
}

size_t peak_decode_scratch_bytes(int B, int H, int W) { return (size_t)B * H * W * sizeof(unsigned long long); }

int peak_decode(const float* heat9, const float* feat, int B, int H, int W, int FC, const int* tile_meta, float cut_off,
                float page_w, float page_h, int max_peaks, int* count, int* total, float* loc, float* gfeat, void* scratch,
                cudaStream_t s) {
  FTC_REQUIRE(max_peaks == 1024 || max_peaks == 2048 || max_peaks == 4096, "max_peaks must be 1024/2048/4096");
  FTC_CHECK_CUDA(cudaMemsetAsync(total, 0, sizeof(int) * B, s));
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(scratch);   // [B][H*W]
  dim3 grid(ceil_div(H * W, 256), B);
  peak_collect_kernel<<<grid, 256, 0, s>>>(heat9, H, W, tile_meta, cut_off, page_w, page_h, total, cand);
  FTC_POST_LAUNCH();
  if (max_peaks == 1024)
    peak_emit_kernel<1024><<<B, 1024, 0, s>>>(heat9, feat, H, W, FC, tile_meta, max_peaks, count, total, cand, loc, gfeat);
  else if (max_peaks == 2048)
    peak_emit_kernel<2048><<<B, 1024, 0, s>>>(heat9, feat, H, W, FC, tile_meta, max_peaks, count, total, cand, loc, gfeat);
  else
    peak_emit_kernel<4096><<<B, 1024, 0, s>>>(heat9, feat, H, W, FC, tile_meta, max_peaks, count, total, cand, loc, gfeat);
  FTC_POST_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// weight packing
template <typename T>
__global__ void pack_conv_weight_kernel(T* __restrict__ dst, const float* __restrict__ src, int O, int Itot, int kh,
                                        int kw, int c_off, int C, int k_off, int Kpad, int o_off,
                                        const float* __restrict__ cscale) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t per_o = (int64_t)kh * kw * C;
  if (idx >= (int64_t)O * per_o) return;
  int o = idx / per_o;
  int r = idx - (int64_t)o * per_o;
  int t = r / C, c = r - t * C;
  int ky = t / kw, kx = t - ky * kw;
  float v = src[(((int64_t)o * Itot + c_off + c) * kh + ky) * kw + kx];
  if (cscale) v *= cscale[c];
  dst[(int64_t)(o_off + o) * Kpad + k_off + r] = from_f<T>(v);
}

int pack_conv_weight(void* dst, int dtype, const float* src, int O, int Itot, int kh, int kw, int c_off, int C, int k_off,
                     int Kpad, int o_off, const float* cscale, cudaStream_t s) {
  int64_t total = (int64_t)O * kh * kw * C;
  int grid = (int)((total + 255) / 256);
  if (dtype == DT_F32)
    pack_conv_weight_kernel<float><<<grid, 256, 0, s>>>((float*)dst, src, O, Itot, kh, kw, c_off, C, k_off, Kpad, o_off, cscale);
  else
    pack_conv_weight_kernel<bf16><<<grid, 256, 0, s>>>((bf16*)dst, src, O, Itot, kh, kw, c_off, C, k_off, Kpad, o_off, cscale);
  FTC_POST_LAUNCH();
  return 0;
}

__global__ void bn_fold_kernel(float* scale, float* bias, const float* gamma, const float* beta, const float* mean,
                               const float* var, float eps, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float sc = gamma[c] / sqrtf(var[c] + eps);
  scale[c] = sc;
  bias[c] = beta[c] - mean[c] * sc;
}

int bn_fold(float* scale, float* bias, const float* gamma, const float* beta, const float* mean, const float* var,
            float eps, int C, cudaStream_t s) {
  bn_fold_kernel<<<ceil_div(C, 256), 256, 0, s>>>(scale, bias, gamma, beta, mean, var, eps, C);
  FTC_POST_LAUNCH();
  return 0;
}

// one block per output channel n: T[tap] = sum_c w[n][c_off+c][tap] * shift[c]; then the 9 border cases
__global__ void __launch_bounds__(256) leaf_bias_table_kernel(float* __restrict__ tab, int ld, const float* __restrict__ w,
                                                              int Itot, int c_off, int C, const float* __restrict__ shift,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ bias) {
  __shared__ float red[9][256];
  __shared__ float T[9];
  const int n = blockIdx.x, tid = threadIdx.x;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  for (int c = tid; c < C; c += blockDim.x) {
    const float* wp = w + ((int64_t)n * Itot + c_off + c) * 9;
    float sh = shift[c];
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[t] = fmaf(wp[t], sh, acc[t]);
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) red[t][tid] = acc[t];
  __syncthreads();
  if (tid < 9) {
    float s = 0.f;
    for (int i = 0; i < 256; ++i) s += red[tid][i];
    T[tid] = s;
  }
  __syncthreads();
  if (tid < 9) {
    int rc = tid / 3, cc = tid % 3;   // 0: first row/col, 1: interior, 2: last row/col
    float s = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      if ((rc == 0 && ky == 0) || (rc == 2 && ky == 2)) continue;
      for (int kx = 0; kx < 3; ++kx) {
        if ((cc == 0 && kx == 0) || (cc == 2 && kx == 2)) continue;
        s += T[ky * 3 + kx];
      }
    }
    tab[(int64_t)tid * ld + n] = bias[n] + scale[n] * s;
  }
}

int leaf_bias_table(float* tab, int ld, const float* w, int O, int Itot, int c_off, int C, const float* shift,
                    const float* scale, const float* bias, cudaStream_t s) {
  leaf_bias_table_kernel<<<O, 256, 0, s>>>(tab, ld, w, Itot, c_off, C, shift, scale, bias);
  FTC_POST_LAUNCH();
  return 0;
}

__global__ void fill_f32_kernel(float* p, float v, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
int fill_f32(float* p, float v, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  fill_f32_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(p, v, n);
  FTC_POST_LAUNCH();
  return 0;
}

// dst[c][r] = src[r][c]  (src is [R][C])
__global__ void transpose_f32_kernel(float* dst, const float* src, int R, int C) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * C) return;
  int r = i / C, c = i - (int64_t)r * C;
  dst[(int64_t)c * R + r] = src[i];
}
int transpose_f32(float* dst, const float* src, int R, int C, cudaStream_t s) {
  int64_t n = (int64_t)R * C;
  transpose_f32_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(dst, src, R, C);
  FTC_POST_LAUNCH();
  return 0;
}

}  // namespace ftc
