// SIMT implicit-GEMM convolution (CUDA cores, fp32 accumulate).  Works for fp32 and bf16 tensors.
// This is the fp32 parity path (config #1 of BASELINE.json) and the cross-check for the tcgen05
// kernel; it shares ConvGemmParams, the weight packing and the k-chunk table with conv_gemm_tc.cu.
#include "conv_gemm.cuh"

namespace ftc {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

template <typename T>
__global__ void __launch_bounds__(NT) conv_gemm_simt_kernel(const ConvGemmParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int g = blockIdx.z;
  const T* srcA = reinterpret_cast<const T*>(p.srcA);
  const T* srcB = reinterpret_cast<const T*>(p.srcB);
  const T* wgt = reinterpret_cast<const T*>(p.w);

  // loader role: threads 0..127 load A chunks, 128..255 load W chunks
  const bool loadA = tid < 128;
  const int lrow = tid & 63;           // row (A) or n (W) inside the tile
  const int lch = (tid >> 6) & 1;      // which 8-chunk of the BK=16 block
  int lb = 0, iy0 = 0, ix0 = 0;
  bool row_ok = false;
  if (loadA) {
    int m = m0 + lrow;
    row_ok = m < p.M;
    if (row_ok) {
      int hw = p.Ho * p.Wo;
      lb = m / hw;
      int r = m - lb * hw;
      int oy = r / p.Wo, ox = r - oy * p.Wo;
      iy0 = oy * p.stride - p.pad;
      ix0 = ox * p.stride - p.pad;
    }
  } else {
    row_ok = (n0 + lrow) < p.N;
  }

  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < p.K; k0 += BK) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    const int kc = (k0 >> 3) + lch;
    if (loadA) {
      uint32_t e = p.ktab[kc];
      if (row_ok && (e & KT_VALID)) {
        int iy = iy0 + kt_ky(e), ix = ix0 + kt_kx(e);
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
          int64_t pix = (int64_t)(lb * p.H + iy) * p.W + ix;
          int c = kt_c(e);
          if (e & KT_SRCB) {
            load8(srcB + pix * p.b_pix_stride + p.b_ch_off + g * p.b_group_stride + c, v);
          } else {
            load8(srcA + pix * p.a_pix_stride + p.a_ch_off + c, v);
            if (p.a_scale) {
              const float* s = p.a_scale + (int64_t)lb * p.a_scale_stride + c;
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= s[j];
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) As[lch * 8 + j][lrow] = v[j];
    } else {
      if (row_ok) load8(wgt + ((int64_t)g * p.N + n0 + lrow) * p.K + (int64_t)kc * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) Bs[lch * 8 + j][lrow] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
  const int hw = p.Ho * p.Wo;
  const T* res1 = reinterpret_cast<const T*>(p.res1);
  const T* res2 = reinterpret_cast<const T*>(p.res2);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    int b = m / hw;
    int r = m - b * hw;
    int oy = r / p.Wo, ox = r - oy * p.Wo;
    int cs = 0;
    if (p.ncase == 9) cs = (oy == 0 ? 0 : (oy == p.Ho - 1 ? 2 : 1)) * 3 + (ox == 0 ? 0 : (ox == p.Wo - 1 ? 2 : 1));
    float vals[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      float v = acc[i][j];
      if (n < p.N) {
        int gn = g * p.N + n;
        if (p.scale) v *= p.scale[gn];
        if (p.bias_tab) v += p.bias_tab[(int64_t)cs * p.G * p.N + gn];
        if (p.act != ACT_SWIGLU) v = apply_act<true>(v, p.act);   // residuals are added after the activation
        if (res1) {
          int64_t rr = p.res1_row_mod ? (m % p.res1_row_mod) : m;
          v += to_f(res1[rr * p.res1_stride + gn]);
        }
        if (res2) v += to_f(res2[(int64_t)m * p.res2_stride + gn]);
      }
      vals[j] = v;
    }
    if (p.act == ACT_SWIGLU) {
      // pairs (x1, xg) -> x1 * silu(xg); tx*4 is even so pairs never straddle threads
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        int n = n0 + tx * 4 + j;
        if (n + 1 < p.n_valid[g] + 0 && n + 1 < p.N) {
          float o = vals[j] * silu_precise(vals[j + 1]);
          reinterpret_cast<T*>(p.out)[(int64_t)m * p.out_stride + p.out_ch_base[g] + (n >> 1)] = from_f<T>(o);
        }
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= p.N || n >= p.n_valid[g]) continue;
      int ch = p.out_ch_base[g] + n;
      if (p.out_layout == OUT_NCHW_F32) {
        reinterpret_cast<float*>(p.out)[(((int64_t)b * p.out_stride + ch) * p.Ho + oy) * p.Wo + ox] = vals[j];
      } else if (p.out_layout == OUT_NHWC_F32) {
        reinterpret_cast<float*>(p.out)[(int64_t)m * p.out_stride + ch] = vals[j];
      } else {
        reinterpret_cast<T*>(p.out)[(int64_t)m * p.out_stride + ch] = from_f<T>(vals[j]);
      }
    }
  }
}

}  // namespace

int conv_gemm_simt(const ConvGemmParams& p, cudaStream_t stream) {
  FTC_REQUIRE(p.K % BK == 0, "K must be padded to a multiple of 16");
  FTC_REQUIRE(p.G >= 1 && p.G <= MAX_GROUPS, "groups out of range");
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), p.G);
  if (p.dtype == DT_F32)
    conv_gemm_simt_kernel<float><<<grid, NT, 0, stream>>>(p);
  else
    conv_gemm_simt_kernel<bf16><<<grid, NT, 0, stream>>>(p);
  FTC_POST_LAUNCH();
  return 0;
}

}  // namespace ftc
