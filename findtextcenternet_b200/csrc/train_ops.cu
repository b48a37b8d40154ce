// Train-step building blocks of the detector (train1.py:128-170: forward in train mode, backward), first correct path:
// CUDA-core kernels with fp32 math over NHWC fp32 / bf16 tensors.  The forward convolutions of a train step run through
// the same implicit-GEMM kernels as inference (raw output: no folded BatchNorm); this file adds what inference never
// needs -- batch statistics, the BatchNorm + activation pair and its backward, the weight / data gradients of the dense and
// depthwise convolutions, the squeeze-excitation gate and its backward, and the adjoint of the bilinear upsample.
//
// Reference semantics (the modules autograd differentiates in train1.py):
//   nn.BatchNorm2d / BatchNorm1d in train mode (torch batch_norm: batch mean, BIASED variance for the normalisation, UNBIASED
//     variance into running_var), torchvision Conv2dNormActivation (ops/misc.py:69-126), SiLU, nn.GELU (erf),
//   torchvision MBConv / FusedMBConv (efficientnet.py:105-231), SqueezeExcitation (ops/misc.py:225-261),
//   nn.UpsamplingBilinear2d(scale_factor=2) (align_corners=True; models/detector.py:167-186).
#include <algorithm>
#include <stdlib.h>

#include "../../include/ftc_b200.h"
#include "common.cuh"
#ifndef FTC_EMU
#include "detector_ops.cuh"
#endif

namespace ftc {

namespace {

// ------------------------------------------------------------------------------------------------
// derivative of the activation at pre-activation z
__device__ __forceinline__ float act_grad(float z, int act) {
  if (act == ACT_SILU) {
    float s = sigmoid_precise(z);
    return s * (1.f + z * (1.f - s));
  }
  if (act == ACT_GELU) {   // d/dz [ z * Phi(z) ] = Phi(z) + z * phi(z)
    float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
    float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    return cdf + z * pdf;
  }
  return 1.f;
}

// bf16-storage variant: the derivative only has to be good to bf16 rounding (2^-9).  SiLU': one ex2 + one rcp; GELU': the cdf
// through the same 3-term tanh fit as the forward epilogue (abs. error 2.5e-5), one ex2 for the pdf.  (The precise version costs
// ~40 instructions per element and made the BatchNorm backward kernels issue-bound: 4x off the HBM roofline.)
__device__ __forceinline__ float act_grad_fast(float z, int act) {
#ifdef FTC_EMU
  return act_grad(z, act);
#else
  if (act == ACT_SILU) {
    const float s = __fdividef(1.f, 1.f + __expf(-z));
    return s * fmaf(z, 1.f - s, 1.f);
  }
  if (act == ACT_GELU) {
    const float z2 = fminf(z * z, 64.f);
    const float u = z * fmaf(z2, fmaf(z2, -3.51516792e-04f, 3.70056461e-02f), 7.97507884e-01f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
    const float cdf = fmaf(0.5f, t, 0.5f);
    return fmaf(z, 0.39894228040143267794f * __expf(-0.5f * z2), cdf);
  }
  return 1.f;
#endif
}
template <typename T> __device__ __forceinline__ float act_grad_t(float z, int act) {
  return sizeof(T) == 2 ? act_grad_fast(z, act) : act_grad(z, act);
}

// ------------------------------------------------------------------------------------------------
// Column reductions over a row-major [rows, C] matrix (NHWC activations: rows = B*H*W).
// Stage 1: CTA = (64 channels) x (4 row lanes); each CTA owns a contiguous chunk of rows and writes fp32 partials
// part[q][chunk][C]; stage 2: one thread per channel sums the partials in double.  Deterministic.
constexpr int RED_CH = 64, RED_LANES = 4, RED_THREADS = RED_CH * RED_LANES;

struct BnArgs {
  const float* mean; const float* var; const float* gamma; const float* beta;
  float eps; int act;
};

// MODE 0: (sum x, sum x^2) ; MODE 1: (sum dz, sum dz*xhat) with dz = dy * act'(gamma*xhat+beta)
template <typename T, int MODE>
__global__ void __launch_bounds__(RED_THREADS) col_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy, int64_t rows,
                                                                 int C, int64_t rows_per_chunk, float* __restrict__ part,
                                                                 BnArgs bn) {
  __shared__ float sm[2][RED_LANES][RED_CH];
  const int cl = threadIdx.x % RED_CH, lane = threadIdx.x / RED_CH;
  const int c = blockIdx.x * RED_CH + cl;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(rows, r0 + rows_per_chunk);
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    float mean = 0.f, rstd = 1.f, g = 1.f, bta = 0.f;
    if (MODE == 1) { mean = bn.mean[c]; rstd = rsqrtf(bn.var[c] + bn.eps); g = bn.gamma[c]; bta = bn.beta[c]; }
    for (int64_t r = r0 + lane; r < r1; r += RED_LANES) {
      float v = to_f(x[r * C + c]);
      if (MODE == 0) {
        s0 += v;
        s1 = fmaf(v, v, s1);
      } else {
        float xh = (v - mean) * rstd;
        float dz = to_f(dy[r * C + c]) * act_grad(fmaf(g, xh, bta), bn.act);
        s0 += dz;
        s1 = fmaf(dz, xh, s1);
      }
    }
  }
  sm[0][lane][cl] = s0;
  sm[1][lane][cl] = s1;
  __syncthreads();
  if (lane == 0 && c < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int l = 0; l < RED_LANES; ++l) { a += sm[0][l][cl]; b += sm[1][l][cl]; }
    const int64_t nchunk = gridDim.y;
    part[((int64_t)0 * nchunk + blockIdx.y) * C + c] = a;
    part[((int64_t)1 * nchunk + blockIdx.y) * C + c] = b;
  }
}

// FIN 0: out0 = mean, out1 = biased variance ; FIN 1: out0 = sum0, out1 = sum1
// CTA = 32 channels x 32 chunk lanes: lane ly adds chunks ly, ly + 32, ... in double, the 32 lane sums are then added in lane
// order by one thread per channel -- a fixed summation order (deterministic) with 32-way parallelism over the up-to-4096
// partials.  (One thread per channel walking all partials was latency-bound: ~50 us per launch, 724 launches per train step.)
struct RunningStats {           // nn.BatchNorm train-mode side effects (torch batch_norm: momentum update, UNBIASED variance)
  float* mean;                  // running_mean [C] or nullptr
  float* var;                   // running_var [C]
  long long* count;             // num_batches_tracked (scalar) or nullptr
  float momentum;
};

template <int FIN>
__global__ void __launch_bounds__(1024) col_reduce_finish_kernel(const float* __restrict__ part, int nchunk, int C, int64_t rows,
                                                                 float* __restrict__ out0, float* __restrict__ out1, RunningStats rs) {
  __shared__ double sa[32][33], sb[32][33];
  const int cx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  double a = 0.0, b = 0.0;
  if (c < C) {
    for (int k = ly; k < nchunk; k += 32) {
      a += (double)part[((int64_t)0 * nchunk + k) * C + c];
      b += (double)part[((int64_t)1 * nchunk + k) * C + c];
    }
  }
  sa[ly][cx] = a; sb[ly][cx] = b;
  __syncthreads();
  if (ly != 0 || c >= C) return;
  a = 0.0; b = 0.0;
  for (int j = 0; j < 32; ++j) { a += sa[j][cx]; b += sb[j][cx]; }
  if (FIN == 0) {
    double m = a / (double)rows;
    double v = b / (double)rows - m * m;
    const float mf = (float)m, vf = (float)(v > 0.0 ? v : 0.0);
    out0[c] = mf;
    out1[c] = vf;
    if (rs.mean != nullptr) {     // running = running * (1 - momentum) + momentum * batch statistic (variance: n / (n - 1) corrected)
      const float unbiased = vf * ((float)rows / (float)(rows > 1 ? rows - 1 : 1));
      rs.mean[c] = fmaf(rs.momentum, mf, rs.mean[c] * (1.0f - rs.momentum));
      rs.var[c] = fmaf(rs.momentum, unbiased, rs.var[c] * (1.0f - rs.momentum));
      if (c == 0 && rs.count != nullptr) *rs.count += 1;
    }
  } else {
    out0[c] = (float)a;
    out1[c] = (float)b;
  }
}

int red_chunks(int64_t rows) {
  // ~512 rows per CTA lane group keeps fp32 partial sums short; at most 4096 chunks
  int64_t n = (rows + 511) / 512;
  return (int)std::max<int64_t>(1, std::min<int64_t>(n, 4096));
}

// y = act(gamma * xhat + beta) (+ residual)
template <typename T>
__global__ void __launch_bounds__(256) bn_act_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total, int C, BnArgs bn,
                                                     const T* __restrict__ residual) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    float xh = (to_f(x[i]) - bn.mean[c]) * rsqrtf(bn.var[c] + bn.eps);
    float v = apply_act<true>(fmaf(bn.gamma[c], xh, bn.beta[c]), bn.act);
    if (residual) v += to_f(residual[i]);
    y[i] = from_f<T>(v);
  }
}

// dx = gamma * rstd * (dz - sum_dz / rows - xhat * sum_dz_xhat / rows)
template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                         int64_t total, int C, float inv_rows, BnArgs bn,
                                                         const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float rstd = rsqrtf(bn.var[c] + bn.eps), g = bn.gamma[c];
    float xh = (to_f(x[i]) - bn.mean[c]) * rstd;
    float dz = to_f(dy[i]) * act_grad(fmaf(g, xh, bn.beta[c]), bn.act);
    float v = g * rstd * (dz - sum_dz[c] * inv_rows - xh * sum_dz_xhat[c] * inv_rows);
    dx[i] = from_f<T>(v);
  }
}

// ---- 16-byte-vector versions of the three BatchNorm kernels (C % 8 == 0, every real layer): a thread owns 8 consecutive
// channels, so its per-channel constants live in registers and every load / store is a full 16-byte piece; the scalar kernels
// above remain for odd channel counts.  Reduction CTA = 8 chunk lanes (64 channels, one 128-byte line of bf16 per row) x 32
// row lanes.
template <typename T, int MODE, int U>
__global__ void __launch_bounds__(256, 3) col_reduce_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy, int64_t rows, int C,
                                                             int64_t rows_per_chunk, float* __restrict__ part, BnArgs bn) {
  __shared__ float sm[2][32][RED_CH + 1];
  const int ck = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * RED_CH + ck * 8;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
  if (c0 < C) {
    float mean[8], rstd[8], g[8], bt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = 0.f; rstd[j] = 1.f; g[j] = 1.f; bt[j] = 0.f;
      if (MODE == 1) { mean[j] = bn.mean[c0 + j]; rstd[j] = rsqrtf(bn.var[c0 + j] + bn.eps); g[j] = bn.gamma[c0 + j]; bt[j] = bn.beta[c0 + j]; }
    }
    // U rows per trip, all loads issued before the arithmetic: 4 x 2 x 16 B in flight per thread (the one-row loop kept one
    // or two loads in flight and ran at ~25 % of the HBM roofline)
    for (int64_t r = r0 + rl; r < r1; r += 32 * U) {
      float v[U][8], d[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + 32 * u;
        if (rr < r1) {
          load8(x + rr * C + c0, v[u]);
          if (MODE == 1) load8(dy + rr * C + c0, d[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (r + 32 * u >= r1) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (MODE == 0) {
            s0[j] += v[u][j];
            s1[j] = fmaf(v[u][j], v[u][j], s1[j]);
          } else {
            const float xh = (v[u][j] - mean[j]) * rstd[j];
            const float dz = d[u][j] * act_grad_t<T>(fmaf(g[j], xh, bt[j]), bn.act);
            s0[j] += dz;
            s1[j] = fmaf(dz, xh, s1[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sm[0][rl][ck * 8 + j] = s0[j]; sm[1][rl][ck * 8 + j] = s1[j]; }
  __syncthreads();
  if (threadIdx.x < 2 * RED_CH) {
    const int q = threadIdx.x / RED_CH, cl = threadIdx.x % RED_CH;
    const int c = blockIdx.x * RED_CH + cl;
    if (c < C) {
      float a = 0.f;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) a += sm[q][l][cl];
      part[((int64_t)q * gridDim.y + blockIdx.y) * C + c] = a;
    }
  }
}

template <typename T, int U>
__global__ void __launch_bounds__(256, 3) bn_act_vec_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int C, BnArgs bn,
                                                         const T* __restrict__ residual) {
  // thread = (chunk lane ck, row lane rl): a warp touches 4 rows x one 64-channel run (full 128-byte lines); grid.y walks the
  // 64-channel slabs, so a thread's eight (scale, shift) pairs are loop invariants
  const int ck = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.y * RED_CH + ck * 8;
  if (c0 >= C) return;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float rs = rsqrtf(bn.var[c0 + j] + bn.eps);
    sc[j] = bn.gamma[c0 + j] * rs;
    sh[j] = bn.beta[c0 + j] - bn.mean[c0 + j] * sc[j];
  }
  const int64_t stride = (int64_t)gridDim.x * 32;
  for (int64_t r = (int64_t)blockIdx.x * 32 + rl; r < rows; r += stride * U) {
    float v[U][8], res[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + stride * u;
      if (rr < rows) {
        load8(x + rr * C + c0, v[u]);
        if (residual) load8(residual + rr * C + c0, res[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + stride * u;
      if (rr >= rows) break;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // gamma * xhat + beta with the mean folded into the shift; bf16 storage: the 1-SFU activations of the inference epilogues
        const float z = fmaf(v[u][j], sc[j], sh[j]);
        v[u][j] = sizeof(T) == 2 ? apply_act<false>(z, bn.act) : apply_act<true>(z, bn.act);
        if (residual) v[u][j] += res[u][j];
      }
      store8(y + rr * C + c0, v[u]);
    }
  }
}

template <typename T, int U>
__global__ void __launch_bounds__(256, 3) bn_act_bwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
                                                             int64_t rows, int C, float inv_rows, BnArgs bn,
                                                             const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat) {
  const int ck = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.y * RED_CH + ck * 8;
  if (c0 >= C) return;
  float mean[8], rstd[8], g[8], bt[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = bn.mean[c0 + j]; rstd[j] = rsqrtf(bn.var[c0 + j] + bn.eps); g[j] = bn.gamma[c0 + j]; bt[j] = bn.beta[c0 + j];
    m1[j] = sum_dz[c0 + j] * inv_rows; m2[j] = sum_dz_xhat[c0 + j] * inv_rows;
  }
  const int64_t stride = (int64_t)gridDim.x * 32;
  for (int64_t r = (int64_t)blockIdx.x * 32 + rl; r < rows; r += stride * U) {
    float v[U][8], d[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + stride * u;
      if (rr < rows) { load8(x + rr * C + c0, v[u]); load8(dy + rr * C + c0, d[u]); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + stride * u;
      if (rr >= rows) break;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (v[u][j] - mean[j]) * rstd[j];
        const float dz = d[u][j] * act_grad_t<T>(fmaf(g[j], xh, bt[j]), bn.act);
        v[u][j] = g[j] * rstd[j] * (dz - m1[j] - xh * m2[j]);
      }
      store8(dx + rr * C + c0, v[u]);
    }
  }
}

// ---- stream versions of the three BatchNorm kernels for bf16 storage (the train step's configuration) -------------------------
// What the one-row-at-a-time kernels above lost, measured (tools/bench_bn.py, ncu r2g): (1) one or two 16-byte loads in flight per
// thread = ~10 KB per SM, against the ~40 KB Little's law asks of a 6.5 TB/s memory system; (2) the activation selected at run time
// and written with IEEE divisions / erff: ~5 000 SASS lines of slow paths, issue slots 45 % busy; (3) a 32-load + 8-rsqrt prologue
// per thread for eight rows of work; (4) row-major grids whose resident CTAs all read the same 128-byte column stripe (one 128 B
// burst per DRAM page).  Here: activation and residual are template parameters with the 1-SFU forms of the inference epilogues
// (tanh.approx), four row groups (4 x 16 B, twice that with dy / residual) in flight per thread, the channel slab is the FASTEST grid
// dimension (resident CTAs sweep whole rows), each CTA walks many row tiles, vector-loaded constants.  CK = 16-byte chunk lanes per
// row: 8 (64-channel slabs) or 4 (C = 32, 96: no idle lanes).
__device__ __forceinline__ float tanh_fast(float x) {
#ifdef FTC_EMU
  return tanhf(x);
#else
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
#endif
}
template <int ACT> __device__ __forceinline__ float act_fwd_fast(float z) {
  if (ACT == ACT_SILU) { const float h = 0.5f * z; return fmaf(h, tanh_fast(h), h); }                // z * sigmoid(z)
  if (ACT == ACT_GELU) {                                                                             // erf form to 2.5e-5
    const float z2 = fminf(z * z, 64.f);
    const float u = z * fmaf(z2, fmaf(z2, -3.51516792e-04f, 3.70056461e-02f), 7.97507884e-01f);
    const float h = 0.5f * z;
    return fmaf(h, tanh_fast(u), h);
  }
  return z;
}
template <int ACT> __device__ __forceinline__ float act_grad_fast_t(float z) {
  if (ACT == ACT_SILU) {
    const float s = fmaf(0.5f, tanh_fast(0.5f * z), 0.5f);
    return s * fmaf(z, 1.f - s, 1.f);
  }
  if (ACT == ACT_GELU) {
    const float z2 = fminf(z * z, 64.f);
    const float u = z * fmaf(z2, fmaf(z2, -3.51516792e-04f, 3.70056461e-02f), 7.97507884e-01f);
    const float cdf = fmaf(0.5f, tanh_fast(u), 0.5f);
#ifdef FTC_EMU
    const float pdf = 0.39894228040143267794f * expf(-0.5f * z2);
#else
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * z2);
#endif
    return fmaf(z, pdf, cdf);
  }
  return 1.f;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xFFFF0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xFFFF0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xFFFF0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ void ldf8(const float* __restrict__ p, float (&v)[8]) { load8(p, v); }

constexpr int BS_U = 4;      // row groups in flight per thread

template <int ACT, bool RES, int CK>
__global__ void __launch_bounds__(256) bn_apply_stream_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, int64_t rows, int C, BnArgs bn,
                                                              const bf16* __restrict__ residual) {
  constexpr int RL = 256 / CK;
  const int ck = threadIdx.x % CK, rl = threadIdx.x / CK;
  const int c0 = (blockIdx.x * CK + ck) * 8;
  if (c0 >= C) return;
  float sc[8], sh[8];
  {
    float m[8], v[8], g[8], b[8];
    ldf8(bn.mean + c0, m); ldf8(bn.var + c0, v); ldf8(bn.gamma + c0, g); ldf8(bn.beta + c0, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = g[j] * rsqrtf(v[j] + bn.eps); sh[j] = fmaf(-m[j], sc[j], b[j]); }
  }
  const int64_t tile = (int64_t)RL * BS_U;
  for (int64_t t0 = (int64_t)blockIdx.y * tile; t0 < rows; t0 += (int64_t)gridDim.y * tile) {
    uint4 xv[BS_U], rv[BS_U];
#pragma unroll
    for (int u = 0; u < BS_U; ++u) {
      const int64_t rr = t0 + u * RL + rl;
      if (rr < rows) {
        xv[u] = *reinterpret_cast<const uint4*>(x + rr * C + c0);
        if (RES) rv[u] = *reinterpret_cast<const uint4*>(residual + rr * C + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < BS_U; ++u) {
      const int64_t rr = t0 + u * RL + rl;
      if (rr < rows) {
        float v[8], r[8];
        unpack8(xv[u], v);
        if (RES) unpack8(rv[u], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = act_fwd_fast<ACT>(fmaf(v[j], sc[j], sh[j]));
          if (RES) v[j] += r[j];
        }
        store8(y + rr * C + c0, v);
      }
    }
  }
}

// MODE 0: sum x, sum x^2;  MODE 1: sum dz, sum dz * xhat with dz = dy * act'(gamma * xhat + beta).  part[q][blockIdx.y][C]
template <int MODE, int ACT, int CK>
__global__ void __launch_bounds__(256) bn_reduce_stream_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, int64_t rows, int C,
                                                               float* __restrict__ part, BnArgs bn, int64_t ldy) {   // ldy: dy row stride
  constexpr int RL = 256 / CK;
  __shared__ float sm[2][RL][CK * 8 + 1];
  const int ck = threadIdx.x % CK, rl = threadIdx.x / CK;
  const int c0 = (blockIdx.x * CK + ck) * 8;
  const bool live = c0 < C;
  float s0[8], s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
  if (live) {
    float a[8], b[8], g[8], bt[8];
    if (MODE == 1) {
      float m[8], v[8];
      ldf8(bn.mean + c0, m); ldf8(bn.var + c0, v); ldf8(bn.gamma + c0, g); ldf8(bn.beta + c0, bt);
#pragma unroll
      for (int j = 0; j < 8; ++j) { a[j] = rsqrtf(v[j] + bn.eps); b[j] = -m[j] * a[j]; }
    }
    const int64_t tile = (int64_t)RL * BS_U;
    for (int64_t t0 = (int64_t)blockIdx.y * tile; t0 < rows; t0 += (int64_t)gridDim.y * tile) {
      uint4 xv[BS_U], dv[BS_U];
#pragma unroll
      for (int u = 0; u < BS_U; ++u) {
        const int64_t rr = t0 + u * RL + rl;
        if (rr < rows) {
          xv[u] = *reinterpret_cast<const uint4*>(x + rr * C + c0);
          if (MODE == 1) dv[u] = *reinterpret_cast<const uint4*>(dy + rr * ldy + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < BS_U; ++u) {
        const int64_t rr = t0 + u * RL + rl;
        if (rr < rows) {
          float v[8], d[8];
          unpack8(xv[u], v);
          if (MODE == 1) unpack8(dv[u], d);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (MODE == 0) {
              s0[j] += v[j];
              s1[j] = fmaf(v[j], v[j], s1[j]);
            } else {
              const float xh = fmaf(v[j], a[j], b[j]);
              const float dz = d[j] * act_grad_fast_t<ACT>(fmaf(g[j], xh, bt[j]));
              s0[j] += dz;
              s1[j] = fmaf(dz, xh, s1[j]);
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sm[0][rl][ck * 8 + j] = s0[j]; sm[1][rl][ck * 8 + j] = s1[j]; }
  __syncthreads();
  if (threadIdx.x < 2 * CK * 8) {
    const int q = threadIdx.x / (CK * 8), cl = threadIdx.x % (CK * 8);
    const int c = blockIdx.x * CK * 8 + cl;
    if (c < C) {
      float acc = 0.f;
#pragma unroll 8
      for (int l = 0; l < RL; ++l) acc += sm[q][l][cl];
      part[((int64_t)q * gridDim.y + blockIdx.y) * C + c] = acc;
    }
  }
}

template <int ACT, int CK>
__global__ void __launch_bounds__(256) bn_bwd_stream_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, bf16* __restrict__ dx,
                                                            int64_t rows, int C, float inv_rows, BnArgs bn,
                                                            const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat, int64_t ldy) {
  constexpr int RL = 256 / CK;
  const int ck = threadIdx.x % CK, rl = threadIdx.x / CK;
  const int c0 = (blockIdx.x * CK + ck) * 8;
  if (c0 >= C) return;
  float a[8], b[8], g[8], bt[8], k[8], m1[8], m2[8];
  {
    float m[8], v[8];
    ldf8(bn.mean + c0, m); ldf8(bn.var + c0, v); ldf8(bn.gamma + c0, g); ldf8(bn.beta + c0, bt);
    ldf8(sum_dz + c0, m1); ldf8(sum_dz_xhat + c0, m2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] = rsqrtf(v[j] + bn.eps); b[j] = -m[j] * a[j]; k[j] = g[j] * a[j];
      m1[j] *= inv_rows; m2[j] *= inv_rows;
    }
  }
  const int64_t tile = (int64_t)RL * BS_U;
  for (int64_t t0 = (int64_t)blockIdx.y * tile; t0 < rows; t0 += (int64_t)gridDim.y * tile) {
    uint4 xv[BS_U], dv[BS_U];
#pragma unroll
    for (int u = 0; u < BS_U; ++u) {
      const int64_t rr = t0 + u * RL + rl;
      if (rr < rows) {
        xv[u] = *reinterpret_cast<const uint4*>(x + rr * C + c0);
        dv[u] = *reinterpret_cast<const uint4*>(dy + rr * ldy + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < BS_U; ++u) {
      const int64_t rr = t0 + u * RL + rl;
      if (rr < rows) {
        float v[8], d[8];
        unpack8(xv[u], v);
        unpack8(dv[u], d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = fmaf(v[j], a[j], b[j]);
          const float dz = d[j] * act_grad_fast_t<ACT>(fmaf(g[j], xh, bt[j]));
          v[j] = k[j] * (dz - m1[j] - xh * m2[j]);
        }
        store8(dx + rr * C + c0, v);
      }
    }
  }
}

// grid of the stream kernels: x = channel slabs (fastest: resident CTAs sweep whole rows), y = row walkers (each strides over row tiles)
inline dim3 stream_grid(int64_t rows, int c, int ck, int64_t max_y, int ctas_per_sm = 6) {
  const int slabs = (c + ck * 8 - 1) / (ck * 8);
  const int64_t tile = (int64_t)(256 / ck) * BS_U;
  int64_t y = (rows + tile - 1) / tile;
  const int64_t cap = std::max<int64_t>(1, (148 * ctas_per_sm + slabs - 1) / slabs);
  y = std::min<int64_t>(std::min<int64_t>(y, cap), std::max<int64_t>(1, max_y));
  return dim3((unsigned)slabs, (unsigned)y);
}
inline int stream_ck(int c) { return (c % 64 == 0) ? 8 : 4; }

#define FTC_BN_ACT_SWITCH(act, CALL)            \
  do {                                          \
    if ((act) == ACT_SILU) { CALL(ACT_SILU); }  \
    else if ((act) == ACT_GELU) { CALL(ACT_GELU); } \
    else { CALL(ACT_NONE); }                    \
  } while (0)

int ew_grid(int64_t total) { return (int)std::min<int64_t>((total + 255) / 256, 148 * 16); }

// ------------------------------------------------------------------------------------------------
// Dense convolution gradients, implicit GEMM on CUDA cores (64 x 64 x 16 tiles, 4 x 4 outputs per thread).
// k x k, pad (k-1)/2, stride 1|2; x [B,H,W,Cin], dy [B,Ho,Wo,Cout] NHWC; weights / gradient fp32 OIHW (the nn.Conv2d
// parameter layout, so the result lands in .grad unchanged).
constexpr int GB = 64, GK = 16, GT = 256;

struct ConvGeom { int B, H, W, Cin, Ho, Wo, Cout, k, stride, pad; };

// dW[co][ci][ky][kx] += sum_m dy[m][co] * x[b, oy*s - p + ky, ox*s - p + kx, ci]; M split over blockIdx.z, fp32 atomics
template <typename T>
__global__ void __launch_bounds__(GT) conv_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, ConvGeom g,
                                                        int64_t M, int64_t m_per_split, float* __restrict__ dw) {
  __shared__ float As[GK][GB + 4];   // dy   [m][co]
  __shared__ float Bs[GK][GB + 4];   // xcol [m][kk]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.y * GB, kk0 = blockIdx.x * GB;
  const int KK = g.k * g.k * g.Cin;
  const int64_t ms = (int64_t)blockIdx.z * m_per_split, me = min(M, ms + m_per_split);
  // loader mapping: column (co or kk) = tid & 63, rows tid >> 6 + 4 i
  const int lc = tid & 63, lr = tid >> 6;
  const int co_l = co0 + lc;
  const int kk_l = kk0 + lc;
  int ky = 0, kx = 0, ci = 0;
  if (kk_l < KK) { int tap = kk_l / g.Cin; ci = kk_l - tap * g.Cin; ky = tap / g.k; kx = tap - ky * g.k; }
  const int hw = g.Ho * g.Wo;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int64_t m0 = ms; m0 < me; m0 += GK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lr + 4 * i;
      const int64_t m = m0 + r;
      float a = 0.f, b = 0.f;
      if (m < me) {
        if (co_l < g.Cout) a = to_f(dy[m * g.Cout + co_l]);
        if (kk_l < KK) {
          int bi = (int)(m / hw);
          int rem = (int)(m - (int64_t)bi * hw);
          int oy = rem / g.Wo, ox = rem - oy * g.Wo;
          int iy = oy * g.stride - g.pad + ky, ix = ox * g.stride - g.pad + kx;
          if (iy >= 0 && iy < g.H && ix >= 0 && ix < g.W) b = to_f(x[(((int64_t)bi * g.H + iy) * g.W + ix) * g.Cin + ci]);
        }
      }
      As[r][lc] = a;
      Bs[r][lc] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kk = kk0 + tx * 4 + j;
      if (kk >= KK) continue;
      int tap = kk / g.Cin, c = kk - tap * g.Cin;
      atomicAdd(dw + ((int64_t)co * g.Cin + c) * g.k * g.k + tap, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on the warp-level tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate) -- STAGED: compiled in, selected
// only with FTC_WGRAD_MMA=1 until its first run on hardware (written in a session without GPU time).
//   dW[co][kk] = sum_m dy[m][co] * xcol[m][kk]: both operands have the reduction index m as their SLOW axis in memory (NHWC),
//   so the smem tiles keep the global layout ([m][co] and [m][kk], rows padded by 16 B) and BOTH fragments come from
//   ldmatrix.trans (the same idiom as the V operand of attention_mma_kernel, transformer_ops.cu).
// CTA = 128 co x 128 kk, 8 warps as 2 (co) x 4 (kk): warp tile 64 x 32 = 4 x 4 mma tiles; 32 pixels per stage, two stages of
// 16-byte cp.async (im2col gather with zero fill); pixel range split over blockIdx.z, fp32 atomics into OIHW.
// warp-level tensor-core primitives: PTX on the device; oracle/emu/cuda_emu.h provides host-thread versions with the PTX ISA's
// documented fragment layouts (same names), so the kernel below also runs under the CPU emulation
#ifndef FTC_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {   // 16 bytes, zero fill if !valid
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(valid ? 16u : 0u) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#endif

constexpr int WM_T = 128, WM_M = 32, WM_LD = WM_T + 8;
__global__ void __launch_bounds__(256) conv_wgrad_mma_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, ConvGeom g,
                                                             int64_t M, int64_t m_per_split, float* __restrict__ dw) {
  __shared__ __align__(16) bf16 sA[2][WM_M][WM_LD];   // dy   [m][co]
  __shared__ __align__(16) bf16 sB[2][WM_M][WM_LD];   // xcol [m][kk]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int co0 = blockIdx.y * WM_T, kk0 = blockIdx.x * WM_T;
  const int KK = g.k * g.k * g.Cin;
  const int64_t ms = (int64_t)blockIdx.z * m_per_split, me = min(M, ms + m_per_split);
  const int hw = g.Ho * g.Wo;
  // loader: 512 16-byte pieces per operand per stage; thread handles pieces tid and tid + 256: row = p >> 4, chunk = p & 15
  int l_row[2], l_co[2], l_ci[2], l_ky[2], l_kx[2];
  bool l_cok[2], l_kok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int p = tid + 256 * i;
    l_row[i] = p >> 4;
    const int ch = (p & 15) * 8;
    l_co[i] = co0 + ch;
    l_cok[i] = l_co[i] < g.Cout;
    const int kk = kk0 + ch;
    l_kok[i] = kk < KK;
    const int tap = l_kok[i] ? kk / g.Cin : 0;
    l_ci[i] = l_kok[i] ? kk - tap * g.Cin : 0;
    l_ky[i] = tap / g.k;
    l_kx[i] = tap - l_ky[i] * g.k;
  }
  auto load_stage = [&](int st, int64_t m0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t m = m0 + l_row[i];
      const bool mok = m < me;
      const int ch = ((tid + 256 * i) & 15) * 8;
      {
        const bool ok = mok && l_cok[i];
        const bf16* src = ok ? dy + m * g.Cout + l_co[i] : dy;
        cp_async16(&sA[st][l_row[i]][ch], src, ok);
      }
      {
        bool ok = mok && l_kok[i];
        const bf16* src = x;
        if (ok) {
          const int bi = (int)(m / hw);
          const int rem = (int)(m - (int64_t)bi * hw);
          const int oy = rem / g.Wo, ox = rem - oy * g.Wo;
          const int iy = oy * g.stride - g.pad + l_ky[i], ix = ox * g.stride - g.pad + l_kx[i];
          ok = iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
          if (ok) src = x + (((int64_t)bi * g.H + iy) * g.W + ix) * g.Cin + l_ci[i];
        }
        cp_async16(&sB[st][l_row[i]][ch], src, ok);
      }
    }
    cp_async_commit();
  };
  const int wy = warp >> 2, wx = warp & 3;
  float acc[4][4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;
  // ldmatrix.trans lane offsets (elements): A: k = (lane & 7) + 8 * (lane >> 4), co block = (lane >> 3) & 1
  //                                        B: k = (lane & 7) + 8 * ((lane >> 3) & 1), kk block = lane >> 4
  const int a_k = (lane & 7) + 8 * (lane >> 4), a_c = 8 * ((lane >> 3) & 1);
  const int b_k = (lane & 7) + 8 * ((lane >> 3) & 1), b_c = 8 * (lane >> 4);
  const int64_t nst = (me - ms + WM_M - 1) / WM_M;
  if (nst > 0) load_stage(0, ms);
  for (int64_t it = 0; it < nst; ++it) {
    const int st = (int)(it & 1);
    if (it + 1 < nst) {
      load_stage(st ^ 1, ms + (it + 1) * WM_M);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < WM_M / 16; ++ks) {
      uint32_t af[4][4], bq[2][4];     // bq[np] = {b0, b1 of n-tile 2np, b0, b1 of n-tile 2np+1}
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) ldmatrix_x4_trans(af[mt], &sA[st][ks * 16 + a_k][wy * 64 + mt * 16 + a_c]);
#pragma unroll
      for (int np = 0; np < 2; ++np) ldmatrix_x4_trans(bq[np], &sB[st][ks * 16 + b_k][wx * 32 + np * 16 + b_c]);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], af[mt], bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
    }
    __syncthreads();
  }
  const int gq = lane >> 2, tq = lane & 3;
  const int kk2 = g.k * g.k;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = co0 + wy * 64 + mt * 16 + gq + 8 * (e >> 1);
        const int kk = kk0 + wx * 32 + nt * 8 + 2 * tq + (e & 1);
        if (co < g.Cout && kk < KK) {
          const int tap = kk / g.Cin, c = kk - tap * g.Cin;
          atomicAdd(dw + ((int64_t)co * g.Cin + c) * kk2 + tap, acc[mt][nt][e]);
        }
      }
}

// dX[b,iy,ix,ci] = sum_{ky,kx,co} dy[b,oy,ox,co] * W[co][ci][ky][kx], oy*s - p + ky = iy, ox*s - p + kx = ix  (+ add)
template <typename T>
__global__ void __launch_bounds__(GT) conv_dgrad_kernel(const T* __restrict__ dy, const float* __restrict__ w, ConvGeom g, int64_t M,
                                                        const T* __restrict__ add, T* __restrict__ dx) {
  __shared__ float As[GK][GB + 4];   // gathered dy [kd][m]
  __shared__ float Bs[GK][GB + 4];   // weights     [kd][ci]
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * GB;
  const int ci0 = blockIdx.y * GB;
  const int KD = g.k * g.k * g.Cout, kk2 = g.k * g.k;
  // loader mapping: kd = tid & 15 (16 consecutive channels of one pixel), rows / columns tid >> 4 + 16 i
  const int lk = tid & 15, lq = tid >> 4;
  const int hw = g.H * g.W;
  int pb[4], py[4], px[4];
  bool pok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + lq + 16 * i;
    pok[i] = m < M;
    int bi = 0, iy = 0, ix = 0;
    if (pok[i]) { bi = (int)(m / hw); int rem = (int)(m - (int64_t)bi * hw); iy = rem / g.W; ix = rem - iy * g.W; }
    pb[i] = bi; py[i] = iy; px[i] = ix;
  }
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int kd0 = 0; kd0 < KD; kd0 += GK) {
    const int kd = kd0 + lk;
    int tap = 0, co = 0, ky = 0, kx = 0;
    const bool kok = kd < KD;
    if (kok) { tap = kd / g.Cout; co = kd - tap * g.Cout; ky = tap / g.k; kx = tap - ky * g.k; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = 0.f;
      if (kok && pok[i]) {
        int ny = py[i] + g.pad - ky, nx = px[i] + g.pad - kx;
        if (ny >= 0 && nx >= 0 && ny % g.stride == 0 && nx % g.stride == 0) {
          int oy = ny / g.stride, ox = nx / g.stride;
          if (oy < g.Ho && ox < g.Wo) a = to_f(dy[(((int64_t)pb[i] * g.Ho + oy) * g.Wo + ox) * g.Cout + co]);
        }
      }
      As[lk][lq + 16 * i] = a;
      float b = 0.f;
      const int c = ci0 + lq + 16 * i;
      if (kok && c < g.Cin) b = w[((int64_t)co * g.Cin + c) * kk2 + tap];
      Bs[lk][lq + 16 * i] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ci0 + tx * 4 + j;
      if (c >= g.Cin) continue;
      float v = acc[i][j];
      if (add) v += to_f(add[m * g.Cin + c]);
      dx[m * g.Cin + c] = from_f<T>(v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Depthwise 3x3 (pad 1, stride 1|2), weights fp32 [9][C] tap-major.  Raw output (train-mode BatchNorm follows).
template <typename T>
__global__ void __launch_bounds__(256) dw_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int Ho,
                                                     int Wo, int stride, const float* __restrict__ w) {
  const int64_t total = (int64_t)B * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int ox = (int)(p % Wo); p /= Wo;
    const int oy = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * stride - 1 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * stride - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        acc = fmaf(to_f(x[(((int64_t)b * H + iy) * W + ix) * C + c]), w[(ky * 3 + kx) * C + c], acc);
      }
    }
    y[i] = from_f<T>(acc);
  }
}

// dx[b,iy,ix,c] = sum_{ky,kx} dy[b,oy,ox,c] * w[ky,kx,c], oy*s - 1 + ky = iy, ox*s - 1 + kx = ix
template <typename T>
__global__ void __launch_bounds__(256) dw_dgrad_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int H, int W, int C,
                                                       int Ho, int Wo, int stride, const float* __restrict__ w) {
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t p = i / C;
    const int ix = (int)(p % W); p /= W;
    const int iy = (int)(p % H);
    const int b = (int)(p / H);
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ny = iy + 1 - ky;
      if (ny < 0 || ny % stride != 0) continue;
      const int oy = ny / stride;
      if (oy >= Ho) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int nx = ix + 1 - kx;
        if (nx < 0 || nx % stride != 0) continue;
        const int ox = nx / stride;
        if (ox >= Wo) continue;
        acc = fmaf(to_f(dy[(((int64_t)b * Ho + oy) * Wo + ox) * C + c]), w[(ky * 3 + kx) * C + c], acc);
      }
    }
    dx[i] = from_f<T>(acc);
  }
}

// 16-byte-vector depthwise forward / data gradient (C % 8 == 0): thread = (chunk lane, pixel lane), blockIdx.y = 64-channel slab,
// so the 9 x 8 tap weights are loop invariants in registers and every access is a full 16-byte piece of a 128-byte run.
// DGRAD = false: y[b,oy,ox] = sum_t x[b, oy*s-1+ky, ox*s-1+kx] * w[t];  DGRAD = true: dx[b,iy,ix] = sum_t dy[b,(iy+1-ky)/s,(ix+1-kx)/s] * w[t]
template <typename T, bool DGRAD>
__global__ void __launch_bounds__(256) dw_vec_kernel(const T* __restrict__ src, T* __restrict__ dst, int B, int Hs, int Ws, int Hd, int Wd,
                                                     int C, int stride, const float* __restrict__ w) {
  const int ck = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c0 = blockIdx.y * RED_CH + ck * 8;
  if (c0 >= C) return;
  float wk[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t) load8(w + (int64_t)t * C + c0, wk[t]);
  const int64_t P = (int64_t)B * Hd * Wd;
  for (int64_t p = (int64_t)blockIdx.x * 32 + pl; p < P; p += (int64_t)gridDim.x * 32) {
    int64_t q = p;
    const int dx_ = (int)(q % Wd); q /= Wd;
    const int dy_ = (int)(q % Hd);
    const int b = (int)(q / Hd);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      int sy;
      if (DGRAD) {
        const int ny = dy_ + 1 - ky;
        if (ny < 0 || ny % stride != 0) continue;
        sy = ny / stride;
      } else {
        sy = dy_ * stride - 1 + ky;
      }
      if (sy < 0 || sy >= Hs) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        int sx;
        if (DGRAD) {
          const int nx = dx_ + 1 - kx;
          if (nx < 0 || nx % stride != 0) continue;
          sx = nx / stride;
        } else {
          sx = dx_ * stride - 1 + kx;
        }
        if (sx < 0 || sx >= Ws) continue;
        float v[8];
        load8(src + (((int64_t)b * Hs + sy) * Ws + sx) * C + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(v[j], wk[ky * 3 + kx][j], acc[j]);
      }
    }
    store8(dst + p * C + c0, acc);
  }
}

// 16-byte-vector depthwise weight gradient (C % 8 == 0): thread = (8-channel chunk lane, pixel lane) as in dw_vec_kernel, CTA = 64
// channels x 32 pixel lanes walking a contiguous pixel range in raster order (neighbouring lanes touch neighbouring pixels: the
// nine shifted reads of x hit L1); 72 fp32 accumulators per thread, reduced over the pixel lanes by warp shuffles + shared
// memory, then one fp32 atomic per (tap, channel) per CTA into the zeroed gradient.
template <typename T>
__global__ void __launch_bounds__(256) dw_wgrad_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int H, int W, int C,
                                                           int Ho, int Wo, int stride, int pix_per_split, float* __restrict__ dw) {
  __shared__ float sm[8][9][RED_CH];
  const int ck = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * RED_CH + ck * 8;
  const int P = B * Ho * Wo;
  const int p0 = blockIdx.y * pix_per_split, p1 = min(P, p0 + pix_per_split);
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  if (c0 < C) {
    if constexpr (sizeof(T) == 2) {
      // bf16: the ten 16-byte loads of a pixel are issued together from clamped addresses (taps outside the image are masked
      // afterwards): with the boundary tests as branches around each load the compiler kept ONE load in flight per thread and
      // the kernel ran at 1.4 TB/s (12.9 ms of a B = 16 step, ncu r02l).  Tried and measured slower: one kernel row per CTA (24
      // accumulators, 4 CTAs per SM, grid z = 3): 11.2 vs 8.8 ms per step -- dy is read three times and the index arithmetic triples
      for (int p = p0 + pl; p < p1; p += 32) {
        const int ox = p % Wo, q = p / Wo;
        const int oy = q % Ho, b = q / Ho;
        const uint4 gv = *reinterpret_cast<const uint4*>(dy + (int64_t)p * C + c0);
        uint4 xv[9];
        unsigned okm = 0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = oy * stride - 1 + ky;
          const int iyc = min(max(iy, 0), H - 1);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * stride - 1 + kx;
            const int ixc = min(max(ix, 0), W - 1);
            xv[ky * 3 + kx] = *reinterpret_cast<const uint4*>(x + (((int64_t)b * H + iyc) * W + ixc) * C + c0);
            okm |= (iy == iyc && ix == ixc) ? (1u << (ky * 3 + kx)) : 0u;
          }
        }
        float g[8];
        unpack8(gv, g);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          float v[8];
          unpack8(xv[t], v);
          const float m = (okm >> t) & 1u ? 1.f : 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(g[j] * m, v[j], acc[t][j]);
        }
      }
    } else {
    for (int p = p0 + pl; p < p1; p += 32) {
      const int ox = p % Wo, q = p / Wo;
      const int oy = q % Ho, b = q / Ho;
      float g[8];
      load8(dy + (int64_t)p * C + c0, g);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride - 1 + ky;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * stride - 1 + kx;
          if (ix < 0 || ix >= W) continue;
          float v[8];
          load8(x + (((int64_t)b * H + iy) * W + ix) * C + c0, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[ky * 3 + kx][j] = fmaf(g[j], v[j], acc[ky * 3 + kx][j]);
        }
      }
    }
    }
  }
  // lanes of a warp: ck = lane & 7, four pixel lanes (lane >> 3): fold them, then the eight warps through shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[t][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) sm[warp][t][ck * 8 + j] = v;
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * RED_CH; i += 256) {
    const int t = i / RED_CH, cc = i - t * RED_CH;
    const int c = blockIdx.x * RED_CH + cc;
    if (c >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) s += sm[l][t][cc];
    atomicAdd(dw + (int64_t)t * C + c, s);
  }
}

// Shared-memory strip version of the depthwise weight gradient (bf16, stride 1, H % 4 == 0): CTA = 64 channels x a run of 4-row
// strips.  Per strip the 6 input rows (zero halo columns / rows included) and the 4 dy rows arrive through cp.async (the loads in
// flight no longer depend on registers: the register-resident version above runs at one or two CTAs per SM and is latency-bound,
// 8.8 ms of a B = 16 step for 1.8 ms of traffic); every thread then takes pixels of the strip from shared memory: 10 conflict-free
// 16-byte reads and 72 FMAs per (pixel, 8 channels).  The 72 accumulators live across all strips of the CTA; the final fold over
// pixel lanes / warps and the one atomic per (tap, channel) per CTA are those of the kernel above.
constexpr int DWS_R = 4;
__global__ void __launch_bounds__(256) dw_wgrad_strip_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy, int B, int H, int W, int C,
                                                             int strips_per_cta, float* __restrict__ dw) {
  extern __shared__ __align__(16) uint8_t dws_raw[];
  const int IW = W + 2;
  bf16* xs = reinterpret_cast<bf16*>(dws_raw);                          // [DWS_R + 2][IW][64]
  bf16* ds = xs + (size_t)(DWS_R + 2) * IW * RED_CH;                      // [DWS_R][W][64]
  const int tid = threadIdx.x, ck = tid & 7, pl = tid >> 3;
  const int cb = blockIdx.x * RED_CH;
  const int spi = H / DWS_R, n_strips = B * spi;
  const int s0 = blockIdx.y * strips_per_cta, s1 = min(n_strips, s0 + strips_per_cta);
  const bool live = cb + ck * 8 < C;
  float acc[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
  for (int st = s0; st < s1; ++st) {
    const int b = st / spi, y0 = (st - b * spi) * DWS_R;
    for (int i = tid; i < (DWS_R + 2) * IW * 8; i += 256) {
      const int pc = i & 7, q = i >> 3;
      const int row = q / IW, col = q - row * IW;
      const int iy = y0 - 1 + row, ix = col - 1;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W && cb + pc * 8 < C;
      const bf16* src = ok ? x + (((int64_t)b * H + iy) * W + ix) * C + cb + pc * 8 : x;
      cp_async16(xs + (size_t)q * RED_CH + pc * 8, src, ok);
    }
    for (int i = tid; i < DWS_R * W * 8; i += 256) {
      const int pc = i & 7, q = i >> 3;
      const bool ok = cb + pc * 8 < C;
      const bf16* src = ok ? dy + (((int64_t)b * H + y0) * W + q) * C + cb + pc * 8 : dy;
      cp_async16(ds + (size_t)q * RED_CH + pc * 8, src, ok);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    if (live) {
      for (int p = pl; p < DWS_R * W; p += 32) {
        const int oy = p / W, ox = p - oy * W;
        float g[8];
        load8(ds + (size_t)p * RED_CH + ck * 8, g);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            float v[8];
            load8(xs + ((size_t)(oy + ky) * IW + ox + kx) * RED_CH + ck * 8, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[ky * 3 + kx][j] = fmaf(g[j], v[j], acc[ky * 3 + kx][j]);
          }
      }
    }
    __syncthreads();                                                      // the strip buffers are refilled next
  }
  float (*sm)[9][RED_CH] = reinterpret_cast<float (*)[9][RED_CH]>(dws_raw);   // 18 KB, the strip buffers are free now
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[t][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8) sm[warp][t][ck * 8 + j] = v;
    }
  __syncthreads();
  for (int i = tid; i < 9 * RED_CH; i += 256) {
    const int t = i / RED_CH, cc = i - t * RED_CH;
    const int c = cb + cc;
    if (c >= C) continue;
    float a = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) a += sm[l][t][cc];
    atomicAdd(dw + (int64_t)t * C + c, a);
  }
}

// dw[t][c] += sum_{b,oy,ox} dy[b,oy,ox,c] * x[b, oy*s-1+ky, ox*s-1+kx, c]; CTA = 32 channels x 8 pixel lanes, pixel range
// split over blockIdx.y, fp32 atomics into a zeroed buffer
template <typename T>
__global__ void __launch_bounds__(256) dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, int B, int H, int W, int C,
                                                       int Ho, int Wo, int stride, int64_t pix_per_split, float* __restrict__ dw) {
  __shared__ float sm[8][9][32];
  const int cl = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int64_t P = (int64_t)B * Ho * Wo;
  const int64_t p0 = (int64_t)blockIdx.y * pix_per_split, p1 = min(P, p0 + pix_per_split);
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  if (c < C) {
    for (int64_t p = p0 + lane; p < p1; p += 8) {
      int64_t q = p;
      const int ox = (int)(q % Wo); q /= Wo;
      const int oy = (int)(q % Ho);
      const int b = (int)(q / Ho);
      const float g = to_f(dy[p * C + c]);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride - 1 + ky;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = ox * stride - 1 + kx;
          if (ix < 0 || ix >= W) continue;
          acc[ky * 3 + kx] = fmaf(g, to_f(x[(((int64_t)b * H + iy) * W + ix) * C + c]), acc[ky * 3 + kx]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) sm[lane][t][cl] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * 32; i += 256) {
    const int t = i / 32, cc = i - t * 32;
    if (blockIdx.x * 32 + cc >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < 8; ++l) s += sm[l][t][cc];
    atomicAdd(dw + (int64_t)t * C + blockIdx.x * 32 + cc, s);
  }
}

// ------------------------------------------------------------------------------------------------
// Squeeze-excitation pieces.  x [B, HW, C].
// spatial sums: out[b][c] = scale * sum_hw f(x, y)  with f = x (MUL 0) or x*y (MUL 1); CTA = (64 channels x 4 lanes) per image
template <typename T, int MUL>
__global__ void __launch_bounds__(RED_THREADS) spatial_sum_kernel(const T* __restrict__ x, const T* __restrict__ y, int HW, int C,
                                                                  float scale, float* __restrict__ out) {
  __shared__ float sm[RED_LANES][RED_CH];
  const int cl = threadIdx.x % RED_CH, lane = threadIdx.x / RED_CH;
  const int c = blockIdx.x * RED_CH + cl, b = blockIdx.y;
  // two-level accumulation (runs of 64 pixels) keeps the fp32 sums of up to 36 864 pixels accurate
  float total = 0.f;
  if (c < C) {
    const T* xb = x + (int64_t)b * HW * C + c;
    const T* yb = MUL ? y + (int64_t)b * HW * C + c : nullptr;
    float run = 0.f;
    int n = 0;
    for (int p = lane; p < HW; p += RED_LANES) {
      float v = to_f(xb[(int64_t)p * C]);
      if (MUL) v *= to_f(yb[(int64_t)p * C]);
      run += v;
      if (++n == 64) { total += run; run = 0.f; n = 0; }
    }
    total += run;
  }
  sm[lane][cl] = total;
  __syncthreads();
  if (lane == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int l = 0; l < RED_LANES; ++l) s += sm[l][cl];
    out[(int64_t)b * C + c] = s * scale;
  }
}

// y = x * s[b][c]     (also StochasticDepth "row" mode with s constant per image)
template <typename T>
__global__ void __launch_bounds__(256) scale_bc_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ y, int HW,
                                                       int C, int64_t total, const float* __restrict__ bias_bc, float bias_mul) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int b = (int)(i / ((int64_t)HW * C));
    float v = to_f(x[i]) * s[(int64_t)b * C + c];
    if (bias_bc) v = fmaf(bias_bc[(int64_t)b * C + c], bias_mul, v);
    y[i] = from_f<T>(v);
  }
}

// 16-byte-vector versions (C % 8 == 0): thread = (chunk lane, row lane) as in the BatchNorm kernels; blockIdx.z = image, so the
// per-(image, channel) factors are loop invariants
#ifdef FTC_EMU
constexpr int SSUM_THREADS = 256;
#else
constexpr int SSUM_THREADS = 1024;    // product sums (SE backward): 128 pixel lanes, (C / 64) x B CTAs are few (384 at 48 x 48 x 1536, B = 16):
                                      // 5.4 -> 3.5 ms per step; the plain sums stay at 256 threads (1024: 1.8 -> 2.4 ms, the serial lane fold)
#endif
template <typename T, int MUL>
__global__ void __launch_bounds__(1024) spatial_sum_vec_kernel(const T* __restrict__ x, const T* __restrict__ y, int HW, int C,
                                                               float scale, float* __restrict__ out) {
  __shared__ float sm[128][RED_CH + 1];
  const int ck = threadIdx.x & 7, rl = threadIdx.x >> 3, nrl = blockDim.x >> 3;
  const int c0 = blockIdx.x * RED_CH + ck * 8, b = blockIdx.y;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < C) {
    const T* xb = x + (int64_t)b * HW * C + c0;
    const T* yb = MUL ? y + (int64_t)b * HW * C + c0 : nullptr;
    // U pixel groups per trip, loads first (one load in flight per thread ran at ~1.5 TB/s); fixed summation order.  U = 2 with the
    // second operand: 1024-thread CTAs leave 64 registers per thread
    constexpr int U = MUL ? 2 : 4;
    for (int p = rl; p < HW; p += U * nrl) {
      float v[U][8], w[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pp = p + nrl * u;
        if (pp < HW) {
          load8(xb + (int64_t)pp * C, v[u]);
          if (MUL) load8(yb + (int64_t)pp * C, w[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (p + nrl * u < HW) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += MUL ? v[u][j] * w[u][j] : v[u][j];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[rl][ck * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < RED_CH) {
    const int c = blockIdx.x * RED_CH + threadIdx.x;
    if (c < C) {
      float a = 0.f;
      for (int l = 0; l < nrl; ++l) a += sm[l][threadIdx.x];
      out[(int64_t)b * C + c] = a * scale;
    }
  }
}
template <typename T>
__global__ void __launch_bounds__(256) scale_bc_vec_kernel(const T* __restrict__ x, const float* __restrict__ s, T* __restrict__ y, int HW,
                                                           int C, const float* __restrict__ bias_bc, float bias_mul) {
  // grid: x = channel slab (fastest: resident CTAs sweep whole rows), y = pixel walkers, z = image; four pixel groups per trip
  const int ck = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * RED_CH + ck * 8, b = blockIdx.z;
  if (c0 >= C) return;
  float sc[8], bi[8];
  load8(s + (int64_t)b * C + c0, sc);
  if (bias_bc) {
    load8(bias_bc + (int64_t)b * C + c0, bi);
#pragma unroll
    for (int j = 0; j < 8; ++j) bi[j] *= bias_mul;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) bi[j] = 0.f;
  }
  const T* xb = x + (int64_t)b * HW * C + c0;
  T* yb = y + (int64_t)b * HW * C + c0;
  for (int p = blockIdx.y * 128 + rl; p < HW; p += gridDim.y * 128) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (p + 32 * u < HW) load8(xb + (int64_t)(p + 32 * u) * C, v[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (p + 32 * u < HW) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[u][j] = fmaf(v[u][j], sc[j], bi[j]);
        store8(yb + (int64_t)(p + 32 * u) * C, v[u]);
      }
    }
  }
}

// SE excitation of one image per CTA: hid_pre = W1 mean + b1 ; hid = silu ; gate_pre = W2 hid + b2 ; gate = sigmoid
// W1 [S][C], W2 [C][S] row-major (the squeezed conv weights of ops/misc.py:247-248)
// The two fully connected layers of SqueezeExcitation are tiny GEMVs per image (C <= 3840, S <= 160).  One CTA per image was
// latency-bound (B = 2: two CTAs walking 2 x 2.4 MB of weights with strided accesses, 0.3 ms per launch, a quarter of the whole
// train step under ncu); every phase is now a grid-filling kernel with coalesced weight rows:
//   fwd 1  warp per (b, s)        hid_pre[b,s] = b1[s] + w1[s,:] . mean[b,:]
//   fwd 2  warp per (b, c)        gate[b,c]    = sigmoid(b2[c] + w2[c,:] . silu(hid_pre[b,:]))
//   bwd 1  CTA  per (b, 32 s)     dgp = dgate*g*(1-g);  dhp[b,s] = silu'(hid_pre) * sum_c w2[c,s] dgp[b,c]   (lanes over s: row-coalesced)
//   bwd 2  thread per (b, c)      dmean[b,c]   = sum_s w1[s,c] dhp[b,s]
// Summation orders are fixed (no atomics).
__global__ void __launch_bounds__(256) se_fc_fwd1_kernel(const float* __restrict__ mean, int C, int S, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, float* __restrict__ hid_pre) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * 8 + warp;
  if (s >= S) return;
  const float* mb = mean + (int64_t)b * C;
  float a = 0.f;
  for (int c = lane; c < C; c += 32) a = fmaf(w1[(int64_t)s * C + c], mb[c], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) hid_pre[(int64_t)b * S + s] = a + b1[s];
}
__global__ void __launch_bounds__(256) se_fc_fwd2_kernel(const float* __restrict__ hid_pre, int C, int S, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, float* __restrict__ gate) {
  extern __shared__ float se_sm[];   // [S] silu(hid_pre[b,:])
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = threadIdx.x; s < S; s += 256) se_sm[s] = silu_precise(hid_pre[(int64_t)b * S + s]);
  __syncthreads();
  for (int c = blockIdx.x * 64 + warp; c < C && c < blockIdx.x * 64 + 64; c += 8) {
    float a = 0.f;
    for (int s = lane; s < S; s += 32) a = fmaf(w2[(int64_t)c * S + s], se_sm[s], a);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) gate[(int64_t)b * C + c] = sigmoid_precise(a + b2[c]);
  }
}
#ifdef FTC_EMU
constexpr int SE_BWD1_THREADS = 256;
#else
constexpr int SE_BWD1_THREADS = 1024;
#endif
__global__ void __launch_bounds__(1024) se_fc_bwd1_kernel(const float* __restrict__ dgate, const float* __restrict__ gate,
                                                          const float* __restrict__ hid_pre, int C, int S,
                                                          const float* __restrict__ w2, float* __restrict__ dgp,
                                                          float* __restrict__ dhp) {
  // 32 warps walk the channels (warp w: c = w, w + 32, ...), eight channels per trip with all loads issued first: C = 3 072 is 12
  // dependent trips instead of the 384 of the first version (one channel per trip, 8 warps: 53 us per launch for a few hundred KB)
  __shared__ float red[32][32];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int s = blockIdx.x * 32 + lane;
  const bool first = blockIdx.x == 0;              // the first s-chunk's CTA also publishes dgp (the weight kernel reads it)
  float a = 0.f;
  for (int cb = warp; cb < C; cb += nw * 8) {
    float gt[8], dg[8], wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = cb + nw * u;
      const bool in = c < C;
      gt[u] = in ? gate[(int64_t)b * C + c] : 0.f;
      dg[u] = in ? dgate[(int64_t)b * C + c] : 0.f;
      wv[u] = (in && s < S) ? w2[(int64_t)c * S + s] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int c = cb + nw * u;
      if (c < C) {
        const float v = dg[u] * gt[u] * (1.f - gt[u]);
        if (first && lane == 0) dgp[(int64_t)b * C + c] = v;
        if (s < S) a = fmaf(wv[u], v, a);
      }
    }
  }
  red[warp][lane] = a;
  __syncthreads();
  if (warp == 0 && s < S) {
    float t = red[0][lane];
    for (int w = 1; w < nw; ++w) t += red[w][lane];
    dhp[(int64_t)b * S + s] = t * act_grad(hid_pre[(int64_t)b * S + s], ACT_SILU);
  }
}
__global__ void __launch_bounds__(256) se_fc_bwd2_kernel(const float* __restrict__ dhp, int C, int S, const float* __restrict__ w1,
                                                         float* __restrict__ dmean) {
  extern __shared__ float se_sm[];   // [S] dhp[b,:]
  const int b = blockIdx.y;
  for (int s = threadIdx.x; s < S; s += 256) se_sm[s] = dhp[(int64_t)b * S + s];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a = fmaf(w1[(int64_t)s * C + c], se_sm[s], a);
  dmean[(int64_t)b * C + c] = a;
}
__global__ void __launch_bounds__(256) se_fc_bwd_weight_kernel(const float* __restrict__ dgp, const float* __restrict__ dhp,
                                                               const float* __restrict__ hid_pre, const float* __restrict__ mean,
                                                               int B, int C, int S, float* __restrict__ dw1, float* __restrict__ db1,
                                                               float* __restrict__ dw2, float* __restrict__ db2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)C * S) return;
  {   // dW2 element (c, s) = i
    const int c = (int)(i / S), s = (int)(i - (int64_t)c * S);
    float a = 0.f, bsum = 0.f;
    for (int b = 0; b < B; ++b) {
      float g = dgp[(int64_t)b * C + c];
      a = fmaf(g, silu_precise(hid_pre[(int64_t)b * S + s]), a);
      bsum += g;
    }
    dw2[i] = a;
    if (s == 0) db2[c] = bsum;
  }
  {   // dW1 element (s, c) = i
    const int s = (int)(i / C), c = (int)(i - (int64_t)s * C);
    float a = 0.f, bsum = 0.f;
    for (int b = 0; b < B; ++b) {
      float g = dhp[(int64_t)b * S + s];
      a = fmaf(g, mean[(int64_t)b * C + c], a);
      bsum += g;
    }
    dw1[i] = a;
    if (c == 0) db1[s] = bsum;
  }
}

// ------------------------------------------------------------------------------------------------
// adjoint of the bilinear x2 upsample (align_corners=True): dx[b,iy,ix,:] = sum over the output pixels whose 2x2 source
// footprint contains (iy, ix), with the forward's own weights.  Source coordinate of output o: f = o * (n-1)/(2n-1)
// (computed as in upsample2x_kernel, detector_ops.cu), so the candidates of input index j are the outputs with floor(f) in
// {j-1, j}: o in [2j-2, 2j+3] clipped.  CTA per input row, warps over pixels, lanes over channels.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int H, int W, int C,
                                                             float sy, float sx, int64_t ldy) {   // ldy: dy pixel stride (>= C: a channel slice of a wider map)
  const int Ho = 2 * H, Wo = 2 * W;
  const int b = blockIdx.y, iy = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // row candidates and their weights (CTA constants)
  float wy[6];
  int oys[6];
  int ny = 0;
  for (int oy = max(0, 2 * iy - 2); oy <= min(Ho - 1, 2 * iy + 3); ++oy) {
    const float fy = sy * oy;
    const int y0 = min((int)fy, H - 1), y1 = min(y0 + 1, H - 1);
    const float ly = fminf(fmaxf(fy - y0, 0.f), 1.f), hy = 1.f - ly;
    float wgt = 0.f;
    if (y0 == iy) wgt += hy;
    if (y1 == iy) wgt += ly;
    if (wgt != 0.f) { wy[ny] = wgt; oys[ny] = oy; ++ny; }
  }
  for (int ix = warp; ix < W; ix += 8) {
    float wx[6];
    int oxs[6];
    int nx = 0;
    for (int ox = max(0, 2 * ix - 2); ox <= min(Wo - 1, 2 * ix + 3); ++ox) {
      const float fx = sx * ox;
      const int x0 = min((int)fx, W - 1), x1 = min(x0 + 1, W - 1);
      const float lx = fminf(fmaxf(fx - x0, 0.f), 1.f), hx = 1.f - lx;
      float wgt = 0.f;
      if (x0 == ix) wgt += hx;
      if (x1 == ix) wgt += lx;
      if (wgt != 0.f) { wx[nx] = wgt; oxs[nx] = ox; ++nx; }
    }
    T* po = dx + (((int64_t)b * H + iy) * W + ix) * C;
    if (VEC) {     // C % 8 == 0: a lane owns 8 consecutive channels, 16-byte loads / stores
      for (int c = lane * 8; c < C; c += 256) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int a = 0; a < ny; ++a) {
          const T* row = dy + (((int64_t)b * Ho + oys[a]) * Wo) * ldy + c;
          for (int k = 0; k < nx; ++k) {
            float v[8];
            load8(row + (int64_t)oxs[k] * ldy, v);
            const float wgt = wy[a] * wx[k];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
          }
        }
        store8(po + c, acc);
      }
      continue;
    }
    for (int c = lane; c < C; c += 32) {
      float acc = 0.f;
      for (int a = 0; a < ny; ++a) {
        const T* row = dy + (((int64_t)b * Ho + oys[a]) * Wo) * ldy + c;
        float r = 0.f;
        for (int k = 0; k < nx; ++k) r = fmaf(wx[k], to_f(row[(int64_t)oxs[k] * ldy]), r);
        acc = fmaf(wy[a], r, acc);
      }
      po[c] = from_f<T>(acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Transformer train-step pieces (train3.py:132-137 differentiates models/transformer.py:58-253).
// LayerNorm over the last axis of [rows, D] with up to two residual inputs folded in (EncoderBlock / DecoderBlock:
// LN(ff + _x + skip), models/transformer.py:149-160,196-211): xs = x (+ r1) (+ r2) is written out because the backward
// needs it; mean / rstd fp32 per row.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ r1, const T* __restrict__ r2,
                                                     T* __restrict__ xs, T* __restrict__ y, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out, int64_t rows, int D,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int64_t base = row * D;
  float s = 0.f;
  for (int e = lane; e < D; e += 32) {
    float v = to_f(x[base + e]);
    if (r1) v += to_f(r1[base + e]);
    if (r2) v += to_f(r2[base + e]);
    if (xs) { xs[base + e] = from_f<T>(v); v = to_f(from_f<T>(v)); }   // statistics of the stored (rounded) sum
    s += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / D;
  const T* src = xs ? xs : x;     // same thread wrote the elements it re-reads
  float q = 0.f;
  for (int e = lane; e < D; e += 32) { float d = to_f(src[base + e]) - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / D + eps);
  for (int e = lane; e < D; e += 32) y[base + e] = from_f<T>(fmaf((to_f(src[base + e]) - mean) * rstd, gamma[e], beta[e]));
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma
template <typename T>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const T* __restrict__ xs, const T* __restrict__ dy, T* __restrict__ dx,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd, int64_t rows,
                                                     int D, const float* __restrict__ gamma) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int64_t base = row * D;
  const float mu = mean[row], rs = rstd[row];
  float a = 0.f, b = 0.f;
  for (int e = lane; e < D; e += 32) {
    float g = to_f(dy[base + e]) * gamma[e];
    float xh = (to_f(xs[base + e]) - mu) * rs;
    a += g;
    b = fmaf(g, xh, b);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  a /= D; b /= D;
  for (int e = lane; e < D; e += 32) {
    float g = to_f(dy[base + e]) * gamma[e];
    float xh = (to_f(xs[base + e]) - mu) * rs;
    dx[base + e] = from_f<T>(rs * (g - a - xh * b));
  }
}

// column partials of (dy, dy * xhat) with per-ROW statistics: same two-stage scheme as col_reduce_kernel
template <typename T>
__global__ void __launch_bounds__(RED_THREADS) ln_col_reduce_kernel(const T* __restrict__ xs, const T* __restrict__ dy,
                                                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                    int64_t rows, int D, int64_t rows_per_chunk,
                                                                    float* __restrict__ part) {
  __shared__ float sm[2][RED_LANES][RED_CH];
  const int cl = threadIdx.x % RED_CH, lane = threadIdx.x / RED_CH;
  const int c = blockIdx.x * RED_CH + cl;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float s0 = 0.f, s1 = 0.f;
  if (c < D) {
    for (int64_t r = r0 + lane; r < r1; r += RED_LANES) {
      float g = to_f(dy[r * D + c]);
      s0 += g;
      s1 = fmaf(g, (to_f(xs[r * D + c]) - mean[r]) * rstd[r], s1);
    }
  }
  sm[0][lane][cl] = s0;
  sm[1][lane][cl] = s1;
  __syncthreads();
  if (lane == 0 && c < D) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int l = 0; l < RED_LANES; ++l) { a += sm[0][l][cl]; b += sm[1][l][cl]; }
    const int64_t nchunk = gridDim.y;
    part[((int64_t)0 * nchunk + blockIdx.y) * D + c] = a;
    part[((int64_t)1 * nchunk + blockIdx.y) * D + c] = b;
  }
}

// SwiGLU gate (models/transformer.py:66-69): h = x1 * silu(xg) and its backward
template <typename T>
__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const T* __restrict__ x1, const T* __restrict__ xg, T* __restrict__ h,
                                                         int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    h[i] = from_f<T>(to_f(x1[i]) * silu_precise(to_f(xg[i])));
}
template <typename T>
__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const T* __restrict__ x1, const T* __restrict__ xg, const T* __restrict__ dh,
                                                         T* __restrict__ dx1, T* __restrict__ dxg, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float g = to_f(xg[i]), d = to_f(dh[i]);
    dx1[i] = from_f<T>(d * silu_precise(g));
    dxg[i] = from_f<T>(d * to_f(x1[i]) * act_grad(g, ACT_SILU));
  }
}

// Decoder token embedding (models/transformer.py:226-233): out[row] = sum_i E_i[token[row] mod m_i]; tables fp32 [m_i][D]
struct Embed3 { const float* e[3]; float* de[3]; int m[3]; };
template <typename T>
__global__ void __launch_bounds__(256) embed3_fwd_kernel(const int64_t* __restrict__ tok, Embed3 tb, T* __restrict__ out,
                                                         int64_t rows, int D) {
  const int64_t total = rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / D;
    const int e = (int)(i - row * D);
    const int64_t t = tok[row];
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) v += tb.e[k][(int64_t)(((t % tb.m[k]) + tb.m[k]) % tb.m[k]) * D + e];
    out[i] = from_f<T>(v);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) embed3_bwd_kernel(const int64_t* __restrict__ tok, Embed3 tb, const T* __restrict__ dy,
                                                         int64_t rows, int D) {
  const int64_t total = rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / D;
    const int e = (int)(i - row * D);
    const int64_t t = tok[row];
    const float g = to_f(dy[i]);
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(tb.de[k] + (int64_t)(((t % tb.m[k]) + tb.m[k]) % tb.m[k]) * D + e, g);
  }
}

// Attention backward (F.scaled_dot_product_attention with an additive key mask, models/transformer.py:133).
// q / dout / dq: [B*Lt, D], k / v / dk / dv: [B*Ls, D], head h at columns h*hd.  Two kernels, no atomics:
//  A (warp per query row): recompute s = scale q.k + mask, p = softmax(s); dP_j = dout.v_j; dS = p * (dP - sum_j p dP);
//    store p and dS rows in the [B,H,Lt,Ls] scratch; dq = scale * sum_j dS_j k_j
//  B (warp per key row):   dv_j = sum_i p_ij dout_i ; dk_j = scale * sum_i dS_ij q_i
constexpr int ATT_MAX_HD = 128;
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_rows_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                            const float* __restrict__ mask, const T* __restrict__ dout,
                                                            float* __restrict__ P, float* __restrict__ dS, float* __restrict__ dq,
                                                            int H, int hd, int Lt, int Ls, int D, float scale) {
  extern __shared__ float att_sm[];   // per warp: q[hd], do[hd], s[Ls], dp[Ls]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= Lt) return;                 // whole warp exits together (no block-level barrier below)
  float* sq = att_sm + (size_t)warp * (2 * hd + 2 * Ls);
  float* sdo = sq + hd;
  float* ss = sdo + hd;
  float* sdp = ss + Ls;
  const int64_t qrow = ((int64_t)b * Lt + i) * D + (int64_t)h * hd;
  for (int e = lane; e < hd; e += 32) { sq[e] = to_f(q[qrow + e]); sdo[e] = to_f(dout[qrow + e]); }
  __syncwarp();
  float mx = -INFINITY;
  for (int j = lane; j < Ls; j += 32) {
    const int64_t krow = ((int64_t)b * Ls + j) * D + (int64_t)h * hd;
    float a = 0.f, c = 0.f;
    for (int e = 0; e < hd; ++e) { a = fmaf(sq[e], to_f(k[krow + e]), a); c = fmaf(sdo[e], to_f(v[krow + e]), c); }
    a = a * scale + (mask ? mask[(int64_t)b * Ls + j] : 0.f);
    ss[j] = a;
    sdp[j] = c;
    mx = fmaxf(mx, a);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float l = 0.f;
  for (int j = lane; j < Ls; j += 32) { float p = expf(ss[j] - mx); ss[j] = p; l += p; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  const float inv = 1.f / l;
  float delta = 0.f;
  for (int j = lane; j < Ls; j += 32) { float p = ss[j] * inv; ss[j] = p; delta = fmaf(p, sdp[j], delta); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
  const int64_t prow = (((int64_t)b * H + h) * Lt + i) * Ls;
  for (int j = lane; j < Ls; j += 32) {
    const float p = ss[j], ds = p * (sdp[j] - delta);
    P[prow + j] = p;
    dS[prow + j] = ds;
    sdp[j] = ds;
  }
  __syncwarp();
  for (int e = lane; e < hd; e += 32) {
    float a = 0.f;
    for (int j = 0; j < Ls; ++j) a = fmaf(sdp[j], to_f(k[((int64_t)b * Ls + j) * D + (int64_t)h * hd + e]), a);
    dq[qrow + e] = a * scale;
  }
}

template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_cols_kernel(const T* __restrict__ q, const T* __restrict__ dout,
                                                            const float* __restrict__ P, const float* __restrict__ dS,
                                                            float* __restrict__ dk, float* __restrict__ dv, int H, int hd, int Lt,
                                                            int Ls, int D, float scale) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 4 + warp, h = blockIdx.y, b = blockIdx.z;
  if (j >= Ls) return;
  const int64_t krow = ((int64_t)b * Ls + j) * D + (int64_t)h * hd;
  const int64_t pbase = (((int64_t)b * H + h) * Lt) * Ls + j;
  for (int e = lane; e < hd; e += 32) {
    float av = 0.f, ak = 0.f;
    for (int i = 0; i < Lt; ++i) {
      const int64_t qrow = ((int64_t)b * Lt + i) * D + (int64_t)h * hd + e;
      av = fmaf(P[pbase + (int64_t)i * Ls], to_f(dout[qrow]), av);
      ak = fmaf(dS[pbase + (int64_t)i * Ls], to_f(q[qrow]), ak);
    }
    dv[krow + e] = av;
    dk[krow + e] = ak * scale;
  }
}

// ------------------------------------------------------------------------------------------------
// Page maps of run_detector (process_ocr_base.py:480-520): every tile contributes sigmoid(heatmap channel) inside its
// validity window (0 outside) and the page keeps the maximum where tiles overlap.  Maps: key (ch 0), textline (ch 3),
// separator (ch 4), code1/2/4/8 (ch 5-8) of the 9-channel heatmap.  Values are >= 0, so the float maximum is an integer
// atomicMax on the bit pattern; HBM-bound (7 of 9 channels read once, page written by atomics into L2).
__global__ void __launch_bounds__(256) page_maps_kernel(const float* __restrict__ heat9, int B, int H, int W,
                                                        const int* __restrict__ tile_meta, int* __restrict__ page, int PH, int PW,
                                                        int scale) {
  const int64_t total = (int64_t)B * 7 * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    int64_t r = i / W;
    const int y = (int)(r % H); r /= H;
    const int m = (int)(r % 7);
    const int b = (int)(r / 7);
    const int* tm = tile_meta + b * 6;
    const int py = tm[1] / scale + y, px = tm[0] / scale + x;
    if (py >= PH || px >= PW) continue;
    const int ch = m == 0 ? 0 : m + 2;
    float v = 0.f;
    if (x >= tm[2] && x < tm[3] && y >= tm[4] && y < tm[5])
      v = (tanhf(heat9[(((int64_t)b * 9 + ch) * H + y) * W + x] * 0.5f) + 1.0f) * 0.5f;   // util_func.py:14 sigmoid
    if (v > 0.f) atomicMax(page + ((int64_t)m * PH + py) * PW + px, __float_as_int(v));
  }
}

template <typename T> const T* cp(const void* p) { return reinterpret_cast<const T*>(p); }
template <typename T> T* mp(void* p) { return reinterpret_cast<T*>(p); }

bool dtype_ok(int dt) { return dt == DT_F32 || dt == DT_BF16; }

// grid.x of the row-walking element kernels (CTA = 32 row lanes x one 64-channel slab, `slabs` slabs in grid.y).  Every thread
// pays a prologue of ~100 instructions for its eight channels' parameters (loads, rsqrt, folding), so it must own several rows:
// ncu (round 2) showed one row per thread on the 48x48 / 24x24 layers -- 36 us for a 28 MB tensor, 1.5 TB/s.  Target eight rows per
// thread, but never fewer than ~2 CTAs per SM in total.
unsigned ew_rows_grid(int64_t rows, int slabs) {
  const int64_t max_x = (rows + 31) / 32;
  int64_t x = (rows + 32 * 8 - 1) / (32 * 8);
  const int64_t want = (148 * 2 + slabs - 1) / slabs;
  if (x < want) x = want;
  if (x > max_x) x = max_x;
  if (x > 148 * 8) x = 148 * 8;
  return (unsigned)(x > 0 ? x : 1);
}

// bf16 BatchNorm kernels: FTC_BN_UNROLL = 0 stream kernels (default) | 1 the earlier one-row-per-trip vector kernels (A/B in
// tools/bench_bn.py; they also serve fp32 and channel counts that are not multiples of 32)
int g_bn_unroll = -1;
int bn_unroll() {
  if (g_bn_unroll >= 0) return g_bn_unroll;
  static const int env = [] { const char* e = getenv("FTC_BN_UNROLL"); return e ? atoi(e) : 0; }();   // 0 = stream kernels (default)
  return env;
}

// staged mma.sync weight gradient: -1 = follow FTC_WGRAD_MMA (read once), 0 / 1 = set by ftc_debug_set_wgrad_mma
int g_wgrad_mma = -1;
bool wgrad_mma_enabled() {
  if (g_wgrad_mma >= 0) return g_wgrad_mma != 0;
  static const bool env = [] { const char* e = getenv("FTC_WGRAD_MMA"); return !e || atoi(e) != 0; }();   // default ON
  return env;
}

// ---- attention backward, one CTA per (batch, head), everything in shared memory (sequences up to 128) -------------------------------
// The row / column kernels above read K, V, Q, dO rows straight from global memory with a different row per lane and keep P / dS in a
// global scratch: 149 of the 177 ms of a train3 step (batch 64, 16 + 16 blocks of d = 512).  Here Q, K, V, dO of one (batch, head) are
// staged once (storage type), S = scale QK^T + mask and dP = dO V^T are formed as 4 x 4 register tiles into two fp32 matrices, a warp
// per row turns them into P and dS = P (dP - sum_j P dP), and dQ = scale dS K, dK = scale dS^T Q, dV = P^T dO are read off the
// shared matrices.  No atomics, fixed summation order.
// two consecutive operand elements as floats (one 32-bit / 64-bit shared-memory load; operand rows are 4-byte aligned: RS is even)
__device__ __forceinline__ float2 ld_pair(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ float2 ld_pair(const bf16* p) {
  const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}

template <typename T>
__global__ void __launch_bounds__(256) attn_bwd_fused_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                             const float* __restrict__ mask, const T* __restrict__ dout,
                                                             float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                                                             int hd, int Lt, int Ls, int D, float scale) {
  extern __shared__ __align__(16) uint8_t afb_raw[];
  const int RS = hd + 2;                                   // operand row stride (elements, even): 4 rows apart = 4 (bf16) / 8 (fp32) banks
  const int PS = Ls + 1;                                   // matrix row stride (floats)
  float* Pm = reinterpret_cast<float*>(afb_raw);           // [Lt][PS]  S, then P
  float* Dm = Pm + (size_t)Lt * PS;                        // [Lt][PS]  dP, then dS
  T* Qs = reinterpret_cast<T*>(Dm + (size_t)Lt * PS);      // [Lt][RS]
  T* Os = Qs + (size_t)Lt * RS;                            // [Lt][RS]  dO
  T* Ks = Os + (size_t)Lt * RS;                            // [Ls][RS]
  T* Vs = Ks + (size_t)Ls * RS;                            // [Ls][RS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, h = blockIdx.x, b = blockIdx.y;
  const int64_t qbase = (int64_t)b * Lt * D + (int64_t)h * hd, kbase = (int64_t)b * Ls * D + (int64_t)h * hd;
  for (int i = tid; i < Lt * hd; i += 256) {
    const int r = i / hd, e = i - r * hd;
    Qs[r * RS + e] = q[qbase + (int64_t)r * D + e];
    Os[r * RS + e] = dout[qbase + (int64_t)r * D + e];
  }
  for (int i = tid; i < Ls * hd; i += 256) {
    const int r = i / hd, e = i - r * hd;
    Ks[r * RS + e] = k[kbase + (int64_t)r * D + e];
    Vs[r * RS + e] = v[kbase + (int64_t)r * D + e];
  }
  __syncthreads();
  // ---- S = scale Q K^T + mask and dP = dO V^T as 4 x 4 register tiles.  A warp takes a patch of 4 (query) x 8 (key) tiles: the
  // eight key-tile lanes read eight different banks, lanes that share a tile row / column read the same word (broadcast). ----
  const int tq = (Lt + 3) / 4, tk = (Ls + 3) / 4;
  const int pq = (tq + 3) / 4, pk = (tk + 7) / 8;
  for (int pidx = warp; pidx < pq * pk; pidx += 8) {
    const int it = (pidx / pk) * 4 + (lane >> 3), jt = (pidx % pk) * 8 + (lane & 7);
    if (it >= tq || jt >= tk) continue;
    const int i0 = it * 4, j0 = jt * 4;
    int ri[4], rj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { ri[a] = min(i0 + a, Lt - 1) * RS; rj[a] = min(j0 + a, Ls - 1) * RS; }
    float sacc[4][4], pacc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) { sacc[a][c] = 0.f; pacc[a][c] = 0.f; }
    for (int e = 0; e < hd; e += 2) {
      float2 qa[4], oa[4], ka[4], va[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        qa[a] = ld_pair(Qs + ri[a] + e); oa[a] = ld_pair(Os + ri[a] + e);
        ka[a] = ld_pair(Ks + rj[a] + e); va[a] = ld_pair(Vs + rj[a] + e);
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          sacc[a][c] = fmaf(qa[a].y, ka[c].y, fmaf(qa[a].x, ka[c].x, sacc[a][c]));
          pacc[a][c] = fmaf(oa[a].y, va[c].y, fmaf(oa[a].x, va[c].x, pacc[a][c]));
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + a, j = j0 + c;
        if (i < Lt && j < Ls) {
          Pm[i * PS + j] = sacc[a][c] * scale + (mask ? mask[(int64_t)b * Ls + j] : 0.f);
          Dm[i * PS + j] = pacc[a][c];
        }
      }
  }
  __syncthreads();
  // ---- rows: P = softmax(S), dS = P (dP - sum_j P dP); one warp per row ----
  for (int i = warp; i < Lt; i += 8) {
    float* pr = Pm + i * PS;
    float* dr = Dm + i * PS;
    float mx = -INFINITY;
    for (int j = lane; j < Ls; j += 32) mx = fmaxf(mx, pr[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float l = 0.f;
    for (int j = lane; j < Ls; j += 32) { const float pv = expf(pr[j] - mx); pr[j] = pv; l += pv; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    const float inv = 1.f / l;
    float delta = 0.f;
    for (int j = lane; j < Ls; j += 32) { const float pv = pr[j] * inv; pr[j] = pv; delta = fmaf(pv, dr[j], delta); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) delta += __shfl_xor_sync(0xffffffffu, delta, o);
    for (int j = lane; j < Ls; j += 32) dr[j] = pr[j] * (dr[j] - delta);
  }
  __syncthreads();
  // ---- dQ = scale dS K: thread = 4 query rows x 2 head-dim columns (8 FMAs per 5 shared-memory loads) ----
  const int he = hd / 2;
  for (int t = tid; t < tq * he; t += 256) {
    const int i0 = (t / he) * 4, e = (t % he) * 2;
    const float* dr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) dr[a] = Dm + min(i0 + a, Lt - 1) * PS;
    float acc[4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a) { acc[a][0] = 0.f; acc[a][1] = 0.f; }
    for (int j = 0; j < Ls; ++j) {
      const float2 kv = ld_pair(Ks + j * RS + e);
#pragma unroll
      for (int a = 0; a < 4; ++a) { const float d = dr[a][j]; acc[a][0] = fmaf(d, kv.x, acc[a][0]); acc[a][1] = fmaf(d, kv.y, acc[a][1]); }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (i0 + a < Lt) {
        dq[qbase + (int64_t)(i0 + a) * D + e] = acc[a][0] * scale;
        dq[qbase + (int64_t)(i0 + a) * D + e + 1] = acc[a][1] * scale;
      }
  }
  // ---- dV = P^T dO, dK = scale dS^T Q: thread = 4 keys x 2 head-dim columns (16 FMAs per 10 loads) ----
  for (int t = tid; t < tk * he; t += 256) {
    const int j0 = (t / he) * 4, e = (t % he) * 2;
    int cj[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) cj[c] = min(j0 + c, Ls - 1);
    float av[4][2], ak[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) { av[c][0] = av[c][1] = 0.f; ak[c][0] = ak[c][1] = 0.f; }
    for (int i = 0; i < Lt; ++i) {
      const float2 ov = ld_pair(Os + i * RS + e), qv = ld_pair(Qs + i * RS + e);
      const float* pr = Pm + i * PS;
      const float* dr = Dm + i * PS;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float pv = pr[cj[c]], dsv = dr[cj[c]];
        av[c][0] = fmaf(pv, ov.x, av[c][0]); av[c][1] = fmaf(pv, ov.y, av[c][1]);
        ak[c][0] = fmaf(dsv, qv.x, ak[c][0]); ak[c][1] = fmaf(dsv, qv.y, ak[c][1]);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (j0 + c < Ls) {
        const int64_t o = kbase + (int64_t)(j0 + c) * D + e;
        dv[o] = av[c][0]; dv[o + 1] = av[c][1];
        dk[o] = ak[c][0] * scale; dk[o + 1] = ak[c][1] * scale;
      }
  }
}

}  // namespace
}  // namespace ftc

using namespace ftc;

extern "C" {

size_t ftc_train_reduce_scratch_bytes(int64_t rows, int c) {
  return (size_t)2 * red_chunks(rows) * (size_t)c * sizeof(float);
}

static int bn_stats_impl(const void* x, int dtype, int64_t rows, int c, float* mean, float* var, void* scratch, RunningStats rs,
                         void* stream);

int ftc_train_bn_stats(const void* x, int dtype, int64_t rows, int c, float* mean, float* var, void* scratch, void* stream) {
  return bn_stats_impl(x, dtype, rows, c, mean, var, scratch, RunningStats{nullptr, nullptr, nullptr, 0.f}, stream);
}

int ftc_train_bn_stats_running(const void* x, int dtype, int64_t rows, int c, float* mean, float* var, void* scratch,
                               float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, void* stream) {
  FTC_REQUIRE(running_mean && running_var, "null running statistics");
  return bn_stats_impl(x, dtype, rows, c, mean, var, scratch,
                       RunningStats{running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked), momentum}, stream);
}

static int bn_stats_impl(const void* x, int dtype, int64_t rows, int c, float* mean, float* var, void* scratch, RunningStats rs,
                         void* stream) {
  FTC_REQUIRE(x && mean && var && scratch && rows > 0 && c > 0 && dtype_ok(dtype), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int nchunk = red_chunks(rows);
  const int64_t rpc = (rows + nchunk - 1) / nchunk;
  dim3 grid(ceil_div(c, RED_CH), nchunk);
  BnArgs bn = {};
  const bool vec = c % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec && dtype == DT_BF16 && c % 32 == 0 && bn_unroll() == 0) {
    const int ck = stream_ck(c);
    const dim3 sg = stream_grid(rows, c, ck, nchunk, 3);   // all CTAs resident; few partial rows for the finish
    if (ck == 8) bn_reduce_stream_kernel<0, ACT_NONE, 8><<<sg, 256, 0, s>>>(cp<bf16>(x), nullptr, rows, c, (float*)scratch, bn, c);
    else bn_reduce_stream_kernel<0, ACT_NONE, 4><<<sg, 256, 0, s>>>(cp<bf16>(x), nullptr, rows, c, (float*)scratch, bn, c);
    FTC_POST_LAUNCH();
    col_reduce_finish_kernel<0><<<ceil_div(c, 32), 1024, 0, s>>>((const float*)scratch, (int)sg.y, c, rows, mean, var, rs);
    FTC_POST_LAUNCH();
    return 0;
  }
  if (vec && dtype == DT_F32)
    col_reduce_vec_kernel<float, 0, 1><<<grid, 256, 0, s>>>(cp<float>(x), nullptr, rows, c, rpc, (float*)scratch, bn);
  else if (vec)
    col_reduce_vec_kernel<bf16, 0, 1><<<grid, 256, 0, s>>>(cp<bf16>(x), nullptr, rows, c, rpc, (float*)scratch, bn);
  else if (dtype == DT_F32)
    col_reduce_kernel<float, 0><<<grid, RED_THREADS, 0, s>>>(cp<float>(x), nullptr, rows, c, rpc, (float*)scratch, bn);
  else
    col_reduce_kernel<bf16, 0><<<grid, RED_THREADS, 0, s>>>(cp<bf16>(x), nullptr, rows, c, rpc, (float*)scratch, bn);
  FTC_POST_LAUNCH();
  col_reduce_finish_kernel<0><<<ceil_div(c, 32), 1024, 0, s>>>((const float*)scratch, nchunk, c, rows, mean, var, rs);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_bn_act(const void* x, void* y, int dtype, int64_t rows, int c, const float* mean, const float* var,
                     const float* gamma, const float* beta, float eps, int act, const void* residual, void* stream) {
  FTC_REQUIRE(x && y && mean && var && gamma && beta && rows > 0 && c > 0 && dtype_ok(dtype), "bad argument");
  FTC_REQUIRE(act == ACT_NONE || act == ACT_SILU || act == ACT_GELU, "activation");
  cudaStream_t s = (cudaStream_t)stream;
  BnArgs bn = {mean, var, gamma, beta, eps, act};
  const int64_t total = rows * c;
  const bool vec = c % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0;
  const dim3 vgrid(ew_rows_grid(rows, ceil_div(c, RED_CH)), ceil_div(c, RED_CH));
  const bool par16 = ((reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(var) | reinterpret_cast<uintptr_t>(gamma) |
                       reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (vec && par16 && dtype == DT_BF16 && c % 32 == 0 && bn_unroll() == 0) {
    const int ck = stream_ck(c);
    const dim3 sg = stream_grid(rows, c, ck, 1 << 30);
#define BN_APPLY(A)                                                                                                                      \
    if (ck == 8) {                                                                                                                       \
      if (residual) bn_apply_stream_kernel<A, true, 8><<<sg, 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), rows, c, bn, cp<bf16>(residual));      \
      else bn_apply_stream_kernel<A, false, 8><<<sg, 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), rows, c, bn, nullptr);                        \
    } else {                                                                                                                             \
      if (residual) bn_apply_stream_kernel<A, true, 4><<<sg, 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), rows, c, bn, cp<bf16>(residual));      \
      else bn_apply_stream_kernel<A, false, 4><<<sg, 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), rows, c, bn, nullptr);                        \
    }
    FTC_BN_ACT_SWITCH(act, BN_APPLY);
#undef BN_APPLY
    FTC_POST_LAUNCH();
    return 0;
  }
  if (vec && dtype == DT_F32)
    bn_act_vec_kernel<float, 1><<<vgrid, 256, 0, s>>>(cp<float>(x), mp<float>(y), rows, c, bn, cp<float>(residual));
  else if (vec)
    bn_act_vec_kernel<bf16, 1><<<vgrid, 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), rows, c, bn, cp<bf16>(residual));
  else if (dtype == DT_F32)
    bn_act_kernel<float><<<ew_grid(total), 256, 0, s>>>(cp<float>(x), mp<float>(y), total, c, bn, cp<float>(residual));
  else
    bn_act_kernel<bf16><<<ew_grid(total), 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), total, c, bn, cp<bf16>(residual));
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_bn_act_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t rows, int c, const float* mean,
                         const float* var, const float* gamma, const float* beta, float eps, int act, float* dbeta,
                         float* dgamma, void* scratch, void* stream) {
  return ftc_train_bn_act_bwd_ld(x, dy, c, dx, dtype, rows, c, mean, var, gamma, beta, eps, act, dbeta, dgamma, scratch, stream);
}

int ftc_train_bn_act_bwd_ld(const void* x, const void* dy, int64_t dy_ld, void* dx, int dtype, int64_t rows, int c, const float* mean,
                            const float* var, const float* gamma, const float* beta, float eps, int act, float* dbeta,
                            float* dgamma, void* scratch, void* stream) {
  FTC_REQUIRE(x && dy && dx && mean && var && gamma && beta && dbeta && dgamma && scratch && rows > 0 && c > 0 && dtype_ok(dtype),
              "bad argument");
  FTC_REQUIRE(dy_ld >= c, "dy row stride below the channel count");
  FTC_REQUIRE(act == ACT_NONE || act == ACT_SILU || act == ACT_GELU, "activation");
  cudaStream_t s = (cudaStream_t)stream;
  BnArgs bn = {mean, var, gamma, beta, eps, act};
  const int nchunk = red_chunks(rows);
  const int64_t rpc = (rows + nchunk - 1) / nchunk;
  dim3 grid(ceil_div(c, RED_CH), nchunk);
  const bool vec = c % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  const bool par16 = ((reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(var) | reinterpret_cast<uintptr_t>(gamma) |
                       reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(dbeta) | reinterpret_cast<uintptr_t>(dgamma)) & 15) == 0;
  const bool stream_path = vec && par16 && dtype == DT_BF16 && c % 32 == 0 && dy_ld % 8 == 0 && bn_unroll() == 0;
  FTC_REQUIRE(stream_path || dy_ld == c, "a row-strided dy is taken by the bf16 stream kernels only (C % 32 == 0): pass a contiguous dy");
  if (stream_path) {
    const int ck = stream_ck(c);
    const dim3 rg = stream_grid(rows, c, ck, nchunk, 6);   // (2 per SM measured slower: 27.5 vs 24.5 ms per step, tools/bench_bn.py)
#define BN_RED(A)                                                                                                                  \
    if (ck == 8) bn_reduce_stream_kernel<1, A, 8><<<rg, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), rows, c, (float*)scratch, bn, dy_ld);        \
    else bn_reduce_stream_kernel<1, A, 4><<<rg, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), rows, c, (float*)scratch, bn, dy_ld);
    FTC_BN_ACT_SWITCH(act, BN_RED);
#undef BN_RED
    FTC_POST_LAUNCH();
    col_reduce_finish_kernel<1><<<ceil_div(c, 32), 1024, 0, s>>>((const float*)scratch, (int)rg.y, c, rows, dbeta, dgamma, RunningStats{nullptr, nullptr, nullptr, 0.f});
    FTC_POST_LAUNCH();
    const float inv_rows_s = (float)(1.0 / (double)rows);
    const dim3 ag = stream_grid(rows, c, ck, 1 << 30);
#define BN_BWD(A)                                                                                                                              \
    if (ck == 8) bn_bwd_stream_kernel<A, 8><<<ag, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), mp<bf16>(dx), rows, c, inv_rows_s, bn, dbeta, dgamma, dy_ld);   \
    else bn_bwd_stream_kernel<A, 4><<<ag, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), mp<bf16>(dx), rows, c, inv_rows_s, bn, dbeta, dgamma, dy_ld);
    FTC_BN_ACT_SWITCH(act, BN_BWD);
#undef BN_BWD
    FTC_POST_LAUNCH();
    return 0;
  }
  if (vec && dtype == DT_F32)
    col_reduce_vec_kernel<float, 1, 1><<<grid, 256, 0, s>>>(cp<float>(x), cp<float>(dy), rows, c, rpc, (float*)scratch, bn);
  else if (vec)
    col_reduce_vec_kernel<bf16, 1, 1><<<grid, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), rows, c, rpc, (float*)scratch, bn);
  else if (dtype == DT_F32)
    col_reduce_kernel<float, 1><<<grid, RED_THREADS, 0, s>>>(cp<float>(x), cp<float>(dy), rows, c, rpc, (float*)scratch, bn);
  else
    col_reduce_kernel<bf16, 1><<<grid, RED_THREADS, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), rows, c, rpc, (float*)scratch, bn);
  FTC_POST_LAUNCH();
  col_reduce_finish_kernel<1><<<ceil_div(c, 32), 1024, 0, s>>>((const float*)scratch, nchunk, c, rows, dbeta, dgamma, RunningStats{nullptr, nullptr, nullptr, 0.f});
  FTC_POST_LAUNCH();
  const int64_t total = rows * c;
  const float inv_rows = (float)(1.0 / (double)rows);
  const dim3 vgrid(ew_rows_grid(rows, ceil_div(c, RED_CH)), ceil_div(c, RED_CH));
  if (vec && dtype == DT_F32)
    bn_act_bwd_vec_kernel<float, 1><<<vgrid, 256, 0, s>>>(cp<float>(x), cp<float>(dy), mp<float>(dx), rows, c, inv_rows, bn, dbeta, dgamma);
  else if (vec)
    bn_act_bwd_vec_kernel<bf16, 1><<<vgrid, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), mp<bf16>(dx), rows, c, inv_rows, bn, dbeta, dgamma);
  else if (dtype == DT_F32)
    bn_act_bwd_kernel<float><<<ew_grid(total), 256, 0, s>>>(cp<float>(x), cp<float>(dy), mp<float>(dx), total, c, inv_rows, bn, dbeta,
                                                           dgamma);
  else
    bn_act_bwd_kernel<bf16><<<ew_grid(total), 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), mp<bf16>(dx), total, c, inv_rows, bn, dbeta,
                                                          dgamma);
  FTC_POST_LAUNCH();
  return 0;
}

static int conv_geom(ConvGeom* g, int batch, int h, int w, int cin, int cout, int ksize, int stride) {
  FTC_REQUIRE(batch > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, "bad geometry");
  FTC_REQUIRE(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
  FTC_REQUIRE(stride == 1 || stride == 2, "stride must be 1 or 2");
  g->B = batch; g->H = h; g->W = w; g->Cin = cin; g->Cout = cout; g->k = ksize; g->stride = stride; g->pad = (ksize - 1) / 2;
  g->Ho = (h - 1) / stride + 1; g->Wo = (w - 1) / stride + 1;
  return 0;
}

int ftc_train_conv2d_wgrad(const void* x, const void* dy, int dtype, int batch, int h, int w, int cin, int cout, int ksize,
                           int stride, float* dw_oihw, void* stream) {
  FTC_REQUIRE(x && dy && dw_oihw && dtype_ok(dtype), "bad argument");
  ConvGeom g;
  int rc = conv_geom(&g, batch, h, w, cin, cout, ksize, stride);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int KK = ksize * ksize * cin;
  const int64_t M = (int64_t)batch * g.Ho * g.Wo;
  FTC_CHECK_CUDA(cudaMemsetAsync(dw_oihw, 0, (size_t)cout * KK * sizeof(float), s));
  const int tiles = ceil_div(KK, GB) * ceil_div(cout, GB);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>((148 * 4 + tiles - 1) / tiles, (M + 255) / 256));
  splits = std::min<int64_t>(splits, 65535);
  int64_t mps = (M + splits - 1) / splits;
  mps = (mps + GK - 1) / GK * GK;
  splits = (M + mps - 1) / mps;
  const bool use_mma = wgrad_mma_enabled();
  if (use_mma && dtype == DT_BF16 && cin % 8 == 0 && cout % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0) {   // 16-byte cp.async pieces
    const int tiles2 = ceil_div(KK, WM_T) * ceil_div(cout, WM_T);
    int64_t sp = std::max<int64_t>(1, std::min<int64_t>((148 * 2 + tiles2 - 1) / tiles2, (M + 255) / 256));
    sp = std::min<int64_t>(sp, 65535);
    int64_t per = (M + sp - 1) / sp;
    per = (per + WM_M - 1) / WM_M * WM_M;
    sp = (M + per - 1) / per;
    dim3 grid2(ceil_div(KK, WM_T), ceil_div(cout, WM_T), (unsigned)sp);
    conv_wgrad_mma_kernel<<<grid2, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), g, M, per, dw_oihw);
    FTC_POST_LAUNCH();
    return 0;
  }
  dim3 grid(ceil_div(KK, GB), ceil_div(cout, GB), (unsigned)splits);
  if (dtype == DT_F32)
    conv_wgrad_kernel<float><<<grid, GT, 0, s>>>(cp<float>(x), cp<float>(dy), g, M, mps, dw_oihw);
  else
    conv_wgrad_kernel<bf16><<<grid, GT, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), g, M, mps, dw_oihw);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_conv2d_dgrad(const void* dy, int dtype, int batch, int h, int w, int cin, int cout, int ksize, int stride,
                           const float* w_oihw, const void* add, void* dx, void* stream) {
  FTC_REQUIRE(dy && w_oihw && dx && dtype_ok(dtype), "bad argument");
  ConvGeom g;
  int rc = conv_geom(&g, batch, h, w, cin, cout, ksize, stride);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t M = (int64_t)batch * h * w;
  FTC_REQUIRE((M + GB - 1) / GB <= 0x7fffffff, "too many pixels");
  dim3 grid((unsigned)((M + GB - 1) / GB), ceil_div(cin, GB));
  if (dtype == DT_F32)
    conv_dgrad_kernel<float><<<grid, GT, 0, s>>>(cp<float>(dy), w_oihw, g, M, cp<float>(add), mp<float>(dx));
  else
    conv_dgrad_kernel<bf16><<<grid, GT, 0, s>>>(cp<bf16>(dy), w_oihw, g, M, cp<bf16>(add), mp<bf16>(dx));
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_dwconv3x3(const void* x, void* y, int dtype, int batch, int h, int w, int c, int stride, const float* w9c,
                        void* stream) {
  FTC_REQUIRE(x && y && w9c && dtype_ok(dtype) && batch > 0 && (stride == 1 || stride == 2), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
  const int64_t total = (int64_t)batch * ho * wo * c;
#ifndef FTC_EMU
  // stride-1 maps up to 48 px wide (78 of the 80 depthwise convs of EfficientNetV2-XL at 768 px): the TMA strip kernel of the
  // inference engine in its plain mode -- each input element crosses L2 -> SM once instead of nine times
  if (stride == 1 && batch <= 65535 && dwconv3x3_se_supported(h, w, c, 1) && (reinterpret_cast<uintptr_t>(x) & 15) == 0)
    return dwconv3x3_raw_strip(x, y, dtype, batch, h, w, c, w9c, 0, s);
#endif
  if (c % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w9c)) & 15) == 0) {
    const dim3 vgrid((unsigned)std::min<int64_t>(((int64_t)batch * ho * wo + 31) / 32, 148 * 8), ceil_div(c, RED_CH));
    if (dtype == DT_F32) dw_vec_kernel<float, false><<<vgrid, 256, 0, s>>>(cp<float>(x), mp<float>(y), batch, h, w, ho, wo, c, stride, w9c);
    else dw_vec_kernel<bf16, false><<<vgrid, 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), batch, h, w, ho, wo, c, stride, w9c);
    FTC_POST_LAUNCH();
    return 0;
  }
  if (dtype == DT_F32)
    dw_fwd_kernel<float><<<ew_grid(total), 256, 0, s>>>(cp<float>(x), mp<float>(y), batch, h, w, c, ho, wo, stride, w9c);
  else
    dw_fwd_kernel<bf16><<<ew_grid(total), 256, 0, s>>>(cp<bf16>(x), mp<bf16>(y), batch, h, w, c, ho, wo, stride, w9c);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_dwconv3x3_dgrad(const void* dy, void* dx, int dtype, int batch, int h, int w, int c, int stride, const float* w9c,
                              void* stream) {
  FTC_REQUIRE(dy && dx && w9c && dtype_ok(dtype) && batch > 0 && (stride == 1 || stride == 2), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
  const int64_t total = (int64_t)batch * h * w * c;
#ifndef FTC_EMU
  // stride 1: the data gradient is the same depthwise convolution with the taps rotated by 180 degrees
  if (stride == 1 && batch <= 65535 && dwconv3x3_se_supported(h, w, c, 1) && (reinterpret_cast<uintptr_t>(dy) & 15) == 0)
    return dwconv3x3_raw_strip(dy, dx, dtype, batch, h, w, c, w9c, 1, s);
#endif
  if (c % 8 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(w9c)) & 15) == 0) {
    const dim3 vgrid((unsigned)std::min<int64_t>(((int64_t)batch * h * w + 31) / 32, 148 * 8), ceil_div(c, RED_CH));
    if (dtype == DT_F32) dw_vec_kernel<float, true><<<vgrid, 256, 0, s>>>(cp<float>(dy), mp<float>(dx), batch, ho, wo, h, w, c, stride, w9c);
    else dw_vec_kernel<bf16, true><<<vgrid, 256, 0, s>>>(cp<bf16>(dy), mp<bf16>(dx), batch, ho, wo, h, w, c, stride, w9c);
    FTC_POST_LAUNCH();
    return 0;
  }
  if (dtype == DT_F32)
    dw_dgrad_kernel<float><<<ew_grid(total), 256, 0, s>>>(cp<float>(dy), mp<float>(dx), batch, h, w, c, ho, wo, stride, w9c);
  else
    dw_dgrad_kernel<bf16><<<ew_grid(total), 256, 0, s>>>(cp<bf16>(dy), mp<bf16>(dx), batch, h, w, c, ho, wo, stride, w9c);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_dwconv3x3_wgrad(const void* x, const void* dy, int dtype, int batch, int h, int w, int c, int stride, float* dw9c,
                              void* stream) {
  FTC_REQUIRE(x && dy && dw9c && dtype_ok(dtype) && batch > 0 && (stride == 1 || stride == 2), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int ho = (h - 1) / stride + 1, wo = (w - 1) / stride + 1;
  FTC_CHECK_CUDA(cudaMemsetAsync(dw9c, 0, (size_t)9 * c * sizeof(float), s));
  const int64_t P = (int64_t)batch * ho * wo;
  static const int env_strip = [] { const char* e = getenv("FTC_DW_WGRAD_STRIP"); return e ? atoi(e) : 1; }();
  if (env_strip && dtype == DT_BF16 && stride == 1 && c % 8 == 0 && h % DWS_R == 0 && w >= 2 && w <= 128 && P < 0x7fffffff &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0) {
    const int slabs = ceil_div(c, RED_CH);
    const int n_strips = batch * (h / DWS_R);
    const size_t smem = ((size_t)(DWS_R + 2) * (w + 2) + (size_t)DWS_R * w) * RED_CH * sizeof(bf16);
    int groups = std::max(1, std::min(n_strips, (148 * 3 + slabs - 1) / slabs));
    const int per_cta = (n_strips + groups - 1) / groups;
    groups = (n_strips + per_cta - 1) / per_cta;
    static bool attr_done = false;
    if (!attr_done) {
      FTC_CHECK_CUDA(cudaFuncSetAttribute(dw_wgrad_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_done = true;
    }
    dw_wgrad_strip_kernel<<<dim3(slabs, groups), 256, std::max<size_t>(smem, 9 * 8 * RED_CH * sizeof(float)), s>>>(cp<bf16>(x), cp<bf16>(dy), batch, h, w, c, per_cta, dw9c);
    FTC_POST_LAUNCH();
    return 0;
  }
  if (c % 8 == 0 && P < 0x7fffffff && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0) {
    const int cb = ceil_div(c, RED_CH);
    int64_t sp = std::max<int64_t>(1, std::min<int64_t>((148 * 6 + cb - 1) / cb, (P + 255) / 256));
    sp = std::min<int64_t>(sp, 65535);
    int64_t per = (P + sp - 1) / sp;
    per = (per + 31) / 32 * 32;
    sp = (P + per - 1) / per;
    dim3 vgrid(cb, (unsigned)sp);
    if (dtype == DT_F32)
      dw_wgrad_vec_kernel<float><<<vgrid, 256, 0, s>>>(cp<float>(x), cp<float>(dy), batch, h, w, c, ho, wo, stride, (int)per, dw9c);
    else
      dw_wgrad_vec_kernel<bf16><<<vgrid, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), batch, h, w, c, ho, wo, stride, (int)per, dw9c);
    FTC_POST_LAUNCH();
    return 0;
  }
  const int cblocks = ceil_div(c, 32);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>((148 * 8 + cblocks - 1) / cblocks, (P + 63) / 64));
  splits = std::min<int64_t>(splits, 65535);
  const int64_t pps = (P + splits - 1) / splits;
  splits = (P + pps - 1) / pps;
  dim3 grid(cblocks, (unsigned)splits);
  if (dtype == DT_F32)
    dw_wgrad_kernel<float><<<grid, 256, 0, s>>>(cp<float>(x), cp<float>(dy), batch, h, w, c, ho, wo, stride, pps, dw9c);
  else
    dw_wgrad_kernel<bf16><<<grid, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(dy), batch, h, w, c, ho, wo, stride, pps, dw9c);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_spatial_sum(const void* x, const void* y, int dtype, int batch, int hw, int c, float scale, float* out,
                          void* stream) {
  FTC_REQUIRE(x && out && dtype_ok(dtype) && batch > 0 && batch <= 65535 && hw > 0 && c > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(ceil_div(c, RED_CH), batch);
  if (c % 8 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    if (dtype == DT_F32) {
      if (y) spatial_sum_vec_kernel<float, 1><<<grid, SSUM_THREADS, 0, s>>>(cp<float>(x), cp<float>(y), hw, c, scale, out);
      else spatial_sum_vec_kernel<float, 0><<<grid, 256, 0, s>>>(cp<float>(x), nullptr, hw, c, scale, out);
    } else {
      if (y) spatial_sum_vec_kernel<bf16, 1><<<grid, SSUM_THREADS, 0, s>>>(cp<bf16>(x), cp<bf16>(y), hw, c, scale, out);
      else spatial_sum_vec_kernel<bf16, 0><<<grid, 256, 0, s>>>(cp<bf16>(x), nullptr, hw, c, scale, out);
    }
    FTC_POST_LAUNCH();
    return 0;
  }
  if (dtype == DT_F32) {
    if (y) spatial_sum_kernel<float, 1><<<grid, RED_THREADS, 0, s>>>(cp<float>(x), cp<float>(y), hw, c, scale, out);
    else spatial_sum_kernel<float, 0><<<grid, RED_THREADS, 0, s>>>(cp<float>(x), nullptr, hw, c, scale, out);
  } else {
    if (y) spatial_sum_kernel<bf16, 1><<<grid, RED_THREADS, 0, s>>>(cp<bf16>(x), cp<bf16>(y), hw, c, scale, out);
    else spatial_sum_kernel<bf16, 0><<<grid, RED_THREADS, 0, s>>>(cp<bf16>(x), nullptr, hw, c, scale, out);
  }
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_scale_bc(const void* x, const float* scale_bc, const float* bias_bc, float bias_mul, void* y, int dtype, int batch,
                       int hw, int c, void* stream) {
  FTC_REQUIRE(x && scale_bc && y && dtype_ok(dtype) && batch > 0 && hw > 0 && c > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t total = (int64_t)batch * hw * c;
  if (c % 8 == 0 && batch <= 65535 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    // pixel walkers per (slab, image): enough CTAs to fill the machine (~6 per SM), each thread >= 4 pixels when the map allows
    const int slabs = ceil_div(c, RED_CH);
    const int want = std::max(1, (148 * 6 + slabs * batch - 1) / (slabs * batch));
    dim3 vgrid(slabs, (unsigned)std::min(ceil_div(hw, 128), want), batch);
    if (dtype == DT_F32) scale_bc_vec_kernel<float><<<vgrid, 256, 0, s>>>(cp<float>(x), scale_bc, mp<float>(y), hw, c, bias_bc, bias_mul);
    else scale_bc_vec_kernel<bf16><<<vgrid, 256, 0, s>>>(cp<bf16>(x), scale_bc, mp<bf16>(y), hw, c, bias_bc, bias_mul);
    FTC_POST_LAUNCH();
    return 0;
  }
  if (dtype == DT_F32)
    scale_bc_kernel<float><<<ew_grid(total), 256, 0, s>>>(cp<float>(x), scale_bc, mp<float>(y), hw, c, total, bias_bc, bias_mul);
  else
    scale_bc_kernel<bf16><<<ew_grid(total), 256, 0, s>>>(cp<bf16>(x), scale_bc, mp<bf16>(y), hw, c, total, bias_bc, bias_mul);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_se_fc(const float* mean, int batch, int c, int sq, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* hid_pre, float* gate, void* stream) {
  FTC_REQUIRE(mean && w1 && b1 && w2 && b2 && hid_pre && gate && batch > 0 && c > 0 && sq > 0 && sq <= 4096, "bad argument");
  FTC_REQUIRE(batch <= 65535, "batch");
  cudaStream_t s = (cudaStream_t)stream;
  se_fc_fwd1_kernel<<<dim3(ceil_div(sq, 8), batch), 256, 0, s>>>(mean, c, sq, w1, b1, hid_pre);
  FTC_POST_LAUNCH();
  se_fc_fwd2_kernel<<<dim3(ceil_div(c, 64), batch), 256, (size_t)sq * sizeof(float), s>>>(hid_pre, c, sq, w2, b2, gate);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_se_fc_bwd(const float* dgate, const float* gate, const float* hid_pre, const float* mean, int batch, int c, int sq,
                        const float* w1, const float* w2, float* dgp, float* dhp, float* dmean, float* dw1, float* db1,
                        float* dw2, float* db2, void* stream) {
  FTC_REQUIRE(dgate && gate && hid_pre && mean && w1 && w2 && dgp && dhp && dmean && dw1 && db1 && dw2 && db2, "null argument");
  FTC_REQUIRE(batch > 0 && batch <= 65535 && c > 0 && sq > 0 && sq <= 4096, "bad geometry");
  cudaStream_t s = (cudaStream_t)stream;
  se_fc_bwd1_kernel<<<dim3(ceil_div(sq, 32), batch), SE_BWD1_THREADS, 0, s>>>(dgate, gate, hid_pre, c, sq, w2, dgp, dhp);
  FTC_POST_LAUNCH();
  se_fc_bwd2_kernel<<<dim3(ceil_div(c, 256), batch), 256, (size_t)sq * sizeof(float), s>>>(dhp, c, sq, w1, dmean);
  FTC_POST_LAUNCH();
  const int64_t n = (int64_t)c * sq;
  se_fc_bwd_weight_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dgp, dhp, hid_pre, mean, batch, c, sq, dw1, db1, dw2, db2);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_upsample2x_bwd(const void* dy, void* dx, int dtype, int batch, int h, int w, int c, void* stream) {
  return ftc_train_upsample2x_bwd_ld(dy, c, dx, dtype, batch, h, w, c, stream);
}

int ftc_train_upsample2x_bwd_ld(const void* dy, int64_t dy_ld, void* dx, int dtype, int batch, int h, int w, int c, void* stream) {
  FTC_REQUIRE(dy && dx && dtype_ok(dtype) && batch > 0 && batch <= 65535 && h > 0 && w > 0 && c > 0, "bad argument");
  FTC_REQUIRE(dy_ld >= c, "dy pixel stride below the channel count");
  const int64_t ldy = dy_ld;
  cudaStream_t s = (cudaStream_t)stream;
  dim3 grid(h, batch);
  const float sy = (float)(h - 1) / (float)(2 * h - 1), sx = (float)(w - 1) / (float)(2 * w - 1);
  const bool vec = c % 8 == 0 && ldy % 8 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0;
  if (dtype == DT_F32) {
    if (vec) upsample2x_bwd_kernel<float, true><<<grid, 256, 0, s>>>(cp<float>(dy), mp<float>(dx), h, w, c, sy, sx, ldy);
    else upsample2x_bwd_kernel<float, false><<<grid, 256, 0, s>>>(cp<float>(dy), mp<float>(dx), h, w, c, sy, sx, ldy);
  } else {
    if (vec) upsample2x_bwd_kernel<bf16, true><<<grid, 256, 0, s>>>(cp<bf16>(dy), mp<bf16>(dx), h, w, c, sy, sx, ldy);
    else upsample2x_bwd_kernel<bf16, false><<<grid, 256, 0, s>>>(cp<bf16>(dy), mp<bf16>(dx), h, w, c, sy, sx, ldy);
  }
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_layernorm(const void* x, const void* r1, const void* r2, void* xs, void* y, float* mean, float* rstd, int dtype,
                        int64_t rows, int d, const float* gamma, const float* beta, float eps, void* stream) {
  FTC_REQUIRE(x && y && mean && rstd && gamma && beta && rows > 0 && d > 0 && dtype_ok(dtype), "bad argument");
  FTC_REQUIRE((r1 == nullptr && r2 == nullptr) || xs != nullptr, "xs is required when residuals are given");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows * 32 + 255) / 256);
  if (dtype == DT_F32)
    ln_fwd_kernel<float><<<grid, 256, 0, s>>>(cp<float>(x), cp<float>(r1), cp<float>(r2), mp<float>(xs), mp<float>(y), mean, rstd, rows, d,
                                            gamma, beta, eps);
  else
    ln_fwd_kernel<bf16><<<grid, 256, 0, s>>>(cp<bf16>(x), cp<bf16>(r1), cp<bf16>(r2), mp<bf16>(xs), mp<bf16>(y), mean, rstd, rows, d, gamma,
                                           beta, eps);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_layernorm_bwd(const void* xs, const void* dy, void* dx, const float* mean, const float* rstd, int dtype, int64_t rows,
                            int d, const float* gamma, float* dgamma, float* dbeta, void* scratch, void* stream) {
  FTC_REQUIRE(xs && dy && dx && mean && rstd && gamma && dgamma && dbeta && scratch && rows > 0 && d > 0 && dtype_ok(dtype),
              "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((rows * 32 + 255) / 256);
  const int nchunk = red_chunks(rows);
  const int64_t rpc = (rows + nchunk - 1) / nchunk;
  dim3 rgrid(ceil_div(d, RED_CH), nchunk);
  if (dtype == DT_F32) {
    ln_bwd_kernel<float><<<grid, 256, 0, s>>>(cp<float>(xs), cp<float>(dy), mp<float>(dx), mean, rstd, rows, d, gamma);
    FTC_POST_LAUNCH();
    ln_col_reduce_kernel<float><<<rgrid, RED_THREADS, 0, s>>>(cp<float>(xs), cp<float>(dy), mean, rstd, rows, d, rpc, (float*)scratch);
  } else {
    ln_bwd_kernel<bf16><<<grid, 256, 0, s>>>(cp<bf16>(xs), cp<bf16>(dy), mp<bf16>(dx), mean, rstd, rows, d, gamma);
    FTC_POST_LAUNCH();
    ln_col_reduce_kernel<bf16><<<rgrid, RED_THREADS, 0, s>>>(cp<bf16>(xs), cp<bf16>(dy), mean, rstd, rows, d, rpc, (float*)scratch);
  }
  FTC_POST_LAUNCH();
  col_reduce_finish_kernel<1><<<ceil_div(d, 32), 1024, 0, s>>>((const float*)scratch, nchunk, d, rows, dbeta, dgamma, RunningStats{nullptr, nullptr, nullptr, 0.f});
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_swiglu(const void* x1, const void* xg, void* h, int dtype, int64_t total, void* stream) {
  FTC_REQUIRE(x1 && xg && h && total > 0 && dtype_ok(dtype), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DT_F32) swiglu_fwd_kernel<float><<<ew_grid(total), 256, 0, s>>>(cp<float>(x1), cp<float>(xg), mp<float>(h), total);
  else swiglu_fwd_kernel<bf16><<<ew_grid(total), 256, 0, s>>>(cp<bf16>(x1), cp<bf16>(xg), mp<bf16>(h), total);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_swiglu_bwd(const void* x1, const void* xg, const void* dh, void* dx1, void* dxg, int dtype, int64_t total,
                         void* stream) {
  FTC_REQUIRE(x1 && xg && dh && dx1 && dxg && total > 0 && dtype_ok(dtype), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DT_F32)
    swiglu_bwd_kernel<float><<<ew_grid(total), 256, 0, s>>>(cp<float>(x1), cp<float>(xg), cp<float>(dh), mp<float>(dx1), mp<float>(dxg), total);
  else
    swiglu_bwd_kernel<bf16><<<ew_grid(total), 256, 0, s>>>(cp<bf16>(x1), cp<bf16>(xg), cp<bf16>(dh), mp<bf16>(dx1), mp<bf16>(dxg), total);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_embed3(const int64_t* tokens, const float* e0, const float* e1, const float* e2, int m0, int m1, int m2, void* out,
                     int dtype, int64_t rows, int d, void* stream) {
  FTC_REQUIRE(tokens && e0 && e1 && e2 && out && rows > 0 && d > 0 && m0 > 0 && m1 > 0 && m2 > 0 && dtype_ok(dtype), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  Embed3 tb = {{e0, e1, e2}, {nullptr, nullptr, nullptr}, {m0, m1, m2}};
  if (dtype == DT_F32) embed3_fwd_kernel<float><<<ew_grid(rows * d), 256, 0, s>>>(tokens, tb, mp<float>(out), rows, d);
  else embed3_fwd_kernel<bf16><<<ew_grid(rows * d), 256, 0, s>>>(tokens, tb, mp<bf16>(out), rows, d);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_train_embed3_bwd(const int64_t* tokens, const void* dy, int dtype, int64_t rows, int d, int m0, int m1, int m2, float* de0,
                         float* de1, float* de2, void* stream) {
  FTC_REQUIRE(tokens && dy && de0 && de1 && de2 && rows > 0 && d > 0 && m0 > 0 && m1 > 0 && m2 > 0 && dtype_ok(dtype), "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  FTC_CHECK_CUDA(cudaMemsetAsync(de0, 0, (size_t)m0 * d * sizeof(float), s));
  FTC_CHECK_CUDA(cudaMemsetAsync(de1, 0, (size_t)m1 * d * sizeof(float), s));
  FTC_CHECK_CUDA(cudaMemsetAsync(de2, 0, (size_t)m2 * d * sizeof(float), s));
  Embed3 tb = {{nullptr, nullptr, nullptr}, {de0, de1, de2}, {m0, m1, m2}};
  if (dtype == DT_F32) embed3_bwd_kernel<float><<<ew_grid(rows * d), 256, 0, s>>>(tokens, tb, cp<float>(dy), rows, d);
  else embed3_bwd_kernel<bf16><<<ew_grid(rows * d), 256, 0, s>>>(tokens, tb, cp<bf16>(dy), rows, d);
  FTC_POST_LAUNCH();
  return 0;
}

size_t ftc_train_attention_bwd_scratch_bytes(int batch, int heads, int lt, int ls) {
  return (size_t)2 * batch * heads * lt * ls * sizeof(float);
}

int ftc_train_attention_bwd(const void* q, const void* k, const void* v, const float* mask, const void* dout, float* dq, float* dk,
                            float* dv, void* scratch, int dtype, int batch, int heads, int hd, int lt, int ls, void* stream) {
  FTC_REQUIRE(q && k && v && dout && dq && dk && dv && scratch && dtype_ok(dtype), "bad argument");
  FTC_REQUIRE(batch > 0 && batch <= 65535 && heads > 0 && heads <= 65535 && hd > 0 && hd <= ATT_MAX_HD && lt > 0 && ls > 0, "bad geometry");
  const size_t smem = (size_t)4 * (2 * hd + 2 * ls) * sizeof(float);
  FTC_REQUIRE(smem <= 200 * 1024, "key sequence too long for the per-warp staging");
  cudaStream_t s = (cudaStream_t)stream;
  const int D = heads * hd;
  const float scale = 1.0f / sqrtf((float)hd);
  {   // short sequences: one CTA per (batch, head) with everything in shared memory
    const size_t es = dtype == DT_F32 ? 4 : 2;
    const size_t fused = (size_t)2 * lt * (ls + 1) * sizeof(float) + (size_t)2 * (lt + ls) * (hd + 2) * es;
    if (lt <= 128 && ls <= 128 && hd % 2 == 0 && fused <= 200 * 1024) {
      dim3 gf(heads, batch);
      if (dtype == DT_F32) {
        static bool done = false;
        if (!done) { FTC_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
        attn_bwd_fused_kernel<float><<<gf, 256, fused, s>>>(cp<float>(q), cp<float>(k), cp<float>(v), mask, cp<float>(dout), dq, dk, dv, hd, lt, ls, D, scale);
      } else {
        static bool done = false;
        if (!done) { FTC_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_fused_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); done = true; }
        attn_bwd_fused_kernel<bf16><<<gf, 256, fused, s>>>(cp<bf16>(q), cp<bf16>(k), cp<bf16>(v), mask, cp<bf16>(dout), dq, dk, dv, hd, lt, ls, D, scale);
      }
      FTC_POST_LAUNCH();
      return 0;
    }
  }
  float* P = (float*)scratch;
  float* dS = P + (size_t)batch * heads * lt * ls;
  dim3 ga(ceil_div(lt, 4), heads, batch), gb(ceil_div(ls, 4), heads, batch);
  if (dtype == DT_F32) {
    if (smem > 48 * 1024) FTC_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_rows_kernel<float><<<ga, 128, smem, s>>>(cp<float>(q), cp<float>(k), cp<float>(v), mask, cp<float>(dout), P, dS, dq, heads, hd, lt,
                                                     ls, D, scale);
    FTC_POST_LAUNCH();
    attn_bwd_cols_kernel<float><<<gb, 128, 0, s>>>(cp<float>(q), cp<float>(dout), P, dS, dk, dv, heads, hd, lt, ls, D, scale);
  } else {
    if (smem > 48 * 1024) FTC_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_rows_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_rows_kernel<bf16><<<ga, 128, smem, s>>>(cp<bf16>(q), cp<bf16>(k), cp<bf16>(v), mask, cp<bf16>(dout), P, dS, dq, heads, hd, lt, ls,
                                                    D, scale);
    FTC_POST_LAUNCH();
    attn_bwd_cols_kernel<bf16><<<gb, 128, 0, s>>>(cp<bf16>(q), cp<bf16>(dout), P, dS, dk, dv, heads, hd, lt, ls, D, scale);
  }
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_page_maps(const float* heat9, int batch, int h, int w, const int* tile_meta, float* page, int page_h4, int page_w4, int scale,
                  void* stream) {
  FTC_REQUIRE(heat9 && tile_meta && page && batch > 0 && h > 0 && w > 0 && page_h4 > 0 && page_w4 > 0 && scale > 0, "bad argument");
  page_maps_kernel<<<ew_grid((int64_t)batch * 7 * h * w), 256, 0, (cudaStream_t)stream>>>(heat9, batch, h, w, tile_meta, (int*)page, page_h4,
                                                                                         page_w4, scale);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_debug_set_bn_unroll(int u) {
  g_bn_unroll = u;
  return 0;
}

int ftc_debug_set_wgrad_mma(int on) {
  g_wgrad_mma = on < 0 ? -1 : (on != 0);
  return 0;
}

}  // extern "C"
