// Training losses of the hot path's two models (reference loss_func.py), fused:
//   heatmap_loss_sums  : ONE pass over heatmap [B,9,H,W] / labelmap [B,5,H,W] / idmap [B,2,H,W] producing the ten partial sums
//                        behind keymap (focal, :74-92), size (weighted Huber, :111-115), textline / separator BCE (:117-118)
//                        and the four weighted code BCEs (:120-126); per-CTA double partials + a one-CTA finish: deterministic.
//   heatmap_loss_grad  : d(sum_i alpha_i * loss_i) / d heatmap in one more pass (needs weight1_count from the first).
//   ce_rows            : softmax cross-entropy of one row per warp for the three residue heads, weighted sums, argmax hits
//                        (loss_function :128-161 at the fmask pixels, loss_function3 :179-213 on [B, L, m] logits).
// fp32 math as in the reference (it casts the focal loss to fp32 explicitly, :79); sums accumulate in double.
#include "../../include/ftc_b200.h"
#include "common.cuh"
#include <math.h>

namespace ftc {
namespace {

constexpr int HL_SUMS = 10;   // 0 keymap, 1 size numerator, 2 weight1 sum, 3 textline, 4 separator, 5-8 code1/2/4/8, 9 unused
constexpr float KEY_TH1 = 0.85f, KEY_TH2 = 0.85f;

__device__ __forceinline__ float softplus_f(float x) {           // torch softplus (beta 1, threshold 20)
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float logsigmoid_f(float x) {         // torch: min(x, 0) - log1p(exp(-|x|))
  return fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float bce_logits_f(float x, float y) { // torch: (1 - y) * x + max(-x, 0) + log1p(exp(-|x|))
  return (1.f - y) * x + fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float huber_f(float a, float b) {     // delta = 1
  const float d = fabsf(a - b);
  return d < 1.f ? 0.5f * d * d : d - 0.5f;
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

struct PixelTerms { float v[HL_SUMS]; };

__device__ __forceinline__ void pixel_terms(const float* __restrict__ heat, const float* __restrict__ label,
                                            const int64_t* __restrict__ idmap, int64_t b, int64_t r, int64_t hw, PixelTerms& t) {
  const float* hp = heat + b * 9 * hw + r;
  const float* lp = label + b * 5 * hw + r;
  const float key = lp[0];
  const float x0 = hp[0];
  // focal loss (heatmap_loss): alpha 2, beta 4, positive iff true >= 1
  const float pred = sigmoid_f(x0);
  float kl;
  if (key >= 1.0f) kl = -logsigmoid_f(x0) * (1.f - pred) * (1.f - pred);
  else {
    const float nw = (1.f - key) * (1.f - key);
    kl = (x0 + softplus_f(-x0)) * pred * pred * (nw * nw);
  }
  t.v[0] = kl;
  const float w1 = fmaxf(key - KEY_TH1, 0.f) / (1.f - KEY_TH1);
  if (key > KEY_TH1) {
    t.v[1] = (huber_f(hp[hw], lp[hw]) + huber_f(hp[2 * hw], lp[2 * hw])) * w1;
    t.v[2] = w1;
  } else { t.v[1] = 0.f; t.v[2] = 0.f; }
  t.v[3] = bce_logits_f(hp[3 * hw], lp[3 * hw]);
  t.v[4] = bce_logits_f(hp[4 * hw], lp[4 * hw]);
  const float w2 = fmaxf(key - KEY_TH2, 0.f) / (1.f - KEY_TH2);
  const int64_t bits = idmap[b * 2 * hw + hw + r];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float y = (bits & (int64_t(1) << i)) ? 1.f : 0.f;
    const float w = 1.f + y * w2 + w2;
    t.v[5 + i] = w * bce_logits_f(hp[(5 + i) * hw], y);
  }
  t.v[9] = 0.f;
}

__global__ void __launch_bounds__(256) heatmap_loss_sums_kernel(const float* __restrict__ heat, const float* __restrict__ label,
                                                                const int64_t* __restrict__ idmap, int64_t B, int64_t hw,
                                                                double* __restrict__ partial) {
  double acc[HL_SUMS];
#pragma unroll
  for (int i = 0; i < HL_SUMS; ++i) acc[i] = 0.0;
  const int64_t total = B * hw;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    PixelTerms t;
    pixel_terms(heat, label, idmap, idx / hw, idx % hw, hw, t);
#pragma unroll
    for (int i = 0; i < HL_SUMS; ++i) acc[i] += (double)t.v[i];
  }
  __shared__ double sh[8][HL_SUMS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < HL_SUMS; ++i) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < HL_SUMS) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
    partial[(int64_t)blockIdx.x * HL_SUMS + threadIdx.x] = v;
  }
}

// losses[0..8] = keymap*10, size, textline, separator, code1, code2, code4, code8, weight1_count (fp32)
__global__ void heatmap_loss_finish_kernel(const double* __restrict__ partial, int nblocks, double n_pix, float* __restrict__ losses) {
  __shared__ double s[HL_SUMS];
  if (threadIdx.x < HL_SUMS) {
    double v = 0.0;
    for (int b = 0; b < nblocks; ++b) v += partial[(int64_t)b * HL_SUMS + threadIdx.x];
    s[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double cnt = s[2] > 1.0 ? s[2] : 1.0;           // torch.maximum(1, weight1.sum())
    losses[0] = (float)(s[0] / n_pix * 10.0);
    losses[1] = (float)(s[1] / cnt);
    losses[2] = (float)(s[3] / n_pix);
    losses[3] = (float)(s[4] / n_pix);
    for (int i = 0; i < 4; ++i) losses[4 + i] = (float)(s[5 + i] / n_pix);
    losses[8] = (float)cnt;
  }
}

// grad[b, c, y, x] = d( sum_i alpha[i] * loss_i ) / d heatmap[b, c, y, x]; alpha: 8 weights in the order of `losses`
__global__ void __launch_bounds__(256) heatmap_loss_grad_kernel(const float* __restrict__ heat, const float* __restrict__ label,
                                                                const int64_t* __restrict__ idmap, int64_t B, int64_t hw,
                                                                const float* __restrict__ alpha, const float* __restrict__ losses,
                                                                float* __restrict__ grad) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * hw) return;
  const int64_t b = idx / hw, r = idx % hw;
  const float* hp = heat + b * 9 * hw + r;
  const float* lp = label + b * 5 * hw + r;
  float* gp = grad + b * 9 * hw + r;
  const float inv_n = 1.f / (float)(B * hw);
  const float key = lp[0];
  {
    const float x = hp[0], p = sigmoid_f(x);
    float g;
    if (key >= 1.0f) {
      // d/dx [ softplus(-x) (1-p)^2 ] = -(1-p)^3 - 2 softplus(-x) (1-p)^2 p
      const float q = 1.f - p;
      g = -q * q * q - 2.f * softplus_f(-x) * q * q * p;
    } else {
      // d/dx [ softplus(x) p^2 ] * nw ; (x + softplus(-x) == softplus(x))
      const float nw = (1.f - key) * (1.f - key);
      g = (p * p * p + 2.f * (x + softplus_f(-x)) * p * p * (1.f - p)) * (nw * nw);
    }
    gp[0] = alpha[0] * 10.f * inv_n * g;
  }
  {
    float g1 = 0.f, g2 = 0.f;
    if (key > KEY_TH1) {
      const float w1 = fmaxf(key - KEY_TH1, 0.f) / (1.f - KEY_TH1) / losses[8];
      const float d1 = hp[hw] - lp[hw], d2 = hp[2 * hw] - lp[2 * hw];
      g1 = fminf(fmaxf(d1, -1.f), 1.f) * w1;
      g2 = fminf(fmaxf(d2, -1.f), 1.f) * w1;
    }
    gp[hw] = alpha[1] * g1;
    gp[2 * hw] = alpha[1] * g2;
  }
  gp[3 * hw] = alpha[2] * inv_n * (sigmoid_f(hp[3 * hw]) - lp[3 * hw]);
  gp[4 * hw] = alpha[3] * inv_n * (sigmoid_f(hp[4 * hw]) - lp[4 * hw]);
  const float w2 = fmaxf(key - KEY_TH2, 0.f) / (1.f - KEY_TH2);
  const int64_t bits = idmap[b * 2 * hw + hw + r];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float y = (bits & (int64_t(1) << i)) ? 1.f : 0.f;
    const float w = 1.f + y * w2 + w2;
    gp[(5 + i) * hw] = alpha[4 + i] * inv_n * w * (sigmoid_f(hp[(5 + i) * hw]) - y);
  }
}

// one warp per row: ce[g] = logsumexp(logits_g[row, :m_g]) - logits_g[row, target % m_g]; hit[g] = argmax == target % m_g
//   out[0] += sum_g w_row * ce[g]  (w_row = weight[row] if select[row] else 0), out[1] += w_row, out[2] += all three hit
//   (counted where count_sel[row]), out[3] += count_sel[row]
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ l0, const float* __restrict__ l1,
                                                      const float* __restrict__ l2, int ld0, int ld1, int ld2, int m0, int m1, int m2,
                                                      const int64_t* __restrict__ target, const float* __restrict__ weight,
                                                      const unsigned char* __restrict__ select,
                                                      const unsigned char* __restrict__ count_sel, int rows,
                                                      double* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bool sel = select == nullptr || select[row] != 0, csel = count_sel == nullptr || count_sel[row] != 0;
  if (!sel && !csel) return;
  const int64_t tgt = target[row];
  const float* lp[3] = {l0 + (int64_t)row * ld0, l1 + (int64_t)row * ld1, l2 + (int64_t)row * ld2};
  const int mm[3] = {m0, m1, m2};
  float ce_sum = 0.f;
  int hits = 0;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int m = mm[g];
    float mx = -INFINITY;
    int am = 0;
    for (int j = lane; j < m; j += 32) { const float v = lp[g][j]; if (v > mx) { mx = v; am = j; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oi = __shfl_xor_sync(0xffffffffu, am, o);
      if (ov > mx || (ov == mx && oi < am)) { mx = ov; am = oi; }     // first maximum, as torch.argmax
    }
    float s = 0.f;
    for (int j = lane; j < m; j += 32) s += expf(lp[g][j] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int t = (int)(((tgt % m) + m) % m);
    ce_sum += (logf(s) + mx) - lp[g][t];
    hits += (am == t) ? 1 : 0;
  }
  if (lane == 0) {
    if (sel) {
      const float w = weight ? weight[row] : 1.f;
      atomicAdd(&out[0], (double)(ce_sum * w));
      atomicAdd(&out[1], (double)w);
    }
    if (csel) {
      if (hits == 3) atomicAdd(&out[2], 1.0);
      atomicAdd(&out[3], 1.0);
    }
  }
}


// backward of ce_rows' out[0] / max(out[1], 1) (the id_loss of loss_function :128-161, the loss of loss_function3): one warp per
// row, d logits_g[row, j] = coef * w_row * (softmax_g(row)[j] - [j == target % m_g]) on the selected rows, 0 elsewhere;
// coef (device scalar) = upstream gradient / max(sum of weights, 1)
__global__ void __launch_bounds__(256) ce_rows_grad_kernel(const float* __restrict__ l0, const float* __restrict__ l1,
                                                           const float* __restrict__ l2, int ld0, int ld1, int ld2, int m0, int m1,
                                                           int m2, const int64_t* __restrict__ target,
                                                           const float* __restrict__ weight, const unsigned char* __restrict__ select,
                                                           int rows, const float* __restrict__ coef, float* __restrict__ g0,
                                                           float* __restrict__ g1, float* __restrict__ g2) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bool sel = select == nullptr || select[row] != 0;
  const float cw = sel ? coef[0] * (weight ? weight[row] : 1.f) : 0.f;
  const int64_t tgt = target[row];
  const float* lp[3] = {l0 + (int64_t)row * ld0, l1 + (int64_t)row * ld1, l2 + (int64_t)row * ld2};
  float* gp[3] = {g0 + (int64_t)row * m0, g1 + (int64_t)row * m1, g2 + (int64_t)row * m2};
  const int mm[3] = {m0, m1, m2};
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const int m = mm[g];
    if (cw == 0.f) {
      for (int j = lane; j < m; j += 32) gp[g][j] = 0.f;
      continue;
    }
    float mx = -INFINITY;
    for (int j = lane; j < m; j += 32) mx = fmaxf(mx, lp[g][j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int j = lane; j < m; j += 32) s += expf(lp[g][j] - mx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.f / s;
    const int t = (int)(((tgt % m) + m) % m);
    for (int j = lane; j < m; j += 32) gp[g][j] = cw * (expf(lp[g][j] - mx) * inv - (j == t ? 1.f : 0.f));
  }
}

}  // namespace
}  // namespace ftc

using namespace ftc;

extern "C" {

size_t ftc_heatmap_loss_scratch_bytes(void) { return (size_t)1024 * HL_SUMS * sizeof(double); }

int ftc_heatmap_loss(const float* heatmap, const float* labelmap, const int64_t* idmap, int batch, int h, int w, float* losses9,
                     void* scratch, void* stream) {
  FTC_REQUIRE(heatmap && labelmap && idmap && losses9 && scratch && batch > 0 && h > 0 && w > 0, "bad argument");
  const int64_t hw = (int64_t)h * w, total = hw * batch;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  cudaStream_t s = (cudaStream_t)stream;
  heatmap_loss_sums_kernel<<<blocks, 256, 0, s>>>(heatmap, labelmap, idmap, batch, hw, (double*)scratch);
  FTC_POST_LAUNCH();
  heatmap_loss_finish_kernel<<<1, 32, 0, s>>>((const double*)scratch, blocks, (double)total, losses9);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_heatmap_loss_grad(const float* heatmap, const float* labelmap, const int64_t* idmap, int batch, int h, int w,
                          const float* alpha8, const float* losses9, float* grad, void* stream) {
  FTC_REQUIRE(heatmap && labelmap && idmap && alpha8 && losses9 && grad && batch > 0, "bad argument");
  const int64_t hw = (int64_t)h * w, total = hw * batch;
  heatmap_loss_grad_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(heatmap, labelmap, idmap, batch, hw, alpha8,
                                                                                         losses9, grad);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_ce_rows(const float* logits0, const float* logits1, const float* logits2, int ld0, int ld1, int ld2, int m0, int m1, int m2,
                const int64_t* target, const float* weight, const unsigned char* select, const unsigned char* count_select, int rows,
                double* out4, void* stream) {
  FTC_REQUIRE(logits0 && logits1 && logits2 && target && out4 && rows >= 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  FTC_CHECK_CUDA(cudaMemsetAsync(out4, 0, 4 * sizeof(double), s));
  if (rows == 0) return 0;
  ce_rows_kernel<<<(rows + 7) / 8, 256, 0, s>>>(logits0, logits1, logits2, ld0, ld1, ld2, m0, m1, m2, target, weight, select, count_select,
                                                rows, out4);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_ce_rows_grad(const float* logits0, const float* logits1, const float* logits2, int ld0, int ld1, int ld2, int m0, int m1,
                     int m2, const int64_t* target, const float* weight, const unsigned char* select, int rows, const float* coef,
                     float* grad0, float* grad1, float* grad2, void* stream) {
  FTC_REQUIRE(logits0 && logits1 && logits2 && target && coef && grad0 && grad1 && grad2 && rows >= 0, "bad argument");
  if (rows == 0) return 0;
  ce_rows_grad_kernel<<<(int)(((int64_t)rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      logits0, logits1, logits2, ld0, ld1, ld2, m0, m1, m2, target, weight, select, rows, coef, grad0, grad1, grad2);
  FTC_POST_LAUNCH();
  return 0;
}

}  // extern "C"
