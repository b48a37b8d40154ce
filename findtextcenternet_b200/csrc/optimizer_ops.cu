// Fused multi-tensor Schedule-Free AdamW step (models/adamw_schedulefree.py:157-184, the foreach branch) in ONE launch:
// per element 4 reads (y, grad, exp_avg_sq, z) + 4 writes instead of ~12 foreach passes over 262 M parameters.
#include "../../include/ftc_b200.h"
#include "common.cuh"

namespace ftc {

struct SfChunk { int tensor; int pad; int64_t offset; };   // one CTA handles elements [offset, offset + CHUNK) of `tensor`

namespace {
constexpr int SF_CHUNK = 8192;

__global__ void __launch_bounds__(256) adamw_sf_step_kernel(const SfChunk* __restrict__ chunks, float* const* __restrict__ ys,
                                                            float* const* __restrict__ grads, float* const* __restrict__ vs,
                                                            float* const* __restrict__ zs, const int64_t* __restrict__ numels,
                                                            float beta2, float one_m_beta2, float bias_correction2, float eps,
                                                            float decay, float lr, float ckp1, float y_alpha, int normalize) {
  const SfChunk c = chunks[blockIdx.x];
  float* y = ys[c.tensor] + c.offset;
  float* g = grads[c.tensor] + c.offset;
  float* v = vs[c.tensor] + c.offset;
  float* z = zs[c.tensor] + c.offset;
  const int64_t left = numels[c.tensor] - c.offset;
  const int n = left < SF_CHUNK ? (int)left : SF_CHUNK;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i];
    float vi = v[i] * beta2;                       // _foreach_mul_(exp_avg_sq, beta2)
    vi = vi + one_m_beta2 * (gi * gi);             // _foreach_addcmul_(exp_avg_sq, grad, grad, value=1-beta2)
    v[i] = vi;
    // RAdam schedule-free skips the Adam normalisation while the SMA is too short (rho_t <= 4): plain (or silent) SGD phase
    float gn = gi;
    if (normalize) gn = gi / (sqrtf(vi / bias_correction2) + eps);   // grad is normalised IN PLACE in the reference
    float yi = y[i];
    if (decay != 0.0f) gn = gn + decay * yi;
    g[i] = gn;
    const float zi = z[i];
    // torch lerp: weight < 0.5 ? start + weight*(end-start) : end - (end-start)*(1-weight)
    const float diff = zi - yi;
    yi = ckp1 < 0.5f ? yi + ckp1 * diff : zi - diff * (1.0f - ckp1);
    yi = yi + y_alpha * gn;
    y[i] = yi;
    z[i] = zi - lr * gn;
  }
}
// ---- graph-replayable variant: the step-dependent scalars live on the device ------------------------------------------
// A CUDA graph bakes kernel arguments in, so a captured optimizer step cannot take lr / bias_correction2 / c_{k+1} by value.
// Here one thread advances the schedule state {k, lr_max, weight_sum} (doubles, as the reference's Python floats,
// models/adamw_schedulefree.py:121-140) and writes the eight fp32 scalars of this step; the update kernel reads them.
//   consts: lr, beta1, beta2, eps, weight_decay, warmup_steps, r, weight_lr_power      state: k, lr_max, weight_sum
__global__ void adamw_sf_schedule_kernel(const double* __restrict__ consts, double* __restrict__ state, float* __restrict__ hyper) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double lr0 = consts[0], beta1 = consts[1], beta2 = consts[2], eps = consts[3], decay = consts[4], warmup = consts[5],
               r = consts[6], wlp = consts[7];
  const double k = state[0];
  const double sched = k < warmup ? (k + 1.0) / warmup : 1.0;
  const double bc2 = 1.0 - pow(beta2, k + 1.0);
  const double lr = lr0 * sched;
  const double lr_max = fmax(lr, state[1]);
  const double weight = pow(k + 1.0, r) * pow(lr_max, wlp);
  const double weight_sum = state[2] + weight;
  const double ckp1 = weight_sum != 0.0 ? weight / weight_sum : 0.0;
  state[0] = k + 1.0; state[1] = lr_max; state[2] = weight_sum;
  hyper[0] = (float)beta2; hyper[1] = (float)(1.0 - beta2); hyper[2] = (float)bc2; hyper[3] = (float)eps;
  hyper[4] = (float)decay; hyper[5] = (float)lr; hyper[6] = (float)ckp1; hyper[7] = (float)(lr * (beta1 * (1.0 - ckp1) - 1.0));
  hyper[8] = 1.0f;       // Adam normalisation on
}

// Schedule-Free RAdam (models/radam_schedulefree.py:138-152): the rectification term of the variance, evaluated on the device in double
// as the reference evaluates it in Python floats.  consts: lr, beta1, beta2, eps, weight_decay, silent_sgd_phase (0 / 1), r,
// weight_lr_power; state: k, lr_max, weight_sum
__global__ void radam_sf_schedule_kernel(const double* __restrict__ consts, double* __restrict__ state, float* __restrict__ hyper) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double lr0 = consts[0], beta1 = consts[1], beta2 = consts[2], eps = consts[3], decay = consts[4], silent = consts[5],
               r = consts[6], wlp = consts[7];
  const double k = state[0], step = k + 1.0;
  const double beta2_t = pow(beta2, step);
  const double bc2 = 1.0 - beta2_t;
  const double rho_inf = 2.0 / (1.0 - beta2) - 1.0;
  const double rho_t = rho_inf - 2.0 * step * beta2_t / bc2;
  const bool adam = rho_t > 4.0;
  const double rect = adam ? sqrt((rho_t - 4.0) * (rho_t - 2.0) * rho_inf / ((rho_inf - 4.0) * (rho_inf - 2.0) * rho_t))
                           : (silent != 0.0 ? 0.0 : 1.0);
  const double lr = lr0 * rect;
  const double lr_max = fmax(lr, state[1]);
  const double weight = pow(step, r) * pow(lr_max, wlp);
  const double weight_sum = state[2] + weight;
  const double ckp1 = weight_sum != 0.0 ? weight / weight_sum : 0.0;
  state[0] = step; state[1] = lr_max; state[2] = weight_sum;
  hyper[0] = (float)beta2; hyper[1] = (float)(1.0 - beta2); hyper[2] = (float)bc2; hyper[3] = (float)eps;
  hyper[4] = (float)decay; hyper[5] = (float)lr; hyper[6] = (float)ckp1; hyper[7] = (float)(lr * (beta1 * (1.0 - ckp1) - 1.0));
  hyper[8] = adam ? 1.0f : 0.0f;
}

__global__ void __launch_bounds__(256) adamw_sf_step_dev_kernel(const SfChunk* __restrict__ chunks, float* const* __restrict__ ys,
                                                                float* const* __restrict__ grads, float* const* __restrict__ vs,
                                                                float* const* __restrict__ zs, const int64_t* __restrict__ numels,
                                                                const float* __restrict__ hyper) {
  const float beta2 = hyper[0], one_m_beta2 = hyper[1], bias_correction2 = hyper[2], eps = hyper[3], decay = hyper[4], lr = hyper[5],
              ckp1 = hyper[6], y_alpha = hyper[7];
  const bool normalize = hyper[8] != 0.0f;
  const SfChunk c = chunks[blockIdx.x];
  float* y = ys[c.tensor] + c.offset;
  float* g = grads[c.tensor] + c.offset;
  float* v = vs[c.tensor] + c.offset;
  float* z = zs[c.tensor] + c.offset;
  const int64_t left = numels[c.tensor] - c.offset;
  const int n = left < SF_CHUNK ? (int)left : SF_CHUNK;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {       // same arithmetic as adamw_sf_step_kernel (normalize = 1)
    const float gi = g[i];
    float vi = v[i] * beta2;
    vi = vi + one_m_beta2 * (gi * gi);
    v[i] = vi;
    float gn = gi;
    if (normalize) gn = gi / (sqrtf(vi / bias_correction2) + eps);
    float yi = y[i];
    if (decay != 0.0f) gn = gn + decay * yi;
    g[i] = gn;
    const float zi = z[i];
    const float diff = zi - yi;
    yi = ckp1 < 0.5f ? yi + ckp1 * diff : zi - diff * (1.0f - ckp1);
    yi = yi + y_alpha * gn;
    y[i] = yi;
    z[i] = zi - lr * gn;
  }
}
}  // namespace
}  // namespace ftc

using namespace ftc;

extern "C" {

int ftc_adamw_sf_chunk_elems(void) { return SF_CHUNK; }

int ftc_adamw_sf_step(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                      const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, double beta1, double beta2,
                      double bias_correction2, double eps, double weight_decay, double lr, double ckp1, void* stream) {
  FTC_REQUIRE(n_chunks >= 0 && chunks && ys && grads && exp_avg_sqs && zs && numels, "bad argument");
  if (n_chunks == 0) return 0;
  adamw_sf_step_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(
      (const SfChunk*)chunks, (float* const*)ys, (float* const*)grads, (float* const*)exp_avg_sqs, (float* const*)zs, numels,
      // scalars arrive as the Python doubles of the reference and are rounded to fp32 once, as torch does for
      // value= / alpha= / weight= arguments of the foreach ops (1 - beta2 in fp32 would differ by 1.3e-5 relative)
      (float)beta2, (float)(1.0 - beta2), (float)bias_correction2, (float)eps, (float)weight_decay, (float)lr, (float)ckp1,
      (float)(lr * (beta1 * (1.0 - ckp1) - 1.0)), 1);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_adamw_sf_step_dev(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                          const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, const double* consts8,
                          double* state3, float* hyper8, void* stream) {
  FTC_REQUIRE(n_chunks >= 0 && chunks && ys && grads && exp_avg_sqs && zs && numels && consts8 && state3 && hyper8, "bad argument");
  adamw_sf_schedule_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(consts8, state3, hyper8);
  FTC_POST_LAUNCH();
  if (n_chunks == 0) return 0;
  adamw_sf_step_dev_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(
      (const SfChunk*)chunks, (float* const*)ys, (float* const*)grads, (float* const*)exp_avg_sqs, (float* const*)zs, numels, hyper8);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_radam_sf_step_dev(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                          const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, const double* consts8,
                          double* state3, float* hyper9, void* stream) {
  FTC_REQUIRE(n_chunks >= 0 && chunks && ys && grads && exp_avg_sqs && zs && numels && consts8 && state3 && hyper9, "bad argument");
  radam_sf_schedule_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(consts8, state3, hyper9);
  FTC_POST_LAUNCH();
  if (n_chunks == 0) return 0;
  adamw_sf_step_dev_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(
      (const SfChunk*)chunks, (float* const*)ys, (float* const*)grads, (float* const*)exp_avg_sqs, (float* const*)zs, numels, hyper9);
  FTC_POST_LAUNCH();
  return 0;
}

int ftc_radam_sf_step(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                      const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, double beta1, double beta2,
                      double bias_correction2, double eps, double weight_decay, double lr, double ckp1, int adam_step,
                      void* stream) {
  FTC_REQUIRE(n_chunks >= 0 && chunks && ys && grads && exp_avg_sqs && zs && numels, "bad argument");
  if (n_chunks == 0) return 0;
  adamw_sf_step_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(
      (const SfChunk*)chunks, (float* const*)ys, (float* const*)grads, (float* const*)exp_avg_sqs, (float* const*)zs, numels,
      (float)beta2, (float)(1.0 - beta2), (float)bias_correction2, (float)eps, (float)weight_decay, (float)lr, (float)ckp1,
      (float)(lr * (beta1 * (1.0 - ckp1) - 1.0)), adam_step ? 1 : 0);
  FTC_POST_LAUNCH();
  return 0;
}

}  // extern "C"
