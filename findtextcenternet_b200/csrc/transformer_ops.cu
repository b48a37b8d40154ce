// Non-GEMM Transformer kernels.  See transformer_ops.cuh for the reference lines each one follows.
#include "transformer_ops.cuh"
#include <math.h>

namespace ftc {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) pad_cast_rows_kernel(const float* __restrict__ in, T* __restrict__ out,
                                                            float* __restrict__ keymask, int M, int Kin, int Kout) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* ip = in + (int64_t)warp * Kin;
  T* op = out + (int64_t)warp * Kout;
  bool nz = false;
  for (int k = lane; k < Kout; k += 32) {
    float v = k < Kin ? ip[k] : 0.f;
    nz |= (v != 0.f);
    op[k] = from_f<T>(v);
  }
  nz = __any_sync(0xffffffffu, nz);
  if (lane == 0 && keymask) keymask[warp] = nz ? 0.f : -INFINITY;
}

// ---------------------------------------------------------------------------------------------------
constexpr int LN_MAX_PER_LANE = 32;   // d <= 1024

template <typename T>
__device__ __forceinline__ void ln_finish(float (&x)[LN_MAX_PER_LANE], int n_per, int lane, int d, const float* gamma,
                                          const float* beta, T* op, float eps) {
  float s = 0.f;
  for (int i = 0; i < n_per; ++i) { int c = lane + 32 * i; if (c < d) s += x[i]; }
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
  for (int i = 0; i < n_per; ++i) { int c = lane + 32 * i; if (c < d) { float t = x[i] - mean; q += t * t; } }
  const float rstd = rsqrtf(warp_sum(q) / (float)d + eps);
  for (int i = 0; i < n_per; ++i) {
    int c = lane + 32 * i;
    if (c < d) op[c] = from_f<T>((x[i] - mean) * rstd * gamma[c] + beta[c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int M, int d, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const T* ip = in + (int64_t)warp * d;
  const int n_per = (d + 31) / 32;
  float x[LN_MAX_PER_LANE];
  for (int i = 0; i < n_per; ++i) { int c = lane + 32 * i; x[i] = c < d ? to_f(ip[c]) : 0.f; }
  ln_finish<T>(x, n_per, lane, d, gamma, beta, out + (int64_t)warp * d, eps);
}

template <typename T>
__global__ void __launch_bounds__(256) decoder_embed_ln_kernel(const int64_t* __restrict__ tokens, const float* __restrict__ e0,
                                                               const float* __restrict__ e1, const float* __restrict__ e2,
                                                               int m0, int m1, int m2, const float* __restrict__ pos,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               T* __restrict__ out, int M, int L, int d, float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int64_t tok = tokens[warp];
  const int l = warp % L;
  // python modulo for (never expected) negative ids
  const int64_t r0 = ((tok % m0) + m0) % m0, r1 = ((tok % m1) + m1) % m1, r2 = ((tok % m2) + m2) % m2;
  const float* p0 = e0 + r0 * d;
  const float* p1 = e1 + r1 * d;
  const float* p2 = e2 + r2 * d;
  const float* pp = pos + (int64_t)l * d;
  const int n_per = (d + 31) / 32;
  float x[LN_MAX_PER_LANE];
  for (int i = 0; i < n_per; ++i) {
    int c = lane + 32 * i;
    x[i] = c < d ? ((p0[c] + p1[c]) + p2[c]) + pp[c] : 0.f;
  }
  ln_finish<T>(x, n_per, lane, d, gamma, beta, out + (int64_t)warp * d, eps);
}

// ---------------------------------------------------------------------------------------------------
// attention: one CTA per (batch, head); K (padded rows) and V in shared memory as fp32; each warp owns query rows.
template <typename T, int HD>
__global__ void __launch_bounds__(256) attention_kernel(const T* __restrict__ q, int q_stride, int q_off,
                                                        const T* __restrict__ k, const T* __restrict__ v, int kv_stride,
                                                        int k_off, int v_off, const float* __restrict__ mask,
                                                        T* __restrict__ out, int out_stride, int heads, int Lt, int Ls) {
  extern __shared__ float sm[];
  constexpr int KP = HD + 1;
  float* Ks = sm;                         // [Ls][HD+1]
  float* Vs = Ks + (size_t)Ls * KP;       // [Ls][HD]
  float* Ms = Vs + (size_t)Ls * HD;       // [Ls]
  float* Sc = Ms + Ls;                    // [8][Ls]
  float* Qs = Sc + 8 * (size_t)Ls;        // [8][HD]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* kb = k + (int64_t)b * Ls * kv_stride + k_off + h * HD;
  const T* vb = v + (int64_t)b * Ls * kv_stride + v_off + h * HD;
  for (int i = tid; i < Ls * HD; i += blockDim.x) {
    int j = i / HD, c = i - j * HD;
    Ks[j * KP + c] = to_f(kb[(int64_t)j * kv_stride + c]);
    Vs[j * HD + c] = to_f(vb[(int64_t)j * kv_stride + c]);
  }
  for (int j = tid; j < Ls; j += blockDim.x) Ms[j] = mask ? mask[(int64_t)b * Ls + j] : 0.f;
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  float* sc = Sc + (size_t)warp * Ls;
  float* qs = Qs + warp * HD;
  for (int i = warp; i < Lt; i += 8) {
    const T* qp = q + ((int64_t)b * Lt + i) * q_stride + q_off + h * HD;
    for (int c = lane; c < HD; c += 32) qs[c] = to_f(qp[c]);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < Ls; j += 32) {
      const float* kr = Ks + j * KP;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < HD; ++c) s = fmaf(qs[c], kr[c], s);
      s = s * scale + Ms[j];
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Ls; j += 32) {
      float e = expf(sc[j] - mx);
      sc[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    const float inv = 1.f / sum;
    T* op = out + ((int64_t)b * Lt + i) * out_stride + h * HD;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 32) {
      const int c = c0 + lane;
      if (c < HD) {
        float a = 0.f;
        for (int j = 0; j < Ls; ++j) a = fmaf(sc[j], Vs[j * HD + c], a);
        op[c] = from_f<T>(a * inv);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// mask-predict step, one warp per position
struct Top3 { float v[3]; int i[3]; };

__device__ __forceinline__ void top3_insert(Top3& t, float v, int i) {
  if (v > t.v[2] || (v == t.v[2] && i < t.i[2])) {
    t.v[2] = v; t.i[2] = i;
    if (t.v[2] > t.v[1] || (t.v[2] == t.v[1] && t.i[2] < t.i[1])) {
      float fv = t.v[1]; int fi = t.i[1]; t.v[1] = t.v[2]; t.i[1] = t.i[2]; t.v[2] = fv; t.i[2] = fi;
      if (t.v[1] > t.v[0] || (t.v[1] == t.v[0] && t.i[1] < t.i[0])) {
        fv = t.v[0]; fi = t.i[0]; t.v[0] = t.v[1]; t.i[0] = t.i[1]; t.v[1] = fv; t.i[1] = fi;
      }
    }
  }
}

constexpr int64_t CRT_M1 = 1091, CRT_M2 = 1093, CRT_M3 = 1097;

__host__ __device__ inline int64_t powmod(int64_t x, int64_t e, int64_t m) {
  int64_t r = 1;
  x %= m;
  while (e > 0) { if (e & 1) r = r * x % m; x = x * x % m; e >>= 1; }
  return r;
}

// Garner form of util_func.py:92-126 (same unique representative in [0, m1*m2*m3))
__device__ __forceinline__ int64_t crt3(int64_t b1, int64_t b2, int64_t b3, int64_t inv12, int64_t inv13, int64_t inv23) {
  int64_t t0 = b1 % CRT_M1;
  int64_t t1 = (((b2 - t0) % CRT_M2) + CRT_M2) % CRT_M2 * inv12 % CRT_M2;
  int64_t u = t0 + t1 * CRT_M1;
  int64_t t2 = (((b3 - u) % CRT_M3) + CRT_M3) % CRT_M3 * inv13 % CRT_M3 * inv23 % CRT_M3;
  return (u + t2 * CRT_M1 * CRT_M2) % (CRT_M1 * CRT_M2 * CRT_M3);
}

__global__ void __launch_bounds__(256) mask_predict_step_kernel(const float* __restrict__ logits, int ld, int head_ld,
                                                                const int64_t* __restrict__ dec_in, int64_t* __restrict__ ids,
                                                                float* __restrict__ prob, int64_t* __restrict__ next_in,
                                                                int* __restrict__ flags, int M, int64_t inv12, int64_t inv13,
                                                                int64_t inv23) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int mods[3] = {(int)CRT_M1, (int)CRT_M2, (int)CRT_M3};
  float tp[3][3];
  int ti[3][3];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const float* lp = logits + (int64_t)warp * ld + g * head_ld;
    const int m = mods[g];
    Top3 t;
    t.v[0] = t.v[1] = t.v[2] = -INFINITY;
    t.i[0] = t.i[1] = t.i[2] = 0x7fffffff;
    float mx = -INFINITY;
    for (int n = lane; n < m; n += 32) { float x = lp[n]; mx = fmaxf(mx, x); top3_insert(t, x, n); }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < m; n += 32) sum += expf(lp[n] - mx);
    sum = warp_sum(sum);
    // merge the per-lane top-3 lists: three rounds of (max value, min index) extraction
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float bv = t.v[0];
      int bi = t.i[0];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (t.i[0] == bi) { t.v[0] = t.v[1]; t.i[0] = t.i[1]; t.v[1] = t.v[2]; t.i[1] = t.i[2]; t.v[2] = -INFINITY; t.i[2] = 0x7fffffff; }
      tp[g][r] = expf(bv - mx) / sum;     // softmax probability of the r-th best class
      ti[g][r] = bi;
    }
  }
  // 27 candidates in itertools.product order (first modulus slowest)
  float bp = -1.f;
  int64_t bid = 0;
  int bc = 0x7fffffff;
  if (lane < 27) {
    const int a = lane / 9, bq = (lane / 3) % 3, c = lane % 3;
    const float p = expf((logf(fmaxf(tp[0][a], 1e-10f)) + logf(fmaxf(tp[1][bq], 1e-10f)) + logf(fmaxf(tp[2][c], 1e-10f))) / 3.0f);
    const int64_t id = crt3(ti[0][a], ti[1][bq], ti[2][c], inv12, inv13, inv23);
    bp = id > 0x3FFFF ? 0.f : p;
    bid = id;
    bc = lane;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float op = __shfl_xor_sync(0xffffffffu, bp, o);
    int oc = __shfl_xor_sync(0xffffffffu, bc, o);
    int64_t oid = __shfl_xor_sync(0xffffffffu, bid, o);
    if (op > bp || (op == bp && oc < bc)) { bp = op; bc = oc; bid = oid; }
  }
  if (lane == 0) {
    ids[warp] = bid;
    prob[warp] = bp;
    const int64_t din = dec_in[warp];
    const bool remask = (bp < 0.9f) || (bid > 0x3FFFF);
    if (din == 3 && bid > 0 && !(bp > 0.99f)) atomicOr(&flags[0], 1);
    if (remask) atomicOr(&flags[1], 1);
    next_in[warp] = remask ? (int64_t)3 : bid;
  }
}

__global__ void interleave_rows_kernel(float* __restrict__ dst, const float* __restrict__ a, const float* __restrict__ b,
                                       int rows, int cols) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  int r = i / cols, c = i - (int64_t)r * cols;
  dst[((int64_t)2 * r) * cols + c] = a[i];
  dst[((int64_t)2 * r + 1) * cols + c] = b[i];
}

template <typename T>
__global__ void cast_f32_kernel(T* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = from_f<T>(src[i]);
}

}  // namespace

int pad_cast_rows(const float* in, void* out, int dtype, float* keymask, int M, int Kin, int Kout, cudaStream_t s) {
  int grid = ceil_div(M, 8);
  if (dtype == DT_F32) pad_cast_rows_kernel<float><<<grid, 256, 0, s>>>(in, (float*)out, keymask, M, Kin, Kout);
  else pad_cast_rows_kernel<bf16><<<grid, 256, 0, s>>>(in, (bf16*)out, keymask, M, Kin, Kout);
  FTC_POST_LAUNCH();
  return 0;
}

int layernorm_rows(const void* in, void* out, int dtype, const float* gamma, const float* beta, int M, int d, float eps,
                   cudaStream_t s) {
  FTC_REQUIRE(d <= 32 * LN_MAX_PER_LANE, "LayerNorm width > 1024");
  int grid = ceil_div(M, 8);
  if (dtype == DT_F32) layernorm_rows_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (float*)out, gamma, beta, M, d, eps);
  else layernorm_rows_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)in, (bf16*)out, gamma, beta, M, d, eps);
  FTC_POST_LAUNCH();
  return 0;
}

int decoder_embed_ln(const int64_t* tokens, const float* e0, const float* e1, const float* e2, int m0, int m1, int m2,
                     const float* pos, const float* gamma, const float* beta, void* out, int dtype, int M, int L, int d,
                     float eps, cudaStream_t s) {
  FTC_REQUIRE(d <= 32 * LN_MAX_PER_LANE, "embed width > 1024");
  int grid = ceil_div(M, 8);
  if (dtype == DT_F32)
    decoder_embed_ln_kernel<float><<<grid, 256, 0, s>>>(tokens, e0, e1, e2, m0, m1, m2, pos, gamma, beta, (float*)out, M, L, d, eps);
  else
    decoder_embed_ln_kernel<bf16><<<grid, 256, 0, s>>>(tokens, e0, e1, e2, m0, m1, m2, pos, gamma, beta, (bf16*)out, M, L, d, eps);
  FTC_POST_LAUNCH();
  return 0;
}

template <typename T, int HD>
static int launch_attention(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off,
                            int v_off, const float* mask, void* out, int out_stride, int B, int heads, int Lt, int Ls,
                            cudaStream_t s) {
  size_t smem = ((size_t)Ls * (HD + 1) + (size_t)Ls * HD + Ls + 8 * (size_t)Ls + 8 * HD) * sizeof(float);
  FTC_REQUIRE(smem <= 227 * 1024, "attention: sequence too long for the shared-memory K/V tile");
  static bool attr_done = false;
  if (!attr_done) {
    FTC_CHECK_CUDA(cudaFuncSetAttribute(attention_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  attention_kernel<T, HD><<<B * heads, 256, smem, s>>>((const T*)q, q_stride, q_off, (const T*)k, (const T*)v, kv_stride, k_off,
                                                       v_off, mask, (T*)out, out_stride, heads, Lt, Ls);
  FTC_POST_LAUNCH();
  return 0;
}

int attention(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off, int v_off,
              const float* mask, void* out, int out_stride, int dtype, int B, int heads, int hd, int Lt, int Ls,
              cudaStream_t s) {
#define ATT(TT, HH) return launch_attention<TT, HH>(q, q_stride, q_off, k, v, kv_stride, k_off, v_off, mask, out, out_stride, B, heads, Lt, Ls, s)
  if (dtype == DT_F32) {
    if (hd == 16) ATT(float, 16);
    if (hd == 32) ATT(float, 32);
    if (hd == 64) ATT(float, 64);
  } else {
    if (hd == 16) ATT(bf16, 16);
    if (hd == 32) ATT(bf16, 32);
    if (hd == 64) ATT(bf16, 64);
  }
#undef ATT
  FTC_REQUIRE(false, "attention: head_dim must be 16, 32 or 64");
}

int mask_predict_step(const float* logits, int ld, int head_ld, const int64_t* dec_in, int64_t* ids, float* prob,
                      int64_t* next_in, int* flags, int M, cudaStream_t s) {
  const int64_t inv12 = powmod(CRT_M1, CRT_M2 - 2, CRT_M2), inv13 = powmod(CRT_M1, CRT_M3 - 2, CRT_M3),
                inv23 = powmod(CRT_M2, CRT_M3 - 2, CRT_M3);
  mask_predict_step_kernel<<<ceil_div(M, 8), 256, 0, s>>>(logits, ld, head_ld, dec_in, ids, prob, next_in, flags, M, inv12,
                                                         inv13, inv23);
  FTC_POST_LAUNCH();
  return 0;
}

int interleave_rows_f32(float* dst, const float* a, const float* b, int rows, int cols, cudaStream_t s) {
  int64_t n = (int64_t)rows * cols;
  interleave_rows_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(dst, a, b, rows, cols);
  FTC_POST_LAUNCH();
  return 0;
}

int cast_f32(void* dst, int dtype, const float* src, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  if (dtype == DT_F32) cast_f32_kernel<float><<<(int)((n + 255) / 256), 256, 0, s>>>((float*)dst, src, n);
  else cast_f32_kernel<bf16><<<(int)((n + 255) / 256), 256, 0, s>>>((bf16*)dst, src, n);
  FTC_POST_LAUNCH();
  return 0;
}

}  // namespace ftc
