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

// bf16 fast path, d = NV * 256: every lane owns NV runs of 8 consecutive channels (16-byte loads / stores), statistics in
// registers.  The generic kernel above reads 2 bytes per load and keeps its row in local memory (26 us per call on
// [25600, 512]; this one is bandwidth-bound).
template <int NV>
__global__ void __launch_bounds__(256) layernorm_rows_vec_kernel(const bf16* __restrict__ in, bf16* __restrict__ out,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 int M, float eps) {
  constexpr int D = NV * 256;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  if (warp >= M) return;
  const bf16* ip = in + (int64_t)warp * D + lane * 8;
  float x[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) load8(ip + i * 256, x[i]);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[i][j];
  const float mean = warp_sum(s) * (1.0f / (float)D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float t = x[i][j] - mean; q = fmaf(t, t, q); }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / (float)D) + eps);
  bf16* op = out + (int64_t)warp * D + lane * 8;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float ga[8], be[8], o[8];
    load8(gamma + i * 256 + lane * 8, ga);
    load8(beta + i * 256 + lane * 8, be);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf((x[i][j] - mean) * rstd, ga[j], be[j]);
    store8(op + i * 256, o);
  }
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
                                                                int64_t inv23, int seq_len, int* __restrict__ seq_flags) {
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
    if (din == 3 && bid > 0 && !(bp > 0.99f)) { atomicOr(&flags[0], 1); if (seq_flags) atomicOr(&seq_flags[2 * (warp / seq_len)], 1); }
    if (remask) { atomicOr(&flags[1], 1); if (seq_flags) atomicOr(&seq_flags[2 * (warp / seq_len) + 1], 1); }
    next_in[warp] = remask ? (int64_t)3 : bid;
  }
}

// Per-sequence bookkeeping of the mask-predict loop (TransformerPredictor.forward, models/transformer.py:326-358, evaluated for
// EACH sequence as the reference evaluates it for its batch of one): a sequence stops at pass k if none of its masked, non-PAD
// positions is below 0.99 ("early stop"), or if nothing of it would be re-masked ("no remask stop", not on the last pass), or
// after the last pass; its code points are frozen then.  One CTA per sequence.  state[b] = {done, passes, reason}.
__global__ void __launch_bounds__(128) mask_predict_advance_kernel(int* __restrict__ seq_flags, int* __restrict__ state,
                                                                   int64_t* __restrict__ dec_in, const int64_t* __restrict__ next_in,
                                                                   const int64_t* __restrict__ ids, int64_t* __restrict__ out_ids,
                                                                   int seq_len, int k, int last, int* __restrict__ n_running) {
  const int b = blockIdx.x;
  int* st = state + 3 * b;
  const int f0 = seq_flags[2 * b], f1 = seq_flags[2 * b + 1];
  __syncthreads();
  if (threadIdx.x == 0) { seq_flags[2 * b] = 0; seq_flags[2 * b + 1] = 0; }       // re-arm for the next pass
  if (st[0]) return;
  int reason = 0;
  bool stop = false;
  if (f0 == 0) { stop = true; reason = 1; }
  else if (k < last && f1 == 0) { stop = true; reason = 2; }
  else if (k == last) stop = true;
  if (stop) {
    for (int i = threadIdx.x; i < seq_len; i += blockDim.x) out_ids[(int64_t)b * seq_len + i] = ids[(int64_t)b * seq_len + i];
    __syncthreads();
    if (threadIdx.x == 0) { st[0] = 1; st[1] = k + 1; st[2] = reason; }
  } else {
    for (int i = threadIdx.x; i < seq_len; i += blockDim.x) dec_in[(int64_t)b * seq_len + i] = next_in[(int64_t)b * seq_len + i];
    if (threadIdx.x == 0) atomicAdd(n_running, 1);
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
  const bool vec_ok = dtype == DT_BF16 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                      ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  if (dtype == DT_F32) layernorm_rows_kernel<float><<<grid, 256, 0, s>>>((const float*)in, (float*)out, gamma, beta, M, d, eps);
  else if (vec_ok && d == 256) FTC_CHECK_CUDA(launch_pdl(layernorm_rows_vec_kernel<1>, dim3(grid), dim3(256), 0, s, (const bf16*)in, (bf16*)out, gamma, beta, M, eps));
  else if (vec_ok && d == 512) FTC_CHECK_CUDA(launch_pdl(layernorm_rows_vec_kernel<2>, dim3(grid), dim3(256), 0, s, (const bf16*)in, (bf16*)out, gamma, beta, M, eps));
  else if (vec_ok && d == 768) FTC_CHECK_CUDA(launch_pdl(layernorm_rows_vec_kernel<3>, dim3(grid), dim3(256), 0, s, (const bf16*)in, (bf16*)out, gamma, beta, M, eps));
  else if (vec_ok && d == 1024) FTC_CHECK_CUDA(launch_pdl(layernorm_rows_vec_kernel<4>, dim3(grid), dim3(256), 0, s, (const bf16*)in, (bf16*)out, gamma, beta, M, eps));
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

// ---------------------------------------------------------------------------------------------------
// bf16 attention on the warp-level tensor cores (mma.sync m16n8k16, fp32 accumulate), flash style: one CTA per
// (64 query rows, head, batch), 4 warps x 16 rows; K and V of the (batch, head) sit in shared memory as bf16 (row stride
// padded by 16 B: conflict-free ldmatrix); keys are walked in chunks of 32 with an online softmax in the exp2 domain; the
// S accumulator fragments are re-packed in registers as the A operand of P.V.  (The CUDA-core kernel above was 67 % of the
// cfg#4 transformer forward.)
template <int HD, int NW>
__global__ void __launch_bounds__(NW * 32) attention_mma_kernel(const bf16* __restrict__ q, int q_stride, int q_off,
                                                            const bf16* __restrict__ k, const bf16* __restrict__ v, int kv_stride,
                                                            int k_off, int v_off, const float* __restrict__ mask,
                                                            bf16* __restrict__ out, int out_stride, int Lt, int Ls, float scale_log2) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  constexpr int RS = HD + 8;                       // padded row stride (elements)
  constexpr int KS = HD / 16;                      // k-steps over the head dim
  constexpr int NO = HD / 8;                       // output n-tiles
  const int Lp = (Ls + 31) & ~31;
  bf16* Ks = reinterpret_cast<bf16*>(att_smem);    // [Lp][RS]
  bf16* Vs = Ks + (size_t)Lp * RS;                 // [Lp][RS]
  float* Ms = reinterpret_cast<float*>(Vs + (size_t)Lp * RS);   // [Lp] additive mask in the exp2 domain (0 / -inf)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * (NW * 16) + warp * 16;
  const bf16* kb = k + (int64_t)b * Ls * kv_stride + k_off + h * HD;
  const bf16* vb = v + (int64_t)b * Ls * kv_stride + v_off + h * HD;
  pdl_launch_dependents();
  pdl_wait();
  constexpr int PPR = HD / 8;                      // 16-byte pieces per row
  for (int i = tid; i < Lp * PPR; i += NW * 32) {
    const int j = i / PPR, pc = i - j * PPR;
    const bool ok = j < Ls;
    const uint32_t dk = (uint32_t)__cvta_generic_to_shared(Ks + (size_t)j * RS + pc * 8);
    const uint32_t dv = (uint32_t)__cvta_generic_to_shared(Vs + (size_t)j * RS + pc * 8);
    const bf16* sk = ok ? kb + (int64_t)j * kv_stride + pc * 8 : kb;
    const bf16* sv = ok ? vb + (int64_t)j * kv_stride + pc * 8 : vb;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dk), "l"(sk), "r"(ok ? 16u : 0u) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dv), "l"(sv), "r"(ok ? 16u : 0u) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int j = tid; j < Lp; j += NW * 32) Ms[j] = j < Ls ? (mask ? mask[(int64_t)b * Ls + j] : 0.f) : -INFINITY;
  // Q fragments straight from global memory (rows clamped; out-of-range rows are not stored)
  const int g = lane >> 2, qd = lane & 3;
  uint32_t qa[KS][4];
  {
    const int r0 = min(q0 + g, Lt - 1), r1 = min(q0 + g + 8, Lt - 1);
    const bf16* qp0 = q + ((int64_t)b * Lt + r0) * q_stride + q_off + h * HD + qd * 2;
    const bf16* qp1 = q + ((int64_t)b * Lt + r1) * q_stride + q_off + h * HD + qd * 2;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qp0 + ks * 16);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qp1 + ks * 16);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(qp0 + ks * 16 + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(qp1 + ks * 16 + 8);
    }
  }
  float o[NO][4];
#pragma unroll
  for (int n = 0; n < NO; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (q0 >= Lt) return;                            // whole warp beyond the sequence (after the only barrier)
  const uint32_t ks_s = (uint32_t)__cvta_generic_to_shared(Ks), vs_s = (uint32_t)__cvta_generic_to_shared(Vs);
  // ldmatrix lane addresses: K (non-transposed): key = lane & 7, head-dim block = lane >> 3;
  //                          V (transposed):     key = (lane & 7) + 8 * ((lane >> 3) & 1), head-dim block = lane >> 4
  const uint32_t k_lane = (uint32_t)(((lane & 7) * RS + (lane >> 3) * 8) * 2);
  const uint32_t v_lane = (uint32_t)((((lane & 7) + 8 * ((lane >> 3) & 1)) * RS + (lane >> 4) * 8) * 2);
  for (int c0 = 0; c0 < Lp; c0 += 32) {
    float sacc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sacc[nt][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < KS; kk += 2) {         // one ldmatrix.x4 = this n-tile's B fragments for two k-steps
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(ks_s + (uint32_t)(((c0 + nt * 8) * RS + kk * 16) * 2) + k_lane) : "memory");
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(sacc[nt][0]), "+f"(sacc[nt][1]), "+f"(sacc[nt][2]), "+f"(sacc[nt][3])
                     : "r"(qa[kk][0]), "r"(qa[kk][1]), "r"(qa[kk][2]), "r"(qa[kk][3]), "r"(b0), "r"(b1));
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(sacc[nt][0]), "+f"(sacc[nt][1]), "+f"(sacc[nt][2]), "+f"(sacc[nt][3])
                     : "r"(qa[kk + 1][0]), "r"(qa[kk + 1][1]), "r"(qa[kk + 1][2]), "r"(qa[kk + 1][3]), "r"(b2), "r"(b3));
      }
    }
    // scale + mask, chunk row maxima (rows g and g + 8; a row lives in the 4 lanes of a quad)
    float cm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float m0 = Ms[c0 + nt * 8 + qd * 2], m1 = Ms[c0 + nt * 8 + qd * 2 + 1];
      sacc[nt][0] = fmaf(sacc[nt][0], scale_log2, m0); sacc[nt][1] = fmaf(sacc[nt][1], scale_log2, m1);
      sacc[nt][2] = fmaf(sacc[nt][2], scale_log2, m0); sacc[nt][3] = fmaf(sacc[nt][3], scale_log2, m1);
      cm[0] = fmaxf(cm[0], fmaxf(sacc[nt][0], sacc[nt][1]));
      cm[1] = fmaxf(cm[1], fmaxf(sacc[nt][2], sacc[nt][3]));
    }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      cm[r] = fmaxf(cm[r], __shfl_xor_sync(0xffffffffu, cm[r], 1));
      cm[r] = fmaxf(cm[r], __shfl_xor_sync(0xffffffffu, cm[r], 2));
      const float mnew = fmaxf(mrow[r], cm[r]);
      const float msafe = mnew == -INFINITY ? 0.f : mnew;       // fully masked so far: keep everything at zero
      alpha[r] = exp2f(mrow[r] - msafe);
      mrow[r] = mnew;
      cm[r] = msafe;
      lrow[r] *= alpha[r];
    }
#pragma unroll
    for (int n = 0; n < NO; ++n) { o[n][0] *= alpha[0]; o[n][1] *= alpha[0]; o[n][2] *= alpha[1]; o[n][3] *= alpha[1]; }
    uint32_t pa[2][4];                              // P as A fragments: k-step j covers n-tiles 2j, 2j+1
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float p0 = exp2f(sacc[nt][0] - cm[0]), p1 = exp2f(sacc[nt][1] - cm[0]);
      const float p2 = exp2f(sacc[nt][2] - cm[1]), p3 = exp2f(sacc[nt][3] - cm[1]);
      lrow[0] += p0 + p1; lrow[1] += p2 + p3;
      __nv_bfloat162 lo = __floats2bfloat162_rn(p0, p1), hi = __floats2bfloat162_rn(p2, p3);
      pa[nt >> 1][(nt & 1) * 2] = *reinterpret_cast<uint32_t*>(&lo);
      pa[nt >> 1][(nt & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&hi);
    }
#pragma unroll
    for (int kc = 0; kc < 2; ++kc)
#pragma unroll
      for (int n = 0; n < NO; n += 2) {             // one ldmatrix.x4.trans = V fragments of two output n-tiles
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(vs_s + (uint32_t)(((c0 + kc * 16) * RS + n * 8) * 2) + v_lane) : "memory");
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(o[n][0]), "+f"(o[n][1]), "+f"(o[n][2]), "+f"(o[n][3])
                     : "r"(pa[kc][0]), "r"(pa[kc][1]), "r"(pa[kc][2]), "r"(pa[kc][3]), "r"(b0), "r"(b1));
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+f"(o[n + 1][0]), "+f"(o[n + 1][1]), "+f"(o[n + 1][2]), "+f"(o[n + 1][3])
                     : "r"(pa[kc][0]), "r"(pa[kc][1]), "r"(pa[kc][2]), "r"(pa[kc][3]), "r"(b2), "r"(b3));
      }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
  const float inv0 = 1.f / lrow[0], inv1 = 1.f / lrow[1];
  const int r0 = q0 + g, r1 = q0 + g + 8;
#pragma unroll
  for (int n = 0; n < NO; ++n) {
    if (r0 < Lt) {
      __nv_bfloat162 t = __floats2bfloat162_rn(o[n][0] * inv0, o[n][1] * inv0);
      *reinterpret_cast<__nv_bfloat162*>(out + ((int64_t)b * Lt + r0) * out_stride + h * HD + n * 8 + qd * 2) = t;
    }
    if (r1 < Lt) {
      __nv_bfloat162 t = __floats2bfloat162_rn(o[n][2] * inv1, o[n][3] * inv1);
      *reinterpret_cast<__nv_bfloat162*>(out + ((int64_t)b * Lt + r1) * out_stride + h * HD + n * 8 + qd * 2) = t;
    }
  }
}

template <int HD>
static int launch_attention_mma(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off,
                                int v_off, const float* mask, void* out, int out_stride, int B, int heads, int Lt, int Ls,
                                cudaStream_t s) {
  const int Lp = (Ls + 31) & ~31;
  const size_t smem = 2 * (size_t)Lp * (HD + 8) * 2 + (size_t)Lp * 4;
  FTC_REQUIRE(smem <= 227 * 1024, "attention: sequence too long for the shared-memory K/V tile");
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  // 8 warps (128 query rows) per CTA when the sequence is longer than 64: K / V of a (batch, head) are staged once instead of
  // once per 64 rows
#define ATT_MMA_LAUNCH(NW_)                                                                                                  \
  do {                                                                                                                       \
    static bool attr_done = false;                                                                                           \
    if (!attr_done) {                                                                                                        \
      FTC_CHECK_CUDA(cudaFuncSetAttribute(attention_mma_kernel<HD, NW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_done = true;                                                                                                      \
    }                                                                                                                        \
    FTC_CHECK_CUDA(launch_pdl(attention_mma_kernel<HD, NW_>, dim3((Lt + NW_ * 16 - 1) / (NW_ * 16), heads, B), dim3(NW_ * 32), smem, s, \
                              (const bf16*)q, q_stride, q_off, (const bf16*)k, (const bf16*)v, kv_stride, k_off, v_off, mask,  \
                              (bf16*)out, out_stride, Lt, Ls, scale_log2));                                                  \
  } while (0)
  if (Lt > 64) ATT_MMA_LAUNCH(8); else ATT_MMA_LAUNCH(4);
#undef ATT_MMA_LAUNCH
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
    // tensor-core path: 4-byte fragment loads / stores need even offsets and strides, cp.async needs 16-byte rows
    const bool mma_ok = q_stride % 2 == 0 && q_off % 2 == 0 && out_stride % 2 == 0 && kv_stride % 8 == 0 && k_off % 8 == 0 &&
                        v_off % 8 == 0 && B <= 65535 && heads <= 65535 && !getenv("FTC_ATT_NO_MMA") &&
                        (reinterpret_cast<uintptr_t>(k) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(q) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    // sequences of up to 128 tokens: tcgen05 kernel (attention_tc.cu); it declines (1) shapes it does not take
    if (mma_ok) {
      const int rc = attention_tc(q, q_stride, q_off, k, v, kv_stride, k_off, v_off, mask, out, out_stride, B, heads, hd, Lt, Ls, s);
      if (rc <= 0) return rc;
    }
    if (mma_ok && hd == 32) return launch_attention_mma<32>(q, q_stride, q_off, k, v, kv_stride, k_off, v_off, mask, out, out_stride, B, heads, Lt, Ls, s);
    if (mma_ok && hd == 64) return launch_attention_mma<64>(q, q_stride, q_off, k, v, kv_stride, k_off, v_off, mask, out, out_stride, B, heads, Lt, Ls, s);
    if (hd == 16) ATT(bf16, 16);
    if (hd == 32) ATT(bf16, 32);
    if (hd == 64) ATT(bf16, 64);
  }
#undef ATT
  FTC_REQUIRE(false, "attention: head_dim must be 16, 32 or 64");
}

int mask_predict_advance(int* seq_flags, int* state, int64_t* dec_in, const int64_t* next_in, const int64_t* ids, int64_t* out_ids,
                         int batch, int seq_len, int k, int last, int* n_running, cudaStream_t s) {
  mask_predict_advance_kernel<<<batch, 128, 0, s>>>(seq_flags, state, dec_in, next_in, ids, out_ids, seq_len, k, last, n_running);
  FTC_POST_LAUNCH();
  return 0;
}

int mask_predict_step(const float* logits, int ld, int head_ld, const int64_t* dec_in, int64_t* ids, float* prob,
                      int64_t* next_in, int* flags, int M, cudaStream_t s, int seq_len, int* seq_flags) {
  const int64_t inv12 = powmod(CRT_M1, CRT_M2 - 2, CRT_M2), inv13 = powmod(CRT_M1, CRT_M3 - 2, CRT_M3),
                inv23 = powmod(CRT_M2, CRT_M3 - 2, CRT_M3);
  mask_predict_step_kernel<<<ceil_div(M, 8), 256, 0, s>>>(logits, ld, head_ld, dec_in, ids, prob, next_in, flags, M, inv12,
                                                         inv13, inv23, seq_len > 0 ? seq_len : 1, seq_flags);
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
