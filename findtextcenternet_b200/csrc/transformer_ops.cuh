// Non-GEMM kernels of the Transformer hot path (models/transformer.py): input pad/cast + key-pad mask,
// decoder embedding, LayerNorm, attention, mask-predict step.  Row-major [rows, d] activations, fp32 or bf16.
#pragma once
#include "common.cuh"

namespace ftc {

// enc_input fp32 [M, Kin] -> out (dtype) [M, Kout] zero padded; keymask[m] = -inf if the row is all zero else 0
// (models/transformer.py:249-250 / :275-276)
int pad_cast_rows(const float* in, void* out, int dtype, float* keymask, int M, int Kin, int Kout, cudaStream_t s);

// out[m] = LayerNorm(in[m]) * gamma + beta, eps 1e-5 (nn.LayerNorm); in/out dtype, may alias
int layernorm_rows(const void* in, void* out, int dtype, const float* gamma, const float* beta, int M, int d, float eps,
                   cudaStream_t s);

// Decoder.forward head (models/transformer.py:226-233): sum_i Embedding_i[token % m_i] + pos[l] -> LayerNorm
int decoder_embed_ln(const int64_t* tokens, const float* e0, const float* e1, const float* e2, int m0, int m1, int m2,
                     const float* pos, const float* gamma, const float* beta, void* out, int dtype, int M, int L, int d,
                     float eps, cudaStream_t s);

// softmax(q k^T / sqrt(hd) + mask) v per (batch, head)  (F.scaled_dot_product_attention, models/transformer.py:133)
//   q: [B*Lt, q_stride] at column q_off + h*hd ; k, v: [B*Ls, kv_stride] at k_off / v_off + h*hd
//   mask: fp32 [B, Ls] additive (0 / -inf) or null ; out: [B*Lt, out_stride] at column h*hd
int attention(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off, int v_off,
              const float* mask, void* out, int out_stride, int dtype, int B, int heads, int hd, int Lt, int Ls,
              cudaStream_t s);

// the same product on tcgen05 / TMEM for bf16 sequences of up to 128 tokens (attention_tc.cu); returns 1 when it declines the shape
int attention_tc(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off, int v_off,
                 const float* mask, void* out, int out_stride, int B, int heads, int hd, int Lt, int Ls, cudaStream_t s);

// one mask-predict decision per position (models/transformer.py:311-324 + util_func.py:92-126 CRT):
//   logits fp32 [M, ld] with the three heads at columns g*head_ld (m_g valid each)
//   -> ids int64 [M], prob fp32 [M]; flags[0] |= any(dec_in==MSK && id>0 && !(p>0.99)); flags[1] |= any(remask);
//   next_in[m] = remask ? MSK : id   (remask = p < 0.9 || id > 0x3FFFF)
int mask_predict_step(const float* logits, int ld, int head_ld, const int64_t* dec_in, int64_t* ids, float* prob,
                      int64_t* next_in, int* flags, int M, cudaStream_t s, int seq_len = 0, int* seq_flags = nullptr);
// per-sequence stop rules of the mask-predict loop (see transformer_ops.cu); seq_flags int [2 * batch], state int [3 * batch]
int mask_predict_advance(int* seq_flags, int* state, int64_t* dec_in, const int64_t* next_in, const int64_t* ids, int64_t* out_ids,
                         int batch, int seq_len, int k, int last, int* n_running, cudaStream_t s);

int interleave_rows_f32(float* dst, const float* a, const float* b, int rows, int cols, cudaStream_t s);
int cast_f32(void* dst, int dtype, const float* src, int64_t n, cudaStream_t s);

}  // namespace ftc
