// Implicit-GEMM convolution / linear op shared by every dense layer of the hot path.
//
//   out[m, g*N + n] = act( (sum_k A[m,k] * W[g][n][k]) * scale[g*N+n] + bias_tab[case(m)][g*N+n] )
//                     + res1[m, n] + res2[m, n]
// (FusedMBConv with expand 1 adds its input AFTER the SiLU, torchvision efficientnet.py:224-230; every other
//  residual user has act = none.  SwiGLU: residuals are added to the (x1, xg) pair columns before gating -- unused.)
//
// m enumerates output pixels (b, oy, ox) of an NHWC tensor; k enumerates (source, ky, kx, c):
// first the taps of source A (shared by all groups), then the taps of source B (per-group channel
// slice) -- this is the Leafmap "cat([y, bn(x)])" conv (models/detector.py:199) without materialising
// the concat.  A per-op chunk table (one int per 8 channels of k) carries (valid, src, ky, kx, c).
// 1x1 convs / nn.Linear are the H=M, W=1 special case.
#pragma once
#include <vector>
#include "common.cuh"

namespace ftc {

constexpr int KCHUNK = 8;          // channels per k-chunk (16 B of bf16)
constexpr int KBLOCK = 64;         // k per pipeline stage (one 128 B swizzle row of bf16)
constexpr int MAX_GROUPS = 9;

// chunk table entry
constexpr uint32_t KT_VALID = 1u << 31;
constexpr uint32_t KT_SRCB = 1u << 30;
__host__ __device__ inline uint32_t kt_make(bool srcb, int ky, int kx, int c) {
  return KT_VALID | (srcb ? KT_SRCB : 0u) | (uint32_t(ky) << 18) | (uint32_t(kx) << 16) | uint32_t(c);
}
__host__ __device__ inline int kt_ky(uint32_t e) { return (e >> 18) & 3; }
__host__ __device__ inline int kt_kx(uint32_t e) { return (e >> 16) & 3; }
__host__ __device__ inline int kt_c(uint32_t e) { return e & 0xFFFF; }

// tile plan of the tcgen05 path (conv_gemm_tc.cu); also fixes the packed-weight layout
struct ConvTcPlan {
  int BN;       // N tile (multiple of 16, <= 256) = UMMA N
  int NT;       // N tiles per group; packed weight rows per group = NT*BN
  int NKB;      // K / KBLOCK
  int stages;   // smem pipeline depth
  int MT;       // 128-row sub-tiles per CTA tile (1 or 2)
  int flags;    // experiment bits (env FTC_TC_FLAGS; results are garbage when >= 4): 1 try_wait suspend hint, 2 epilogue
                // poll backoff, 4 skip B copies, 8 skip A gathers, 16 skip MMAs, 32 skip epilogue math/stores
  int tma;      // operand-A path: TMA_NONE = cp.async im2col gather (conv_gemm_tc.cu); TMA_ROWS / TMA_HALO = tensor-tile
                // TMA loads (conv_gemm_tma.cu).  Fixes the k-block ORDER of the packed weights (see pack_conv_weight_tc)
  int nGA, nGB; // TMA paths: 64-channel chunks of source A / source B; K = 64 * ksize^2 * (nGA + nGB)
  int kb32;     // TMA_HALO with <= 32 input channels: 32-wide k-blocks (64-byte rows, SWIZZLE_64B) instead of zero-padding to 64
};
enum TmaMode : int { TMA_NONE = 0, TMA_ROWS = 1, TMA_HALO = 2 };
constexpr int HALO_TW = 16;        // halo tiles are 16 pixels wide and 8 (MT=1) or 16 (MT=2) rows high

struct ConvGemmParams {
  // geometry
  int B, H, W;          // input spatial (both sources)
  int Ho, Wo;           // output spatial
  int stride, pad;
  int M;                // B*Ho*Wo
  int N;                // output channels per group
  int G;                // groups (source A shared, source B sliced by g*srcB_group_stride)
  int K;                // padded k (multiple of KBLOCK)
  // sources (element type = dtype)
  const void* srcA; int a_pix_stride; int a_ch_off;
  const void* srcB; int b_pix_stride; int b_ch_off; int b_group_stride;
  int CA, CB;           // channels taken from source A / B per tap (plan input of the TMA paths; 0 = not given)
  const uint32_t* ktab; // K / KCHUNK entries
  const float* a_scale; // optional [B, a_scale_stride] multiplier on source A channels (SE), 1x1 only
  int a_scale_stride;
  // weights [G*N][K] (k-major), element type = dtype
  const void* w;
  // epilogue
  const float* scale;     // [G*N] or null (=1)
  const float* bias_tab;  // [ncase][G*N] or null (=0); ncase 1 or 9 (3x3 border cases)
  int ncase;
  int act;
  const void* res1; int res1_stride; int res1_row_mod;   // row index = res_row_mod ? m % res_row_mod : m
  const void* res2; int res2_stride;
  // output
  void* out; int out_layout; int out_stride;   // NHWC: elements per pixel;  NCHW: total channels
  int out_ch_base[MAX_GROUPS];                 // first output channel of group g
  int n_valid[MAX_GROUPS];                     // channels of group g actually stored (<= N)
  int dtype;                                   // DT_F32 / DT_BF16 (sources, weights, NHWC output, residuals)
  ConvTcPlan tc;                               // tcgen05 path only
  unsigned long long* trace;                   // optional clock64 trace of CTA 0 (ftc_debug_set_trace), else null
};

// SIMT (CUDA-core, fp32 accumulate) implementation: any dtype, the parity path.
int conv_gemm_simt(const ConvGemmParams& p, cudaStream_t stream);
// tcgen05/TMEM implementation: bf16 only, the product path.  p.tc must come from conv_gemm_tc_plan and the
// weights must have been packed with pack_conv_weight_tc for the same plan.
//   allow_tma: pick a TMA operand path when the geometry allows it (needs p.CA/p.CB, H, W, stride, pad, pixel strides);
//   the plan may then change K (per-source channel padding to 64): callers must use plan->NKB * 64 as the packed K.
int conv_gemm_tc_plan(const ConvGemmParams& p, ConvTcPlan* plan, bool allow_tma = false);
int conv_gemm_tc(const ConvGemmParams& p, cudaStream_t stream);
int conv_gemm_tma(const ConvGemmParams& p, cudaStream_t stream);   // called by conv_gemm_tc when p.tc.tma != TMA_NONE
// tcgen05 weight image: per (padded row tile, k-block) one [BN rows][64 k] bf16 block in the 128B-swizzled
// K-major shared-memory layout, so that a stage's B operand is ONE contiguous cp.async.bulk.
//   row R = o_off + o (o_off in padded-row space: group g starts at g*NT*BN)
//   halo_order = 0: k = k_off + (ky*kw+kx)*C + c                      (tap-major; im2col kernel and every 1x1)
//   halo_order = 2: k = (kx*3 + ky)*32 + c, 32-wide k-blocks in the 64B-swizzled image (TMA_HALO with plan.kb32)
//   halo_order = 1: k = k_off + ((c/64)*9 + kx*3 + ky)*64 + c%64      (TMA_HALO: per 64-channel chunk and column
//                   shift kx one halo tile serves the three row taps ky; k_off of source B = 9*64*nGA)
int pack_conv_weight_tc(void* dst, const float* src, int O, int Itot, int kh, int kw, int c_off, int C, int k_off,
                        int Kpad, int o_off, int BN, const float* cscale, cudaStream_t s, int halo_order = 0, int dgrad_rows = 0);
size_t conv_tc_weight_bytes(const ConvTcPlan& plan, int G);
void conv_gemm_tc_set_trace(unsigned long long* dev_ptr);
unsigned long long* conv_gemm_trace_ptr();   // current trace buffer (null = off)   // 4 x 1024 u64: MMA wait start/end, producer wait start/end

std::vector<uint32_t> make_ktab(int CA, int CB, int ksize, int* Kout);

// tuning overrides (0 = automatic), initialised from the environment (FTC_TMA_MT, FTC_TMA_FLAGS) and settable through
// ftc_debug_set_gemm_tuning for shape sweeps (tools/bench_gemm.py)
struct GemmTuning { int mt, flags, box_depth, plan_bn, no_bstat, epi8, nb; };
GemmTuning& gemm_tuning();

}  // namespace ftc
