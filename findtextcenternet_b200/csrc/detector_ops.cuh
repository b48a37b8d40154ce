// Bandwidth-bound detector kernels (stem, depthwise+SE, upsample, peak-pick/decode) and weight packing.
#pragma once
#include "common.cuh"

namespace ftc {

// stem: NCHW fp32 image in [0,1] -> (x*2-1) -> conv3x3 s2 p1 -> BN -> SiLU -> NHWC (dtype)
// models/detector.py:218 + torchvision efficientnet.py:271-276
//   nhwc255 = 1: img is the backend-ABI tile layout [B,H,W,3] fp32 in 0..255 (process_ocr_base.py:49-51)
int stem_conv(const float* img, int nhwc255, void* out, int dtype, int B, int H, int W, int Cout,
              const float* w /*[27][Cout]*/, const float* scale, const float* bias, cudaStream_t s);

// depthwise 3x3 (stride 1/2, pad 1) + BN + SiLU on NHWC, and per-(b, spatial tile, c) partial sums of the output for SE
// (torchvision efficientnet.py:137-149; ops/misc.py:251 avgpool).  se_sum (optional): [B][nt][C] fp32 with
// nt = dwconv3x3_tiles(H, W, stride, dtype), every entry written (no atomics, nothing to zero): deterministic squeeze.
int dwconv3x3_tiles(int H, int W, int stride, int dtype);
int dwconv3x3(const void* in, void* out, int dtype, int B, int H, int W, int C, int stride,
              const float* w /*[9][C]*/, const float* scale, const float* bias, float* se_sum /*[B,nt,C]*/,
              cudaStream_t s);

// SE excitation: scale[b,c] = sigmoid(fc2(silu(fc1(sum_t sum[b,t,:] / HW)))) (ops/misc.py:251-261); the nt tile partials are
// added in order.  hid: scratch [B, S] fp32
int se_fc(const float* sum, int nt, float* scale_out, float* hid, int B, int C, int S, float inv_hw, const float* w1 /*[S][C]*/,
          const float* b1, const float* w2t /*[S][C]*/, const float* b2, cudaStream_t s);

// stride-1 depthwise 3x3 + BN + SiLU with the SE squeeze and fc1 folded in (one CTA per image x 32 channels):
//   hid_part[b, g, s] = sum_{c in group g} w1[s, c] * mean_hw(out[b, :, :, c])   (hid_part [B][C/32][S] fp32, every entry written)
bool dwconv3x3_se_supported(int H, int W, int C, int stride);
int dwconv3x3_se(const void* in, void* out, int dtype, int B, int H, int W, int C, const float* w /*[9][C]*/, const float* scale,
                 const float* bias, const float* w1 /*[S][C]*/, int S, float* hid_part, cudaStream_t s, int flip = 0);
// the same strip kernel as a PLAIN stride-1 depthwise 3x3 (scale == nullptr: no BN / SiLU / SE); flip = 1 rotates the taps by 180
// degrees, which makes it the data gradient (train step: ftc_train_dwconv3x3 / ftc_train_dwconv3x3_dgrad)
int dwconv3x3_raw_strip(const void* in, void* out, int dtype, int B, int H, int W, int C, const float* w /*[9][C]*/, int flip, cudaStream_t s);
// scale[b,c] = sigmoid(b2[c] + w2t[:,c] . silu(sum_g hid_part[b,g,:] + b1)), groups added in a fixed order (G = C / 32)
int se_fc2_hid(const float* hid_part, int G, float* scale_out, int B, int C, int S, const float* b1,
               const float* w2t, const float* b2, cudaStream_t s);

// bilinear x2, align_corners=True, NHWC (nn.UpsamplingBilinear2d, models/detector.py:170)
int upsample2x(const void* in, void* out, int dtype, int B, int H, int W, int C, cudaStream_t s);

// Leafmap.top_conv (models/detector.py:188-190) of the small heads (out_dim 1 or 2): 3x3, 192 -> od, +bias.
//   y: NHWC (dtype) with head (head0+i)'s 192 channels at [(head0+i)*192, ...); w: fp32 [sum(od)][9*192] tap-major;
//   out: NCHW fp32 [B, out_ch, H, W], head i writes channels [sum(od[:i]), +od[i])
int head_top_conv(const void* y, int dtype, int pix_stride, int head0, int n_heads, const int* od, const float* w,
                  const float* bias, float* out, int out_ch, int B, int H, int W, cudaStream_t s);

// CenterNetDetector.forward tail (models/detector.py:289-296): heat9 NCHW fp32 -> heat10 NCHW fp32
int peak_pick(const float* heat9, float* heat10, int B, int H, int W, cudaStream_t s);

// per-tile peak compaction + box decode (process_ocr_base.py:498-538), sorted by descending (score, -index).
//   tile_meta[b] = {x_i, y_i, x_min, x_max, y_min, y_max} (page offset and centre-crop mask bounds)
//   loc [B][max_peaks][9] fp32: p, cx, cy, w, h, c1, c2, c4, c8 ; feat [B][max_peaks][100] fp32
//   count[b] = rows written = min(total[b], max_peaks); total[b] = every peak >= cut_off the reference would keep.
//   When total > max_peaks the max_peaks HIGHEST-scoring peaks are kept (deterministic), never an arbitrary subset.
//   scratch: peak_decode_scratch_bytes(B, H, W) (one 8-byte key per map pixel: every pixel may be a candidate)
size_t peak_decode_scratch_bytes(int B, int H, int W);
int peak_decode(const float* heat9, const float* feat, int B, int H, int W, int FC, const int* tile_meta,
                float cut_off, float page_w, float page_h, int max_peaks, int* count, int* total, float* loc,
                float* gfeat, void* scratch, cudaStream_t s);

// ---- weight packing (device side; sources are fp32 torch-layout tensors) ----
// dst[o*Kpad + k_off + (ky*kw+kx)*C + c] = src[((o*Itot + c_off + c)*kh + ky)*kw + kx] * (cscale ? cscale[c] : 1)
int pack_conv_weight(void* dst, int dtype, const float* src, int O, int Itot, int kh, int kw, int c_off, int C,
                     int k_off, int Kpad, int o_off, const float* cscale, cudaStream_t s);
// scale = gamma * rsqrt(var + eps), bias = beta - mean * scale   (both written at [off, off+C))
int bn_fold(float* scale, float* bias, const float* gamma, const float* beta, const float* mean, const float* var,
            float eps, int C, cudaStream_t s);
// Leafmap border-aware bias table (see DESIGN.md "in_bn before zero padding"):
//   tab[case][n] = bias[n] + scale[n] * sum_{taps valid in case} sum_c w[n][c_off+c][ky][kx] * shift[c]
int leaf_bias_table(float* tab /*[9][ld], pre-offset to this group's first column*/, int ld, const float* w, int O,
                    int Itot, int c_off, int C, const float* shift, const float* scale, const float* bias, cudaStream_t s);
int fill_f32(float* p, float v, int64_t n, cudaStream_t s);
// dst[c][r] = src[r][c] for src [R][C]: depthwise [C][9]->[9][C], stem [Cout][27]->[27][Cout], SE fc2 [C][S]->[S][C]
int transpose_f32(float* dst, const float* src, int R, int C, cudaStream_t s);

}  // namespace ftc
