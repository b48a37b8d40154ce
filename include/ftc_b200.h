/* C-ABI of the findtextCenterNet B200 (sm_100a) hot path.
 *
 * The reference (lithium0003/findtextCenterNet) has no FFI: its seam is Python duck typing
 * (process_ocr_base.py:49-55 backend ABI; models/detector.py / models/transformer.py module ABI).
 * This header is the boundary *underneath* the Python mirror of that surface
 * (findtextcenternet_b200/models, findtextcenternet_b200/process_ocr_b200.py): plain pointers and sizes,
 * caller-owned device buffers, every call enqueues on the caller's stream and never allocates device
 * memory or synchronises.  Return value: 0 on success, negative on error (ftc_last_error() has the text).
 * `stream` is a cudaStream_t passed as void*.
 */
#ifndef FTC_B200_H
#define FTC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTC_MAX_STAGES 8
#define FTC_MAX_HEADS 9

enum { FTC_PREC_F32 = 0, FTC_PREC_BF16 = 1 };
/* SIMT: CUDA-core fp32-accumulate parity path.  TCGEN05: tensor-core path, operand A staged by TMA tensor tiles where the
 * geometry allows (1x1; 3x3 stride 1 with W % 16 == 0), else by the cp.async im2col gather.  TCGEN05_IM2COL: tensor-core
 * path with the im2col gather everywhere (A/B comparison of the two operand pipelines). */
enum { FTC_GEMM_SIMT = 0, FTC_GEMM_TCGEN05 = 1, FTC_GEMM_TCGEN05_IM2COL = 2 };
/* layout/range of the `images` argument of ftc_detector_forward:
 *   NCHW_UNIT: [B,3,H,W] fp32 in [0,1]   (models/detector.py:217 CenterNetDetection.forward input)
 *   NHWC_255 : [B,H,W,3] fp32 in 0..255  (process_ocr_base.py:49-51 call_detector input; /255 as process_ocr_torch.py:44) */
enum { FTC_INPUT_NCHW_UNIT = 0, FTC_INPUT_NHWC_255 = 1 };

typedef struct ftc_stage_cfg {
  int fused;   /* 1 = FusedMBConv, 0 = MBConv (torchvision efficientnet.py:105-231) */
  int expand, kernel, stride, cin, cout, layers;
} ftc_stage_cfg;

/* replaces: models/detector.py:12-28 efficientnet_v2_xl() table + :203-215 CenterNetDetection.__init__ */
typedef struct ftc_detector_config {
  int stem_out;
  int n_stages;
  ftc_stage_cfg stages[FTC_MAX_STAGES];
  int last_channel;
  int n_heads;                   /* 9 */
  int head_out[FTC_MAX_HEADS];   /* 1,2,1,1,1,1,1,1,100; the last head is the feature map */
  char head_names[FTC_MAX_HEADS][32]; /* state_dict prefix of each head: "keyheatmap", "sizes", ... (models/detector.py:207-215) */
  int height, width;             /* 768, 768 */
  int precision;                 /* FTC_PREC_* : storage type of activations / GEMM operands */
  int gemm_backend;              /* FTC_GEMM_* */
} ftc_detector_config;

typedef struct ftc_detector ftc_detector;

int ftc_version(void);
const char* ftc_last_error(void);
/* number of kernels this library launched since load (bench.py "gpu_launches") */
int64_t ftc_launch_count(void);

/* ---- detector: CenterNetDetection.forward + CenterNetDetector.forward (models/detector.py:217-230, 289-296) ---- */
int ftc_detector_create(const ftc_detector_config* cfg, ftc_detector** out);
void ftc_detector_destroy(ftc_detector* d);
size_t ftc_detector_weight_bytes(const ftc_detector* d);
size_t ftc_detector_workspace_bytes(const ftc_detector* d, int batch);
/* names[i] = reference state_dict key below `prefix` handed in by the caller (e.g. "backbone.features.0.0.weight"),
 * ptrs[i] = device pointer to the fp32 contiguous tensor.  Folds BN, packs K-major GEMM operands. */
int ftc_detector_pack_weights(ftc_detector* d, int n, const char* const* names, const void* const* ptrs,
                              const int64_t* numels, void* packed, size_t packed_bytes, void* stream);
/* images: [B,3,H,W] fp32 NCHW in [0,1].  heat9: [B,9,H/4,W/4] fp32 NCHW.  feat: [B,F,H/4,W/4] fp32 NCHW.
 * heat10 (optional, may be NULL): [B,10,H/4,W/4] = CenterNetDetector output with the peak channel. */
int ftc_detector_forward(ftc_detector* d, const float* images, int batch, float* heat9, float* feat, float* heat10,
                         void* workspace, size_t workspace_bytes, void* stream);
/* The same forward in two parts, so that the host -> device copy of a batch of tiles (process_ocr_base.py:487-497 hands
 * run_detector host float32 tiles; 226 MB for 32 of them) overlaps the first layers:
 *   FTC_PART_EARLY: stem + backbone.features[1..3] (models/detector.py:139-146) for images [image0, image0 + n_images) of the
 *                   batch -- call it per chunk of images as soon as that chunk has arrived in `images` (device, full-batch base
 *                   pointer); heat9 / feat / heat10 are not touched (may be NULL);
 *   FTC_PART_REST : everything after features[3] for the whole batch (image0 = 0, n_images = batch), outputs as ftc_detector_forward.
 * EARLY over every image followed by REST is bit-identical to ftc_detector_forward. */
enum { FTC_PART_EARLY = 1, FTC_PART_REST = 2 };
int ftc_detector_forward_part(ftc_detector* d, const float* images, int batch, int part, int image0, int n_images, float* heat9,
                              float* feat, float* heat10, void* workspace, size_t workspace_bytes, void* stream);

int ftc_detector_set_input_format(ftc_detector* d, int fmt);
/* measurement: same forward with a CUDA-event pair around every op of the plan (synchronises the stream).
 * op_ms / op_flop (2*MAC for `batch` images) / op_kind (0 stem, 1 dense 3x3, 2 1x1, 3 depthwise, 4 SE, 5 upsample) */
int ftc_detector_num_ops(const ftc_detector* d);
int ftc_detector_forward_timed(ftc_detector* d, const float* images, int batch, float* heat9, float* feat, void* workspace,
                               size_t workspace_bytes, void* stream, int max_ops, float* op_ms, double* op_flop,
                               int* op_kind);
/* introspection (tests): device pointer of backbone tap `tap` (0..3 = x1..x4, NHWC, engine dtype) inside `workspace`
 * after a forward of `batch` images; *channels / *hw receive its shape. */
int ftc_detector_tap(const ftc_detector* d, int tap, int batch, void* workspace, void** ptr, int* channels, int* h, int* w);

/* ---- per-tile peak compaction + box decode (process_ocr_base.py:498-538) ----
 * tile_meta: int32 [B][6] = {offset_x, offset_y, mask_xmin, mask_xmax, mask_ymin, mask_ymax} (device)
 * count: int32 [B] rows written = min(total, max_peaks); total: int32 [B] = every peak the reference loop would keep (it has no
 * cap); when total > max_peaks the max_peaks highest-scoring peaks are returned (deterministic) and the caller sees the overflow.
 * Rows are ordered by descending score, ties by ascending flat pixel index.  loc: fp32 [B][max_peaks][9]; gfeat: fp32
 * [B][max_peaks][F]; scratch: ftc_peak_decode_scratch_bytes(batch, h, w); max_peaks in {1024, 2048, 4096}. */
size_t ftc_peak_decode_scratch_bytes(int batch, int h, int w);
int ftc_peak_decode(const float* heat9, const float* feat, int batch, int h, int w, int feat_ch, const int* tile_meta,
                    float cut_off, float page_w, float page_h, int max_peaks, int* count, int* total, float* loc, float* gfeat,
                    void* scratch, void* stream);
int ftc_peak_pick(const float* heat9, float* heat10, int batch, int h, int w, void* stream);

/* ---- Transformer: Encoder / Decoder / TransformerPredictor (models/transformer.py) ---- */
/* replaces: models/transformer.py:255-264 ModelDimensions */
typedef struct ftc_transformer_config {
  int enc_input_dim;   /* 106 */
  int embed_dim, head_num, enc_blocks, dec_blocks, max_enc_len, max_dec_len;
  int precision;       /* FTC_PREC_* */
  int gemm_backend;    /* FTC_GEMM_* */
} ftc_transformer_config;
typedef struct ftc_transformer ftc_transformer;

int ftc_transformer_create(const ftc_transformer_config* cfg, ftc_transformer** out);
void ftc_transformer_destroy(ftc_transformer* t);
size_t ftc_transformer_weight_bytes(const ftc_transformer* t);
size_t ftc_transformer_workspace_bytes(const ftc_transformer* t, int batch, int enc_len, int dec_len);
/* names = reference Transformer.state_dict() keys ("encoder.blocks.0.mha.q_proj.weight", ...), fp32 device tensors */
int ftc_transformer_pack_weights(ftc_transformer* t, int n, const char* const* names, const void* const* ptrs,
                                 const int64_t* numels, void* packed, size_t packed_bytes, void* stream);
/* logits rows are [3 * head_stride] fp32: head g (modulus 1091/1093/1097) at columns [g*head_stride, g*head_stride+m_g) */
int ftc_transformer_logit_stride(void);
int ftc_transformer_head_stride(void);
/* Transformer.forward (models/transformer.py:248-253): enc_input fp32 [B,Le,enc_input_dim], dec_input int64 [B,Ld]
 * -> logits fp32 [B*Ld, logit_stride].  Rows of enc_input that are all zero are masked keys. */
int ftc_transformer_forward(ftc_transformer* t, const float* enc_input, const int64_t* dec_input, int batch, int enc_len,
                            int dec_len, float* logits, void* workspace, size_t workspace_bytes, void* stream);
/* TransformerPredictor.forward (models/transformer.py:274-360): encoder once, then <= max_passes (reference: 8)
 * mask-predict decoder passes with the reference's two data-dependent exits.  out_ids int64 [B,Ld] (device).
 * This call synchronises the stream once per pass (the reference does it twice per pass).
 * stop_reason: 0 = ran all passes, 1 = "early stop" (:326), 2 = "no remask stop" (:356). */
int ftc_transformer_predict(ftc_transformer* t, const float* enc_input, int batch, int enc_len, int dec_len, int64_t* out_ids,
                            int max_passes, int* passes_run, int* stop_reason, void* workspace, size_t workspace_bytes,
                            void* stream);
/* The same loop for a batch of INDEPENDENT sequences (all feature chunks of a page in one call, process_ocr_base.py:187-283): each
 * sequence stops by the rules the reference applies to its batch of one (it always decodes chunk by chunk, :235), so the result
 * equals the chunk-by-chunk loop.  seq_state: device int32 [3 * batch] = {1, passes run, stop reason} per sequence on return;
 * scratch_i32: device int32 [2 * batch + 4].  Synchronises once per pass. */
int ftc_transformer_predict_each(ftc_transformer* t, const float* enc_input, int batch, int enc_len, int dec_len, int64_t* out_ids,
                                 int max_passes, int* seq_state, int* scratch_i32, void* workspace, size_t workspace_bytes,
                                 void* stream);
/* one mask-predict decision per position on fp32 logits (models/transformer.py:311-324 + util_func.py:92-126) */
int ftc_mask_predict_step(const float* logits, int ld, int head_ld, const int64_t* dec_in, int64_t* ids, float* prob,
                          int64_t* next_in, int* flags, int rows, void* stream);

/* ---- train step: fused multi-tensor Schedule-Free AdamW update (models/adamw_schedulefree.py:157-184) ----
 * chunks: device array of {int32 tensor, int32 pad, int64 offset} (one CTA per ftc_adamw_sf_chunk_elems() elements);
 * ys/grads/exp_avg_sqs/zs: device arrays of fp32 device pointers, numels: device int64 per tensor.
 * Updates exp_avg_sq, grad (normalised in place, as the reference does), y (= the parameter) and z. */
int ftc_adamw_sf_chunk_elems(void);
int ftc_adamw_sf_step(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                      const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, double beta1, double beta2,
                      double bias_correction2, double eps, double weight_decay, double lr, double ckp1, void* stream);
/* Graph-replayable form of ftc_adamw_sf_step: nothing step-dependent is passed by value.  consts8 (device double[8]) = lr, beta1,
 * beta2, eps, weight_decay, warmup_steps, r, weight_lr_power; state3 (device double[3]) = k, lr_max, weight_sum (the param-group
 * entries of models/adamw_schedulefree.py:121-140), advanced by the call; hyper8: device fp32[9] scratch (9 since round 2: [8] = Adam normalisation flag).  Two launches: a
 * one-thread schedule kernel, then the same element update as ftc_adamw_sf_step.  A CUDA graph that contains this call performs
 * one further optimizer step per replay. */
int ftc_adamw_sf_step_dev(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                          const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, const double* consts8,
                          double* state3, float* hyper8, void* stream);
/* Schedule-Free RAdam (models/radam_schedulefree.py:109-236, train3.py:121): same update with the rectified lr computed by
 * the caller; adam_step = 0 during the early phase (rho_t <= 4), where the gradient is NOT normalised (:182-190) */
/* graph-replayable Schedule-Free RAdam: consts8 (device double[8]) = lr, beta1, beta2, eps, weight_decay, silent_sgd_phase (0 / 1), r,
 * weight_lr_power; state3 (device double[3]) = k, lr_max, weight_sum, advanced by the call; hyper9: device fp32[9] scratch.  The
 * rectification schedule (:138-152) runs on the device in double. */
int ftc_radam_sf_step_dev(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                          const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, const double* consts8,
                          double* state3, float* hyper9, void* stream);
int ftc_radam_sf_step(int n_chunks, const void* chunks, const void* const* ys, const void* const* grads,
                      const void* const* exp_avg_sqs, const void* const* zs, const int64_t* numels, double beta1, double beta2,
                      double bias_correction2, double eps, double weight_decay, double lr, double ckp1, int adam_step,
                      void* stream);

/* ---- train step: losses (loss_func.py) ----
 * ftc_heatmap_loss: loss_function :94-126 (the map part).  heatmap fp32 [B,9,H,W] logits, labelmap fp32 [B,5,H,W], idmap int64
 * [B,2,H,W]; losses9 (device fp32[9]) = keymap_loss (focal, x10), size_loss, textline_loss, separator_loss, code1/2/4/8_loss,
 * weight1_count.  scratch: ftc_heatmap_loss_scratch_bytes().  Deterministic (per-CTA double partials, one-CTA finish).
 * ftc_heatmap_loss_grad: grad[B,9,H,W] = d(sum_i alpha8[i] * losses9[i]) / d heatmap (alpha8: device fp32[8], e.g. CoV weights).
 * ftc_ce_rows: three-head residue cross entropy of `rows` rows (loss_function :128-161 on the fmask pixels, loss_function3
 * :179-213): out4 (device double[4]) = { sum_rows w * (ce0+ce1+ce2), sum_rows w, #rows with all three argmax == target % m,
 * #rows counted }; w = weight[row] (or 1) where select[row] (or all); hits/rows are counted where count_select[row] (or all). */
size_t ftc_heatmap_loss_scratch_bytes(void);
int ftc_heatmap_loss(const float* heatmap, const float* labelmap, const int64_t* idmap, int batch, int h, int w, float* losses9,
                     void* scratch, void* stream);
int ftc_heatmap_loss_grad(const float* heatmap, const float* labelmap, const int64_t* idmap, int batch, int h, int w,
                          const float* alpha8, const float* losses9, float* grad, void* stream);
int ftc_ce_rows(const float* logits0, const float* logits1, const float* logits2, int ld0, int ld1, int ld2, int m0, int m1, int m2,
                const int64_t* target, const float* weight, const unsigned char* select, const unsigned char* count_select, int rows,
                double* out4, void* stream);
/* backward of ftc_ce_rows' weighted mean  out4[0] / max(out4[1], 1)  w.r.t. the three logit matrices (dense fp32 [rows, m_i]):
 * grad_i[row, j] = coef * w_row * (softmax_i(row)[j] - [j == target % m_i]) on the selected rows, 0 elsewhere; coef: device fp32
 * scalar = upstream gradient / max(out4[1], 1)  (autograd of F.cross_entropy in loss_func.py:139-147, 191-197) */
int ftc_ce_rows_grad(const float* logits0, const float* logits1, const float* logits2, int ld0, int ld1, int ld2, int m0, int m1,
                     int m2, const int64_t* target, const float* weight, const unsigned char* select, int rows, const float* coef,
                     float* grad0, float* grad1, float* grad2, void* stream);

/* ---- single ops (unit-test / building-block entry points) ---- */
/* dense conv (k in {1,3}) or linear as implicit GEMM on NHWC activations.
 * x: [B,H,W,Cin] (dtype), w_oihw: fp32 [Cout,Cin,k,k] (packed on the fly into `wpack`),
 * scale/bias: fp32 [Cout] or NULL, residual: [B,Ho,Wo,Cout] (dtype) or NULL, out: [B,Ho,Wo,Cout] (dtype). */
int ftc_op_conv2d(const void* x, int dtype, int batch, int h, int w, int cin, const float* w_oihw, int cout, int ksize,
                  int stride, const float* scale, const float* bias, int act, const void* residual,
                  const float* a_scale, void* out, void* wpack, size_t wpack_bytes, int backend, void* stream);
/* data gradient of a stride-1 bf16 convolution on the tcgen05 kernel: dx [B,h,w,fwd_cin] = conv(dy [B,h,w,dy_ch], W') with W' =
 * the forward weight (fp32 OIHW [fwd_cout][fwd_cin][k][k], as nn.Conv2d stores it) read with the channel roles swapped and the taps
 * rotated 180 degrees INSIDE the weight-pack kernel (no flip / transpose / contiguous copies); dy_ch >= fwd_cout: extra (zero-padded)
 * dy channels meet zero weights.  wpack: ftc_op_conv2d_wpack_bytes(dy_ch, fwd_cin, ksize).  (torch autograd's conv backward-input of
 * train1.py:170 loss.backward().) */
int ftc_op_conv2d_dgrad(const void* dy, int batch, int h, int w, int dy_ch, const float* w_fwd_oihw, int fwd_cout, int fwd_cin,
                        int ksize, void* dx, void* wpack, size_t wpack_bytes, void* stream);
size_t ftc_op_conv2d_wpack_bytes(int cin, int cout, int ksize);
/* depthwise 3x3 + BN + SiLU (torchvision efficientnet.py:137-149).  se_sum (optional): fp32 [batch][tiles][c] receives one partial
 * spatial sum per CTA tile, tiles = ftc_op_dwconv3x3_tiles(h, w, stride, dtype); every entry is written, no atomics (the SE
 * squeeze is bit-reproducible), ftc_op_se_fc adds the tiles in order. */
int ftc_op_dwconv3x3_tiles(int h, int w, int stride, int dtype);
int ftc_op_dwconv3x3(const void* x, void* out, int dtype, int batch, int h, int w, int c, int stride,
                     const float* w9c, const float* scale, const float* bias, float* se_sum, void* stream);
/* SqueezeExcitation gate (ops/misc.py:251-261) from the tile sums above; hid: fp32 scratch [batch, s] */
int ftc_op_se_fc(const float* sum, int tiles, float* scale_out, float* hid, int batch, int c, int s, float inv_hw, const float* w1,
                 const float* b1, const float* w2t, const float* b2, void* stream);
/* debug: device buffer of 4096 u64 receiving clock64 stamps of CTA 0 of every following tcgen05 conv launch
 * ([0,1024) MMA full-wait start, [1024,2048) end, [2048,3072) producer empty-wait start, [3072,4096) end); NULL = off */
int ftc_debug_set_trace(void* dev_u64_4096);
/* debug / tuning: overrides of the tcgen05 GEMM launch heuristics (0 = automatic): mt 1|2 rows-per-tile/128, flags = ablation
 * bits, box_depth 1|2, plan_bn = forced n-tile (must divide N), no_bstat = 1 disables the weight-stationary schedule */
int ftc_debug_set_gemm_tuning(int mt, int flags, int box_depth, int plan_bn, int no_bstat);
/* debug / tuning: average CUDA-event time (ms, L2 flushed before each run) of one bf16 1x1-conv shape on the tcgen05 path:
 * out[batch*hw, n] = act(x[batch*hw, k] (* se[batch, k]) W^T * scale + bias) (+ res) */
int ftc_debug_bench_gemm(int batch, int hw, int k, int n, int act, int use_se, int use_res, int iters, float* ms_out);
/* same for a 3x3 stride-1 pad-1 conv on [batch, h, w, cin] */
int ftc_debug_bench_conv3x3(int batch, int h, int w, int cin, int cout, int act, int use_res, int iters, float* ms_out);
int ftc_op_upsample2x(const void* x, void* out, int dtype, int batch, int h, int w, int c, void* stream);
/* MBConv middle (torchvision efficientnet.py:137-149 + ops/misc.py:251-261), stride 1: depthwise 3x3 + BN + SiLU with the SE
 * squeeze and fc1 folded into the same kernel, then fc2 + sigmoid.  hid_part: fp32 scratch [batch, c / 32, s] (one fc1 share per
 * 32-channel CTA, added in a fixed order: no atomics, bit-reproducible); scale_out: fp32 [batch, c].  Returns an error for unsupported geometry
 * (needs w <= 48, w even, h % 8 == 0, c % 32 == 0). */
int ftc_op_dwconv3x3_se(const void* x, void* out, int dtype, int batch, int h, int w, int c, const float* w9c,
                        const float* scale, const float* bias, const float* w1, const float* b1, const float* w2t,
                        const float* b2, int s, float* hid_part, float* scale_out, void* stream);
/* Leafmap.top_conv of the 1-/2-channel heads (models/detector.py:188-190): y NHWC (dtype) with head i at channels
 * [i*192, i*192+192), w fp32 [sum(od)][9*192] tap-major, out NCHW fp32 [batch, sum(od), h, w] */
int ftc_op_head_top_conv(const void* y, int dtype, int pix_stride, int n_heads, const int* od, const float* w,
                         const float* bias, float* out, int batch, int h, int wd, void* stream);
/* softmax(q k^T / sqrt(hd) + mask) v per (batch, head) (models/transformer.py:133): q [B*Lt, q_stride] at column
 * q_off + h*hd, k/v [B*Ls, kv_stride] at k_off/v_off + h*hd, mask fp32 [B, Ls] additive or NULL, out [B*Lt, out_stride] */
int ftc_op_attention(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off, int v_off,
                     const float* mask, void* out, int out_stride, int dtype, int batch, int heads, int hd, int lt, int ls,
                     void* stream);


/* ---- train step: forward in train mode + backward of the detector blocks (train1.py:128-170 differentiates
 * models/detector.py:217-268 through torch autograd; here each autograd node is one of the entries below, see
 * findtextcenternet_b200/train_ops.py).  First correct path: CUDA-core kernels, fp32 math, NHWC fp32 / bf16 tensors (dtype =
 * FTC_PREC_*).  The forward convolutions of a train step are ftc_op_conv2d with scale = bias = NULL (raw output).
 *
 * BatchNorm in train mode (nn.BatchNorm2d / BatchNorm1d, torch batch_norm) over x [rows, c] (rows = B*H*W):
 *   ftc_train_bn_stats   mean[c], var[c] = batch mean and BIASED variance (fp32); scratch: ftc_train_reduce_scratch_bytes
 *   ftc_train_bn_act     y = act(gamma * (x - mean) * rsqrt(var + eps) + beta) (+ residual); act = none | SiLU | GELU(erf)
 *                        (torchvision Conv2dNormActivation ops/misc.py:69-126; Leafmap conv-BN-GELU models/detector.py:162-186)
 *   ftc_train_bn_act_bwd dbeta[c] = sum dz, dgamma[c] = sum dz * xhat, dx = gamma * rstd * (dz - dbeta/rows - xhat * dgamma/rows)
 *                        with dz = dy * act'(gamma * xhat + beta) recomputed from x (the residual's gradient is dy itself) */
size_t ftc_train_reduce_scratch_bytes(int64_t rows, int c);
int ftc_train_bn_stats(const void* x, int dtype, int64_t rows, int c, float* mean, float* var, void* scratch, void* stream);
/* the same with nn.BatchNorm's train-mode side effects folded into the finishing kernel (six tiny launches per layer otherwise):
 * running_mean = (1 - momentum) * running_mean + momentum * mean, running_var likewise with the UNBIASED batch variance,
 * *num_batches_tracked += 1 (may be NULL) */
int ftc_train_bn_stats_running(const void* x, int dtype, int64_t rows, int c, float* mean, float* var, void* scratch,
                               float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, void* stream);
int ftc_train_bn_act(const void* x, void* y, int dtype, int64_t rows, int c, const float* mean, const float* var,
                     const float* gamma, const float* beta, float eps, int act, const void* residual, void* stream);
int ftc_train_bn_act_bwd(const void* x, const void* dy, void* dx, int dtype, int64_t rows, int c, const float* mean,
                         const float* var, const float* gamma, const float* beta, float eps, int act, float* dbeta,
                         float* dgamma, void* scratch, void* stream);
/* the same with a ROW-STRIDED dy (dy_ld elements between rows, >= c): the gradient of a channel slice of a wider map -- what
 * torch.cat's backward hands to each of its inputs (Leafmap: cat([y, bn(x)]), models/detector.py:199) -- is consumed in place instead
 * of being copied into a contiguous tensor first.  bf16 with c % 32 == 0 (the stream kernels); otherwise dy_ld must equal c. */
int ftc_train_bn_act_bwd_ld(const void* x, const void* dy, int64_t dy_ld, void* dx, int dtype, int64_t rows, int c, const float* mean,
                            const float* var, const float* gamma, const float* beta, float eps, int act, float* dbeta,
                            float* dgamma, void* scratch, void* stream);
/* gradients of nn.Conv2d(k = 1 | 3, padding = (k-1)/2, stride = 1 | 2, bias-free): x [batch,h,w,cin], dy [batch,ho,wo,cout] NHWC;
 * weights and their gradient fp32 OIHW (the parameter's own layout).  wgrad OVERWRITES dw_oihw (fp32 atomics over pixel
 * splits); dgrad writes dx = conv_transpose(dy, w) (+ add, e.g. the gradient arriving over a residual connection). */
int ftc_train_conv2d_wgrad(const void* x, const void* dy, int dtype, int batch, int h, int w, int cin, int cout, int ksize,
                           int stride, float* dw_oihw, void* stream);
/* The same weight gradient with caller-provided scratch: bf16 stride-1 convolutions / linears (cin, cout multiples of 8) run on the
 * tcgen05 tensor cores -- dy and the input as MN-major UMMA operands straight from TMA tensor boxes, split over pixel tiles, fp32
 * partial tiles in `scratch` added in a fixed order (deterministic, no atomics); other shapes fall back to ftc_train_conv2d_wgrad.
 * scratch_bytes >= ftc_train_conv2d_wgrad_scratch_bytes(...) (0 = the tcgen05 kernel does not take the shape). */
size_t ftc_train_conv2d_wgrad_scratch_bytes(int batch, int h, int w, int cin, int cout, int ksize, int stride);
int ftc_train_conv2d_wgrad_ws(const void* x, const void* dy, int dtype, int batch, int h, int w, int cin, int cout, int ksize,
                              int stride, float* dw_oihw, void* scratch, size_t scratch_bytes, void* stream);
int ftc_train_conv2d_dgrad(const void* dy, int dtype, int batch, int h, int w, int cin, int cout, int ksize, int stride,
                           const float* w_oihw, const void* add, void* dx, void* stream);
/* depthwise 3x3 (pad 1, stride 1 | 2) of MBConv (efficientnet.py:136-147): raw forward, data gradient, weight gradient;
 * weights fp32 [9][c] tap-major (as ftc_op_dwconv3x3); h, w are the INPUT extents; wgrad overwrites dw9c */
int ftc_train_dwconv3x3(const void* x, void* y, int dtype, int batch, int h, int w, int c, int stride, const float* w9c,
                        void* stream);
int ftc_train_dwconv3x3_dgrad(const void* dy, void* dx, int dtype, int batch, int h, int w, int c, int stride, const float* w9c,
                              void* stream);
int ftc_train_dwconv3x3_wgrad(const void* x, const void* dy, int dtype, int batch, int h, int w, int c, int stride, float* dw9c,
                              void* stream);
/* SqueezeExcitation (torchvision ops/misc.py:225-261) on x [batch, hw, c]:
 *   ftc_train_spatial_sum  out[b][c] = scale * sum_hw x (* y when y != NULL): the squeeze mean (scale = 1/hw) and, in the
 *                          backward, dgate = sum_hw dy * x
 *   ftc_train_se_fc        hid_pre = W1 mean + b1, gate = sigmoid(W2 silu(hid_pre) + b2); W1 [sq][c], W2 [c][sq]
 *   ftc_train_scale_bc     y = x * scale_bc[b][c] (+ bias_mul * bias_bc[b][c]): the excitation, its data gradient
 *                          dx = dy * gate + dmean / hw, and StochasticDepth "row" mode (scale constant per image)
 *   ftc_train_se_fc_bwd    dgate -> dmean [batch][c] and dW1, db1, dW2, db2 (overwritten); dgp [batch][c], dhp [batch][sq] scratch */
int ftc_train_spatial_sum(const void* x, const void* y, int dtype, int batch, int hw, int c, float scale, float* out,
                          void* stream);
int ftc_train_scale_bc(const void* x, const float* scale_bc, const float* bias_bc, float bias_mul, void* y, int dtype, int batch,
                       int hw, int c, void* stream);
int ftc_train_se_fc(const float* mean, int batch, int c, int sq, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* hid_pre, float* gate, void* stream);
int ftc_train_se_fc_bwd(const float* dgate, const float* gate, const float* hid_pre, const float* mean, int batch, int c, int sq,
                        const float* w1, const float* w2, float* dgp, float* dhp, float* dmean, float* dw1, float* db1,
                        float* dw2, float* db2, void* stream);
/* adjoint of nn.UpsamplingBilinear2d(scale_factor=2) (align_corners=True, models/detector.py:167-186): dy [batch,2h,2w,c] ->
 * dx [batch,h,w,c] */
int ftc_train_upsample2x_bwd(const void* dy, void* dx, int dtype, int batch, int h, int w, int c, void* stream);
/* the same with a pixel-strided dy (dy_ld elements between pixels, >= c) */
int ftc_train_upsample2x_bwd_ld(const void* dy, int64_t dy_ld, void* dx, int dtype, int batch, int h, int w, int c, void* stream);

/* ---- train step of the Transformer (train3.py:132-137 differentiates models/transformer.py:58-253; dropout = 0 as in
 * ModelDimensions :257-264).  Linear layers are ftc_op_conv2d / ftc_train_conv2d_{wgrad,dgrad} on [rows,1,1,C] tensors.
 *   ftc_train_layernorm      y = LN(x (+ r1) (+ r2)) * gamma + beta over the last axis (nn.LayerNorm, eps 1e-5; the block outputs
 *                            LN(ff + _x + skip) :149-160,196-211); xs = the summed input (required when r1 / r2 are given; the
 *                            backward reads it), mean / rstd fp32 [rows]
 *   ftc_train_layernorm_bwd  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma (also the gradient of r1, r2);
 *                            dgamma[d] = sum_rows dy * xhat, dbeta[d] = sum_rows dy; scratch: ftc_train_reduce_scratch_bytes(rows, d)
 *   ftc_train_swiglu(_bwd)   h = x1 * silu(xg) (SwiGLU.forward :66-69) and dx1 = dh * silu(xg), dxg = dh * x1 * silu'(xg)
 *   ftc_train_embed3(_bwd)   out[row] = sum_i E_i[token[row] mod m_i] (Decoder.forward :226-233, fp32 tables [m_i][d]); the backward
 *                            OVERWRITES the three table gradients (zero + fp32 atomics)
 *   ftc_train_attention_bwd  backward of softmax(q k^T / sqrt(hd) + mask) v (F.scaled_dot_product_attention :133; forward =
 *                            ftc_op_attention): q / dout [batch*lt, heads*hd], k / v [batch*ls, heads*hd] (dtype), mask fp32
 *                            [batch, ls] additive or NULL; dq / dk / dv fp32, same shapes; scratch holds the recomputed
 *                            probabilities and dS: ftc_train_attention_bwd_scratch_bytes */
int ftc_train_layernorm(const void* x, const void* r1, const void* r2, void* xs, void* y, float* mean, float* rstd, int dtype,
                        int64_t rows, int d, const float* gamma, const float* beta, float eps, void* stream);
int ftc_train_layernorm_bwd(const void* xs, const void* dy, void* dx, const float* mean, const float* rstd, int dtype, int64_t rows,
                            int d, const float* gamma, float* dgamma, float* dbeta, void* scratch, void* stream);
int ftc_train_swiglu(const void* x1, const void* xg, void* h, int dtype, int64_t total, void* stream);
int ftc_train_swiglu_bwd(const void* x1, const void* xg, const void* dh, void* dx1, void* dxg, int dtype, int64_t total,
                         void* stream);
int ftc_train_embed3(const int64_t* tokens, const float* e0, const float* e1, const float* e2, int m0, int m1, int m2, void* out,
                     int dtype, int64_t rows, int d, void* stream);
int ftc_train_embed3_bwd(const int64_t* tokens, const void* dy, int dtype, int64_t rows, int d, int m0, int m1, int m2, float* de0,
                         float* de1, float* de2, void* stream);
size_t ftc_train_attention_bwd_scratch_bytes(int batch, int heads, int lt, int ls);
int ftc_train_attention_bwd(const void* q, const void* k, const void* v, const float* mask, const void* dout, float* dq, float* dk,
                            float* dv, void* scratch, int dtype, int batch, int heads, int hd, int lt, int ls, void* stream);

/* ---- page maps of run_detector (process_ocr_base.py:480-520), on the device: for every tile b (tile_meta as in ftc_peak_decode:
 * offset_x, offset_y, x_min, x_max, y_min, y_max) v = sigmoid(heat9[b, ch]) inside the validity window, 0 outside, and
 * page[m][offset_y/scale + y][offset_x/scale + x] = max(page, v) for m = key (ch 0), textline (3), separator (4), code1/2/4/8
 * (5-8).  page: fp32 [7][page_h4][page_w4], ZEROED by the caller; tiles may overlap (atomic maximum). */
int ftc_page_maps(const float* heat9, int batch, int h, int w, const int* tile_meta, float* page, int page_h4, int page_w4, int scale,
                  void* stream);

/* ---- page-level box selection of run_detector (process_ocr_base.py:540-650, imageHist :652-693), on the device ----
 * ftc_box_hists: for each of the n candidate boxes (loc fp32 [n][9] = p, cx, cy, w, h, c1, c2, c4, c8 as ftc_peak_decode writes
 * them) the reference's two histogram scores on the padded page (uint8 [page_h][page_w][3]): hists[0][i] = "loose" crop of the
 * threshold pass (:543-556, Python wrap-around slice semantics included), hists[1][i] = "tight" crop of the greedy pass (:571-576);
 * device double [2][n], bit-identical to numpy.  The threshold is median(hists[0]) / 5 (caller).
 * ftc_select_boxes: the greedy pass over `order` (candidate indices by descending score), the separator veto and the 3x3 code-map
 * maximum.  seps_all fp32 [h4][w4], code_all fp32 [4][h4][w4] (rows 2 and 3..6 of ftc_page_maps).  Outputs: *n_out accepted boxes,
 * sel_idx int32 [n] their candidate indices in acceptance order, out_loc fp32 [n][9], out_gf fp32 [n][feat_ch] (first *n_out rows).
 * One CTA (the greedy loop is sequential); scratch: ftc_select_boxes_scratch_bytes(n). */
int ftc_box_hists(const unsigned char* page, int page_h, int page_w, const float* loc, int n, double* hists, void* stream);
size_t ftc_select_boxes_scratch_bytes(int n);
int ftc_select_boxes(const float* loc, const float* gfeat, int feat_ch, const int* order, int n, const double* tight, double th,
                     const float* seps_all, const float* code_all, int h4, int w4, int scale, int* n_out, int* sel_idx, float* out_loc,
                     float* out_gf, void* scratch, size_t scratch_bytes, void* stream);

/* ---- train1 input pipeline on the device: the reference's per-sample Cython routine
 * dataset/processer.pyx::transform_crop (:260-454: inverse_partial :124-135, position rotation :345-355, center_map :137-163,
 * box_map :165-186, id_map :188-206, nearest / bilinear affine crop :387-409, textline / separator map crop :411-425), the blank
 * sample of process() (:662-666), random_salt of dataset/data_detector.py:17-26 and the colour compositing of
 * random_mono / random_single / random_double / random_background (:675-887), for a whole batch in four launches.
 * Every random decision is an INPUT (the host draws them in the reference's order, findtextcenternet_b200/dataset/processer.py),
 * so the same parameters give the reference's arrays: images, textline / separator maps, id maps and minsize bit for bit
 * (float32 operations rounded one by one, the double-promoted sub-expressions of the generated C evaluated in double),
 * the Gaussian centre map and the log-size maps to 1 ulp of libm's expf / logf.
 * ftc_crop_sample lives in DEVICE memory (array of `batch`); its pointers are device pointers. */
typedef struct {
  const unsigned char* image;     /* uint8 [im_h][im_w] page (ink high) */
  const unsigned char* textline;  /* uint8 [im_h2][im_w2] half-resolution text-line mask */
  const unsigned char* sepline;   /* uint8 [im_h2][im_w2] half-resolution separator mask */
  const unsigned char* bgimg;     /* color_mode 2: uint8 [bg_h][bg_w][3] background photograph */
  const unsigned char* salt;      /* optional uint8 [salt_h][salt_w] noise cells: 0 -> ink 0, 1 -> keep, 2 -> ink 1 (NaN cell) */
  int im_h, im_w, im_h2, im_w2;
  int box_begin, box_count;       /* this sample's rows of position / codelist */
  float rot[9];                   /* GetMatrix of the page (:308): rotates the box corners */
  float inv[9], inv2[9];          /* inverse affine of the page / of the half-resolution masks (:314-326) */
  int inv_i, inv_j, inv_h, inv_w; /* inverse_partial rectangle: rows [i, i + h), columns [j, j + w) are inverted */
  int cidx;                       /* box whose rotated centre anchors the crop (:359-364); box_count == 0: startx0 / starty0 */
  float woffset, hoffset, startx0, starty0;
  int nearest;                    /* 1: nearest-neighbour crop (:390-394) */
  int blank;                      /* 1: process()'s all-zero sample */
  int color_mode;                 /* 0: gray output [768][768]; 1: fg1 / fg2 (inside rect) over bg; 2: fg1 over the bgimg crop, clamped */
  float fg1[3], fg2[3], bg[3];
  int rect_top, rect_bottom, rect_left, rect_right;   /* random_double's inner rectangle (exclusive bounds); empty otherwise */
  int bg_h, bg_w, bg_startx, bg_starty;
  int salt_s, salt_h, salt_w;     /* cell size in pixels and grid shape */
} ftc_crop_sample;

int ftc_crop_sample_bytes(void);   /* sizeof(ftc_crop_sample): binding check */
size_t ftc_crop_scratch_bytes(int batch, int total_boxes);
/* position fp32 [total_boxes][4] (cx, cy, w, h in page pixels), codelist int32 [total_boxes][2]; out_image fp32
 * [batch][out_channels][768][768] (out_channels 1: every sample color_mode 0; 3: every sample color_mode 1 / 2), out_map fp32
 * [batch][5][192][192] (centre, log w, log h, textline, separator), out_idmap int32 [batch][2][192][192], out_minsize fp32 [batch]. */
int ftc_crop_batch(const ftc_crop_sample* samples, int batch, const float* position, const int* codelist, int total_boxes,
                   float* out_image, int out_channels, float* out_map, int* out_idmap, float* out_minsize, void* scratch,
                   size_t scratch_bytes, void* stream);

/* ---- random_distortion of the train1 input pipeline (dataset/data_detector.py:28-42) on the device, in place on the batch
 * image fp32 [batch][3][768][768]: (1) additive Gaussian noise im = float(double(im) + alpha * N(0,1)), clipped to [0,1]; (2) either a
 * Gaussian blur (scipy.ndimage.gaussian_filter semantics: the scalar sigma filters ALL THREE axes, the colour axis included, mode
 * 'reflect', radius int(4 sigma + 0.5), each axis pass accumulated in double in correlate1d's order -- centre tap, then symmetric pairs
 * from the outermost inwards -- and rounded to fp32 before the next axis), clipped, or (3) an unsharp mask im + k * (im - blur_5), clipped.
 * The random decisions are inputs (ftc_distort_sample, DEVICE array of `batch`); `weights` is a device double [batch][64] table, row b
 * = the normalised Gaussian taps w[0..radius] (w[0] = centre) the host computed as scipy does.  Noise comes from a counter-based
 * generator (Philox4x32-10, Box-Muller) keyed by noise_seed, or -- for parity tests -- from `noise` (device double
 * [batch][3][768][768] standard normals, may be NULL).  scratch: 2 * batch * 3 * 768 * 768 floats. */
typedef struct {
  int noise_on;                 /* 1: add noise */
  int mode;                     /* 0 none, 1 blur, 2 unsharp */
  int radius;                   /* taps each side (<= 63) */
  int pad_;
  double alpha;                 /* noise amplitude */
  unsigned long long noise_seed;
  float unsharp_k;              /* 10 * rng.random() as float32 */
  float pad2_;
} ftc_distort_sample;

size_t ftc_distort_scratch_bytes(int batch);
int ftc_distort_batch(float* image, int batch, const ftc_distort_sample* samples, const double* weights, const double* noise,
                      void* scratch, size_t scratch_bytes, void* stream);

/* debug / staging: route bf16 weight gradients (cin, cout multiples of 8) through the mma.sync kernel: 1 on, 0 off, -1 follow the
 * FTC_WGRAD_MMA environment variable (default; off when unset) */
int ftc_debug_set_wgrad_mma(int on);
/* debug / tuning: bf16 BatchNorm train kernels: 0 = stream kernels (default), 1 = the earlier one-row-per-trip vector kernels;
 * -1 = follow FTC_BN_UNROLL */
int ftc_debug_set_bn_unroll(int u);
/* debug / staging: the tcgen05 weight gradient: 0 off, 1 on (three N = 64 row-tap instructions per column shift), 2 on with the
 * three row taps fused into one N = 192 instruction, -1 follow the FTC_WGRAD_TC environment variable (default 2) */
int ftc_debug_set_wgrad_tc(int mode);

#ifdef __cplusplus
}
#endif
#endif /* FTC_B200_H */
