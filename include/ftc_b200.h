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
enum { FTC_GEMM_SIMT = 0, FTC_GEMM_TCGEN05 = 1 };

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

/* introspection (tests): device pointer of backbone tap `tap` (0..3 = x1..x4, NHWC, engine dtype) inside `workspace`
 * after a forward of `batch` images; *channels / *hw receive its shape. */
int ftc_detector_tap(const ftc_detector* d, int tap, int batch, void* workspace, void** ptr, int* channels, int* h, int* w);

/* ---- per-tile peak compaction + box decode (process_ocr_base.py:498-538) ----
 * tile_meta: int32 [B][6] = {offset_x, offset_y, mask_xmin, mask_xmax, mask_ymin, mask_ymax} (device)
 * count: int32 [B]; loc: fp32 [B][max_peaks][9]; gfeat: fp32 [B][max_peaks][F]; scratch: 8*B*max_peaks bytes */
int ftc_peak_decode(const float* heat9, const float* feat, int batch, int h, int w, int feat_ch, const int* tile_meta,
                    float cut_off, float page_w, float page_h, int max_peaks, int* count, float* loc, float* gfeat,
                    void* scratch, void* stream);
int ftc_peak_pick(const float* heat9, float* heat10, int batch, int h, int w, void* stream);

/* ---- single ops (unit-test / building-block entry points) ---- */
/* dense conv (k in {1,3}) or linear as implicit GEMM on NHWC activations.
 * x: [B,H,W,Cin] (dtype), w_oihw: fp32 [Cout,Cin,k,k] (packed on the fly into `wpack`),
 * scale/bias: fp32 [Cout] or NULL, residual: [B,Ho,Wo,Cout] (dtype) or NULL, out: [B,Ho,Wo,Cout] (dtype). */
int ftc_op_conv2d(const void* x, int dtype, int batch, int h, int w, int cin, const float* w_oihw, int cout, int ksize,
                  int stride, const float* scale, const float* bias, int act, const void* residual,
                  const float* a_scale, void* out, void* wpack, size_t wpack_bytes, int backend, void* stream);
size_t ftc_op_conv2d_wpack_bytes(int cin, int cout, int ksize);
int ftc_op_dwconv3x3(const void* x, void* out, int dtype, int batch, int h, int w, int c, int stride,
                     const float* w9c, const float* scale, const float* bias, float* se_sum, void* stream);
int ftc_op_se_fc(float* sum, float* scale_out, int batch, int c, int s, float inv_hw, const float* w1, const float* b1,
                 const float* w2t, const float* b2, void* stream);
int ftc_op_upsample2x(const void* x, void* out, int dtype, int batch, int h, int w, int c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FTC_B200_H */
