// C-ABI plumbing (error text, launch counter, version) and the single-op entry points of include/ftc_b200.h.
#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/ftc_b200.h"
#include "conv_gemm.cuh"
#include "detector_ops.cuh"
#include "transformer_ops.cuh"

namespace ftc {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { g_err = msg; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("FTC_NO_PDL"); v = (e && atoi(e) != 0) ? 0 : 1; }
  return v;
}

std::vector<uint32_t> make_ktab(int CA, int CB, int ksize, int* Kout) {
  int taps = ksize * ksize;
  int kreal = taps * (CA + CB);
  int K = (kreal + KBLOCK - 1) / KBLOCK * KBLOCK;
  std::vector<uint32_t> t(K / KCHUNK, 0u);
  for (int kc = 0; kc < kreal / KCHUNK; ++kc) {
    int k = kc * KCHUNK;
    if (k < taps * CA) {
      int tap = k / CA, c = k % CA;
      t[kc] = kt_make(false, tap / ksize, tap % ksize, c);
    } else {
      int k2 = k - taps * CA;
      int tap = k2 / CB, c = k2 % CB;
      t[kc] = kt_make(true, tap / ksize, tap % ksize, c);
    }
  }
  *Kout = K;
  return t;
}

}  // namespace ftc

using namespace ftc;

extern "C" {

int ftc_version(void) { return 100; }
const char* ftc_last_error(void) { return g_err.c_str(); }
int64_t ftc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

size_t ftc_peak_decode_scratch_bytes(int batch, int h, int w) { return peak_decode_scratch_bytes(batch, h, w); }

int ftc_peak_decode(const float* heat9, const float* feat, int batch, int h, int w, int feat_ch, const int* tile_meta,
                    float cut_off, float page_w, float page_h, int max_peaks, int* count, int* total, float* loc, float* gfeat,
                    void* scratch, void* stream) {
  FTC_REQUIRE(heat9 && feat && tile_meta && count && total && loc && gfeat && scratch && batch > 0, "bad argument");
  return peak_decode(heat9, feat, batch, h, w, feat_ch, tile_meta, cut_off, page_w, page_h, max_peaks, count, total, loc, gfeat,
                     scratch, (cudaStream_t)stream);
}

int ftc_peak_pick(const float* heat9, float* heat10, int batch, int h, int w, void* stream) {
  FTC_REQUIRE(heat9 && heat10 && batch > 0, "bad argument");
  return peak_pick(heat9, heat10, batch, h, w, (cudaStream_t)stream);
}

size_t ftc_op_conv2d_wpack_bytes(int cin, int cout, int ksize) {
  int K = 0;
  std::vector<uint32_t> kt = make_ktab(cin, 0, ksize, &K);
  int npad = (cout + 15) / 16 * 16;
  int Ktma = ksize * ksize * 64 * ((cin + 63) / 64);     // TMA paths pad the channels to whole 64-wide chunks
  if (Ktma > K) K = Ktma;
  return align_up((size_t)npad * K * 4, 256) + align_up(kt.size() * 4, 256) + 256;
}

static int op_conv2d_impl(const void* x, int dtype, int batch, int h, int w, int cin, const float* w_oihw, int cout, int ksize,
                          int stride, const float* scale, const float* bias, int act, const void* residual,
                          const float* a_scale, void* out, void* wpack, size_t wpack_bytes, int backend, void* stream, int dgrad_rows);

int ftc_op_conv2d(const void* x, int dtype, int batch, int h, int w, int cin, const float* w_oihw, int cout, int ksize,
                  int stride, const float* scale, const float* bias, int act, const void* residual,
                  const float* a_scale, void* out, void* wpack, size_t wpack_bytes, int backend, void* stream) {
  return op_conv2d_impl(x, dtype, batch, h, w, cin, w_oihw, cout, ksize, stride, scale, bias, act, residual, a_scale, out, wpack,
                        wpack_bytes, backend, stream, 0);
}

int ftc_op_conv2d_dgrad(const void* dy, int batch, int h, int w, int dy_ch, const float* w_fwd_oihw, int fwd_cout, int fwd_cin,
                        int ksize, void* dx, void* wpack, size_t wpack_bytes, void* stream) {
  FTC_REQUIRE(fwd_cout > 0 && fwd_cout <= dy_ch, "ftc_op_conv2d_dgrad: dy has fewer channels than the forward convolution's outputs");
  return op_conv2d_impl(dy, DT_BF16, batch, h, w, dy_ch, w_fwd_oihw, fwd_cin, ksize, 1, nullptr, nullptr, ACT_NONE, nullptr, nullptr, dx,
                        wpack, wpack_bytes, FTC_GEMM_TCGEN05, stream, fwd_cout);
}

static int op_conv2d_impl(const void* x, int dtype, int batch, int h, int w, int cin, const float* w_oihw, int cout, int ksize,
                          int stride, const float* scale, const float* bias, int act, const void* residual,
                          const float* a_scale, void* out, void* wpack, size_t wpack_bytes, int backend, void* stream, int dgrad_rows) {
  FTC_REQUIRE(x && w_oihw && out && wpack, "null argument");
  FTC_REQUIRE(dgrad_rows == 0 || backend == FTC_GEMM_TCGEN05, "data-gradient weight form is packed by the tcgen05 path only");
  FTC_REQUIRE(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
  FTC_REQUIRE(cin % 8 == 0, "cin must be a multiple of 8");
  FTC_REQUIRE(wpack_bytes >= ftc_op_conv2d_wpack_bytes(cin, cout, ksize), "wpack too small");
  FTC_REQUIRE(a_scale == nullptr || ksize == 1, "a_scale only for 1x1");
  cudaStream_t s = (cudaStream_t)stream;
  int K = 0;
  std::vector<uint32_t> kt = make_ktab(cin, 0, ksize, &K);
  const size_t es = dtype == DT_BF16 ? 2 : 4;
  const int npad = (cout + 15) / 16 * 16;
  char* wp = (char*)wpack;
  const int Kmax = std::max(K, ksize * ksize * 64 * ((cin + 63) / 64));
  // the im2col k-table depends on (cin, ksize) only: one device copy per (device, cin, ksize), made at first use and kept for
  // the life of the library -- the call itself then neither copies from pageable host memory nor synchronises, so a whole train
  // step of these calls can be captured into a CUDA graph (first uses must happen before the capture: warm-up step)
  const uint32_t* ktab_d = nullptr;
  {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, uint32_t*> cache;
    int devid = 0;
    FTC_CHECK_CUDA(cudaGetDevice(&devid));
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(devid, cin, ksize);
    auto it = cache.find(key);
    if (it == cache.end()) {
      uint32_t* d = nullptr;
      FTC_CHECK_CUDA(cudaMalloc((void**)&d, kt.size() * 4));
      FTC_CHECK_CUDA(cudaMemcpy(d, kt.data(), kt.size() * 4, cudaMemcpyHostToDevice));
      it = cache.emplace(key, d).first;
    }
    ktab_d = it->second;
  }
  FTC_CHECK_CUDA(cudaMemsetAsync(wp, 0, (size_t)npad * Kmax * es, s));
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.H = h; p.W = w; p.stride = stride; p.pad = (ksize - 1) / 2;
  p.Ho = (h - 1) / stride + 1; p.Wo = (w - 1) / stride + 1;
  p.M = batch * p.Ho * p.Wo; p.N = cout; p.G = 1; p.K = K;
  p.srcA = x; p.a_pix_stride = cin; p.ktab = ktab_d; p.CA = cin; p.CB = 0;
  p.a_scale = a_scale; p.a_scale_stride = cin;
  p.w = wp; p.scale = scale; p.bias_tab = bias; p.ncase = 1; p.act = act;
  p.res1 = residual; p.res1_stride = cout;
  p.out = out; p.out_layout = OUT_NHWC; p.out_stride = cout;
  p.out_ch_base[0] = 0; p.n_valid[0] = cout;
  p.dtype = dtype;
  int rc;
  if (backend == FTC_GEMM_TCGEN05 || backend == FTC_GEMM_TCGEN05_IM2COL) {
    FTC_REQUIRE(dtype == DT_BF16, "tcgen05 backend needs bf16");
    ConvTcPlan plan;
    rc = conv_gemm_tc_plan(p, &plan, backend == FTC_GEMM_TCGEN05);
    if (rc) return rc;
    p.K = plan.NKB * KBLOCK;
    rc = pack_conv_weight_tc(wp, w_oihw, cout, cin, ksize, ksize, 0, cin, 0, p.K, 0, plan.BN, nullptr, s,
                             plan.tma == TMA_HALO ? (plan.kb32 ? 2 : 1) : 0, dgrad_rows);
    if (rc) return rc;
    p.tc = plan;
    rc = conv_gemm_tc(p, s);
  } else {
    rc = pack_conv_weight(wp, dtype, w_oihw, cout, cin, ksize, ksize, 0, cin, 0, K, 0, nullptr, s);
    if (rc) return rc;
    rc = conv_gemm_simt(p, s);
  }
  return rc;
}

int ftc_op_dwconv3x3(const void* x, void* out, int dtype, int batch, int h, int w, int c, int stride,
                     const float* w9c, const float* scale, const float* bias, float* se_sum, void* stream) {
  FTC_REQUIRE(x && out && w9c && scale && bias, "null argument");
  return dwconv3x3(x, out, dtype, batch, h, w, c, stride, w9c, scale, bias, se_sum, (cudaStream_t)stream);
}

int ftc_op_dwconv3x3_tiles(int h, int w, int stride, int dtype) { return dwconv3x3_tiles(h, w, stride, dtype); }

int ftc_op_se_fc(const float* sum, int tiles, float* scale_out, float* hid, int batch, int c, int s, float inv_hw, const float* w1,
                 const float* b1, const float* w2t, const float* b2, void* stream) {
  FTC_REQUIRE(sum && scale_out && hid && w1 && b1 && w2t && b2 && tiles >= 1, "null argument");
  return se_fc(sum, tiles, scale_out, hid, batch, c, s, inv_hw, w1, b1, w2t, b2, (cudaStream_t)stream);
}

int ftc_mask_predict_step(const float* logits, int ld, int head_ld, const int64_t* dec_in, int64_t* ids, float* prob,
                          int64_t* next_in, int* flags, int rows, void* stream) {
  FTC_REQUIRE(logits && dec_in && ids && prob && next_in && flags && rows > 0, "bad argument");
  return mask_predict_step(logits, ld, head_ld, dec_in, ids, prob, next_in, flags, rows, (cudaStream_t)stream);
}

int ftc_debug_set_gemm_tuning(int mt, int flags, int box_depth, int plan_bn, int no_bstat) {
  GemmTuning& t = gemm_tuning();
  t.mt = mt; t.flags = flags & 0xFFFFFF; t.box_depth = box_depth; t.plan_bn = plan_bn; t.no_bstat = no_bstat & 1;
  t.epi8 = (no_bstat >> 1) & 1;   // bit 1 of no_bstat: keep 8 epilogue warps
  t.nb = (no_bstat >> 4) & 15;    // bits 4-7: forced weight-ring depth of the rows path (0 = automatic)
  return 0;
}

// timing harness for one tcgen05 1x1-conv / linear shape (debug; allocates and frees its own buffers):
// out[M,N] = act(x[M,K] (* se[b,K]) W^T * scale + bias) (+ res); M = batch * hw rows
static int debug_bench_conv(int batch, int h, int w, int ksize, int k, int n, int act, int use_se, int use_res, int iters, float* ms_out);

int ftc_debug_bench_gemm(int batch, int hw, int k, int n, int act, int use_se, int use_res, int iters, float* ms_out) {
  return debug_bench_conv(batch, hw, 1, 1, k, n, act, use_se, use_res, iters, ms_out);
}

int ftc_debug_bench_conv3x3(int batch, int h, int w, int cin, int cout, int act, int use_res, int iters, float* ms_out) {
  return debug_bench_conv(batch, h, w, 3, cin, cout, act, 0, use_res, iters, ms_out);
}

static int debug_bench_conv(int batch, int h, int w, int ksize, int k, int n, int act, int use_se, int use_res, int iters, float* ms_out) {
  const int hw = h * w;
  FTC_REQUIRE(batch > 0 && hw > 0 && k % 8 == 0 && n > 0 && iters > 0 && ms_out, "bad argument");
  const int M = batch * hw;
  void *x = nullptr, *out = nullptr, *res = nullptr, *wp = nullptr, *flush = nullptr;
  float *wf = nullptr, *sc = nullptr, *bi = nullptr, *se = nullptr;
  const size_t flush_bytes = 256u << 20;
  FTC_CHECK_CUDA(cudaMalloc(&x, (size_t)M * k * 2)); FTC_CHECK_CUDA(cudaMalloc(&out, (size_t)M * n * 2));
  FTC_CHECK_CUDA(cudaMalloc(&res, (size_t)M * n * 2)); FTC_CHECK_CUDA(cudaMalloc(&wf, (size_t)n * k * 9 * 4));
  FTC_CHECK_CUDA(cudaMalloc(&sc, (size_t)n * 4)); FTC_CHECK_CUDA(cudaMalloc(&bi, (size_t)n * 4));
  FTC_CHECK_CUDA(cudaMalloc(&se, (size_t)batch * k * 4)); FTC_CHECK_CUDA(cudaMalloc(&flush, flush_bytes));
  FTC_CHECK_CUDA(cudaMemset(x, 0x3c, (size_t)M * k * 2)); FTC_CHECK_CUDA(cudaMemset(res, 0x3c, (size_t)M * n * 2));
  FTC_CHECK_CUDA(cudaMemset(wf, 0x3b, (size_t)n * k * 9 * 4)); FTC_CHECK_CUDA(cudaMemset(sc, 0x3c, (size_t)n * 4));
  FTC_CHECK_CUDA(cudaMemset(bi, 0x3c, (size_t)n * 4)); FTC_CHECK_CUDA(cudaMemset(se, 0x3c, (size_t)batch * k * 4));
  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = batch; p.H = h; p.W = w; p.stride = 1; p.pad = (ksize - 1) / 2; p.Ho = h; p.Wo = w;
  p.M = M; p.N = n; p.G = 1;
  p.srcA = x; p.a_pix_stride = k; p.CA = k; p.CB = 0;
  int K = 0;
  std::vector<uint32_t> kt = make_ktab(k, 0, ksize, &K);
  p.K = K;
  ConvTcPlan plan;
  int rc = conv_gemm_tc_plan(p, &plan, true);
  if (rc) return rc;
  p.K = plan.NKB * KBLOCK;
  const size_t wbytes = conv_tc_weight_bytes(plan, 1);
  uint32_t* ktd = nullptr;
  FTC_CHECK_CUDA(cudaMalloc(&wp, wbytes)); FTC_CHECK_CUDA(cudaMalloc((void**)&ktd, kt.size() * 4));
  FTC_CHECK_CUDA(cudaMemset(wp, 0, wbytes));
  FTC_CHECK_CUDA(cudaMemcpy(ktd, kt.data(), kt.size() * 4, cudaMemcpyHostToDevice));
  rc = pack_conv_weight_tc(wp, wf, n, k, ksize, ksize, 0, k, 0, p.K, 0, plan.BN, nullptr, 0, plan.tma == TMA_HALO ? (plan.kb32 ? 2 : 1) : 0);
  if (rc) return rc;
  p.ktab = ktd; p.w = wp; p.scale = sc; p.bias_tab = bi; p.ncase = 1; p.act = act;
  p.a_scale = use_se ? se : nullptr; p.a_scale_stride = k;
  p.res1 = use_res ? res : nullptr; p.res1_stride = n;
  p.out = out; p.out_layout = OUT_NHWC; p.out_stride = n; p.out_ch_base[0] = 0; p.n_valid[0] = n;
  p.dtype = DT_BF16; p.tc = plan;
  cudaEvent_t e0, e1;
  FTC_CHECK_CUDA(cudaEventCreate(&e0)); FTC_CHECK_CUDA(cudaEventCreate(&e1));
  float total = 0.f;
  for (int it = -2; it < iters && rc == 0; ++it) {
    FTC_CHECK_CUDA(cudaMemsetAsync(flush, it & 1, flush_bytes, 0));    // evict: the real producer's output is larger than L2
    FTC_CHECK_CUDA(cudaEventRecord(e0, 0));
    rc = conv_gemm_tc(p, 0);
    FTC_CHECK_CUDA(cudaEventRecord(e1, 0));
    FTC_CHECK_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    FTC_CHECK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (it >= 0) total += ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(x); cudaFree(out); cudaFree(res); cudaFree(wf); cudaFree(sc); cudaFree(bi); cudaFree(se); cudaFree(flush); cudaFree(wp); cudaFree(ktd);
  if (rc) return rc;
  *ms_out = total / iters;
  return 0;
}

int ftc_debug_set_trace(void* dev_u64_4096) {
  conv_gemm_tc_set_trace((unsigned long long*)dev_u64_4096);
  return 0;
}

int ftc_op_dwconv3x3_se(const void* x, void* out, int dtype, int batch, int h, int w, int c, const float* w9c,
                        const float* scale, const float* bias, const float* w1, const float* b1, const float* w2t,
                        const float* b2, int s, float* hid_part, float* scale_out, void* stream) {
  FTC_REQUIRE(x && out && w9c && scale && bias && w1 && b1 && w2t && b2 && hid_part && scale_out, "null argument");
  FTC_REQUIRE(dwconv3x3_se_supported(h, w, c, 1), "dwconv3x3_se: unsupported geometry");
  int rc = dwconv3x3_se(x, out, dtype, batch, h, w, c, w9c, scale, bias, w1, s, hid_part, (cudaStream_t)stream);
  if (rc) return rc;
  return se_fc2_hid(hid_part, c / 32, scale_out, batch, c, s, b1, w2t, b2, (cudaStream_t)stream);
}

int ftc_op_head_top_conv(const void* y, int dtype, int pix_stride, int n_heads, const int* od, const float* w,
                         const float* bias, float* out, int batch, int h, int wd, void* stream) {
  FTC_REQUIRE(y && od && w && bias && out && n_heads >= 1 && n_heads <= 8, "bad argument");
  int out_ch = 0;
  for (int i = 0; i < n_heads; ++i) out_ch += od[i];
  return head_top_conv(y, dtype, pix_stride, 0, n_heads, od, w, bias, out, out_ch, batch, h, wd, (cudaStream_t)stream);
}

int ftc_op_attention(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off, int v_off,
                     const float* mask, void* out, int out_stride, int dtype, int batch, int heads, int hd, int lt, int ls,
                     void* stream) {
  FTC_REQUIRE(q && k && v && out && batch > 0 && heads > 0 && lt > 0 && ls > 0, "bad argument");
  return attention(q, q_stride, q_off, k, v, kv_stride, k_off, v_off, mask, out, out_stride, dtype, batch, heads, hd, lt, ls,
                   (cudaStream_t)stream);
}

int ftc_op_upsample2x(const void* x, void* out, int dtype, int batch, int h, int w, int c, void* stream) {
  FTC_REQUIRE(x && out, "null argument");
  return upsample2x(x, out, dtype, batch, h, w, c, (cudaStream_t)stream);
}

}  // extern "C"
