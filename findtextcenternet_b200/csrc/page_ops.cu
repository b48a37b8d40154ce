// Page-level box selection of run_detector on the device (process_ocr_base.py:540-650 + imageHist :652-693), SURVEY.md 8 row f1.
//
//   box_hist_kernel      per candidate box the two histogram scores of the reference: "loose" (threshold pass :543-556, crop bounds
//                        int(c -+ s/2) - 1 / + 2 with Python's wrap-around slice semantics) and "tight" (greedy pass :571-576, crop
//                        clipped to the page): 256-bin histograms of the three colour channels in shared memory, then the
//                        two-means gap of imageHist.cluster_dist per channel, one warp per channel, in double precision with the
//                        reference's operation order (integer sums are exact, every division is IEEE double: bit-identical).
//   select_boxes_kernel  the greedy pass (:559-619): candidates in descending score; a candidate is dropped if its tight score is
//                        below the threshold, if its IoU with an accepted box exceeds 0.5, if an intersection exceeds 75 % of its
//                        area, or if the accepted boxes it touches cover more than half of its int(w) x int(h) pixel grid.  The
//                        loop over candidates is inherently sequential (one CTA); each step is parallel over the accepted boxes
//                        (overlap tests, block max-reduce) and over the grid cells (exact coverage count).  Then the separator
//                        veto (:621-631) and the 3x3 maximum of the page code maps (:641-658).
// Double arithmetic uses the non-contracting intrinsics (__dadd_rn ...): an FMA would round differently from numpy.
// Plain SIMT on purpose: oracle/emu compiles this file for host threads and tests/test_emu_kernels.py runs it against the
// reference-pinned oracle (oracle/detector_oracle.py::select_boxes) on the CPU.
#include "../../include/ftc_b200.h"
#include "common.cuh"

#ifdef FTC_EMU
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
#endif

namespace ftc {
namespace {

// Python slice [lo:hi] on an axis of length S -> [lo', hi') (negative bounds wrap around once, everything is clamped)
__device__ __forceinline__ void py_slice(long long lo, long long hi, int S, int* a, int* b) {
  if (lo < 0) { lo += S; if (lo < 0) lo = 0; } else if (lo > S) lo = S;
  if (hi < 0) { hi += S; if (hi < 0) hi = 0; } else if (hi > S) hi = S;
  *a = (int)lo; *b = (int)(hi > lo ? hi : lo);
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// imageHist.cluster_dist on one 256-bin histogram; executed by one full warp, lane l owns bins [8 l, 8 l + 8)
__device__ double two_means_gap_warp(const int* __restrict__ hist, int lane) {
  long long h[8], m[8];
  long long tot = 0, mass = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h[i] = hist[lane * 8 + i];
    m[i] = h[i] * (long long)(lane * 8 + i);
    tot += h[i]; mass += m[i];
  }
  tot = warp_sum(tot); mass = warp_sum(mass);
  if (tot == 0) return 0.0;
  const int split = (int)(__dadd_rn(__ddiv_rn((double)mass, (double)tot), 0.5));
  long long lo_n = 0, hi_n = 0, lo_m = 0, hi_m = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (lane * 8 + i < split) { lo_n += h[i]; lo_m += m[i]; } else { hi_n += h[i]; hi_m += m[i]; }
  }
  lo_n = warp_sum(lo_n); hi_n = warp_sum(hi_n); lo_m = warp_sum(lo_m); hi_m = warp_sum(hi_m);
  if (lo_n == 0 || hi_n == 0) return 0.0;
  double c_lo = __ddiv_rn((double)lo_m, (double)lo_n), c_hi = __ddiv_rn((double)hi_m, (double)hi_n);
  double prev = 256.0, cur = fabs(__dsub_rn(c_lo, c_hi));
  for (int it = 0; it < 4096 && prev != cur; ++it) {
    prev = cur;
    lo_n = hi_n = lo_m = hi_m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double bin = (double)(lane * 8 + i);
      if (fabs(__dsub_rn(bin, c_lo)) < fabs(__dsub_rn(bin, c_hi))) { lo_n += h[i]; lo_m += m[i]; } else { hi_n += h[i]; hi_m += m[i]; }
    }
    lo_n = warp_sum(lo_n); hi_n = warp_sum(hi_n); lo_m = warp_sum(lo_m); hi_m = warp_sum(hi_m);
    if (lo_n == 0 || hi_n == 0) return 0.0;
    c_lo = __ddiv_rn((double)lo_m, (double)lo_n); c_hi = __ddiv_rn((double)hi_m, (double)hi_n);
    cur = fabs(__dsub_rn(c_lo, c_hi));
  }
  return prev;
}

// grid (n boxes, 2 variants); page: uint8 [H][W][3]; loc: fp32 [n][9] (p, cx, cy, w, h, ...); out: double [2][n]
__global__ void __launch_bounds__(128) box_hist_kernel(const unsigned char* __restrict__ page, int H, int W, const float* __restrict__ loc,
                                                       int n, double* __restrict__ out) {
  __shared__ int hist[3][256];
  __shared__ double gap[3];
  const int i = blockIdx.x, variant = blockIdx.y, tid = threadIdx.x;
  for (int k = tid; k < 3 * 256; k += blockDim.x) (&hist[0][0])[k] = 0;
  __syncthreads();
  const double cx = (double)loc[i * 9 + 1], cy = (double)loc[i * 9 + 2], w = (double)loc[i * 9 + 3], h = (double)loc[i * 9 + 4];
  const double hw = __dmul_rn(w, 0.5), hh = __dmul_rn(h, 0.5);         // w / 2: exact either way
  long long x0 = (long long)__dsub_rn(cx, hw), x1 = (long long)__dadd_rn(cx, hw);
  long long y0 = (long long)__dsub_rn(cy, hh), y1 = (long long)__dadd_rn(cy, hh);
  if (variant == 0) { x0 -= 1; x1 += 2; y0 -= 1; y1 += 2; }
  else {
    x0 = x0 > 0 ? x0 : 0; y0 = y0 > 0 ? y0 : 0;
    x1 = (x1 + 1 < W - 1) ? x1 + 1 : W - 1;
    y1 = (y1 + 1 < H - 1) ? y1 + 1 : H - 1;
  }
  int xa, xb, ya, yb;
  py_slice(x0, x1, W, &xa, &xb);
  py_slice(y0, y1, H, &ya, &yb);
  const int bw = xb - xa, bh = yb - ya;
  const long long npx = (long long)bw * bh;
  for (long long k = tid; k < npx; k += blockDim.x) {
    const int yy = ya + (int)(k / bw), xx = xa + (int)(k % bw);
    const unsigned char* px = page + ((long long)yy * W + xx) * 3;
    atomicAdd(&hist[0][px[0]], 1); atomicAdd(&hist[1][px[1]], 1); atomicAdd(&hist[2][px[2]], 1);
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < 3) {
    const double g = two_means_gap_warp(hist[warp], lane);
    if (lane == 0) gap[warp] = g;
  }
  __syncthreads();
  if (tid == 0) {
    double best = -1.0;
    for (int c = 0; c < 3; ++c) best = gap[c] > best ? gap[c] : best;
    out[(long long)variant * n + i] = best;
  }
}

constexpr int SB_THREADS = 256;

struct SelectArgs {
  const float* loc;           // [n][9]
  const float* gfeat;         // [n][fc]
  const int* order;           // [n] candidate indices in descending score (ties: ascending index)
  const double* tight;        // [n]
  double th;                  // median(loose) / 5 (NaN when there are no candidates: nothing is filtered)
  const float* seps;          // [h4][w4]
  const float* code;          // [4][h4][w4]
  int n, fc, h4, w4, scale;
  int* n_out;                 // [1]
  int* sel_idx;               // [n] accepted candidate indices in acceptance order, then compacted after the separator veto
  float* out_loc;             // [n][9]
  float* out_gf;              // [n][fc]
  double* acc;                // scratch [n][4]: cx, cy, w, h of the accepted boxes
  int* rects;                 // scratch [n][4]: a0, a1, b0, b1 of the accepted boxes the current candidate touches
  int* keep;                  // scratch [n]
};

__device__ __forceinline__ double block_max(double v, double* red, int tid) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double r = red[0];
  for (int k = 1; k < SB_THREADS / 32; ++k) r = red[k] > r ? red[k] : r;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SB_THREADS) select_boxes_kernel(const SelectArgs a) {
  __shared__ double red[SB_THREADS / 32];
  __shared__ int s_nrect, s_cov, s_m;
  const int tid = threadIdx.x;
  if (tid == 0) s_m = 0;
  __syncthreads();
  for (int oi = 0; oi < a.n; ++oi) {
    const int i = a.order[oi];
    if (a.tight[i] < a.th) continue;                     // warp-uniform: every thread reads the same values
    const double cx = (double)a.loc[i * 9 + 1], cy = (double)a.loc[i * 9 + 2], w = (double)a.loc[i * 9 + 3], h = (double)a.loc[i * 9 + 4];
    const double x_lo = __dsub_rn(cx, __dmul_rn(w, 0.5)), x_hi = __dadd_rn(cx, __dmul_rn(w, 0.5));
    const double y_lo = __dsub_rn(cy, __dmul_rn(h, 0.5)), y_hi = __dadd_rn(cy, __dmul_rn(h, 0.5));
    const double area0 = __dmul_rn(w, h);
    const int m = s_m;
    if (tid == 0) { s_nrect = 0; s_cov = 0; }
    __syncthreads();
    double my_iou = 0.0, my_inter = 0.0;
    for (int j = tid; j < m; j += SB_THREADS) {
      const double cx1 = a.acc[j * 4 + 0], cy1 = a.acc[j * 4 + 1], w1 = a.acc[j * 4 + 2], h1 = a.acc[j * 4 + 3];
      const double lx = fmax(x_lo, __dsub_rn(cx1, __dmul_rn(w1, 0.5))), ly = fmax(y_lo, __dsub_rn(cy1, __dmul_rn(h1, 0.5)));
      const double hx = fmin(x_hi, __dadd_rn(cx1, __dmul_rn(w1, 0.5))), hy = fmin(y_hi, __dadd_rn(cy1, __dmul_rn(h1, 0.5)));
      const double iw = fmax(__dsub_rn(hx, lx), 0.0), ih = fmax(__dsub_rn(hy, ly), 0.0);
      const double inter = __dmul_rn(iw, ih);
      const double uni = __dsub_rn(__dadd_rn(area0, __dmul_rn(w1, h1)), inter);
      const double iou = uni > 0.0 ? __ddiv_rn(inter, uni) : 0.0;
      my_iou = iou > my_iou ? iou : my_iou;
      my_inter = inter > my_inter ? inter : my_inter;
      if (iou > 0.0) {
        const int r = atomicAdd(&s_nrect, 1);
        a.rects[r * 4 + 0] = (int)__dsub_rn(lx, x_lo);
        a.rects[r * 4 + 1] = (int)__dsub_rn(hx, x_lo) + 1;
        a.rects[r * 4 + 2] = (int)__dsub_rn(ly, y_lo);
        a.rects[r * 4 + 3] = (int)__dsub_rn(hy, y_lo) + 1;
      }
    }
    const double max_iou = block_max(my_iou, red, tid);
    const double max_inter = block_max(my_inter, red, tid);      // (the barriers inside also publish s_nrect / rects)
    bool drop = m > 0 && (max_iou > 0.5 || max_inter > __dmul_rn(area0, 0.75));
    if (!drop && m > 0) {
      const int nrect = s_nrect;
      const int gw = (int)w, gh = (int)h;
      const long long cells = (long long)gw * gh;
      if (nrect > 0 && cells > 0) {
        int cov = 0;
        for (long long c = tid; c < cells; c += SB_THREADS) {
          const int ga = (int)(c / gh), gb = (int)(c % gh);
          bool hit = false;
          for (int r = 0; r < nrect && !hit; ++r)
            hit = ga >= a.rects[r * 4 + 0] && ga < a.rects[r * 4 + 1] && gb >= a.rects[r * 4 + 2] && gb < a.rects[r * 4 + 3];
          cov += hit ? 1 : 0;
        }
        cov = warp_sum(cov);
        if ((tid & 31) == 0 && cov) atomicAdd(&s_cov, cov);
        __syncthreads();
        drop = 2LL * s_cov > cells;                      // np.mean(fill_map) > 0.5
      }
    }
    __syncthreads();
    if (!drop) {
      if (tid == 0) {
        a.acc[m * 4 + 0] = cx; a.acc[m * 4 + 1] = cy; a.acc[m * 4 + 2] = w; a.acc[m * 4 + 3] = h;
        a.sel_idx[m] = i;
        s_m = m + 1;
      }
    }
    __syncthreads();
  }
  // ---- separator veto (centre pixel of the quarter-resolution separator map > 0.5), order-preserving compaction
  const int m = s_m;
  for (int j = tid; j < m; j += SB_THREADS) {
    const int i = a.sel_idx[j];
    const double cx = (double)a.loc[i * 9 + 1], cy = (double)a.loc[i * 9 + 2];
    const int x = (int)__ddiv_rn(cx, (double)a.scale), y = (int)__ddiv_rn(cy, (double)a.scale);
    bool ok = true;
    if (x >= 0 && x < a.w4 && y >= 0 && y < a.h4) ok = !(a.seps[(long long)y * a.w4 + x] > 0.5f);
    a.keep[j] = ok ? 1 : 0;
  }
  __syncthreads();
  if (tid == 0) {
    int k = 0;
    for (int j = 0; j < m; ++j)
      if (a.keep[j]) { a.keep[k] = a.sel_idx[j]; ++k; }
    s_m = k;
    *a.n_out = k;
  }
  __syncthreads();
  const int kept = s_m;
  for (int j = tid; j < kept; j += SB_THREADS) a.sel_idx[j] = a.keep[j];
  __syncthreads();
  // ---- output rows: code probabilities := max(own, 3x3 neighbourhood of the page code maps); glyph features gathered
  for (int j = tid; j < kept; j += SB_THREADS) {
    const int i = a.sel_idx[j];
    float row[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) row[c] = a.loc[i * 9 + c];
    const double cx = (double)row[1], cy = (double)row[2];
    const double qx = __ddiv_rn(cx, (double)a.scale), qy = __ddiv_rn(cy, (double)a.scale);
    const int x = (int)qx, y = (int)qy;
    if (x >= 0 && x < a.w4 && y >= 0 && y < a.h4) {
      int x0 = (int)__dsub_rn(qx, 1.0), y0 = (int)__dsub_rn(qy, 1.0);
      int x1 = (int)__dadd_rn(qx, 1.0) + 1, y1 = (int)__dadd_rn(qy, 1.0) + 1;
      x0 = x0 > 0 ? x0 : 0; y0 = y0 > 0 ? y0 : 0;
      x1 = x1 < a.w4 ? x1 : a.w4; y1 = y1 < a.h4 ? y1 : a.h4;
      for (int k = 0; k < 4; ++k) {
        const float* cm = a.code + (long long)k * a.h4 * a.w4;
        float mx = row[5 + k];
        for (int yy = y0; yy < y1; ++yy)
          for (int xx = x0; xx < x1; ++xx) { const float v = cm[(long long)yy * a.w4 + xx]; mx = v > mx ? v : mx; }
        row[5 + k] = mx;
      }
    }
#pragma unroll
    for (int c = 0; c < 9; ++c) a.out_loc[(long long)j * 9 + c] = row[c];
  }
  for (long long e = tid; e < (long long)kept * a.fc; e += SB_THREADS) {
    const int j = (int)(e / a.fc), c = (int)(e % a.fc);
    a.out_gf[e] = a.gfeat[(long long)a.sel_idx[j] * a.fc + c];
  }
}

}  // namespace
}  // namespace ftc

using namespace ftc;

extern "C" {

int ftc_box_hists(const unsigned char* page, int page_h, int page_w, const float* loc, int n, double* hists, void* stream) {
  FTC_REQUIRE(page && loc && hists && page_h > 0 && page_w > 0 && n >= 0, "bad argument");
  if (n == 0) return 0;
  FTC_REQUIRE(n <= 65535 * 32, "too many boxes");
  box_hist_kernel<<<dim3((unsigned)n, 2), 128, 0, (cudaStream_t)stream>>>(page, page_h, page_w, loc, n, hists);
  FTC_POST_LAUNCH();
  return 0;
}

size_t ftc_select_boxes_scratch_bytes(int n) { return (size_t)(n > 0 ? n : 1) * (4 * sizeof(double) + 4 * sizeof(int) + sizeof(int)) + 64; }

int ftc_select_boxes(const float* loc, const float* gfeat, int feat_ch, const int* order, int n, const double* tight, double th,
                     const float* seps_all, const float* code_all, int h4, int w4, int scale, int* n_out, int* sel_idx, float* out_loc,
                     float* out_gf, void* scratch, size_t scratch_bytes, void* stream) {
  FTC_REQUIRE(loc && gfeat && order && tight && seps_all && code_all && n_out && sel_idx && out_loc && out_gf && scratch, "null argument");
  FTC_REQUIRE(n >= 0 && feat_ch > 0 && h4 > 0 && w4 > 0 && scale > 0, "bad geometry");
  FTC_REQUIRE(scratch_bytes >= ftc_select_boxes_scratch_bytes(n), "scratch too small");
  SelectArgs a;
  a.loc = loc; a.gfeat = gfeat; a.order = order; a.tight = tight; a.th = th; a.seps = seps_all; a.code = code_all;
  a.n = n; a.fc = feat_ch; a.h4 = h4; a.w4 = w4; a.scale = scale;
  a.n_out = n_out; a.sel_idx = sel_idx; a.out_loc = out_loc; a.out_gf = out_gf;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  a.acc = reinterpret_cast<double*>(scratch);
  a.rects = reinterpret_cast<int*>(reinterpret_cast<char*>(scratch) + nn * 4 * sizeof(double));
  a.keep = a.rects + nn * 4;
  select_boxes_kernel<<<1, SB_THREADS, 0, (cudaStream_t)stream>>>(a);
  FTC_POST_LAUNCH();
  return 0;
}

}  // extern "C"
