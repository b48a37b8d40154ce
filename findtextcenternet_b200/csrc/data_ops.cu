// train1 input pipeline on the device: the reference's per-sample Cython routine
// dataset/processer.pyx::transform_crop (:260-454) + process() (:655-673) + the colour compositing functions (:675-887) +
// random_salt (dataset/data_detector.py:17-26), for a whole batch in four launches.
//
//   crop_init_kernel     centre map = 0, log-size maps = +inf, id maps = 0
//   crop_prepare_kernel  one CTA per sample: box corners through the page's affine matrix (:345-355), the crop origin from the
//                        anchor box (:358-366), per-box crop coordinates + in-crop flag, minsize in the reference's sequential
//                        order (:371-385)
//   crop_label_kernel    one CTA per box: separable Gaussian (center_map :137-163) merged with atomicMax on the float bit patterns
//                        (values in (0, 1]), ellipse footprint (box_map :165-186, id_map :188-206) merged with atomicMin / atomicMax:
//                        maximum and minimum are order-independent, so the result is deterministic
//   crop_image_kernel    thread per output pixel: inverse affine, inverse_partial folded into the pixel fetch, nearest or bilinear
//                        gather from the uint8 page (L2-resident), salt cells, colour compositing; coalesced plane writes
//   crop_maps_kernel     thread per map pixel: bilinear textline / separator crop, +inf -> 0 on the log-size maps
//
// Bound: HBM writes (7.1 MB of fp32 image + 1.0 MB of maps per sample; the page bytes a crop touches are ~0.6 MB).
// Arithmetic follows the generated C of the Cython source operation by operation: float32 with round-to-nearest intrinsics (no
// FMA contraction), and double where the source promotes ("1 - dx" is emitted as 1.0 - dx, "rx + 0.5", the colour blend).
// expf / logf are evaluated in double and rounded (<= 1 ulp from any libm).
// Plain SIMT on purpose: the CPU kernel-emulation build compiles this file for host threads (tests/test_processer.py).
#include "../../include/ftc_b200.h"
#include "common.cuh"

#ifdef FTC_EMU
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  std::lock_guard<std::mutex> g(emu::atomic_lock);
  unsigned old = *p; if (v > old) *p = v; return old;
}
static inline int atomicMin(int* p, int v) {
  std::lock_guard<std::mutex> g(emu::atomic_lock);
  int old = *p; if (v < old) *p = v; return old;
}
#endif

namespace ftc {
namespace {

constexpr int CW = 768, CH = 768, CS = 4, MW = CW / CS, MH = CH / CS;     // util_func.py:6-8

struct BoxRec { float cx, cy, w, h; int flag, sample, code1, code2; };

// vector_dot (:75-86): v = 0; v += a[k] * b[k] for b = (x, y, 1), float32, one rounding per operation
__device__ __forceinline__ void vdot(const float* a, float x, float y, float* rx, float* ry) {
  *rx = __fadd_rn(__fadd_rn(__fmul_rn(a[0], x), __fmul_rn(a[1], y)), a[2]);
  *ry = __fadd_rn(__fadd_rn(__fmul_rn(a[3], x), __fmul_rn(a[4], y)), a[5]);
}

__global__ void crop_init_kernel(float* __restrict__ out_map, int* __restrict__ out_idmap, int batch) {
  const long long per = (long long)MH * MW;
  const long long total = (long long)batch * 5 * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)((i / per) % 5);
    if (ch == 0) out_map[i] = 0.f;
    else if (ch <= 2) out_map[i] = INFINITY;
    else {                                   // reuse the index space of channels 3, 4 for the two id-map planes
      const long long b = i / (5 * per), r = i % per;
      out_idmap[(b * 2 + (ch - 3)) * per + r] = 0;
    }
  }
}

__global__ void crop_prepare_kernel(const ftc_crop_sample* __restrict__ samples, const float* __restrict__ position,
                                    const int* __restrict__ codelist, BoxRec* __restrict__ rec, float* __restrict__ start,
                                    float* __restrict__ out_minsize) {
  const int b = blockIdx.x;
  const ftc_crop_sample& s = samples[b];
  __shared__ float sh_start[2];
  const int n = s.blank ? 0 : s.box_count;
  // rotated boxes (cx, cy, w, h) in page coordinates
  for (int i = threadIdx.x; i < s.box_count; i += blockDim.x) {
    const float* p = position + (size_t)(s.box_begin + i) * 4;
    const float hw = __fdiv_rn(p[2], 2.f), hh = __fdiv_rn(p[3], 2.f);
    float xr1, yr1, xr2, yr2;
    vdot(s.rot, __fsub_rn(p[0], hw), __fsub_rn(p[1], hh), &xr1, &yr1);
    vdot(s.rot, __fadd_rn(p[0], hw), __fadd_rn(p[1], hh), &xr2, &yr2);
    BoxRec r;
    r.cx = __fdiv_rn(__fadd_rn(xr1, xr2), 2.f);
    r.cy = __fdiv_rn(__fadd_rn(yr1, yr2), 2.f);
    r.w = __fsub_rn(xr2, xr1);
    r.h = __fsub_rn(yr2, yr1);
    r.flag = 0; r.sample = b;
    r.code1 = codelist[(size_t)(s.box_begin + i) * 2];
    r.code2 = codelist[(size_t)(s.box_begin + i) * 2 + 1];
    rec[s.box_begin + i] = r;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sx = s.startx0, sy = s.starty0;
    if (s.box_count > 0) {
      const BoxRec& a = rec[s.box_begin + s.cidx];
      sx = __fsub_rn(a.cx, s.woffset);
      sy = __fsub_rn(a.cy, s.hoffset);
    }
    sh_start[0] = sx; sh_start[1] = sy;
    start[b * 2] = sx; start[b * 2 + 1] = sy;
  }
  __syncthreads();
  const float sx = sh_start[0], sy = sh_start[1];
  for (int i = threadIdx.x; i < s.box_count; i += blockDim.x) {
    BoxRec& r = rec[s.box_begin + i];
    r.cx = __fsub_rn(r.cx, sx);
    r.cy = __fsub_rn(r.cy, sy);
    r.flag = (i < n && r.cx > 0.f && r.cx < (float)CW && r.cy > 0.f && r.cy < (float)CH) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {            // the reference's running minimum restarts whenever it is <= 0 (:382-385)
    float m = 0.f;
    for (int i = 0; i < n; ++i) {
      const BoxRec& r = rec[s.box_begin + i];
      if (!r.flag) continue;
      const float v = r.h > r.w ? r.h : r.w;
      if (m <= 0.f) m = v; else m = v < m ? v : m;
    }
    out_minsize[b] = m;
  }
}

// gkern (:40-47): ax = i - (l - 1) / 2.0 (double, stored float); exp argument ((-0.5 * ax) * ax) / (double)(sig * sig) -> float
__device__ __forceinline__ float gauss_tap(int i, int l, float sig) {
  const float ax = (float)__dsub_rn((double)(float)i, __ddiv_rn((double)(float)(l - 1), 2.0));
  const double arg = __ddiv_rn(__dmul_rn(__dmul_rn(-0.5, (double)ax), (double)ax), (double)__fmul_rn(sig, sig));
  return (float)exp((double)(float)arg);
}

__global__ void crop_label_kernel(const BoxRec* __restrict__ rec, float* __restrict__ out_map, int* __restrict__ out_idmap) {
  const BoxRec r = rec[blockIdx.x];
  if (!r.flag) return;
  __shared__ float gx[MW], gy[MH];
  const long long per = (long long)MH * MW;
  float* center = out_map + (size_t)r.sample * 5 * per;
  float* box0 = center + per;
  float* box1 = center + 2 * per;
  int* id0 = out_idmap + (size_t)r.sample * 2 * per;
  int* id1 = id0 + per;
  // ---- center_map (:137-163) ----
  {
    const float cx = __fdiv_rn(r.cx, (float)CS), cy = __fdiv_rn(r.cy, (float)CS);
    const float w = __fdiv_rn(r.w, (float)CS), h = __fdiv_rn(r.h, (float)CS);
    const float w2 = __fdiv_rn(w, 2.f), h2 = __fdiv_rn(h, 2.f);
    const float fix_w = 1.f > w2 ? 1.f : w2, fix_h = 1.f > h2 ? 1.f : h2;
    const double kw = __dmul_rn((double)fix_w, 1.5), kh = __dmul_rn((double)fix_h, 1.5);
    const int ks = (int)(kh > kw ? kh : kw);
    const float std_x = __fdiv_rn(fix_w, 4.f), std_y = __fdiv_rn(fix_h, 4.f);
    const int L = ks * 2 + 1;
    const int xi = (int)roundf(cx), yi = (int)roundf(cy);
    const int x0 = xi - ks, y0 = yi - ks;
    const int xa = x0 > 0 ? x0 : 0, xb = (x0 + L) < MW ? (x0 + L) : MW;
    const int ya = y0 > 0 ? y0 : 0, yb = (y0 + L) < MH ? (y0 + L) : MH;
    for (int x = xa + (int)threadIdx.x; x < xb; x += blockDim.x) gx[x] = gauss_tap(x - x0, L, std_x);
    for (int y = ya + (int)threadIdx.x; y < yb; y += blockDim.x) gy[y] = gauss_tap(y - y0, L, std_y);
    __syncthreads();
    const int ww = xb - xa, wh = yb - ya;
    if (ww > 0 && wh > 0)
      for (int i = threadIdx.x; i < ww * wh; i += blockDim.x) {
        const int x = xa + i % ww, y = ya + i / ww;
        const float v = __fmul_rn(gy[y], gx[x]);
        atomicMax(reinterpret_cast<unsigned*>(center + (size_t)y * MW + x), __float_as_uint(v));   // v >= 0
      }
  }
  // ---- box_map / id_map footprint (:165-206) ----
  {
    const float w10 = __fdiv_rn(r.w, 10.f), h10 = __fdiv_rn(r.h, 10.f);
    const float fix_w = (float)CS > w10 ? (float)CS : w10, fix_h = (float)CS > h10 ? (float)CS : h10;
    const float sizex = (float)__dadd_rn((double)(float)log((double)__fdiv_rn(r.w, 1024.f)), 3.0);
    const float sizey = (float)__dadd_rn((double)(float)log((double)__fdiv_rn(r.h, 1024.f)), 3.0);
    int xmin = (int)__fdiv_rn(__fsub_rn(r.cx, fix_w), (float)CS) - 2; xmin = xmin > 0 ? xmin : 0;
    int xmax = (int)__fdiv_rn(__fadd_rn(r.cx, fix_w), (float)CS) + 2; xmax = xmax < MW ? xmax : MW;
    int ymin = (int)__fdiv_rn(__fsub_rn(r.cy, fix_h), (float)CS) - 2; ymin = ymin > 0 ? ymin : 0;
    int ymax = (int)__fdiv_rn(__fadd_rn(r.cy, fix_h), (float)CS) + 2; ymax = ymax < MH ? ymax : MH;
    const int ww = xmax - xmin, wh = ymax - ymin;
    if (ww > 0 && wh > 0)
      for (int i = threadIdx.x; i < ww * wh; i += blockDim.x) {
        const int xi = xmin + i % ww, yi = ymin + i / ww;
        const float x = __fsub_rn((float)(xi * CS), r.cx), y = __fsub_rn((float)(yi * CS), r.cy);
        const float qx = __fdiv_rn(x, fix_w), qy = __fdiv_rn(y, fix_h);
        if (__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)) < 1.f) {
          const size_t idx = (size_t)yi * MW + xi;
          // float minimum through integer atomics: non-negative floats order like signed ints, negative ones like reversed unsigned
          if (sizex == sizex) {
            if (sizex >= 0.f) atomicMin(reinterpret_cast<int*>(box0 + idx), __float_as_int(sizex));
            else atomicMax(reinterpret_cast<unsigned*>(box0 + idx), __float_as_uint(sizex));
          }
          if (sizey == sizey) {
            if (sizey >= 0.f) atomicMin(reinterpret_cast<int*>(box1 + idx), __float_as_int(sizey));
            else atomicMax(reinterpret_cast<unsigned*>(box1 + idx), __float_as_uint(sizey));
          }
          atomicMax(id0 + idx, r.code1);
          atomicMax(id1 + idx, r.code2);
        }
      }
  }
}

// getpixel (:208-211) with inverse_partial (:124-135) folded in.  lut[v] = v / 255 rounded to nearest (the IEEE division the
// reference performs per fetch; as an instruction sequence it was ~60 of the kernel's ~290 instructions per pixel)
struct PageView { const unsigned char* image; int im_h, im_w, inv_i0, inv_i1, inv_j0, inv_j1; };
__device__ __forceinline__ float page_pixel(const PageView& s, const float* __restrict__ lut, int x, int y) {
  if ((unsigned)x >= (unsigned)s.im_w || (unsigned)y >= (unsigned)s.im_h) return 0.f;
  int v = s.image[(size_t)y * s.im_w + x];
  if (y >= s.inv_i0 && y < s.inv_i1 && x >= s.inv_j0 && x < s.inv_j1) v = 255 - v;
  return lut[v];
}
__device__ __forceinline__ float mask_pixel(const unsigned char* img, int im_h, int im_w, int x, int y) {
  if (x < 0 || x >= im_w || y < 0 || y >= im_h) return 0.f;
  return __fdiv_rn((float)img[(size_t)y * im_w + x], 255.f);
}

// bilinear weights (:396-401): w11 = (1.0 - dx) * (1.0 - dy), w21 = dx * (1.0 - dy), w12 = (1.0 - dx) * dy in double, w22 = dx * dy in float
__device__ __forceinline__ void bilinear_weights(float rx, float ry, float* w11, float* w21, float* w12, float* w22) {
  const float dx = __fsub_rn(rx, floorf(rx)), dy = __fsub_rn(ry, floorf(ry));
  const double ex = __dsub_rn(1.0, (double)dx), ey = __dsub_rn(1.0, (double)dy);
  *w11 = (float)__dmul_rn(ex, ey);
  *w21 = (float)__dmul_rn((double)dx, ey);
  *w12 = (float)__dmul_rn(ex, (double)dy);
  *w22 = __fmul_rn(dx, dy);
}

// CTA = 256 consecutive pixels of IMG_ROWS consecutive rows of one sample: the sample's parameters are read once per thread
// (registers) and amortised over the rows; every store is a fully coalesced 1 KB run of one output plane.
constexpr int IMG_ROWS = 8;
__global__ void __launch_bounds__(256) crop_image_kernel(const ftc_crop_sample* __restrict__ samples, const float* __restrict__ start,
                                                         float* __restrict__ out_image, int out_channels, int rows_per_cta) {
  __shared__ float lut[256];
  for (int v = threadIdx.x; v < 256; v += blockDim.x) lut[v] = __fdiv_rn((float)v, 255.f);
  __syncthreads();
  const long long per = (long long)CH * CW;
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= CW) return;
  const ftc_crop_sample& s = samples[b];
  const int blank = s.blank, nearest = s.nearest, mode = s.color_mode;
  PageView pv;
  pv.image = s.image; pv.im_h = s.im_h; pv.im_w = s.im_w;
  pv.inv_i0 = s.inv_i; pv.inv_i1 = s.inv_i + s.inv_h; pv.inv_j0 = s.inv_j; pv.inv_j1 = s.inv_j + s.inv_w;
  float inv[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) inv[i] = s.inv[i];
  const float sx = start[b * 2], sy = start[b * 2 + 1];
  const unsigned char* salt = s.salt;
  const int salt_s = salt ? s.salt_s : 1, salt_w = s.salt_w;
  float fg1[3], fg2[3], bg[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { fg1[c] = s.fg1[c]; fg2[c] = s.fg2[c]; bg[c] = s.bg[c]; }
  const int rt = s.rect_top, rb = s.rect_bottom, rl = s.rect_left, rr = s.rect_right;
  const unsigned char* bgimg = s.bgimg;
  const int bg_h = s.bg_h, bg_w = s.bg_w, bg_sx = s.bg_startx, bg_sy = s.bg_starty;
  const float fx = __fadd_rn((float)x, sx);
  const int y0 = blockIdx.y * rows_per_cta;
  for (int y = y0; y < y0 + rows_per_cta && y < CH; ++y) {
    float a = 0.f;
    if (!blank) {
      float rx, ry;
      vdot(inv, fx, __fadd_rn((float)y, sy), &rx, &ry);
      if (nearest) {
        a = page_pixel(pv, lut, (int)__dadd_rn((double)rx, 0.5), (int)__dadd_rn((double)ry, 0.5));
      } else {
        float w11, w21, w12, w22;
        bilinear_weights(rx, ry, &w11, &w21, &w12, &w22);
        const int ix = (int)rx, iy = (int)ry;
        a = __fmul_rn(w11, page_pixel(pv, lut, ix, iy));
        a = __fadd_rn(a, __fmul_rn(w21, page_pixel(pv, lut, ix + 1, iy)));
        a = __fadd_rn(a, __fmul_rn(w12, page_pixel(pv, lut, ix, iy + 1)));
        a = __fadd_rn(a, __fmul_rn(w22, page_pixel(pv, lut, ix + 1, iy + 1)));
      }
    }
    if (salt != nullptr) {             // random_salt (data_detector.py:17-26): x * noise, NaN cells -> 1
      const int c = salt[(size_t)(y / salt_s) * salt_w + x / salt_s];
      a = c == 0 ? 0.f : (c == 2 ? 1.f : a);
    }
    const long long rem = (long long)y * CW + x;
    if (out_channels == 1) { out_image[(size_t)b * per + rem] = a; continue; }
    float* o = out_image + (size_t)b * 3 * per + rem;
    const double na = __dsub_rn(1.0, (double)a);
    if (mode == 2) {                   // random_background (:690-741)
      const int yi = y + bg_sy, xi = x + bg_sx;
      const bool in = yi >= 0 && yi < bg_h && xi >= 0 && xi < bg_w;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float bgv = in ? lut[bgimg[((size_t)yi * bg_w + xi) * 3 + c]] : 0.f;
        double v = __dadd_rn((double)__fmul_rn(a, fg1[c]), __dmul_rn(na, (double)bgv));
        v = v < 1.0 ? v : 1.0;         // max(0, min(1, v)) as the generated comparisons evaluate it
        v = v > 0.0 ? v : 0.0;
        o[(size_t)c * per] = (float)v;
      }
    } else {                           // random_mono / random_single / random_double (:745-887)
      const bool inner = x > rl && x < rr && y > rt && y < rb;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float fg = inner ? fg2[c] : fg1[c];
        o[(size_t)c * per] = (float)__dadd_rn((double)__fmul_rn(a, fg), __dmul_rn(na, (double)bg[c]));
      }
    }
  }
}

__global__ void crop_maps_kernel(const ftc_crop_sample* __restrict__ samples, const float* __restrict__ start,
                                 float* __restrict__ out_map, int batch) {
  const long long per = (long long)MH * MW;
  const long long total = (long long)batch * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const int rem = (int)(i % per);
    const int y = rem / MW, x = rem % MW;
    const ftc_crop_sample& s = samples[b];
    float* m = out_map + (size_t)b * 5 * per + rem;
    if (s.blank) {
      m[0] = 0.f; m[per] = 0.f; m[2 * per] = 0.f; m[3 * per] = 0.f; m[4 * per] = 0.f;
      continue;
    }
    // +inf (no box) -> 0 (:441-442)
    const float b0 = m[per], b1 = m[2 * per];
    m[per] = (b0 - b0 == 0.f) ? b0 : 0.f;
    m[2 * per] = (b1 - b1 == 0.f) ? b1 : 0.f;
    float rx, ry;
    const float px = __fadd_rn(__fmul_rn((float)x, (float)(CS / 2)), __fdiv_rn(start[b * 2], 2.f));
    const float py = __fadd_rn(__fmul_rn((float)y, (float)(CS / 2)), __fdiv_rn(start[b * 2 + 1], 2.f));
    vdot(s.inv2, px, py, &rx, &ry);
    float w11, w21, w12, w22;
    bilinear_weights(rx, ry, &w11, &w21, &w12, &w22);
    const int ix = (int)rx, iy = (int)ry;
    float t = __fmul_rn(w11, mask_pixel(s.textline, s.im_h2, s.im_w2, ix, iy));
    t = __fadd_rn(t, __fmul_rn(w21, mask_pixel(s.textline, s.im_h2, s.im_w2, ix + 1, iy)));
    t = __fadd_rn(t, __fmul_rn(w12, mask_pixel(s.textline, s.im_h2, s.im_w2, ix, iy + 1)));
    t = __fadd_rn(t, __fmul_rn(w22, mask_pixel(s.textline, s.im_h2, s.im_w2, ix + 1, iy + 1)));
    float p = __fmul_rn(w11, mask_pixel(s.sepline, s.im_h2, s.im_w2, ix, iy));
    p = __fadd_rn(p, __fmul_rn(w21, mask_pixel(s.sepline, s.im_h2, s.im_w2, ix + 1, iy)));
    p = __fadd_rn(p, __fmul_rn(w12, mask_pixel(s.sepline, s.im_h2, s.im_w2, ix, iy + 1)));
    p = __fadd_rn(p, __fmul_rn(w22, mask_pixel(s.sepline, s.im_h2, s.im_w2, ix + 1, iy + 1)));
    m[3 * per] = t;
    m[4 * per] = p;
  }
}


// ---------------------------------------------------------------------------------------------------------------------------
// random_distortion (dataset/data_detector.py:28-42)
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

constexpr long long IMG3 = 3LL * CH * CW;

// im = float(double(im) + alpha * z), clipped to [0, 1]; z: given table or Philox + Box-Muller (two normals per counter pair)
__global__ void distort_noise_kernel(float* __restrict__ image, const ftc_distort_sample* __restrict__ samples,
                                     const double* __restrict__ noise, int batch) {
  const long long total = (long long)batch * IMG3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / IMG3);
    const ftc_distort_sample& s = samples[b];
    if (!s.noise_on) continue;
    double z;
    if (noise != nullptr) {
      z = noise[i];
    } else {
      const long long e = i - (long long)b * IMG3;
      unsigned r[4];
      philox4x32_10((unsigned)(e >> 1), (unsigned)((e >> 1) >> 32), (unsigned)b, 0u, (unsigned)s.noise_seed, (unsigned)(s.noise_seed >> 32), r);
      const float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      const float rad = sqrtf(-2.0f * logf(u1));
      float sn, cs;
      sincosf(6.283185307179586f * u2, &sn, &cs);
      z = (double)(rad * ((e & 1) ? sn : cs));
    }
    float v = (float)__dadd_rn((double)image[i], __dmul_rn(s.alpha, z));
    v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    image[i] = v;
  }
}

// one axis of scipy.ndimage.gaussian_filter (correlate1d, symmetric kernel, mode 'reflect'): in -> out, fp32 storage, double sum
// tmp = in[0] * w[0]; for j = R .. 1: tmp += (in[-j] + in[+j]) * w[j]
__global__ void distort_gauss_axis_kernel(const float* __restrict__ in, float* __restrict__ out, const ftc_distort_sample* __restrict__ samples,
                                          const double* __restrict__ weights, int axis, int batch) {
  const long long total = (long long)batch * IMG3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / IMG3);
    const ftc_distort_sample& s = samples[b];
    if (s.mode == 0) continue;
    const int e = (int)(i - (long long)b * IMG3);
    const int c = e / (CH * CW), y = (e / CW) % CH, x = e % CW;
    const int n = axis == 0 ? 3 : (axis == 1 ? CH : CW);
    const int pos = axis == 0 ? c : (axis == 1 ? y : x);
    const long long stride = axis == 0 ? (long long)CH * CW : (axis == 1 ? CW : 1);
    const float* line = in + (i - (long long)pos * stride);
    const double* w = weights + (size_t)b * 64;
    double tmp = __dmul_rn((double)line[(long long)pos * stride], w[0]);
    const int R = s.radius;
    if (pos - R >= 0 && pos + R < n) {             // interior: no reflection arithmetic (all but a 2R-wide frame of the image axes)
      const float* lp = line + (long long)pos * stride;
      for (int j = R; j >= 1; --j)
        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)lp[-(long long)j * stride], (double)lp[(long long)j * stride]), w[j]));
    } else {
      for (int j = R; j >= 1; --j) {
        int lo = pos - j, hi = pos + j;
        // half-sample symmetric reflection with period 2n (lines shorter than the radius reflect repeatedly: the 3-element colour axis)
        lo %= 2 * n; if (lo < 0) lo += 2 * n; if (lo >= n) lo = 2 * n - 1 - lo;
        hi %= 2 * n; if (hi >= n) hi = 2 * n - 1 - hi;
        tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn((double)line[(long long)lo * stride], (double)line[(long long)hi * stride]), w[j]));
      }
    }
    out[i] = (float)tmp;
  }
}

// axis 0 (the 3-element colour axis): thread per pixel, the three channels in registers; the reflected index of c +- j follows the
// period-6 pattern 0 1 2 2 1 0, looked up from j mod 6 (one small modulo per tap instead of two 64-bit divisions per tap and channel)
__global__ void distort_gauss_color_kernel(const float* __restrict__ in, float* __restrict__ out, const ftc_distort_sample* __restrict__ samples,
                                           const double* __restrict__ weights, int batch) {
  const long long per = (long long)CH * CW;
  const long long total = (long long)batch * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    const ftc_distort_sample& s = samples[b];
    if (s.mode == 0) continue;
    const long long base = (long long)b * IMG3 + (i - (long long)b * per);
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = (double)in[base + c * per];
    const double* w = weights + (size_t)b * 64;
    double t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = __dmul_rn(v[c], w[0]);
    for (int j = s.radius; j >= 1; --j) {
      const int m = j % 6;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int lo = (c - m + 6) % 6, hi = (c + m) % 6;          // compile-time c: small constant arithmetic
        lo = lo >= 3 ? 5 - lo : lo; hi = hi >= 3 ? 5 - hi : hi;
        const double a = lo == 0 ? v[0] : (lo == 1 ? v[1] : v[2]);
        const double d = hi == 0 ? v[0] : (hi == 1 ? v[1] : v[2]);
        t[c] = __dadd_rn(t[c], __dmul_rn(__dadd_rn(a, d), w[j]));
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[base + c * per] = (float)t[c];
  }
}

// mode 1: im = clip(blur); mode 2: im = clip(im + k * (im - blur)) in float32 (numpy: weak Python scalar, float32 arrays)
__global__ void distort_combine_kernel(float* __restrict__ image, const float* __restrict__ blur, const ftc_distort_sample* __restrict__ samples,
                                       int batch) {
  const long long total = (long long)batch * IMG3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const ftc_distort_sample& s = samples[(int)(i / IMG3)];
    if (s.mode == 0) continue;
    float v = blur[i];
    if (s.mode == 2) {
      const float im = image[i];
      v = __fadd_rn(im, __fmul_rn(s.unsharp_k, __fsub_rn(im, v)));
    }
    v = v < 0.f ? 0.f : (v > 1.f ? 1.f : v);
    image[i] = v;
  }
}

#ifdef FTC_EMU
constexpr int kThreads = 32, kMaxBlocks = 4, kBoxThreads = 32;     // one OS thread per CUDA thread on the host
#else
constexpr int kThreads = 256, kMaxBlocks = 148 * 16, kBoxThreads = 128;
#endif

inline int grid_for(long long n) {
  long long g = (n + kThreads - 1) / kThreads;
  return (int)(g < 1 ? 1 : (g > kMaxBlocks ? kMaxBlocks : g));
}

}  // namespace
}  // namespace ftc

using namespace ftc;

extern "C" int ftc_crop_sample_bytes(void) { return (int)sizeof(ftc_crop_sample); }

extern "C" size_t ftc_crop_scratch_bytes(int batch, int total_boxes) {
  return (size_t)(total_boxes > 0 ? total_boxes : 1) * sizeof(BoxRec) + (size_t)(batch > 0 ? batch : 1) * 2 * sizeof(float) + 256;
}

extern "C" int ftc_crop_batch(const ftc_crop_sample* samples, int batch, const float* position, const int* codelist, int total_boxes,
                              float* out_image, int out_channels, float* out_map, int* out_idmap, float* out_minsize, void* scratch,
                              size_t scratch_bytes, void* stream) {
  FTC_REQUIRE(batch > 0 && total_boxes >= 0, "ftc_crop_batch: batch / boxes");
  FTC_REQUIRE(out_channels == 1 || out_channels == 3, "ftc_crop_batch: out_channels is 1 (gray) or 3 (composited)");
  FTC_REQUIRE(samples && out_image && out_map && out_idmap && out_minsize && scratch, "ftc_crop_batch: null pointer");
  FTC_REQUIRE(total_boxes == 0 || (position && codelist), "ftc_crop_batch: boxes without position / codelist");
  FTC_REQUIRE(scratch_bytes >= ftc_crop_scratch_bytes(batch, total_boxes), "ftc_crop_batch: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  uintptr_t p = ((uintptr_t)scratch + 127) & ~(uintptr_t)127;
  BoxRec* rec = reinterpret_cast<BoxRec*>(p);
  float* start = reinterpret_cast<float*>(p + (size_t)(total_boxes > 0 ? total_boxes : 1) * sizeof(BoxRec));
  FTC_CHECK_CUDA(cudaMemsetAsync(rec, 0, (size_t)(total_boxes > 0 ? total_boxes : 1) * sizeof(BoxRec), s));   // boxes no sample owns: flag 0
  crop_init_kernel<<<grid_for((long long)batch * 5 * MH * MW), kThreads, 0, s>>>(out_map, out_idmap, batch);
  FTC_POST_LAUNCH();
  crop_prepare_kernel<<<batch, kBoxThreads, 0, s>>>(samples, position, codelist, rec, start, out_minsize);
  FTC_POST_LAUNCH();
  if (total_boxes > 0) {
    crop_label_kernel<<<total_boxes, kBoxThreads, 0, s>>>(rec, out_map, out_idmap);
    FTC_POST_LAUNCH();
  }
  {
#ifdef FTC_EMU
    const int tx = 32, rows = 96;      // few, fat CTAs on host threads
#else
    const int tx = 256, rows = IMG_ROWS;
#endif
    FTC_REQUIRE(batch <= 65535, "ftc_crop_batch: batch");
    crop_image_kernel<<<dim3((CW + tx - 1) / tx, (CH + rows - 1) / rows, batch), tx, 0, s>>>(samples, start, out_image, out_channels, rows);
  }
  FTC_POST_LAUNCH();
  crop_maps_kernel<<<grid_for((long long)batch * MH * MW), kThreads, 0, s>>>(samples, start, out_map, batch);
  FTC_POST_LAUNCH();
  return 0;
}

extern "C" size_t ftc_distort_scratch_bytes(int batch) { return (size_t)2 * (size_t)(batch > 0 ? batch : 1) * 3 * CH * CW * sizeof(float) + 256; }

extern "C" int ftc_distort_batch(float* image, int batch, const ftc_distort_sample* samples, const double* weights, const double* noise,
                                 void* scratch, size_t scratch_bytes, void* stream) {
  FTC_REQUIRE(image && samples && weights && scratch && batch > 0, "ftc_distort_batch: bad argument");
  FTC_REQUIRE(scratch_bytes >= ftc_distort_scratch_bytes(batch), "ftc_distort_batch: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  float* t1 = reinterpret_cast<float*>(((uintptr_t)scratch + 127) & ~(uintptr_t)127);
  float* t2 = t1 + (size_t)batch * 3 * CH * CW;
  const int grid = grid_for((long long)batch * IMG3 / 4);
  distort_noise_kernel<<<grid, kThreads, 0, s>>>(image, samples, noise, batch);
  FTC_POST_LAUNCH();
  distort_gauss_color_kernel<<<grid_for((long long)batch * CH * CW / 2), kThreads, 0, s>>>(image, t1, samples, weights, batch);
  FTC_POST_LAUNCH();
  distort_gauss_axis_kernel<<<grid, kThreads, 0, s>>>(t1, t2, samples, weights, 1, batch);
  FTC_POST_LAUNCH();
  distort_gauss_axis_kernel<<<grid, kThreads, 0, s>>>(t2, t1, samples, weights, 2, batch);
  FTC_POST_LAUNCH();
  distort_combine_kernel<<<grid, kThreads, 0, s>>>(image, t1, samples, batch);
  FTC_POST_LAUNCH();
  return 0;
}
