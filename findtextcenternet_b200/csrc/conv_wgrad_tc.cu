// Weight gradient of nn.Conv2d (k = 1 | 3, stride 1, pad (k-1)/2) and nn.Linear on the 5th-generation tensor cores.
//
//   dW[co][tap][ci] = sum over output pixels m of  dy[m][co] * x[m + shift(tap)][ci]          (train1.py:151 loss.backward())
//
// As a tcgen05 GEMM the REDUCTION index is the pixel: D[M = 128 co][N = (tap, ci)] += A[co][16 px] * B[(tap, ci)][16 px]^T.
// Both operands are NHWC activations whose channel axis is contiguous, i.e. the M / N index is the fast axis in memory:
// "MN-major" UMMA operands (instruction-descriptor bits 15 / 16).  A TMA tensor box of (64 channels x pixels) lands in shared
// memory as rows of 128 B per pixel with the 128-byte swizzle -- exactly the canonical MN-major SWIZZLE_128B layout
//   ((8, n), (8, k)) : ((1, LBO), (8, SBO))  in 16-byte units  (cute/atom/mma_traits_sm100.hpp "make_umma_desc<Major::MN>")
// with SBO = 1024 B (eight pixels) and LBO = the distance between two 64-channel chunks.  No transposition pass, no im2col.
//
// 3x3: per column shift kx ONE halo box of the input, (rows + 2) x cols x 64 ch with the conv padding supplied by TMA's
// out-of-bounds zero fill, serves the three row taps: tap ky starts ky image rows = ky * cols * 128 B further (cols is a
// multiple of 8, so the swizzle phase is preserved).  With FUSE_KY the three row taps are ONE instruction: N = 192 whose three
// 64-channel "chunks" are LBO = cols * 128 B apart (they overlap in memory; the descriptor is an address generator).
// TMEM holds 128 lanes x 512 fp32 columns, one 64-ci chunk x 9 taps would need 576: a CTA owns the taps of two column shifts
// (tap group 0: kx = 0, 1 -> 384 columns) or of the third (tap group 1: kx = 2 -> 192 columns).
// 1x1 / Linear: N = up to 512 input channels (two 256-column accumulators share each dy stage).
//
// Split-K over pixel tiles fills the machine (M is 16 x 192 x 192 pixels, the output a few MB); every split writes its fp32
// partial tile to scratch and a second kernel adds the splits IN ORDER and scatters into the OIHW gradient: deterministic, no
// atomics.  Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5 epilogue (TMEM lane quarter =
// warp % 4).
#include <algorithm>

#include "../../include/ftc_b200.h"
#include "tc_common.cuh"
#include "tma_util.cuh"

namespace ftc {
// -1: follow FTC_WGRAD_TC (default 2), 0 off, 1 on (three N = 64 row-tap instructions per column shift), 2 on + fused row taps
int g_wgrad_tc = -1;
namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_STAGES_MAX = 4;

struct WgradTcParams {
  int ksize;                  // 1 or 3
  int Cin, Cout, CinP;        // CinP = Cin rounded up to 64 (column pitch of a tap in the partial tile)
  int B, H, W;                // output == input extents (stride 1); for 1x1 / Linear: H = 1, B = 1, W = rows
  int R, Cw;                  // pixel tile: R rows x Cw columns (3x3);  R * Cw = P pixels (1x1: P rows of the flat matrix)
  int P;                      // pixels per tile (multiple of 16)
  int tiles_x, tiles_y;       // pixel tiles per image
  int n_ptiles;               // total pixel tiles = B * tiles_y * tiles_x
  int splits, tiles_per_split;
  int n_mt;                   // co tiles of 128
  int n_items;                // 3x3: ci chunks (64) x 2 tap groups; 1x1: ci blocks of ncb chunks
  int ncb;                    // 1x1: 64-channel chunks per CTA (<= 8)
  int stages;
  int fuse_ky;                // 3x3: one N = 192 instruction per column shift (else three N = 64)
  uint32_t a_bytes, x_bytes, stage_bytes;   // per stage: dy (2 chunks), one x box, everything
  int64_t part_ld;            // floats per co row of a partial tile = taps * CinP
  int64_t part_split;         // floats per split = CoutP * part_ld
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void tma_load_4d_wg(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ WgradTcParams p, const __grid_constant__ CUtensorMap tmDy,
                     const __grid_constant__ CUtensorMap tmX, float* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t wg_smem_raw[];
  const uint32_t raw = smem_u32(wg_smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* smem = wg_smem_raw + (sbase - raw);
  const uint32_t bar0 = sbase + (uint32_t)p.stages * p.stage_bytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (WG_STAGES_MAX + s); };
  const uint32_t acc_bar = bar0 + 8u * (2 * WG_STAGES_MAX);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + (size_t)p.stages * p.stage_bytes + 8 * (2 * WG_STAGES_MAX + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item
  const int mt = blockIdx.x % p.n_mt;
  const int item = blockIdx.x / p.n_mt;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.n_ptiles, t_begin + p.tiles_per_split);
  const int co0 = mt * 128;
  const bool k3 = p.ksize == 3;
  int ci0, nkx, kx0, nchunk;
  if (k3) {
    const int chunk = item >> 1, tg = item & 1;
    ci0 = chunk * 64; kx0 = tg ? 2 : 0; nkx = tg ? 1 : 2; nchunk = 1;
  } else {
    ci0 = item * p.ncb * 64; kx0 = 0; nkx = 1;
    nchunk = min(p.ncb, (p.CinP - ci0) / 64);
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmDy)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmX)) : "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(acc_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  const uint32_t chunk_bytes = (uint32_t)p.P * 128u;     // one 64-channel chunk of a P-pixel tile
  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = sbase + (uint32_t)s * p.stage_bytes;
        mbar_arrive_expect_tx(full_bar(s), p.a_bytes + (k3 ? (uint32_t)nkx * p.x_bytes : (uint32_t)nchunk * chunk_bytes));
        if (k3) {
          const int b = t / (p.tiles_x * p.tiles_y);
          const int r = t - b * (p.tiles_x * p.tiles_y);
          const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
          const int y0 = ty * p.R, x0 = tx * p.Cw;
          tma_load_4d_wg(st, &tmDy, co0, x0, y0, b, full_bar(s));
          tma_load_4d_wg(st + chunk_bytes, &tmDy, co0 + 64, x0, y0, b, full_bar(s));
          for (int j = 0; j < nkx; ++j)
            tma_load_4d_wg(st + p.a_bytes + (uint32_t)j * p.x_bytes, &tmX, ci0, x0 + (kx0 + j) - 1, y0 - 1, b, full_bar(s));
        } else {
          const int m0 = t * p.P;
          tma_load_4d_wg(st, &tmDy, co0, m0, 0, 0, full_bar(s));
          tma_load_4d_wg(st + chunk_bytes, &tmDy, co0 + 64, m0, 0, 0, full_bar(s));
          for (int j = 0; j < nchunk; ++j)
            tma_load_4d_wg(st + p.a_bytes + (uint32_t)j * chunk_bytes, &tmX, ci0 + j * 64, m0, 0, 0, full_bar(s));
        }
        s = (s + 1 == p.stages) ? 0 : s + 1;
        ph ^= (s == 0) ? 1u : 0u;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    const bool leader = elect_one_sync();
    const uint32_t idesc_base = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 4) << 24);
    const int ksteps = p.P / 16;
    int s = 0;
    uint32_t ph = 0;
    uint32_t accum = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t st = sbase + (uint32_t)s * p.stage_bytes;
      if (leader) {
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint32_t koff = (uint32_t)kk * 2048u;                    // 16 pixels x 128 B
          const uint64_t adesc = umma_desc_mn_sw128(st + koff, chunk_bytes);
          if (k3) {
            const uint32_t row_b = (uint32_t)p.Cw * 128u;                 // one image row of the halo box
            for (int j = 0; j < nkx; ++j) {
              const uint32_t xb = st + p.a_bytes + (uint32_t)j * p.x_bytes + koff;
              if (p.fuse_ky) {
                const uint32_t idesc = idesc_base | ((uint32_t)(192 >> 3) << 17);
                umma_f16(tmem_base + (uint32_t)(j * 192), adesc, umma_desc_mn_sw128(xb, row_b), idesc, accum);
              } else {
                const uint32_t idesc = idesc_base | ((uint32_t)(64 >> 3) << 17);
                for (int ky = 0; ky < 3; ++ky)
                  umma_f16(tmem_base + (uint32_t)(j * 192 + ky * 64), adesc, umma_desc_mn_sw128(xb + (uint32_t)ky * row_b, chunk_bytes),
                           idesc, accum);
              }
            }
          } else {
            // N = 64 * nchunk columns, at most 256 per instruction
            for (int c = 0; c < nchunk; c += 4) {
              const int nc = min(4, nchunk - c);
              const uint32_t idesc = idesc_base | ((uint32_t)((64 * nc) >> 3) << 17);
              umma_f16(tmem_base + (uint32_t)(c * 64), adesc, umma_desc_mn_sw128(st + p.a_bytes + (uint32_t)c * chunk_bytes + koff, chunk_bytes),
                       idesc, accum);
            }
          }
          accum = 1u;
        }
        umma_commit(empty_bar(s));                 // the stage may be refilled when these MMAs have read it
      }
      __syncwarp();
      s = (s + 1 == p.stages) ? 0 : s + 1;
      ph ^= (s == 0) ? 1u : 0u;
    }
    if (leader) umma_commit(acc_bar);              // accumulators complete
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    mbar_wait(acc_bar, 0, 64);
    tc_fence_after();
    const int co = co0 + q * 32 + lane;
    const int ncols = k3 ? nkx * 192 : nchunk * 64;
    float* prow = part + (int64_t)split * p.part_split + (int64_t)co * p.part_ld;
    const bool row_ok = co < p.Cout;
    const bool empty = t_end <= t_begin;            // a split without pixel tiles contributes zeros (the plan never makes one)
    for (int c0 = 0; c0 < ncols; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      if (!row_ok) continue;
      if (empty) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0u;
      }
      // column -> (tap, ci): 3x3 accumulators are [kx - kx0][ky][64 ci]; the partial tile is [tap = ky * 3 + kx][CinP]
      int64_t col;
      if (k3) {
        const int j = c0 / 192, rem = c0 - j * 192, ky = rem >> 6, cc = rem & 63;
        col = (int64_t)(ky * 3 + kx0 + j) * p.CinP + ci0 + cc;
      } else {
        col = ci0 + c0;
      }
      float4* dst = reinterpret_cast<float4*>(prow + col);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                             __uint_as_float(v[4 * i + 3]));
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// dW[co][ci][tap] (fp32 OIHW) = sum over splits, in order, of part[split][co][tap * CinP + ci]
__global__ void __launch_bounds__(256) conv_wgrad_tc_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, int Cout,
                                                                   int Cin, int CinP, int taps, int splits, int64_t part_ld,
                                                                   int64_t part_split) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)Cout * Cin) return;
  const int co = (int)(i / Cin), ci = (int)(i - (int64_t)co * Cin);
  const float* src = part + (int64_t)co * part_ld + ci;
  float* dst = dw + i * taps;
  for (int t = 0; t < taps; ++t) {
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += src[(int64_t)s * part_split + (int64_t)t * CinP];
    dst[t] = a;
  }
}

 
int wgrad_tc_mode() {
  if (g_wgrad_tc >= 0) return g_wgrad_tc;
  static const int env = [] { const char* e = getenv("FTC_WGRAD_TC"); return e ? atoi(e) : 2; }();   // default: on, row taps fused (20 / 20 hardware parity cases green in round 2)
  return env;
}

bool plan(WgradTcParams& p, int batch, int h, int w, int cin, int cout, int ksize, int stride) {
  const int mode = wgrad_tc_mode();
  if (mode <= 0 || stride != 1 || (ksize != 1 && ksize != 3) || cin % 8 || cout % 8 || cin < 16 || cout < 16) return false;
  memset(&p, 0, sizeof(p));
  p.ksize = ksize; p.Cin = cin; p.Cout = cout; p.CinP = (cin + 63) / 64 * 64;
  p.n_mt = (cout + 127) / 128;
  p.fuse_ky = mode >= 2;
  const int64_t M = (int64_t)batch * h * w;
  if (ksize == 3) {
    if (w < 8 || h < 4) return false;
    p.B = batch; p.H = h; p.W = w;
    p.Cw = (w % 16 == 0 || w > 24) ? 16 : 8;
    p.R = 8;
    p.P = p.R * p.Cw;
    p.tiles_x = (w + p.Cw - 1) / p.Cw; p.tiles_y = (h + p.R - 1) / p.R;
    p.n_ptiles = batch * p.tiles_x * p.tiles_y;
    p.n_items = (p.CinP / 64) * 2;
    p.ncb = 1;
    p.a_bytes = 2u * (uint32_t)p.P * 128u;
    p.x_bytes = (uint32_t)(p.R + 2) * p.Cw * 128u;
    p.stage_bytes = p.a_bytes + 2u * p.x_bytes;
  } else {
    if (M < 64) return false;
    p.B = 1; p.H = 1; p.W = (int)M;
    if (M > 0x7fffffff) return false;
    p.P = 128; p.R = 1; p.Cw = 128;
    p.tiles_x = (int)((M + p.P - 1) / p.P); p.tiles_y = 1;
    p.n_ptiles = p.tiles_x;
    const int nch = p.CinP / 64;
    p.ncb = std::min(nch, 4);                     // 256 input channels per CTA: 96 KB stages, two of them
    p.n_items = (nch + p.ncb - 1) / p.ncb;
    p.a_bytes = 2u * (uint32_t)p.P * 128u;
    p.x_bytes = (uint32_t)p.P * 128u;
    p.stage_bytes = p.a_bytes + (uint32_t)p.ncb * p.x_bytes;
  }
  p.stages = std::min<int>(WG_STAGES_MAX, (int)((224u * 1024u) / p.stage_bytes));
  if (p.stages < 2) return false;
  // split-K over pixel tiles: about two waves of CTAs, at least 4 pixel tiles per CTA
  const int work = p.n_mt * p.n_items;
  int splits = std::max(1, std::min((2 * 148 + work - 1) / work, p.n_ptiles / 4));
  splits = std::min(splits, 512);
  p.tiles_per_split = (p.n_ptiles + splits - 1) / splits;
  p.splits = (p.n_ptiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.part_ld = (int64_t)ksize * ksize * p.CinP;
  p.part_split = (int64_t)p.n_mt * 128 * p.part_ld;
  return true;
}

}  // namespace

size_t conv_wgrad_tc_scratch_bytes(int batch, int h, int w, int cin, int cout, int ksize, int stride) {
  WgradTcParams p;
  if (!plan(p, batch, h, w, cin, cout, ksize, stride)) return 0;
  return (size_t)p.splits * p.part_split * sizeof(float);
}

// returns 1 when the shape is not handled here (caller falls back), 0 on success, < 0 on error
int conv_wgrad_tc(const void* x, const void* dy, int batch, int h, int w, int cin, int cout, int ksize, int stride, float* dw_oihw,
                  void* scratch, size_t scratch_bytes, cudaStream_t s) {
  WgradTcParams p;
  if (!plan(p, batch, h, w, cin, cout, ksize, stride)) return 1;
  if (scratch == nullptr || scratch_bytes < (size_t)p.splits * p.part_split * sizeof(float)) return 1;
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(scratch)) & 15) != 0) return 1;
  alignas(64) CUtensorMap tmDy, tmX;
  int rc;
  if (ksize == 3) {
    rc = tma_encode_nhwc(&tmDy, dy, DT_BF16, cout, cout, w, h, batch, 64, p.Cw, p.R, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = tma_encode_nhwc(&tmX, x, DT_BF16, cin, cin, w, h, batch, 64, p.Cw, p.R + 2, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    rc = tma_encode_nhwc(&tmDy, dy, DT_BF16, cout, cout, p.W, 1, 1, 64, p.P, 1, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = tma_encode_nhwc(&tmX, x, DT_BF16, cin, cin, p.W, 1, 1, 64, p.P, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 8 * (2 * WG_STAGES_MAX + 1) + 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    FTC_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_done = true;
  }
  FTC_REQUIRE(smem <= 227 * 1024, "wgrad tile does not fit shared memory");
  dim3 grid((unsigned)(p.n_mt * p.n_items), (unsigned)p.splits);
  conv_wgrad_tc_kernel<<<grid, WG_THREADS, smem, s>>>(p, tmDy, tmX, reinterpret_cast<float*>(scratch));
  FTC_POST_LAUNCH();
  const int64_t n = (int64_t)cout * cin;
  conv_wgrad_tc_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float*>(scratch), dw_oihw, cout, cin,
                                                                          p.CinP, ksize * ksize, p.splits, p.part_ld, p.part_split);
  FTC_POST_LAUNCH();
  return 0;
}

}  // namespace ftc

using namespace ftc;

extern "C" {

size_t ftc_train_conv2d_wgrad_scratch_bytes(int batch, int h, int w, int cin, int cout, int ksize, int stride) {
  return conv_wgrad_tc_scratch_bytes(batch, h, w, cin, cout, ksize, stride);
}

int ftc_train_conv2d_wgrad_ws(const void* x, const void* dy, int dtype, int batch, int h, int w, int cin, int cout, int ksize,
                              int stride, float* dw_oihw, void* scratch, size_t scratch_bytes, void* stream) {
  FTC_REQUIRE(x && dy && dw_oihw, "null argument");
  if (dtype == DT_BF16 && batch > 0 && h > 0 && w > 0) {
    const int rc = conv_wgrad_tc(x, dy, batch, h, w, cin, cout, ksize, stride, dw_oihw, scratch, scratch_bytes, (cudaStream_t)stream);
    if (rc <= 0) return rc;          // done (0) or failed (< 0); 1 = shape not handled by the tcgen05 kernel
  }
  return ftc_train_conv2d_wgrad(x, dy, dtype, batch, h, w, cin, cout, ksize, stride, dw_oihw, stream);
}

int ftc_debug_set_wgrad_tc(int mode) {
  g_wgrad_tc = mode;
  return 0;
}

}  // extern "C"
