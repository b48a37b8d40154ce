// Scaled-dot-product attention on the 5th-generation tensor cores (tcgen05 + TMEM), sequences of up to 128 tokens:
//   softmax(q k^T / sqrt(hd) + mask) v per (batch, head)          (F.scaled_dot_product_attention, models/transformer.py:133)
//
// Work item = (batch element, 64-channel slab) = two heads of 32 or one head of 64 channels; a TASK is one head of an item.
// Persistent CTAs (one per SM) walk their items through a three-stage shared-memory ring:
//   * Q, K, V of the slab arrive as THREE TMA tensor boxes of (64 channels x 128 rows), 128-byte swizzled: Q and K are K-major
//     UMMA operands as they land (a head of 32 channels is a 64-byte offset inside the swizzled row), V is the MN-major B
//     operand of P.V as it lands (keys are the reduction index: the layout of the weight-gradient kernel, conv_wgrad_tc.cu).
//     Rows beyond the sequence are TMA zero fill (map {channels, L, 1, batch}).
//   * S = Q K^T: one tcgen05.mma series (M = 128 query rows, N = keys rounded up to 16, K = hd) into TMEM.
//   * softmax: a group of 8 warps, TWO THREADS PER QUERY ROW (TMEM lane; each takes half of the keys): tcgen05.ld into
//     registers, scale + key mask, row maximum (halves exchanged through shared memory), exp2 (one MUFU op per element: the
//     unit that bounds this kernel), row sum, and P goes to shared memory as bf16 in the K-major 128-byte-swizzled layout (the
//     A operand of the second product).
//   * O = P V: second tcgen05.mma series (M = 128, N = 64, K = keys) into TMEM; the same threads read their half of the row back,
//     divide by the row sum and store bf16.  For hd = 32 the N = 64 product also forms P_h V_{other head}: those columns are
//     not read.
// Two softmax groups run independent pipelines (hd 32: group g = head g of every item; hd 64: items alternate), each with its
// own TMEM columns, its own P tile and its own MMA-issuing thread, so one group's exponentials run under the other group's
// products.  Inside a group the tasks are software-pipelined: TMEM holds TWO score buffers per group (the 64 output columns
// of a task overwrite the head of its own, already consumed, score buffer), the score product of task n + 1 is issued before
// P.V of task n, and the group reads / scales / maximises the scores of task n + 1 while the tensor core runs P.V of task n.
// (Measured steps, cfg4 batch 256, per call under ncu: one thread driving TMA and both groups' MMAs 63 us -- ~4 k cycles of
// dependent single-thread issue per task, samples spread evenly over its code; one issuer per group + a TMA warp 57 us; the
// mma.sync kernel 56 us.)  Warps 0..15 softmax / epilogue (TMEM lane quarter = warp % 4), warps 16 / 17 MMA issuers, warp 18 TMA.
#include "../../include/ftc_b200.h"
#include "tc_common.cuh"
#include "tma_util.cuh"
#include "transformer_ops.cuh"

namespace ftc {
namespace {

constexpr int AT_ROWS = 128;                 // query rows / key rows per box
constexpr uint32_t AT_TILE = 128u * 128u;    // bytes of one (64 ch x 128 rows) bf16 box
constexpr uint32_t AT_STAGE = 3u * AT_TILE;  // Q, K, V
constexpr int AT_NSTAGE = 3;
constexpr int AT_THREADS = 19 * 32;            // 16 softmax warps, 2 MMA issuers, 1 TMA producer
// barrier slots
constexpr int AB_FULL = 0, AB_EMPTY = 3, AB_SREADY = 6 /* [group][S buffer] */, AB_PREADY = 10, AB_OREADY = 12, AB_OCONS = 14, AB_COUNT = 16;
// shared memory behind the operand tiles (floats): key mask [2 groups][2][128], row-max halves and row-sum halves [2][2][2][128]
constexpr uint32_t AT_MS_FLOATS = 2 * 2 * 128, AT_X_FLOATS = 2 * 2 * 2 * 128;
constexpr uint32_t AT_TAIL_OFF = AT_NSTAGE * AT_STAGE + 2u * 2u * AT_TILE;
constexpr uint32_t AT_SMEM = 1024 + AT_TAIL_OFF + (AT_MS_FLOATS + 2 * AT_X_FLOATS) * 4 + 8 * AB_COUNT + 16;

struct AttTcParams {
  int B, heads, Lt, Ls, Np;                  // Np = keys rounded up to 16
  int n_slabs, n_items;
  const float* mask;                         // [B, Ls] additive or null
  bf16* out; int out_stride;
  float scale_log2;
};

__device__ __forceinline__ uint64_t att_desc_mn_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((AT_TILE >> 4) & 0x3FFFu) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void att_tma_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(0), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int HD>
__global__ void __launch_bounds__(AT_THREADS, 1)
attention_tc_kernel(const __grid_constant__ AttTcParams p, const __grid_constant__ CUtensorMap tmQ,
                    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV) {
  constexpr int NH = 64 / HD;                // heads (tasks) per item
  constexpr int KS = HD / 16;                // k-steps of the score product
  constexpr int W_MMA = 16, W_TMA = 18;      // MMA issuer warps (one per group), TMA producer warp (+ TMEM allocation)
  constexpr int OC = HD / 2;                 // output columns per thread
  extern __shared__ __align__(1024) uint8_t at_smem_raw[];
  const uint32_t raw = smem_u32(at_smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;
  uint8_t* smem = at_smem_raw + (sbase - raw);
  // layout: 3 stages x (Q, K, V) | P[2 groups][2 chunks of 64 keys] | key mask | row-max halves | row-sum halves | barriers
  const uint32_t p_base = sbase + AT_NSTAGE * AT_STAGE;
  float* Ms = reinterpret_cast<float*>(smem + AT_TAIL_OFF);
  float* Xm = Ms + AT_MS_FLOATS;
  float* Xl = Xm + AT_X_FLOATS;
  const uint32_t bar0 = sbase + AT_TAIL_OFF + (AT_MS_FLOATS + 2 * AT_X_FLOATS) * 4;
  auto bar = [&](int slot, int i) { return bar0 + 8u * (uint32_t)(slot + i); };
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + AT_TAIL_OFF + (AT_MS_FLOATS + 2 * AT_X_FLOATS) * 4 + 8 * AB_COUNT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == W_TMA) {
    if (lane == 0) {
      for (int i = 0; i < AT_NSTAGE; ++i) { mbar_init(bar(AB_FULL, i), 1); mbar_init(bar(AB_EMPTY, i), NH); }
      for (int i = 0; i < 4; ++i) mbar_init(bar(AB_SREADY, i), 1);
      for (int i = 0; i < 2; ++i) { mbar_init(bar(AB_PREADY, i), 256); mbar_init(bar(AB_OREADY, i), 1); mbar_init(bar(AB_OCONS, i), 256); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_launch_dependents();
  pdl_wait();

  const int first = blockIdx.x, step = gridDim.x;
  const int n_my = first < p.n_items ? (p.n_items - first + step - 1) / step : 0;   // items of this CTA (local index it)
  // group g runs head g of every item (hd 32) or the items it = 2 n + g (hd 64): its n-th task is item IT0 + n * ITS
  if (warp == W_TMA) {
    // ---------------------------------------------------------------- TMA producer: three boxes per item, three stages ahead
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 1;                             // parity that passes on a fresh barrier
      int item = first;
      for (int j = 0; j < n_my; ++j, item += step) {
        mbar_wait(bar(AB_EMPTY, stage), ph, 0, 2000);   // hinted waits: a spinning control warp takes issue slots from the softmax warps of its scheduler
        const int b = item / p.n_slabs, slab = item - b * p.n_slabs;
        const uint32_t st = sbase + (uint32_t)stage * AT_STAGE, fb = bar(AB_FULL, stage);
        mbar_arrive_expect_tx(fb, AT_STAGE);
        att_tma_3d(st, &tmQ, slab * 64, 0, b, fb);
        att_tma_3d(st + AT_TILE, &tmK, slab * 64, 0, b, fb);
        att_tma_3d(st + 2u * AT_TILE, &tmV, slab * 64, 0, b, fb);
        if (++stage == AT_NSTAGE) { stage = 0; ph ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= W_MMA) {
    // ---------------------------------------------------------------- MMA issuer of group g (one elected thread)
    const int g = warp - W_MMA;
    const int its = NH == 2 ? 1 : 2, it0 = NH == 2 ? 0 : g, hh = NH == 2 ? g : 0;
    const int n_tasks = NH == 2 ? n_my : (n_my - g + 1) / 2;
    if (lane == 0 && n_tasks > 0) {
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Np >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const int ksteps_o = p.Np >> 4;
      // TMEM columns of group g: two score buffers S[k] at g * 256 + k * 128; the output of task n overwrites the first 64
      // columns of ITS score buffer S[n & 1] (read out by the group before it announced P)
      const uint32_t t_g = tmem_base + (uint32_t)(g * 256);
      // descriptors advance in 16-byte units in their low word (shared-memory addresses stay below 2^18)
      const uint64_t dq0 = umma_desc_sw128(sbase + (uint32_t)(hh * HD) * 2u);
      const uint64_t dk0 = umma_desc_sw128(sbase + AT_TILE + (uint32_t)(hh * HD) * 2u);
      const uint64_t dv0 = att_desc_mn_sw128(sbase + 2u * AT_TILE);
      const uint64_t dp0 = umma_desc_sw128(p_base + (uint32_t)g * 2u * AT_TILE);
      int stage = it0 % AT_NSTAGE;                 // stage / parity of the item of the NEXT score product
      uint32_t ph = 0;
      int pv_stage = stage;                        // stage of the item of the next P.V product
      auto issue_s = [&](int k) {                  // score product of task k into S[k & 1]
        mbar_wait(bar(AB_FULL, stage), ph, 0, 1000);
        tc_fence_after();
        const uint64_t so = (uint64_t)((uint32_t)stage * (AT_STAGE >> 4));
#pragma unroll
        for (int q = 0; q < KS; ++q)
          umma_f16(t_g + (uint32_t)((k & 1) * 128), dq0 + so + (uint64_t)(q * 2), dk0 + so + (uint64_t)(q * 2), idesc_s, q ? 1u : 0u);
        umma_commit(bar(AB_SREADY, g * 2 + (k & 1)));
        stage += its;
        if (stage >= AT_NSTAGE) { stage -= AT_NSTAGE; ph ^= 1u; }
      };
      issue_s(0);
      if (n_tasks > 1) issue_s(1);
      for (int n = 0; n < n_tasks; ++n) {
        mbar_wait(bar(AB_PREADY, g), (uint32_t)(n & 1), 0, 1000);
        tc_fence_after();
        const uint64_t vo = dv0 + (uint64_t)((uint32_t)pv_stage * (AT_STAGE >> 4));
        const uint32_t t_o = t_g + (uint32_t)((n & 1) * 128);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < ksteps_o)
            umma_f16(t_o, dp0 + (uint64_t)((k >> 2) * (AT_TILE >> 4) + (k & 3) * 2), vo + (uint64_t)(k * (2048 >> 4)), idesc_o, k ? 1u : 0u);
        umma_commit(bar(AB_OREADY, g));
        umma_commit(bar(AB_EMPTY, pv_stage));      // this head's reads of the stage are done when these MMAs complete
        pv_stage += its;
        if (pv_stage >= AT_NSTAGE) pv_stage -= AT_NSTAGE;
        // score product n + 2 reuses S[n & 1]: the group must have read O(n) out of it
        if (n + 2 < n_tasks) {
          mbar_wait(bar(AB_OCONS, g), (uint32_t)(n & 1), 0, 1000);
          issue_s(n + 2);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- softmax + epilogue: two threads per query row
    const int g = warp >> 3, hf = (warp >> 2) & 1, quarter = warp & 3;
    const int r = quarter * 32 + lane;             // query row = TMEM lane
    const int gt = threadIdx.x - g * 256;          // thread index inside the group
    const int its = NH == 2 ? 1 : 2, it0 = NH == 2 ? 0 : g, hh = NH == 2 ? g : 0;
    const int n_tasks = NH == 2 ? n_my : (n_my - g + 1) / 2;
    const int nch = p.Np >> 4, n0 = (nch + 1) >> 1;
    const int c_begin = hf ? n0 : 0, my_n = hf ? nch - n0 : n0;   // this thread's 16-key chunks
    const uint32_t t_g = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * 256);
    uint8_t* prow = smem + AT_NSTAGE * AT_STAGE + (uint32_t)g * 2u * AT_TILE + (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
    float* ms_g = Ms + g * 256;                    // [2][128]
    float* xm_g = Xm + g * 512;                    // [2][2 halves][128]
    float* xl_g = Xl + g * 512;
    const bool key_thread = gt < 128;
    const float* mask_col = (p.mask && gt < p.Ls) ? p.mask + gt : nullptr;
    const float mask_pad = gt < p.Ls ? 0.f : -INFINITY;
    const int item0 = first + it0 * step, item_inc = its * step;
    auto mask_of = [&](int k) -> float {           // key mask value of task k for key gt
      return mask_col ? __ldg(mask_col + (int64_t)((item0 + k * item_inc) / p.n_slabs) * p.Ls) : mask_pad;
    };
    if (n_tasks > 0 && key_thread) ms_g[gt] = mask_of(0);
    float mreg = (n_tasks > 1 && key_thread) ? mask_of(1) : mask_pad;
    asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
    uint32_t sv[64];
    float mx = -INFINITY;
    const f32x2 scale2 = pk2(p.scale_log2, p.scale_log2);
    // phase A of task k: scores out of TMEM, scale + key mask, this thread's half of the row maximum; also publishes the key mask
    // of task k + 1 (loaded one task ahead)
    auto phase_a = [&](int k) {
      const float* ms = ms_g + (k & 1) * 128 + c_begin * 16;
      mbar_wait(bar(AB_SREADY, g * 2 + (k & 1)), (uint32_t)((k >> 1) & 1), 0, 4000);
      tc_fence_after();
      const uint32_t t_s = t_g + (uint32_t)((k & 1) * 128 + c_begin * 16);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < my_n) tmem_ld16(t_s + (uint32_t)(c * 16), *reinterpret_cast<uint32_t(*)[16]>(&sv[c * 16]));
      tmem_ld_wait();
      mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < my_n) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 mk = *reinterpret_cast<const float4*>(ms + c * 16 + j);
            const f32x2 a = ffma2(pk2(__uint_as_float(sv[c * 16 + j]), __uint_as_float(sv[c * 16 + j + 1])), scale2, pk2(mk.x, mk.y));
            const f32x2 bq = ffma2(pk2(__uint_as_float(sv[c * 16 + j + 2]), __uint_as_float(sv[c * 16 + j + 3])), scale2, pk2(mk.z, mk.w));
            float a0, a1, b0, b1;
            upk2(a, a0, a1); upk2(bq, b0, b1);
            sv[c * 16 + j] = __float_as_uint(a0); sv[c * 16 + j + 1] = __float_as_uint(a1);
            sv[c * 16 + j + 2] = __float_as_uint(b0); sv[c * 16 + j + 3] = __float_as_uint(b1);
            mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(b0, b1)));
          }
        }
      xm_g[(k & 1) * 256 + hf * 128 + r] = mx;
      if (k + 1 < n_tasks && key_thread) ms_g[((k + 1) & 1) * 128 + gt] = mreg;
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
      mx = fmaxf(mx, xm_g[(k & 1) * 256 + (hf ^ 1) * 128 + r]);
    };
    if (n_tasks > 0) phase_a(0);
    for (int n = 0; n < n_tasks; ++n) {
      const int par = n & 1;
      if (n + 2 < n_tasks && key_thread) mreg = mask_of(n + 2);   // in flight under the exponentials; published by phase A of n + 1
      // ---- phase B: exponentials, row-sum half, P tile
      const float msafe = mx == -INFINITY ? 0.f : mx;       // fully masked row: every p is 0, the row sum 0 (NaN output, as SDPA)
      const f32x2 negm2 = pk2(-msafe, -msafe);
      f32x2 l2 = pk2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c < my_n) {
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float d0, d1;
            upk2(fadd2(pk2(__uint_as_float(sv[c * 16 + 2 * j]), __uint_as_float(sv[c * 16 + 2 * j + 1])), negm2), d0, d1);
            const float p0 = ex2_approx(d0), p1 = ex2_approx(d1);
            l2 = fadd2(l2, pk2(p0, p1));
            w[j] = pack_bf16x2(p0, p1);
          }
          // keys cg*16 .. cg*16+15 = two 16-byte chunks of row r in the K-major 128B-swizzled P tile (64 keys per 128-byte row)
          const int cg = c_begin + c;
          uint8_t* pc = prow + (uint32_t)(cg >> 2) * AT_TILE;
          const int c16 = (cg & 3) * 2;
          *reinterpret_cast<uint4*>(pc + (uint32_t)((c16 ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
          *reinterpret_cast<uint4*>(pc + (uint32_t)(((c16 + 1) ^ (r & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
        }
      float l;
      { float la, lb; upk2(l2, la, lb); l = la + lb; }
      float* xl = xl_g + par * 256;
      xl[hf * 128 + r] = l;
      tc_fence_before();
      fence_proxy_async();                                   // P (generic-proxy stores) -> visible to the tensor core's reads
      mbar_arrive(bar(AB_PREADY, g));
      // ---- phase A of the NEXT task runs under this task's P.V product (its scores are in the other S buffer)
      if (n + 1 < n_tasks) phase_a(n + 1);
      // ---- phase C: O row (this thread's half of the head's channels), divide by the row sum, store
      mbar_wait(bar(AB_OREADY, g), (uint32_t)par, 0, 4000);
      tc_fence_after();
      uint32_t ov[OC];
      const uint32_t t_o = t_g + (uint32_t)(par * 128 + hh * HD + hf * OC);
#pragma unroll
      for (int c = 0; c < OC / 16; ++c) tmem_ld16(t_o + (uint32_t)(c * 16), *reinterpret_cast<uint32_t(*)[16]>(&ov[c * 16]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar(AB_OCONS, g));
      asm volatile("bar.sync %0, 256;" ::"r"(3 + g) : "memory");   // both halves of every row sum are in shared memory
      l += xl[(hf ^ 1) * 128 + r];
      if (r < p.Lt) {
        const int item = item0 + n * item_inc;
        const int b = item / p.n_slabs, slab = item - b * p.n_slabs;
        const float inv = 1.f / l;
        bf16* op = p.out + ((int64_t)b * p.Lt + r) * p.out_stride + (slab * NH + hh) * HD + hf * OC;
#pragma unroll
        for (int c = 0; c < OC; c += 8) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ov[c]) * inv, __uint_as_float(ov[c + 1]) * inv);
          o.y = pack_bf16x2(__uint_as_float(ov[c + 2]) * inv, __uint_as_float(ov[c + 3]) * inv);
          o.z = pack_bf16x2(__uint_as_float(ov[c + 4]) * inv, __uint_as_float(ov[c + 5]) * inv);
          o.w = pack_bf16x2(__uint_as_float(ov[c + 6]) * inv, __uint_as_float(ov[c + 7]) * inv);
          *reinterpret_cast<uint4*>(op + c) = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == W_TMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int g_att_sms = 0;

template <int HD>
int launch_attention_tc(const AttTcParams& p, const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, cudaStream_t s) {
  const size_t smem = AT_SMEM;
  static bool attr_done = false;
  if (!attr_done) {
    FTC_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  const int grid = p.n_items < g_att_sms ? p.n_items : g_att_sms;
  FTC_CHECK_CUDA(launch_pdl(attention_tc_kernel<HD>, dim3(grid), dim3(AT_THREADS), smem, s, p, tmQ, tmK, tmV));
  FTC_POST_LAUNCH();
  return 0;
}

}  // namespace

// returns 1 when the shape is not one this kernel takes (the caller falls back to the mma.sync kernel), 0 on success, < 0 on error
int attention_tc(const void* q, int q_stride, int q_off, const void* k, const void* v, int kv_stride, int k_off, int v_off,
                 const float* mask, void* out, int out_stride, int B, int heads, int hd, int Lt, int Ls, cudaStream_t s) {
  static const int env = [] { const char* e = getenv("FTC_ATT_TC"); return e ? atoi(e) : 1; }();
  if (!env) return 1;
  const int C = heads * hd;
  if ((hd != 32 && hd != 64) || C % 64 != 0 || Lt > AT_ROWS || Ls > AT_ROWS || Lt < 1 || Ls < 1) return 1;
  if (q_stride % 8 || q_off % 8 || kv_stride % 8 || k_off % 8 || v_off % 8 || out_stride % 8) return 1;
  const bf16* qb = reinterpret_cast<const bf16*>(q) + q_off;
  const bf16* kb = reinterpret_cast<const bf16*>(k) + k_off;
  const bf16* vb = reinterpret_cast<const bf16*>(v) + v_off;
  if ((reinterpret_cast<uintptr_t>(qb) & 15) || (reinterpret_cast<uintptr_t>(kb) & 15) || (reinterpret_cast<uintptr_t>(vb) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 15))
    return 1;
  if (g_att_sms == 0) {
    int dev = 0;
    FTC_CHECK_CUDA(cudaGetDevice(&dev));
    FTC_CHECK_CUDA(cudaDeviceGetAttribute(&g_att_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  alignas(64) CUtensorMap tmQ, tmK, tmV;
  int rc = tma_encode_nhwc(&tmQ, qb, DT_BF16, (uint64_t)C, (uint64_t)q_stride, (uint64_t)Lt, 1, (uint64_t)B, 64, AT_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = tma_encode_nhwc(&tmK, kb, DT_BF16, (uint64_t)C, (uint64_t)kv_stride, (uint64_t)Ls, 1, (uint64_t)B, 64, AT_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = tma_encode_nhwc(&tmV, vb, DT_BF16, (uint64_t)C, (uint64_t)kv_stride, (uint64_t)Ls, 1, (uint64_t)B, 64, AT_ROWS, 1, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  AttTcParams p;
  p.B = B; p.heads = heads; p.Lt = Lt; p.Ls = Ls; p.Np = (Ls + 15) & ~15;
  p.n_slabs = C / 64; p.n_items = B * p.n_slabs;
  p.mask = mask; p.out = reinterpret_cast<bf16*>(out); p.out_stride = out_stride;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
  return hd == 32 ? launch_attention_tc<32>(p, tmQ, tmK, tmV, s) : launch_attention_tc<64>(p, tmQ, tmK, tmV, s);
}

}  // namespace ftc
