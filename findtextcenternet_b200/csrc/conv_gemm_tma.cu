// tcgen05 / TMEM implicit-GEMM convolution with TMA tensor-tile operand staging (sm_100a, bf16 -> fp32 accumulate).
//
// Same math, parameters and epilogue as conv_gemm_tc.cu; what changes is how operand A reaches shared memory:
//   TMA_ROWS (1x1 conv / Linear): A is the [M, C] activation matrix; one cp.async.bulk.tensor box of (128*MT rows x 64
//             channels) per k-block lands directly in the 128B-swizzled K-major layout tcgen05.mma reads.
//   TMA_HALO (3x3, stride 1, pad 1, W % 16 == 0): the CTA tile is a 16-wide x 8*MT-high pixel patch of ONE image.  For
//             each (64-channel chunk, column shift kx) ONE box of (8*MT+2 rows x 16 px x 64 ch) is loaded -- TMA's
//             out-of-bounds zero fill IS the conv padding -- and serves the three row taps ky = 0,1,2 as three UMMA
//             descriptors whose start address differs by 16 px * 128 B = 2 KB (a multiple of the 1 KB swizzle atom).
//             Operand A therefore crosses L2->SM 3*(1+2/(8*MT)) times instead of the 9 times of an im2col gather.
// Roles (14 warps, one persistent CTA per SM): warps 0-7 epilogue, warp 8 TMA producer (one lane), warp 9 MMA issuer
// (one lane), warps 10-13 SE scalers (only when the A operand carries a per-(image, channel) squeeze-excite factor:
// they rescale the landed tile in shared memory between the TMA completion and the MMA).  A and B (weights, one
// cp.async.bulk of the pre-swizzled block per k-block) travel through two independent mbarrier rings, so a halo tile
// stays resident for its three k-blocks while weight blocks stream underneath.
#include <cuda.h>

#include "tc_common.cuh"

namespace ftc {

namespace {

constexpr int TM_BM = 128;
constexpr int TM_EPI_WARPS = 8;
constexpr int TM_EPI_THREADS = TM_EPI_WARPS * 32;
constexpr int TM_TMA_WARP = 8, TM_MMA_WARP = 9, TM_SCALE_WARP0 = 10, TM_SCALE_WARPS = 4;
constexpr int TM_THREADS = (TM_EPI_WARPS + 2 + TM_SCALE_WARPS) * 32;   // 448
constexpr int TM_STAGED_FLOATS = 10 * 256;     // per-tile scale[BN] + bias[ncase<=9][BN]
constexpr int TM_MAX_SLOTS = 12;   // ring depth limit per operand (12: a whole 3x3 single-chunk weight set can stay resident)
constexpr int TM_MAX_EPI_WARPS = TM_EPI_WARPS + TM_SCALE_WARPS;     // without SE the scaler warps join the staged epilogue
constexpr int TM_NBARS = 5 * TM_MAX_SLOTS + 4 + 2 * TM_MAX_EPI_WARPS;   // + per-warp residual ring (2 deep)
constexpr uint32_t TM_BOX_BYTES = 32 * 64;   // one staged epilogue box: 32 rows x 32 bf16 channels
constexpr uint32_t TM_SUB_BYTES = TM_BM * 128;  // one 128-row sub-tile of A = 16 KB

struct TmaLaunch {
  int nA, nB;             // ring depths (A slots, B slots)
  uint32_t a_slot_bytes;  // bytes of one A slot
  int NKG, nsub;          // k-groups per tile; k-blocks per k-group (1 rows, 3 halo)
  int tiles_x, tiles_y;   // halo: 16 x (8*MT) tiles per image
  int nbuf;               // TMEM accumulator sets (2: epilogue of tile i overlaps main loop of tile i+1)
  int se_tab;             // SE: channels per image of the shared-memory bf16 scale table (0 = per-row global loads)
  int staged;             // epilogue: 1 = bf16 NHWC output staged in shared memory and written with TMA tensor stores
  int res_tma;            // staged epilogue: 1 = the residual tile is prefetched with TMA loads (per-warp ring)
  int box_depth;          // staged epilogue: boxes per warp for the output (and again for the residual ring): 1 or 2
  int bstat;              // weight-stationary schedule: a CTA walks a CONTIGUOUS range of tiles in (group, n-tile)-major
                          // order and keeps the whole [BN x K] weight tile resident in its nB = NKB slots across m-tiles
  int m_tiles;
  int n_epi;              // epilogue warps: 8, or 12 when the (idle) scaler warps join the staged epilogue
  // exact division by multiply-shift (q = (t * magic) >> 40, t < 2^23, divisor <= 2^17): a tile decode costs four integer
  // divisions in every role of every CTA per tile, ~120 instructions each time with the hardware-less 32-bit divide
  unsigned long long mg_per_m, mg_nt, mg_per_img, mg_tiles_x, mg_m_tiles;
};
__host__ __device__ inline unsigned long long div_magic(int d) { return ((1ull << 40) + (unsigned long long)d - 1) / (unsigned long long)d; }
__device__ __forceinline__ int fast_div(int t, unsigned long long magic) { return (int)(((unsigned long long)(unsigned)t * magic) >> 40); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// pipeline trace of CTA 0 (-DFTC_ABLATION builds only): role r writes stamp j of its i-th tile to trace[r * 1024 + i * 4 + j]
#ifdef FTC_ABLATION
#define FTC_TRACE(role, i, j) do { if (p.trace && blockIdx.x == 0 && (i) < 256) p.trace[(role) * 1024 + (i) * 4 + (j)] = clock64(); } while (0)
#else
#define FTC_TRACE(role, i, j) do { } while (0)
#endif
struct TmaTile { int m0, b, y0, x0, g, nt; };
template <int MT, bool HALO>
__device__ __forceinline__ TmaTile decode_tma_tile(int tile, int NT, int G, const TmaLaunch& L) {
  const int per_m = NT * G;
  int mt, rest;
  if (L.bstat) { rest = fast_div(tile, L.mg_m_tiles); mt = tile - rest * L.m_tiles; }
  else { mt = fast_div(tile, L.mg_per_m); rest = tile - mt * per_m; }
  TmaTile t;
  t.g = fast_div(rest, L.mg_nt);
  t.nt = rest - t.g * NT;
  t.m0 = mt * (MT * TM_BM);
  t.b = 0; t.y0 = 0; t.x0 = 0;
  if (HALO) {
    const int per_img = L.tiles_x * L.tiles_y;
    t.b = fast_div(mt, L.mg_per_img);
    const int r = mt - t.b * per_img;
    const int ty = fast_div(r, L.mg_tiles_x);
    t.y0 = ty * (8 * MT);
    t.x0 = (r - ty * L.tiles_x) * HALO_TW;
  }
  return t;
}

template <bool SE, int MT, bool HALO>
__global__ void __launch_bounds__(TM_THREADS, 1)
conv_gemm_tma_kernel(const __grid_constant__ ConvGemmParams p, const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                     const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ TmaLaunch L, const int num_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;          // SWIZZLE_128B needs 1024 B alignment
  uint8_t* smem = smem_raw + (sbase - raw);
  const int BN = p.tc.BN, NT = p.tc.NT, G = p.G;
  const bool kb32 = HALO && p.tc.kb32 != 0;               // 32-channel k-blocks: 64-byte operand rows, SWIZZLE_64B
  const uint32_t rowb = kb32 ? 64u : 128u;
  const uint32_t b_bytes = (uint32_t)BN * rowb;
  const uint32_t a_ring = sbase;
  const uint32_t b_ring = sbase + (uint32_t)L.nA * L.a_slot_bytes;
  const uint32_t ring_bytes = (uint32_t)L.nA * L.a_slot_bytes + (uint32_t)L.nB * b_bytes;
  // staged epilogue: per epilogue warp two output boxes (+ two residual boxes), 2 KB each, right behind the rings
  const uint32_t box_base = sbase + ring_bytes;
  const uint32_t box_bytes = L.staged ? (uint32_t)L.n_epi * (L.res_tma ? 2u : 1u) * (uint32_t)L.box_depth * TM_BOX_BYTES : 0u;
  const uint32_t bar0 = box_base + box_bytes;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (TM_MAX_SLOTS + s); };
  auto a_raw = [&](int s) { return bar0 + 8u * (2 * TM_MAX_SLOTS + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (3 * TM_MAX_SLOTS + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (4 * TM_MAX_SLOTS + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (5 * TM_MAX_SLOTS + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (5 * TM_MAX_SLOTS + 2 + a); };
  auto res_bar = [&](int w, int i) { return bar0 + 8u * (5 * TM_MAX_SLOTS + 4 + 2 * w + i); };
  uint8_t* after_bars = smem + ring_bytes + box_bytes + 8 * TM_NBARS;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(after_bars);
  float* staged = reinterpret_cast<float*>(after_bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == TM_TMA_WARP && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (L.staged) tma_prefetch_desc(&tmOut);
    if (L.res_tma) tma_prefetch_desc(&tmRes);
  }
  if (warp == TM_MMA_WARP) {
    if (lane == 0) {
      for (int s = 0; s < L.nA; ++s) {
        mbar_init(a_full(s), SE ? TM_SCALE_WARPS : 1);
        mbar_init(a_empty(s), 1);
        mbar_init(a_raw(s), 1);
      }
      for (int s = 0; s < L.nB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), (uint32_t)L.n_epi); }
      for (int w = 0; w < TM_MAX_EPI_WARPS; ++w) { mbar_init(res_bar(w, 0), 1); mbar_init(res_bar(w, 1), 1); }
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
  pdl_launch_dependents();
  pdl_wait();                 // everything above (barriers, TMEM, descriptor prefetch) overlapped the previous kernel's tail
  const int hw = p.Ho * p.Wo;
  const int NKG = L.NKG, nsub = L.nsub;
  int t_first = blockIdx.x, t_last = num_tiles, t_step = gridDim.x;
  if (L.bstat) {
    t_first = (int)((long long)blockIdx.x * num_tiles / gridDim.x);
    t_last = (int)((long long)(blockIdx.x + 1) * num_tiles / gridDim.x);
    t_step = 1;
  }

  if (warp == TM_TMA_WARP) {
    // ------------------------------------------------------------------ TMA producer (A tensor tiles + B bulk copies)
    if (lane == 0) {
      const bf16* wgt = reinterpret_cast<const bf16*>(p.w);
      const int nGA = p.tc.nGA;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int prev_gn = -1;
      int tr_i = 0;
      for (int tile = t_first; tile < t_last; tile += t_step, ++tr_i) {
        FTC_TRACE(0, tr_i, 0);
        const TmaTile tc = decode_tma_tile<MT, HALO>(tile, NT, G, L);
        const size_t kel = kb32 ? 32 : 64;                  // elements per weight row of a k-block
        const bf16* wtile = wgt + (size_t)(tc.g * NT + tc.nt) * p.tc.NKB * ((size_t)BN * kel);
        // weight-stationary: slot kb holds k-block kb; it is (re)loaded only when the CTA moves to another n-tile.  The
        // ring bookkeeping below then advances exactly one lap per reload, so the phase logic is the ring's own.
        const bool loadB = !L.bstat || (tc.g * NT + tc.nt) != prev_gn;
        prev_gn = tc.g * NT + tc.nt;
        int kb = 0;
        for (int kg = 0; kg < NKG; ++kg) {
          int chunk = kg, dx = 0;
          bool srcb = false;
          if (HALO) {
            const int q = kg / 3;
            dx = kg - 3 * q;
            srcb = q >= nGA;
            chunk = srcb ? q - nGA : q;
          }
          mbar_wait(a_empty(as), aph ^ 1u);
          if (kg == 0) FTC_TRACE(0, tr_i, 1);
          if (kg == NKG - 1) FTC_TRACE(0, tr_i, 2);
          const uint32_t bar = SE ? a_raw(as) : a_full(as);
          const CUtensorMap* tm = srcb ? &tmB : &tmA;
          const int c = (srcb ? p.b_ch_off + tc.g * p.b_group_stride : p.a_ch_off) + chunk * 64;
          if FTC_ABL(512) mbar_arrive(bar);                  // ablation: no operand-A traffic
          else {
            mbar_arrive_expect_tx(bar, L.a_slot_bytes);
            if (HALO) tma_load_4d(a_ring + (uint32_t)as * L.a_slot_bytes, tm, c, tc.x0 + dx - 1, tc.y0 - 1, tc.b, bar);
            else tma_load_4d(a_ring + (uint32_t)as * L.a_slot_bytes, tm, c, tc.m0, 0, 0, bar);
          }
          as = (as + 1 == L.nA) ? 0 : as + 1;
          aph ^= (as == 0) ? 1u : 0u;
          for (int sub = 0; sub < nsub && loadB; ++sub, ++kb) {
            mbar_wait(b_empty(bs), bph ^ 1u);
            mbar_arrive_expect_tx(b_full(bs), b_bytes);
            bulk_copy_g2s(b_ring + (uint32_t)bs * b_bytes, wtile + (size_t)kb * BN * kel, b_bytes, b_full(bs));
            bs = (bs + 1 == L.nB) ? 0 : bs + 1;
            bph ^= (bs == 0) ? 1u : 0u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == TM_MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp walks the loop (barrier waits, descriptor arithmetic: warp-uniform, so it stays on the uniform datapath)
    // and one elected lane issues the tcgen05 instructions.  With everything under `if (lane == 0)` the compiler moved
    // every descriptor through ELECT / R2UR.BROADCAST sequences: ~17 dependent instructions per MMA, which made the N = 32
    // convs of stage 1 MMA-ISSUE bound (72 MMAs of 16 tensor-cycles each per tile).
    const bool leader = elect_one_sync();
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM_BM >> 4) << 24);
      // loop-invariant descriptor pieces: swizzle mode / SBO / version bits, and every byte offset in 16-byte units
      const uint64_t desc0 = kb32 ? umma_desc_sw64(0) : umma_desc_sw128(0);
      const uint32_t a_ring16 = a_ring >> 4, b_ring16 = b_ring >> 4, a_slot16 = L.a_slot_bytes >> 4, b16 = b_bytes >> 4;
      const uint32_t tap16 = (HALO_TW * rowb) >> 4, sub16 = (TM_BM * rowb) >> 4;
      int ksteps = kb32 ? 2 : 4;
      asm volatile("" : "+r"(ksteps));   // keep it in a register: the compiler otherwise re-reads the plan flag per k-block
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      uint32_t titer = 0;
      int prev_gn = -1;
      for (int tile = t_first; tile < t_last; tile += t_step, ++titer) {
        const uint32_t buf = L.nbuf == 2 ? (titer & 1u) : 0u;
        const uint32_t use = L.nbuf == 2 ? (titer >> 1) : titer;
        bool newB = true, lastB = true;       // weight-stationary: wait for the weights once, release them on the last m-tile
        if (L.bstat) {
          const int gn = fast_div(tile, L.mg_m_tiles);
          newB = gn != prev_gn;
          prev_gn = gn;
          lastB = tile + 1 >= t_last || fast_div(tile + 1, L.mg_m_tiles) != gn;
        }
        if (leader) FTC_TRACE(1, (int)titer, 0);
        mbar_wait(tempty_bar(buf), (use & 1u) ^ 1u);
        if (leader) FTC_TRACE(1, (int)titer, 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * (uint32_t)(MT * BN);
        uint32_t accum = 0;
        for (int kg = 0; kg < NKG; ++kg) {
          mbar_wait(a_full(as), aph);
          if (leader && kg == 0) FTC_TRACE(1, (int)titer, 2);
          const uint64_t a_desc = desc0 + (uint64_t)(a_ring16 + (uint32_t)as * a_slot16);
          for (int sub = 0; sub < nsub; ++sub) {
            if (newB) mbar_wait(b_full(bs), bph);
            tc_fence_after();
            // descriptors advance by plain 64-bit adds in 16-byte units (shared addresses stay below 256 KB, so the 14-bit
            // start-address field never carries): the issuing warp's dependent uniform-register chain per k-block is what
            // paces small-N tiles, not the tensor pipe
            const uint64_t bdesc = desc0 + (uint64_t)(b_ring16 + (uint32_t)bs * b16);
            const uint64_t adesc = a_desc + (HALO ? (uint64_t)((uint32_t)sub * tap16) : 0ull);
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // (K = 16 bf16 = 32 B) steps inside the swizzle row: +2 in 16-byte units
              if (k >= ksteps) break;       // 64-byte rows hold two steps
              if (FTC_ABL(8192) && k >= 1) break;         // (ablation: one K step per k-block)
#pragma unroll
              for (int h = 0; h < MT; ++h) {
                if (FTC_ABL(32768) && h >= 1) break;      // (ablation: one accumulator)
                if (leader && !FTC_ABL(4096))             // (ablation bit: no MMAs, commits still fire)
                umma_f16(d_tmem + (uint32_t)h * (uint32_t)BN, adesc + (uint64_t)((uint32_t)h * sub16 + 2u * k), bdesc + (uint64_t)(2 * k),
                         idesc, (k == 0 && h < MT) ? accum : 1u);
              }
              accum = 1u;
            }
            if (lastB && leader) umma_commit(b_empty(bs));
            bs = (bs + 1 == L.nB) ? 0 : bs + 1;
            if (!L.bstat || lastB) bph ^= (bs == 0) ? 1u : 0u;
          }
          if (leader) umma_commit(a_empty(as));
          as = (as + 1 == L.nA) ? 0 : as + 1;
          aph ^= (as == 0) ? 1u : 0u;
        }
        if (leader) umma_commit(tfull_bar(buf));
        if (leader) FTC_TRACE(1, (int)titer, 3);
      }
    }
    __syncwarp();
  } else if (warp >= TM_SCALE_WARP0 && (SE || L.n_epi <= TM_EPI_WARPS)) {
    // ------------------------------------------------------------------ SE scalers (rows mode only)
    if (SE) {
      const int t = threadIdx.x - TM_SCALE_WARP0 * 32;   // 0..127
      const int j = t & 7;                               // 16-byte chunk (8 channels) of the 128-byte row
      const int rb = t >> 3;                             // rows rb + 16*i
      const uint32_t row_off = (uint32_t)(rb >> 3) * 1024u + (uint32_t)(rb & 7) * 128u + (uint32_t)((j ^ (rb & 7)) << 4);
      constexpr int ROWS = 8 * MT;
      int as = 0;
      uint32_t aph = 0;
      if (L.se_tab) {
        // fast path (a tile spans at most two images): the scale rows of both images are staged once per tile in shared
        // memory as bf16 and applied with packed bf16 multiplies -- 7 instructions per 16-byte chunk instead of ~45
        // (two dependent global loads, 8 unpack / FMUL / pack).  The scale is rounded to bf16 (the activation it
        // multiplies already is); the product is rounded once, as before.
        __nv_bfloat16* sc_tab = reinterpret_cast<__nv_bfloat16*>(staged + TM_STAGED_FLOATS);
        const int CAp = L.se_tab;
        for (int tile = t_first; tile < t_last; tile += t_step) {
          const TmaTile tc = decode_tma_tile<MT, HALO>(tile, NT, G, L);
          const int b0 = tc.m0 / hw;
          const int nb = (b0 + 1) * hw - tc.m0;          // tile rows [0, nb) are image b0, the rest image b0 + 1
          const int b1 = min(b0 + 1, p.B - 1);
          asm volatile("bar.sync 2, %0;" ::"n"(TM_SCALE_WARPS * 32) : "memory");   // previous tile's table reads are done
          for (int idx = t * 4; idx < 2 * CAp; idx += TM_SCALE_WARPS * 32 * 4) {
            const int second = idx >= CAp ? 1 : 0;
            const int c = idx - second * CAp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < p.CA) v = __ldg(reinterpret_cast<const float4*>(p.a_scale + (int64_t)(second ? b1 : b0) * p.a_scale_stride + c));
            __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(sc_tab + idx);
            dst[0] = __floats2bfloat162_rn(v.x, v.y);
            dst[1] = __floats2bfloat162_rn(v.z, v.w);
          }
          asm volatile("bar.sync 2, %0;" ::"n"(TM_SCALE_WARPS * 32) : "memory");
          for (int kg = 0; kg < NKG; ++kg) {
            const int c = kg * 64 + j * 8;
            const uint4 s0 = *reinterpret_cast<const uint4*>(sc_tab + c);
            const uint4 s1 = *reinterpret_cast<const uint4*>(sc_tab + CAp + c);
            mbar_wait(a_raw(as), aph);
            if (!FTC_ABL(64)) {
              const uint32_t a_dst = a_ring + (uint32_t)as * L.a_slot_bytes + row_off;
#pragma unroll
              for (int i = 0; i < ROWS; ++i) {
                const uint32_t addr = a_dst + (uint32_t)i * 2048u;
                const uint4 sc = (rb + 16 * i < nb) ? s0 : s1;
                uint4 u;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
                const __nv_bfloat162* sh = reinterpret_cast<const __nv_bfloat162*>(&sc);
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __hmul2(h[e], sh[e]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(as));
            as = (as + 1 == L.nA) ? 0 : as + 1;
            aph ^= (as == 0) ? 1u : 0u;
          }
        }
      } else
      for (int tile = t_first; tile < t_last; tile += t_step) {
        const TmaTile tc = decode_tma_tile<MT, HALO>(tile, NT, G, L);
        int img[ROWS];
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          int m = tc.m0 + rb + 16 * i;
          m = m < p.M ? m : p.M - 1;
          img[i] = m / hw;
        }
        for (int kg = 0; kg < NKG; ++kg) {
          mbar_wait(a_raw(as), aph);
          const int c = kg * 64 + j * 8;
          if (c < p.CA && !FTC_ABL(64)) {
            const uint32_t a_dst = a_ring + (uint32_t)as * L.a_slot_bytes + row_off;
#pragma unroll
            for (int i = 0; i < ROWS; ++i) {
              const uint32_t addr = a_dst + (uint32_t)i * 2048u;
              uint4 u;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
              const float4* sp = reinterpret_cast<const float4*>(p.a_scale + (int64_t)img[i] * p.a_scale_stride + c);
              const float4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
              float2 f;
              f = __bfloat1622float2(h[0]); h[0] = __floats2bfloat162_rn(f.x * s0.x, f.y * s0.y);
              f = __bfloat1622float2(h[1]); h[1] = __floats2bfloat162_rn(f.x * s0.z, f.y * s0.w);
              f = __bfloat1622float2(h[2]); h[2] = __floats2bfloat162_rn(f.x * s1.x, f.y * s1.y);
              f = __bfloat1622float2(h[3]); h[3] = __floats2bfloat162_rn(f.x * s1.z, f.y * s1.w);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full(as));
          as = (as + 1 == L.nA) ? 0 : as + 1;
          aph ^= (as == 0) ? 1u : 0u;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps: 4 lane quarters x 2 halves)
    const int q = warp & 3;                    // TMEM lane quarter this warp may read (hardware: warp id % 4)
    const int half = warp >> 2;                // direct path: MT == 1: even / odd 16-column chunks; MT == 2: 128-row sub-tile
    const int ewarp = warp < TM_EPI_WARPS ? warp : warp - 2;       // dense epilogue-warp index (warps 10-13 -> 8-11)
    const int ew = warp < TM_EPI_WARPS ? (warp >> 2) : 2;          // index among the warps that share this lane quarter
    const int etid = ewarp * 32 + lane;
    const int n_ethreads = L.n_epi * 32;
    const int row = (MT == 2 ? half * TM_BM : 0) + q * 32 + lane;
    const bf16* res1 = reinterpret_cast<const bf16*>(p.res1);
    const bf16* res2 = reinterpret_cast<const bf16*>(p.res2);
    float* sscale = staged;                    // [BN]
    float* sbias = staged + 256;               // [ncase][BN]
    uint32_t titer = 0;
    if (L.staged) {
      // ---- staged epilogue: every warp owns 32 accumulator rows and walks 32-column chunks.  A chunk is converted to
      // bf16 in a 2 KB shared-memory box (64-byte rows, SWIZZLE_64B so that the 32 row-owning lanes spread over all
      // banks) and leaves through ONE TMA tensor store, which writes whole 64-byte row segments (the direct path wrote
      // 32 scattered 16-byte pieces per store instruction and was LSU-bound).  The residual tile comes in the same way
      // through a two-deep per-warp TMA ring that is issued before the accumulator is waited for.
      // Work items: (128-row sub-tile h, 32-channel box i) of the tile; the 2 (with SE scalers busy) or 3 warps that share a
      // TMEM lane quarter take them round-robin.
      const uint32_t depth = (uint32_t)L.box_depth, dmask = depth - 1u, dshift = depth >> 1;   // depth 1 or 2
      const uint32_t my_boxes = box_base + (uint32_t)ewarp * (L.res_tma ? 2u : 1u) * depth * TM_BOX_BYTES;
      const uint32_t out_box0 = my_boxes, res_box0 = my_boxes + depth * TM_BOX_BYTES;
      const uint32_t sw = (uint32_t)((lane >> 1) & 3);                 // SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3
      const uint32_t row_b = (uint32_t)lane * 64u;
      // a chunk = one output box of 32 channels = 32 accumulator columns (64 for SwiGLU, which gates column pairs)
      const int act = p.act;                                          // hoisted out of the item loop (was re-read from the
      const bool has_res2 = res2 != nullptr;                          // parameter bank per item)
      const int cw = act == ACT_SWIGLU ? 64 : 32;
      const int per_sub = BN / cw, n_items = MT * per_sub, n_share = L.n_epi >> 2;
      const bool res_tma = L.res_tma != 0, res1_direct = p.res1 != nullptr && !res_tma, nine = p.ncase == 9, deep = depth == 2;
      const bool any_direct = !HALO && (res1_direct || has_res2);
      const uint32_t swo[4] = {(0u ^ sw) << 4, (1u ^ sw) << 4, (2u ^ sw) << 4, (3u ^ sw) << 4};   // swizzled 16-byte chunk offsets
      const int Ho1 = p.Ho - 1, Wo1 = p.Wo - 1;
      uint32_t n_out = 0, n_res_issued = 0, n_res_used = 0;
      // SiLU as h + h*tanh(h), h = x/2: the 1/2 is folded into the staged BN scale / bias (exact: a power of two)
      const float act_pre = act == ACT_SILU ? 0.5f : 1.f;
      int staged_gn = -1;
      // Residual boxes are fetched by a load cursor that runs `depth` items AHEAD of the consumer across tile boundaries: the
      // box of the next tile is already in flight while this tile's main loop runs (short-K convs exposed the ~2-3k-cycle
      // HBM latency of a load issued at tile start on every tile).
      int ld_tile = t_first, ld_it = ew;
      auto issue_next_res = [&]() {
        if (ld_tile >= t_last || ew >= n_items) return;
        const TmaTile t2 = decode_tma_tile<MT, HALO>(ld_tile, NT, G, L);
        const int h = (MT == 2 && ld_it >= per_sub) ? 1 : 0, c0 = (ld_it - h * per_sub) * cw;   // MT <= 2: no division
        const uint32_t slot = n_res_issued & dmask;
        mbar_arrive_expect_tx(res_bar(ewarp, slot), TM_BOX_BYTES);
        tma_load_4d(res_box0 + slot * TM_BOX_BYTES, &tmRes, t2.g * p.N + t2.nt * BN + c0, HALO ? t2.x0 : t2.m0 + h * TM_BM + q * 32,
                    HALO ? t2.y0 + h * 8 + q * 2 : 0, HALO ? t2.b : 0, res_bar(ewarp, slot));
        ++n_res_issued;
        ld_it += n_share;
        if (ld_it >= n_items) { ld_it = ew; ld_tile += t_step; }
      };
      if (L.res_tma && lane == 0)
        for (uint32_t d = 0; d < depth; ++d) issue_next_res();
      for (int tile = t_first; tile < t_last; tile += t_step, ++titer) {
        const TmaTile tc = decode_tma_tile<MT, HALO>(tile, NT, G, L);
        const uint32_t buf = L.nbuf == 2 ? (titer & 1u) : 0u;
        const uint32_t use = L.nbuf == 2 ? (titer >> 1) : titer;
        const int ch_out = p.out_ch_base[tc.g] + (act == ACT_SWIGLU ? (tc.nt * BN) >> 1 : tc.nt * BN);
        // box coordinates of the 32 rows (q, sub-tile h) of this warp
        auto box_k1 = [&](int h) { return HALO ? tc.x0 : tc.m0 + h * TM_BM + q * 32; };
        auto box_k2 = [&](int h) { return HALO ? tc.y0 + h * 8 + q * 2 : 0; };
        const int k3 = HALO ? tc.b : 0;
        // per-column scale / bias of this (group, n-tile): staged once and kept while consecutive tiles share it (always
        // under the weight-stationary schedule; whenever NT * G == 1)
        if (tc.g * NT + tc.nt != staged_gn) {
          staged_gn = tc.g * NT + tc.nt;
          asm volatile("bar.sync 1, %0;" ::"r"(n_ethreads) : "memory");
          const int ncol0 = tc.nt * BN;
          for (int c = etid; c < BN; c += n_ethreads) {
            const int n = ncol0 + c;
            sscale[c] = act_pre * ((p.scale && n < p.N) ? __ldg(p.scale + (int64_t)tc.g * p.N + n) : 1.f);
          }
          for (int c = etid; c < p.ncase * BN; c += n_ethreads) {
            const int cs_ = c / BN, cc = c - cs_ * BN;
            const int n = ncol0 + cc;
            sbias[c] = act_pre * ((p.bias_tab && n < p.N) ? __ldg(p.bias_tab + ((int64_t)cs_ * G + tc.g) * p.N + n) : 0.f);
          }
          asm volatile("bar.sync 1, %0;" ::"r"(n_ethreads) : "memory");
        }
        if (warp == 0 && lane == 0) FTC_TRACE(2, (int)titer, 0);
        if (lane == 0) mbar_wait(tfull_bar(buf), use & 1u);
        __syncwarp();
        if (warp == 0 && lane == 0) FTC_TRACE(2, (int)titer, 1);
        tc_fence_after();
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (uint32_t)(MT * BN);
        auto t_item = [&](int it) {                     // TMEM address of an item: sub-tile h at column h * BN, box at c0
          const int h = (MT == 2 && it >= per_sub) ? 1 : 0;
          return t_base + (uint32_t)(h * BN + (it - h * per_sub) * cw);
        };
        uint32_t raw[32];
        bool released = false;
        if (act != ACT_SWIGLU && ew < n_items) {
          __syncwarp();
          tmem_ld32(t_item(ew), raw);
        }
        for (int it = ew; it < n_items; it += n_share) {
          const int h = (MT == 2 && it >= per_sub) ? 1 : 0, c0 = (it - h * per_sub) * cw;
          const uint32_t t_addr = t_base + (uint32_t)(h * BN);
          const int k1 = box_k1(h), k2 = box_k2(h);
          const int row_t = h * TM_BM + q * 32 + lane;          // this lane's row inside the tile
          const int m_row = (HALO ? 0 : tc.m0) + row_t;         // rows mode: output row (direct residual loads)
          int cs = 0;
          if (nine) {
            const int oy = tc.y0 + (row_t >> 4), ox = tc.x0 + (row_t & 15);   // ncase 9 only occurs on halo (3x3) tiles
            cs = (oy == 0 ? 0 : (oy == Ho1 ? 2 : 1)) * 3 + (ox == 0 ? 0 : (ox == Wo1 ? 2 : 1));
          }
          f32x2 vv[16];
          if (act == ACT_SWIGLU) {
            // interleaved (x1, xg) column pairs -> x1 * silu(xg): 64 accumulator columns give this box's 32 outputs
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              __syncwarp();
              tmem_ld32(t_addr + (uint32_t)(c0 + hh * 32), raw);
              tmem_ld_wait();
              const ulonglong2* sb = reinterpret_cast<const ulonglong2*>(sbias + cs * BN + c0 + hh * 32);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const ulonglong2 bi = sb[e];
                float x1a, xga, x1b, xgb, ta, tb;
                upk2(fadd2(pk2(__uint_as_float(raw[4 * e]), __uint_as_float(raw[4 * e + 1])), bi.x), x1a, xga);
                upk2(fadd2(pk2(__uint_as_float(raw[4 * e + 2]), __uint_as_float(raw[4 * e + 3])), bi.y), x1b, xgb);
                const float ha = 0.5f * xga, hb = 0.5f * xgb;
                asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(ha));
                asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(hb));
                vv[hh * 8 + e] = pk2(x1a * fmaf(ha, ta, ha), x1b * fmaf(hb, tb, hb));
              }
            }
          } else {
          // software-pipelined TMEM reads: chunk i was requested one iteration ago (or before the loop); chunk i+1 is
          // requested as soon as the scale/bias FMAs have consumed the registers, so its latency hides behind the
          // activation / pack / store work of chunk i (the TMEM read port moves only 64 B/clk per SM)
          tmem_ld_wait();
          if (it + n_share >= n_items) {
            // this warp's last accumulator read of the tile is in registers: hand the TMEM buffer back NOW, so that the MMA
            // of the tile after next does not wait for the activation / pack / store work below
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(buf));
            released = true;
          }
          // packed fp32 math (FFMA2): the epilogue warps are issue-bound on the short-K convs
          {
            const ulonglong2* ss = reinterpret_cast<const ulonglong2*>(sscale + c0);
            const ulonglong2* sb = reinterpret_cast<const ulonglong2*>(sbias + cs * BN + c0);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const ulonglong2 sc = ss[e], bi = sb[e];
              vv[2 * e] = ffma2(pk2(__uint_as_float(raw[4 * e]), __uint_as_float(raw[4 * e + 1])), sc.x, bi.x);
              vv[2 * e + 1] = ffma2(pk2(__uint_as_float(raw[4 * e + 2]), __uint_as_float(raw[4 * e + 3])), sc.y, bi.y);
            }
          }
          if (it + n_share < n_items) {
            __syncwarp();
            tmem_ld32(t_item(it + n_share), raw);
          }
          if (act == ACT_SILU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float h0, h1, t0, t1;
              upk2(vv[e], h0, h1);
              if FTC_ABL(8192) { t0 = h0; t1 = h1; }      // ablation: no SFU
              else {
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
              }
              vv[e] = ffma2(vv[e], pk2(t0, t1), vv[e]);
            }
          } else if (act == ACT_GELU) {
            const f32x2 ca = pk2(7.97507884e-01f, 7.97507884e-01f), cb2 = pk2(3.70056461e-02f, 3.70056461e-02f);
            const f32x2 cc = pk2(-3.51516792e-04f, -3.51516792e-04f), half2 = pk2(0.5f, 0.5f);
#pragma unroll
            for (int e = 0; e < 16; ++e) {                 // gelu_tanh3 on pairs
              float q0, q1, t0, t1;
              upk2(fmul2(vv[e], vv[e]), q0, q1);
              const f32x2 x2 = pk2(fminf(q0, 64.f), fminf(q1, 64.f));
              const f32x2 u = fmul2(vv[e], ffma2(x2, ffma2(x2, cc, cb2), ca));
              upk2(u, q0, q1);
              asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(q0));
              asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(q1));
              const f32x2 h = fmul2(vv[e], half2);
              vv[e] = ffma2(h, pk2(t0, t1), h);
            }
          }
          }
          // residuals that cannot come through the TMA ring (row-periodic tables, a second residual): 64 bytes per lane
          auto add_direct = [&](const bf16* base, int64_t rrow, int stride) {
            const uint4* rp = reinterpret_cast<const uint4*>(base + rrow * stride + (int64_t)tc.g * p.N + tc.nt * BN + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = __ldg(rp + j);
              const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                vv[4 * j + e] = fadd2(vv[4 * j + e], pk2(__uint_as_float(uu[e] << 16), __uint_as_float(uu[e] & 0xFFFF0000u)));
            }
          };
          if (any_direct && m_row < p.M) {
            if (res1_direct) add_direct(res1, p.res1_row_mod ? (int64_t)(m_row % p.res1_row_mod) : (int64_t)m_row, p.res1_stride);
            if (has_res2) add_direct(res2, (int64_t)m_row, p.res2_stride);
          }
          if (res_tma) {
            const uint32_t slot = n_res_used & dmask;
            mbar_wait(res_bar(ewarp, slot), (n_res_used >> dshift) & 1u);
            const uint32_t rb_ = res_box0 + slot * TM_BOX_BYTES + row_b;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(rb_ + swo[j]) : "memory");
              const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
              for (int e = 0; e < 4; ++e)
                vv[4 * j + e] = fadd2(vv[4 * j + e], pk2(__uint_as_float(uu[e] << 16), __uint_as_float(uu[e] & 0xFFFF0000u)));
            }
            ++n_res_used;
            __syncwarp();                               // every lane has read the box before it is refilled
            if (lane == 0) issue_next_res();
          }
          if (!FTC_ABL(32)) {
            const uint32_t ob = out_box0 + (n_out & dmask) * TM_BOX_BYTES;
            if (lane == 0) {                            // the store that last read this box is done
              if (deep) bulk_wait_read<1>(); else bulk_wait_read<0>();
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 u;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int e = 0; e < 4; ++e) { float lo, hi; upk2(vv[4 * j + e], lo, hi); h[e] = __floats2bfloat162_rn(lo, hi); }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ob + row_b + swo[j]), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
            }
            if (!FTC_ABL(16384)) {                  // ablation bit: keep the st.shared, drop fence + TMA store
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_4d(&tmOut, ob, ch_out + (act == ACT_SWIGLU ? c0 >> 1 : c0), k1, k2, k3);
                bulk_commit();
              }
            }
            ++n_out;
          }
        }
        if (warp == 0 && lane == 0) FTC_TRACE(2, (int)titer, 2);
        if (!released) {                               // SwiGLU path, or a warp without items in this tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(buf));
        }
      }
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    } else
    for (int tile = t_first; tile < t_last; tile += t_step, ++titer) {
      const TmaTile tc = decode_tma_tile<MT, HALO>(tile, NT, G, L);
      const uint32_t buf = L.nbuf == 2 ? (titer & 1u) : 0u;
      const uint32_t use = L.nbuf == 2 ? (titer >> 1) : titer;
      // stage this tile's per-column scale / bias (the previous tile's readers are past this barrier)
      asm volatile("bar.sync 1, %0;" ::"n"(TM_EPI_THREADS) : "memory");
      {
        const int ncol0 = tc.nt * BN;
        for (int c = etid; c < BN; c += TM_EPI_THREADS) {
          const int n = ncol0 + c;
          sscale[c] = (p.scale && n < p.N) ? __ldg(p.scale + (int64_t)tc.g * p.N + n) : 1.f;
        }
        for (int c = etid; c < p.ncase * BN; c += TM_EPI_THREADS) {
          const int cs_ = c / BN, cc = c - cs_ * BN;
          const int n = ncol0 + cc;
          sbias[c] = (p.bias_tab && n < p.N) ? __ldg(p.bias_tab + ((int64_t)cs_ * G + tc.g) * p.N + n) : 0.f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TM_EPI_THREADS) : "memory");
      int m, b = 0, oy = 0, ox = 0;
      bool mvalid = true;
      if (HALO) {
        b = tc.b; oy = tc.y0 + (row >> 4); ox = tc.x0 + (row & 15);
        m = (b * p.Ho + oy) * p.Wo + ox;
        mvalid = ox < p.Wo;                      // ragged last tile column (W % 16 != 0)
      } else {
        m = tc.m0 + row;
        mvalid = m < p.M;
        if (mvalid) { b = m / hw; const int r = m - b * hw; oy = r / p.Wo; ox = r - oy * p.Wo; }
      }
      int cs = 0;
      if (p.ncase == 9) cs = (oy == 0 ? 0 : (oy == p.Ho - 1 ? 2 : 1)) * 3 + (ox == 0 ? 0 : (ox == p.Wo - 1 ? 2 : 1));
      const int64_t r1row = p.res1_row_mod ? (m % p.res1_row_mod) : m;
      const int nvalid = min(p.N, p.n_valid[tc.g]);
      const int chb = p.out_ch_base[tc.g];
      if (lane == 0) mbar_wait(tfull_bar(buf), use & 1u);
      __syncwarp();
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (buf * (uint32_t)MT + (MT == 2 ? (uint32_t)half : 0u)) * (uint32_t)BN;
      for (int c0 = (MT == 1 ? half * 16 : 0); c0 < BN; c0 += (MT == 1 ? 32 : 16)) {
        const int n0 = tc.nt * BN + c0;
        if (n0 >= p.N) break;                       // warp-uniform
        if FTC_ABL(256) continue;             // ablation: no TMEM reads, no epilogue math
        uint32_t raw16[16];
        __syncwarp();                               // tcgen05.ld is warp-collective: reconverge first
        tmem_ld16(t_addr + (uint32_t)c0, raw16);
        tmem_ld_wait();
        if (mvalid)
          epilogue_store(p, raw16, tc.g, n0, m, b, oy, ox, hw, r1row, nvalid, chb, sscale + c0, sbias + cs * BN + c0, res1, res2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TM_MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
int g_tma_sms = 0;

int encode_map(CUtensorMap* tm, const void* base, uint64_t channels, uint64_t pix_stride, uint64_t d1, uint64_t d2, uint64_t d3,
               uint32_t box1, uint32_t box2, uint32_t box0 = 64) {
  FTC_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA tensor must be 16-byte aligned");
  FTC_REQUIRE(pix_stride % 8 == 0, "TMA pixel stride must be a multiple of 16 bytes");
  cuuint64_t dims[4] = {channels, d1, d2, d3};
  cuuint64_t strides[3] = {pix_stride * 2, pix_stride * 2 * d1, pix_stride * 2 * d1 * d2};
  cuuint32_t box[4] = {box0, box1, box2, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // operand tiles: 64 channels = 128-byte rows, SWIZZLE_128B (the UMMA K-major layout); epilogue boxes: 32 channels =
  // 64-byte rows, SWIZZLE_64B
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, box0 == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return -2;
  }
  return 0;
}

}  // namespace

int conv_gemm_tma(const ConvGemmParams& p_in, cudaStream_t stream) {
  const int env_mt = gemm_tuning().mt, env_flags = gemm_tuning().flags;
  ConvGemmParams p = p_in;
  p.tc.flags = env_flags;
  p.trace = conv_gemm_trace_ptr();
  FTC_REQUIRE(p.dtype == DT_BF16, "tcgen05 path is bf16 only");
  FTC_REQUIRE(p.G >= 1 && p.G <= MAX_GROUPS, "groups out of range");
  FTC_REQUIRE(p.tc.BN >= 16 && p.tc.BN <= 256 && p.tc.BN % 16 == 0, "bad tc plan");
  FTC_REQUIRE(p.act != ACT_SWIGLU || p.out_layout == OUT_NHWC, "swiglu needs NHWC out");
  FTC_REQUIRE(p.stride == 1, "TMA paths are stride 1");
  const bool halo = p.tc.tma == TMA_HALO;
  const bool se = p.a_scale != nullptr;
  FTC_REQUIRE(!(se && halo), "SE operand scaling is a 1x1 feature");
  FTC_REQUIRE(p.tc.nGA + p.tc.nGB > 0, "TMA plan without sources");
  FTC_REQUIRE(p.tc.nGA == 0 || p.srcA, "source A missing");
  FTC_REQUIRE(p.tc.nGB == 0 || p.srcB, "source B missing");
  if (g_encode == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FTC_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    FTC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    g_encode = (EncodeTiledFn)fn;
    int dev = 0;
    FTC_CHECK_CUDA(cudaGetDevice(&dev));
    FTC_CHECK_CUDA(cudaDeviceGetAttribute(&g_tma_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  TmaLaunch L;
  memset(&L, 0, sizeof(L));
  // 256-row tiles halve the weight traffic per MAC.  They keep two accumulator sets (epilogue overlapped with the next
  // main loop) only for BN <= 128; wider tiles take them when the main loop is long enough to amortise the exposed
  // epilogue.  Either way there must be about a wave of tiles left.
  const long tiles2 = (long)ceil_div(p.M, 2 * TM_BM) * p.tc.NT * p.G;
  // (measured, tools/bench_gemm.py: with a single accumulator set the exposed epilogue costs more than the saved weight
  //  traffic up to ~30 k-blocks: 3x3 96->384 18 k-blocks 196 vs 224 us, 1x1 K=1536 84 vs 87 us; K=3072 69 vs 73 us the other way)
  int MT = ((p.tc.BN <= 128 || p.tc.NKB >= 32) && tiles2 >= 120) ? 2 : 1;
  if (env_mt == 1 || env_mt == 2) MT = env_mt;
  const bool kb32 = halo && p.tc.kb32 != 0;
  const uint32_t rowb = kb32 ? 64u : 128u;
  const uint32_t b_bytes = (uint32_t)p.tc.BN * rowb;
  if (!halo && !se && MT == 2 && p.tc.NKB <= TM_MAX_SLOTS && env_mt == 0) {
    // prefer the weight-stationary schedule (below) with 128-row tiles over 256-row tiles that cannot hold the weights
    const size_t fixed0 = 1024 + 8 * TM_NBARS + 16 + TM_STAGED_FLOATS * 4 + 256 + (size_t)TM_EPI_WARPS * (p.res1 ? 4 : 2) * TM_BOX_BYTES;
    const size_t avail0 = 227 * 1024 - fixed0, wbytes = (size_t)p.tc.NKB * b_bytes;
    if (wbytes + 3 * 2 * (size_t)TM_SUB_BYTES > avail0 && wbytes + 3 * (size_t)TM_SUB_BYTES <= avail0) MT = 1;
  }
  int m_tiles;
  if (halo) {
    if (p.H % 16 != 0) MT = 1;             // 8-row tiles; a ragged last tile COLUMN is fine: TMA zero-fills loads and clips stores
    FTC_REQUIRE(p.pad == 1 && p.H % (8 * MT) == 0 && p.Ho == p.H && p.Wo == p.W, "halo geometry");
    FTC_REQUIRE(p.tc.NKB == 9 * (p.tc.nGA + p.tc.nGB), "halo plan does not match K");
    L.tiles_x = ceil_div(p.W, HALO_TW); L.tiles_y = p.H / (8 * MT);
    L.NKG = 3 * (p.tc.nGA + p.tc.nGB); L.nsub = 3;
    L.a_slot_bytes = (uint32_t)(8 * MT + 2) * HALO_TW * rowb;
    m_tiles = p.B * L.tiles_x * L.tiles_y;
  } else {
    FTC_REQUIRE(p.pad == 0 && p.tc.nGB == 0 && p.tc.NKB == p.tc.nGA, "rows plan does not match K");
    L.tiles_x = L.tiles_y = 1;
    L.NKG = p.tc.nGA; L.nsub = 1;
    L.a_slot_bytes = (uint32_t)MT * TM_SUB_BYTES;
    m_tiles = ceil_div(p.M, TM_BM * MT);
  }
  L.nbuf = (2 * MT * p.tc.BN <= 512) ? 2 : 1;
  // SE scale table: two images x (channels padded to 64) bf16, when a tile cannot span more than two images
  if (se && p.Ho * p.Wo >= MT * TM_BM && p.tc.nGA * 64 <= 4096 && p.a_scale_stride % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(p.a_scale) & 15) == 0)
    L.se_tab = p.tc.nGA * 64;
  // staged (TMA-store) epilogue: plain bf16 NHWC outputs whose n-tiles are whole 32-channel boxes
  {
    const bool swiglu = p.act == ACT_SWIGLU;
    bool ok = p.out_layout == OUT_NHWC && p.tc.BN % (swiglu ? 64 : 32) == 0 && p.tc.BN * p.tc.NT == p.N &&
              (p.ncase == 1 || halo) && p.out_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 &&
              !(env_flags & 1024);
    for (int g = 0; g < p.G && ok; ++g) ok = p.n_valid[g] == p.N && p.out_ch_base[g] % 8 == 0;
    if (ok && swiglu) ok = p.res1 == nullptr && p.res2 == nullptr && p.scale == nullptr;
    if (ok && p.res1) ok = p.res1_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.res1) & 15) == 0 && (p.G * p.N) % 8 == 0;
    if (ok && p.res2) ok = !halo && p.res2_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(p.res2) & 15) == 0;
    if (ok && p.res1 && p.res1_row_mod) ok = !halo;          // periodic tables are read directly (rows mode only)
    L.staged = ok ? 1 : 0;
    L.res_tma = (ok && p.res1 && p.res1_row_mod == 0) ? 1 : 0;
  }
  // SE kernels need a third A slot (TMA -> scaler -> MMA hand-off) more than deep epilogue staging: their main loops are
  // long (K >= 768), so single boxes cost nothing there
  // (the weight-stationary kernels also take single boxes: the 2 KB per warp buy another activation slot, measured 90 vs 103 us)
  const bool want_bstat = !halo && !se && p.tc.NKB <= TM_MAX_SLOTS && p.tc.NKB <= 4;
  L.box_depth = (se || want_bstat) ? 1 : 2;
  if (gemm_tuning().box_depth == 1 || gemm_tuning().box_depth == 2) L.box_depth = gemm_tuning().box_depth;
  // without SE the four scaler warps would idle: they join the staged epilogue (3 instead of 2 warps per TMEM lane quarter;
  // the epilogue of the short-K convs is latency-bound with 2)
  L.n_epi = (L.staged && !se && !L.res_tma && !gemm_tuning().epi8) ? TM_MAX_EPI_WARPS : TM_EPI_WARPS;
  const size_t box_bytes = L.staged ? (size_t)L.n_epi * (L.res_tma ? 2 : 1) * L.box_depth * TM_BOX_BYTES : 0;
  const size_t fixed = 1024 + 8 * TM_NBARS + 16 + TM_STAGED_FLOATS * 4 + 256 + (size_t)L.se_tab * 4 + box_bytes;
  const size_t avail = 227 * 1024 - fixed;
  if (halo && !se && p.tc.NT * p.G == 1 && p.tc.NKB <= TM_MAX_SLOTS && !gemm_tuning().no_bstat &&
      (size_t)p.tc.NKB * b_bytes <= 48 * 1024 && (size_t)p.tc.NKB * b_bytes + 3 * (size_t)L.a_slot_bytes <= avail &&
      m_tiles >= 4 * g_tma_sms) {
    // weight-stationary 3x3 (stage 1: 32 -> 32 channels, 36 KB of weights): the tcgen05 trace showed the single producer
    // thread as the bottleneck of this conv - 12 TMA / bulk operations per 256-pixel tile at ~450 cycles each against
    // ~1.2k cycles of MMA.  With the nine weight blocks resident only the three halo boxes per tile are left.
    L.bstat = 1;
    L.nB = p.tc.NKB;
    size_t na = (avail - (size_t)L.nB * b_bytes) / L.a_slot_bytes;
    L.nA = na > TM_MAX_SLOTS ? TM_MAX_SLOTS : (int)na;
  } else if (halo) {
    L.nA = (3 * (size_t)L.a_slot_bytes + 4 * (size_t)b_bytes <= avail) ? 3 : 2;
    size_t nb = (avail - (size_t)L.nA * L.a_slot_bytes) / b_bytes;
    L.nB = nb > TM_MAX_SLOTS ? TM_MAX_SLOTS : (int)nb;
    FTC_REQUIRE(L.nB >= 3, "smem budget (halo)");
    // small-N convs (stage 1: 4 KB weight blocks) leave most of the shared memory unused: deepen the A ring so that more than
    // one tile's halo boxes are in flight (a tile needs 3 * (nGA + nGB) of them)
    while (L.nA < TM_MAX_SLOTS && (size_t)(L.nA + 1) * L.a_slot_bytes + (size_t)L.nB * b_bytes <= avail) ++L.nA;
  } else {
    // weight-stationary schedule (short K, many m-tiles per n-tile: the MBConv expand convs): the [BN x K] weight tile
    // stays in shared memory and only activations stream, which removes the dominant L2 -> SM traffic term
    // N*K*(M/tile rows) of these L2-bandwidth-bound GEMMs
    if (!se && p.tc.NKB <= TM_MAX_SLOTS && !(env_flags & 2048) && !gemm_tuning().no_bstat &&
        (size_t)p.tc.NKB * b_bytes + 3 * (size_t)L.a_slot_bytes <= avail && (long)m_tiles * p.tc.NT * p.G >= 4L * g_tma_sms) {
      L.bstat = 1;
      L.m_tiles = m_tiles;
      L.nB = p.tc.NKB;
      size_t na = (avail - (size_t)L.nB * b_bytes) / L.a_slot_bytes;
      L.nA = na > TM_MAX_SLOTS ? TM_MAX_SLOTS : (int)na;
    } else {
    size_t n = avail / ((size_t)L.a_slot_bytes + b_bytes);
    L.nA = L.nB = n > TM_MAX_SLOTS ? TM_MAX_SLOTS : (int)n;
    FTC_REQUIRE(L.nA >= 2, "smem budget (rows)");

    // spend what is left on one more slot of either ring (A first: with SE it feeds a three-party hand-off)
    size_t left = avail - (size_t)L.nA * ((size_t)L.a_slot_bytes + b_bytes);
    if (L.nA < TM_MAX_SLOTS && left >= L.a_slot_bytes) { ++L.nA; left -= L.a_slot_bytes; }
    if (L.nB < TM_MAX_SLOTS && left >= b_bytes) { ++L.nB; left -= b_bytes; }
    if (gemm_tuning().nb >= 2 && gemm_tuning().nb <= L.nB) {       // tuning: shallower weight ring, deeper activation ring
      L.nB = gemm_tuning().nb;
      size_t na = (avail - (size_t)L.nB * b_bytes) / L.a_slot_bytes;
      L.nA = na > TM_MAX_SLOTS ? TM_MAX_SLOTS : (int)na;
    }
    }
  }
  size_t smem = fixed + (size_t)L.nA * L.a_slot_bytes + (size_t)L.nB * b_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: the CTA owns all 512 TMEM columns
  FTC_REQUIRE(smem <= 227 * 1024, "smem budget");

  alignas(64) CUtensorMap tmA, tmB;
  const bf16* a_base = reinterpret_cast<const bf16*>(p.srcA);
  const bf16* b_base = reinterpret_cast<const bf16*>(p.srcB);
  int rc = 0;
  if (halo) {
    if (p.tc.nGA) rc = encode_map(&tmA, a_base, (uint64_t)(p.a_ch_off + p.CA), p.a_pix_stride, p.W, p.H, p.B, HALO_TW, 8 * MT + 2, kb32 ? 32 : 64);
    if (rc) return rc;
    if (p.tc.nGB) rc = encode_map(&tmB, b_base, (uint64_t)p.b_pix_stride, p.b_pix_stride, p.W, p.H, p.B, HALO_TW, 8 * MT + 2);
    if (rc) return rc;
    if (!p.tc.nGA) tmA = tmB;
    if (!p.tc.nGB) tmB = tmA;
  } else {
    rc = encode_map(&tmA, a_base, (uint64_t)(p.a_ch_off + p.CA), p.a_pix_stride, (uint64_t)p.M, 1, 1, TM_BM * MT, 1);
    if (rc) return rc;
    tmB = tmA;
  }
  alignas(64) CUtensorMap tmOut = tmA, tmRes = tmA;
  if (L.staged) {
    if (halo) rc = encode_map(&tmOut, p.out, (uint64_t)p.out_stride, p.out_stride, p.Wo, p.Ho, p.B, HALO_TW, 2, 32);
    else rc = encode_map(&tmOut, p.out, (uint64_t)p.out_stride, p.out_stride, (uint64_t)p.M, 1, 1, 32, 1, 32);
    if (rc) return rc;
    if (L.res_tma) {
      if (halo) rc = encode_map(&tmRes, p.res1, (uint64_t)p.res1_stride, p.res1_stride, p.Wo, p.Ho, p.B, HALO_TW, 2, 32);
      else rc = encode_map(&tmRes, p.res1, (uint64_t)p.res1_stride, p.res1_stride, (uint64_t)p.M, 1, 1, 32, 1, 32);
      if (rc) return rc;
    }
  }
  const int num_tiles = m_tiles * p.tc.NT * p.G;
  FTC_REQUIRE(num_tiles < (1 << 23) && m_tiles < (1 << 17), "tile count exceeds the multiply-shift division range");
  L.m_tiles = m_tiles;
  L.mg_per_m = div_magic(p.tc.NT * p.G); L.mg_nt = div_magic(p.tc.NT); L.mg_per_img = div_magic(L.tiles_x * L.tiles_y);
  L.mg_tiles_x = div_magic(L.tiles_x); L.mg_m_tiles = div_magic(m_tiles);
  const int grid = num_tiles < g_tma_sms ? num_tiles : g_tma_sms;
  p.tc.MT = MT;
#define TMA_LAUNCH(SE_, MT_, HALO_)                                                                                   \
  do {                                                                                                                \
    static bool attr_done = false;                                                                                    \
    if (!attr_done) {                                                                                                 \
      FTC_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tma_kernel<SE_, MT_, HALO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          227 * 1024));                                                               \
      attr_done = true;                                                                                               \
    }                                                                                                                 \
    FTC_CHECK_CUDA(launch_pdl(conv_gemm_tma_kernel<SE_, MT_, HALO_>, dim3(grid), dim3(TM_THREADS), smem, stream, p, tmA, tmB, tmOut, tmRes, L, num_tiles)); \
  } while (0)
  if (halo) { if (MT == 2) TMA_LAUNCH(false, 2, true); else TMA_LAUNCH(false, 1, true); }
  else if (se) { if (MT == 2) TMA_LAUNCH(true, 2, false); else TMA_LAUNCH(true, 1, false); }
  else { if (MT == 2) TMA_LAUNCH(false, 2, false); else TMA_LAUNCH(false, 1, false); }
#undef TMA_LAUNCH
  FTC_POST_LAUNCH();
  return 0;
}

}  // namespace ftc
