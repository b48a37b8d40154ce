// tcgen05 / TMEM implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate in tensor memory).
//
// One persistent CTA per SM, 9 warps, three roles connected by mbarrier pipelines:
//   warps 0-3  A producers : gather the im2col rows of a 128-pixel tile straight from the NHWC activation
//                            tensor(s) with 16-byte cp.async into the 128B-swizzled K-major smem layout that
//                            tcgen05.mma reads (zero fill for padding / ragged tails; the Leafmap "concat" is just
//                            a second source pointer in the k-chunk table).  Thread 0 also issues the stage's
//                            weight block as ONE cp.async.bulk (TMA bulk copy): weights are pre-packed in
//                            exactly the smem image (pack_conv_weight_tc).
//   warp  4    MMA issuer  : one thread, 4 x tcgen05.mma (M=128, N=BN, K=16) per 64-wide k-block, accumulating
//                            in one of two TMEM accumulators; tcgen05.commit releases smem stages / publishes
//                            the accumulator.
//   warps 5-8  epilogue    : tcgen05.ld the accumulator (32 lanes per warp), folded-BN scale + (border-aware)
//                            bias + residual(s) + SiLU / erf-GELU / SwiGLU, store NHWC bf16 or NCHW fp32.
// The epilogue of tile i overlaps the main loop of tile i+1 (double-buffered TMEM).
#include "tc_common.cuh"

namespace ftc {

namespace {

constexpr int TC_BM = 128;
// warp roles: 0-7 epilogue, 8 MMA issuer, 9-16 producers (producers get the HIGHEST warp ids: the SM's issue arbiter
// favours high warp ids, and the producers' address-generation chain is the critical path of the pipeline)
constexpr int PROD_WARPS = 8;
constexpr int MMA_WARP = 8, PROD_WARP0 = 9;
constexpr int TC_THREADS = (8 + 1 + PROD_WARPS) * 32;   // 544
constexpr int TC_MAX_KTAB = 1536;         // k-chunk table entries staged in shared memory (K <= 12288)
constexpr uint32_t A_STAGE_BYTES = TC_BM * 128;
constexpr int TC_MAX_STAGES = 6;



struct TileCoord { int m0, g, nt; };
__device__ __forceinline__ TileCoord decode_tile(int tile, int NT, int G, int rows_per_tile) {
  int per_m = NT * G;
  int mt = tile / per_m;
  int rest = tile - mt * per_m;
  TileCoord t;
  t.m0 = mt * rows_per_tile;
  t.g = rest / NT;
  t.nt = rest - t.g * NT;
  return t;
}


constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int STAGED_FLOATS = 10 * 256;   // per-tile scale[BN] + bias[ncase<=9][BN]

// MT = number of 128-row sub-tiles per CTA tile (1: 128 x BN with double-buffered TMEM accumulators so the epilogue
// overlaps the next main loop; 2: 256 x BN, two accumulators sharing every B stage -> 30 % fewer smem bytes per MMA
// cycle, used when K is long enough that the exposed epilogue is small)
template <bool SE, int MT>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_gemm_tc_kernel(const __grid_constant__ ConvGemmParams p,
                                                                      const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sbase = (raw + 1023u) & ~1023u;          // SWIZZLE_128B needs 1024 B alignment
  uint8_t* smem = smem_raw + (sbase - raw);
  const int S = p.tc.stages, BN = p.tc.BN, NKB = p.tc.NKB, NT = p.tc.NT;
  const uint32_t b_bytes = (uint32_t)BN * 128u;
  constexpr uint32_t A_BYTES = (uint32_t)MT * A_STAGE_BYTES;
  const uint32_t stage_bytes = A_BYTES + b_bytes;
  const uint32_t bar0 = sbase + (uint32_t)S * stage_bytes;   // 8 B each: full[S], empty[S], tfull[2], tempty[2]
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * S + 2 + a); };
  uint8_t* after_bars = smem + (size_t)S * stage_bytes + 8 * (2 * S + 4);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(after_bars);
  float* staged = reinterpret_cast<float*>(after_bars + 16);      // [STAGED_FLOATS], 16 B aligned
  uint32_t* sktab = reinterpret_cast<uint32_t*>(staged + STAGED_FLOATS);   // [NKB*8] k-chunk table copy
  for (int i = threadIdx.x; i < NKB * 8; i += TC_THREADS) sktab[i] = p.ktab[i];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == MMA_WARP) {
    if (lane == 0) {
      // full: the B expect_tx arrive + either 256 async per-thread arrivals or (SE) one arrival per producer warp
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1 + (SE ? PROD_WARPS : PROD_WARPS * 32)); mbar_init(empty_bar(s), 1); }
      for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
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
  pdl_wait();
  const int G = p.G;
  const int hw = p.Ho * p.Wo;
  const uint32_t hint_ns = (p.tc.flags & 1) ? 100000u : 0u;          // experiment knobs (FTC_TC_FLAGS)
  const uint32_t epi_backoff_ns = (p.tc.flags & 2) ? 200u : 0u;

  if (warp >= PROD_WARP0) {
    // ------------------------------------------------------------------ A producers (+ B bulk copy)
    // All 8 producer warps cooperate on every k-block: thread t owns chunk j = t & 7 of rows (t >> 3) + 32*i.
    // Plain convs publish through cp.async.mbarrier.arrive.noinc (the barrier receives the thread's arrival when its
    // copies land), so producers only ever block on the stage-empty barrier.  SE convs keep `lag` k-blocks in flight,
    // then scale the landed operand in place and publish.
    constexpr int ROWS = 4 * MT;                   // rows per thread
    const int t = threadIdx.x - PROD_WARP0 * 32;   // 0..255
    const int j = t & 7;
    const int rbase = t >> 3;                      // 0..31
    const uint32_t row_off = (uint32_t)(rbase >> 3) * 1024u + (uint32_t)(rbase & 7) * 128u + (uint32_t)((j ^ (rbase & 7)) << 4);
    auto dst_off = [&](int i) { return (uint32_t)(i >> 2) * A_STAGE_BYTES + row_off + (uint32_t)(i & 3) * 4096u; };
    const bf16* srcA = reinterpret_cast<const bf16*>(p.srcA);
    const bf16* srcB = reinterpret_cast<const bf16*>(p.srcB);
    const bf16* wgt = reinterpret_cast<const bf16*>(p.w);
    const bf16* any_src = srcA ? srcA : srcB;      // dereferenceable address for zero-byte (fill-only) copies
    int stage = 0;
    uint32_t phase = 0;
    int arr_stage = 0;               // SE: oldest stage whose A fill this thread still has to publish
    int arr_kb = 0;                  //     its k-block index inside the current tile (the scale needs the channel)
    uint32_t pending = 0;            //     committed-but-unpublished cp.async groups
    int img[ROWS];
    int trace_n = 0;
    const int lag = S - 2 > 2 ? 2 : (S - 2 < 1 ? 1 : S - 2);

    auto publish = [&]() {           // SE only: the oldest group has landed (caller waited)
      const uint32_t e = sktab[arr_kb * 8 + j];
      if (e & KT_VALID) {
        const int c = kt_c(e);
        const uint32_t a_dst = sbase + (uint32_t)arr_stage * stage_bytes;
#pragma unroll
        for (int i = 0; i < ROWS; ++i) {
          const uint32_t addr = a_dst + dst_off(i);
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
      if (lane == 0) mbar_arrive(full_bar(arr_stage));
      arr_stage = (arr_stage + 1 == S) ? 0 : arr_stage + 1;
      ++arr_kb;
      --pending;
    };

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(tile, NT, G, MT * TC_BM);
      // per row: pixel offset of tap (0,0) and a 9-bit mask of the taps that fall inside the image, so the k-loop
      // spends ~5 instructions per 16-byte gather (test bit, 2 selects, one 64-bit multiply-add, cp.async)
      int pixoff[ROWS];
      uint32_t vmask[ROWS];
#pragma unroll
      for (int i = 0; i < ROWS; ++i) {
        const int m = tc.m0 + (i >> 2) * TC_BM + rbase + 32 * (i & 3);
        pixoff[i] = 0; vmask[i] = 0u; img[i] = 0;
        if (m < p.M) {
          const int b = m / hw;
          const int r = m - b * hw;
          const int oy = r / p.Wo, ox = r - oy * p.Wo;
          const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
          pixoff[i] = (b * p.H + iy0) * p.W + ix0;
          img[i] = b;
          uint32_t ym = 0, xm = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            ym |= (iy0 + k >= 0 && iy0 + k < p.H) ? (1u << k) : 0u;
            xm |= (ix0 + k >= 0 && ix0 + k < p.W) ? (1u << k) : 0u;
          }
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
            if (ym & (1u << ky)) vmask[i] |= xm << (3 * ky);
        }
      }
      const bf16* wtile = wgt + (size_t)(tc.g * NT + tc.nt) * NKB * ((size_t)BN * 64);
      if (SE) arr_kb = 0;            // SE drains at tile end, so pending == 0 here
      for (int kb = 0; kb < NKB; ++kb) {
        const uint32_t e = sktab[kb * 8 + j];
        const bool srcb = (e & KT_SRCB) != 0;
        const int ky = kt_ky(e), kx = kt_kx(e);
        const int pstride = srcb ? p.b_pix_stride : p.a_pix_stride;
        const uint32_t tapbit = (e & KT_VALID) ? (1u << (ky * 3 + kx)) : 0u;
        // address of this chunk for a row whose tap-(0,0) pixel offset is 0
        const bf16* kbase = tapbit == 0u ? any_src
                                         : (srcb ? (srcB + p.b_ch_off + tc.g * p.b_group_stride) : (srcA + p.a_ch_off)) + kt_c(e) +
                                               (int64_t)(ky * p.W + kx) * pstride;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && t == 0 && trace_n < 1024;
        if (tr) p.trace[2048 + trace_n] = clock64();
        if (lane == 0) mbar_wait(empty_bar(stage), phase ^ 1u, 0, hint_ns);
        if (tr) p.trace[3072 + trace_n++] = clock64();
        __syncwarp();
        const uint32_t a_dst = sbase + (uint32_t)stage * stage_bytes;
        if (t == 0) {
          if (p.tc.flags & 4) mbar_arrive(full_bar(stage));       // experiment: no B traffic
          else {
            mbar_arrive_expect_tx(full_bar(stage), b_bytes);
            bulk_copy_g2s(a_dst + A_BYTES, wtile + (size_t)kb * BN * 64, b_bytes, full_bar(stage));
          }
        }
        if (!(p.tc.flags & 8)) {                                   // experiment: no A traffic
#pragma unroll
          for (int i = 0; i < ROWS; ++i) {
            const bool ok = (vmask[i] & tapbit) != 0u;
            const bf16* src = kbase + (int64_t)(ok ? pixoff[i] : 0) * pstride;
            cp_async16(a_dst + dst_off(i), src, ok ? 16u : 0u);
          }
        }
        if (!SE) {
          cp_async_mbar_arrive_noinc(full_bar(stage));   // never blocks: only the empty barrier paces the producers
        } else {
          cp_async_commit();
          ++pending;
          if (pending > (uint32_t)lag) {     // keep `lag` k-blocks of gathers in flight, publish the oldest
            if (lag == 1) cp_async_wait<1>(); else cp_async_wait<2>();
            publish();
          }
        }
        stage = (stage + 1 == S) ? 0 : stage + 1;
        phase ^= (stage == 0) ? 1u : 0u;
      }
      if (SE) {      // the scale rows (img[]) belong to this tile: publish everything before moving on
        cp_async_wait<0>();
        while (pending > 0) publish();
      }
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer (whole warp walks the loop, one elected
    // lane issues: descriptor arithmetic stays warp-uniform, see conv_gemm_tma.cu)
    const bool leader = elect_one_sync();
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0;
      int trace_n = 0;
      uint32_t phase = 0;
      uint32_t titer = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
        // MT == 1: two accumulators alternate (epilogue overlaps the next main loop); MT == 2: both belong to this tile
        const uint32_t acc = MT == 1 ? (titer & 1u) : 0u;
        const uint32_t use = MT == 1 ? (titer >> 1) : titer;
        mbar_wait(tempty_bar(acc), (use & 1u) ^ 1u, 0, hint_ns);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)BN;
        for (int kb = 0; kb < NKB; ++kb) {
          const bool tr = leader && p.trace != nullptr && blockIdx.x == 0 && trace_n < 1024;
          if (tr) p.trace[trace_n] = clock64();
          mbar_wait(full_bar(stage), phase, 0, hint_ns);
          if (tr) p.trace[1024 + trace_n++] = clock64();
          tc_fence_after();
          const uint32_t a_addr = sbase + (uint32_t)stage * stage_bytes;
          const uint64_t bdesc = umma_desc_sw128(a_addr + A_BYTES);
          if (leader && !(p.tc.flags & 16))                         // (experiment bit: no MMA)
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 4 x (K = 16 bf16 = 32 B) inside the 128 B swizzle row: +2 in 16-byte units
#pragma unroll
            for (int h = 0; h < MT; ++h)
              umma_f16(d_tmem + (uint32_t)h * (uint32_t)BN, umma_desc_sw128(a_addr + (uint32_t)h * A_STAGE_BYTES) + (uint64_t)(2 * k),
                       bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (leader) umma_commit(empty_bar(stage));
          stage = (stage + 1 == S) ? 0 : stage + 1;
          phase ^= (stage == 0) ? 1u : 0u;
        }
        if (leader) umma_commit(tfull_bar(acc));
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps: 4 lane quarters x 2 column halves)
    const int ew = warp;                       // 0..7
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int half = ew >> 2;                  // MT == 1: which 16-column chunks (even / odd); MT == 2: which 128-row sub-tile
    const int etid = threadIdx.x;              // 0..255
    const int row = (MT == 2 ? half * TC_BM : 0) + q * 32 + lane;
    const bf16* res1 = reinterpret_cast<const bf16*>(p.res1);
    const bf16* res2 = reinterpret_cast<const bf16*>(p.res2);
    float* sscale = staged;                    // [BN]
    float* sbias = staged + 256;               // [ncase][BN]
    uint32_t titer = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++titer) {
      const TileCoord tc = decode_tile(tile, NT, G, MT * TC_BM);
      const uint32_t acc = MT == 1 ? (titer & 1u) : 0u;
      const uint32_t use = MT == 1 ? (titer >> 1) : titer;
      // stage this tile's per-column scale / bias (the previous tile's readers are past this barrier)
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      {
        const int ncol0 = tc.nt * BN;
        for (int c = etid; c < BN; c += EPI_THREADS) {
          const int n = ncol0 + c;
          sscale[c] = (p.scale && n < p.N) ? __ldg(p.scale + (int64_t)tc.g * p.N + n) : 1.f;
        }
        for (int c = etid; c < p.ncase * BN; c += EPI_THREADS) {
          const int cs_ = c / BN, cc = c - cs_ * BN;
          const int n = ncol0 + cc;
          sbias[c] = (p.bias_tab && n < p.N) ? __ldg(p.bias_tab + ((int64_t)cs_ * G + tc.g) * p.N + n) : 0.f;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
      if (lane == 0) mbar_wait(tfull_bar(acc), use & 1u, epi_backoff_ns, hint_ns);
      __syncwarp();
      tc_fence_after();
      const int m = tc.m0 + row;
      const bool mvalid = m < p.M;
      int b = 0, oy = 0, ox = 0;
      if (mvalid) { b = m / hw; int r = m - b * hw; oy = r / p.Wo; ox = r - oy * p.Wo; }
      int cs = 0;
      if (p.ncase == 9) cs = (oy == 0 ? 0 : (oy == p.Ho - 1 ? 2 : 1)) * 3 + (ox == 0 ? 0 : (ox == p.Wo - 1 ? 2 : 1));
      const int64_t r1row = p.res1_row_mod ? (m % p.res1_row_mod) : m;
      const int nvalid = min(p.N, p.n_valid[tc.g]);
      const int chb = p.out_ch_base[tc.g];
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (MT == 1 ? acc : (uint32_t)half) * (uint32_t)BN;
      for (int c0 = (MT == 1 ? half * 16 : 0); c0 < BN; c0 += (MT == 1 ? 32 : 16)) {
        const int n0 = tc.nt * BN + c0;
        if (n0 >= p.N) break;                       // warp-uniform
        uint32_t raw16[16];
        __syncwarp();                               // tcgen05.ld is warp-collective: reconverge first
        tmem_ld16(t_addr + (uint32_t)c0, raw16);
        tmem_ld_wait();
        if (mvalid && !(p.tc.flags & 32))                            // experiment: no epilogue math/stores
          epilogue_store(p, raw16, tc.g, n0, m, b, oy, ox, hw, r1row, nvalid, chb, sscale + c0, sbias + cs * BN + c0, res1, res2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// dgrad_rows > 0: src is the FORWARD convolution's OIHW weight [dgrad_rows][O][kh][kw]; the packed operand is the data gradient's
// (channel roles swapped, taps rotated 180 degrees): element (o, c, ky, kx) = src[c][o][kh-1-ky][kw-1-kx], zero for c >= dgrad_rows
// (dy channels padded up to the kernel's granularity).  Saves the flip / transpose / contiguous copies per convolution and step.
__global__ void pack_conv_weight_tc_kernel(bf16* __restrict__ dst, const float* __restrict__ src, int O, int Itot, int kh,
                                           int kw, int c_off, int C, int k_off, int NKB, int o_off, int BN,
                                           const float* __restrict__ cscale, int halo_order, int dgrad_rows) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t per_o = (int64_t)kh * kw * C;
  if (idx >= (int64_t)O * per_o) return;
  int o = idx / per_o;
  int r = idx - (int64_t)o * per_o;
  int t = r / C, c = r - t * C;
  int ky = t / kw, kx = t - ky * kw;
  float v;
  if (dgrad_rows > 0) v = c < dgrad_rows ? src[(((int64_t)c * O + o) * kh + (kh - 1 - ky)) * kw + (kw - 1 - kx)] : 0.f;
  else v = src[(((int64_t)o * Itot + c_off + c) * kh + ky) * kw + kx];
  if (cscale) v *= cscale[c];
  int R = o_off + o;
  int tile = R / BN, rr = R - tile * BN;
  if (halo_order == 2) {
    // 32-wide k-blocks, 64-byte rows, SWIZZLE_64B: 16-byte chunk ^= (row >> 1) & 3, 8-row groups of 512 B
    const int k = (kx * 3 + ky) * 32 + c;
    const int kb = k >> 5, kk = k & 31, chunk = kk >> 3, within = kk & 7;
    const int64_t off = ((int64_t)tile * NKB + kb) * ((int64_t)BN * 32) + (rr >> 3) * 256 + (rr & 7) * 32 + ((chunk ^ ((rr >> 1) & 3)) << 3) + within;
    dst[off] = __float2bfloat16_rn(v);
    return;
  }
  int k = halo_order ? k_off + ((c >> 6) * 9 + kx * 3 + ky) * 64 + (c & 63) : k_off + r;
  int kb = k >> 6, kk = k & 63;
  int chunk = kk >> 3, within = kk & 7;
  int64_t off = ((int64_t)tile * NKB + kb) * ((int64_t)BN * 64) + (rr >> 3) * 512 + (rr & 7) * 64 + ((chunk ^ (rr & 7)) << 3) + within;
  dst[off] = __float2bfloat16_rn(v);
}

int g_num_sms = 0;
unsigned long long* g_trace = nullptr;

}  // namespace

static int tc_stages(int MT, int BN) {
  int stage_bytes = MT * (int)A_STAGE_BYTES + BN * 128;
  int s = (208 * 1024) / stage_bytes;
  return s > TC_MAX_STAGES ? TC_MAX_STAGES : (s < 2 ? 2 : s);
}

int conv_gemm_tc_plan(const ConvGemmParams& p, ConvTcPlan* plan, bool allow_tma) {
  static int env_no_tma = -1;
  if (env_no_tma < 0) { const char* e = getenv("FTC_NO_TMA"); env_no_tma = e ? atoi(e) : 0; }   // 1: all, 2: rows, 4: halo
  FTC_REQUIRE(p.K % KBLOCK == 0 && p.K > 0, "K must be a positive multiple of 64");
  FTC_REQUIRE(p.N >= 1, "N");
  int best_bn = 0, best_nt = 0;
  long best_pad = -1;
  int nt_min = (p.N + 255) / 256;
  for (int nt = nt_min; nt <= nt_min + 3; ++nt) {
    int bn = ((p.N + nt - 1) / nt + 15) / 16 * 16;
    if (bn > 256) continue;
    long pad = (long)bn * nt;
    if (best_pad < 0 || pad < best_pad) { best_pad = pad; best_bn = bn; best_nt = nt; }
  }
  FTC_REQUIRE(best_bn > 0, "no N tiling");
  if (gemm_tuning().plan_bn > 0 && p.N % gemm_tuning().plan_bn == 0) { best_bn = gemm_tuning().plan_bn; best_nt = p.N / best_bn; }
  plan->BN = best_bn;
  plan->NT = best_nt;
  plan->NKB = p.K / KBLOCK;
  plan->flags = 0;
  plan->MT = 1;
  plan->stages = tc_stages(1, best_bn);
  plan->tma = TMA_NONE; plan->nGA = plan->nGB = 0; plan->kb32 = 0;
  if (allow_tma && !(env_no_tma & 1) && (p.CA + p.CB) > 0 && p.stride == 1) {
    const bool strides_ok = (p.CA == 0 || p.a_pix_stride % 8 == 0) && (p.CB == 0 || (p.b_pix_stride % 8 == 0 && p.b_group_stride % 8 == 0));
    const int nGA = (p.CA + 63) / 64, nGB = (p.CB + 63) / 64;
    // a chunk may only run past its source's channels when the tensor ENDS there (TMA zero-fills out of bounds):
    // source B slices of a wider tensor must be whole chunks
    const bool chunks_ok = (p.CB % 64 == 0 || p.G == 1) ;
    if (p.pad == 0 && p.CB == 0 && strides_ok && !(env_no_tma & 2)) {
      plan->tma = TMA_ROWS; plan->nGA = nGA; plan->nGB = 0; plan->NKB = nGA;
      // K <= 256 with many n-tiles (MBConv expand 192->768, 256->1536): 192-wide n-tiles keep the resident weight tile at
      // <= 96 KB under the weight-stationary schedule, which leaves room for a deeper activation ring (measured 89 vs 103 us)
      if (nGA <= 4 && p.N % 192 == 0 && p.N >= 768 && plan->BN > 192 && gemm_tuning().plan_bn == 0) { plan->BN = 192; plan->NT = p.N / 192; }
      // (K in (256, 512] with 128-wide n-tiles so that [128 x K] stays resident was measured slower: 0.127 vs 0.078 ms at
      // K = 512, N = 3072, M = 18432; removed)
    } else if (p.pad == 1 && strides_ok && chunks_ok && p.H % 8 == 0 && !(env_no_tma & 4)) {
      plan->tma = TMA_HALO; plan->nGA = nGA; plan->nGB = nGB; plan->NKB = 9 * (nGA + nGB);
      // <= 32 input channels (stage 1): 32-wide k-blocks; padding to 64 doubled the MMA count and every tcgen05.mma streams
      // its whole A operand from shared memory whatever N is (65 cycles each at N = 32: tools/trace_tma.py)
      if (p.CB == 0 && p.CA == 32 && !getenv("FTC_NO_KB32")) plan->kb32 = 1;
    }
  }
  return 0;
}

void conv_gemm_tc_set_trace(unsigned long long* dev_ptr) { g_trace = dev_ptr; }
unsigned long long* conv_gemm_trace_ptr() { return g_trace; }

GemmTuning& gemm_tuning() {
  static GemmTuning t = [] {
    GemmTuning v{0, 0, 0, 0, 0, 0, 0};
    const char* e = getenv("FTC_TMA_MT"); v.mt = e ? atoi(e) : 0;
    e = getenv("FTC_TMA_BOX_DEPTH"); v.box_depth = e ? atoi(e) : 0;
    e = getenv("FTC_TMA_EPI8"); v.epi8 = e ? atoi(e) : 0;
    e = getenv("FTC_TMA_FLAGS"); v.flags = e ? atoi(e) : 0;     // ablations (results are garbage): 32 no stores, 64 no SE
    return v;                                                   // scaling, 128 no residual loads, 256 no epilogue at all
  }();
  return t;
}

size_t conv_tc_weight_bytes(const ConvTcPlan& plan, int G) {
  return (size_t)G * plan.NT * plan.NKB * plan.BN * 128;
}

int pack_conv_weight_tc(void* dst, const float* src, int O, int Itot, int kh, int kw, int c_off, int C, int k_off, int Kpad,
                        int o_off, int BN, const float* cscale, cudaStream_t s, int halo_order, int dgrad_rows) {
  int64_t total = (int64_t)O * kh * kw * C;
  int grid = (int)((total + 255) / 256);
  FTC_REQUIRE(!halo_order || (kh == 3 && kw == 3), "halo order is for 3x3 kernels");
  pack_conv_weight_tc_kernel<<<grid, 256, 0, s>>>((bf16*)dst, src, O, Itot, kh, kw, c_off, C, k_off, Kpad / KBLOCK, o_off, BN,
                                                  cscale, halo_order, dgrad_rows);
  FTC_POST_LAUNCH();
  return 0;
}

int conv_gemm_tc(const ConvGemmParams& p_in, cudaStream_t stream) {
  if (p_in.tc.tma != TMA_NONE) return conv_gemm_tma(p_in, stream);
  static int env_flags = -1;
  if (env_flags < 0) { const char* e = getenv("FTC_TC_FLAGS"); env_flags = e ? atoi(e) : 0; }
  ConvGemmParams p = p_in;
  p.tc.flags = env_flags;
  p.trace = g_trace;
  FTC_REQUIRE(p.dtype == DT_BF16, "tcgen05 path is bf16 only");
  FTC_REQUIRE(p.G >= 1 && p.G <= MAX_GROUPS, "groups out of range");
  FTC_REQUIRE(p.tc.BN >= 16 && p.tc.BN <= 256 && p.tc.BN % 16 == 0, "bad tc plan");
  FTC_REQUIRE(p.tc.NKB * KBLOCK == p.K, "tc plan does not match K");
  FTC_REQUIRE(p.H < 65000 && p.W < 65000, "spatial size");
  FTC_REQUIRE(p.act != ACT_SWIGLU || p.out_layout == OUT_NHWC, "swiglu needs NHWC out");
  if (g_num_sms == 0) {
    int dev = 0;
    FTC_CHECK_CUDA(cudaGetDevice(&dev));
    FTC_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const bool se = p.a_scale != nullptr;
  // tile height is a launch-time choice (the weight image only depends on BN): 256-row tiles when the main loop is
  // long enough to amortise the non-overlapped epilogue and there are enough rows to fill the machine
  p.tc.MT = (p.tc.NKB >= 12 && p.M >= 2 * TC_BM * 128 && !(p.tc.flags & 512)) ? 2 : 1;
  p.tc.stages = tc_stages(p.tc.MT, p.tc.BN);
  const int m_tiles = ceil_div(p.M, TC_BM * p.tc.MT);
  const int num_tiles = m_tiles * p.tc.NT * p.G;
  size_t smem = (size_t)p.tc.stages * ((size_t)p.tc.MT * A_STAGE_BYTES + (size_t)p.tc.BN * 128) + 1024 + 256 + STAGED_FLOATS * 4 + TC_MAX_KTAB * 4;
  FTC_REQUIRE(p.tc.NKB * 8 <= TC_MAX_KTAB, "K too large for the shared-memory chunk table");
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: the CTA owns all 512 TMEM columns
  FTC_REQUIRE(smem <= 227 * 1024, "smem budget");
  int grid = num_tiles < g_num_sms ? num_tiles : g_num_sms;
#define TC_LAUNCH(SE_, MT_)                                                                                        \
  do {                                                                                                             \
    static bool attr_done = false;                                                                                 \
    if (!attr_done) {                                                                                              \
      FTC_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_tc_kernel<SE_, MT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
      attr_done = true;                                                                                            \
    }                                                                                                              \
    FTC_CHECK_CUDA(launch_pdl(conv_gemm_tc_kernel<SE_, MT_>, dim3(grid), dim3(TC_THREADS), smem, stream, p, num_tiles)); \
  } while (0)
  if (se) { if (p.tc.MT == 2) TC_LAUNCH(true, 2); else TC_LAUNCH(true, 1); }
  else { if (p.tc.MT == 2) TC_LAUNCH(false, 2); else TC_LAUNCH(false, 1); }
#undef TC_LAUNCH
  FTC_POST_LAUNCH();
  return 0;
}

}  // namespace ftc
