// Detector engine: static plan of the EfficientNetV2 + 9xLeafmap forward over the ops in
// conv_gemm*.cu / detector_ops.cu.  The plan is built once from the stage table; weights are folded/packed
// on the device from the reference state_dict tensors; forward() only fills pointers and launches.
//
// Reference call sequence being replaced: models/detector.py:217-230 (CenterNetDetection.forward),
// :139-146 (BackboneModel.forward), :192-201 (Leafmap.forward), :289-296 (CenterNetDetector.forward).
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/ftc_b200.h"
#include "conv_gemm.cuh"
#include "detector_ops.cuh"

namespace ftc {

namespace {

enum Buf : int { BUF_X0 = 0, BUF_X1, BUF_E, BUF_D, BUF_T1, BUF_T2, BUF_T3, BUF_T4, BUF_YA, BUF_YB, BUF_COUNT,
                 BUF_EXT_HEAT9 = 100, BUF_EXT_FEAT = 101, BUF_NONE = -1 };

struct Lookup {
  std::map<std::string, std::pair<const float*, int64_t>> t;
  const float* get(const std::string& name, int64_t numel) const {
    auto it = t.find(name);
    if (it == t.end()) { set_error("missing tensor: " + name); return nullptr; }
    if (it->second.second != numel) {
      set_error("tensor " + name + " has " + std::to_string(it->second.second) + " elements, expected " + std::to_string(numel));
      return nullptr;
    }
    return it->second.first;
  }
};

struct GemmOp {
  // sources
  int bufA = BUF_NONE, CA = 0, a_pix_stride = 0;
  int bufB = BUF_NONE, CB = 0, b_pix_stride = 0, b_ch_off = 0, b_group_stride = 0;
  int H = 0, W = 0, Ho = 0, Wo = 0, ksize = 1, stride = 1;
  int N = 0, G = 1, K = 0;
  bool se = false;
  int act = ACT_NONE;
  int ncase = 1;
  int bufRes = BUF_NONE, res_stride = 0;
  int bufOut = BUF_NONE, out_layout = OUT_NHWC, out_stride = 0;
  int out_ch_base[MAX_GROUPS] = {0}, n_valid[MAX_GROUPS] = {0};
  bool has_scale = true;
  size_t w_off = 0, scale_off = 0, bias_off = 0, ktab_off = 0;
  std::vector<uint32_t> ktab;
  ConvTcPlan tc{0, 0, 0, 0, 1, 0};
};

struct Op {
  enum Type { STEM, GEMM, DW, SE, UP, TOPS } type;
  GemmOp g;                          // GEMM
  // STEM / DW / SE / UP / TOPS
  int n_heads = 0, head0 = 0, od[8] = {0}, pix_stride = 0, out_ch = 0;
  int bufIn = BUF_NONE, bufOut = BUF_NONE, C = 0, H = 0, W = 0, stride = 1, S = 0;
  size_t w_off = 0, scale_off = 0, bias_off = 0, w2_off = 0, b1_off = 0, b2_off = 0;
  bool fused_se = false;             // DW: squeeze + fc1 folded into the depthwise kernel (w1 at se_w1_off); SE: fc2 only
  int parity = 0;                    // unfused SE: number of spatial tiles (partial sums) of the depthwise kernel
  size_t se_w1_off = 0;
};

}  // namespace

}  // namespace ftc

using namespace ftc;

struct ftc_detector {
  ftc_detector_config cfg;
  int dtype = DT_F32;
  size_t esize = 4;
  std::vector<Op> ops;
  size_t buf_elems[BUF_COUNT] = {0};   // per image
  size_t se_c_max = 0;
  size_t se_part_max = 1, se_hid_part_max = 1;   // floats per image: unfused tile partial sums / fused fc1 shares
  size_t weight_bytes = 0;
  size_t scratch_off = 0;              // packing scratch (floats) inside the packed buffer
  std::vector<std::function<int(const Lookup&, char*, cudaStream_t)>> pack_tasks;
  char* packed = nullptr;              // bound at pack time
  int input_format = 0;                // FTC_INPUT_*
  int Hq = 0, Wq = 0;                  // output resolution (H/4)
  int early_end = 0;                   // ops [0, early_end) = stem + features[1..3]: their surviving outputs (taps x1, x2) live in buffers of their own
  int tapC[4] = {0, 0, 0, 0}, tapH[4] = {0, 0, 0, 0}, tapW[4] = {0, 0, 0, 0};

  size_t walloc(size_t bytes) { size_t o = weight_bytes; weight_bytes = align_up(weight_bytes + bytes, 256); return o; }
  void need(int buf, size_t elems) { if (buf >= 0 && buf < BUF_COUNT && elems > buf_elems[buf]) buf_elems[buf] = elems; }

  bool use_tc = false, allow_tma = false;
  // weight bytes of a GEMM op in the layout of the selected backend (fills g.tc for tcgen05)
  size_t gemm_weight_bytes(GemmOp& g) {
    if (use_tc) {
      ConvGemmParams q; memset(&q, 0, sizeof(q)); q.N = g.N; q.K = g.K; q.G = g.G;
      q.H = g.H; q.W = g.W; q.stride = g.stride; q.pad = (g.ksize - 1) / 2;
      q.CA = g.CA; q.CB = g.CB; q.a_pix_stride = g.a_pix_stride; q.b_pix_stride = g.b_pix_stride; q.b_group_stride = g.b_group_stride;
      conv_gemm_tc_plan(q, &g.tc, allow_tma);
      g.K = g.tc.NKB * KBLOCK;           // the TMA paths pad the channels of each source to whole 64-wide chunks
      return conv_tc_weight_bytes(g.tc, g.G);
    }
    return (size_t)g.G * g.N * g.K * esize;
  }
  // pack channels [c_off, c_off+C) of an OIHW fp32 weight as group `grp`'s rows, k columns from k_off
  static int pack_w(const GemmOp& gc, int dt, bool tc, char* base, const float* w, int O, int Itot, int ks, int c_off, int C,
                    int k_off, int grp, const float* cscale, cudaStream_t s) {
    if (tc) {
      const bool halo = gc.tc.tma == TMA_HALO;
      if (halo && k_off) k_off = 9 * KBLOCK * gc.tc.nGA;     // second source starts after source A's padded chunks
      return pack_conv_weight_tc(base + gc.w_off, w, O, Itot, ks, ks, c_off, C, k_off, gc.K, grp * gc.tc.NT * gc.tc.BN, gc.tc.BN,
                                 cscale, s, halo ? (gc.tc.kb32 ? 2 : 1) : 0);
    }
    return pack_conv_weight(base + gc.w_off, dt, w, O, Itot, ks, ks, c_off, C, k_off, gc.K, grp * gc.N, cscale, s);
  }

  // ---- plan helpers ------------------------------------------------------------------------
  // dense conv + folded BN (+act, +residual, +SE) reading `in` (C channels, HxW) -> `out`
  void add_conv_bn(const std::string& p, int in, int out, int Cin, int Cout, int H, int W, int ksize, int stride, int act,
                   int res, bool se, float eps) {
    Op op; op.type = Op::GEMM;
    GemmOp& g = op.g;
    g.bufA = in; g.CA = Cin; g.a_pix_stride = Cin;
    g.H = H; g.W = W; g.ksize = ksize; g.stride = stride;
    g.Ho = (H - 1) / stride + 1; g.Wo = (W - 1) / stride + 1;
    g.N = Cout; g.G = 1; g.se = se; g.act = act;
    g.bufRes = res; g.res_stride = Cout;
    g.bufOut = out; g.out_layout = OUT_NHWC; g.out_stride = Cout;
    g.out_ch_base[0] = 0; g.n_valid[0] = Cout;
    g.ktab = make_ktab(Cin, 0, ksize, &g.K);
    g.w_off = walloc(gemm_weight_bytes(g));
    g.scale_off = walloc((size_t)Cout * 4);
    g.bias_off = walloc((size_t)Cout * 4);
    g.ktab_off = walloc(g.ktab.size() * 4);
    need(out, (size_t)g.Ho * g.Wo * Cout);
    const int dt = dtype; const GemmOp gc = g; const bool tc = use_tc;
    pack_tasks.push_back([=](const Lookup& L, char* base, cudaStream_t s) -> int {
      const float* w = L.get(p + ".0.weight", (int64_t)Cout * Cin * ksize * ksize);
      const float* ga = L.get(p + ".1.weight", Cout); const float* be = L.get(p + ".1.bias", Cout);
      const float* mu = L.get(p + ".1.running_mean", Cout); const float* va = L.get(p + ".1.running_var", Cout);
      if (!w || !ga || !be || !mu || !va) return -1;
      int rc = pack_w(gc, dt, tc, base, w, Cout, Cin, ksize, 0, Cin, 0, 0, nullptr, s);
      if (rc) return rc;
      rc = bn_fold((float*)(base + gc.scale_off), (float*)(base + gc.bias_off), ga, be, mu, va, eps, Cout, s);
      if (rc) return rc;
      FTC_CHECK_CUDA(cudaMemcpyAsync(base + gc.ktab_off, gc.ktab.data(), gc.ktab.size() * 4, cudaMemcpyHostToDevice, s));
      return 0;
    });
    ops.push_back(std::move(op));
  }

  int build();
};

int ftc_detector::build() {
  const ftc_detector_config& c = cfg;
  dtype = c.precision == FTC_PREC_BF16 ? DT_BF16 : DT_F32;
  esize = dtype == DT_BF16 ? 2 : 4;
  use_tc = c.gemm_backend == FTC_GEMM_TCGEN05 || c.gemm_backend == FTC_GEMM_TCGEN05_IM2COL;
  allow_tma = c.gemm_backend == FTC_GEMM_TCGEN05;
  FTC_REQUIRE(!use_tc || dtype == DT_BF16, "the tcgen05 backend needs FTC_PREC_BF16");
  FTC_REQUIRE(c.height % 32 == 0 && c.width % 32 == 0, "input size must be a multiple of 32");
  FTC_REQUIRE(c.n_stages >= 5 && c.n_stages <= FTC_MAX_STAGES, "stage count");
  FTC_REQUIRE(c.n_heads >= 1 && c.n_heads <= FTC_MAX_HEADS, "head count");
  const float EPS_BB = 1e-3f, EPS_HEAD = 1e-5f;
  scratch_off = walloc(4 * 4096 * sizeof(float));   // in_bn scale / shift scratch while packing

  // ---- stem ----
  int H = c.height / 2, W = c.width / 2;
  {
    Op op; op.type = Op::STEM; op.bufOut = BUF_X0; op.C = c.stem_out; op.H = c.height; op.W = c.width;
    op.w_off = walloc(27 * c.stem_out * 4); op.scale_off = walloc(c.stem_out * 4); op.bias_off = walloc(c.stem_out * 4);
    need(BUF_X0, (size_t)H * W * c.stem_out);
    const Op oc = op; const int Cout = c.stem_out;
    pack_tasks.push_back([=](const Lookup& L, char* base, cudaStream_t s) -> int {
      const std::string p = "backbone.features.0";
      const float* w = L.get(p + ".0.weight", (int64_t)Cout * 27);
      const float* ga = L.get(p + ".1.weight", Cout); const float* be = L.get(p + ".1.bias", Cout);
      const float* mu = L.get(p + ".1.running_mean", Cout); const float* va = L.get(p + ".1.running_var", Cout);
      if (!w || !ga || !be || !mu || !va) return -1;
      int rc = transpose_f32((float*)(base + oc.w_off), w, Cout, 27, s);
      if (rc) return rc;
      return bn_fold((float*)(base + oc.scale_off), (float*)(base + oc.bias_off), ga, be, mu, va, EPS_BB, Cout, s);
    });
    ops.push_back(op);
  }

  // ---- backbone stages ----
  int cur = BUF_X0;
  int n_fused_se = 0;
  int tap_bufs[3] = {BUF_T1, BUF_T2, BUF_T3};
  int tap_C[4] = {0, 0, 0, 0}, tap_H[4] = {0, 0, 0, 0};
  int ntap = 0;
  for (int si = 0; si < c.n_stages; ++si) {
    const ftc_stage_cfg& st = c.stages[si];
    FTC_REQUIRE(st.kernel == 3, "only 3x3 stages (all EfficientNetV2 configs)");
    const bool is_tap_stage = (si + 1 == 2 || si + 1 == 3 || si + 1 == 5);   // features[2], [3], [5]
    for (int li = 0; li < st.layers; ++li) {
      int cin = li == 0 ? st.cin : st.cout;
      int stride = li == 0 ? st.stride : 1;
      int exp = cin * st.expand;
      bool res = stride == 1 && cin == st.cout;
      int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
      int out = (is_tap_stage && li == st.layers - 1) ? tap_bufs[ntap] : (cur == BUF_X0 ? BUF_X1 : BUF_X0);
      std::string p = "backbone.features." + std::to_string(si + 1) + "." + std::to_string(li) + ".block";
      if (st.fused) {
        if (exp != cin) {
          add_conv_bn(p + ".0", cur, BUF_E, cin, exp, H, W, 3, stride, ACT_SILU, BUF_NONE, false, EPS_BB);
          add_conv_bn(p + ".1", BUF_E, out, exp, st.cout, Ho, Wo, 1, 1, ACT_NONE, res ? cur : BUF_NONE, false, EPS_BB);
        } else {
          add_conv_bn(p + ".0", cur, out, cin, st.cout, H, W, 3, stride, ACT_SILU, res ? cur : BUF_NONE, false, EPS_BB);
        }
      } else {
        int sq = cin / 4 > 1 ? cin / 4 : 1;
        add_conv_bn(p + ".0", cur, BUF_E, cin, exp, H, W, 1, 1, ACT_SILU, BUF_NONE, false, EPS_BB);
        const bool fused_se = dwconv3x3_se_supported(H, W, exp, stride) && !getenv("FTC_NO_DW_STRIP");
        const size_t se_w1_off = walloc((size_t)sq * exp * 4);
        // unfused depthwise kernel: `parity` carries its spatial tile count (se_fc1 adds that many partial sums in order)
        const int parity = fused_se ? 0 : dwconv3x3_tiles(H, W, stride, dtype);
        if (fused_se) { if ((size_t)(exp / 32) * sq > se_hid_part_max) se_hid_part_max = (size_t)(exp / 32) * sq; }
        else if ((size_t)parity * exp > se_part_max) se_part_max = (size_t)parity * exp;
        {
          Op op; op.type = Op::DW; op.bufIn = BUF_E; op.bufOut = BUF_D; op.C = exp; op.H = H; op.W = W; op.stride = stride;
          op.fused_se = fused_se; op.parity = parity; op.se_w1_off = se_w1_off; op.S = sq;
          op.w_off = walloc(9 * exp * 4); op.scale_off = walloc(exp * 4); op.bias_off = walloc(exp * 4);
          need(BUF_D, (size_t)Ho * Wo * exp);
          if ((size_t)exp > se_c_max) se_c_max = exp;
          const Op oc = op; const std::string q = p + ".1";
          pack_tasks.push_back([=](const Lookup& L, char* base, cudaStream_t s) -> int {
            const float* w = L.get(q + ".0.weight", (int64_t)exp * 9);
            const float* ga = L.get(q + ".1.weight", exp); const float* be = L.get(q + ".1.bias", exp);
            const float* mu = L.get(q + ".1.running_mean", exp); const float* va = L.get(q + ".1.running_var", exp);
            if (!w || !ga || !be || !mu || !va) return -1;
            int rc = transpose_f32((float*)(base + oc.w_off), w, exp, 9, s);
            if (rc) return rc;
            return bn_fold((float*)(base + oc.scale_off), (float*)(base + oc.bias_off), ga, be, mu, va, EPS_BB, exp, s);
          });
          ops.push_back(op);
        }
        {
          Op op; op.type = Op::SE; op.C = exp; op.S = sq; op.H = Ho; op.W = Wo;
          op.fused_se = fused_se; op.parity = parity;
          op.w_off = se_w1_off; op.b1_off = walloc(sq * 4);
          op.w2_off = walloc((size_t)sq * exp * 4); op.b2_off = walloc(exp * 4);
          const Op oc = op; const std::string q = p + ".2";
          pack_tasks.push_back([=](const Lookup& L, char* base, cudaStream_t s) -> int {
            const float* w1 = L.get(q + ".fc1.weight", (int64_t)sq * exp); const float* b1 = L.get(q + ".fc1.bias", sq);
            const float* w2 = L.get(q + ".fc2.weight", (int64_t)sq * exp); const float* b2 = L.get(q + ".fc2.bias", exp);
            if (!w1 || !b1 || !w2 || !b2) return -1;
            FTC_CHECK_CUDA(cudaMemcpyAsync(base + oc.w_off, w1, (size_t)sq * exp * 4, cudaMemcpyDeviceToDevice, s));
            FTC_CHECK_CUDA(cudaMemcpyAsync(base + oc.b1_off, b1, sq * 4, cudaMemcpyDeviceToDevice, s));
            FTC_CHECK_CUDA(cudaMemcpyAsync(base + oc.b2_off, b2, exp * 4, cudaMemcpyDeviceToDevice, s));
            return transpose_f32((float*)(base + oc.w2_off), w2, exp, sq, s);   // [C][S] -> [S][C]
          });
          ops.push_back(op);
        }
        add_conv_bn(p + ".3", BUF_D, out, exp, st.cout, Ho, Wo, 1, 1, ACT_NONE, res ? cur : BUF_NONE, true, EPS_BB);
      }
      cur = out; H = Ho; W = Wo;
    }
    if (is_tap_stage) { tap_C[ntap] = st.cout; tap_H[ntap] = H; ++ntap; }
    if (si + 1 == 3) early_end = (int)ops.size();
  }
  FTC_REQUIRE(ntap == 3, "expected three intermediate taps");
  add_conv_bn("backbone.features." + std::to_string(c.n_stages + 1), cur, BUF_T4, c.stages[c.n_stages - 1].cout,
              c.last_channel, H, W, 1, 1, ACT_SILU, BUF_NONE, false, EPS_BB);
  tap_C[3] = c.last_channel; tap_H[3] = H;
  const int tap_buf4[4] = {BUF_T1, BUF_T2, BUF_T3, BUF_T4};
  const int tap_W[4] = {c.width / 4, c.width / 8, c.width / 16, c.width / 32};
  for (int i = 0; i < 4; ++i) FTC_REQUIRE(tap_H[i] == c.height / (4 << i), "unexpected tap resolution");
  Hq = c.height / 4; Wq = c.width / 4;
  for (int i = 0; i < 4; ++i) { tapC[i] = tap_C[i]; tapH[i] = tap_H[i]; tapW[i] = tap_W[i]; }

  // ---- nine Leafmap heads, batched: N = n_heads * 192 per level ----
  const int NH = c.n_heads, CD = 192, NT = NH * CD;
  int ybuf = BUF_YA, ubuf = BUF_YB;
  for (int lv = 0; lv < 4; ++lv) {
    const int ti = 3 - lv;                  // tap index (x4 first)
    const int Ct = tap_C[ti], Hh = tap_H[ti], Ww = tap_W[ti];
    Op op; op.type = Op::GEMM;
    GemmOp& g = op.g;
    g.bufA = tap_buf4[ti]; g.CA = Ct; g.a_pix_stride = Ct;
    if (lv > 0) { g.bufB = ubuf; g.CB = CD; g.b_pix_stride = NT; g.b_ch_off = 0; g.b_group_stride = CD; }
    g.H = Hh; g.W = Ww; g.Ho = Hh; g.Wo = Ww; g.ksize = 3; g.stride = 1;
    g.N = CD; g.G = NH; g.act = ACT_GELU; g.ncase = 9;
    g.bufOut = ybuf; g.out_layout = OUT_NHWC; g.out_stride = NT;
    for (int h = 0; h < NH; ++h) { g.out_ch_base[h] = h * CD; g.n_valid[h] = CD; }
    g.ktab = make_ktab(Ct, lv > 0 ? CD : 0, 3, &g.K);
    g.w_off = walloc(gemm_weight_bytes(g));
    g.scale_off = walloc((size_t)NT * 4);
    g.bias_off = walloc((size_t)9 * NT * 4);
    g.ktab_off = walloc(g.ktab.size() * 4);
    need(ybuf, (size_t)Hh * Ww * NT);
    const GemmOp gc = g; const int dt = dtype; const size_t sc_off = scratch_off; const bool tc = use_tc;
    FTC_REQUIRE(Ct <= 4096, "tap too wide for packing scratch");
    for (int h = 0; h < NH; ++h) {
      const std::string hp = std::string(c.head_names[h]);
      pack_tasks.push_back([=](const Lookup& L, char* base, cudaStream_t s) -> int {
        const int Itot = Ct + (lv > 0 ? CD : 0);
        const std::string up = hp + ".upsamplers." + std::to_string(lv);
        const std::string ib = hp + ".in_bn." + std::to_string(ti);
        const float* w = L.get(up + ".0.weight", (int64_t)CD * Itot * 9);
        const float* ga = L.get(up + ".1.weight", CD); const float* be = L.get(up + ".1.bias", CD);
        const float* mu = L.get(up + ".1.running_mean", CD); const float* va = L.get(up + ".1.running_var", CD);
        const float* iga = L.get(ib + ".weight", Ct); const float* ibe = L.get(ib + ".bias", Ct);
        const float* imu = L.get(ib + ".running_mean", Ct); const float* iva = L.get(ib + ".running_var", Ct);
        if (!w || !ga || !be || !mu || !va || !iga || !ibe || !imu || !iva) return -1;
        float* in_scale = (float*)(base + sc_off);
        float* in_shift = in_scale + 4096;
        float* tmp_bias = in_scale + 8192;
        float* scale = (float*)(base + gc.scale_off) + h * CD;
        int rc = bn_fold(in_scale, in_shift, iga, ibe, imu, iva, EPS_HEAD, Ct, s);
        if (rc) return rc;
        // tap part: channels [Itot-Ct, Itot) of the concatenated input, in_bn scale folded into the weights
        rc = pack_w(gc, dt, tc, base, w, CD, Itot, 3, Itot - Ct, Ct, 0, h, in_scale, s);
        if (rc) return rc;
        if (lv > 0) {
          rc = pack_w(gc, dt, tc, base, w, CD, Itot, 3, 0, CD, 9 * Ct, h, nullptr, s);
          if (rc) return rc;
        }
        rc = bn_fold(scale, tmp_bias, ga, be, mu, va, EPS_HEAD, CD, s);
        if (rc) return rc;
        return leaf_bias_table((float*)(base + gc.bias_off) + h * CD, NT, w, CD, Itot, Itot - Ct, Ct, in_shift, scale,
                               tmp_bias, s);
      });
    }
    pack_tasks.push_back([=](const Lookup&, char* base, cudaStream_t s) -> int {
      FTC_CHECK_CUDA(cudaMemcpyAsync(base + gc.ktab_off, gc.ktab.data(), gc.ktab.size() * 4, cudaMemcpyHostToDevice, s));
      return 0;
    });
    ops.push_back(std::move(op));
    if (lv < 3) {
      Op up; up.type = Op::UP; up.bufIn = ybuf; up.bufOut = ubuf; up.C = NT; up.H = Hh; up.W = Ww;
      need(ubuf, (size_t)4 * Hh * Ww * NT);
      ops.push_back(up);
      // next level reads ubuf as source B and writes ybuf (already consumed by the upsample)
    }
  }

  // ---- top convs: 3x3 192 -> out_dim (+bias), NCHW fp32 outputs ----
  auto add_top = [&](int h0, int nh, int Npad, int ext_buf, int total_ch) {
    Op op; op.type = Op::GEMM;
    GemmOp& g = op.g;
    g.bufB = ybuf; g.CB = CD; g.b_pix_stride = NT; g.b_ch_off = h0 * CD; g.b_group_stride = CD;
    g.H = Hq; g.W = Wq; g.Ho = Hq; g.Wo = Wq; g.ksize = 3; g.stride = 1;
    g.N = Npad; g.G = nh; g.act = ACT_NONE; g.ncase = 1; g.has_scale = false;
    g.bufOut = ext_buf; g.out_layout = OUT_NCHW_F32; g.out_stride = total_ch;
    int base_ch = 0;
    for (int h = 0; h < nh; ++h) { g.out_ch_base[h] = base_ch; g.n_valid[h] = c.head_out[h0 + h]; base_ch += c.head_out[h0 + h]; }
    g.ktab = make_ktab(0, CD, 3, &g.K);
    g.w_off = walloc(gemm_weight_bytes(g));
    g.bias_off = walloc((size_t)nh * Npad * 4);
    g.ktab_off = walloc(g.ktab.size() * 4);
    const GemmOp gc = g; const int dt = dtype; const bool tc = use_tc;
    for (int h = 0; h < nh; ++h) {
      const std::string hp = std::string(c.head_names[h0 + h]);
      const int od = c.head_out[h0 + h];
      pack_tasks.push_back([=](const Lookup& L, char* base, cudaStream_t s) -> int {
        const float* w = L.get(hp + ".top_conv.0.weight", (int64_t)od * CD * 9);
        const float* b = L.get(hp + ".top_conv.0.bias", od);
        if (!w || !b) return -1;
        int rc = pack_w(gc, dt, tc, base, w, od, CD, 3, 0, CD, 0, h, nullptr, s);
        if (rc) return rc;
        FTC_CHECK_CUDA(cudaMemcpyAsync((float*)(base + gc.bias_off) + h * Npad, b, od * 4, cudaMemcpyDeviceToDevice, s));
        return 0;
      });
    }
    pack_tasks.push_back([=](const Lookup&, char* base, cudaStream_t s) -> int {
      FTC_CHECK_CUDA(cudaMemcpyAsync(base + gc.ktab_off, gc.ktab.data(), gc.ktab.size() * 4, cudaMemcpyHostToDevice, s));
      return 0;
    });
    ops.push_back(std::move(op));
  };
  int heat_ch = 0;
  for (int h = 0; h < NH - 1; ++h) { heat_ch += c.head_out[h]; FTC_REQUIRE(c.head_out[h] <= 2, "small head out_dim <= 2"); }
  {
    // the eight small heads: dedicated bandwidth kernel (head_top_conv), fp32 tap-major weights (as one grouped N = 16 tcgen05
    // GEMM they took 3.1 instead of 1.5 ms; that route was removed)
    Op op; op.type = Op::TOPS; op.bufIn = ybuf; op.H = Hq; op.W = Wq; op.n_heads = NH - 1; op.head0 = 0; op.pix_stride = NT;
    op.out_ch = heat_ch;
    for (int h = 0; h < NH - 1; ++h) op.od[h] = c.head_out[h];
    op.w_off = walloc((size_t)heat_ch * 9 * CD * 4);
    op.bias_off = walloc((size_t)heat_ch * 4);
    const Op oc = op;
    int row = 0;
    for (int h = 0; h < NH - 1; ++h) {
      const std::string hp = std::string(c.head_names[h]);
      const int od = c.head_out[h], r0 = row;
      pack_tasks.push_back([=](const Lookup& L, char* base, cudaStream_t s) -> int {
        const float* w = L.get(hp + ".top_conv.0.weight", (int64_t)od * CD * 9);
        const float* b = L.get(hp + ".top_conv.0.bias", od);
        if (!w || !b) return -1;
        int rc = pack_conv_weight(base + oc.w_off, DT_F32, w, od, CD, 3, 3, 0, CD, 0, 9 * CD, r0, nullptr, s);
        if (rc) return rc;
        FTC_CHECK_CUDA(cudaMemcpyAsync((float*)(base + oc.bias_off) + r0, b, od * 4, cudaMemcpyDeviceToDevice, s));
        return 0;
      });
      row += od;
    }
    ops.push_back(op);
  }
  int fpad = (c.head_out[NH - 1] + 15) / 16 * 16;
  add_top(NH - 1, 1, fpad, BUF_EXT_FEAT, c.head_out[NH - 1]);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// flops (2*MAC) of one op for `B` images
static double op_flops(const Op& op, int B) {
  switch (op.type) {
    case Op::STEM: return 2.0 * B * (op.H / 2) * (op.W / 2) * 27.0 * op.C;
    case Op::DW: { int Ho = (op.H - 1) / op.stride + 1, Wo = (op.W - 1) / op.stride + 1; return 2.0 * B * Ho * Wo * 9.0 * op.C; }
    case Op::SE: return 4.0 * B * (double)op.C * op.S;
    case Op::UP: return 0.0;
    case Op::TOPS: return 2.0 * B * op.H * op.W * 9.0 * 192.0 * op.out_ch;
    case Op::GEMM: {
      const GemmOp& g = op.g;
      double kreal = (double)g.ksize * g.ksize * (g.CA + g.CB);
      double n = 0;
      for (int i = 0; i < g.G; ++i) n += g.n_valid[i];
      return 2.0 * B * g.Ho * g.Wo * kreal * n;
    }
  }
  return 0.0;
}

// Runs ops [op_lo, op_hi) on images [i0, i0 + nb) of a batch of B (every activation tensor is [B][pixels][channels]: a range of
// images is a pointer offset per tensor).  i0 > 0 / nb < B only for the early ops (op_hi <= early_end): their intermediate
// tensors share the ping-pong buffers with different per-image sizes, so image ranges of DIFFERENT ops may overlap there; the
// early ops of one range run to completion before the next range starts, and what survives them (taps x1, x2) sits in buffers
// that hold one tensor only.
static int detector_forward_impl(ftc_detector* d, const float* images, int B, float* heat9, float* feat, float* heat10,
                                 void* workspace, size_t ws_bytes, cudaStream_t s, cudaEvent_t* ev = nullptr, int op_lo = 0,
                                 int op_hi = -1, int i0 = 0, int nb = -1) {
  FTC_REQUIRE(d->packed != nullptr, "ftc_detector_pack_weights must be called before forward");
  FTC_REQUIRE(ws_bytes >= ftc_detector_workspace_bytes(d, B), "workspace too small");
  const int n_ops = (int)d->ops.size();
  if (op_hi < 0) op_hi = n_ops;
  if (nb < 0) nb = B;
  FTC_REQUIRE(op_lo >= 0 && op_lo <= op_hi && op_hi <= n_ops && i0 >= 0 && nb > 0 && i0 + nb <= B, "bad op / image range");
  FTC_REQUIRE((i0 == 0 && nb == B) || op_hi <= d->early_end, "image ranges are for the early ops only");
  char* ws = (char*)workspace;
  char* bufp[BUF_COUNT];
  size_t off = 0;
  for (int i = 0; i < BUF_COUNT; ++i) { bufp[i] = ws + off; off += align_up(d->buf_elems[i] * B * d->esize, 256); }
  // SE scratch: every entry a kernel reads was written by the producing kernel of the same block (no accumulators to zero, no
  // atomics: the squeeze is bit-reproducible and independent of the batch size)
  float* se_sum = (float*)(ws + off); off += align_up(d->se_part_max * B * 4, 256);     // [B][tiles][C] partial sums (unfused path)
  float* se_scale = (float*)(ws + off); off += align_up(d->se_c_max * B * 4, 256);
  float* se_hid = (float*)(ws + off); off += align_up((size_t)256 * B * 4, 256);
  float* se_hid2 = (float*)(ws + off); off += align_up(d->se_hid_part_max * B * 4, 256);  // [B][C/32][S] fc1 shares (fused path)
  const size_t es = d->esize;
  auto bp = [&](int id, size_t pix = 0, size_t stride = 0) -> void* {      // image i0 of the [B][pix][stride] tensor in buffer id
    if (id == BUF_NONE) return nullptr;
    if (id == BUF_EXT_HEAT9) return heat9;
    if (id == BUF_EXT_FEAT) return feat;
    return bufp[id] + (size_t)i0 * pix * stride * es;
  };
  char* P = d->packed;
  for (int oi = op_lo; oi < op_hi; ++oi) {
    const Op& op = d->ops[oi];
    int rc = 0;
    if (ev) FTC_CHECK_CUDA(cudaEventRecord(ev[oi], s));
    switch (op.type) {
      case Op::STEM:
        rc = stem_conv(images + (size_t)i0 * 3 * op.H * op.W, d->input_format, bp(op.bufOut, (size_t)(op.H / 2) * (op.W / 2), op.C), d->dtype, nb,
                       op.H, op.W, op.C, (const float*)(P + op.w_off), (const float*)(P + op.scale_off), (const float*)(P + op.bias_off), s);
        break;
      case Op::DW: {
        const int Ho = (op.H - 1) / op.stride + 1, Wo = (op.W - 1) / op.stride + 1;
        void* in = bp(op.bufIn, (size_t)op.H * op.W, op.C);
        void* out = bp(op.bufOut, (size_t)Ho * Wo, op.C);
        if (op.fused_se)
          rc = dwconv3x3_se(in, out, d->dtype, nb, op.H, op.W, op.C, (const float*)(P + op.w_off),
                            (const float*)(P + op.scale_off), (const float*)(P + op.bias_off), (const float*)(P + op.se_w1_off),
                            op.S, se_hid2, s);
        else
          rc = dwconv3x3(in, out, d->dtype, nb, op.H, op.W, op.C, op.stride, (const float*)(P + op.w_off),
                         (const float*)(P + op.scale_off), (const float*)(P + op.bias_off), se_sum, s);
        break;
      }
      case Op::SE:
        if (op.fused_se)
          rc = se_fc2_hid(se_hid2, op.C / 32, se_scale, nb, op.C,
                          op.S, (const float*)(P + op.b1_off), (const float*)(P + op.w2_off), (const float*)(P + op.b2_off), s);
        else
          rc = se_fc(se_sum, op.parity /* = tile count of the depthwise kernel */, se_scale, se_hid, nb, op.C, op.S, 1.0f / (float)(op.H * op.W), (const float*)(P + op.w_off),
                     (const float*)(P + op.b1_off), (const float*)(P + op.w2_off), (const float*)(P + op.b2_off), s);
        break;
      case Op::TOPS:
        rc = head_top_conv(bp(op.bufIn), d->dtype, op.pix_stride, op.head0, op.n_heads, op.od, (const float*)(P + op.w_off),
                           (const float*)(P + op.bias_off), heat9, op.out_ch, nb, op.H, op.W, s);
        break;
      case Op::UP:
        rc = upsample2x(bp(op.bufIn), bp(op.bufOut), d->dtype, nb, op.H, op.W, op.C, s);
        break;
      case Op::GEMM: {
        const GemmOp& g = op.g;
        ConvGemmParams p;
        memset(&p, 0, sizeof(p));
        p.B = nb; p.H = g.H; p.W = g.W; p.Ho = g.Ho; p.Wo = g.Wo; p.stride = g.stride; p.pad = (g.ksize - 1) / 2;
        p.M = nb * g.Ho * g.Wo; p.N = g.N; p.G = g.G; p.K = g.K;
        p.srcA = bp(g.bufA, (size_t)g.H * g.W, g.a_pix_stride); p.a_pix_stride = g.a_pix_stride; p.a_ch_off = 0;
        p.srcB = bp(g.bufB, (size_t)g.H * g.W, g.b_pix_stride); p.b_pix_stride = g.b_pix_stride; p.b_ch_off = g.b_ch_off; p.b_group_stride = g.b_group_stride;
        p.CA = g.CA; p.CB = g.CB;
        p.ktab = (const uint32_t*)(P + g.ktab_off);
        p.a_scale = g.se ? se_scale : nullptr; p.a_scale_stride = g.CA;
        p.w = P + g.w_off;
        p.scale = g.has_scale ? (const float*)(P + g.scale_off) : nullptr;
        p.bias_tab = (const float*)(P + g.bias_off); p.ncase = g.ncase;
        p.act = g.act;
        p.res1 = bp(g.bufRes, (size_t)g.Ho * g.Wo, g.res_stride); p.res1_stride = g.res_stride; p.res1_row_mod = 0;
        p.res2 = nullptr; p.res2_stride = 0;
        p.out = g.out_layout == OUT_NHWC ? bp(g.bufOut, (size_t)g.Ho * g.Wo, g.out_stride) : bp(g.bufOut);
        p.out_layout = g.out_layout; p.out_stride = g.out_stride;
        for (int i = 0; i < MAX_GROUPS; ++i) { p.out_ch_base[i] = g.out_ch_base[i]; p.n_valid[i] = g.n_valid[i]; }
        p.dtype = d->dtype;
        p.tc = g.tc;
        if (d->use_tc) rc = conv_gemm_tc(p, s);
        else rc = conv_gemm_simt(p, s);
        break;
      }
    }
    if (rc) return rc;
  }
  if (ev && op_hi == n_ops) FTC_CHECK_CUDA(cudaEventRecord(ev[n_ops], s));
  if (heat10 && op_hi == n_ops) return peak_pick(heat9, heat10, B, d->Hq, d->Wq, s);
  return 0;
}

extern "C" {

int ftc_detector_create(const ftc_detector_config* cfg, ftc_detector** out) {
  FTC_REQUIRE(cfg && out, "null argument");
  ftc_detector* d = new ftc_detector();
  d->cfg = *cfg;
  int rc = d->build();
  if (rc) { delete d; return rc; }
  *out = d;
  return 0;
}

void ftc_detector_destroy(ftc_detector* d) { delete d; }

size_t ftc_detector_weight_bytes(const ftc_detector* d) { return d ? d->weight_bytes : 0; }

size_t ftc_detector_workspace_bytes(const ftc_detector* d, int batch) {
  if (!d) return 0;
  size_t off = 0;
  for (int i = 0; i < BUF_COUNT; ++i) off += align_up(d->buf_elems[i] * batch * d->esize, 256);
  off += align_up(d->se_part_max * batch * 4, 256) + align_up(d->se_c_max * batch * 4, 256) + align_up((size_t)256 * batch * 4, 256) +
         align_up(d->se_hid_part_max * batch * 4, 256);
  return off + 256;
}

int ftc_detector_set_input_format(ftc_detector* d, int fmt) {
  FTC_REQUIRE(d && (fmt == FTC_INPUT_NCHW_UNIT || fmt == FTC_INPUT_NHWC_255), "bad input format");
  d->input_format = fmt;
  return 0;
}

int ftc_detector_num_ops(const ftc_detector* d) { return d ? (int)d->ops.size() : 0; }

int ftc_detector_forward_timed(ftc_detector* d, const float* images, int batch, float* heat9, float* feat, void* workspace,
                               size_t workspace_bytes, void* stream, int max_ops, float* op_ms, double* op_flop,
                               int* op_kind) {
  FTC_REQUIRE(d && images && heat9 && feat && workspace && op_ms && op_flop && op_kind, "bad argument");
  const int n = (int)d->ops.size();
  FTC_REQUIRE(max_ops >= n, "op arrays too small");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) FTC_CHECK_CUDA(cudaEventCreate(&e));
  int rc = detector_forward_impl(d, images, batch, heat9, feat, nullptr, workspace, workspace_bytes, s, ev.data());
  if (rc == 0) {
    FTC_CHECK_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) {
      FTC_CHECK_CUDA(cudaEventElapsedTime(&op_ms[i], ev[i], ev[i + 1]));
      op_flop[i] = op_flops(d->ops[i], batch);
      const Op& op = d->ops[i];
      // kind: 0 stem, 1 dense 3x3 conv, 2 1x1 conv, 3 depthwise, 4 SE, 5 upsample
      op_kind[i] = op.type == Op::STEM ? 0 : op.type == Op::GEMM ? (op.g.ksize == 3 ? 1 : 2) : op.type == Op::DW ? 3
                   : op.type == Op::SE ? 4 : op.type == Op::UP ? 5 : 6;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int ftc_detector_tap(const ftc_detector* d, int tap, int batch, void* workspace, void** ptr, int* channels, int* h, int* w) {
  FTC_REQUIRE(d && workspace && ptr && tap >= 0 && tap < 4, "bad argument");
  const int ids[4] = {BUF_T1, BUF_T2, BUF_T3, BUF_T4};
  size_t off = 0;
  for (int i = 0; i < ids[tap]; ++i) off += align_up(d->buf_elems[i] * batch * d->esize, 256);
  *ptr = (char*)workspace + off;
  if (channels) *channels = d->tapC[tap];
  if (h) *h = d->tapH[tap];
  if (w) *w = d->tapW[tap];
  return 0;
}

int ftc_detector_pack_weights(ftc_detector* d, int n, const char* const* names, const void* const* ptrs,
                              const int64_t* numels, void* packed, size_t packed_bytes, void* stream) {
  FTC_REQUIRE(d && names && ptrs && numels && packed, "null argument");
  FTC_REQUIRE(packed_bytes >= d->weight_bytes, "packed buffer too small");
  cudaStream_t s = (cudaStream_t)stream;
  Lookup L;
  for (int i = 0; i < n; ++i) L.t[names[i]] = {(const float*)ptrs[i], numels[i]};
  FTC_CHECK_CUDA(cudaMemsetAsync(packed, 0, d->weight_bytes, s));
  for (auto& task : d->pack_tasks) {
    int rc = task(L, (char*)packed, s);
    if (rc) return rc;
  }
  // ktab uploads come from host vectors owned by the plan: make sure they are consumed before returning
  FTC_CHECK_CUDA(cudaStreamSynchronize(s));
  d->packed = (char*)packed;
  return 0;
}

int ftc_detector_forward_part(ftc_detector* d, const float* images, int batch, int part, int image0, int n_images, float* heat9,
                              float* feat, float* heat10, void* workspace, size_t workspace_bytes, void* stream) {
  FTC_REQUIRE(d && images && workspace && batch > 0 && (part == FTC_PART_EARLY || part == FTC_PART_REST), "bad argument");
  if (part == FTC_PART_EARLY)
    return detector_forward_impl(d, images, batch, nullptr, nullptr, nullptr, workspace, workspace_bytes, (cudaStream_t)stream, nullptr,
                                 0, d->early_end, image0, n_images);
  FTC_REQUIRE(heat9 && feat, "bad argument");
  return detector_forward_impl(d, images, batch, heat9, feat, heat10, workspace, workspace_bytes, (cudaStream_t)stream, nullptr,
                               d->early_end, -1, 0, batch);
}

int ftc_detector_forward(ftc_detector* d, const float* images, int batch, float* heat9, float* feat, float* heat10,
                         void* workspace, size_t workspace_bytes, void* stream) {
  FTC_REQUIRE(d && images && heat9 && feat && workspace && batch > 0, "bad argument");
  return detector_forward_impl(d, images, batch, heat9, feat, heat10, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
