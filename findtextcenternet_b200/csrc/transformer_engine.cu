// Transformer engine: Encoder / Decoder / mask-predict loop of the reference (models/transformer.py) as a static
// plan over the shared GEMM kernels (conv_gemm_*.cu, "1x1 conv" with H=W=1) and the kernels in transformer_ops.cu.
//
//  * (x + P) W^T is evaluated as x W^T + (P W^T): the position tables P W^T are folded once at pack time and enter the
//    projection GEMM's epilogue as a row-periodic residual (res1_row_mod = sequence length).  Q, K and V of a
//    self-attention share one GEMM (N = 3d); cross-attention K/V of every decoder layer are computed once per
//    encoder pass and reused by all (<= 8) mask-predict decoder passes.
//  * w1 / wg of SwiGLU are row-interleaved into one GEMM whose epilogue emits w1(x) * silu(wg(x)).
//  * residual adds (x + _x + skip) ride in the out-proj / w2 GEMM epilogues; LayerNorm is a row kernel.
//  * the three output heads are one grouped GEMM writing fp32 logits [rows, 3*HEAD_LD].
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../../include/ftc_b200.h"
#include "conv_gemm.cuh"
#include "detector_ops.cuh"
#include "transformer_ops.cuh"

namespace ftc {
namespace {

constexpr int HEAD_LD = 1104;          // >= max modulus (1097), multiple of 16
constexpr int MODS[3] = {1091, 1093, 1097};
constexpr float LN_EPS = 1e-5f;

struct TLookup {
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

// one nn.Linear (or a fusion of several) in engine form
struct Lin {
  int K = 0, Kpad = 0, N = 0, G = 1;       // N per group
  int n_valid[MAX_GROUPS] = {0};
  size_t w_off = 0, bias_off = 0, ktab_off = 0, tab_off = 0;   // tab: row-periodic residual table [rows][G*N] (dtype)
  bool has_bias = false, has_tab = false;
  int tab_rows = 0;
  std::vector<uint32_t> ktab;
  ConvTcPlan tc{0, 0, 0, 0, 1, 0};
};

struct LNw { size_t g_off = 0, b_off = 0; };

struct EncBlock { Lin qkv, out, w1g, w2; LNw n1, n2; };
struct DecBlock { Lin qkv, out, cq, ckv, cout, w1g, w2; LNw n1, n2, n3; };

}  // namespace
}  // namespace ftc

using namespace ftc;

struct ftc_transformer {
  ftc_transformer_config cfg;
  int dtype = DT_F32;
  size_t esize = 4;
  bool use_tc = false;
  int d = 0, heads = 0, hd = 0, kin_pad = 0;
  size_t weight_bytes = 0;
  size_t scratch_off = 0, scratch_floats = 0;
  Lin enc_embed, heads_out;
  LNw enc_norm, dec_norm;
  size_t enc_pos_off = 0;                       // encoder pos table [max_enc][d] (dtype) = res table of enc_embed
  size_t dec_pos_off = 0, dec_emb_off[3] = {0, 0, 0};   // fp32 tables for the decoder embedding kernel
  std::vector<EncBlock> enc;
  std::vector<DecBlock> dec;
  std::vector<std::function<int(const TLookup&, char*, cudaStream_t)>> pack_tasks;
  char* packed = nullptr;
  // state of the last encode() (lives in the caller's workspace)
  int enc_B = 0, enc_L = 0;

  size_t walloc(size_t bytes) { size_t o = weight_bytes; weight_bytes = align_up(weight_bytes + bytes, 256); return o; }

  void plan_lin(Lin& l, int K, int N, int G, bool bias, int tab_rows) {
    l.K = K; l.N = N; l.G = G; l.has_bias = bias; l.has_tab = tab_rows > 0; l.tab_rows = tab_rows;
    for (int g = 0; g < G; ++g) l.n_valid[g] = N;
    l.ktab = make_ktab(K, 0, 1, &l.Kpad);
    if (use_tc) {
      ConvGemmParams q; memset(&q, 0, sizeof(q)); q.N = N; q.K = l.Kpad;
      q.CA = K; q.stride = 1; q.pad = 0; q.a_pix_stride = 8;      // rows-mode TMA operand path (the real row stride is a launch argument)
      conv_gemm_tc_plan(q, &l.tc, !getenv("FTC_TF_NO_TMA"));
      l.w_off = walloc(conv_tc_weight_bytes(l.tc, G));
    } else {
      l.w_off = walloc((size_t)G * N * l.Kpad * esize);
    }
    l.ktab_off = walloc(l.ktab.size() * 4);
    if (bias) l.bias_off = walloc((size_t)G * N * 4);
    if (tab_rows > 0) l.tab_off = walloc((size_t)tab_rows * G * N * esize);
  }
  // rows [row0, row0+O) of group `grp` <- fp32 weight [O][K]
  int pack_rows(const Lin& l, char* base, const float* w, int O, int row0, int grp, cudaStream_t s) const {
    if (use_tc)
      return pack_conv_weight_tc(base + l.w_off, w, O, l.K, 1, 1, 0, l.K, 0, l.Kpad, grp * l.tc.NT * l.tc.BN + row0, l.tc.BN,
                                 nullptr, s);
    return pack_conv_weight(base + l.w_off, dtype, w, O, l.K, 1, 1, 0, l.K, 0, l.Kpad, grp * l.N + row0, nullptr, s);
  }
  int upload_ktab(const Lin& l, char* base, cudaStream_t s) const {
    FTC_CHECK_CUDA(cudaMemcpyAsync(base + l.ktab_off, l.ktab.data(), l.ktab.size() * 4, cudaMemcpyHostToDevice, s));
    return 0;
  }
  // table[r][col0 .. col0+d) = P[r] . W^T  (fp32 SIMT GEMM straight on the state_dict tensors), then cast to dtype
  int fold_pos(const Lin& l, char* base, const float* P, int rows, const float* W, int col0, cudaStream_t s) const {
    float* scratch = (float*)(base + scratch_off);
    FTC_REQUIRE((size_t)rows * d + 64 <= scratch_floats, "pos-fold scratch too small");
    ConvGemmParams p; memset(&p, 0, sizeof(p));
    p.B = rows; p.H = 1; p.W = 1; p.Ho = 1; p.Wo = 1; p.stride = 1; p.pad = 0;
    p.M = rows; p.N = d; p.G = 1; p.K = d;
    p.srcA = P; p.a_pix_stride = d;
    // a d-wide 1x1 chunk table lives at the end of the scratch area
    std::vector<uint32_t> kt; int K2 = 0; kt = make_ktab(d, 0, 1, &K2);
    FTC_REQUIRE(K2 == d, "embed_dim must be a multiple of 64");
    uint32_t* ktd = (uint32_t*)(scratch + scratch_floats);
    FTC_CHECK_CUDA(cudaMemcpyAsync(ktd, kt.data(), kt.size() * 4, cudaMemcpyHostToDevice, s));
    FTC_CHECK_CUDA(cudaStreamSynchronize(s));   // kt is a host temporary
    p.ktab = ktd;
    p.w = W; p.act = ACT_NONE; p.ncase = 1;
    p.out = scratch; p.out_layout = OUT_NHWC; p.out_stride = d; p.out_ch_base[0] = 0; p.n_valid[0] = d;
    p.dtype = DT_F32;
    int rc = conv_gemm_simt(p, s);
    if (rc) return rc;
    // scatter into the [rows][G*N] table at column col0
    const int ld = l.G * l.N;
    for (int r = 0; r < rows; ++r) {   // row-wise cast (rows <= a few hundred, pack time only)
      rc = cast_f32(base + l.tab_off + ((size_t)r * ld + col0) * esize, dtype, scratch + (size_t)r * d, d, s);
      if (rc) return rc;
    }
    return 0;
  }

  void plan_ln(LNw& n) { n.g_off = walloc((size_t)d * 4); n.b_off = walloc((size_t)d * 4); }
  void task_ln(const LNw n, const std::string name) {
    const int dd = d;
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      const float* g = L.get(name + ".weight", dd); const float* b = L.get(name + ".bias", dd);
      if (!g || !b) return -1;
      FTC_CHECK_CUDA(cudaMemcpyAsync(base + n.g_off, g, dd * 4, cudaMemcpyDeviceToDevice, s));
      FTC_CHECK_CUDA(cudaMemcpyAsync(base + n.b_off, b, dd * 4, cudaMemcpyDeviceToDevice, s));
      return 0;
    });
  }
  // self-attention fused QKV (keys use pos_emb_q: models/transformer.py:100-108)
  void task_qkv(const Lin l, const std::string p, int maxlen) {
    const int dd = d;
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      const float* wq = L.get(p + ".q_proj.weight", (int64_t)dd * dd); const float* wk = L.get(p + ".k_proj.weight", (int64_t)dd * dd);
      const float* wv = L.get(p + ".v_proj.weight", (int64_t)dd * dd); const float* pq = L.get(p + ".pos_emb_q.encoding", (int64_t)maxlen * dd);
      if (!wq || !wk || !wv || !pq) return -1;
      int rc;
      if ((rc = pack_rows(l, base, wq, dd, 0, 0, s))) return rc;
      if ((rc = pack_rows(l, base, wk, dd, dd, 0, s))) return rc;
      if ((rc = pack_rows(l, base, wv, dd, 2 * dd, 0, s))) return rc;
      if ((rc = upload_ktab(l, base, s))) return rc;
      FTC_CHECK_CUDA(cudaMemsetAsync(base + l.tab_off, 0, (size_t)l.tab_rows * l.N * esize, s));
      if ((rc = fold_pos(l, base, pq, maxlen, wq, 0, s))) return rc;
      return fold_pos(l, base, pq, maxlen, wk, dd, s);
    });
  }
  void task_plain(const Lin l, const std::string wname, const std::string bname) {
    const int K = l.K, N = l.N;
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      const float* w = L.get(wname, (int64_t)N * K);
      if (!w) return -1;
      int rc;
      if ((rc = pack_rows(l, base, w, N, 0, 0, s))) return rc;
      if ((rc = upload_ktab(l, base, s))) return rc;
      if (!bname.empty()) {
        const float* b = L.get(bname, N);
        if (!b) return -1;
        FTC_CHECK_CUDA(cudaMemcpyAsync(base + l.bias_off, b, (size_t)N * 4, cudaMemcpyDeviceToDevice, s));
      }
      return 0;
    });
  }
  // SwiGLU first layer: rows (2i, 2i+1) = (w1[i], wg[i]) so the GEMM epilogue can gate in registers
  void task_w1g(const Lin l, const std::string p) {
    const int dd = d;
    const size_t so = scratch_off;
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      const float* w1 = L.get(p + ".w1.weight", (int64_t)2 * dd * dd); const float* b1 = L.get(p + ".w1.bias", 2 * dd);
      const float* wg = L.get(p + ".wg.weight", (int64_t)2 * dd * dd); const float* bg = L.get(p + ".wg.bias", 2 * dd);
      if (!w1 || !b1 || !wg || !bg) return -1;
      float* scratch = (float*)(base + so);
      int rc;
      if ((rc = interleave_rows_f32(scratch, w1, wg, 2 * dd, dd, s))) return rc;
      if ((rc = pack_rows(l, base, scratch, 4 * dd, 0, 0, s))) return rc;
      if ((rc = interleave_rows_f32((float*)(base + l.bias_off), b1, bg, 2 * dd, 1, s))) return rc;
      return upload_ktab(l, base, s);
    });
  }

  int build();
};

int ftc_transformer::build() {
  const ftc_transformer_config& c = cfg;
  dtype = c.precision == FTC_PREC_BF16 ? DT_BF16 : DT_F32;
  esize = dtype == DT_BF16 ? 2 : 4;
  use_tc = c.gemm_backend == FTC_GEMM_TCGEN05 || c.gemm_backend == FTC_GEMM_TCGEN05_IM2COL;
  FTC_REQUIRE(!use_tc || dtype == DT_BF16, "the tcgen05 backend needs FTC_PREC_BF16");
  d = c.embed_dim; heads = c.head_num;
  FTC_REQUIRE(d % 64 == 0 && d <= 1024, "embed_dim must be a multiple of 64, <= 1024");
  FTC_REQUIRE(heads > 0 && d % heads == 0, "embed_dim % head_num");
  hd = d / heads;
  FTC_REQUIRE(hd == 16 || hd == 32 || hd == 64, "head_dim must be 16, 32 or 64");
  kin_pad = (c.enc_input_dim + 7) / 8 * 8;
  const int maxlen = c.max_enc_len > c.max_dec_len ? c.max_enc_len : c.max_dec_len;
  scratch_floats = (size_t)4 * d * d > (size_t)maxlen * d + 64 ? (size_t)4 * d * d : (size_t)maxlen * d + 64;
  scratch_off = walloc(scratch_floats * 4 + (size_t)d / 8 * 4 + 256);

  // ---- encoder ----
  plan_lin(enc_embed, kin_pad, d, 1, false, c.max_enc_len);
  {
    const Lin l = enc_embed; const int kin = c.enc_input_dim, kp = kin_pad, dd = d, rows = c.max_enc_len; const size_t so = scratch_off;
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      const float* w = L.get("encoder.embed.weight", (int64_t)dd * kin);
      const float* pe = L.get("encoder.pos_emb.encoding", (int64_t)rows * dd);
      if (!w || !pe) return -1;
      // pad the weight's K from enc_input_dim to a multiple of 8 (zero columns)
      float* scratch = (float*)(base + so);
      FTC_CHECK_CUDA(cudaMemsetAsync(scratch, 0, (size_t)dd * kp * 4, s));
      FTC_CHECK_CUDA(cudaMemcpy2DAsync(scratch, (size_t)kp * 4, w, (size_t)kin * 4, (size_t)kin * 4, dd, cudaMemcpyDeviceToDevice, s));
      int rc;
      if ((rc = pack_rows(l, base, scratch, dd, 0, 0, s))) return rc;
      if ((rc = upload_ktab(l, base, s))) return rc;
      return cast_f32(base + l.tab_off, dtype, pe, (int64_t)rows * dd, s);
    });
  }
  plan_ln(enc_norm); task_ln(enc_norm, "encoder.norm");
  enc.resize(c.enc_blocks);
  for (int i = 0; i < c.enc_blocks; ++i) {
    EncBlock& b = enc[i];
    const std::string p = "encoder.blocks." + std::to_string(i);
    plan_lin(b.qkv, d, 3 * d, 1, false, c.max_enc_len); task_qkv(b.qkv, p + ".mha", c.max_enc_len);
    plan_lin(b.out, d, d, 1, false, 0); task_plain(b.out, p + ".mha.out_proj.weight", "");
    plan_ln(b.n1); task_ln(b.n1, p + ".norm1");
    plan_lin(b.w1g, d, 4 * d, 1, true, 0); task_w1g(b.w1g, p + ".ff");
    plan_lin(b.w2, 2 * d, d, 1, true, 0); task_plain(b.w2, p + ".ff.w2.weight", p + ".ff.w2.bias");
    plan_ln(b.n2); task_ln(b.n2, p + ".norm2");
  }
  // ---- decoder ----
  for (int i = 0; i < 3; ++i) {
    dec_emb_off[i] = walloc((size_t)MODS[i] * d * 4);
    const size_t off = dec_emb_off[i]; const int m = MODS[i], dd = d; const std::string name = "decoder.embed." + std::to_string(i) + ".weight";
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      const float* e = L.get(name, (int64_t)m * dd);
      if (!e) return -1;
      FTC_CHECK_CUDA(cudaMemcpyAsync(base + off, e, (size_t)m * dd * 4, cudaMemcpyDeviceToDevice, s));
      return 0;
    });
  }
  dec_pos_off = walloc((size_t)c.max_dec_len * d * 4);
  {
    const size_t off = dec_pos_off; const int rows = c.max_dec_len, dd = d;
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      const float* pe = L.get("decoder.pos_emb.encoding", (int64_t)rows * dd);
      if (!pe) return -1;
      FTC_CHECK_CUDA(cudaMemcpyAsync(base + off, pe, (size_t)rows * dd * 4, cudaMemcpyDeviceToDevice, s));
      return 0;
    });
  }
  plan_ln(dec_norm); task_ln(dec_norm, "decoder.norm");
  dec.resize(c.dec_blocks);
  for (int i = 0; i < c.dec_blocks; ++i) {
    DecBlock& b = dec[i];
    const std::string p = "decoder.blocks." + std::to_string(i);
    plan_lin(b.qkv, d, 3 * d, 1, false, c.max_dec_len); task_qkv(b.qkv, p + ".self_attn", c.max_dec_len);
    plan_lin(b.out, d, d, 1, false, 0); task_plain(b.out, p + ".self_attn.out_proj.weight", "");
    plan_ln(b.n1); task_ln(b.n1, p + ".norm1");
    // cross attention: q = (x + Pq) Wq ; k = (enc + Pk) Wk ; v = enc Wv   (pos tables have max_dec_len rows)
    plan_lin(b.cq, d, d, 1, false, c.max_dec_len);
    plan_lin(b.ckv, d, 2 * d, 1, false, c.max_dec_len);
    {
      const Lin lq = b.cq, lkv = b.ckv; const int dd = d, rows = c.max_dec_len; const std::string cp = p + ".cross_attn";
      pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
        const float* wq = L.get(cp + ".q_proj.weight", (int64_t)dd * dd); const float* wk = L.get(cp + ".k_proj.weight", (int64_t)dd * dd);
        const float* wv = L.get(cp + ".v_proj.weight", (int64_t)dd * dd);
        const float* pq = L.get(cp + ".pos_emb_q.encoding", (int64_t)rows * dd); const float* pk = L.get(cp + ".pos_emb_k.encoding", (int64_t)rows * dd);
        if (!wq || !wk || !wv || !pq || !pk) return -1;
        int rc;
        if ((rc = pack_rows(lq, base, wq, dd, 0, 0, s))) return rc;
        if ((rc = upload_ktab(lq, base, s))) return rc;
        if ((rc = fold_pos(lq, base, pq, rows, wq, 0, s))) return rc;
        if ((rc = pack_rows(lkv, base, wk, dd, 0, 0, s))) return rc;
        if ((rc = pack_rows(lkv, base, wv, dd, dd, 0, s))) return rc;
        if ((rc = upload_ktab(lkv, base, s))) return rc;
        FTC_CHECK_CUDA(cudaMemsetAsync(base + lkv.tab_off, 0, (size_t)lkv.tab_rows * lkv.N * esize, s));
        return fold_pos(lkv, base, pk, rows, wk, 0, s);
      });
    }
    plan_lin(b.cout, d, d, 1, false, 0); task_plain(b.cout, p + ".cross_attn.out_proj.weight", "");
    plan_ln(b.n2); task_ln(b.n2, p + ".norm2");
    plan_lin(b.w1g, d, 4 * d, 1, true, 0); task_w1g(b.w1g, p + ".ff");
    plan_lin(b.w2, 2 * d, d, 1, true, 0); task_plain(b.w2, p + ".ff.w2.weight", p + ".ff.w2.bias");
    plan_ln(b.n3); task_ln(b.n3, p + ".norm3");
  }
  // ---- three output heads as one grouped GEMM ----
  plan_lin(heads_out, d, HEAD_LD, 3, true, 0);
  for (int g = 0; g < 3; ++g) heads_out.n_valid[g] = MODS[g];
  {
    const Lin l = heads_out; const int dd = d;
    pack_tasks.push_back([=](const TLookup& L, char* base, cudaStream_t s) -> int {
      FTC_CHECK_CUDA(cudaMemsetAsync(base + l.bias_off, 0, (size_t)3 * HEAD_LD * 4, s));
      for (int g = 0; g < 3; ++g) {
        const std::string n = "decoder.out_layers." + std::to_string(g);
        const float* w = L.get(n + ".weight", (int64_t)MODS[g] * dd); const float* b = L.get(n + ".bias", MODS[g]);
        if (!w || !b) return -1;
        int rc = pack_rows(l, base, w, MODS[g], 0, g, s);
        if (rc) return rc;
        FTC_CHECK_CUDA(cudaMemcpyAsync(base + l.bias_off + (size_t)g * HEAD_LD * 4, b, (size_t)MODS[g] * 4, cudaMemcpyDeviceToDevice, s));
      }
      return upload_ktab(l, base, s);
    });
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
namespace {

struct Workspace {
  char* base; size_t off = 0, cap;
  Workspace(void* p, size_t c) : base((char*)p), cap(c) {}
  void* take(size_t bytes) { void* r = base + off; off += align_up(bytes, 256); return r; }
};

struct Bufs {
  void *enc_in, *xa, *x1, *x2, *tmp, *qkv, *att, *ffh, *enc_out, *crosskv;
  float* keymask; float* logits;
  int64_t *dec_in, *ids, *next_in; float* prob; int* flags;
  size_t cross_stride;   // bytes per decoder layer
};

Bufs carve(const ftc_transformer* t, int B, int Le, int Ld, void* ws, size_t ws_bytes, size_t* used) {
  Workspace w(ws, ws_bytes);
  const size_t es = t->esize, d = t->d;
  const size_t Me = (size_t)B * Le, Md = (size_t)B * Ld, Mx = Me > Md ? Me : Md;
  Bufs b;
  b.enc_in = w.take(Me * t->kin_pad * es);
  b.keymask = (float*)w.take(Me * 4);
  b.xa = w.take(Mx * d * es); b.x1 = w.take(Mx * d * es); b.x2 = w.take(Mx * d * es); b.tmp = w.take(Mx * d * es);
  b.qkv = w.take(Mx * 3 * d * es); b.att = w.take(Mx * d * es); b.ffh = w.take(Mx * 2 * d * es);
  b.enc_out = w.take(Me * d * es);
  b.cross_stride = align_up(Me * 2 * d * es, 256);
  b.crosskv = w.take(b.cross_stride * t->cfg.dec_blocks);
  b.logits = (float*)w.take(Md * 3 * HEAD_LD * 4);
  b.dec_in = (int64_t*)w.take(Md * 8); b.ids = (int64_t*)w.take(Md * 8); b.next_in = (int64_t*)w.take(Md * 8);
  b.prob = (float*)w.take(Md * 4); b.flags = (int*)w.take(256);
  *used = w.off;
  return b;
}

// out[M, ...] = act(A[M,K] W^T + bias) + table[m % mod] + res1 + res2
int run_lin(const ftc_transformer* t, const Lin& l, const void* A, int a_stride, int M, int act, const void* res_full,
            int res_stride, const void* res2, int res2_stride, int table_mod, void* out, int out_stride, int out_layout,
            cudaStream_t s) {
  ConvGemmParams p; memset(&p, 0, sizeof(p));
  char* P = t->packed;
  p.B = M; p.H = 1; p.W = 1; p.Ho = 1; p.Wo = 1; p.stride = 1; p.pad = 0;
  p.M = M; p.N = l.N; p.G = l.G; p.K = l.Kpad;
  p.srcA = A; p.a_pix_stride = a_stride; p.CA = l.K;
  p.ktab = (const uint32_t*)(P + l.ktab_off);
  p.w = P + l.w_off;
  p.scale = nullptr;
  p.bias_tab = l.has_bias ? (const float*)(P + l.bias_off) : nullptr; p.ncase = 1;
  p.act = act;
  if (table_mod > 0) {
    FTC_REQUIRE(l.has_tab && table_mod <= l.tab_rows, "sequence longer than the positional table");
    FTC_REQUIRE(res2 == nullptr, "table + two residuals unsupported");
    p.res1 = P + l.tab_off; p.res1_stride = l.G * l.N; p.res1_row_mod = table_mod;
    p.res2 = res_full; p.res2_stride = res_stride;
  } else {
    p.res1 = res_full; p.res1_stride = res_stride; p.res1_row_mod = 0;
    p.res2 = res2; p.res2_stride = res2_stride;
  }
  p.out = out; p.out_layout = out_layout; p.out_stride = out_stride;
  for (int g = 0; g < l.G; ++g) { p.out_ch_base[g] = g * (act == ACT_SWIGLU ? l.N / 2 : l.N); p.n_valid[g] = l.n_valid[g]; }
  p.dtype = t->dtype;
  p.tc = l.tc;
  return t->use_tc ? conv_gemm_tc(p, s) : conv_gemm_simt(p, s);
}

#define RUN(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

int encode_impl(ftc_transformer* t, const float* enc_input, int B, int Le, const Bufs& b, cudaStream_t s) {
  const int d = t->d, M = B * Le, dt = t->dtype;
  char* P = t->packed;
  RUN(pad_cast_rows(enc_input, b.enc_in, dt, b.keymask, M, t->cfg.enc_input_dim, t->kin_pad, s));
  // embed + pos -> LN
  RUN(run_lin(t, t->enc_embed, b.enc_in, t->kin_pad, M, ACT_NONE, nullptr, 0, nullptr, 0, Le, b.tmp, d, OUT_NHWC, s));
  RUN(layernorm_rows(b.tmp, b.xa, dt, (const float*)(P + t->enc_norm.g_off), (const float*)(P + t->enc_norm.b_off), M, d, LN_EPS, s));
  for (const EncBlock& e : t->enc) {
    RUN(run_lin(t, e.qkv, b.xa, d, M, ACT_NONE, nullptr, 0, nullptr, 0, Le, b.qkv, 3 * d, OUT_NHWC, s));
    RUN(attention(b.qkv, 3 * d, 0, b.qkv, b.qkv, 3 * d, d, 2 * d, b.keymask, b.att, d, dt, B, t->heads, t->hd, Le, Le, s));
    RUN(run_lin(t, e.out, b.att, d, M, ACT_NONE, b.xa, d, nullptr, 0, 0, b.tmp, d, OUT_NHWC, s));
    RUN(layernorm_rows(b.tmp, b.x1, dt, (const float*)(P + e.n1.g_off), (const float*)(P + e.n1.b_off), M, d, LN_EPS, s));
    RUN(run_lin(t, e.w1g, b.x1, d, M, ACT_SWIGLU, nullptr, 0, nullptr, 0, 0, b.ffh, 2 * d, OUT_NHWC, s));
    RUN(run_lin(t, e.w2, b.ffh, 2 * d, M, ACT_NONE, b.x1, d, b.xa, d, 0, b.tmp, d, OUT_NHWC, s));
    RUN(layernorm_rows(b.tmp, b.xa, dt, (const float*)(P + e.n2.g_off), (const float*)(P + e.n2.b_off), M, d, LN_EPS, s));
  }
  FTC_CHECK_CUDA(cudaMemcpyAsync(b.enc_out, b.xa, (size_t)M * d * t->esize, cudaMemcpyDeviceToDevice, s));
  // cross-attention K/V of every decoder layer, once per encoder pass
  for (size_t i = 0; i < t->dec.size(); ++i)
    RUN(run_lin(t, t->dec[i].ckv, b.enc_out, d, M, ACT_NONE, nullptr, 0, nullptr, 0, Le, (char*)b.crosskv + i * b.cross_stride,
                2 * d, OUT_NHWC, s));
  t->enc_B = B; t->enc_L = Le;
  return 0;
}

int decode_impl(ftc_transformer* t, const int64_t* tokens, int B, int Ld, int Le, const Bufs& b, float* logits, cudaStream_t s) {
  const int d = t->d, M = B * Ld, dt = t->dtype;
  char* P = t->packed;
  RUN(decoder_embed_ln(tokens, (const float*)(P + t->dec_emb_off[0]), (const float*)(P + t->dec_emb_off[1]),
                       (const float*)(P + t->dec_emb_off[2]), MODS[0], MODS[1], MODS[2], (const float*)(P + t->dec_pos_off),
                       (const float*)(P + t->dec_norm.g_off), (const float*)(P + t->dec_norm.b_off), b.xa, dt, M, Ld, d, LN_EPS, s));
  for (size_t i = 0; i < t->dec.size(); ++i) {
    const DecBlock& e = t->dec[i];
    const void* ckv = (char*)b.crosskv + i * b.cross_stride;
    RUN(run_lin(t, e.qkv, b.xa, d, M, ACT_NONE, nullptr, 0, nullptr, 0, Ld, b.qkv, 3 * d, OUT_NHWC, s));
    RUN(attention(b.qkv, 3 * d, 0, b.qkv, b.qkv, 3 * d, d, 2 * d, nullptr, b.att, d, dt, B, t->heads, t->hd, Ld, Ld, s));
    RUN(run_lin(t, e.out, b.att, d, M, ACT_NONE, b.xa, d, nullptr, 0, 0, b.tmp, d, OUT_NHWC, s));
    RUN(layernorm_rows(b.tmp, b.x1, dt, (const float*)(P + e.n1.g_off), (const float*)(P + e.n1.b_off), M, d, LN_EPS, s));
    RUN(run_lin(t, e.cq, b.x1, d, M, ACT_NONE, nullptr, 0, nullptr, 0, Ld, b.qkv, d, OUT_NHWC, s));
    RUN(attention(b.qkv, d, 0, ckv, ckv, 2 * d, 0, d, b.keymask, b.att, d, dt, B, t->heads, t->hd, Ld, Le, s));
    RUN(run_lin(t, e.cout, b.att, d, M, ACT_NONE, b.x1, d, nullptr, 0, 0, b.tmp, d, OUT_NHWC, s));
    RUN(layernorm_rows(b.tmp, b.x2, dt, (const float*)(P + e.n2.g_off), (const float*)(P + e.n2.b_off), M, d, LN_EPS, s));
    RUN(run_lin(t, e.w1g, b.x2, d, M, ACT_SWIGLU, nullptr, 0, nullptr, 0, 0, b.ffh, 2 * d, OUT_NHWC, s));
    RUN(run_lin(t, e.w2, b.ffh, 2 * d, M, ACT_NONE, b.x2, d, b.xa, d, 0, b.tmp, d, OUT_NHWC, s));
    RUN(layernorm_rows(b.tmp, b.xa, dt, (const float*)(P + e.n3.g_off), (const float*)(P + e.n3.b_off), M, d, LN_EPS, s));
  }
  RUN(run_lin(t, t->heads_out, b.xa, d, M, ACT_NONE, nullptr, 0, nullptr, 0, 0, logits, 3 * HEAD_LD, OUT_NHWC_F32, s));
  return 0;
}

}  // namespace

extern "C" {

int ftc_transformer_create(const ftc_transformer_config* cfg, ftc_transformer** out) {
  FTC_REQUIRE(cfg && out, "null argument");
  ftc_transformer* t = new ftc_transformer();
  t->cfg = *cfg;
  int rc = t->build();
  if (rc) { delete t; return rc; }
  *out = t;
  return 0;
}

void ftc_transformer_destroy(ftc_transformer* t) { delete t; }
size_t ftc_transformer_weight_bytes(const ftc_transformer* t) { return t ? t->weight_bytes : 0; }
int ftc_transformer_logit_stride(void) { return 3 * HEAD_LD; }
int ftc_transformer_head_stride(void) { return HEAD_LD; }

size_t ftc_transformer_workspace_bytes(const ftc_transformer* t, int batch, int enc_len, int dec_len) {
  if (!t) return 0;
  size_t used = 0;
  carve(t, batch, enc_len, dec_len, nullptr, 0, &used);
  return used + 256;
}

int ftc_transformer_pack_weights(ftc_transformer* t, int n, const char* const* names, const void* const* ptrs,
                                 const int64_t* numels, void* packed, size_t packed_bytes, void* stream) {
  FTC_REQUIRE(t && names && ptrs && numels && packed, "null argument");
  FTC_REQUIRE(packed_bytes >= t->weight_bytes, "packed buffer too small");
  cudaStream_t s = (cudaStream_t)stream;
  TLookup L;
  for (int i = 0; i < n; ++i) L.t[names[i]] = {(const float*)ptrs[i], numels[i]};
  FTC_CHECK_CUDA(cudaMemsetAsync(packed, 0, t->weight_bytes, s));
  for (auto& task : t->pack_tasks) {
    int rc = task(L, (char*)packed, s);
    if (rc) return rc;
  }
  FTC_CHECK_CUDA(cudaStreamSynchronize(s));
  t->packed = (char*)packed;
  return 0;
}

int ftc_transformer_forward(ftc_transformer* t, const float* enc_input, const int64_t* dec_input, int batch, int enc_len,
                            int dec_len, float* logits, void* workspace, size_t workspace_bytes, void* stream) {
  FTC_REQUIRE(t && enc_input && dec_input && logits && workspace && batch > 0, "bad argument");
  FTC_REQUIRE(t->packed, "ftc_transformer_pack_weights must be called first");
  FTC_REQUIRE(enc_len <= t->cfg.max_enc_len && dec_len <= t->cfg.max_dec_len && enc_len <= t->cfg.max_dec_len,
              "sequence longer than the positional tables");
  FTC_REQUIRE(workspace_bytes >= ftc_transformer_workspace_bytes(t, batch, enc_len, dec_len), "workspace too small");
  size_t used;
  Bufs b = carve(t, batch, enc_len, dec_len, workspace, workspace_bytes, &used);
  cudaStream_t s = (cudaStream_t)stream;
  RUN(encode_impl(t, enc_input, batch, enc_len, b, s));
  return decode_impl(t, dec_input, batch, dec_len, enc_len, b, logits, s);
}

int ftc_transformer_predict(ftc_transformer* t, const float* enc_input, int batch, int enc_len, int dec_len, int64_t* out_ids,
                            int max_passes, int* passes_run, int* stop_reason, void* workspace, size_t workspace_bytes,
                            void* stream) {
  FTC_REQUIRE(t && enc_input && out_ids && workspace && batch > 0 && max_passes > 0, "bad argument");
  FTC_REQUIRE(t->packed, "ftc_transformer_pack_weights must be called first");
  FTC_REQUIRE(enc_len <= t->cfg.max_enc_len && dec_len <= t->cfg.max_dec_len && enc_len <= t->cfg.max_dec_len,
              "sequence longer than the positional tables");
  FTC_REQUIRE(workspace_bytes >= ftc_transformer_workspace_bytes(t, batch, enc_len, dec_len), "workspace too small");
  size_t used;
  Bufs b = carve(t, batch, enc_len, dec_len, workspace, workspace_bytes, &used);
  cudaStream_t s = (cudaStream_t)stream;
  const int M = batch * dec_len;
  RUN(encode_impl(t, enc_input, batch, enc_len, b, s));
  // decoder_input[:, :] = MSK (models/transformer.py:278-279)
  {
    std::vector<int64_t> init((size_t)M, 3);
    FTC_CHECK_CUDA(cudaMemcpyAsync(b.dec_in, init.data(), (size_t)M * 8, cudaMemcpyHostToDevice, s));
    FTC_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  int reason = 0, k = 0;
  for (k = 0; k < max_passes; ++k) {
    FTC_CHECK_CUDA(cudaMemsetAsync(b.flags, 0, 8, s));
    RUN(decode_impl(t, b.dec_in, batch, dec_len, enc_len, b, b.logits, s));
    RUN(mask_predict_step(b.logits, 3 * HEAD_LD, HEAD_LD, b.dec_in, b.ids, b.prob, b.next_in, b.flags, M, s));
    int flags[2] = {0, 0};
    FTC_CHECK_CUDA(cudaMemcpyAsync(flags, b.flags, 8, cudaMemcpyDeviceToHost, s));
    FTC_CHECK_CUDA(cudaStreamSynchronize(s));      // the reference takes the same data-dependent exits (:326, :356)
    if (flags[0] == 0) { reason = 1; break; }      // "[k early stop]"
    if (k < max_passes - 1) {
      if (flags[1] == 0) { reason = 2; break; }    // "[k no remask stop]"
      FTC_CHECK_CUDA(cudaMemcpyAsync(b.dec_in, b.next_in, (size_t)M * 8, cudaMemcpyDeviceToDevice, s));
    }
  }
  FTC_CHECK_CUDA(cudaMemcpyAsync(out_ids, b.ids, (size_t)M * 8, cudaMemcpyDeviceToDevice, s));
  if (passes_run) *passes_run = k < max_passes ? k + 1 : max_passes;
  if (stop_reason) *stop_reason = reason;
  return 0;
}

// TransformerPredictor.forward for a batch of INDEPENDENT sequences: every sequence follows the stop rules the reference applies
// to a batch of one (models/transformer.py:326, :356 test torch.all / torch.any over whatever batch they are given, so a batched
// reference call couples its sequences; process_ocr_base.py:235 always calls it with one chunk).  The chunks of a page decoded
// here in one call therefore give exactly the code points of the reference's chunk-by-chunk loop.  seq_state: device int32
// [3 * batch] = {done, passes, stop reason} per sequence on return; scratch_i32: device int32 [2 * batch + 4].
int ftc_transformer_predict_each(ftc_transformer* t, const float* enc_input, int batch, int enc_len, int dec_len, int64_t* out_ids,
                                 int max_passes, int* seq_state, int* scratch_i32, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  FTC_REQUIRE(t && enc_input && out_ids && seq_state && scratch_i32 && workspace && batch > 0 && max_passes > 0, "bad argument");
  FTC_REQUIRE(t->packed, "ftc_transformer_pack_weights must be called first");
  FTC_REQUIRE(enc_len <= t->cfg.max_enc_len && dec_len <= t->cfg.max_dec_len && enc_len <= t->cfg.max_dec_len,
              "sequence longer than the positional tables");
  FTC_REQUIRE(workspace_bytes >= ftc_transformer_workspace_bytes(t, batch, enc_len, dec_len), "workspace too small");
  size_t used;
  Bufs b = carve(t, batch, enc_len, dec_len, workspace, workspace_bytes, &used);
  cudaStream_t s = (cudaStream_t)stream;
  const int M = batch * dec_len;
  RUN(encode_impl(t, enc_input, batch, enc_len, b, s));
  {
    std::vector<int64_t> init((size_t)M, 3);
    FTC_CHECK_CUDA(cudaMemcpyAsync(b.dec_in, init.data(), (size_t)M * 8, cudaMemcpyHostToDevice, s));
    FTC_CHECK_CUDA(cudaStreamSynchronize(s));
  }
  int* seq_flags = scratch_i32;
  int* n_running = scratch_i32 + 2 * batch;
  FTC_CHECK_CUDA(cudaMemsetAsync(scratch_i32, 0, sizeof(int) * (2 * (size_t)batch + 4), s));
  FTC_CHECK_CUDA(cudaMemsetAsync(seq_state, 0, sizeof(int) * 3 * (size_t)batch, s));
  for (int k = 0; k < max_passes; ++k) {
    FTC_CHECK_CUDA(cudaMemsetAsync(b.flags, 0, 8, s));
    FTC_CHECK_CUDA(cudaMemsetAsync(n_running, 0, sizeof(int), s));
    RUN(decode_impl(t, b.dec_in, batch, dec_len, enc_len, b, b.logits, s));
    RUN(mask_predict_step(b.logits, 3 * HEAD_LD, HEAD_LD, b.dec_in, b.ids, b.prob, b.next_in, b.flags, M, s, dec_len, seq_flags));
    RUN(mask_predict_advance(seq_flags, seq_state, b.dec_in, b.next_in, b.ids, out_ids, batch, dec_len, k, max_passes - 1, n_running, s));
    int running = 0;
    FTC_CHECK_CUDA(cudaMemcpyAsync(&running, n_running, sizeof(int), cudaMemcpyDeviceToHost, s));
    FTC_CHECK_CUDA(cudaStreamSynchronize(s));
    if (running == 0) break;
  }
  return 0;
}

}  // extern "C"
