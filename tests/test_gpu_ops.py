"""GPU: single-op C-ABI entry points vs plain PyTorch fp32 references of the same op.
Tolerances: fp32 SIMT path 1e-4 rel-L2 (accumulation order only); bf16 paths 1e-2 rel-L2 (operands are rounded to
bf16 on both sides, the reference accumulates in fp32, the output is rounded to bf16 once)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _ops():
    from findtextcenternet_b200 import _lib, _ops
    return _lib, _ops


def _act(y, act):
    if act == 1:
        return F.silu(y)
    if act == 2:
        return F.gelu(y)
    return y


CONV_CASES = [
    # B, H, W, Cin, Cout, k, stride, act, residual
    (2, 24, 24, 64, 96, 3, 1, 1, False),
    (1, 17, 13, 32, 32, 3, 1, 1, True),       # ragged M (221 rows), Cin=32 (4 chunks per tap)
    (2, 32, 32, 32, 128, 3, 2, 1, False),     # stride 2
    (1, 24, 24, 96, 384, 3, 1, 1, False),     # Cin=96: k-blocks straddle taps; N=384 -> 2 x 192 tiles
    (2, 24, 24, 256, 64, 1, 1, 0, True),      # 1x1 project + residual
    (1, 12, 12, 640, 1280, 1, 1, 1, False),   # N=1280 -> 5 x 256
    (1, 48, 48, 192, 16, 3, 1, 0, False),     # tiny N
    (3, 10, 10, 8, 24, 3, 1, 2, False),       # K=72 -> padded k-block, GELU
    (4, 96, 96, 128, 192, 3, 1, 2, True),     # M=36864, K=1152: 256-row tile path (two accumulators per B stage)
    (3, 111, 100, 128, 96, 3, 1, 1, False),   # M=33300: ragged 256-row tiles
    (2, 130, 128, 768, 320, 1, 1, 0, True),   # 1x1, K=768, N=320 -> 2 x 160, 256-row tiles
    (2, 32, 48, 96, 192, 3, 1, 2, True),      # TMA halo tiles (W % 16 == 0), Cin=96 -> second chunk half out of bounds
    (3, 16, 16, 64, 100, 3, 1, 0, False),     # halo, one tile per image, N=100 -> 112
    (2, 128, 128, 128, 96, 3, 1, 1, True),    # halo, M=32768: 16x16 tiles (two accumulators), double-buffered TMEM
    (1, 33, 7, 96, 40, 1, 1, 1, False),       # TMA rows: ragged M=231, Cin=96
    (2, 32, 48, 32, 32, 3, 1, 1, True),       # halo, Cin=32: 32-wide k-blocks (SWIZZLE_64B operands), residual, weight-stationary off (few tiles)
    (3, 64, 64, 32, 64, 3, 1, 1, False),      # halo, Cin=32, two accumulators
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("mode", ["simt_f32", "simt_bf16", "tc_bf16", "tc_im2col_bf16"])
def test_conv2d(case, mode):
    _lib, ops = _ops()
    b, h, w, cin, cout, k, stride, act, use_res = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn(b, h, w, cin, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / np.sqrt(cin * k * k)
    scale = 0.5 + torch.rand(cout, generator=g)
    bias = 0.3 * torch.randn(cout, generator=g)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    res = torch.randn(b, ho, wo, cout, generator=g) if use_res else None
    dt = torch.float32 if mode == "simt_f32" else torch.bfloat16
    xq, wq = x.to(dt).float(), (wt.to(dt).float() if dt == torch.bfloat16 else wt)
    resq = None if res is None else res.to(dt).float()
    ref = F.conv2d(xq.permute(0, 3, 1, 2), wq, None, stride, (k - 1) // 2)
    ref = ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    ref = _act(ref, act)                      # residual joins after the activation (FusedMBConv semantics)
    if resq is not None:
        ref = ref + resq.permute(0, 3, 1, 2)
    ref = ref.permute(0, 2, 3, 1)
    backend = {"tc_bf16": _lib.GEMM_TCGEN05, "tc_im2col_bf16": _lib.GEMM_TCGEN05_IM2COL}.get(mode, _lib.GEMM_SIMT)
    out = ops.conv2d(x.to(dt).cuda(), wt, stride, scale, bias, act, None if res is None else res.to(dt).cuda(),
                     None, backend)
    torch.cuda.synchronize()
    tol = 1e-4 if mode == "simt_f32" else 1e-2
    assert out.shape == ref.shape
    assert rel_l2(out.float().cpu().numpy(), ref.numpy()) < tol


@pytest.mark.parametrize("shape", [(3, 12, 12, 384, 96), (8, 64, 64, 768, 128)])   # 2nd: M=32768, 12 k-blocks -> 256-row tiles
@pytest.mark.parametrize("mode", ["simt_f32", "tc_bf16", "tc_im2col_bf16"])
def test_conv2d_se_scaled_operand(mode, shape):
    """1x1 project conv with the SE excitation applied to the A operand (torchvision efficientnet.py MBConv:
    block[2] scale then block[3] conv)."""
    _lib, ops = _ops()
    g = torch.Generator().manual_seed(3)
    b, h, w, cin, cout = shape
    dt = torch.float32 if mode == "simt_f32" else torch.bfloat16
    x = torch.randn(b, h, w, cin, generator=g)
    wt = torch.randn(cout, cin, 1, 1, generator=g) / np.sqrt(cin)
    a_scale = torch.rand(b, cin, generator=g)
    xs = x.to(dt).float() * a_scale.view(b, 1, 1, cin)
    if dt == torch.bfloat16:
        xs = xs.to(dt).float()
        wq = wt.to(dt).float()
    else:
        wq = wt
    ref = F.conv2d(xs.permute(0, 3, 1, 2), wq).permute(0, 2, 3, 1)
    backend = {"tc_bf16": _lib.GEMM_TCGEN05, "tc_im2col_bf16": _lib.GEMM_TCGEN05_IM2COL}.get(mode, _lib.GEMM_SIMT)
    out = ops.conv2d(x.to(dt).cuda(), wt, 1, None, None, 0, None, a_scale.cuda(), backend)
    assert rel_l2(out.float().cpu().numpy(), ref.numpy()) < (1e-4 if mode == "simt_f32" else 1e-2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("stride", [1, 2])
def test_dwconv_se(dt, stride):
    _lib, ops = _ops()
    g = torch.Generator().manual_seed(5)
    b, h, w, c = 2, 24, 24, 384
    x = torch.randn(b, h, w, c, generator=g).to(dt)
    wt = torch.randn(c, 1, 3, 3, generator=g) / 3
    scale = 0.5 + torch.rand(c, generator=g)
    bias = 0.2 * torch.randn(c, generator=g)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, stride, 1, 1, c)
    ref = F.silu(ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    out, se = ops.dwconv3x3(x.cuda(), wt.view(c, 9).t().contiguous().cuda(), scale.cuda(), bias.cuda(), stride, True)
    out2, se2 = ops.dwconv3x3(x.cuda(), wt.view(c, 9).t().contiguous().cuda(), scale.cuda(), bias.cuda(), stride, True)
    assert torch.equal(se, se2) and torch.equal(out, out2)       # no atomics: bit-reproducible squeeze
    tol = 1e-5 if dt == torch.float32 else 1e-2
    assert rel_l2(out.float().cpu().numpy(), ref.permute(0, 2, 3, 1).numpy()) < tol
    assert rel_l2(se.sum(1).cpu().numpy(), ref.sum((2, 3)).numpy()) < (1e-4 if dt == torch.float32 else 2e-2)
    # SE excitation (ops/misc.py:251-261)
    s = 96
    w1 = torch.randn(s, c, generator=g) / np.sqrt(c); b1 = 0.1 * torch.randn(s, generator=g)
    w2 = torch.randn(c, s, generator=g) / np.sqrt(s); b2 = 0.1 * torch.randn(c, generator=g)
    mean = ref.mean((2, 3))
    ref_scale = torch.sigmoid(F.linear(F.silu(F.linear(mean, w1, b1)), w2, b2))
    ho = ref.shape[2]
    got = ops.se_fc(se, 1.0 / (ho * ho), w1.cuda(), b1.cuda(), w2.t().contiguous().cuda(), b2.cuda())
    assert rel_l2(got.cpu().numpy(), ref_scale.numpy()) < (1e-4 if dt == torch.float32 else 1e-2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 48, 48, 192), (3, 24, 24, 96), (1, 16, 12, 32)])
def test_dwconv_se_folded(dt, shape):
    """Strip depthwise kernel with folded SE squeeze + fc1, then fc2 (torchvision MBConv middle, stride 1)."""
    _lib, ops = _ops()
    g = torch.Generator().manual_seed(15)
    b, h, w, c = shape
    s = max(c // 16, 4)
    x = torch.randn(b, h, w, c, generator=g).to(dt)
    wt = torch.randn(c, 1, 3, 3, generator=g) / 3
    scale = 0.5 + torch.rand(c, generator=g)
    bias = 0.2 * torch.randn(c, generator=g)
    w1 = torch.randn(s, c, generator=g) / np.sqrt(c); b1 = 0.1 * torch.randn(s, generator=g)
    w2 = torch.randn(c, s, generator=g) / np.sqrt(s); b2 = 0.1 * torch.randn(c, generator=g)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, 1, 1, 1, c)
    ref = F.silu(ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1))
    ref_scale = torch.sigmoid(F.linear(F.silu(F.linear(ref.mean((2, 3)), w1, b1)), w2, b2))
    out, sc = ops.dwconv3x3_se(x.cuda(), wt.view(c, 9).t().contiguous().cuda(), scale.cuda(), bias.cuda(), w1.cuda(), b1.cuda(),
                               w2.t().contiguous().cuda(), b2.cuda())
    tol = 1e-5 if dt == torch.float32 else 1e-2
    assert rel_l2(out.float().cpu().numpy(), ref.permute(0, 2, 3, 1).numpy()) < tol
    assert rel_l2(sc.cpu().numpy(), ref_scale.numpy()) < (1e-4 if dt == torch.float32 else 1e-2)
    out2, sc2 = ops.dwconv3x3_se(x.cuda(), wt.view(c, 9).t().contiguous().cuda(), scale.cuda(), bias.cuda(), w1.cuda(), b1.cuda(),
                                 w2.t().contiguous().cuda(), b2.cuda())
    assert torch.equal(sc, sc2) and torch.equal(out, out2)       # fixed-order fc1 reduce: bit-reproducible excitation


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_head_top_conv(dt):
    """Leafmap.top_conv of the small heads (models/detector.py:188-190) against F.conv2d."""
    _lib, ops = _ops()
    g = torch.Generator().manual_seed(16)
    od = [1, 2, 1]
    b, h, w = 2, 24, 40
    y = torch.randn(b, h, w, len(od) * 192, generator=g).to(dt)
    ws = [torch.randn(o, 192, 3, 3, generator=g) / 40 for o in od]
    bs = [0.1 * torch.randn(o, generator=g) for o in od]
    refs = []
    for i, (wt, bb) in enumerate(zip(ws, bs)):
        yi = y[..., i * 192:(i + 1) * 192].float().permute(0, 3, 1, 2)
        wq = wt.to(dt).float() if dt == torch.bfloat16 else wt
        refs.append(F.conv2d(yi, wq, bb, 1, 1))
    ref = torch.cat(refs, 1)
    rows = torch.cat([wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1) for wt in ws], 0).contiguous()   # [rows][tap][c]
    out = ops.head_top_conv(y.cuda(), od, rows.cuda(), torch.cat(bs).cuda())
    assert rel_l2(out.cpu().numpy(), ref.numpy()) < (1e-5 if dt == torch.float32 else 5e-3)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cfg", [(3, 100, 100, 16, 32), (2, 100, 37, 16, 32), (2, 150, 400, 12, 64), (1, 17, 65, 4, 64),
                                 (2, 128, 128, 16, 32), (5, 1, 16, 2, 32), (40, 100, 100, 16, 32), (21, 128, 100, 12, 64)])
def test_attention_matches_sdpa(dt, cfg):
    """F.scaled_dot_product_attention with an additive key-pad mask (models/transformer.py:126-134).  bf16 with both sequence
    lengths <= 128 runs the tcgen05 kernel (csrc/attention_tc.cu: hd 32 = two heads per work item, hd 64 = one; 320 / 252 work
    items = several per persistent CTA in the last two cases); longer sequences the mma.sync kernel; fp32 the CUDA-core kernel."""
    _lib, ops = _ops()
    b, lt, ls, heads, hd = cfg
    g = torch.Generator().manual_seed(17)
    d = heads * hd
    q = torch.randn(b, lt, d, generator=g).to(dt); k = torch.randn(b, ls, d, generator=g).to(dt); v = torch.randn(b, ls, d, generator=g).to(dt)
    mask = torch.zeros(b, ls)
    for i in range(b):
        mask[i, max(1, ls - 1 - 3 * (i % 8)):] = float("-inf")
    def split(t, l):
        return t.float().view(b, l, heads, hd).transpose(1, 2)
    ref = F.scaled_dot_product_attention(split(q, lt), split(k, ls), split(v, ls), attn_mask=mask.view(b, 1, 1, ls))
    ref = ref.transpose(1, 2).reshape(b, lt, d)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda(), heads, mask.cuda())
    assert rel_l2(out.float().cpu().numpy(), ref.numpy()) < (1e-5 if dt == torch.float32 else 1e-2)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_upsample2x_align_corners(dt):
    _lib, ops = _ops()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 24, 24, 192, generator=g).to(dt)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=True)
    out = ops.upsample2x(x.cuda())
    assert rel_l2(out.float().cpu().numpy(), ref.permute(0, 2, 3, 1).numpy()) < (1e-5 if dt == torch.float32 else 5e-3)


def test_peak_pick_and_decode_match_oracle(golden_detector):
    """Bit-exact peak index set; decoded boxes vs the numpy oracle of process_ocr_base.py:498-538."""
    from findtextcenternet_b200 import _ops as ops
    from findtextcenternet_b200.engine import peak_decode
    from oracle import detector_oracle as DO
    h10 = golden_detector["rand0_heatmap10"]
    heat9 = torch.from_numpy(np.concatenate([h10[0:1], h10[2:]], 0)[None].copy()).cuda()
    got10 = ops.peak_pick(heat9).cpu().numpy()[0]
    assert np.array_equal(got10, h10)          # includes the -inf pattern
    # plateau ties and borders: quantised random map
    g = torch.Generator().manual_seed(9)
    q = torch.round(torch.randn(2, 9, 64, 48, generator=g) * 2) / 2
    got = ops.peak_pick(q.cuda()).cpu()
    ref = DO.peak_pick(q)
    assert torch.equal(got, ref)
    # decode: synthetic feature map so the gather is checked too
    feat = torch.randn(1, 100, 192, 192, generator=g)
    for (x_i, y_i, pw, ph) in [(0, 0, 768, 768), (460, 920, 2148, 2148)]:
        mask = DO.tile_mask(x_i, y_i, pw, ph)
        ys, xs = np.nonzero(mask)
        meta = torch.tensor([[x_i, y_i, xs.min(), xs.max() + 1, ys.min(), ys.max() + 1]], dtype=torch.int32).cuda()
        count, loc, gf, total = peak_decode(heat9, feat.cuda(), meta, pw, ph, 0.4, 4096)
        n = int(count[0])
        assert int(total[0]) == n
        ref_loc, ref_gf = DO.decode_tile(h10, feat[0].numpy(), x_i, y_i, pw, ph)
        assert n == len(ref_loc) and n > 10
        loc = loc[0, :n].cpu().numpy(); gf = gf[0, :n].cpu().numpy()
        # identical peak SET (the reference's argsort is unstable, so order among equal scores is unspecified);
        # device order must be non-increasing in score
        assert np.all(np.diff(loc[:, 0]) <= 0)
        key = lambda l: (int(l[1]), int(l[2]))
        ref_by = {key(l): i for i, l in enumerate(ref_loc)}
        assert sorted(ref_by) == sorted(key(l) for l in loc)
        perm = np.array([ref_by[key(l)] for l in loc])
        np.testing.assert_allclose(loc, ref_loc[perm], rtol=2e-5, atol=1e-6)
        assert np.array_equal(gf, ref_gf[perm])


@pytest.mark.parametrize("max_peaks", [1024, 2048, 4096])
def test_peak_decode_overflow_keeps_top_scores(max_peaks):
    """A dense tile (every other pixel a local maximum: 9 216 peaks, more than any max_peaks) must return the max_peaks
    HIGHEST-scoring peaks of the reference loop (process_ocr_base.py:519-538 has no cap), deterministically, and report the
    uncapped total."""
    from findtextcenternet_b200.engine import peak_decode
    from oracle import detector_oracle as DO
    g = torch.Generator().manual_seed(31)
    heat9 = torch.full((2, 9, 192, 192), -4.0)
    heat9[:, 0, ::2, ::2] = torch.rand(2, 96, 96, generator=g) * 4.0       # isolated maxima, sigmoid in [0.5, 0.98]
    heat9[:, 1:3] = 0.3 * torch.randn(2, 2, 192, 192, generator=g)          # w, h ~ 50 px
    heat9[:, 3:] = torch.randn(2, 6, 192, 192, generator=g)
    feat = torch.randn(2, 100, 192, 192, generator=g)
    meta = torch.tensor([[0, 0, 0, 192, 0, 192], [0, 0, 0, 192, 0, 192]], dtype=torch.int32).cuda()
    runs = [peak_decode(heat9.cuda(), feat.cuda(), meta, 768, 768, 0.4, max_peaks) for _ in range(2)]
    for a, b in zip(runs[0], runs[1]):
        assert torch.equal(a, b)                                            # run-to-run identical, overflow included
    count, loc, gf, total = (t.cpu().numpy() for t in runs[0])
    h10 = DO.peak_pick(heat9).numpy()
    for b in range(2):
        ref_loc, ref_gf = DO.decode_tile(h10[b], feat[b].numpy(), 0, 0, 768, 768)
        assert total[b] == len(ref_loc) == 96 * 96 and count[b] == max_peaks
        got = loc[b, :max_peaks]
        assert np.all(np.diff(got[:, 0]) <= 0)
        # the kept set is the reference's top max_peaks by score (scores within one fp32 ulp of the cut may fall either side:
        # tanhf on the device vs numpy's tanh)
        key = lambda l: (int(l[1]), int(l[2]))
        ref_by = {key(l): i for i, l in enumerate(ref_loc)}
        cut = ref_loc[max_peaks - 1, 0]
        assert all(key(l) in ref_by for l in got) and len({key(l) for l in got}) == max_peaks
        assert got[:, 0].min() >= cut - 1e-6
        assert {key(l) for l in ref_loc if l[0] > cut + 1e-6} <= {key(l) for l in got}
        perm = np.array([ref_by[key(l)] for l in got])
        np.testing.assert_allclose(got, ref_loc[perm], rtol=2e-5, atol=1e-6)
        assert np.array_equal(gf[b, :max_peaks], ref_gf[perm])
