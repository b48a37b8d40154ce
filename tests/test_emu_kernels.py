"""CPU: the REAL kernel source of csrc/train_ops.cu / csrc/loss_ops.cu executed on host threads through the CUDA-on-CPU shim
(oracle/emu: one OS thread per CUDA thread, barriers for __syncthreads / __syncwarp / shuffles) and compared with
oracle/train_oracle.py.  This checks what a compile cannot: index arithmetic, shared-memory reductions, warp shuffles,
split / atomic accumulation, launch geometry -- for the kernels that had no GPU run yet as much as for those that had.
Shapes are tiny (every CUDA thread is an OS thread)."""
import ctypes as C
import os
import sys

import pytest
import torch

from conftest import rel_l2
from oracle import train_oracle as TO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = {torch.float32: 0, torch.bfloat16: 1}


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, os.path.join(ROOT, "oracle", "emu"))
    import build_emu
    lib = C.CDLL(build_emu.build())
    lib.ftc_last_error.restype = C.c_char_p
    lib.ftc_train_reduce_scratch_bytes.restype = C.c_size_t
    lib.ftc_train_reduce_scratch_bytes.argtypes = [C.c_int64, C.c_int]
    lib.ftc_train_attention_bwd_scratch_bytes.restype = C.c_size_t
    return lib


def P(t):
    return C.c_void_p(None if t is None else t.data_ptr())


def ok(lib, rc):
    assert rc == 0, lib.ftc_last_error()


def rnd(*shape, seed=0, scale=1.0, dt=torch.float32):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).to(dt)


def scratch(lib, rows, c):
    return torch.empty(lib.ftc_train_reduce_scratch_bytes(rows, c) // 4)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,c,act", [(37, 5, 1), (600, 70, 2), (9, 130, 0), (50, 72, 1), (300, 200, 2), (5, 8, 0), (2, 64, 1),
                                        # bf16 with C % 32 == 0 takes the stream kernels (4 row groups in flight, 4 / 8 chunk lanes)
                                        (300, 96, 1), (700, 128, 2), (1100, 32, 0), (513, 64, 1)])
def test_emu_batchnorm_kernels(emu, dt, rows, c, act):
    x = (rnd(rows, c, seed=1) * 1.7 + 0.3).to(dt)
    gamma, beta, res, dy = rnd(c, seed=2), rnd(c, seed=3), rnd(rows, c, seed=4, dt=dt), rnd(rows, c, seed=5, dt=dt)
    mean, var, sc = torch.empty(c), torch.empty(c), scratch(emu, rows, c)
    ok(emu, emu.ftc_train_bn_stats(P(x), DT[dt], C.c_int64(rows), c, P(mean), P(var), P(sc), None))
    m0, v0 = TO.bn_stats(x.float())
    assert rel_l2(mean, m0) < 1e-5 and rel_l2(var, v0) < 1e-4
    y = torch.empty_like(x)
    ok(emu, emu.ftc_train_bn_act(P(x), P(y), DT[dt], C.c_int64(rows), c, P(m0), P(v0), P(gamma), P(beta), C.c_float(1e-3), act, P(res), None))
    tol = 2e-5 if dt == torch.float32 else 2e-2
    assert rel_l2(y.float(), TO.bn_act(x.float(), m0, v0, gamma, beta, 1e-3, act, res.float())) < tol
    dx, dbeta, dgamma = torch.empty_like(x), torch.empty(c), torch.empty(c)
    ok(emu, emu.ftc_train_bn_act_bwd(P(x), P(dy), P(dx), DT[dt], C.c_int64(rows), c, P(m0), P(v0), P(gamma), P(beta), C.c_float(1e-3), act,
                                     P(dbeta), P(dgamma), P(sc), None))
    dx0, dg0, db0 = TO.bn_act_bwd(x.float(), dy.float(), m0, v0, gamma, beta, 1e-3, act)
    gtol = 1e-4 if (dt == torch.float32 or c % 32) else 5e-4     # stream kernels: 1-SFU activation forms (tanh fit of erf: 2.5e-5 abs)
    assert rel_l2(dgamma, dg0) < gtol and rel_l2(dbeta, db0) < gtol and rel_l2(dx.float(), dx0) < max(tol, 1e-4)
    if c % 32 == 0:                                               # no-residual instantiation
        ok(emu, emu.ftc_train_bn_act(P(x), P(y), DT[dt], C.c_int64(rows), c, P(m0), P(v0), P(gamma), P(beta), C.c_float(1e-3), act, None, None))
        assert rel_l2(y.float(), TO.bn_act(x.float(), m0, v0, gamma, beta, 1e-3, act, None)) < tol


@pytest.mark.parametrize("b,h,w,cin,cout,k,stride", [(2, 6, 5, 8, 16, 3, 1), (1, 7, 6, 3, 70, 3, 2), (2, 4, 4, 72, 5, 1, 1)])
def test_emu_conv_gradients(emu, b, h, w, cin, cout, k, stride):
    x, wt = rnd(b, h, w, cin, seed=1), rnd(cout, cin, k, k, seed=2, scale=0.2)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    dy, add = rnd(b, ho, wo, cout, seed=3), rnd(b, h, w, cin, seed=4)
    dw = torch.empty(cout, cin, k, k)
    ok(emu, emu.ftc_train_conv2d_wgrad(P(x), P(dy), 0, b, h, w, cin, cout, k, stride, P(dw), None))
    assert rel_l2(dw, TO.conv2d_wgrad(x, dy, k, stride)) < 2e-5
    dx = torch.empty(b, h, w, cin)
    ok(emu, emu.ftc_train_conv2d_dgrad(P(dy), 0, b, h, w, cin, cout, k, stride, P(wt), P(add), P(dx), None))
    assert rel_l2(dx, TO.conv2d_dgrad(dy, wt, h, w, stride, add)) < 2e-5


@pytest.mark.parametrize("stride,h,w,c", [(1, 5, 4, 40), (2, 7, 6, 9)])
def test_emu_depthwise_and_se_and_upsample(emu, stride, h, w, c):
    b = 2
    x, w9c = rnd(b, h, w, c, seed=1), rnd(9, c, seed=2, scale=0.3)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    dy = rnd(b, ho, wo, c, seed=3)
    y, dx, dw = torch.empty(b, ho, wo, c), torch.empty(b, h, w, c), torch.empty(9, c)
    ok(emu, emu.ftc_train_dwconv3x3(P(x), P(y), 0, b, h, w, c, stride, P(w9c), None))
    ok(emu, emu.ftc_train_dwconv3x3_dgrad(P(dy), P(dx), 0, b, h, w, c, stride, P(w9c), None))
    ok(emu, emu.ftc_train_dwconv3x3_wgrad(P(x), P(dy), 0, b, h, w, c, stride, P(dw), None))
    assert rel_l2(y, TO.dwconv3x3_raw(x, w9c, stride)) < 1e-5 and rel_l2(dx, TO.dwconv3x3_dgrad(dy, w9c, h, w, stride)) < 1e-5
    assert rel_l2(dw, TO.dwconv3x3_wgrad(x, dy, stride)) < 1e-5
    # squeeze-excitation pieces on the same tensor
    s = 3
    hw = h * w
    w1, b1, w2, b2 = rnd(s, c, seed=5, scale=0.3), rnd(s, seed=6), rnd(c, s, seed=7, scale=0.5), rnd(c, seed=8)
    mean = torch.empty(b, c)
    ok(emu, emu.ftc_train_spatial_sum(P(x), None, 0, b, hw, c, C.c_float(1.0 / hw), P(mean), None))
    mean0 = TO.spatial_sum(x, None, 1.0 / hw)
    assert rel_l2(mean, mean0) < 1e-5
    hid, gate = torch.empty(b, s), torch.empty(b, c)
    ok(emu, emu.ftc_train_se_fc(P(mean0), b, c, s, P(w1), P(b1), P(w2), P(b2), P(hid), P(gate), None))
    hid0, gate0 = TO.se_fc_train(mean0, w1, b1, w2, b2)
    assert rel_l2(hid, hid0) < 1e-5 and rel_l2(gate, gate0) < 1e-5
    g = rnd(b, h, w, c, seed=9)
    dgate = torch.empty(b, c)
    ok(emu, emu.ftc_train_spatial_sum(P(g), P(x), 0, b, hw, c, C.c_float(1.0), P(dgate), None))
    dgate0 = TO.spatial_sum(g, x, 1.0)
    assert rel_l2(dgate, dgate0) < 1e-5
    outs = [torch.empty(b, c), torch.empty(b, s), torch.empty(b, c), torch.empty(s, c), torch.empty(s), torch.empty(c, s), torch.empty(c)]
    ok(emu, emu.ftc_train_se_fc_bwd(P(dgate0), P(gate0), P(hid0), P(mean0), b, c, s, P(w1), P(w2), *[P(t) for t in outs], None))
    for got, ref in zip(outs[2:], TO.se_fc_train_bwd(dgate0, gate0, hid0, mean0, w1, w2)):
        assert rel_l2(got, ref) < 1e-5
    dxs = torch.empty(b, h, w, c)
    ok(emu, emu.ftc_train_scale_bc(P(g), P(gate0), P(outs[2]), C.c_float(1.0 / hw), P(dxs), 0, b, hw, c, None))
    assert rel_l2(dxs, TO.scale_bc(g, gate0, outs[2], 1.0 / hw)) < 1e-5
    # upsample adjoint
    du = rnd(b, 2 * h, 2 * w, c, seed=10)
    dxu = torch.empty(b, h, w, c)
    ok(emu, emu.ftc_train_upsample2x_bwd(P(du), P(dxu), 0, b, h, w, c, None))
    assert rel_l2(dxu, TO.upsample2x_bwd(du)) < 1e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_emu_layernorm_swiglu_embed(emu, dt):
    rows, d = 13, 72
    x, r1, r2, dy = (rnd(rows, d, seed=s, dt=dt) for s in (1, 2, 3, 6))
    gamma, beta = rnd(d, seed=4) + 1.0, rnd(d, seed=5)
    tol = 2e-5 if dt == torch.float32 else 2e-2
    for res in ((None, None), (r1, None), (r1, r2)):
        y, xs, mean, rstd = torch.empty_like(x), torch.empty_like(x), torch.empty(rows), torch.empty(rows)
        has = res[0] is not None
        ok(emu, emu.ftc_train_layernorm(P(x), P(res[0]), P(res[1]), P(xs) if has else None, P(y), P(mean), P(rstd), DT[dt],
                                        C.c_int64(rows), d, P(gamma), P(beta), C.c_float(1e-5), None))
        y0, xs0, mean0, rstd0 = TO.layernorm_train(x, gamma, beta, 1e-5, res[0], res[1])
        assert rel_l2(y.float(), y0.float()) < tol and rel_l2(mean, mean0) < 1e-4 + tol and rel_l2(rstd, rstd0) < 1e-4 + tol
        if has:
            assert rel_l2(xs.float(), xs0.float()) < tol
        dx, dg, db, sc = torch.empty_like(x), torch.empty(d), torch.empty(d), scratch(emu, rows, d)
        ok(emu, emu.ftc_train_layernorm_bwd(P(xs0), P(dy), P(dx), P(mean0), P(rstd0), DT[dt], C.c_int64(rows), d, P(gamma), P(dg), P(db),
                                            P(sc), None))
        dx0, dg0, db0 = TO.layernorm_train_bwd(xs0, dy, mean0, rstd0, gamma)
        assert rel_l2(dx.float(), dx0.float()) < max(tol, 1e-4) and rel_l2(dg, dg0) < 1e-4 and rel_l2(db, db0) < 1e-4
    a, b, dh = (rnd(7, 33, seed=s, dt=dt) for s in (7, 8, 9))
    h, da, dbb = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
    ok(emu, emu.ftc_train_swiglu(P(a), P(b), P(h), DT[dt], C.c_int64(a.numel()), None))
    ok(emu, emu.ftc_train_swiglu_bwd(P(a), P(b), P(dh), P(da), P(dbb), DT[dt], C.c_int64(a.numel()), None))
    da0, db0 = TO.swiglu_bwd(a, b, dh)
    assert rel_l2(h.float(), TO.swiglu(a, b).float()) < tol and rel_l2(da.float(), da0.float()) < tol and rel_l2(dbb.float(), db0.float()) < tol
    ms = (11, 13, 17)
    tabs = [rnd(m, 24, seed=20 + m) for m in ms]
    tok = torch.randint(0, 5000, (3, 9), generator=torch.Generator().manual_seed(7))
    e = torch.empty(3, 9, 24, dtype=dt)
    ok(emu, emu.ftc_train_embed3(P(tok), P(tabs[0]), P(tabs[1]), P(tabs[2]), *ms, P(e), DT[dt], C.c_int64(27), 24, None))
    assert rel_l2(e.float(), TO.embed3(tok, tabs, dt).float()) < tol
    de = rnd(3, 9, 24, seed=11, dt=dt)
    dts = [torch.empty(m, 24) for m in ms]
    ok(emu, emu.ftc_train_embed3_bwd(P(tok), P(de), DT[dt], C.c_int64(27), 24, *ms, *[P(t) for t in dts], None))
    for got, ref in zip(dts, TO.embed3_bwd(tok, de, ms)):
        assert rel_l2(got, ref) < 1e-5


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("b,heads,hd,lt,ls,masked", [(1, 2, 16, 5, 5, False), (2, 2, 32, 6, 37, True),
                                                     (2, 3, 32, 50, 43, True), (1, 1, 64, 9, 128, False),    # fused kernel (<= 128), ragged 4 x 4 tiles
                                                     (1, 2, 16, 7, 150, False)])                             # > 128 keys: row / column kernels
def test_emu_attention_backward(emu, dt, b, heads, hd, lt, ls, masked):
    d = heads * hd
    q, do = rnd(b, lt, d, seed=1, dt=dt), rnd(b, lt, d, seed=4, dt=dt)
    k, v = rnd(b, ls, d, seed=2, dt=dt), rnd(b, ls, d, seed=3, dt=dt)
    mask = None
    if masked:
        mask = torch.zeros(b, ls)
        mask[0, 30:] = float("-inf")
        mask[1, 9:] = float("-inf")
    dq, dk, dv = torch.empty(b, lt, d), torch.empty(b, ls, d), torch.empty(b, ls, d)
    sc = torch.empty(emu.ftc_train_attention_bwd_scratch_bytes(b, heads, lt, ls) // 4)
    ok(emu, emu.ftc_train_attention_bwd(P(q), P(k), P(v), P(mask), P(do), P(dq), P(dk), P(dv), P(sc), DT[dt], b, heads, hd, lt, ls, None))
    for got, ref in zip((dq, dk, dv), TO.attention_bwd(q.float(), k.float(), v.float(), do.float(), heads, mask)):
        assert rel_l2(got, ref) < 1e-4


def test_emu_ce_rows_grad(emu):
    rows, ms = 9, (1091, 1093, 1097)
    ls = [rnd(rows, m, seed=m) for m in ms]
    tgt = torch.randint(0, 0x3FFFF, (rows,), generator=torch.Generator().manual_seed(1))
    wgt = torch.rand(rows, generator=torch.Generator().manual_seed(2))
    sel = (torch.rand(rows, generator=torch.Generator().manual_seed(3)) < 0.7).to(torch.uint8)
    coef = torch.tensor([0.37])
    gs = [torch.empty_like(l) for l in ls]
    ok(emu, emu.ftc_ce_rows_grad(P(ls[0]), P(ls[1]), P(ls[2]), *ms, *ms, P(tgt), P(wgt), P(sel), rows, P(coef), *[P(g) for g in gs], None))
    for l, m, g in zip(ls, ms, gs):
        ref = 0.37 * (wgt * sel)[:, None] * (torch.softmax(l, -1) - torch.nn.functional.one_hot(tgt % m, m).float())
        assert rel_l2(g, ref) < 1e-5


@pytest.mark.parametrize("b,h,w,cin,cout,k,stride", [(2, 6, 5, 8, 16, 3, 1), (1, 9, 7, 24, 40, 3, 2), (3, 5, 4, 16, 8, 1, 1),
                                                     (1, 4, 5, 136, 200, 1, 1), (1, 6, 6, 16, 136, 3, 1)])
def test_emu_mma_weight_gradient(emu, monkeypatch, b, h, w, cin, cout, k, stride):
    """The staged mma.sync weight-gradient kernel (FTC_WGRAD_MMA=1): its tiling, cp.async im2col loader, ldmatrix.trans lane
    addressing, fragment-to-output mapping and split-pixel accumulation, executed with host versions of the three PTX
    primitives that follow the PTX ISA fragment layouts."""
    import subprocess
    code = f"""
import ctypes as C, sys, torch
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
from oracle import train_oracle as TO
lib = C.CDLL({os.path.join(ROOT, 'oracle', '_ref', 'libftc_emu.so')!r})
g = torch.Generator().manual_seed(1)
x = torch.randn({b}, {h}, {w}, {cin}, generator=g).to(torch.bfloat16)
ho, wo = ({h} - 1) // {stride} + 1, ({w} - 1) // {stride} + 1
dy = torch.randn({b}, ho, wo, {cout}, generator=g).to(torch.bfloat16)
dw = torch.empty({cout}, {cin}, {k}, {k})
rc = lib.ftc_train_conv2d_wgrad(C.c_void_p(x.data_ptr()), C.c_void_p(dy.data_ptr()), 1, {b}, {h}, {w}, {cin}, {cout}, {k}, {stride},
                                C.c_void_p(dw.data_ptr()), None)
assert rc == 0
ref = TO.conv2d_wgrad(x.float(), dy.float(), {k}, {stride})
err = float((dw.double() - ref.double()).norm() / ref.double().norm())
assert err < 2e-5, err
print("ok", err)
"""
    # the kernel is selected by an environment switch read once per process: run each case in a fresh interpreter
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FTC_WGRAD_MMA="1"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_emu_edge_cases(emu):
    """Degenerate geometry: one row / one pixel / one channel, extents below a tile, zero variance, a single key."""
    # BatchNorm over a single row: variance 0 -> rstd = 1/sqrt(eps); backward gives dx = 0 (the mean absorbs everything)
    x, gamma, beta, dy = rnd(1, 3, seed=1), rnd(3, seed=2), rnd(3, seed=3), rnd(1, 3, seed=4)
    mean, var, sc = torch.empty(3), torch.empty(3), scratch(emu, 1, 3)
    ok(emu, emu.ftc_train_bn_stats(P(x), 0, C.c_int64(1), 3, P(mean), P(var), P(sc), None))
    assert torch.allclose(mean, x[0]) and float(var.abs().max()) == 0.0
    dx, db, dg = torch.empty_like(x), torch.empty(3), torch.empty(3)
    ok(emu, emu.ftc_train_bn_act_bwd(P(x), P(dy), P(dx), 0, C.c_int64(1), 3, P(mean), P(var), P(gamma), P(beta), C.c_float(1e-3), 0,
                                     P(db), P(dg), P(sc), None))
    assert float(dx.abs().max()) < 1e-6 and torch.allclose(db, dy[0]) and float(dg.abs().max()) < 1e-6
    # 1x1 image, 3x3 conv (only the centre tap sees data), one output channel
    x, dy = rnd(1, 1, 1, 8, seed=5), rnd(1, 1, 1, 1, seed=6)
    dw = torch.empty(1, 8, 3, 3)
    ok(emu, emu.ftc_train_conv2d_wgrad(P(x), P(dy), 0, 1, 1, 1, 8, 1, 3, 1, P(dw), None))
    ref = torch.zeros(1, 8, 3, 3)
    ref[0, :, 1, 1] = x[0, 0, 0] * dy[0, 0, 0, 0]
    assert torch.allclose(dw, ref, atol=1e-6)
    wt = rnd(1, 8, 3, 3, seed=7)
    dxx = torch.empty(1, 1, 1, 8)
    ok(emu, emu.ftc_train_conv2d_dgrad(P(dy), 0, 1, 1, 1, 8, 1, 3, 1, P(wt), None, P(dxx), None))
    assert torch.allclose(dxx[0, 0, 0], wt[0, :, 1, 1] * dy[0, 0, 0, 0], atol=1e-6)
    # attention with one query and one key: softmax = 1 -> dq = dk = 0, dv = dout
    q, k, v, do = (rnd(1, 1, 16, seed=s) for s in (8, 9, 10, 11))
    dq, dk, dv = torch.empty(1, 1, 16), torch.empty(1, 1, 16), torch.empty(1, 1, 16)
    sc = torch.empty(emu.ftc_train_attention_bwd_scratch_bytes(1, 1, 1, 1) // 4)
    ok(emu, emu.ftc_train_attention_bwd(P(q), P(k), P(v), None, P(do), P(dq), P(dk), P(dv), P(sc), 0, 1, 1, 16, 1, 1, None))
    assert float(dq.abs().max()) < 1e-6 and float(dk.abs().max()) < 1e-6 and torch.allclose(dv, do)
    # embedding of negative tokens wraps like Python's % (the reference uses torch's %, same convention)
    tabs = [rnd(m, 8, seed=m) for m in (5, 7, 11)]
    tok = torch.tensor([[-1, 0, 12]])
    e = torch.empty(1, 3, 8)
    ok(emu, emu.ftc_train_embed3(P(tok), P(tabs[0]), P(tabs[1]), P(tabs[2]), 5, 7, 11, P(e), 0, C.c_int64(3), 8, None))
    assert rel_l2(e, TO.embed3(tok, tabs, torch.float32)) < 1e-6
    # bad arguments fail loudly instead of launching
    assert emu.ftc_train_conv2d_wgrad(P(x), P(dy), 0, 1, 1, 1, 8, 1, 5, 1, P(dw), None) != 0 and b"ksize" in emu.ftc_last_error()
    assert emu.ftc_train_bn_stats(None, 0, C.c_int64(1), 3, P(mean), P(var), P(sc), None) != 0


def test_emu_page_maps(emu):
    """ftc_page_maps (tile sigmoid * validity window, atomic maximum over overlapping tiles) on host threads vs the oracle, on
    a shrunken geometry (tiles 12x12 at stride 7 on a 19x19 page map) so that every CUDA thread can be an OS thread."""
    import numpy as np
    g = torch.Generator().manual_seed(5)
    h = w = 12
    offs = [(0, 0), (28, 0), (0, 28), (28, 28)]                # in image pixels (scale 4): 7 map pixels apart
    heat = torch.randn(4, 9, h, w, generator=g) * 2
    meta = torch.tensor([[0, 0, 0, 9, 0, 9], [28, 0, 3, 12, 0, 9], [0, 28, 0, 9, 3, 12], [28, 28, 3, 12, 3, 12]], dtype=torch.int32)
    page = torch.zeros(7, 19, 19)
    ok(emu, emu.ftc_page_maps(P(heat), 4, h, w, P(meta), P(page), 19, 19, 4, None))
    ref = np.zeros((7, 19, 19), dtype=np.float32)
    for b in range(4):
        ox, oy, x0, x1, y0, y1 = (int(v) for v in meta[b])
        mask = np.zeros((h, w), dtype=bool)
        mask[y0:y1, x0:x1] = True
        for m, ch in enumerate([0, 3, 4, 5, 6, 7, 8]):
            p = ((np.tanh(heat[b, ch].numpy() / 2) + 1) / 2) * mask
            ref[m, oy // 4:oy // 4 + h, ox // 4:ox // 4 + w] = np.maximum(p, ref[m, oy // 4:oy // 4 + h, ox // 4:ox // 4 + w])
    assert np.abs(page.numpy() - ref).max() < 1e-6


@pytest.fixture()
def ops_on_emu(emu, monkeypatch):
    """findtextcenternet_b200._ops (the product's ctypes wrappers) pointed at the emulated library: checks the wrappers' argument
    order and the ctypes signatures of _lib.SYMBOLS on CPU.  Test-only patches: CUDA stream / device context / is_cuda."""
    import contextlib
    from findtextcenternet_b200 import _lib, _ops
    for name, (res, args) in _lib.SYMBOLS.items():
        if hasattr(emu, name):
            fn = getattr(emu, name)
            fn.restype, fn.argtypes = res, args
    monkeypatch.setattr(_lib, "_lib", emu)
    monkeypatch.setattr(_ops, "_s", lambda t: None)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    yield _ops
    for name in _lib.SYMBOLS:        # leave the shared fixture library untyped again for the raw-pointer tests
        if hasattr(emu, name):
            getattr(emu, name).argtypes = None
    emu.ftc_train_reduce_scratch_bytes.restype = C.c_size_t
    emu.ftc_train_reduce_scratch_bytes.argtypes = [C.c_int64, C.c_int]
    emu.ftc_train_attention_bwd_scratch_bytes.restype = C.c_size_t
    emu.ftc_last_error.restype = C.c_char_p


def test_ops_wrappers_through_emulated_library(ops_on_emu):
    K = ops_on_emu
    # transformer pieces (no GPU run yet): wrapper argument order and ctypes signatures
    x, r1, dy = rnd(2, 5, 24, seed=1), rnd(2, 5, 24, seed=2), rnd(2, 5, 24, seed=3)
    gamma, beta = rnd(24, seed=4) + 1.0, rnd(24, seed=5)
    y, xs, mean, rstd = K.layernorm_train(x, gamma, beta, 1e-5, r1, None)
    y0, xs0, mean0, rstd0 = TO.layernorm_train(x, gamma, beta, 1e-5, r1, None)
    assert rel_l2(y, y0) < 1e-5 and rel_l2(xs, xs0) < 1e-6 and rel_l2(mean, mean0) < 1e-5 and rel_l2(rstd, rstd0) < 1e-5
    for got, ref in zip(K.layernorm_train_bwd(xs0, dy, mean0, rstd0, gamma), TO.layernorm_train_bwd(xs0, dy, mean0, rstd0, gamma)):
        assert rel_l2(got, ref) < 1e-4
    a, b = rnd(3, 10, seed=6), rnd(3, 10, seed=7)
    assert rel_l2(K.swiglu(a, b), TO.swiglu(a, b)) < 1e-5
    for got, ref in zip(K.swiglu_bwd(a, b, dy[0, :3, :10].contiguous()), TO.swiglu_bwd(a, b, dy[0, :3, :10])):
        assert rel_l2(got, ref) < 1e-5
    tabs = [rnd(m, 8, seed=m) for m in (11, 13, 17)]
    tok = torch.randint(0, 999, (2, 6), generator=torch.Generator().manual_seed(8))
    assert rel_l2(K.embed3(tok, tabs, torch.float32), TO.embed3(tok, tabs, torch.float32)) < 1e-6
    de = rnd(2, 6, 8, seed=9)
    for got, ref in zip(K.embed3_bwd(tok, de, (11, 13, 17)), TO.embed3_bwd(tok, de, (11, 13, 17))):
        assert rel_l2(got, ref) < 1e-6
    q, k, v, do = rnd(2, 4, 32, seed=10), rnd(2, 7, 32, seed=11), rnd(2, 7, 32, seed=12), rnd(2, 4, 32, seed=13)
    mask = torch.zeros(2, 7)
    mask[1, 5:] = float("-inf")
    for got, ref in zip(K.attention_bwd(q, k, v, do, 2, mask), TO.attention_bwd(q, k, v, do, 2, mask)):
        assert rel_l2(got, ref) < 1e-4
    # detector-side wrappers (already run on a B200) through the same route, as a control of the method
    xx, dyy = rnd(2, 4, 4, 8, seed=14), rnd(2, 4, 4, 16, seed=15)
    assert rel_l2(K.conv2d_wgrad(xx, dyy, 3, 1), TO.conv2d_wgrad(xx, dyy, 3, 1)) < 1e-5
    m, vv = K.bn_stats(xx)
    m0, v0 = TO.bn_stats(xx)
    assert rel_l2(m, m0) < 1e-5 and rel_l2(vv, v0) < 1e-4
    for got, ref in zip(K.bn_act_bwd(xx, rnd(2, 4, 4, 8, seed=16), m0, v0, rnd(8, seed=17), rnd(8, seed=18), 1e-3, 1),
                        TO.bn_act_bwd(xx, rnd(2, 4, 4, 8, seed=16), m0, v0, rnd(8, seed=17), rnd(8, seed=18), 1e-3, 1)):
        assert rel_l2(got, ref) < 1e-4


def test_ce_loss_backward_wrapper_through_emulated_library(ops_on_emu, monkeypatch):
    """loss_func._CEMean (forward ftc_ce_rows + backward ftc_ce_rows_grad, the id_loss / train3 loss) against torch autograd."""
    from findtextcenternet_b200 import loss_func as LF
    monkeypatch.setattr(LF, "_s", lambda t: None)
    rows = 7
    ls = [rnd(rows, m, seed=m).requires_grad_() for m in LF.modulo_list]
    tgt = torch.randint(0, 0x3FFFF, (rows,), generator=torch.Generator().manual_seed(1))
    wgt = torch.rand(rows, generator=torch.Generator().manual_seed(2))
    sel = torch.rand(rows, generator=torch.Generator().manual_seed(3)) < 0.7
    loss, o = LF._CEMean.apply(ls[0], ls[1], ls[2], tgt, wgt, sel, sel, True)
    (loss * 1.7).backward()
    ref_in = [l.detach().clone().requires_grad_() for l in ls]
    ce = sum(torch.nn.functional.cross_entropy(l, tgt % m, reduction="none") for l, m in zip(ref_in, LF.modulo_list))
    ref = (ce * wgt)[sel].sum() / torch.clamp_min(wgt[sel].sum(), 1.0)
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    for a, b in zip(ls, ref_in):
        assert rel_l2(a.grad, b.grad) < 1e-5


def test_transformer_train_graph_through_product_wrappers_on_emu(ops_on_emu, monkeypatch):
    """The golden train3 forward + backward (tests/golden/train_transformer_seed0.npz) once more, now through the PRODUCT wrappers
    of _ops.py running the emulated kernels for everything train-specific (LayerNorm, SwiGLU, embeddings, attention backward,
    conv weight / data gradients, bias sums); only the two forward ops whose kernels are not emulated (GEMM forward, attention
    forward - both B200-verified) come from the oracle."""
    import numpy as np
    from conftest import GOLDEN
    from findtextcenternet_b200 import synthetic, train_ops
    from findtextcenternet_b200.models.transformer import Transformer
    from test_train_oracle import TF_DIMS, check_gradients_against_golden

    class Hybrid:
        def __getattr__(self, name):
            if name in ("conv2d", "attention"):
                return getattr(TO, name)
            return getattr(ops_on_emu, name)

    monkeypatch.setattr(train_ops, "K", Hybrid())
    monkeypatch.setattr(train_ops, "_need_cuda", lambda t, what: None)
    gold = np.load(os.path.join(GOLDEN, "train_transformer_seed0.npz"))
    model = Transformer(**TF_DIMS, dropout=0.0)
    model.load_state_dict(synthetic.transformer_state_dict(0, **TF_DIMS))
    model.set_precision("fp32").train()
    outs = model(torch.from_numpy(gold["enc"]), torch.from_numpy(gold["dec"]))
    for i in range(3):
        assert rel_l2(outs[i].detach(), gold[f"out{i}"]) < 1e-4
    sum((o * torch.from_numpy(gold[f"w{i}"])).sum() for i, o in enumerate(outs)).backward()
    check_gradients_against_golden(gold, dict(model.named_parameters()), 1e-3)


def test_emu_bf16_storage_of_the_bandwidth_kernels(emu):
    """bf16 activations through the depthwise / squeeze-excitation / upsample-adjoint kernels (vector and scalar variants)."""
    dt = torch.bfloat16
    for c in (16, 12):          # 16: the 16-byte-vector kernels; 12: the scalar ones
        b, h, w = 2, 6, 5
        x, w9c, dy = rnd(b, h, w, c, seed=1, dt=dt), rnd(9, c, seed=2, scale=0.3), rnd(b, h, w, c, seed=3, dt=dt)
        y, dx, dw = torch.empty_like(x), torch.empty_like(x), torch.empty(9, c)
        ok(emu, emu.ftc_train_dwconv3x3(P(x), P(y), 1, b, h, w, c, 1, P(w9c), None))
        ok(emu, emu.ftc_train_dwconv3x3_dgrad(P(dy), P(dx), 1, b, h, w, c, 1, P(w9c), None))
        ok(emu, emu.ftc_train_dwconv3x3_wgrad(P(x), P(dy), 1, b, h, w, c, 1, P(dw), None))
        assert rel_l2(y.float(), TO.dwconv3x3_raw(x.float(), w9c, 1)) < 1e-2
        assert rel_l2(dx.float(), TO.dwconv3x3_dgrad(dy.float(), w9c, h, w, 1)) < 1e-2
        assert rel_l2(dw, TO.dwconv3x3_wgrad(x.float(), dy.float(), 1)) < 1e-5
        mean, dgate = torch.empty(b, c), torch.empty(b, c)
        ok(emu, emu.ftc_train_spatial_sum(P(x), None, 1, b, h * w, c, C.c_float(1.0 / (h * w)), P(mean), None))
        ok(emu, emu.ftc_train_spatial_sum(P(dy), P(x), 1, b, h * w, c, C.c_float(1.0), P(dgate), None))
        assert rel_l2(mean, TO.spatial_sum(x.float(), None, 1.0 / (h * w))) < 1e-5
        assert rel_l2(dgate, TO.spatial_sum(dy.float(), x.float(), 1.0)) < 1e-5
        gate, bias = torch.rand(b, c, generator=torch.Generator().manual_seed(4)), rnd(b, c, seed=5)
        out = torch.empty_like(x)
        ok(emu, emu.ftc_train_scale_bc(P(x), P(gate), P(bias), C.c_float(0.25), P(out), 1, b, h * w, c, None))
        assert rel_l2(out.float(), TO.scale_bc(x.float(), gate, bias, 0.25)) < 1e-2
        du = rnd(b, 2 * h, 2 * w, c, seed=6, dt=dt)
        dxu = torch.empty_like(x)
        ok(emu, emu.ftc_train_upsample2x_bwd(P(du), P(dxu), 1, b, h, w, c, None))
        assert rel_l2(dxu.float(), TO.upsample2x_bwd(du.float())) < 1e-2


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_detector_blocks_through_product_wrappers_on_emu(ops_on_emu, monkeypatch, dt):
    """A FusedMBConv block, an MBConv block (depthwise stride 2 + squeeze-excitation), a residual MBConv block with a pinned
    StochasticDepth draw and a conv-BN-GELU-upsample head stage, forward + backward, through train_ops.py -> _ops.py -> the
    emulated kernels (incl. the 16-byte-vector BatchNorm / SE kernels) against the same graph on the pure oracle."""
    from findtextcenternet_b200 import arch, train_ops
    from findtextcenternet_b200.models._tree import Node, populate

    def build():
        torch.manual_seed(0)
        root = Node()
        specs = list(arch._conv_bn("a.block.0", 16, 64, 3)) + list(arch._conv_bn("a.block.1", 64, 24, 1))           # fused, expand 4
        for name, cin, cout in (("b", 24, 32), ("c", 32, 32)):
            exp = cin * 2
            specs += list(arch._conv_bn(f"{name}.block.0", cin, exp, 1)) + list(arch._conv_bn(f"{name}.block.1", exp, exp, 3, groups=exp))
            specs += [arch.ParamSpec(f"{name}.block.2.fc1.weight", (8, exp, 1, 1), "conv", exp), arch.ParamSpec(f"{name}.block.2.fc1.bias", (8,), "bias"),
                      arch.ParamSpec(f"{name}.block.2.fc2.weight", (exp, 8, 1, 1), "conv", 8), arch.ParamSpec(f"{name}.block.2.fc2.bias", (exp,), "bias")]
            specs += list(arch._conv_bn(f"{name}.block.3", exp, cout, 1))
        specs += list(arch._conv_bn("up", 32, 16, 3))
        populate(root, specs, backbone_prefix="\0")
        return root

    def run(K, root, x, noise):
        monkeypatch.setattr(train_ops, "K", K)
        sa = arch.StageCfg(True, 4, 3, 1, 16, 24, 1)
        sb = arch.StageCfg(False, 2, 3, 2, 24, 32, 1)
        sc = arch.StageCfg(False, 2, 3, 1, 32, 32, 1)
        y = train_ops._block(x, root.a, sa, 16, 1, 0.0, None)
        y = train_ops._block(y, root.b, sb, 24, 2, 0.0, None)
        y = train_ops._block(y, root.c, sc, 32, 1, 0.3, noise)          # residual + pinned StochasticDepth noise
        y = train_ops._Upsample2x.apply(train_ops.conv_bn_act(y, root.up, arch.HEAD_BN_EPS, train_ops._lib.ACT_GELU))
        w = torch.randn(y.shape, generator=torch.Generator().manual_seed(3))
        (y.float() * w).sum().backward()
        return y.detach().float(), {n: p.grad.clone() for n, p in root.named_parameters()}

    class Hybrid:
        def __getattr__(self, name):
            # the tensor-core convolution entry points are not in the emulation library: convolutions (and the "is there a fused
            # data-gradient convolution" probe of train_ops) resolve on the oracle namespace
            return getattr(TO, name) if name in ("conv2d", "upsample2x", "conv2d_dgrad_tc") else getattr(ops_on_emu, name)

    x = torch.randn(2, 8, 8, 16, generator=torch.Generator().manual_seed(1)).to(dt)
    noise = torch.tensor([0.0, 1.0 / 0.7])
    ref_root, emu_root = build(), build()
    y_ref, g_ref = run(TO, ref_root, x.clone().requires_grad_(), noise)
    y_emu, g_emu = run(Hybrid(), emu_root, x.clone().requires_grad_(), noise)
    tol = 1e-4 if dt == torch.float32 else 6e-2
    assert rel_l2(y_emu, y_ref) < tol
    # BatchNorm shifts that feed a 1x1 conv + batch-statistics layer have exactly zero true gradient (both sides return rounding
    # noise there): the error is measured against the largest gradient of the graph as well as against the tensor itself
    scale = max(float(g.double().norm()) for g in g_ref.values())
    gtol = tol + (0.0 if dt == torch.float32 else 0.1)
    for n in g_ref:
        err = float((g_emu[n].double() - g_ref[n].double()).norm())
        assert err <= gtol * float(g_ref[n].double().norm()) + (1e-5 if dt == torch.float32 else 1e-3) * scale, (n, err)
    for (n, b_ref), (_, b_emu) in zip(ref_root.named_buffers(), emu_root.named_buffers()):
        btol = 1e-3 + (0 if dt == torch.float32 else 2e-2)      # + absolute floor: means of exactly centred inputs are pure noise
        assert float((b_emu.double() - b_ref.double()).norm()) <= btol * float(b_ref.double().norm()) + (1e-5 if dt == torch.float32 else 1e-3), n


# ---- page-level box selection (csrc/page_ops.cu) on host threads vs the reference-pinned oracle --------------------------------
def _select_on_emu(emu, loc32, gf, tight, th, seps, code):
    import numpy as np
    n = loc32.shape[0]
    order = torch.from_numpy(np.argsort(-loc32[:, 0].astype(np.float64), kind="stable").astype(np.int32))
    loc_t, gf_t, tight_t = torch.from_numpy(loc32.copy()), torch.from_numpy(gf.copy()), torch.from_numpy(tight.copy())
    seps_t, code_t = torch.from_numpy(seps.copy()), torch.from_numpy(np.ascontiguousarray(code))
    emu.ftc_select_boxes_scratch_bytes.restype = C.c_size_t
    nb = emu.ftc_select_boxes_scratch_bytes(n)
    scr = torch.empty(nb, dtype=torch.uint8)
    n_out, sel = torch.zeros(1, dtype=torch.int32), torch.zeros(n, dtype=torch.int32)
    out_loc, out_gf = torch.zeros(n, 9), torch.zeros(n, gf.shape[1])
    ok(emu, emu.ftc_select_boxes(P(loc_t), P(gf_t), gf.shape[1], P(order), n, P(tight_t), C.c_double(th), P(seps_t), P(code_t),
                                 seps.shape[0], seps.shape[1], 4, P(n_out), P(sel), P(out_loc), P(out_gf), P(scr), C.c_size_t(nb), None))
    m = int(n_out[0])
    return out_loc[:m].numpy(), out_gf[:m].numpy(), sel[:m].numpy()


def test_emu_box_hists_and_select_boxes_real_page():
    """ftc_box_hists + ftc_select_boxes (kernel source on host threads) on the 402 per-tile peaks of the 4-tile golden page: the
    histogram scores are bit-identical to the oracle's numpy imageHist, and the selected boxes equal the UNMODIFIED reference
    run_detector's final output (tests/golden/page4_seed0.npz), bit for bit."""
    import numpy as np
    from conftest import GOLDEN
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.process_ocr_b200 import page_tiles
    from oracle import detector_oracle as DO
    sys.path.insert(0, os.path.join(ROOT, "oracle", "emu"))
    import build_emu
    emu = C.CDLL(build_emu.build())
    emu.ftc_last_error.restype = C.c_char_p
    gold = np.load(os.path.join(GOLDEN, "page4_seed0.npz"))
    h, w = (int(v) for v in gold["image_hw"])
    page, _ = page_tiles(synthetic.page_image(int(gold["seed"]), h, w))
    loc32, gf = gold["pre_locations"], gold["pre_glyphfeatures"]
    n = loc32.shape[0]
    page_t = torch.from_numpy(np.ascontiguousarray(page))
    hists = torch.zeros(2, n, dtype=torch.float64)
    loc_t = torch.from_numpy(loc32.copy())
    ok(emu, emu.ftc_box_hists(P(page_t), page.shape[0], page.shape[1], P(loc_t), n, P(hists), None))
    loose, tight = DO.box_hists(loc32.astype(np.float64), page.astype(np.float32))
    assert np.array_equal(hists[0].numpy(), loose) and np.array_equal(hists[1].numpy(), tight)
    assert len(set(np.round(loose, 3))) > 20                     # the fixture has ink, blank and noise boxes
    th = float(np.median(loose) / 5)
    maps7 = gold["maps7"]
    out_loc, out_gf, sel = _select_on_emu(emu, loc32, gf, tight, th, maps7[2], maps7[3:7])
    assert out_loc.shape == gold["locations"].shape
    assert np.array_equal(out_loc, gold["locations"]) and np.array_equal(out_gf, gold["glyphfeatures"])


def test_emu_select_boxes_dense_page():
    """The greedy kernel on the dense stub page (1 150 candidates, 374 survivors in the reference): IoU / 75 % / fill-map /
    histogram / separator / code-maximum branches all fire; output == reference run_detector golden."""
    import numpy as np
    from test_oracle_golden import _dense_page
    from oracle import detector_oracle as DO
    sys.path.insert(0, os.path.join(ROOT, "oracle", "emu"))
    import build_emu
    emu = C.CDLL(build_emu.build())
    emu.ftc_last_error.restype = C.c_char_p
    gold, page, offsets, heat10, feats = _dense_page()
    ph, pw = page.shape[:2]
    pre = [DO.decode_tile(heat10[i], feats[i], x, y, pw, ph) for i, (x, y) in enumerate(offsets)]
    loc = np.concatenate([p[0] for p in pre]).astype(np.float32)
    gf = np.concatenate([p[1] for p in pre]).astype(np.float32)
    maps7 = DO.page_maps(np.concatenate([heat10[:, :1], heat10[:, 2:]], 1), offsets, pw, ph)
    loose, tight = DO.box_hists(loc.astype(np.float64), page.astype(np.float32))
    out_loc, out_gf, sel = _select_on_emu(emu, loc, gf, tight, float(np.median(loose) / 5), maps7[2], maps7[3:7])
    assert out_loc.shape == gold["locations"].shape and np.array_equal(out_loc, gold["locations"])
    assert np.allclose(out_gf.astype(np.float64).sum(1), gold["glyphfeatures_sum"], rtol=0, atol=1e-9)


def test_emu_bn_stats_running_update(emu):
    """ftc_train_bn_stats_running: batch statistics AND nn.BatchNorm's train-mode side effects (momentum update with the unbiased
    variance, num_batches_tracked) in one call == torch.nn.functional.batch_norm(training=True)."""
    import torch.nn.functional as F
    rows, c = 300, 72
    x = rnd(rows, c, seed=1) * 1.7 + 0.3
    rm, rv = rnd(c, seed=2), rnd(c, seed=3).abs() + 0.5
    nbt = torch.tensor(7, dtype=torch.int64)
    rm0, rv0 = rm.clone(), rv.clone()
    F.batch_norm(x, rm0, rv0, None, None, True, 0.1, 1e-3)
    mean, var, sc = torch.empty(c), torch.empty(c), scratch(emu, rows, c)
    ok(emu, emu.ftc_train_bn_stats_running(P(x), 0, C.c_int64(rows), c, P(mean), P(var), P(sc), P(rm), P(rv), P(nbt), C.c_float(0.1), None))
    m0, v0 = TO.bn_stats(x)
    assert rel_l2(mean, m0) < 1e-5 and rel_l2(var, v0) < 1e-4
    assert rel_l2(rm, rm0) < 1e-6 and rel_l2(rv, rv0) < 1e-6 and int(nbt) == 8


def test_emu_row_strided_dy_is_read_in_place(emu):
    """The gradient of a channel slice of a wider map (torch.cat's backward) goes to the BatchNorm / upsample backward kernels with its
    row stride instead of a contiguous copy: same numbers as the contiguous call."""
    rows, c, wide = 300, 64, 160
    x = (rnd(rows, c, seed=1) * 1.3).to(torch.bfloat16)
    dwide = rnd(rows, wide, seed=2, dt=torch.bfloat16)
    dy = dwide[:, 96:160]                                   # channel slice: row stride 160, 16-byte aligned offset
    gamma, beta = rnd(c, seed=3), rnd(c, seed=4)
    m0, v0 = TO.bn_stats(x.float())
    sc = scratch(emu, rows, c)
    outs = []
    for d, ld in ((dy.contiguous(), c), (dy, wide)):
        dx, dbeta, dgamma = torch.empty_like(x), torch.empty(c), torch.empty(c)
        ok(emu, emu.ftc_train_bn_act_bwd_ld(P(x), C.c_void_p(d.data_ptr()), C.c_int64(ld), P(dx), 1, C.c_int64(rows), c, P(m0), P(v0), P(gamma),
                                            P(beta), C.c_float(1e-3), 1, P(dbeta), P(dgamma), P(sc), None))
        outs.append((dx.clone(), dbeta.clone(), dgamma.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    b_, h, w, cc = 2, 3, 4, 16
    duw = rnd(b_, 2 * h, 2 * w, 40, seed=5, dt=torch.bfloat16)
    du = duw[..., 8:24]
    res = []
    for d, ld in ((du.contiguous(), cc), (du, 40)):
        dxu = torch.empty(b_, h, w, cc, dtype=torch.bfloat16)
        ok(emu, emu.ftc_train_upsample2x_bwd_ld(C.c_void_p(d.data_ptr()), C.c_int64(ld), P(dxu), 1, b_, h, w, cc, None))
        res.append(dxu.clone())
    assert torch.equal(res[0], res[1])
    from findtextcenternet_b200 import _ops
    assert _ops._channel_slice_ld(dy, c) == wide and _ops._channel_slice_ld(dy.contiguous(), c) == 0
    assert _ops._channel_slice_ld(du, cc) == 40 and _ops._channel_slice_ld(duw[:, ::2, :, 8:24], cc) == 0


@pytest.mark.parametrize("b,h,w,c", [(2, 8, 6, 72), (1, 4, 10, 64), (3, 12, 5, 8)])
def test_emu_depthwise_weight_gradient_strip_kernel(emu, b, h, w, c):
    """bf16, stride 1, H % 4 == 0: the shared-memory strip kernel (cp.async staging, zero halo) against the oracle"""
    x = rnd(b, h, w, c, seed=1, dt=torch.bfloat16)
    dy = rnd(b, h, w, c, seed=2, dt=torch.bfloat16)
    dw = torch.empty(9, c)
    ok(emu, emu.ftc_train_dwconv3x3_wgrad(P(x), P(dy), 1, b, h, w, c, 1, P(dw), None))
    assert rel_l2(dw, TO.dwconv3x3_wgrad(x.float(), dy.float(), 1)) < 1e-5
