"""CPU: the train-step oracle (oracle/train_oracle.py) against torch autograd over the modules the reference differentiates,
and the host-side autograd graph of findtextcenternet_b200/train_ops.py (kernel namespace swapped for the oracle) against the
reference-generated golden of a full train-mode forward + backward (tests/golden/train_xl64_seed0.npz)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, rel_l2
from oracle import train_oracle as TO


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("k,stride,h,w", [(1, 1, 5, 7), (3, 1, 6, 5), (3, 2, 8, 6), (3, 2, 7, 9), (1, 2, 6, 6)])
def test_conv_grads_match_autograd(k, stride, h, w):
    g = torch.Generator().manual_seed(k * 10 + stride)
    x = torch.randn(2, 5, h, w, generator=g, requires_grad=True)
    wt = torch.randn(6, 5, k, k, generator=g, requires_grad=True)
    y = F.conv2d(x, wt, None, stride, (k - 1) // 2)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    assert rel_l2(TO.conv2d(nhwc(x.detach()), wt, stride), nhwc(y.detach())) < 1e-6
    assert rel_l2(TO.conv2d_wgrad(nhwc(x.detach()), nhwc(dy), k, stride), wt.grad) < 1e-6
    add = torch.randn(2, h, w, 5, generator=g)
    dx = TO.conv2d_dgrad(nhwc(dy), wt, h, w, stride, add)
    assert rel_l2(dx, nhwc(x.grad) + add) < 1e-6


@pytest.mark.parametrize("act", [TO.ACT_NONE, TO.ACT_SILU, TO.ACT_GELU])
def test_bn_act_matches_autograd(act):
    g = torch.Generator().manual_seed(act)
    x = (torch.randn(3, 6, 5, 4, generator=g) * 2 + 0.5).requires_grad_()
    gamma = torch.randn(6, generator=g).requires_grad_()
    beta = torch.randn(6, generator=g).requires_grad_()
    res = torch.randn(3, 6, 5, 4, generator=g)
    rm, rv = torch.zeros(6), torch.ones(6)
    z = F.batch_norm(x, rm, rv, gamma, beta, True, 0.1, 1e-3)
    y = {TO.ACT_NONE: lambda t: t, TO.ACT_SILU: F.silu, TO.ACT_GELU: F.gelu}[act](z) + res
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xs = nhwc(x.detach())
    mean, var = TO.bn_stats(xs)
    n = xs.numel() // 6
    assert rel_l2(mean * 0.1, rm) < 1e-5 and rel_l2(0.9 + 0.1 * var * n / (n - 1), rv) < 1e-5
    assert rel_l2(TO.bn_act(xs, mean, var, gamma, beta, 1e-3, act, nhwc(res)), nhwc(y.detach())) < 1e-6
    dx, dgamma, dbeta = TO.bn_act_bwd(xs, nhwc(dy), mean, var, gamma, beta, 1e-3, act)
    assert rel_l2(dx, nhwc(x.grad)) < 1e-5
    assert rel_l2(dgamma, gamma.grad) < 1e-5 and rel_l2(dbeta, beta.grad) < 1e-5


@pytest.mark.parametrize("stride,h,w", [(1, 6, 5), (2, 8, 6), (2, 7, 5)])
def test_depthwise_matches_autograd(stride, h, w):
    g = torch.Generator().manual_seed(stride)
    c = 7
    x = torch.randn(2, c, h, w, generator=g, requires_grad=True)
    wt = torch.randn(c, 1, 3, 3, generator=g, requires_grad=True)
    y = F.conv2d(x, wt, None, stride, 1, 1, c)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    w9c = wt.detach().reshape(c, 9).t().contiguous()
    assert rel_l2(TO.dwconv3x3_raw(nhwc(x.detach()), w9c, stride), nhwc(y.detach())) < 1e-6
    assert rel_l2(TO.dwconv3x3_dgrad(nhwc(dy), w9c, h, w, stride), nhwc(x.grad)) < 1e-6
    assert rel_l2(TO.dwconv3x3_wgrad(nhwc(x.detach()), nhwc(dy), stride).t().reshape(c, 1, 3, 3), wt.grad) < 1e-6


def test_squeeze_excite_matches_torchvision_autograd():
    from torchvision.ops.misc import SqueezeExcitation
    from functools import partial
    torch.manual_seed(0)
    c, s = 12, 3
    se = SqueezeExcitation(c, s, activation=partial(torch.nn.SiLU, inplace=True))   # efficientnet.py:149
    x = torch.randn(2, c, 5, 4, requires_grad=True)
    y = se(x)
    dy = torch.randn(y.shape)
    y.backward(dy)
    xs, dys = nhwc(x.detach()), nhwc(dy)
    w1, w2 = se.fc1.weight.detach().reshape(s, c), se.fc2.weight.detach().reshape(c, s)
    mean = TO.spatial_sum(xs, None, 1.0 / 20)
    hid_pre, gate = TO.se_fc_train(mean, w1, se.fc1.bias.detach(), w2, se.fc2.bias.detach())
    assert rel_l2(TO.scale_bc(xs, gate), nhwc(y.detach())) < 1e-6
    dgate = TO.spatial_sum(dys, xs, 1.0)
    dmean, dw1, db1, dw2, db2 = TO.se_fc_train_bwd(dgate, gate, hid_pre, mean, w1, w2)
    assert rel_l2(TO.scale_bc(dys, gate, dmean, 1.0 / 20), nhwc(x.grad)) < 1e-5
    assert rel_l2(dw1, se.fc1.weight.grad.reshape(s, c)) < 1e-5 and rel_l2(db1, se.fc1.bias.grad) < 1e-5
    assert rel_l2(dw2, se.fc2.weight.grad.reshape(c, s)) < 1e-5 and rel_l2(db2, se.fc2.bias.grad) < 1e-5


@pytest.mark.parametrize("h,w", [(1, 1), (2, 3), (6, 6), (24, 5)])
def test_upsample_adjoint_matches_autograd(h, w):
    g = torch.Generator().manual_seed(h)
    x = torch.randn(2, 3, h, w, generator=g, requires_grad=True)
    y = torch.nn.UpsamplingBilinear2d(scale_factor=2)(x)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    assert rel_l2(TO.upsample2x(nhwc(x.detach())), nhwc(y.detach())) < 1e-6
    assert rel_l2(TO.upsample2x_bwd(nhwc(dy)), nhwc(x.grad)) < 1e-6


def test_upsample_adjoint_candidate_window():
    """csrc/train_ops.cu::upsample2x_bwd_kernel only visits outputs 2j-2 .. 2j+3 for input index j: every output whose
    footprint touches j must lie inside that window."""
    for n in (1, 2, 3, 6, 12, 24, 48, 96, 192):
        m = TO._interp_matrix(n)
        for j in range(n):
            touched = torch.nonzero(m[:, j]).flatten().tolist()
            assert touched and min(touched) >= 2 * j - 2 and max(touched) <= 2 * j + 3, (n, j, touched)


@pytest.fixture()
def oracle_kernels(monkeypatch):
    """Route the autograd nodes of train_ops.py through the CPU oracle (checks the graph logic, not the kernels)."""
    from findtextcenternet_b200 import train_ops

    class Ns:
        pass

    ns = Ns()
    for name in ("conv2d", "conv2d_wgrad", "conv2d_dgrad", "bn_stats", "bn_act", "bn_act_bwd", "dwconv3x3_raw", "dwconv3x3_dgrad",
                 "dwconv3x3_wgrad", "spatial_sum", "scale_bc", "se_fc_train", "se_fc_train_bwd", "upsample2x", "upsample2x_bwd",
                 "layernorm_train", "layernorm_train_bwd", "swiglu", "swiglu_bwd", "embed3", "embed3_bwd", "attention", "attention_bwd"):
        setattr(ns, name, getattr(TO, name))
    monkeypatch.setattr(train_ops, "K", ns)
    # the product wrappers refuse CPU tensors; the graph test runs them on CPU on purpose
    monkeypatch.setattr(train_ops, "_need_cuda", lambda t, what: None)
    return train_ops


def test_row_scale_node(oracle_kernels):
    x = torch.randn(3, 4, 5, 6, requires_grad=True)
    noise = torch.tensor([0.0, 1.25, 1.25])
    y = oracle_kernels._RowScale.apply(x, noise)
    y.backward(torch.ones_like(y))
    assert torch.allclose(y, x.detach() * noise[:, None, None, None])
    assert torch.allclose(x.grad, noise[:, None, None, None].expand_as(x))


def _golden_train():
    p = os.path.join(GOLDEN, "train_xl64_seed0.npz")
    if not os.path.exists(p):
        pytest.skip("tests/golden/train_xl64_seed0.npz missing (oracle/make_golden_train.py)")
    return np.load(p)


@pytest.mark.slow
def test_train_graph_matches_reference_golden(oracle_kernels):
    """Full TextDetectorModel train-mode forward + backward (StochasticDepth off, as in the golden) through train_ops.py with
    oracle kernels == the unmodified reference under torch autograd: outputs, updated BatchNorm buffers, every gradient."""
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.models.detector import TextDetectorModel
    gold = _golden_train()
    model = TextDetectorModel(pre_weights=False)
    model.load_state_dict(synthetic.detector_state_dict(0))
    model.detector.set_precision("fp32")
    model.decoder.precision = "fp32"
    model.train()
    x = torch.from_numpy(gold["x"])
    fmask = torch.from_numpy(gold["fmask"])
    heat, feat = oracle_kernels.detection_train_forward(model.detector, x, sd_prob=0.0)
    f = feat.permute(0, 2, 3, 1).flatten(0, -2)
    dec = model.decoder(f[fmask])
    assert rel_l2(heat.detach(), gold["heatmap"]) < 1e-3     # fp32 reference through ~110 batch-statistics layers
    for i in range(3):
        assert rel_l2(dec[i].detach(), gold[f"dec{i}"]) < 1e-3
    loss = (heat * torch.from_numpy(gold["w_heat"])).sum() + sum((dec[i] * torch.from_numpy(gold[f"w_dec{i}"])).sum() for i in range(3))
    loss.backward()
    names = [str(n) for n in gold["grad_names"]]
    params = dict(model.named_parameters())
    assert set(names) == set(params)
    # truth = the reference in float64; grad_fp32_err[i] = || reference fp32 gradient - truth ||: the noise any fp32-storage
    # implementation carries (BatchNorm betas feeding another batch-statistics layer have ~zero true gradients)
    bad = []
    for i, n in enumerate(names):
        g = params[n].grad
        assert g is not None, n
        ref_norm, ref_dot, noise = float(gold["grad_norm"][i]), float(gold["grad_dot"][i]), float(gold["grad_fp32_err"][i])
        tol = 2e-3 * ref_norm + 4.0 * noise + 1e-9
        probe = torch.from_numpy(synthetic_probe(g.shape, i))
        e_norm = abs(float(g.double().norm()) - ref_norm)
        e_dot = abs(float((g.double() * probe).sum()) - ref_dot)
        if e_norm > tol or e_dot > 8.0 * tol:
            bad.append((n, e_norm, e_dot, tol))
    assert not bad, bad[:10]
    for k in [k for k in gold.files if k.startswith("full/")]:
        i = names.index(k[5:])
        tol = 2e-3 + 4.0 * float(gold["grad_fp32_err"][i]) / float(gold["grad_norm"][i])
        assert rel_l2(params[k[5:]].grad, gold[k]) < tol, (k, tol)
    bufs = dict(model.named_buffers())
    for k in [k for k in gold.files if k.startswith("buf/")]:
        assert rel_l2(bufs[k[4:]].double(), gold[k]) < 1e-5, k


def synthetic_probe(shape, i):
    """Deterministic +-1 probe tensor used to fingerprint a gradient with one dot product (same in make_golden_train.py)."""
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.int64)
    v = (((idx * 2654435761 + i * 40503) >> 7) & 1).astype(np.float64) * 2 - 1
    return v.reshape(shape)


# ---- Transformer pieces ---------------------------------------------------------------------------------------------
def test_layernorm_swiglu_embed_match_autograd():
    g = torch.Generator().manual_seed(3)
    d = 24
    x, r1, r2 = (torch.randn(3, 5, d, generator=g, requires_grad=True) for _ in range(3))
    gamma, beta = torch.randn(d, generator=g).requires_grad_(), torch.randn(d, generator=g).requires_grad_()
    y = F.layer_norm(x + r1 + r2, [d], gamma, beta, 1e-5)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    y0, xs, mean, rstd = TO.layernorm_train(x.detach(), gamma, beta, 1e-5, r1.detach(), r2.detach())
    assert rel_l2(y0, y.detach()) < 1e-6
    dx, dg, db = TO.layernorm_train_bwd(xs, dy, mean, rstd, gamma)
    assert rel_l2(dx, x.grad) < 1e-5 and rel_l2(dx, r2.grad) < 1e-5
    assert rel_l2(dg, gamma.grad) < 1e-5 and rel_l2(db, beta.grad) < 1e-5
    a, b = torch.randn(4, 7, generator=g, requires_grad=True), torch.randn(4, 7, generator=g, requires_grad=True)
    h = a * F.silu(b)
    dh = torch.randn(h.shape, generator=g)
    h.backward(dh)
    assert rel_l2(TO.swiglu(a.detach(), b.detach()), h.detach()) < 1e-6
    da, dbb = TO.swiglu_bwd(a.detach(), b.detach(), dh)
    assert rel_l2(da, a.grad) < 1e-6 and rel_l2(dbb, b.grad) < 1e-6
    tabs = [torch.randn(m, 6, generator=g, requires_grad=True) for m in (11, 13, 17)]
    tok = torch.randint(0, 5000, (3, 9), generator=g)
    e = sum(t[tok % t.shape[0]] for t in tabs)
    de = torch.randn(e.shape, generator=g)
    e.backward(de)
    assert rel_l2(TO.embed3(tok, tabs, torch.float32), e.detach()) < 1e-6
    for got, t in zip(TO.embed3_bwd(tok, de, (11, 13, 17)), tabs):
        assert rel_l2(got, t.grad) < 1e-6


@pytest.mark.parametrize("lt,ls,masked", [(5, 5, False), (7, 9, True)])
def test_attention_backward_matches_sdpa_autograd(lt, ls, masked):
    g = torch.Generator().manual_seed(lt)
    b, heads, hd = 2, 3, 8
    q = torch.randn(b, lt, heads * hd, generator=g, requires_grad=True)
    k = torch.randn(b, ls, heads * hd, generator=g, requires_grad=True)
    v = torch.randn(b, ls, heads * hd, generator=g, requires_grad=True)
    mask = None
    if masked:
        mask = torch.zeros(b, ls)
        mask[0, 6:] = float("-inf")
        mask[1, 3:] = float("-inf")
    sp = lambda t, l: t.view(b, l, heads, hd).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(q, lt), sp(k, ls), sp(v, ls), None if mask is None else mask[:, None, None, :])
    o = o.transpose(1, 2).reshape(b, lt, heads * hd)
    do = torch.randn(o.shape, generator=g)
    o.backward(do)
    assert rel_l2(TO.attention(q.detach(), k.detach(), v.detach(), heads, mask), o.detach()) < 1e-6
    dq, dk, dv = TO.attention_bwd(q.detach(), k.detach(), v.detach(), do, heads, mask)
    assert rel_l2(dq, q.grad) < 1e-5 and rel_l2(dk, k.grad) < 1e-5 and rel_l2(dv, v.grad) < 1e-5


TF_DIMS = dict(enc_input_dim=106, embed_dim=64, head_num=4, enc_block_num=2, dec_block_num=2, max_enc_seq_len=24,
               max_dec_seq_len=24)


def check_gradients_against_golden(gold, params, rel_tol):
    """Every parameter gradient vs the float64 reference fingerprints (L2 norm, +-1-probe dot product), tolerance =
    rel_tol * norm + 4 x the reference's own fp32 noise for that tensor; parameters the reference leaves without a gradient
    (norm -1) must have none here either."""
    names = [str(n) for n in gold["grad_names"]]
    assert set(names) == set(params)
    bad = []
    for i, n in enumerate(names):
        g = params[n].grad
        ref_norm, ref_dot, noise = float(gold["grad_norm"][i]), float(gold["grad_dot"][i]), float(gold["grad_fp32_err"][i])
        if ref_norm < 0:
            if g is not None and float(g.abs().max()) != 0.0:
                bad.append((n, "reference has no gradient"))
            continue
        if g is None:
            bad.append((n, "missing"))
            continue
        g = g.double().cpu()
        tol = rel_tol * ref_norm + 4.0 * noise + 1e-9
        e_norm = abs(float(g.norm()) - ref_norm)
        e_dot = abs(float((g * torch.from_numpy(synthetic_probe(g.shape, i))).sum()) - ref_dot)
        if e_norm > tol or e_dot > 8.0 * tol:
            bad.append((n, e_norm, e_dot, tol))
    assert not bad, (len(bad), bad[:10])
    for k in [k for k in gold.files if k.startswith("full/")]:
        i = names.index(k[5:])
        tol = rel_tol + 4.0 * float(gold["grad_fp32_err"][i]) / float(gold["grad_norm"][i])
        assert rel_l2(params[k[5:]].grad.cpu(), gold[k]) < tol, (k, tol)


def test_transformer_train_graph_matches_reference_golden(oracle_kernels):
    """Transformer.forward in train mode + backward through train_ops.py with oracle kernels == the unmodified reference in
    float64 (tests/golden/train_transformer_seed0.npz): logits and all 96 parameter gradients (incl. the learnable position
    tables and the three residue embeddings); pos_emb_k of the self-attention layers gets no gradient on either side."""
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.models.transformer import Transformer
    p = os.path.join(GOLDEN, "train_transformer_seed0.npz")
    if not os.path.exists(p):
        pytest.skip("golden missing (oracle/make_golden_train.py transformer)")
    gold = np.load(p)
    model = Transformer(**TF_DIMS, dropout=0.0)
    model.load_state_dict(synthetic.transformer_state_dict(0, **TF_DIMS))
    model.set_precision("fp32").train()
    outs = model(torch.from_numpy(gold["enc"]), torch.from_numpy(gold["dec"]))
    for i in range(3):
        assert rel_l2(outs[i].detach(), gold[f"out{i}"]) < 1e-5
    sum((o * torch.from_numpy(gold[f"w{i}"])).sum() for i, o in enumerate(outs)).backward()
    check_gradients_against_golden(gold, dict(model.named_parameters()), 1e-4)


def test_part_predictors_match_reference(oracle_kernels):
    """TransformerEncoderPredictor / TransformerDecoderPredictor(Splited) (export-side wrappers, models/transformer.py:362-404)
    layer by layer on the train kernels (oracle-backed here) vs the reference's own wrappers on the same fp32 model."""
    from findtextcenternet_b200 import synthetic
    from findtextcenternet_b200.models.transformer import (Transformer, TransformerDecoderPredictor,
                                                            TransformerDecoderPredictorSplited, TransformerEncoderPredictor)
    gold = np.load(os.path.join(GOLDEN, "train_transformer_seed0.npz"))
    model = Transformer(**TF_DIMS, dropout=0.0)
    model.load_state_dict(synthetic.transformer_state_dict(0, **TF_DIMS))
    model.eval()
    enc, dec = torch.from_numpy(gold["enc"]), torch.from_numpy(gold["dec"])
    km = torch.where(torch.all(enc == 0, dim=-1)[:, None, None, :], float("-inf"), 0)
    ep, dp, ds = (TransformerEncoderPredictor(model.encoder), TransformerDecoderPredictor(model.decoder),
                  TransformerDecoderPredictorSplited(model.decoder))
    for m in (ep, dp, ds):
        m.precision = "fp32"
    enc_out = ep(enc, km)
    assert rel_l2(enc_out, gold["pred_enc_out"]) < 1e-5
    probs = dp(enc_out, dec, km)
    split = ds(enc_out, dec % 1091, dec % 1093, dec % 1097, km)
    for i in range(3):
        assert rel_l2(probs[i].max(-1).values, gold[f"pred_probs{i}_max"]) < 1e-4
        assert torch.equal(probs[i].argmax(-1), torch.from_numpy(gold[f"pred_probs{i}_argmax"]))
        assert rel_l2(split[i], probs[i]) < 1e-6
        assert torch.allclose(probs[i].sum(-1), torch.ones(probs[i].shape[:-1]), atol=1e-5)
