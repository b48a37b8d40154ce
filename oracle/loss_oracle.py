"""ORACLE — test infrastructure only (never imported by the product path).

fp32 torch-CPU restatement of the reference losses: ``heatmap_loss`` loss_func.py:74-92, ``loss_function`` :94-177,
``loss_function3`` :179-213.  Written term by term from those lines (per-pixel formulas, explicit sums) so that it is an
independent statement of the arithmetic the CUDA kernels implement.  Pinned by tests/golden/loss_seed0.npz, which
oracle/make_golden.py produces by running the unmodified reference functions."""
from __future__ import annotations

import torch
import torch.nn.functional as F

MODULO = [1091, 1093, 1097]


def heatmap_losses(labelmap, idmap, heatmap):
    """-> dict of the eight map losses + weight1_count (loss_func.py:94-126)."""
    lab, hm = labelmap.float(), heatmap.float()
    key = lab[:, 0]
    x = hm[:, 0]
    p = torch.sigmoid(x)
    pos = key >= 1.0
    pos_loss = -F.logsigmoid(x) * (1 - p) ** 2
    neg_loss = (x + F.softplus(-x)) * p ** 2 * (1 - key) ** 4
    keymap = torch.where(pos, pos_loss, neg_loss).mean() * 10.
    m1 = key > 0.85
    w1 = torch.clamp_min(key - 0.85, 0) / (1 - 0.85)
    cnt = torch.clamp_min(w1[m1].sum(), 1.0)
    hub = F.huber_loss(hm[:, 1], lab[:, 1], reduction="none") + F.huber_loss(hm[:, 2], lab[:, 2], reduction="none")
    size = (hub * w1)[m1].sum() / cnt
    out = {"keymap_loss": keymap, "size_loss": size,
           "textline_loss": F.binary_cross_entropy_with_logits(hm[:, 3], lab[:, 3]),
           "separator_loss": F.binary_cross_entropy_with_logits(hm[:, 4], lab[:, 4]), "weight1_count": cnt}
    w2 = w1                                         # key_th2 == key_th1
    for i in range(4):
        y = ((idmap[:, 1] & (1 << i)) > 0).float()
        w = 1 + y * w2 + w2
        out["code%d_loss" % (1 << i)] = (w * F.binary_cross_entropy_with_logits(hm[:, 5 + i], y, reduction="none")).mean()
    return out


def ce_rows(logits, target, weight, select, count_select):
    """-> (sum w*ce over the three heads, sum w, #rows with three hits, #rows counted)."""
    ce = 0
    hits = 0
    for lg, m in zip(logits, MODULO):
        t = target % m
        ce = ce + (torch.logsumexp(lg.float(), -1) - lg.float().gather(-1, t[:, None])[:, 0])
        hits = hits + (lg.argmax(-1) == t).long()
    w = torch.ones_like(ce) if weight is None else weight
    # the two sums stay tensors: torch autograd over this oracle is the reference for the analytic backward kernels
    return (ce * w)[select].sum(), w[select].sum(), int((hits == 3)[count_select].sum()), int(count_select.sum())


def loss_function(fmask, labelmap, idmap, heatmap, decoder_outputs):
    out = heatmap_losses(labelmap, idmap, heatmap)
    keyv = labelmap[:, 0].flatten()[fmask].float()
    tid = idmap[:, 0].flatten()[fmask]
    w3 = torch.clamp_min(keyv - 0.99, 0) / (1 - 0.99)
    s, ws, c, n = ce_rows(decoder_outputs, tid, w3, (keyv > 0.99) & (tid > 0), (keyv == 1) & (tid > 0))
    out["id_loss"] = s / torch.clamp_min(ws, 1.0)
    out["loss"] = sum(out[k] for k in out if k.endswith("_loss"))
    out["correct"], out["total"] = c, n
    return out


def loss_function3(outputs, labelcode, mask):
    flat = [o.reshape(-1, o.shape[-1]) for o in outputs]
    m = mask.reshape(-1)
    s, ws, c, n = ce_rows(flat, labelcode.reshape(-1), None, m, m)
    return {"loss": s / ws, "correct": c, "total": n}
