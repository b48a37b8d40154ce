"""ORACLE — test infrastructure only (never imported by the product path).

CPU fp32 restatement of the reference Transformer hot path as functional torch ops over a
``state_dict``:

  * ``mha``               <- models/transformer.py:99-137 (MultiheadAttn.forward)
  * ``swiglu``            <- :66-71
  * ``encoder_forward``   <- :149-160, :173-180
  * ``decoder_forward``   <- :195-211, :225-238
  * ``transformer_forward`` <- :248-253
  * ``predictor_forward`` <- :274-360 (TransformerPredictor.forward, mask-predict loop)
  * ``calc_predid_np``    <- util_func.py:92-126 (CRT over moduli 1091/1093/1097), int64 numpy

Pinned against the unmodified reference by tests/golden/transformer_*.npz (oracle/make_golden.py).
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from findtextcenternet_b200 import arch

M1, M2, M3 = arch.MODULO_LIST
M_ALL = M1 * M2 * M3   # 1 308 131 911


def calc_predid_np(b1, b2, b3):
    """Garner CRT: the unique x in [0, m1*m2*m3) with x = b_i (mod m_i).  int64 arrays."""
    b1 = np.asarray(b1, dtype=np.int64); b2 = np.asarray(b2, dtype=np.int64); b3 = np.asarray(b3, dtype=np.int64)
    inv12 = pow(M1, M2 - 2, M2)
    inv13 = pow(M1, M3 - 2, M3)
    inv23 = pow(M2, M3 - 2, M3)
    t0 = b1 % M1
    t1 = ((b2 - t0) % M2) * inv12 % M2
    t2 = ((b3 - (t0 + t1 * M1)) % M3) * inv13 % M3 * inv23 % M3
    return (t0 + t1 * M1 + t2 * M1 * M2) % M_ALL


def calc_predid_t(b1, b2, b3):
    return torch.from_numpy(calc_predid_np(b1.numpy(), b2.numpy(), b3.numpy()))


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def mha(sd, p, heads, query, key=None, key_mask=None):
    if key is None:
        key = query
        pk = sd[p + ".pos_emb_q.encoding"]
    else:
        pk = sd[p + ".pos_emb_k.encoding"]
    value = key
    b, lt, d = query.shape
    ls = key.shape[1]
    hd = d // heads
    q = F.linear(query + sd[p + ".pos_emb_q.encoding"][:lt], sd[p + ".q_proj.weight"])
    k = F.linear(key + pk[:ls], sd[p + ".k_proj.weight"])
    v = F.linear(value, sd[p + ".v_proj.weight"])
    q = q.view(b, lt, heads, hd).transpose(1, 2)
    k = k.view(b, ls, heads, hd).transpose(1, 2)
    v = v.view(b, ls, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if key_mask is not None:
        s = s + key_mask[:, :, :, :ls]
    a = torch.softmax(s, dim=-1) @ v
    a = a.transpose(1, 2).reshape(b, lt, d)
    return F.linear(a, sd[p + ".out_proj.weight"])


def swiglu(sd, p, x):
    x1 = F.linear(x, sd[p + ".w1.weight"], sd[p + ".w1.bias"])
    xg = F.silu(F.linear(x, sd[p + ".wg.weight"], sd[p + ".wg.bias"]))
    return F.linear(x1 * xg, sd[p + ".w2.weight"], sd[p + ".w2.bias"])


def key_mask_of(enc_input):
    km = torch.all(enc_input == 0, dim=-1)
    return torch.where(km[:, None, None, :], float("-inf"), 0.0)


def encoder_forward(sd, heads, x, key_mask, n_blocks):
    x = F.linear(x, sd["encoder.embed.weight"])
    x = x + sd["encoder.pos_emb.encoding"][: x.shape[1]]
    x = _ln(sd, "encoder.norm", x)
    for i in range(n_blocks):
        p = f"encoder.blocks.{i}"
        skip = x
        x = _ln(sd, p + ".norm1", mha(sd, p + ".mha", heads, x, key_mask=key_mask) + skip)
        _x = x
        x = _ln(sd, p + ".norm2", swiglu(sd, p + ".ff", x) + _x + skip)
    return x


def decoder_forward(sd, heads, tokens, enc_out, key_mask, n_blocks):
    x = None
    for i, m in enumerate(arch.MODULO_LIST):
        e = F.embedding(tokens % m, sd[f"decoder.embed.{i}.weight"])
        x = e if x is None else x + e
    x = x + sd["decoder.pos_emb.encoding"][: x.shape[1]]
    x = _ln(sd, "decoder.norm", x)
    for i in range(n_blocks):
        p = f"decoder.blocks.{i}"
        skip = x
        x = _ln(sd, p + ".norm1", mha(sd, p + ".self_attn", heads, x) + skip)
        _x = x
        x = _ln(sd, p + ".norm2", mha(sd, p + ".cross_attn", heads, x, enc_out, key_mask) + _x)
        _x = x
        x = _ln(sd, p + ".norm3", swiglu(sd, p + ".ff", x) + _x + skip)
    return [F.linear(x, sd[f"decoder.out_layers.{i}.weight"], sd[f"decoder.out_layers.{i}.bias"])
            for i in range(len(arch.MODULO_LIST))]


def _count(sd, prefix):
    n = 0
    while f"{prefix}.{n}.norm1.weight" in sd:
        n += 1
    return n


def transformer_forward(sd, heads, enc_input, dec_input):
    with torch.no_grad():
        km = key_mask_of(enc_input)
        enc = encoder_forward(sd, heads, enc_input, km, _count(sd, "encoder.blocks"))
        return decoder_forward(sd, heads, dec_input, enc, km, _count(sd, "decoder.blocks"))


def mask_predict_step(outputs: List[torch.Tensor]):
    """One decoder pass's logits -> (decoder_output int64 [B,L], pred_p [B,L]); :311-324."""
    listp, listi = [], []
    for o in outputs:
        p = torch.softmax(o, dim=-1)
        tp, ti = torch.topk(p, 3)
        listp.append(tp.permute(2, 0, 1))
        listi.append(ti.permute(2, 0, 1))
    ids = torch.stack([torch.stack(x) for x in itertools.product(*listi)]).transpose(0, 1)   # [3,27,B,L]
    pp = torch.stack([torch.stack(x) for x in itertools.product(*listp)]).transpose(0, 1)
    pp = pp.clamp_min(1e-10).log().mean(dim=0).exp()                                       # [27,B,L]
    out = calc_predid_t(*ids)
    pp[out > 0x3FFFF] = 0
    maxi = torch.argmax(pp, dim=0)
    out = torch.gather(out, 0, maxi.unsqueeze(0))[0]
    pp = torch.gather(pp, 0, maxi.unsqueeze(0))[0]
    return out, pp


def predictor_forward(sd, heads, enc_input, max_decoderlen=arch.MAX_DECODERLEN, rep_count=8, trace=None):
    """TransformerPredictor.forward; returns int64 [B, max_decoderlen].  ``trace`` (list) collects
    (k, decoder_output, pred_p) per pass."""
    with torch.no_grad():
        km = key_mask_of(enc_input)
        ne, nd = _count(sd, "encoder.blocks"), _count(sd, "decoder.blocks")
        enc = encoder_forward(sd, heads, enc_input, km, ne)
        dec_in = torch.full((enc_input.shape[0], max_decoderlen), arch.DECODER_MSK, dtype=torch.long)
        out = None
        for k in range(rep_count):
            outputs = decoder_forward(sd, heads, dec_in, enc, km, nd)
            out, pp = mask_predict_step(outputs)
            if trace is not None:
                trace.append((k, out.clone(), pp.clone()))
            if torch.all(pp[torch.logical_and(dec_in == arch.DECODER_MSK, out > 0)] > 0.99):
                break
            if k < rep_count - 1:
                remask = torch.logical_or(pp < 0.9, out > 0x3FFFF)
                if not torch.any(remask):
                    break
                dec_in = torch.where(remask, arch.DECODER_MSK, out)
        return out
