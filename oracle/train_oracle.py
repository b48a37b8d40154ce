"""CPU oracle of the train-step building blocks (TEST INFRASTRUCTURE: imported by tests/ only, never by the product path).

Each function restates, with explicit formulas on CPU tensors (no autograd inside), the kernel of the same name in
``findtextcenternet_b200/_ops.py`` / ``csrc/train_ops.cu`` -- same signatures, NHWC layout -- so that
(a) ``tests/test_train_oracle.py`` can pin every formula to the reference's semantics (torch autograd over the modules the
    reference differentiates: nn.Conv2d, nn.BatchNorm2d in train mode, SiLU / GELU, torchvision SqueezeExcitation,
    nn.UpsamplingBilinear2d; models/detector.py:148-254, torchvision efficientnet.py:105-231, ops/misc.py:225-261),
(b) the host-side autograd graph of ``findtextcenternet_b200/train_ops.py`` can be checked on CPU against the unmodified
    reference model (``train_ops.K`` swapped for this module), and
(c) the GPU tests compare each CUDA kernel with the function here on the same seeded inputs.

Parity pin: reference-generated goldens (``oracle/make_golden_train.py`` -> ``tests/golden/train_*.npz``) plus direct
comparison with torch autograd in the CPU tests; the reference owns no tests or fixtures of its own.
"""
from __future__ import annotations

import math

import torch

ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2


def _act(z, act):
    if act == ACT_SILU:
        return z * torch.sigmoid(z)
    if act == ACT_GELU:
        return 0.5 * z * (1.0 + torch.erf(z / math.sqrt(2.0)))
    return z


def _act_grad(z, act):
    if act == ACT_SILU:
        s = torch.sigmoid(z)
        return s * (1.0 + z * (1.0 - s))
    if act == ACT_GELU:
        cdf = 0.5 * (1.0 + torch.erf(z / math.sqrt(2.0)))
        pdf = torch.exp(-0.5 * z * z) / math.sqrt(2.0 * math.pi)
        return cdf + z * pdf
    return torch.ones_like(z)


def _out_hw(h, w, stride):
    return (h - 1) // stride + 1, (w - 1) // stride + 1


def _taps(x, k, stride):
    """x [B,H,W,C] -> list over (ky,kx) of the strided, zero-padded views [B,Ho,Wo,C] a k x k / pad (k-1)//2 conv reads."""
    b, h, w, c = x.shape
    p = (k - 1) // 2
    ho, wo = _out_hw(h, w, stride)
    xp = torch.zeros(b, h + 2 * p + stride, w + 2 * p + stride, c, dtype=x.dtype)
    xp[:, p:p + h, p:p + w] = x
    return [xp[:, ky:ky + (ho - 1) * stride + 1:stride, kx:kx + (wo - 1) * stride + 1:stride] for ky in range(k) for kx in range(k)]


# ---- dense convolution (nn.Conv2d, padding (k-1)//2) -------------------------------------------------------------------
def conv2d(x, w_oihw, stride=1, scale=None, bias=None, act=ACT_NONE, residual=None, a_scale=None, backend=0):
    dt = x.dtype
    x = x.double()
    w = w_oihw.detach().double()
    k = w.shape[-1]
    if a_scale is not None:
        x = x * a_scale.double()[:, None, None, :]
    y = 0
    for t, v in enumerate(_taps(x, k, stride)):
        y = y + v @ w[:, :, t // k, t % k].t()
    if scale is not None:
        y = y * scale.double()
    if bias is not None:
        y = y + bias.detach().double()
    y = _act(y, act)
    if residual is not None:
        y = y + residual.double()
    return y.to(dt)


def conv2d_wgrad(x, dy, ksize, stride=1):
    """dW[co,ci,ky,kx] = sum_{b,oy,ox} dy[b,oy,ox,co] x[b, oy*s-p+ky, ox*s-p+kx, ci]"""
    cin, cout = x.shape[-1], dy.shape[-1]
    dw = torch.zeros(cout, cin, ksize, ksize, dtype=torch.float64)
    d2 = dy.double().reshape(-1, cout)
    for t, v in enumerate(_taps(x.double(), ksize, stride)):
        dw[:, :, t // ksize, t % ksize] = d2.t() @ v.reshape(-1, cin)
    return dw.float()


def conv2d_dgrad(dy, w_oihw, h, w, stride=1, add=None):
    """dX[b, oy*s-p+ky, ox*s-p+kx, ci] += sum_co dy[b,oy,ox,co] W[co,ci,ky,kx]   (scatter form of the kernel's gather)"""
    dt = dy.dtype
    wd = w_oihw.detach().double()
    cout, cin, k, _ = wd.shape
    b = dy.shape[0]
    p = (k - 1) // 2
    ho, wo = _out_hw(h, w, stride)
    acc = torch.zeros(b, h + 2 * p + stride, w + 2 * p + stride, cin, dtype=torch.float64)
    d = dy.double()
    for ky in range(k):
        for kx in range(k):
            acc[:, ky:ky + (ho - 1) * stride + 1:stride, kx:kx + (wo - 1) * stride + 1:stride] += d @ wd[:, :, ky, kx]
    dx = acc[:, p:p + h, p:p + w]
    if add is not None:
        dx = dx + add.double()
    return dx.to(dt).contiguous()


# ---- BatchNorm (train mode) + activation -------------------------------------------------------------------------------
def bn_stats(x):
    x2 = x.double().reshape(-1, x.shape[-1])
    mean = x2.mean(0)
    var = (x2 * x2).mean(0) - mean * mean
    return mean.float(), var.clamp_min(0).float()


def bn_act(x, mean, var, gamma, beta, eps, act, residual=None):
    xh = (x.double() - mean.double()) / torch.sqrt(var.double() + eps)
    y = _act(gamma.detach().double() * xh + beta.detach().double(), act)
    if residual is not None:
        y = y + residual.double()
    return y.to(x.dtype)


def bn_act_bwd(x, dy, mean, var, gamma, beta, eps, act):
    """dz = dy act'(z); dbeta = sum dz; dgamma = sum dz xhat; dx = gamma rstd (dz - dbeta/n - xhat dgamma/n)"""
    c = x.shape[-1]
    n = x.numel() // c
    rstd = 1.0 / torch.sqrt(var.double() + eps)
    g = gamma.detach().double()
    xh = (x.double() - mean.double()) * rstd
    dz = dy.double() * _act_grad(g * xh + beta.detach().double(), act)
    dbeta = dz.reshape(-1, c).sum(0)
    dgamma = (dz * xh).reshape(-1, c).sum(0)
    dx = g * rstd * (dz - dbeta / n - xh * dgamma / n)
    return dx.to(x.dtype), dgamma.float(), dbeta.float()


# ---- depthwise 3x3, weights [9, C] tap-major ---------------------------------------------------------------------------
def dwconv3x3_raw(x, w9c, stride=1):
    y = 0
    for t, v in enumerate(_taps(x.double(), 3, stride)):
        y = y + v * w9c[t].double()
    return y.to(x.dtype)


def dwconv3x3_dgrad(dy, w9c, h, w, stride=1):
    b, _, _, c = dy.shape
    ho, wo = _out_hw(h, w, stride)
    acc = torch.zeros(b, h + 2 + stride, w + 2 + stride, c, dtype=torch.float64)
    d = dy.double()
    for ky in range(3):
        for kx in range(3):
            acc[:, ky:ky + (ho - 1) * stride + 1:stride, kx:kx + (wo - 1) * stride + 1:stride] += d * w9c[ky * 3 + kx].double()
    return acc[:, 1:1 + h, 1:1 + w].to(dy.dtype).contiguous()


def dwconv3x3_wgrad(x, dy, stride=1):
    c = x.shape[-1]
    d = dy.double()
    return torch.stack([(v * d).reshape(-1, c).sum(0) for v in _taps(x.double(), 3, stride)]).float()


# ---- squeeze-excitation ------------------------------------------------------------------------------------------------
def spatial_sum(x, y=None, scale=1.0):
    b, c = x.shape[0], x.shape[-1]
    v = x.double() if y is None else x.double() * y.double()
    return (v.reshape(b, -1, c).sum(1) * scale).float()


def scale_bc(x, scale, bias=None, bias_mul=0.0):
    b, c = x.shape[0], x.shape[-1]
    shp = [b] + [1] * (x.dim() - 2) + [c]
    y = x.double() * scale.double().reshape(shp)
    if bias is not None:
        y = y + bias_mul * bias.double().reshape(shp)
    return y.to(x.dtype)


def se_fc_train(mean, w1, b1, w2, b2):
    hid_pre = mean.double() @ w1.double().t() + b1.double()
    gate = torch.sigmoid(_act(hid_pre, ACT_SILU) @ w2.double().t() + b2.double())
    return hid_pre.float(), gate.float()


def se_fc_train_bwd(dgate, gate, hid_pre, mean, w1, w2):
    g = gate.double()
    dgp = dgate.double() * g * (1.0 - g)
    hid = _act(hid_pre.double(), ACT_SILU)
    dhp = (dgp @ w2.double()) * _act_grad(hid_pre.double(), ACT_SILU)
    dmean = dhp @ w1.double()
    return (dmean.float(), (dhp.t() @ mean.double()).float(), dhp.sum(0).float(), (dgp.t() @ hid).float(), dgp.sum(0).float())


# ---- bilinear x2, align_corners=True -----------------------------------------------------------------------------------
def _interp_matrix(n):
    """[2n, n] matrix of the 1-D interpolation: source coordinate o * (n-1)/(2n-1)."""
    m = torch.zeros(2 * n, n, dtype=torch.float64)
    for o in range(2 * n):
        f = o * (n - 1) / (2 * n - 1) if n > 1 else 0.0
        i0 = min(int(f), n - 1)
        i1 = min(i0 + 1, n - 1)
        l = min(max(f - i0, 0.0), 1.0)
        m[o, i0] += 1.0 - l
        m[o, i1] += l
    return m


def upsample2x(x):
    my, mx = _interp_matrix(x.shape[1]), _interp_matrix(x.shape[2])
    return torch.einsum("oh,pw,bhwc->bopc", my, mx, x.double()).to(x.dtype).contiguous()


def upsample2x_bwd(dy):
    my, mx = _interp_matrix(dy.shape[1] // 2), _interp_matrix(dy.shape[2] // 2)
    return torch.einsum("oh,pw,bopc->bhwc", my, mx, dy.double()).to(dy.dtype).contiguous()


# ---- Transformer pieces (models/transformer.py:58-253) -----------------------------------------------------------------
def layernorm_train(x, gamma, beta, eps=1e-5, r1=None, r2=None):
    xs = x.double()
    if r1 is not None:
        xs = xs + r1.double()
    if r2 is not None:
        xs = xs + r2.double()
    xs = xs.to(x.dtype)                       # the kernel stores the sum in the activation dtype and normalises that
    v = xs.double()
    mean = v.mean(-1)
    rstd = 1.0 / torch.sqrt(((v - mean[..., None]) ** 2).mean(-1) + eps)
    y = (v - mean[..., None]) * rstd[..., None] * gamma.detach().double() + beta.detach().double()
    return y.to(x.dtype), xs, mean.reshape(-1).float(), rstd.reshape(-1).float()


def layernorm_train_bwd(xs, dy, mean, rstd, gamma):
    d = xs.shape[-1]
    v = xs.double().reshape(-1, d)
    xh = (v - mean.double()[:, None]) * rstd.double()[:, None]
    g = dy.double().reshape(-1, d) * gamma.detach().double()
    dx = rstd.double()[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    dyf = dy.double().reshape(-1, d)
    return dx.reshape(xs.shape).to(xs.dtype), (dyf * xh).sum(0).float(), dyf.sum(0).float()


def swiglu(x1, xg):
    return (x1.double() * _act(xg.double(), ACT_SILU)).to(x1.dtype)


def swiglu_bwd(x1, xg, dh):
    d = dh.double()
    return (d * _act(xg.double(), ACT_SILU)).to(x1.dtype), (d * x1.double() * _act_grad(xg.double(), ACT_SILU)).to(x1.dtype)


def embed3(tokens, tables, dtype):
    out = 0
    for e in tables:
        out = out + e.detach().double()[tokens % e.shape[0]]
    return out.to(dtype)


def embed3_bwd(tokens, dy, ms):
    d = dy.shape[-1]
    outs = []
    for m in ms:
        g = torch.zeros(m, d, dtype=torch.float64)
        g.index_add_(0, (tokens % m).reshape(-1), dy.double().reshape(-1, d))
        outs.append(g.float())
    return outs


def _heads(t, heads):
    b, l, d = t.shape
    return t.double().reshape(b, l, heads, d // heads).permute(0, 2, 1, 3)


def _probs(q, k, heads, mask):
    qh, kh = _heads(q, heads), _heads(k, heads)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(qh.shape[-1])
    if mask is not None:
        s = s + mask.double()[:, None, None, :]
    return torch.softmax(s, -1), qh, kh


def attention(q, k, v, heads, mask=None):
    p, _, _ = _probs(q, k, heads, mask)
    o = p @ _heads(v, heads)
    return o.permute(0, 2, 1, 3).reshape(q.shape).to(q.dtype)


def attention_bwd(q, k, v, dout, heads, mask=None):
    """dV = P^T dO; dP = dO V^T; dS = P (dP - rowsum(P dP)); dQ = dS K / sqrt(hd); dK = dS^T Q / sqrt(hd)"""
    p, qh, kh = _probs(q, k, heads, mask)
    vh, doh = _heads(v, heads), _heads(dout, heads)
    dv = p.transpose(-1, -2) @ doh
    dp = doh @ vh.transpose(-1, -2)
    ds = p * (dp - (p * dp).sum(-1, keepdim=True))
    sc = 1.0 / math.sqrt(qh.shape[-1])
    dq, dk = ds @ kh * sc, ds.transpose(-1, -2) @ qh * sc
    back = lambda t, ref: t.permute(0, 2, 1, 3).reshape(ref.shape).float()
    return back(dq, q), back(dk, k), back(dv, v)
