"""Thin torch-tensor wrappers over the single-op C-ABI entry points (unit tests and building blocks).
NHWC activations; fp32 or bf16.  No CPU path."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


FUSED_RUNNING_STATS = True      # train_ops.batchnorm_act: BatchNorm running statistics updated inside ftc_train_bn_stats_running


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.DT_F32
    if t.dtype == torch.bfloat16:
        return _lib.DT_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _s(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def conv2d(x: torch.Tensor, w_oihw: torch.Tensor, stride: int = 1, scale=None, bias=None, act: int = _lib.ACT_NONE,
           residual=None, a_scale=None, backend: int = _lib.GEMM_SIMT) -> torch.Tensor:
    """x [B,H,W,Cin] NHWC -> [B,Ho,Wo,Cout]; w fp32 [Cout,Cin,k,k]; scale/bias fp32 [Cout]."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous()
    b, h, w, cin = x.shape
    cout, _, k, _ = w_oihw.shape
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    out = torch.empty(b, ho, wo, cout, dtype=x.dtype, device=x.device)
    nbytes = int(lib.ftc_op_conv2d_wpack_bytes(cin, cout, k))
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    w32 = w_oihw.to(device=x.device, dtype=torch.float32).contiguous()
    keep = [t if t is None else t.to(device=x.device, dtype=torch.float32).contiguous() for t in (scale, bias, a_scale)]
    res = None if residual is None else residual.contiguous()
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_op_conv2d(x.data_ptr(), _dt(x), b, h, w, cin, w32.data_ptr(), cout, k, stride, _p(keep[0]),
                                     _p(keep[1]), act, _p(res), _p(keep[2]), out.data_ptr(), wpack.data_ptr(), nbytes,
                                     backend, _s(x)), "ftc_op_conv2d")
    return out


def conv2d_dgrad_tc(dy: torch.Tensor, w_fwd_oihw: torch.Tensor) -> torch.Tensor:
    """Stride-1 data gradient on the tcgen05 kernel: dy bf16 [B,H,W,C] (C >= Cout, extra channels zero) and the FORWARD weight fp32
    [Cout,Cin,k,k] -> dx [B,H,W,Cin]; the transposed / rotated operand is formed by the weight-pack kernel."""
    lib = _lib.load()
    assert dy.is_cuda and dy.is_contiguous() and dy.dtype == torch.bfloat16
    b, h, w, c = dy.shape
    cout, cin, k, _ = w_fwd_oihw.shape
    w32 = w_fwd_oihw.detach().to(device=dy.device, dtype=torch.float32).contiguous()
    dx = torch.empty(b, h, w, cin, dtype=dy.dtype, device=dy.device)
    nbytes = int(lib.ftc_op_conv2d_wpack_bytes(c, cin, k))
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=dy.device)
    with torch.cuda.device(dy.device):
        _lib.check(lib.ftc_op_conv2d_dgrad(dy.data_ptr(), b, h, w, c, w32.data_ptr(), cout, cin, k, dx.data_ptr(), wpack.data_ptr(),
                                           nbytes, _s(dy)), "ftc_op_conv2d_dgrad")
    return dx


def dwconv3x3(x: torch.Tensor, w9c: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, stride: int = 1,
              want_se_sum: bool = False):
    """Depthwise 3x3 + BN + SiLU.  With ``want_se_sum`` also the per-tile spatial sums [B, tiles, C] fp32 (the SE squeeze,
    written without atomics; ``se_fc`` adds the tiles in order) -> (out, se_parts)."""
    lib = _lib.load()
    b, h, w, c = x.shape
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    out = torch.empty(b, ho, wo, c, dtype=x.dtype, device=x.device)
    parts = None
    if want_se_sum:
        nt = int(lib.ftc_op_dwconv3x3_tiles(h, w, stride, _dt(x)))
        parts = torch.empty(b, nt, c, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_op_dwconv3x3(x.data_ptr(), out.data_ptr(), _dt(x), b, h, w, c, stride, w9c.data_ptr(),
                                        scale.data_ptr(), bias.data_ptr(), _p(parts), _s(x)), "ftc_op_dwconv3x3")
    return (out, parts) if want_se_sum else out


def se_fc(se_parts: torch.Tensor, inv_hw: float, w1, b1, w2t, b2) -> torch.Tensor:
    """se_parts fp32 [B, tiles, C] (or [B, C]) spatial sums -> SE gate [B, C]."""
    lib = _lib.load()
    if se_parts.dim() == 2:
        se_parts = se_parts[:, None]
    se_parts = se_parts.contiguous()
    b, nt, c = se_parts.shape
    s = w1.shape[0]
    out = torch.empty(b, c, dtype=torch.float32, device=se_parts.device)
    hid = torch.empty(b, s, dtype=torch.float32, device=se_parts.device)
    with torch.cuda.device(se_parts.device):
        _lib.check(lib.ftc_op_se_fc(se_parts.data_ptr(), nt, out.data_ptr(), hid.data_ptr(), b, c, s, inv_hw, w1.data_ptr(),
                                    b1.data_ptr(), w2t.data_ptr(), b2.data_ptr(), _s(se_parts)), "ftc_op_se_fc")
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    b, h, w, c = x.shape
    out = torch.empty(b, 2 * h, 2 * w, c, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_op_upsample2x(x.data_ptr(), out.data_ptr(), _dt(x), b, h, w, c, _s(x)), "ftc_op_upsample2x")
    return out


def peak_pick(heat9: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    b, _, h, w = heat9.shape
    out = torch.empty(b, 10, h, w, dtype=torch.float32, device=heat9.device)
    with torch.cuda.device(heat9.device):
        _lib.check(lib.ftc_peak_pick(heat9.contiguous().data_ptr(), out.data_ptr(), b, h, w, _s(heat9)), "ftc_peak_pick")
    return out


def dwconv3x3_se(x: torch.Tensor, w9c, scale, bias, w1, b1, w2t, b2):
    """Stride-1 MBConv middle: depthwise 3x3 + BN + SiLU with folded SE squeeze/fc1, then fc2 -> (out, se_scale [B, C])."""
    lib = _lib.load()
    b, h, w, c = x.shape
    s = w1.shape[0]
    out = torch.empty_like(x)
    hid = torch.empty(b, c // 32, s, dtype=torch.float32, device=x.device)    # one fc1 share per 32-channel CTA
    sc = torch.empty(b, c, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_op_dwconv3x3_se(x.data_ptr(), out.data_ptr(), _dt(x), b, h, w, c, w9c.data_ptr(), scale.data_ptr(),
                                           bias.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2t.data_ptr(), b2.data_ptr(), s,
                                           hid.data_ptr(), sc.data_ptr(), _s(x)), "ftc_op_dwconv3x3_se")
    return out, sc


def head_top_conv(y: torch.Tensor, od, w_rows: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """y [B,H,W,n_heads*192] NHWC -> NCHW fp32 [B, sum(od), H, W]; w_rows fp32 [sum(od), 9*192] tap-major."""
    import ctypes
    lib = _lib.load()
    b, h, w, ct = y.shape
    arr = (ctypes.c_int * len(od))(*od)
    out = torch.empty(b, sum(od), h, w, dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _lib.check(lib.ftc_op_head_top_conv(y.data_ptr(), _dt(y), ct, len(od), arr, w_rows.data_ptr(), bias.data_ptr(),
                                            out.data_ptr(), b, h, w, _s(y)), "ftc_op_head_top_conv")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B, Lt, d], k/v [B, Ls, d] (contiguous) -> [B, Lt, d]; mask fp32 [B, Ls] additive."""
    lib = _lib.load()
    b, lt, d = q.shape
    ls = k.shape[1]
    out = torch.empty_like(q)
    with torch.cuda.device(q.device):
        _lib.check(lib.ftc_op_attention(q.data_ptr(), d, 0, k.data_ptr(), v.data_ptr(), d, 0, 0, _p(mask), out.data_ptr(), d,
                                        _dt(q), b, heads, d // heads, lt, ls, _s(q)), "ftc_op_attention")
    return out


# ---- train-step building blocks (csrc/train_ops.cu; include/ftc_b200.h "train step: forward in train mode + backward") ----
def _f32(t: torch.Tensor, dev) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32).contiguous()


def _reduce_scratch(rows: int, c: int, dev) -> torch.Tensor:
    n = int(_lib.load().ftc_train_reduce_scratch_bytes(rows, c))
    return torch.empty(n // 4, dtype=torch.float32, device=dev)


def bn_stats(x: torch.Tensor, running=None, momentum: float = 0.1):
    """x [..., C] contiguous -> (batch mean, biased batch variance) fp32 [C].  ``running`` = (running_mean, running_var,
    num_batches_tracked) fp32 / fp32 / int64 device tensors: updated in place by the same kernel (nn.BatchNorm train mode)."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous()
    c = x.shape[-1]
    rows = x.numel() // c
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    var = torch.empty(c, dtype=torch.float32, device=x.device)
    scratch = _reduce_scratch(rows, c, x.device)
    with torch.cuda.device(x.device):
        if running is None:
            _lib.check(lib.ftc_train_bn_stats(x.data_ptr(), _dt(x), rows, c, mean.data_ptr(), var.data_ptr(), scratch.data_ptr(),
                                              _s(x)), "ftc_train_bn_stats")
        else:
            rm, rv, nbt = running
            assert rm.dtype == torch.float32 and rv.dtype == torch.float32 and rm.is_contiguous() and rv.is_contiguous()
            assert nbt is None or nbt.dtype == torch.int64
            _lib.check(lib.ftc_train_bn_stats_running(x.data_ptr(), _dt(x), rows, c, mean.data_ptr(), var.data_ptr(), scratch.data_ptr(),
                                                      rm.data_ptr(), rv.data_ptr(), _p(nbt), momentum, _s(x)), "ftc_train_bn_stats_running")
            torch.autograd.graph.increment_version([t for t in (rm, rv, nbt) if t is not None])
    return mean, var


def bn_act(x, mean, var, gamma, beta, eps: float, act: int, residual=None) -> torch.Tensor:
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous()
    c = x.shape[-1]
    rows = x.numel() // c
    y = torch.empty_like(x)
    res = None if residual is None else residual.contiguous()
    g, b = _f32(gamma, x.device), _f32(beta, x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_train_bn_act(x.data_ptr(), y.data_ptr(), _dt(x), rows, c, mean.data_ptr(), var.data_ptr(),
                                        g.data_ptr(), b.data_ptr(), eps, act, _p(res), _s(x)), "ftc_train_bn_act")
    return y


def _channel_slice_ld(t: torch.Tensor, c: int) -> int:
    """Row stride of ``t`` when it is a channel slice ``wide[..., a:a+c]`` of a contiguous tensor (16-byte aligned), else 0."""
    st, sh = t.stride(), t.shape
    if t.dim() < 2 or st[-1] != 1 or sh[-1] != c:
        return 0
    ld = st[-2]
    if ld <= c or ld % 8 or t.data_ptr() % 16:
        return 0
    for i in range(t.dim() - 2):
        if st[i] != st[i + 1] * sh[i + 1]:
            return 0
    return int(ld)


def bn_act_bwd(x, dy, mean, var, gamma, beta, eps: float, act: int):
    """-> (dx, dgamma, dbeta).  A dy that is a channel slice of a wider map (torch.cat's backward) is read in place by the bf16
    stream kernels."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous()
    c = x.shape[-1]
    ld = c
    if not dy.is_contiguous():
        sl = _channel_slice_ld(dy, c) if (x.dtype == torch.bfloat16 and c % 32 == 0 and hasattr(lib, "ftc_train_bn_act_bwd_ld")) else 0
        if sl:
            ld = sl
        else:
            dy = dy.contiguous()
    rows = x.numel() // c
    dx = torch.empty_like(x)
    dbeta = torch.empty(c, dtype=torch.float32, device=x.device)
    dgamma = torch.empty(c, dtype=torch.float32, device=x.device)
    scratch = _reduce_scratch(rows, c, x.device)
    g, b = _f32(gamma, x.device), _f32(beta, x.device)
    with torch.cuda.device(x.device):
        if ld != c:
            _lib.check(lib.ftc_train_bn_act_bwd_ld(x.data_ptr(), dy.data_ptr(), ld, dx.data_ptr(), _dt(x), rows, c, mean.data_ptr(),
                                                   var.data_ptr(), g.data_ptr(), b.data_ptr(), eps, act, dbeta.data_ptr(),
                                                   dgamma.data_ptr(), scratch.data_ptr(), _s(x)), "ftc_train_bn_act_bwd_ld")
        else:
            _lib.check(lib.ftc_train_bn_act_bwd(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), _dt(x), rows, c, mean.data_ptr(),
                                                var.data_ptr(), g.data_ptr(), b.data_ptr(), eps, act, dbeta.data_ptr(),
                                                dgamma.data_ptr(), scratch.data_ptr(), _s(x)), "ftc_train_bn_act_bwd")
    return dx, dgamma, dbeta


def conv2d_wgrad(x: torch.Tensor, dy: torch.Tensor, ksize: int, stride: int = 1) -> torch.Tensor:
    """x [B,H,W,Cin], dy [B,Ho,Wo,Cout] NHWC -> dW fp32 [Cout,Cin,k,k]."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous() and dy.dtype == x.dtype
    dy = dy.contiguous()
    b, h, w, cin = x.shape
    cout = dy.shape[-1]
    assert dy.shape == (b, (h - 1) // stride + 1, (w - 1) // stride + 1, cout), (x.shape, dy.shape)
    dw = torch.empty(cout, cin, ksize, ksize, dtype=torch.float32, device=x.device)
    # bf16 shapes the tcgen05 kernel takes need scratch for its per-split partial tiles (0 bytes = not taken: CUDA-core / mma.sync
    # kernels of the plain entry point; the CPU kernel-emulation library has only that one)
    nbytes = 0
    if x.dtype == torch.bfloat16 and hasattr(lib, "ftc_train_conv2d_wgrad_scratch_bytes"):
        nbytes = int(lib.ftc_train_conv2d_wgrad_scratch_bytes(b, h, w, cin, cout, ksize, stride))
    with torch.cuda.device(x.device):
        if nbytes:
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
            _lib.check(lib.ftc_train_conv2d_wgrad_ws(x.data_ptr(), dy.data_ptr(), _dt(x), b, h, w, cin, cout, ksize, stride,
                                                     dw.data_ptr(), scratch.data_ptr(), nbytes, _s(x)), "ftc_train_conv2d_wgrad_ws")
        else:
            _lib.check(lib.ftc_train_conv2d_wgrad(x.data_ptr(), dy.data_ptr(), _dt(x), b, h, w, cin, cout, ksize, stride,
                                                  dw.data_ptr(), _s(x)), "ftc_train_conv2d_wgrad")
    return dw


def conv2d_dgrad(dy: torch.Tensor, w_oihw: torch.Tensor, h: int, w: int, stride: int = 1, add=None) -> torch.Tensor:
    """dy [B,Ho,Wo,Cout] NHWC, weights [Cout,Cin,k,k] -> dx [B,h,w,Cin] (+ add)."""
    lib = _lib.load()
    assert dy.is_cuda
    dy = dy.contiguous()
    b = dy.shape[0]
    cout, cin, k, _ = w_oihw.shape
    assert dy.shape == (b, (h - 1) // stride + 1, (w - 1) // stride + 1, cout), (dy.shape, w_oihw.shape, h, w)
    w32 = _f32(w_oihw, dy.device)
    dx = torch.empty(b, h, w, cin, dtype=dy.dtype, device=dy.device)
    ad = None if add is None else add.contiguous()
    with torch.cuda.device(dy.device):
        _lib.check(lib.ftc_train_conv2d_dgrad(dy.data_ptr(), _dt(dy), b, h, w, cin, cout, k, stride, w32.data_ptr(), _p(ad),
                                              dx.data_ptr(), _s(dy)), "ftc_train_conv2d_dgrad")
    return dx


def dwconv3x3_raw(x: torch.Tensor, w9c: torch.Tensor, stride: int = 1) -> torch.Tensor:
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous()
    b, h, w, c = x.shape
    out = torch.empty(b, (h - 1) // stride + 1, (w - 1) // stride + 1, c, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_train_dwconv3x3(x.data_ptr(), out.data_ptr(), _dt(x), b, h, w, c, stride, w9c.data_ptr(), _s(x)),
                   "ftc_train_dwconv3x3")
    return out


def dwconv3x3_dgrad(dy: torch.Tensor, w9c: torch.Tensor, h: int, w: int, stride: int = 1) -> torch.Tensor:
    lib = _lib.load()
    dy = dy.contiguous()
    b, _, _, c = dy.shape
    dx = torch.empty(b, h, w, c, dtype=dy.dtype, device=dy.device)
    with torch.cuda.device(dy.device):
        _lib.check(lib.ftc_train_dwconv3x3_dgrad(dy.data_ptr(), dx.data_ptr(), _dt(dy), b, h, w, c, stride, w9c.data_ptr(),
                                                 _s(dy)), "ftc_train_dwconv3x3_dgrad")
    return dx


def dwconv3x3_wgrad(x: torch.Tensor, dy: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """-> dW fp32 [9, C] tap-major."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous() and dy.dtype == x.dtype
    dy = dy.contiguous()
    b, h, w, c = x.shape
    dw = torch.empty(9, c, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_train_dwconv3x3_wgrad(x.data_ptr(), dy.data_ptr(), _dt(x), b, h, w, c, stride, dw.data_ptr(), _s(x)),
                   "ftc_train_dwconv3x3_wgrad")
    return dw


def spatial_sum(x: torch.Tensor, y: Optional[torch.Tensor] = None, scale: float = 1.0) -> torch.Tensor:
    """x (, y) [B, ..., C] -> fp32 [B, C] = scale * sum over the middle axes of x (* y)."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous()
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    yy = None if y is None else y.contiguous()
    out = torch.empty(b, c, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_train_spatial_sum(x.data_ptr(), _p(yy), _dt(x), b, hw, c, scale, out.data_ptr(), _s(x)),
                   "ftc_train_spatial_sum")
    return out


def scale_bc(x: torch.Tensor, scale: torch.Tensor, bias: Optional[torch.Tensor] = None, bias_mul: float = 0.0) -> torch.Tensor:
    """y = x * scale[b, c] (+ bias_mul * bias[b, c]); scale / bias fp32 [B, C]."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous() and scale.dtype == torch.float32 and scale.is_contiguous()
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_train_scale_bc(x.data_ptr(), scale.data_ptr(), _p(bias), bias_mul, y.data_ptr(), _dt(x), b, hw, c,
                                          _s(x)), "ftc_train_scale_bc")
    return y


def se_fc_train(mean: torch.Tensor, w1, b1, w2, b2):
    """mean fp32 [B, C]; W1 [S, C], W2 [C, S] -> (hid_pre [B, S], gate [B, C])."""
    lib = _lib.load()
    b, c = mean.shape
    s = w1.shape[0]
    hid = torch.empty(b, s, dtype=torch.float32, device=mean.device)
    gate = torch.empty(b, c, dtype=torch.float32, device=mean.device)
    with torch.cuda.device(mean.device):
        _lib.check(lib.ftc_train_se_fc(mean.data_ptr(), b, c, s, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                       hid.data_ptr(), gate.data_ptr(), _s(mean)), "ftc_train_se_fc")
    return hid, gate


def se_fc_train_bwd(dgate, gate, hid_pre, mean, w1, w2):
    """-> (dmean [B, C], dW1 [S, C], db1 [S], dW2 [C, S], db2 [C])"""
    lib = _lib.load()
    b, c = mean.shape
    s = w1.shape[0]
    dev = mean.device
    f = dict(dtype=torch.float32, device=dev)
    dgp, dhp, dmean = torch.empty(b, c, **f), torch.empty(b, s, **f), torch.empty(b, c, **f)
    dw1, db1, dw2, db2 = torch.empty(s, c, **f), torch.empty(s, **f), torch.empty(c, s, **f), torch.empty(c, **f)
    dgate = dgate.contiguous()
    with torch.cuda.device(dev):
        _lib.check(lib.ftc_train_se_fc_bwd(dgate.data_ptr(), gate.data_ptr(), hid_pre.data_ptr(), mean.data_ptr(), b, c, s,
                                           w1.data_ptr(), w2.data_ptr(), dgp.data_ptr(), dhp.data_ptr(), dmean.data_ptr(),
                                           dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(), _s(mean)),
                   "ftc_train_se_fc_bwd")
    return dmean, dw1, db1, dw2, db2


def upsample2x_bwd(dy: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    b, ho, wo, c = dy.shape
    ld = c
    if not dy.is_contiguous():
        sl = _channel_slice_ld(dy, c) if hasattr(lib, "ftc_train_upsample2x_bwd_ld") else 0
        if sl:
            ld = sl             # channel slice of a wider map (torch.cat's backward): read in place
        else:
            dy = dy.contiguous()
    dx = torch.empty(b, ho // 2, wo // 2, c, dtype=dy.dtype, device=dy.device)
    with torch.cuda.device(dy.device):
        if ld != c:
            _lib.check(lib.ftc_train_upsample2x_bwd_ld(dy.data_ptr(), ld, dx.data_ptr(), _dt(dy), b, ho // 2, wo // 2, c, _s(dy)),
                       "ftc_train_upsample2x_bwd_ld")
        else:
            _lib.check(lib.ftc_train_upsample2x_bwd(dy.data_ptr(), dx.data_ptr(), _dt(dy), b, ho // 2, wo // 2, c, _s(dy)),
                       "ftc_train_upsample2x_bwd")
    return dx


# ---- Transformer train-step pieces ----
def layernorm_train(x, gamma, beta, eps: float = 1e-5, r1=None, r2=None):
    """y = LN(x (+ r1) (+ r2)); -> (y, xs, mean, rstd) with xs the summed input (x itself without residuals)."""
    lib = _lib.load()
    assert x.is_cuda and x.is_contiguous()
    d = x.shape[-1]
    rows = x.numel() // d
    y = torch.empty_like(x)
    has_res = r1 is not None or r2 is not None
    xs = torch.empty_like(x) if has_res else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    g, b = _f32(gamma, x.device), _f32(beta, x.device)
    a1 = None if r1 is None else r1.contiguous()
    a2 = None if r2 is None else r2.contiguous()
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_train_layernorm(x.data_ptr(), _p(a1), _p(a2), _p(xs), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                           _dt(x), rows, d, g.data_ptr(), b.data_ptr(), eps, _s(x)), "ftc_train_layernorm")
    return y, (xs if has_res else x), mean, rstd


def layernorm_train_bwd(xs, dy, mean, rstd, gamma):
    """-> (dx, dgamma, dbeta)"""
    lib = _lib.load()
    dy = dy.contiguous()
    d = xs.shape[-1]
    rows = xs.numel() // d
    dx = torch.empty_like(xs)
    dgamma = torch.empty(d, dtype=torch.float32, device=xs.device)
    dbeta = torch.empty(d, dtype=torch.float32, device=xs.device)
    scratch = _reduce_scratch(rows, d, xs.device)
    g = _f32(gamma, xs.device)
    with torch.cuda.device(xs.device):
        _lib.check(lib.ftc_train_layernorm_bwd(xs.data_ptr(), dy.data_ptr(), dx.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                               _dt(xs), rows, d, g.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                               scratch.data_ptr(), _s(xs)), "ftc_train_layernorm_bwd")
    return dx, dgamma, dbeta


def swiglu(x1, xg):
    lib = _lib.load()
    assert x1.is_cuda and x1.is_contiguous() and xg.is_contiguous() and x1.dtype == xg.dtype
    h = torch.empty_like(x1)
    with torch.cuda.device(x1.device):
        _lib.check(lib.ftc_train_swiglu(x1.data_ptr(), xg.data_ptr(), h.data_ptr(), _dt(x1), x1.numel(), _s(x1)), "ftc_train_swiglu")
    return h


def swiglu_bwd(x1, xg, dh):
    lib = _lib.load()
    dh = dh.contiguous()
    dx1, dxg = torch.empty_like(x1), torch.empty_like(xg)
    with torch.cuda.device(x1.device):
        _lib.check(lib.ftc_train_swiglu_bwd(x1.data_ptr(), xg.data_ptr(), dh.data_ptr(), dx1.data_ptr(), dxg.data_ptr(), _dt(x1),
                                            x1.numel(), _s(x1)), "ftc_train_swiglu_bwd")
    return dx1, dxg


def embed3(tokens, tables, dtype):
    """tokens int64 [...]; tables: three fp32 [m_i, D] -> [..., D] (dtype) = sum_i tables[i][tokens mod m_i]."""
    lib = _lib.load()
    assert tokens.is_cuda and tokens.dtype == torch.int64
    tok = tokens.contiguous()
    t = [_f32(e, tok.device) for e in tables]
    d = t[0].shape[1]
    out = torch.empty(*tok.shape, d, dtype=dtype, device=tok.device)
    with torch.cuda.device(tok.device):
        _lib.check(lib.ftc_train_embed3(tok.data_ptr(), t[0].data_ptr(), t[1].data_ptr(), t[2].data_ptr(), t[0].shape[0],
                                        t[1].shape[0], t[2].shape[0], out.data_ptr(), _dt(out), tok.numel(), d, _s(tok)),
                   "ftc_train_embed3")
    return out


def embed3_bwd(tokens, dy, ms):
    """-> three fp32 [m_i, D] table gradients"""
    lib = _lib.load()
    tok = tokens.contiguous()
    dy = dy.contiguous()
    d = dy.shape[-1]
    outs = [torch.empty(m, d, dtype=torch.float32, device=dy.device) for m in ms]
    with torch.cuda.device(dy.device):
        _lib.check(lib.ftc_train_embed3_bwd(tok.data_ptr(), dy.data_ptr(), _dt(dy), tok.numel(), d, ms[0], ms[1], ms[2],
                                            outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), _s(dy)),
                   "ftc_train_embed3_bwd")
    return outs


def attention_bwd(q, k, v, dout, heads: int, mask=None):
    """q / dout [B, Lt, D], k / v [B, Ls, D] contiguous, mask fp32 [B, Ls] additive -> (dq, dk, dv) fp32."""
    lib = _lib.load()
    assert q.is_cuda and q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    dout = dout.contiguous()
    b, lt, d = q.shape
    ls = k.shape[1]
    f = dict(dtype=torch.float32, device=q.device)
    dq, dk, dv = torch.empty(b, lt, d, **f), torch.empty(b, ls, d, **f), torch.empty(b, ls, d, **f)
    scratch = torch.empty(int(lib.ftc_train_attention_bwd_scratch_bytes(b, heads, lt, ls)) // 4, **f)
    with torch.cuda.device(q.device):
        _lib.check(lib.ftc_train_attention_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), _p(mask), dout.data_ptr(), dq.data_ptr(),
                                               dk.data_ptr(), dv.data_ptr(), scratch.data_ptr(), _dt(q), b, heads, d // heads, lt, ls,
                                               _s(q)), "ftc_train_attention_bwd")
    return dq, dk, dv
