"""Thin torch-tensor wrappers over the single-op C-ABI entry points (unit tests and building blocks).
NHWC activations; fp32 or bf16.  No CPU path."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


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


def dwconv3x3(x: torch.Tensor, w9c: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, stride: int = 1,
              se_sum: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    b, h, w, c = x.shape
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    out = torch.empty(b, ho, wo, c, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.ftc_op_dwconv3x3(x.data_ptr(), out.data_ptr(), _dt(x), b, h, w, c, stride, w9c.data_ptr(),
                                        scale.data_ptr(), bias.data_ptr(), _p(se_sum), _s(x)), "ftc_op_dwconv3x3")
    return out


def se_fc(se_sum: torch.Tensor, inv_hw: float, w1, b1, w2t, b2) -> torch.Tensor:
    lib = _lib.load()
    b, c = se_sum.shape
    s = w1.shape[0]
    out = torch.empty(b, c, dtype=torch.float32, device=se_sum.device)
    hid = torch.empty(b, s, dtype=torch.float32, device=se_sum.device)
    with torch.cuda.device(se_sum.device):
        _lib.check(lib.ftc_op_se_fc(se_sum.data_ptr(), out.data_ptr(), hid.data_ptr(), b, c, s, inv_hw, w1.data_ptr(), b1.data_ptr(),
                                    w2t.data_ptr(), b2.data_ptr(), _s(se_sum)), "ftc_op_se_fc")
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
    hid = torch.zeros(b, s, dtype=torch.float32, device=x.device)
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
