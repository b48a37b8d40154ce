"""Train-mode forward + backward of the detector (train1.py:128-170 calls ``model(image, fmask)`` in train mode and
``loss.backward()``): each autograd node is one C-ABI kernel call of include/ftc_b200.h ("train step" section,
csrc/train_ops.cu) or one of the shared implicit-GEMM convolution kernels; torch provides the tape, device memory and the
layout glue (NCHW<->NHWC permutes, the Leafmap channel concat, the fmask gather) only.  No CPU path: CPU tensors raise.

Reference semantics: torchvision ``Conv2dNormActivation`` / ``FusedMBConv`` / ``MBConv`` / ``SqueezeExcitation`` /
``StochasticDepth("row")`` (efficientnet.py:105-231, ops/misc.py:69-126,225-261, ops/stochastic_depth.py),
``Leafmap.forward`` (models/detector.py:192-201), ``SimpleDecoder`` (:232-254), ``nn.BatchNorm2d`` in train mode
(batch statistics, momentum 0.1, unbiased running variance).

Status: first correct path (CUDA-core gradients in fp32 / bf16 storage; stride-1 data gradients of bf16 convolutions reuse
the tcgen05 forward kernel with transposed, flipped weights).  ``K`` is the kernel namespace; the CPU tests swap in
``oracle.train_oracle`` to check the host-side graph logic against the reference's autograd.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

from . import _lib, _ops, arch

K = _ops
BN_MOMENTUM = 0.1              # nn.BatchNorm2d / BatchNorm1d default; torchvision passes only eps
STOCHASTIC_DEPTH_PROB = 0.2    # torchvision EfficientNet.__init__ default (the reference does not override it)


def _need_cuda(t: Tensor, what: str) -> None:
    """The product path has no CPU route (the CPU graph test swaps this check out together with the kernel namespace)."""
    if not t.is_cuda:
        raise RuntimeError(f"findtextcenternet_b200 {what}: input must be a CUDA tensor (no CPU path)")


def _backend(x: Tensor, cin: int = 64, cout: int = 64) -> int:
    """tcgen05 for bf16 convolutions with GEMM-sized channel counts; the stem (3 -> 32) and the 1-/2-channel top convs are
    bandwidth-bound slivers and stay on the CUDA-core kernel (as in the inference plan)."""
    if x.dtype != torch.bfloat16 or cin < 32 or cout < 16:
        return _lib.GEMM_SIMT
    return _lib.GEMM_TCGEN05


def _pad_last(t: Tensor, mult: int, at_least: int = 0) -> Tensor:
    """Zero-pad the channel (last) axis up to a multiple of ``mult`` (and at least ``at_least``)."""
    c = t.shape[-1]
    target = max((c + mult - 1) // mult * mult, at_least)
    return t if target == c else torch.nn.functional.pad(t, (0, target - c))


class _Conv2d(Function):
    """nn.Conv2d(k in {1,3}, padding=(k-1)//2, stride, bias optional) on NHWC tensors; weight OIHW.

    bf16 tensors keep every GEMM-shaped piece on the tensor cores.  Channel counts the tcgen05 kernels do not take directly are
    zero-padded (zero channels contribute nothing): the 1- / 2-channel top convs run as 16-output-channel GEMMs, the data and
    weight gradients of the 1- / 2- / 100- / 1091- / 1093- / 1097-channel outputs see dy padded to a multiple of 8 (>= 32 as the
    input of the data-gradient convolution).  The data gradient of a stride-2 convolution is the stride-1 data gradient of dy
    with zeros inserted between its pixels (4x the minimal arithmetic, but on tcgen05 instead of CUDA cores)."""

    @staticmethod
    def forward(ctx, x, w, bias, stride):
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.has_bias = stride, bias is not None
        cout, cin = w.shape[0], w.shape[1]
        if x.dtype == torch.bfloat16 and cin >= 32 and cout < 16 and stride == 1:
            wp = torch.nn.functional.pad(w.detach(), (0, 0, 0, 0, 0, 0, 0, 16 - cout))
            bp = None if bias is None else torch.nn.functional.pad(bias.detach(), (0, 16 - cout))
            return K.conv2d(x, wp, 1, None, bp, _lib.ACT_NONE, None, None, _lib.GEMM_TCGEN05)[..., :cout].contiguous()
        return K.conv2d(x, w, stride, None, bias, _lib.ACT_NONE, None, None, _backend(x, w.shape[1], w.shape[0]))

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        cout, cin, k, _ = w.shape
        dx = dw = db = None
        tc = dy.dtype == torch.bfloat16 and x.shape[-1] % 8 == 0
        dyp = _pad_last(dy, 8, 16) if tc else dy         # shared by the data and the weight gradient (tcgen05 wgrad: >= 16 rows)
        if ctx.needs_input_grad[0]:
            if tc and cin >= 16 and (ctx.stride == 1 or k == 3):
                # data gradient = a stride-1 convolution of dy with the taps rotated 180 degrees and the channel roles swapped:
                # runs on the tcgen05 forward kernel
                d = _pad_last(dyp, 8, 32)
                if ctx.stride == 2:
                    up = torch.zeros(d.shape[0], x.shape[1], x.shape[2], d.shape[-1], dtype=d.dtype, device=d.device)
                    up[:, ::2, ::2] = d                   # U[2 oy, 2 ox] = dy[oy, ox]; iy = 2 oy + ky - 1
                    d = up
                if hasattr(K, "conv2d_dgrad_tc"):         # the weight-pack kernel reads the forward weight transposed + rotated
                    dx = K.conv2d_dgrad_tc(d.contiguous(), w)
                else:                                     # kernel namespaces without it (CPU graph tests): explicit operand
                    wt = w.detach().float().flip(2, 3).transpose(0, 1)
                    if d.shape[-1] != cout:
                        wt = torch.nn.functional.pad(wt, (0, 0, 0, 0, 0, d.shape[-1] - cout))
                    dx = K.conv2d(d, wt.contiguous(), 1, None, None, _lib.ACT_NONE, None, None, _lib.GEMM_TCGEN05)
            else:
                dx = K.conv2d_dgrad(dy, w, x.shape[1], x.shape[2], ctx.stride)
        if ctx.needs_input_grad[1]:
            dw = K.conv2d_wgrad(x, dyp, k, ctx.stride)[:cout].to(w.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            mean, _ = K.bn_stats(dy)
            db = mean * float(dy.numel() // cout)
        return dx, dw, db, None


class _BNAct(Function):
    """y = act(BatchNorm_train(x)) (+ residual); mean / var are the batch statistics of x (their dependence on x is part of
    the backward formula, autograd sees them as constants)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, mean, var, eps, act, residual):
        ctx.save_for_backward(x, gamma, beta, mean, var)
        ctx.eps, ctx.act, ctx.has_res = eps, act, residual is not None
        return K.bn_act(x, mean, var, gamma, beta, eps, act, residual)

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, var = ctx.saved_tensors
        dx, dgamma, dbeta = K.bn_act_bwd(x, dy, mean, var, gamma, beta, ctx.eps, ctx.act)     # dy may be a channel slice (cat backward)
        return dx, dgamma.to(gamma.dtype), dbeta.to(beta.dtype), None, None, None, None, (dy if ctx.has_res else None)


class _DwConv3x3(Function):
    @staticmethod
    def forward(ctx, x, w, stride):
        w9c = w.detach().float().reshape(w.shape[0], 9).t().contiguous()
        ctx.save_for_backward(x, w9c)
        ctx.stride, ctx.wshape, ctx.wdtype = stride, w.shape, w.dtype
        return K.dwconv3x3_raw(x, w9c, stride)

    @staticmethod
    def backward(ctx, dy):
        x, w9c = ctx.saved_tensors
        dy = dy.contiguous()
        dx = K.dwconv3x3_dgrad(dy, w9c, x.shape[1], x.shape[2], ctx.stride) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            dw = K.dwconv3x3_wgrad(x, dy, ctx.stride).t().reshape(ctx.wshape).to(ctx.wdtype)
        return dx, dw, None


class _SqueezeExcite(Function):
    """torchvision SqueezeExcitation: x * sigmoid(fc2(silu(fc1(mean_hw(x)))))."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        b, h, w, c = x.shape
        s = w1.shape[0]
        w1m, w2m = w1.detach().float().reshape(s, c).contiguous(), w2.detach().float().reshape(c, s).contiguous()
        mean = K.spatial_sum(x, None, 1.0 / (h * w))
        hid_pre, gate = K.se_fc_train(mean, w1m, b1.detach().float().contiguous(), w2m, b2.detach().float().contiguous())
        ctx.save_for_backward(x, mean, hid_pre, gate, w1m, w2m)
        ctx.shapes = (w1.shape, w2.shape, w1.dtype)
        return K.scale_bc(x, gate)

    @staticmethod
    def backward(ctx, dy):
        x, mean, hid_pre, gate, w1m, w2m = ctx.saved_tensors
        dy = dy.contiguous()
        b, h, w, c = x.shape
        dgate = K.spatial_sum(dy, x, 1.0)
        dmean, dw1, db1, dw2, db2 = K.se_fc_train_bwd(dgate, gate, hid_pre, mean, w1m, w2m)
        dx = K.scale_bc(dy, gate, dmean, 1.0 / (h * w))
        s1, s2, dt = ctx.shapes
        return dx, dw1.reshape(s1).to(dt), db1.to(dt), dw2.reshape(s2).to(dt), db2.to(dt)


class _Upsample2x(Function):
    @staticmethod
    def forward(ctx, x):
        return K.upsample2x(x)

    @staticmethod
    def backward(ctx, dy):
        return K.upsample2x_bwd(dy)          # dy may be a channel slice of a wider map (cat backward): read in place


class _RowScale(Function):
    """StochasticDepth "row" mode: x * noise[b] (noise = bernoulli(1 - p) / (1 - p), ops/stochastic_depth.py:32-39)."""

    @staticmethod
    def forward(ctx, x, noise):
        sc = noise.float().reshape(-1, 1).expand(x.shape[0], x.shape[-1]).contiguous()
        ctx.save_for_backward(sc)
        return K.scale_bc(x, sc)

    @staticmethod
    def backward(ctx, dy):
        (sc,) = ctx.saved_tensors
        return K.scale_bc(dy.contiguous(), sc), None


# ---- layer helpers over the parameter containers of models/_tree.py ------------------------------------------------
def _sub(node, name):
    return getattr(node, str(name))


def select_rows(flat: Tensor, fmask: Tensor, count: int) -> Tensor:
    """``flat[fmask]`` for the boolean pixel mask of ``TextDetectorModel.get_fmask`` (models/detector.py:270-281).  Boolean
    indexing asks the device how many rows are selected (a host synchronisation, illegal while a CUDA graph is being captured);
    under capture the mask's population is taken to be ``count`` (get_fmask sets exactly 1024 * batch pixels) and the rows are
    gathered through ``nonzero_static`` - same rows, same order, static shapes."""
    if fmask.dtype == torch.bool and fmask.is_cuda and torch.cuda.is_current_stream_capturing():
        idx = torch.nonzero_static(fmask, size=count)[:, 0]
        return flat.index_select(0, idx)
    return flat[fmask]


def conv2d(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, stride: int = 1) -> Tensor:
    return _Conv2d.apply(x, w, bias, stride)


def batchnorm_act(x: Tensor, bn, eps: float, act: int, residual: Optional[Tensor] = None) -> Tensor:
    """Train-mode BatchNorm (+ activation, + residual after it); updates running_mean / running_var /
    num_batches_tracked like torch (momentum 0.1, unbiased variance into running_var)."""
    with torch.no_grad():
        rm, rv, nbt = bn.running_mean, bn.running_var, bn.num_batches_tracked
        if getattr(K, "FUSED_RUNNING_STATS", False) and rm.dtype == torch.float32 and rv.dtype == torch.float32 and nbt.is_cuda:
            mean, var = K.bn_stats(x, (rm, rv, nbt), BN_MOMENTUM)       # running statistics updated by the finishing kernel
        else:
            mean, var = K.bn_stats(x)
            n = x.numel() // x.shape[-1]
            rm.mul_(1 - BN_MOMENTUM).add_(mean.to(rm.dtype), alpha=BN_MOMENTUM)
            unbiased = var * (float(n) / max(n - 1, 1))
            rv.mul_(1 - BN_MOMENTUM).add_(unbiased.to(rv.dtype), alpha=BN_MOMENTUM)
            bn.num_batches_tracked += 1
    return _BNAct.apply(x, bn.weight, bn.bias, mean, var, eps, act, residual)


def conv_bn_act(x: Tensor, node, eps: float, act: int, stride: int = 1, residual: Optional[Tensor] = None) -> Tensor:
    """torchvision Conv2dNormActivation: node.0 = conv (no bias), node.1 = BatchNorm2d."""
    w = _sub(node, 0).weight
    if w.shape[1] == 1 and w.shape[0] > 1 and w.shape[-1] == 3 and x.shape[-1] == w.shape[0]:
        y = _DwConv3x3.apply(x, w, stride)
    else:
        y = conv2d(x, w, None, stride)
    return batchnorm_act(y, _sub(node, 1), eps, act, residual)


def _stochastic_depth(branch: Tensor, inp: Tensor, p: float, noise: Optional[Tensor]) -> Tensor:
    """result = StochasticDepth(p, "row")(branch) + input (efficientnet.py:166-169)."""
    if p == 0.0 and noise is None:
        return None   # caller fuses the add into the preceding kernel
    if noise is None:
        survival = 1.0 - p
        noise = torch.empty(branch.shape[0], dtype=torch.float32, device=branch.device).bernoulli_(survival)
        if survival > 0.0:
            noise.div_(survival)
    return _RowScale.apply(branch, noise) + inp


def _block(x: Tensor, blk, st: arch.StageCfg, cin: int, stride: int, sd_prob: float, noise: Optional[Tensor]) -> Tensor:
    eps = arch.BACKBONE_BN_EPS
    res = x if (stride == 1 and cin == st.cout) else None
    fuse_res = res is not None and sd_prob == 0.0 and noise is None
    r = res if fuse_res else None
    b = blk.block
    if st.fused:
        if st.expand != 1:
            y = conv_bn_act(x, _sub(b, 0), eps, _lib.ACT_SILU, stride)
            y = conv_bn_act(y, _sub(b, 1), eps, _lib.ACT_NONE, 1, r)
        else:
            y = conv_bn_act(x, _sub(b, 0), eps, _lib.ACT_SILU, stride, r)     # residual after the SiLU
    else:
        y = conv_bn_act(x, _sub(b, 0), eps, _lib.ACT_SILU, 1)
        y = conv_bn_act(y, _sub(b, 1), eps, _lib.ACT_SILU, stride)
        se = _sub(b, 2)
        y = _SqueezeExcite.apply(y, se.fc1.weight, se.fc1.bias, se.fc2.weight, se.fc2.bias)
        y = conv_bn_act(y, _sub(b, 3), eps, _lib.ACT_NONE, 1, r)
    if res is not None and not fuse_res:
        y = _stochastic_depth(y, res, sd_prob, noise)
    return y


def backbone_train_forward(features, x_nhwc: Tensor, model_size: str = "xl", sd_prob: float = STOCHASTIC_DEPTH_PROB,
                           sd_noise: Optional[dict] = None) -> List[Tensor]:
    """BackboneModel.forward (models/detector.py:139-146) in train mode -> taps [x1, x2, x3, x4] (NHWC).
    sd_noise: optional {block index: noise [B]} to pin the StochasticDepth draws (tests)."""
    _, stages, _ = arch.backbone_cfg(model_size)
    total = sum(st.layers for st in stages)
    cpad = (-x_nhwc.shape[-1]) % 8
    stem = _sub(features, 0)
    w0 = _sub(stem, 0).weight
    if cpad:   # the GEMM kernels take channel counts in multiples of 8: zero channels, zero weights
        x_nhwc = torch.nn.functional.pad(x_nhwc, (0, cpad))
        w0 = torch.nn.functional.pad(w0, (0, 0, 0, 0, 0, cpad))
    y = batchnorm_act(conv2d(x_nhwc.contiguous(), w0, None, 2), _sub(stem, 1), arch.BACKBONE_BN_EPS, _lib.ACT_SILU)
    taps = []
    bid = 0
    for si, st in enumerate(stages, start=1):
        stage = _sub(features, si)
        for li in range(st.layers):
            cin = st.cin if li == 0 else st.cout
            stride = st.stride if li == 0 else 1
            p = sd_prob * float(bid) / total
            noise = None if sd_noise is None else sd_noise.get(bid)
            y = _block(y, _sub(stage, li), st, cin, stride, p, noise)
            bid += 1
        if si in arch.TAP_FEATURE_IDX:
            taps.append(y)
    y = conv_bn_act(y, _sub(features, len(stages) + 1), arch.BACKBONE_BN_EPS, _lib.ACT_SILU)
    taps.append(y)
    return taps


def leafmap_train_forward(leaf, taps: List[Tensor]) -> Tensor:
    """Leafmap.forward (models/detector.py:192-201) in train mode; NHWC in, NHWC [B,H,W,out_dim] out."""
    y = None
    n = len(taps)
    for i in range(n):
        x = taps[n - 1 - i]
        x = batchnorm_act(x, _sub(leaf.in_bn, n - 1 - i), arch.HEAD_BN_EPS, _lib.ACT_NONE)
        if y is not None:
            x = torch.cat([y, x], dim=-1)
        up = _sub(leaf.upsamplers, i)
        y = conv_bn_act(x, up, arch.HEAD_BN_EPS, _lib.ACT_GELU)
        if i < n - 1:
            y = _Upsample2x.apply(y)
    top = _sub(leaf.top_conv, 0)
    return conv2d(y, top.weight, top.bias, 1)


def detection_train_forward(det, x: Tensor, sd_prob: Optional[float] = None, sd_noise: Optional[dict] = None
                            ) -> Tuple[Tensor, Tensor]:
    """CenterNetDetection.forward (models/detector.py:217-230) in train mode: x NCHW in [0,1] -> (heatmap [B,9,H/4,W/4],
    feature [B,100,H/4,W/4]) fp32 NCHW with autograd history.  StochasticDepth probability: ``sd_prob``, else the module's
    ``stochastic_depth_prob`` attribute, else torchvision's 0.2."""
    _need_cuda(x, "detector")
    if sd_prob is None:
        sd_prob = getattr(det, "stochastic_depth_prob", STOCHASTIC_DEPTH_PROB)
    dt = torch.float32 if det.precision == "fp32" else torch.bfloat16
    xh = (x.float() * 2 - 1).permute(0, 2, 3, 1).to(dt).contiguous()
    taps = backbone_train_forward(det.backbone.features, xh, det.model_size, sd_prob, sd_noise)
    outs = [leafmap_train_forward(getattr(det, name), taps).float() for name, _ in arch.HEADS]
    heat = torch.cat(outs[:-1], dim=-1).permute(0, 3, 1, 2).contiguous()
    feat = outs[-1].permute(0, 3, 1, 2).contiguous()
    return heat, feat


def simple_decoder_train_forward(dec, x: Tensor) -> List[Tensor]:
    """SimpleDecoder.forward (models/detector.py:250-254) in train mode: x [N,100] -> 3 x [N, modulo] fp32."""
    _need_cuda(x, "SimpleDecoder")
    from .engine import default_precision
    prec = getattr(dec, "precision", None) or default_precision()
    dt = torch.float32 if prec == "fp32" else torch.bfloat16
    n = x.shape[0]
    cpad = (-x.shape[1]) % 8
    xp = torch.nn.functional.pad(x.to(dt), (0, cpad)).reshape(n, 1, 1, -1).contiguous()
    outs = []
    for i, m in enumerate(arch.MODULO_LIST):
        blk = _sub(dec.blocks, i)
        w0 = torch.nn.functional.pad(_sub(blk, 0).weight, (0, cpad))[:, :, None, None]
        y = batchnorm_act(conv2d(xp, w0), _sub(blk, 1), arch.HEAD_BN_EPS, _lib.ACT_GELU)
        y = batchnorm_act(conv2d(y, _sub(blk, 3).weight[:, :, None, None]), _sub(blk, 4), arch.HEAD_BN_EPS, _lib.ACT_GELU)
        lin = _sub(blk, 6)
        o = conv2d(y, lin.weight[:, :, None, None], lin.bias)
        outs.append(o.reshape(n, m).float())
    return outs


# ====================================================================================================================
# Transformer train step (train3.py:132-137: ``outputs = model(encoder_input, decoder_input)`` in train mode, then
# ``loss_function3`` and ``backward()``).  dropout = 0 (ModelDimensions default, models/transformer.py:257-264): train-mode
# arithmetic equals eval-mode arithmetic, what is added is the tape.
class _LayerNorm(Function):
    """y = nn.LayerNorm(x (+ r1) (+ r2)); the residual sums of EncoderBlock / DecoderBlock are folded into the kernel."""

    @staticmethod
    def forward(ctx, x, gamma, beta, r1, r2, eps):
        y, xs, mean, rstd = K.layernorm_train(x, gamma, beta, eps, r1, r2)
        ctx.save_for_backward(xs, mean, rstd, gamma)
        ctx.res = (r1 is not None, r2 is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, mean, rstd, gamma = ctx.saved_tensors
        dx, dgamma, dbeta = K.layernorm_train_bwd(xs, dy.contiguous(), mean, rstd, gamma)
        return dx, dgamma.to(gamma.dtype), dbeta.to(gamma.dtype), (dx if ctx.res[0] else None), (dx if ctx.res[1] else None), None


class _SwiGLUGate(Function):
    @staticmethod
    def forward(ctx, x1, xg):
        ctx.save_for_backward(x1, xg)
        return K.swiglu(x1, xg)

    @staticmethod
    def backward(ctx, dh):
        x1, xg = ctx.saved_tensors
        return K.swiglu_bwd(x1, xg, dh.contiguous())


class _Embed3(Function):
    @staticmethod
    def forward(ctx, tokens, e0, e1, e2, dtype):
        ctx.save_for_backward(tokens)
        ctx.ms = (e0.shape[0], e1.shape[0], e2.shape[0])
        ctx.dt = e0.dtype
        return K.embed3(tokens, (e0, e1, e2), dtype)

    @staticmethod
    def backward(ctx, dy):
        (tokens,) = ctx.saved_tensors
        d0, d1, d2 = K.embed3_bwd(tokens, dy.contiguous(), ctx.ms)
        return None, d0.to(ctx.dt), d1.to(ctx.dt), d2.to(ctx.dt), None


class _Attention(Function):
    """F.scaled_dot_product_attention(q, k, v, additive key mask) on [B, L, heads*hd] tensors."""

    @staticmethod
    def forward(ctx, q, k, v, mask, heads):
        ctx.save_for_backward(q, k, v, mask)
        ctx.heads = heads
        return K.attention(q, k, v, heads, mask)

    @staticmethod
    def backward(ctx, dout):
        q, k, v, mask = ctx.saved_tensors
        dq, dk, dv = K.attention_bwd(q, k, v, dout.contiguous(), ctx.heads, mask)
        return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype), None, None


def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """nn.Linear on [..., C] through the shared GEMM kernels ([rows,1,1,C] 1x1 convolution); C padded to a multiple of 8."""
    shp = x.shape
    cpad = (-shp[-1]) % 8
    if cpad:
        x = torch.nn.functional.pad(x, (0, cpad))
        weight = torch.nn.functional.pad(weight, (0, cpad))
    y = conv2d(x.reshape(-1, 1, 1, x.shape[-1]).contiguous(), weight[:, :, None, None], bias)
    return y.reshape(*shp[:-1], weight.shape[0])


def _pos(x: Tensor, table: Tensor) -> Tensor:
    """PositionalEncoding.forward (models/transformer.py:45-56): x + encoding[:L] (a learnable parameter)."""
    return x + table[: x.shape[1]].unsqueeze(0).to(x.dtype)


def _mha(m, heads: int, query: Tensor, key: Optional[Tensor], mask: Optional[Tensor]) -> Tensor:
    """MultiheadAttn.forward (models/transformer.py:99-137): positions on q and k only, v = raw key input, no biases."""
    if key is None:
        key, pos_k = query, m.pos_emb_q.encoding
    else:
        pos_k = m.pos_emb_k.encoding
    q = linear(_pos(query, m.pos_emb_q.encoding), m.q_proj.weight)
    k = linear(_pos(key, pos_k), m.k_proj.weight)
    v = linear(key, m.v_proj.weight)
    a = _Attention.apply(q.contiguous(), k.contiguous(), v.contiguous(), mask, heads)
    return linear(a, m.out_proj.weight)


def _ff(ff, x: Tensor) -> Tensor:
    """SwiGLU.forward (models/transformer.py:66-71)."""
    h = _SwiGLUGate.apply(linear(x, ff.w1.weight, ff.w1.bias).contiguous(), linear(x, ff.wg.weight, ff.wg.bias).contiguous())
    return linear(h, ff.w2.weight, ff.w2.bias)


def _ln(x: Tensor, norm, r1=None, r2=None) -> Tensor:
    return _LayerNorm.apply(x.contiguous(), norm.weight, norm.bias, r1, r2, 1e-5)


def key_pad_mask(enc_input: Tensor) -> Tensor:
    """Transformer.forward (models/transformer.py:249-250): additive fp32 mask [B, Le], -inf on all-zero encoder rows."""
    key_pad = torch.all(enc_input == 0, dim=-1)
    return torch.zeros(key_pad.shape, dtype=torch.float32, device=enc_input.device).masked_fill_(key_pad, float("-inf"))


def encoder_forward(enc, enc_input: Tensor, mask: Optional[Tensor], dt: torch.dtype) -> Tensor:
    """Encoder.forward (models/transformer.py:173-180) layer by layer on the train kernels -> [B, Le, d] (dt)."""
    heads = enc.head_num
    x = _pos(linear(enc_input.to(dt), enc.embed.weight), enc.pos_emb.encoding)
    x = _ln(x, enc.norm)
    for i in range(enc.block_num):
        blk = _sub(enc.blocks, i)
        skip = x
        x = _ln(_mha(blk.mha, heads, x, None, mask), blk.norm1, skip)
        x = _ln(_ff(blk.ff, x), blk.norm2, x, skip)
    return x


def decoder_forward(dec, dec_input, y: Tensor, mask: Optional[Tensor], dt: torch.dtype) -> List[Tensor]:
    """Decoder.forward (models/transformer.py:225-238) -> 3 x [B, Ld, m_i] fp32 logits.  dec_input: int64 tokens [B, Ld], or the
    three residue tensors of DecoderSplited.forward (:371-383)."""
    heads = dec.head_num
    emb = [_sub(dec.embed, i).weight for i in range(3)]
    if isinstance(dec_input, (list, tuple)):     # already reduced modulo m_i: one lookup per table (the other two tables zero)
        x = None
        for i, r in enumerate(dec_input):
            tabs = [e if j == i else torch.zeros_like(e) for j, e in enumerate(emb)]
            part = _Embed3.apply(r.to(torch.int64), tabs[0], tabs[1], tabs[2], dt)
            x = part if x is None else x + part
    else:
        x = _Embed3.apply(dec_input.to(torch.int64), emb[0], emb[1], emb[2], dt)
    x = _ln(_pos(x, dec.pos_emb.encoding), dec.norm)
    for i in range(dec.block_num):
        blk = _sub(dec.blocks, i)
        skip = x
        x = _ln(_mha(blk.self_attn, heads, x, None, None), blk.norm1, skip)
        x = _ln(_mha(blk.cross_attn, heads, x, y, mask), blk.norm2, x)
        x = _ln(_ff(blk.ff, x), blk.norm3, x, skip)
    return [linear(x, _sub(dec.out_layers, i).weight, _sub(dec.out_layers, i).bias).float() for i in range(3)]


def transformer_train_forward(model, enc_input: Tensor, dec_input: Tensor) -> List[Tensor]:
    """Transformer.forward (models/transformer.py:248-253) in train mode with a tape -> 3 x [B, Ld, m_i] fp32."""
    _need_cuda(enc_input, "transformer")
    if getattr(model, "dropout", 0.0):
        raise NotImplementedError("findtextcenternet_b200: train-mode dropout > 0 is not built (ModelDimensions default is 0.0)")
    dt = torch.float32 if model.precision == "fp32" else torch.bfloat16
    mask = key_pad_mask(enc_input)
    y = encoder_forward(model.encoder, enc_input, mask, dt)
    return decoder_forward(model.decoder, dec_input, y, mask, dt)
