"""Training losses — mirror of the reference ``loss_func.py`` (same names, arguments and result keys) over the fused CUDA
kernels of csrc/loss_ops.cu.  CUDA tensors only (no CPU fallback).

* ``loss_function(fmask, labelmap, idmap, heatmap, decoder_outputs)``  — loss_func.py:94-177 (train1)
* ``loss_function3(outputs, labelcode, mask)``                          — loss_func.py:179-213 (train3)
* ``CoVWeightingLoss``                                                   — loss_func.py:8-72
* ``heatmap_loss_grad(...)``: d(sum_i alpha_i loss_i)/d heatmap for the eight map losses (the analytic backward of the map part).
"""
from __future__ import annotations

import torch

from . import _lib

modulo_list = [1091, 1093, 1097]          # util_func.py:5
MAP_LOSSES = ["keymap_loss", "size_loss", "textline_loss", "separator_loss", "code1_loss", "code2_loss", "code4_loss", "code8_loss"]


def _s(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("findtextcenternet_b200.loss_func runs on CUDA tensors only (no CPU fallback)")


def heatmap_losses(labelmap: torch.Tensor, idmap: torch.Tensor, heatmap: torch.Tensor) -> torch.Tensor:
    """fp32[9] on device: the eight map losses in MAP_LOSSES order + weight1_count."""
    lib = _lib.load()
    _need_cuda(labelmap, idmap, heatmap)
    b, c, h, w = heatmap.shape
    assert c == 9 and labelmap.shape == (b, 5, h, w) and idmap.shape == (b, 2, h, w)
    hm = heatmap.detach().float().contiguous(); lm = labelmap.float().contiguous(); im = idmap.to(torch.int64).contiguous()
    out = torch.empty(9, dtype=torch.float32, device=hm.device)
    scratch = torch.empty(int(lib.ftc_heatmap_loss_scratch_bytes()), dtype=torch.uint8, device=hm.device)
    with torch.cuda.device(hm.device):
        _lib.check(lib.ftc_heatmap_loss(hm.data_ptr(), lm.data_ptr(), im.data_ptr(), b, h, w, out.data_ptr(), scratch.data_ptr(), _s(hm)),
                   "ftc_heatmap_loss")
    return out


def heatmap_loss_grad(labelmap, idmap, heatmap, alphas: torch.Tensor, losses9: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(labelmap, idmap, heatmap)
    b, c, h, w = heatmap.shape
    hm = heatmap.detach().float().contiguous(); lm = labelmap.float().contiguous(); im = idmap.to(torch.int64).contiguous()
    al = alphas.to(device=hm.device, dtype=torch.float32).contiguous()
    grad = torch.empty_like(hm)
    with torch.cuda.device(hm.device):
        _lib.check(lib.ftc_heatmap_loss_grad(hm.data_ptr(), lm.data_ptr(), im.data_ptr(), b, h, w, al.data_ptr(), losses9.data_ptr(),
                                             grad.data_ptr(), _s(hm)), "ftc_heatmap_loss_grad")
    return grad


class _HeatmapLosses(torch.autograd.Function):
    """fp32[9] map losses with the analytic backward of csrc/loss_ops.cu (ftc_heatmap_loss_grad): the upstream gradient of the
    eight losses is exactly the kernel's alpha vector."""

    @staticmethod
    def forward(ctx, heatmap, labelmap, idmap):
        out = heatmap_losses(labelmap, idmap, heatmap)
        ctx.save_for_backward(heatmap, labelmap, idmap, out)
        return out

    @staticmethod
    def backward(ctx, g):
        heatmap, labelmap, idmap, out = ctx.saved_tensors
        grad = heatmap_loss_grad(labelmap, idmap, heatmap, g[:8].float().contiguous(), out)
        return grad.to(heatmap.dtype), None, None


class _CEMean(torch.autograd.Function):
    """out4[0] / max(out4[1], 1) of ftc_ce_rows with ftc_ce_rows_grad as its backward; also returns the detached out4."""

    @staticmethod
    def forward(ctx, l0, l1, l2, target, weight, select, count_select, clamp_den):
        o = _ce_rows((l0, l1, l2), target, weight, select, count_select)
        den = torch.clamp_min(o[1], 1.0) if clamp_den else o[1]
        ctx.save_for_backward(l0, l1, l2, target, weight, select, den)
        ctx.mark_non_differentiable(o)
        return (o[0] / den).float(), o

    @staticmethod
    def backward(ctx, g, _):
        lib = _lib.load()
        l0, l1, l2, target, weight, select, den = ctx.saved_tensors
        ls = [l.detach().float().contiguous() for l in (l0, l1, l2)]
        rows = ls[0].shape[0]
        grads = [torch.empty_like(l) for l in ls]
        coef = (g.double() / den).float().reshape(1).contiguous()
        tg = target.to(torch.int64).contiguous()
        wt = None if weight is None else weight.float().contiguous()
        se = None if select is None else select.to(torch.uint8).contiguous()
        p = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(coef.device):
            _lib.check(lib.ftc_ce_rows_grad(ls[0].data_ptr(), ls[1].data_ptr(), ls[2].data_ptr(), ls[0].stride(0), ls[1].stride(0),
                                            ls[2].stride(0), modulo_list[0], modulo_list[1], modulo_list[2], tg.data_ptr(), p(wt),
                                            p(se), rows, coef.data_ptr(), grads[0].data_ptr(), grads[1].data_ptr(),
                                            grads[2].data_ptr(), _s(coef)), "ftc_ce_rows_grad")
        return grads[0].to(l0.dtype), grads[1].to(l1.dtype), grads[2].to(l2.dtype), None, None, None, None, None


def _ce_rows(logits, target, weight, select, count_select) -> torch.Tensor:
    lib = _lib.load()
    ls = [l.detach().float() for l in logits]
    ls = [l if l.stride(-1) == 1 else l.contiguous() for l in ls]
    rows = ls[0].shape[0]
    out = torch.zeros(4, dtype=torch.float64, device=ls[0].device)
    tg = target.to(torch.int64).contiguous()
    wt = None if weight is None else weight.float().contiguous()
    se = None if select is None else select.to(torch.uint8).contiguous()
    cs = None if count_select is None else count_select.to(torch.uint8).contiguous()
    p = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(out.device):
        _lib.check(lib.ftc_ce_rows(ls[0].data_ptr(), ls[1].data_ptr(), ls[2].data_ptr(), ls[0].stride(0), ls[1].stride(0), ls[2].stride(0),
                                   modulo_list[0], modulo_list[1], modulo_list[2], tg.data_ptr(), p(wt), p(se), p(cs), rows, out.data_ptr(),
                                   _s(out)), "ftc_ce_rows")
    return out


def loss_function(fmask, labelmap, idmap, heatmap, decoder_outputs):
    """loss_func.py:94-177.  Returns the reference's dict (0-d CUDA tensors)."""
    _need_cuda(fmask, labelmap, idmap, heatmap, *decoder_outputs)
    key_th3 = 0.99
    m9 = _HeatmapLosses.apply(heatmap, labelmap, idmap)      # differentiable w.r.t. heatmap (train1.py:151 loss.backward())
    from .train_ops import select_rows
    count = min(1024 * labelmap.shape[0], labelmap[:, 0].numel())      # population of get_fmask's mask (used under graph capture)
    keyvals = select_rows(labelmap[:, 0].flatten(), fmask, count).float()
    target_id = select_rows(idmap[:, 0].flatten(), fmask, count)
    pos = target_id > 0
    weight3 = torch.clamp_min(keyvals - key_th3, 0.) / (1 - key_th3)
    id_loss, o = _CEMean.apply(decoder_outputs[0], decoder_outputs[1], decoder_outputs[2], target_id, weight3,
                               (keyvals > key_th3) & pos, (keyvals == 1) & pos, True)
    res = {k: m9[i] for i, k in enumerate(MAP_LOSSES)}
    res["id_loss"] = id_loss
    res["loss"] = m9[:8].sum() + id_loss
    res["correct"] = o[2].to(torch.int64)
    res["total"] = o[3].to(torch.int64)
    return res


def loss_function3(outputs, labelcode, mask):
    """loss_func.py:179-213: outputs = 3 x [B, L, m_i] logits, labelcode int64 [B, L], mask bool [B, L]."""
    _need_cuda(labelcode, mask, *outputs)
    flat = [o.reshape(-1, o.shape[-1]) for o in outputs]
    m = mask.reshape(-1)
    loss, o = _CEMean.apply(flat[0], flat[1], flat[2], labelcode.reshape(-1), None, m, m, False)
    return {"loss": loss, "correct": o[2].to(torch.int64), "total": o[3].to(torch.int64)}


class CoVWeightingLoss(torch.nn.Module):
    """loss_func.py:8-72 (Multi-Loss Weighting with Coefficient of Variations): Welford statistics of the loss ratios.

    Same arithmetic and attribute names as the reference, but the iteration-dependent pieces (first-iteration L0, the uniform
    weights of iterations 0 and 1, ``mean_param``) are selected by a DEVICE-resident iteration counter and every statistic is
    updated in place, so one call is a fixed sequence of kernels on fixed storage: a train step containing it can be captured
    into a CUDA graph and replayed (findtextcenternet_b200/train.py::Train1Graph) and still advance the statistics."""

    def __init__(self, *args, **kwargs) -> None:
        self.device = kwargs.pop("device", "cpu")
        self.losses = kwargs.pop("losses", [])
        self.num_losses = len(self.losses)
        super().__init__(*args, **kwargs)
        self.current_iter = -1           # host mirror of the device counter (the reference's attribute)
        z = lambda: torch.zeros((self.num_losses,), dtype=torch.float32, device=self.device)
        self.alphas, self.running_mean_L, self.running_mean_l, self.running_S_l = z(), z(), z(), z()
        self.running_std_l = z()
        self._it = torch.full((), -1.0, dtype=torch.float64, device=self.device)

    def forward(self, losses):
        L = torch.stack([losses[key].detach().to(torch.float32) for key in self.losses])
        if not self.train:          # (sic) the reference tests the bound method, which is always truthy (loss_func.py:30)
            return torch.sum(L)
        self.current_iter += 1
        it = self._it.add_(1.0)                                   # device copy of current_iter
        first = it == 0
        L0 = torch.where(first, L, self.running_mean_L)
        l = L / L0
        ls = self.running_std_l / self.running_mean_l
        uniform = torch.full_like(L, 1.0 / self.num_losses) if self.num_losses else L
        self.alphas.copy_(torch.where(it <= 1, uniform, ls / torch.sum(ls)))
        # Python-double arithmetic of the reference (1 - 1 / (n + 1), then 1 - that), rounded to fp32 once as torch does for
        # scalar operands
        mp64 = torch.where(first, torch.zeros_like(it), 1.0 - 1.0 / (it + 1.0))
        mean_param, one_m = mp64.float(), (1.0 - mp64).float()
        new_mean_l = mean_param * self.running_mean_l + one_m * l
        self.running_S_l.add_((l - self.running_mean_l) * (l - new_mean_l))
        self.running_mean_l.copy_(new_mean_l)
        running_variance_l = self.running_S_l / (it + 1.0).float()
        self.running_std_l.copy_(torch.sqrt(running_variance_l.clamp_min(1e-16)))
        self.running_mean_L.copy_(mean_param * self.running_mean_L + one_m * L)
        alphas = self.alphas.clone()         # the weights of THIS call (the statistics above already moved on)
        return sum(alphas[i] * losses[key].to(torch.float32) for i, key in enumerate(self.losses))
