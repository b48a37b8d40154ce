"""Drop-in mirror of the reference ``models/radam_schedulefree.py::RAdamScheduleFree`` (train3.py:121) whose ``step()`` is one
fused multi-tensor CUDA launch (csrc/optimizer_ops.cu).  Same constructor arguments, param-group keys, per-parameter state and
``train()`` / ``eval()`` protocol as the reference; the rectification schedule (:138-152) is evaluated on the host in Python
floats exactly as the reference does."""
from __future__ import annotations

from typing import Callable, Optional, Tuple, Union

import torch

from .. import _lib
from .adamw_schedulefree import AdamWScheduleFree


class RAdamScheduleFree(AdamWScheduleFree):
    def __init__(self, params, lr: Union[float, torch.Tensor] = 0.0025, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0, r: float = 0.0, weight_lr_power: float = 2.0,
                 foreach: Optional[bool] = True, silent_sgd_phase: bool = True):
        defaults = dict(lr=lr, betas=betas, eps=eps, r=r, k=0, train_mode=False, weight_sum=0.0, lr_max=-1.0, scheduled_lr=0.0,
                        weight_lr_power=weight_lr_power, weight_decay=weight_decay, foreach=foreach,
                        silent_sgd_phase=silent_sgd_phase)
        torch.optim.Optimizer.__init__(self, params, defaults)
        self._tables = {}

    _DEV_STEP = "ftc_radam_sf_step_dev"

    def _graph_consts(self, group) -> list:
        beta1, beta2 = group["betas"]
        return [float(group["lr"]), beta1, beta2, group["eps"], group["weight_decay"], 1.0 if group["silent_sgd_phase"] else 0.0,
                group["r"], group["weight_lr_power"]]

    @torch.no_grad()
    def step(self, closure: Optional[Callable[[], float]] = None) -> Optional[float]:
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            if closure is not None:
                raise RuntimeError("closure is not supported inside a CUDA graph capture")
            self._step_captured()          # device-side rectification schedule (ftc_radam_sf_step_dev)
            return None
        if not self.param_groups[0]["train_mode"]:
            raise Exception("Optimizer was not in train mode when step is called. Please insert .train() and .eval() calls "
                            "on the optimizer. See documentation for details.")
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            eps = group["eps"]
            beta1, beta2 = group["betas"]
            decay = group["weight_decay"]
            k = group["k"]
            step = k + 1
            r = group["r"]
            beta2_t = beta2 ** step
            bias_correction2 = 1 - beta2_t
            rho_inf = 2 / (1 - beta2) - 1                                   # maximum length of the approximated SMA
            rho_t = rho_inf - 2 * step * beta2_t / bias_correction2
            rect = (((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t)) ** 0.5
                    if rho_t > 4.0 else float(not group["silent_sgd_phase"]))
            lr = float(group["lr"]) * rect
            group["scheduled_lr"] = lr
            lr_max = group["lr_max"] = max(lr, group["lr_max"])
            weight = (step ** r) * (lr_max ** group["weight_lr_power"])
            weight_sum = group["weight_sum"] = group["weight_sum"] + weight
            try:
                ckp1 = weight / weight_sum
            except ZeroDivisionError:
                ckp1 = 0
            active = [p for p in group["params"] if p.grad is not None]
            for p in active:
                if not p.is_cuda:
                    raise RuntimeError("findtextcenternet_b200 RAdamScheduleFree: parameters must live on a CUDA device")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("fused RAdamScheduleFree needs contiguous fp32 parameters and gradients")
                if "z" not in self.state[p]:
                    self.state[p]["z"] = torch.clone(p, memory_format=torch.preserve_format)
                    self.state[p]["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if active:
                tab = self._table(gi, active)
                dev = active[0].device
                with torch.cuda.device(dev):
                    _lib.check(lib.ftc_radam_sf_step(tab["n"], tab["chunks"].data_ptr(), tab["ys"].data_ptr(), tab["gs"].data_ptr(),
                                                     tab["vs"].data_ptr(), tab["zs"].data_ptr(), tab["numels"].data_ptr(),
                                                     beta1, beta2, bias_correction2, eps, decay, lr, ckp1, int(rho_t > 4.0),
                                                     torch.cuda.current_stream(dev).cuda_stream), "ftc_radam_sf_step")
                # the kernel wrote p / grad through raw pointers: tell autograd (and the engines' "did a weight change?" version keys)
                torch.autograd.graph.increment_version(active)
                torch.autograd.graph.increment_version([p.grad for p in active])
            group["k"] = k + 1
        return loss
