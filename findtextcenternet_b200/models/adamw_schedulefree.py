"""Drop-in mirror of the reference ``models/adamw_schedulefree.py::AdamWScheduleFree`` (Meta schedule-free AdamW) whose
``step()`` is ONE fused multi-tensor CUDA launch (csrc/optimizer_ops.cu) instead of ~12 ``torch._foreach_*`` passes.
Same constructor arguments, param-group keys, per-parameter state (``z``, ``exp_avg_sq``) and ``train()`` / ``eval()``
protocol (models/adamw_schedulefree.py:77-103), so checkpoints and training scripts are interchangeable."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple, Union

import numpy as np
import torch
import torch.optim

from .. import _lib


class AdamWScheduleFree(torch.optim.Optimizer):
    def __init__(self, params, lr: Union[float, torch.Tensor] = 0.0025, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0, warmup_steps: int = 0, r: float = 0.0,
                 weight_lr_power: float = 2.0, foreach: Optional[bool] = True):
        defaults = dict(lr=lr, betas=betas, eps=eps, r=r, k=0, warmup_steps=warmup_steps, train_mode=False, weight_sum=0.0,
                        lr_max=-1.0, scheduled_lr=0.0, weight_lr_power=weight_lr_power, weight_decay=weight_decay,
                        foreach=foreach)
        super().__init__(params, defaults)
        self._tables = {}

    @torch.no_grad()
    def eval(self):
        for group in self.param_groups:
            beta1, _ = group["betas"]
            if group["train_mode"]:
                for p in group["params"]:
                    state = self.state[p]
                    if "z" in state:
                        p.lerp_(end=state["z"].to(p.device), weight=1 - 1 / beta1)   # p <- x
                group["train_mode"] = False

    @torch.no_grad()
    def train(self):
        for group in self.param_groups:
            beta1, _ = group["betas"]
            if not group["train_mode"]:
                for p in group["params"]:
                    state = self.state[p]
                    if "z" in state:
                        p.lerp_(end=state["z"].to(p.device), weight=1 - beta1)       # p <- y
                group["train_mode"] = True

    def _table(self, gi: int, active):
        """Device pointer / chunk tables of a parameter group (rebuilt when any buffer moves)."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(), self.state[p]["z"].data_ptr(),
                     p.numel()) for p in active)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1]
        lib = _lib.load()
        dev = active[0].device
        ce = int(lib.ftc_adamw_sf_chunk_elems())
        chunks = []
        for ti, p in enumerate(active):
            for off in range(0, p.numel(), ce):
                chunks.append((ti, 0, off))
        ch = np.zeros(len(chunks), dtype=np.dtype([("t", "<i4"), ("pad", "<i4"), ("off", "<i8")]))
        for i, (ti, _, off) in enumerate(chunks):
            ch[i] = (ti, 0, off)
        tab = dict(
            n=len(chunks),
            chunks=torch.from_numpy(ch.view(np.uint8).copy()).to(dev),
            ys=torch.tensor([k[0] for k in key], dtype=torch.int64, device=dev),
            gs=torch.tensor([k[1] for k in key], dtype=torch.int64, device=dev),
            vs=torch.tensor([k[2] for k in key], dtype=torch.int64, device=dev),
            zs=torch.tensor([k[3] for k in key], dtype=torch.int64, device=dev),
            numels=torch.tensor([k[4] for k in key], dtype=torch.int64, device=dev),
        )
        self._tables[gi] = (key, tab)
        return tab

    # ---- CUDA-graph support ------------------------------------------------------------------------------------------
    def prepare_graph(self) -> None:
        """Call once BEFORE capturing a CUDA graph that contains ``step()`` (after at least one eager step, so that the
        per-parameter state and gradients exist at their final addresses).  Moves the step-dependent schedule state (k, lr_max,
        weight_sum) of every param group to the device; from then on ``step()`` under capture records
        ``ftc_adamw_sf_step_dev`` (schedule computed on the device), and every replay of the graph is one more optimizer step.
        ``sync_from_graph()`` copies the device state back into ``param_groups`` (checkpointing, switching back to eager)."""
        self._graph_state = {}
        for gi, group in enumerate(self.param_groups):
            active = [p for p in group["params"] if p.grad is not None]
            if not active:
                continue
            for p in active:
                if "z" not in self.state[p]:
                    raise RuntimeError("prepare_graph(): run one eager step first (optimizer state not initialised)")
            dev = active[0].device
            beta1, beta2 = group["betas"]
            consts = torch.tensor(self._graph_consts(group), dtype=torch.float64, device=dev)
            state = torch.tensor([float(group["k"]), float(group["lr_max"]), float(group["weight_sum"])], dtype=torch.float64, device=dev)
            hyper = torch.zeros(9, dtype=torch.float32, device=dev)
            self._graph_state[gi] = dict(consts=consts, state=state, hyper=hyper, table=self._table(gi, active), active=active)

    _DEV_STEP = "ftc_adamw_sf_step_dev"

    def _graph_consts(self, group) -> list:
        """the eight schedule constants the device schedule kernel reads (ftc_adamw_sf_step_dev)"""
        beta1, beta2 = group["betas"]
        return [float(group["lr"]), beta1, beta2, group["eps"], group["weight_decay"], float(group["warmup_steps"]), group["r"],
                group["weight_lr_power"]]

    def sync_from_graph(self) -> None:
        for gi, gs in getattr(self, "_graph_state", {}).items():
            k, lr_max, weight_sum = (float(v) for v in gs["state"].cpu())
            group = self.param_groups[gi]
            group["k"], group["lr_max"], group["weight_sum"] = int(round(k)), lr_max, weight_sum

    def _step_captured(self) -> None:
        lib = _lib.load()
        gstate = getattr(self, "_graph_state", None)
        if not gstate:
            raise RuntimeError("AdamWScheduleFree.step() inside a CUDA graph capture needs prepare_graph() before the capture")
        for gi, gs in gstate.items():
            active, tab = gs["active"], gs["table"]
            key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(), self.state[p]["z"].data_ptr(),
                         p.numel()) for p in active)
            if key != self._tables[gi][0]:
                raise RuntimeError("a parameter / gradient buffer moved since prepare_graph(): gradients must live in static "
                                   "storage (shard.FlatGradients) for a captured optimizer step")
            dev = active[0].device
            with torch.cuda.device(dev):
                _lib.check(getattr(lib, self._DEV_STEP)(tab["n"], tab["chunks"].data_ptr(), tab["ys"].data_ptr(), tab["gs"].data_ptr(),
                                                        tab["vs"].data_ptr(), tab["zs"].data_ptr(), tab["numels"].data_ptr(),
                                                        gs["consts"].data_ptr(), gs["state"].data_ptr(), gs["hyper"].data_ptr(),
                                                        torch.cuda.current_stream(dev).cuda_stream), self._DEV_STEP)
            torch.autograd.graph.increment_version(active)
            torch.autograd.graph.increment_version([p.grad for p in active])

    @torch.no_grad()
    def step(self, closure: Optional[Callable[[], float]] = None) -> Optional[float]:
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            if closure is not None:
                raise RuntimeError("closure is not supported inside a CUDA graph capture")
            self._step_captured()
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
            r = group["r"]
            warmup_steps = group["warmup_steps"]
            sched = (k + 1) / warmup_steps if k < warmup_steps else 1.0
            bias_correction2 = 1 - beta2 ** (k + 1)
            lr = float(group["lr"]) * sched
            group["scheduled_lr"] = lr
            lr_max = group["lr_max"] = max(lr, group["lr_max"])
            weight = ((k + 1) ** r) * (lr_max ** group["weight_lr_power"])
            weight_sum = group["weight_sum"] = group["weight_sum"] + weight
            try:
                ckp1 = weight / weight_sum
            except ZeroDivisionError:
                ckp1 = 0
            active = [p for p in group["params"] if p.grad is not None]
            for p in active:
                if not p.is_cuda:
                    raise RuntimeError("findtextcenternet_b200 AdamWScheduleFree: parameters must live on a CUDA device")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("fused AdamWScheduleFree needs contiguous fp32 parameters and gradients")
                if "z" not in self.state[p]:
                    self.state[p]["z"] = torch.clone(p, memory_format=torch.preserve_format)
                    self.state[p]["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if active:
                tab = self._table(gi, active)
                dev = active[0].device
                with torch.cuda.device(dev):
                    _lib.check(lib.ftc_adamw_sf_step(tab["n"], tab["chunks"].data_ptr(), tab["ys"].data_ptr(), tab["gs"].data_ptr(),
                                                     tab["vs"].data_ptr(), tab["zs"].data_ptr(), tab["numels"].data_ptr(),
                                                     beta1, beta2, bias_correction2, eps, decay, lr, ckp1,
                                                     torch.cuda.current_stream(dev).cuda_stream), "ftc_adamw_sf_step")
                # the kernel wrote p / grad through raw pointers: tell autograd (and the engines' "did a weight change?" version keys)
                torch.autograd.graph.increment_version(active)
                torch.autograd.graph.increment_version([p.grad for p in active])
            group["k"] = k + 1
        return loss
