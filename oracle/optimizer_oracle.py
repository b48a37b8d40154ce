"""ORACLE — test infrastructure only (never imported by the product path).

numpy fp32 restatement of one ``AdamWScheduleFree.step`` (reference models/adamw_schedulefree.py:105-214, foreach
branch :157-184) for a list of parameters.  Pinned by tests/golden/optimizer_seed0.npz, which oracle/make_golden.py
produces by running the unmodified reference optimizer."""
from __future__ import annotations

import numpy as np


def lerp(start, end, w):
    w = np.float32(w)
    diff = end - start
    return np.where(w < 0.5, start + w * diff, end - diff * (np.float32(1) - w)).astype(np.float32)


class AdamWScheduleFreeOracle:
    def __init__(self, params, lr=0.0025, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, warmup_steps=0, r=0.0,
                 weight_lr_power=2.0):
        self.y = [np.array(p, dtype=np.float32) for p in params]      # train mode: the parameter tensor holds y
        self.z = [p.copy() for p in self.y]
        self.v = [np.zeros_like(p) for p in self.y]
        self.lr, self.betas, self.eps, self.decay = lr, betas, eps, weight_decay
        self.warmup, self.r, self.wlp = warmup_steps, r, weight_lr_power
        self.k, self.weight_sum, self.lr_max = 0, 0.0, -1.0

    def step(self, grads):
        beta1, beta2 = self.betas
        k = self.k
        sched = (k + 1) / self.warmup if k < self.warmup else 1.0
        bc2 = 1 - beta2 ** (k + 1)
        lr = self.lr * sched
        self.lr_max = max(lr, self.lr_max)
        weight = ((k + 1) ** self.r) * (self.lr_max ** self.wlp)
        self.weight_sum += weight
        ckp1 = weight / self.weight_sum if self.weight_sum != 0 else 0
        f = np.float32
        for i, g in enumerate(grads):
            g = np.asarray(g, dtype=np.float32)
            v = self.v[i] * f(beta2)
            v = v + f(1 - beta2) * (g * g)
            self.v[i] = v.astype(np.float32)
            denom = np.sqrt(self.v[i] / f(bc2)) + f(self.eps)
            gn = (g / denom).astype(np.float32)
            if self.decay != 0:
                gn = gn + f(self.decay) * self.y[i]
            y = lerp(self.y[i], self.z[i], ckp1)
            self.y[i] = (y + f(lr * (beta1 * (1 - ckp1) - 1)) * gn).astype(np.float32)
            self.z[i] = (self.z[i] - f(lr) * gn).astype(np.float32)
        self.k += 1

    def eval_params(self):
        """x = y lerp z with weight 1 - 1/beta1 (optimizer.eval(), :77-89)."""
        beta1 = self.betas[0]
        return [lerp(y, z, 1 - 1 / beta1) for y, z in zip(self.y, self.z)]


class RAdamScheduleFreeOracle(AdamWScheduleFreeOracle):
    """numpy fp32 restatement of ``RAdamScheduleFree.step`` (reference models/radam_schedulefree.py:109-236): the AdamW update
    with the rectified learning rate (:138-152) and no normalisation while rho_t <= 4 (:182-190).  Pinned by
    tests/golden/optimizer_radam_seed0.npz (reference run)."""

    def __init__(self, params, lr=0.0025, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, r=0.0, weight_lr_power=2.0,
                 silent_sgd_phase=True):
        super().__init__(params, lr, betas, eps, weight_decay, 0, r, weight_lr_power)
        self.silent = silent_sgd_phase

    def step(self, grads):
        beta1, beta2 = self.betas
        step = self.k + 1
        beta2_t = beta2 ** step
        bc2 = 1 - beta2_t
        rho_inf = 2 / (1 - beta2) - 1
        rho_t = rho_inf - 2 * step * beta2_t / bc2
        rect = (((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t)) ** 0.5
                if rho_t > 4.0 else float(not self.silent))
        lr = self.lr * rect
        self.lr_max = max(lr, self.lr_max)
        weight = (step ** self.r) * (self.lr_max ** self.wlp)
        self.weight_sum += weight
        ckp1 = weight / self.weight_sum if self.weight_sum != 0 else 0
        f = np.float32
        for i, g in enumerate(grads):
            g = np.asarray(g, dtype=np.float32)
            v = self.v[i] * f(beta2)
            v = v + f(1 - beta2) * (g * g)
            self.v[i] = v.astype(np.float32)
            gn = g
            if rho_t > 4.0:
                gn = (g / (np.sqrt(self.v[i] / f(bc2)) + f(self.eps))).astype(np.float32)
            if self.decay != 0:
                gn = gn + f(self.decay) * self.y[i]
            y = lerp(self.y[i], self.z[i], ckp1)
            self.y[i] = (y + f(lr * (beta1 * (1 - ckp1) - 1)) * gn).astype(np.float32)
            self.z[i] = (self.z[i] - f(lr) * gn).astype(np.float32)
        self.k += 1
