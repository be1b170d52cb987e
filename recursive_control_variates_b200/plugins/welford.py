"""plugins/welford.py of the reference: EMA-Welford statistics of (x = H, y = F).

State lives in float32 CUDA tensors; `update` runs inside the fused CV kernel
(rcv_cv_update) when called through CVIntegrator / MetaIntegrator, so this
class mostly holds the buffers and the counter `n` the optimisation loop reads
(utils.py:152)."""
from __future__ import annotations

import torch

STAT_KEYS = ("mean_x", "mean_y", "var_x", "var_y", "cov")


class StatisticsEstimator:
    """Means, variances and covariance with Adam-style bias correction (welford.py:38-78)."""

    def __init__(self, beta1: float = 0.9, beta2: float = 0.999):
        self.beta1, self.beta2 = float(beta1), float(beta2)
        self.n = 0
        self.mean_x = self.mean_y = self.var_x = self.var_y = self.cov = None

    def allocate(self, like: torch.Tensor):
        for k in STAT_KEYS:
            setattr(self, k, torch.zeros_like(like))

    def update(self, x: torch.Tensor, y: torch.Tensor, ctx):
        """welford.py:59-75 through the CV kernel with the weight / image parts disabled."""
        if self.mean_x is None or self.mean_x.shape != x.shape:
            self.allocate(x)
        scratch = {"img": torch.zeros_like(x), "v_n": torch.zeros_like(x), "w_s": torch.zeros_like(x)}
        scratch.update({k: getattr(self, k) for k in STAT_KEYS})
        ctx.cv_update(x.contiguous(), y.contiguous(), scratch, 1, 1 << 30, self.n, self.beta1, self.beta2, False)
        self.n += 1

    def get(self):
        """welford.py:77-78."""
        c = 1.0 - self.beta2 ** self.n
        return self.var_x / c, self.var_y / c, self.cov / c
