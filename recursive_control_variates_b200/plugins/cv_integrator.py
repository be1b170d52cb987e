"""plugins/cv_integrator.py of the reference: the path-space recursive control variate
around a two-state inner integrator.  The weight recursion, the image
combination and the Welford update are one fused kernel (rcv_cv_update)."""
from __future__ import annotations

import torch

from .twostatevolpath import TwoStatePRBVolpathIntegrator
from .welford import STAT_KEYS, StatisticsEstimator


class CVIntegrator:
    def __init__(self, props):
        self.integrator = props.get("integrator")
        if type(self.integrator) not in (TwoStatePRBVolpathIntegrator,):
            raise ValueError("CV integrator expects a 'twostate' integrator as nested integrator !")   # cv_integrator.py:19-20
        self.adjoint_integrator = props.get("adjoint_integrator", self.integrator)
        self.warmup = int(props.get("warmup", 1))
        self.beta1 = float(props.get("beta1", 0.9))
        self.beta2 = float(props.get("beta2", 0.999))
        self.pre_update = False
        self.reset()

    def reset(self):
        self.init = False

    def init_buffers(self, like: torch.Tensor = None):
        """cv_integrator.py:32-45 (buffers are allocated at the first render, when the film size is known)."""
        if self.init:
            return
        self.stats = StatisticsEstimator(self.beta1, self.beta2)
        self.img = self.v_n = self.w_s = self.F = self.H = None
        self.it = 0
        self.init = True

    def _cv_step(self, ctx):
        """cv_integrator.py:60-77 on self.H / self.F."""
        if self.img is None or self.img.shape != self.H.shape:
            self.img, self.v_n, self.w_s = (torch.zeros_like(self.H) for _ in range(3))
            self.stats.allocate(self.H)
        state = {"img": self.img, "v_n": self.v_n, "w_s": self.w_s}
        state.update({k: getattr(self.stats, k) for k in STAT_KEYS})
        ctx.cv_update(self.H, self.F, state, self.it, self.warmup, self.stats.n, self.beta1, self.beta2, self.pre_update)
        if self.it > 0:
            self.stats.n += 1
        self.it += 1

    def render(self, scene, sensor: int = 0, seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True):
        assert isinstance(sensor, int)
        self.init_buffers()
        self.H, self.F = self.integrator.render_twostates(scene, sensor, seed, spp)     # :58
        self._cv_step(self.integrator._ctx(scene))
        return self.img

    def render_forward(self, scene, params, sensor=0, seed=0, spp=0):
        return self.adjoint_integrator.render_forward(scene, params, sensor, seed, spp)

    def render_backward(self, scene, params, grad_in, sensor=0, seed=0, spp=0):
        """Gradients do not flow through w_s, H or the previous image (cv_integrator.py:87-94)."""
        return self.adjoint_integrator.render_backward(scene, params, grad_in, sensor, seed, spp)
