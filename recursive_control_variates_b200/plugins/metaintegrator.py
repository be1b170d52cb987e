"""plugins/metaintegrator.py of the reference: the generalised wrapper that
optimize.py instantiates (`'type': 'meta'`).  Methods `baseline` and `cv_ps`
are on the hot path; `cv_pss`, `denoise` (OptiX denoiser) and `force_baseline`
are outside it and rejected."""
from __future__ import annotations

import torch

from .cv_integrator import CVIntegrator


class MetaIntegrator(CVIntegrator):
    def __init__(self, props):
        self.denoise = bool(props.get("denoise", False))
        self.force_baseline = bool(props.get("force_baseline", False))
        if self.denoise or self.force_baseline:
            raise NotImplementedError("denoise / force_baseline are outside the hot path (SURVEY.md §2.1 #6)")
        self.integrator = props.get("integrator")
        self.adjoint_integrator = props.get("adjoint_integrator", self.integrator)
        self.method = props.get("method", "baseline")                 # metaintegrator.py:27
        if self.method not in ("baseline", "cv_ps"):
            raise NotImplementedError(f"method {self.method!r} is not supported (baseline, cv_ps)")
        if self.method == "cv_ps" and not hasattr(self.integrator, "render_twostates"):
            raise ValueError("method cv_ps needs a two-state integrator (twostateprbvolpath)")
        self.pre_update = bool(props.get("pre_update", False))
        self.warmup = int(props.get("warmup", 1))
        self.beta1 = float(props.get("beta1", 0.9))
        self.beta2 = float(props.get("beta2", 0.999))
        self.reset()

    def render(self, scene, sensor: int = 0, seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True):
        assert isinstance(sensor, int)
        self.init_buffers()
        if self.method == "cv_ps":                                    # metaintegrator.py:68-69, 93-113
            self.H, self.F = self.integrator.render_twostates(scene, sensor, seed, spp)
            self._cv_step(self.integrator._ctx(scene))
        else:                                                         # :87-88, :114-115
            with torch.no_grad():
                self.img = self.integrator.render(scene, sensor, seed, spp, develop, evaluate)
        return self.img
