"""plugins/metaintegrator.py of the reference: the generalised wrapper that
optimize.py instantiates (`'type': 'meta'`).  Methods `baseline`, `cv_ps` (path-space control
variate: one correlated two-state render) and `cv_pss` (primary-sample-space control variate: the
previous state is re-rendered with the next iteration's seed, metaintegrator.py:90-91,117-119) and
the `force_baseline` / `pre_update` switches are supported; `denoise` needs the OptiX denoiser and
AOVs and is outside the hot path."""
from __future__ import annotations

import torch

from .cv_integrator import CVIntegrator


class MetaIntegrator(CVIntegrator):
    def __init__(self, props):
        self.denoise = bool(props.get("denoise", False))
        self.force_baseline = bool(props.get("force_baseline", False))          # metaintegrator.py:15
        if self.denoise:
            raise NotImplementedError("denoise (OptiX denoiser + AOV integrator) is outside the hot path (SURVEY.md §2.1 #6)")
        self.integrator = props.get("integrator")
        self.adjoint_integrator = props.get("adjoint_integrator", self.integrator)
        self.method = props.get("method", "baseline")                 # metaintegrator.py:27
        if self.method not in ("baseline", "cv_ps", "cv_pss"):
            raise NotImplementedError(f"method {self.method!r} is not supported (baseline, cv_ps, cv_pss)")
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
        img_noisy = None
        if self.method == "cv_ps":                                    # metaintegrator.py:68-69
            self.H, self.F = self.integrator.render_twostates(scene, sensor, seed, spp)
        else:                                                         # :87-88
            with torch.no_grad():
                img_noisy = self.integrator.render(scene, sensor, seed, spp, develop, evaluate)
        if self.method == "cv_pss":                                   # :90-91
            self.F = img_noisy
            if self.H is None:                                        # H = TensorXf(0.0) before the first re-render (:51)
                self.H = torch.zeros_like(self.F)
        if "cv" in self.method:                                       # :93-113
            self._cv_step(self.integrator._ctx(scene))
        else:                                                         # :114-115
            self.img = img_noisy
        if self.method == "cv_pss":
            # re-render the current state with the seed of the next iteration (:117-119)
            with torch.no_grad():
                self.H = self.integrator.render(scene, sensor, seed + 1, spp, develop, evaluate)
        if self.force_baseline:
            # the noisy image, "only so we can look at images with the different methods, while still taking the
            # same steps" (:124-128)
            if self.method != "cv_ps":
                return img_noisy
            with torch.no_grad():
                return self.integrator.render(scene, sensor, seed, spp, develop, evaluate)
        return self.img
