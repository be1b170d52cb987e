"""plugins/twostatevolpath.py of the reference: the correlated two-state volumetric
path tracer.  `render_twostates` keeps the reference's signature and returns
(img_old, img_new); the tracing runs in librcv_b200 (rcv_render_twostates)."""
from __future__ import annotations

import torch

from .. import parallel
from .twostatemedium import TwoStateMedium


class PRBVolpathIntegrator:
    """Stand-in for mi.ad.integrators.prbvolpath.PRBVolpathIntegrator, the base class of
    the reference plugin (twostatevolpath.py:11) and the `prbvolpath` baseline of
    experiments.py:130: a single-state render of the current parameters with the same
    null-scattering estimator (the two-state weights reduce to it when old == new)."""

    def __init__(self, props):
        self.max_depth = int(props.get("max_depth", 6))
        self.rr_depth = int(props.get("rr_depth", 5))
        self.hide_emitters = bool(props.get("hide_emitters", False))
        if self.max_depth < 0:
            self.max_depth = 0x7FFFFFFF          # -1 == infinity
        self.primal_state = 1                    # render theta_i

    def _ctx(self, scene):
        return scene.context(self.max_depth, self.rr_depth, self.hide_emitters, True)

    def render(self, scene, sensor: int = 0, seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True):
        assert isinstance(sensor, int)
        ctx = self._ctx(scene)
        rank, world = parallel.world()
        b, e = parallel.shard_spp(spp, rank, world)
        acc = ctx.render_primal_accum(self.primal_state, seed, spp, b, e)
        parallel.allreduce_sum_([acc])
        return ctx.develop(acc)

    def render_backward(self, scene, params, grad_in, sensor=0, seed=0, spp=0):
        """RBIntegrator.render_backward driving upstream prbvolpath's sample() (Primal, then Backward on the same
        stream) — the adjoint every integrator derived from it inherits, TwoStatePRBVolpathIntegrator included
        (twostatevolpath.py:11), and the one scenes with surfaces or area lights need (volpathsimple.py:24-30 refuses
        them).  Runs in librcv_b200 (rcv_render_backward, RCV_ADJOINT_PRBVOLPATH); returns {parameter key: gradient}."""
        from .. import _abi
        assert isinstance(sensor, int)
        ctx = self._ctx(scene)
        rank, world = parallel.world()
        b, e = parallel.shard_spp(spp, rank, world)
        with torch.no_grad():
            ctx.grad_zero()
            film_w = None
            if world > 1:        # the film weights of all samples: every rank adds its own, one small all-reduce
                film_w = ctx.film_weights(seed, spp, b, e)
                parallel.allreduce_sum_([film_w])
            ctx.render_backward(seed, spp, grad_in.contiguous(), b, e, film_w=film_w, adjoint=_abi.RCV_ADJOINT_PRBVOLPATH)
            parallel.allreduce_sum_([ctx.grad_packed()])
            g_sigma, g_albedo = ctx.grad_read()
        prefix = params.prefix if params is not None else ""
        return {prefix + "sigma_t.data": g_sigma, prefix + "albedo.data": g_albedo}

    def to_string(self):
        return f"PRBVolpathIntegrator[max_depth = {self.max_depth}]"


class TwoStatePRBVolpathIntegrator(PRBVolpathIntegrator):
    def __init__(self, props):
        super().__init__(props)
        self.xyz = props.get("xyz", False)                           # twostatevolpath.py:64
        if self.xyz:
            raise ValueError("xyz output is not supported")
        self.is_checked = False
        self.primal_image_old = self.primal_image_new = None

    def check_scene(self, scene):
        """twostatevolpath.py:67-78."""
        if self.is_checked:
            return
        self.is_checked = True
        has_twostate = any(isinstance(m, TwoStateMedium)
                           for shape in scene.shapes() for m in (shape.interior_medium(), shape.exterior_medium()))
        if not has_twostate:
            raise RuntimeError("No TwoStateMedium found in the scene!")

    def render_twostates(self, scene, sensor=0, seed=0, spp=1):
        """twostatevolpath.py:91-138: one traced sample set, two developed images."""
        self.check_scene(scene)
        assert isinstance(sensor, int)
        ctx = self._ctx(scene)
        rank, world = parallel.world()
        b, e = parallel.shard_spp(spp, rank, world)
        with torch.no_grad():
            acc_old, acc_new = ctx.render_twostates_accum(seed, spp, b, e)
            parallel.allreduce_sum_([acc_old, acc_new])              # film accumulators, before the weight division
            self.primal_image_old = ctx.develop(acc_old)
            self.primal_image_new = ctx.develop(acc_new)
        return self.primal_image_old, self.primal_image_new

    def to_string(self):
        return f"TwoStatePRBVolpathIntegrator[max_depth = {self.max_depth}]"
