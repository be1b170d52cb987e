"""plugins/volpathsimple.py of the reference: the simplified volumetric path tracer
used as `adjoint_integrator` by all volume experiments (experiments.py:131,245,364).
Primal render and the two-pass path-replay adjoint run in librcv_b200
(rcv_render_adjoint_primal / rcv_render_backward)."""
from __future__ import annotations

import torch

from .. import parallel


def get_single_medium(scene):
    """volpathsimple.py:9-20."""
    shapes = scene.shapes()
    assert len(shapes) == 1, f"Not supported: more than 1 shape in the scene (found {len(shapes)})."
    medium = shapes[0].interior_medium()
    assert medium is not None, "Expected a single shape with an interior medium."
    return medium


class VolpathSimpleIntegrator:
    def __init__(self, props=None):
        props = props or {}
        self.max_depth = int(props.get("max_depth", 6))
        self.rr_depth = int(props.get("rr_depth", 5))
        if self.max_depth < 0:
            self.max_depth = 0x7FFFFFFF
        self.hide_emitters = bool(props.get("hide_emitters", False))       # volpathsimple.py:35
        self.use_nee = bool(props.get("use_nee", True))                    # :37
        # Differential ratio tracking (:39-47), on by default as in the reference.  Mitsuba's
        # sample_interaction_drt is not part of the reference's sources: librcv_b200 restates it from
        # the published algorithm, so this estimator is unbiased but its parity unpinned.
        self.use_drt = bool(props.get("use_drt", True))
        self.use_drt_subsampling = bool(props.get("use_drt_subsampling", True))
        self.use_drt_mis = bool(props.get("use_drt_mis", True))

    def _ctx(self, scene):
        get_single_medium(scene)
        return scene.context(self.max_depth, self.rr_depth, self.hide_emitters, self.use_nee)

    def render(self, scene, sensor: int = 0, seed: int = 0, spp: int = 0, develop: bool = True, evaluate: bool = True):
        assert isinstance(sensor, int)
        ctx = self._ctx(scene)
        rank, world = parallel.world()
        b, e = parallel.shard_spp(spp, rank, world)
        acc = ctx.render_adjoint_primal_accum(seed, spp, b, e)
        parallel.allreduce_sum_([acc])
        return ctx.develop(acc)

    def render_backward(self, scene, params, grad_in, sensor=0, seed=0, spp=0):
        """RBIntegrator.render_backward: returns {parameter key: gradient}; mitsuba_compat.render
        hands them to autograd, which accumulates them into the optimiser's variables."""
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
            ctx.set_drt(self.use_drt, self.use_drt_subsampling, self.use_drt_mis)
            ctx.render_backward(seed, spp, grad_in.contiguous(), b, e, film_w=film_w)
            parallel.allreduce_sum_([ctx.grad_packed()])                   # gradient grids before the Adam step
            g_sigma, g_albedo = ctx.grad_read()
        prefix = params.prefix if params is not None else ""
        return {prefix + "sigma_t.data": g_sigma, prefix + "albedo.data": g_albedo}

    def render_forward(self, scene, params, sensor=0, seed=0, spp=0):
        raise NotImplementedError("forward-mode differentiation is not part of the hot path")

    def to_string(self):
        return f"VolpathSimpleIntegrator[max_depth = {self.max_depth}]"
