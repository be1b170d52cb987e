"""optimize.py of the reference: the optimisation loop (`run_opt`), unchanged in
structure, on top of the mitsuba_compat shim and the B200 kernels."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import mitsuba_compat as mi
from . import plugins  # noqa: F401  (registers the integrators)
from .utils import (Stopwatch, checkpoint, clamp, initialize_parameters, l1_loss, l2_loss, precondition, render_reference,
                    render_display, save_final_state, upsample)


def run_opt(params):
    scene = params["scene"]
    scene_ref = params.get("scene_ref", scene)
    loss_name = params.get("loss", "L1")
    if loss_name == "L1":
        loss_func = l1_loss
    elif loss_name == "L2":
        loss_func = l2_loss
    else:
        raise ValueError("Unknown loss function")                      # VGG: outside the hot path
    method = params.get("method", "baseline")
    output_dir = params.get("output_dir")
    n_steps = params.get("n_steps", 0)
    lr = params.get("lr", 1e-2)
    spp = params.get("spp", 1)
    spp_grad = params.get("spp_grad", spp)
    benchmark = params.get("benchmark", False)
    assert len(scene.sensors()) == 1
    if output_dir and not os.path.isdir(output_dir):
        os.makedirs(output_dir)

    scene_params = mi.traverse(scene)
    recomp_freq = params.get("recomp_freq", 10)
    result_dict = {"loss": np.zeros(1 + n_steps // recomp_freq, dtype=np.float32),
                   "var": np.zeros(1 + n_steps // recomp_freq, dtype=np.float32),
                   "runtime": np.zeros((n_steps, 2), dtype=np.float32)}
    if "integrator" not in params:
        raise ValueError("No integrator specified")
    base_integrator = params["integrator"]
    integrator = mi.load_dict({                                        # optimize.py:55-66
        "type": "meta", "method": method, "denoise": params.get("denoise", False),
        "integrator": base_integrator, "adjoint_integrator": params.get("adjoint_integrator", base_integrator),
        "beta1": params.get("beta1", 0.9), "beta2": params.get("beta2", 0.999),
        "force_baseline": params.get("force_baseline", False), "pre_update": params.get("pre_update", False)})
    upsample_steps = [int(n_steps * w) for w in params.get("upsample", [])]
    schedule_steps = [int(n_steps * w) for w in params.get("schedule", [])]

    img_ref = params["img_ref"] if "img_ref" in params else render_reference(params, scene_ref, base_integrator)
    opt = mi.Adam(lr=lr)
    initialize_parameters(params, opt, scene, scene_params)
    watch = Stopwatch()
    for i in range(n_steps):                                           # optimize.py:104-143
        active_benchmark = benchmark and i > 16
        if active_benchmark:
            watch.start()
        scene_params.update(opt)
        img = mi.render(scene, scene_params, integrator=integrator, seed=i, spp=spp, spp_grad=spp_grad)
        loss = loss_func(img, img_ref)
        if active_benchmark:
            result_dict["runtime"][i, 0] = watch.stop()
        checkpoint(scene, params, i, integrator, loss_func, img, img_ref, result_dict)
        if active_benchmark:
            watch.start()
        loss.backward()
        precondition(params, opt)
        opt.step()
        clamp(params, opt)
        if active_benchmark:
            result_dict["runtime"][i, 1] = watch.stop()
        if i in upsample_steps:
            upsample(scene, params, opt, integrator)
        if i in schedule_steps:
            lr *= 0.5
            opt.set_learning_rate(lr)
    with torch.no_grad():
        scene_params.update(opt)
    if output_dir:
        from .exrio import write_exr
        img_final = render_display(params, scene, base_integrator)           # optimize.py:148-152
        write_exr(os.path.join(output_dir, "img_final.exr"), img_final)
        save_final_state(params, opt, output_dir)
    result_dict["opt"] = opt
    return result_dict
