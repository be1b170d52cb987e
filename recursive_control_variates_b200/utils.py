"""utils.py of the reference (loss functions, parameter initialisation, clamp, upsample,
reference render, checkpoint metrics) on top of the B200 kernels.  File formats of the
reference (EXR images, Mitsuba .vol grids) are outside the hot path: metrics and final
parameters are returned / saved as .npz instead."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _abi
from . import mitsuba_compat as mi
from .plugins.volpathsimple import get_single_medium  # noqa: F401  (utils.py:4 re-export)


class _LossOp(torch.autograd.Function):
    """loss value + d loss / d img in one kernel (rcv_loss_grad); utils.py:21-25, optimize.py:110-122."""

    @staticmethod
    def forward(ctx, x, y, kind):
        c = mi._any_context()
        loss, grad = c.loss_grad(kind, x.contiguous(), y.contiguous(), True)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def l2_loss(x, y):
    return _LossOp.apply(x, y, _abi.RCV_LOSS_L2)


def l1_loss(x, y):
    return _LossOp.apply(x, y, _abi.RCV_LOSS_L1)


class Stopwatch:
    """Replacement of utils.runtime() (sum of Dr.Jit kernel times, utils.py:16-19): CUDA-event
    time of everything enqueued on the current stream since start(), in milliseconds."""

    def __init__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)

    def start(self):
        self.e0.record()

    def stop(self) -> float:
        self.e1.record()
        self.e1.synchronize()
        return self.e0.elapsed_time(self.e1)


def adjust_majorant_res_factor(scene, density_res, res_factor=8):
    """utils.py:27-43: the largest factor <= 8 that leaves a supergrid of at least 4 cells per side, else none.
    (Upstream requires the grid resolution to be divisible by the factor; so does rcv_medium_rotate.)  The start value
    is the reference's; these kernels prefer coarser supergrids at high resolutions (256^3: factor 32 is 13 % faster
    than no supergrid, factor 8 is 14 % slower — profiles/README.md), which `params["majorant_res_factor"]` selects."""
    min_side = int(min(density_res[:3]))
    while res_factor > 1 and (min_side // res_factor) < 4:
        res_factor -= 1
    if res_factor <= 1:
        res_factor = 0
    medium = scene.medium()
    if medium.majorant_resolution_factor() != res_factor:
        medium.set_majorant_resolution_factor(res_factor)


def initialize_parameters(params, opt, scene, scene_params):
    """utils.py:62-87.  Runs twice for cv_ps so that old == new == init."""
    for _ in range(2 if params["method"] == "cv_ps" else 1):
        for key, param in params["variables"].items():
            init_state = param["init"]
            if "sigma_t" in key:
                assert init_state.shape[-1] == 1
                if params.get("use_majorant_supergrid", False):
                    adjust_majorant_res_factor(scene, init_state.shape, params.get("majorant_res_factor", 8))   # utils.py:75-76
            opt[key] = init_state
            scene_params[key] = opt[key]
            if "lr_factor" in param:
                opt.set_learning_rate({key: params.get("lr", 1e-2) * param["lr_factor"]})
        scene_params.update()


def render_reference(params, scene, integrator):
    """utils.py:98-111 without the on-disk cache."""
    ref_passes = params.get("ref_passes", 1)
    spp_ref = params.get("spp_ref", params.get("spp", 1))
    img_ref = None
    for j in range(ref_passes):
        img = mi.render(scene, seed=17843 + j, integrator=integrator, spp=spp_ref) / ref_passes
        img_ref = img if img_ref is None else img_ref + img
    return img_ref


def render_display(params, scene, integrator):
    """utils.py:113-127: the final state at the visualisation resolution `final_res` (default: the training film),
    averaged over ref_passes renders of spp_ref samples."""
    film = scene.sensors()[0].film()
    opt_res = film.size()
    film.set_size(params.get("final_res", opt_res))
    try:
        img = None
        with torch.no_grad():
            for j in range(params.get("ref_passes", 1)):
                part = mi.render(scene, seed=17843 + j, integrator=integrator, spp=params.get("spp_ref", params.get("spp", 1)))
                part = part / params.get("ref_passes", 1)
                img = part if img is None else img + part
    finally:
        film.set_size(opt_res)
    return img


def checkpoint(scene, params, i, integrator, loss_func, img, img_ref, result_dict):
    """utils.py:129-165: loss / variance against a higher-spp render every recomp_freq iterations; with `save`, the
    images of that iteration (EXR); at `bias_steps`, a ground-truth render of the current state."""
    recomp_freq = params.get("recomp_freq", 10)
    spp_inf = params.get("spp_inf", 64)
    save = params.get("save", False) and params.get("output_dir")
    output_dir = params.get("output_dir")
    if i % recomp_freq == 0:
        with torch.no_grad():
            img_inf = mi.render(scene, integrator=integrator.integrator, seed=i + 1, spp=spp_inf)
            if save:
                save_img(img_inf, output_dir, "img_inf", i // recomp_freq)
            result_dict["loss"][i // recomp_freq] = float(loss_func(img_ref, img_inf))
            result_dict["var"][i // recomp_freq] = float(torch.mean((img_inf - img.detach()) ** 2))
        if save:
            save_img(integrator.img, output_dir, "img", i // recomp_freq)
            if params.get("method", "baseline") != "baseline" and integrator.stats.n > integrator.warmup:
                save_img(integrator.w_s, output_dir, "weights", i // recomp_freq)
                save_img(integrator.H, output_dir, "img_H", i // recomp_freq)
                save_img(integrator.F, output_dir, "img_F", i // recomp_freq)
    if "bias_steps" in params and i in params["bias_steps"] and output_dir:
        with torch.no_grad():
            ref_passes = params.get("ref_passes", 1)
            spp_ref = params.get("spp_ref", params.get("spp", 1))
            img_gt = None
            for j in range(ref_passes):
                part = mi.render(scene, seed=17843 + j, integrator=integrator, spp=spp_ref) / ref_passes
                img_gt = part if img_gt is None else img_gt + part
            save_img(img_gt, output_dir, "bias_gt", i)
            save_img(img.detach(), output_dir, "bias_img", i)


def precondition(params, opt):
    for _, param in params["variables"].items():
        if "largesteps" in param:
            raise NotImplementedError("large-steps preconditioning is outside the hot path (SURVEY.md §2.1 #14)")


def clamp(params, opt):
    """utils.py:172-175."""
    ctx = mi._any_context()
    for key, param in params["variables"].items():
        if "clamp" in param:
            ctx.clamp(opt[key].data, float(param["clamp"][0]), float(param["clamp"][1]))


def upsample(scene, params, opt, integrator):
    """utils.py:177-192: nearest x2 upsampling of every grid, then integrator.reset()."""
    ctx = mi._any_context()
    for key in params["variables"]:
        old = opt[key]
        if old.dim() != 4:
            raise ValueError(f"Upsampling expects a 4D tensor. Got {old.dim()}.")
        opt[key] = ctx.upsample2x(old.data.contiguous())
        if ".sigma_t." in key and params.get("use_majorant_supergrid", False):   # utils.py:187-188
            adjust_majorant_res_factor(scene, opt[key].shape, params.get("majorant_res_factor", 8))
    integrator.reset()


def save_img(img, output_dir, img_name, it):
    """utils.py:8-14: <output_dir>/<img_name>/<it:04d>.exr (an uncompressed float32 OpenEXR file)."""
    from .exrio import write_exr
    filename = os.path.join(output_dir, img_name, f"{it:04d}.exr")
    os.makedirs(os.path.dirname(filename), exist_ok=True)
    write_exr(filename, img)
    return filename


def save_final_state(params, opt, output_dir):
    """utils.py:45-60: 4-D grids as Mitsuba `.vol` files (mi.VolumeGrid.write), everything else into
    final_state.json; 3-D textures as `.exr` (mi.Bitmap.write upstream)."""
    import json
    from .exrio import write_exr
    from .volio import write_vol
    os.makedirs(output_dir, exist_ok=True)
    final_state = {}
    for key in params["variables"]:
        param = opt[key].detach().cpu()
        name = key.replace(".", "_")
        if param.dim() == 4:
            write_vol(os.path.join(output_dir, f"{name}_final.vol"), param)
        elif param.dim() == 3:
            write_exr(os.path.join(output_dir, f"{name}_final.exr"), param)
        else:
            final_state[key] = param.numpy().tolist()
    if final_state:
        with open(os.path.join(output_dir, "final_state.json"), "w") as f:
            json.dump(final_state, f)
