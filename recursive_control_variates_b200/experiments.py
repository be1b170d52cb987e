"""experiments.py of the reference, restricted to what the hot path needs: scene /
hyper-parameter dictionaries.  The reference's scenes need downloaded assets
(README.md:114-123); here the media are the closed-form synthetic grids of
BASELINE.json's configs in the same scene layout as scene_janga
(experiments.py:20-133): unit cube with a null BSDF, heterogeneous interior,
perspective camera at distance 2, constant emitter."""
from __future__ import annotations

import os

import torch

from . import mitsuba_compat as mi
from . import plugins  # noqa: F401
from .scene import look_at
from .synthetic import make_medium

OUTPUT_DIR = os.path.join(os.path.dirname(os.path.dirname(__file__)), "output")


def _batch_sensor(film, views):
    """The reference's sensor set-up (experiments.py:28-50): `views` perspective cameras on a circle around the
    medium, side by side on one film of width views * film[0]."""
    import math
    sensor = {"type": "batch",
              "film": {"type": "hdrfilm", "width": film[0] * views, "height": film[1], "rfilter": {"type": "box"}},
              "sampler": {"type": "independent", "sample_count": 1}}
    for i in range(views):
        theta = 2 * math.pi / views * i + 0.1
        sensor[f"sensor_{i:02d}"] = {"type": "perspective", "fov": 45,
                                     "to_world": look_at([0.5 + 2 * math.cos(theta), 0.5, 0.5 + 2 * math.sin(theta)],
                                                         [0.5, 0.5, 0.5], [0, 1, 0])}
    return sensor


def _scene_dict(method, target_res, film, kind, scale=20.0, views=1):
    sig, alb = make_medium(target_res, kind, 0.0)
    medium = {"type": "heterogeneous",
              "sigma_t": {"type": "gridvolume", "data": torch.from_numpy(sig)},
              "albedo": {"type": "gridvolume", "data": torch.from_numpy(alb)},
              "scale": scale, "sample_emitters": True, "has_spectral_extinction": False, "majorant_resolution_factor": 0}
    obj = {"type": "cube", "bsdf": {"type": "null"}}
    obj["interior"] = {"type": "twostatemedium", "old": medium, "new": medium, "incoming": medium} if method == "cv_ps" else medium
    if views > 1:
        return {"type": "scene", "object": obj, "emitter": {"type": "constant", "radiance": 1.0},
                "sensor": _batch_sensor(film, views)}
    return {"type": "scene", "object": obj,
            "emitter": {"type": "constant", "radiance": 1.0},
            "sensor": {"type": "perspective", "fov": 45,
                       "to_world": look_at([0.5, 0.5, -2.0], [0.5, 0.5, 0.5], [0, 1, 0]),
                       "film": {"type": "hdrfilm", "width": film[0], "height": film[1], "rfilter": {"type": "box"}},
                       "sampler": {"type": "independent", "sample_count": 1}}}


def scene_synthetic(method, output_dir, grid=64, init_grid=16, film=(128, 128), spp=16, kind="cube", n_steps=100,
                    upsample=(0.16, 0.48), supergrid=False, views=1):
    """Coarse-to-fine recovery of a synthetic density/albedo cube (C1 of BASELINE.json by default):
    the target is the `grid`^3 medium; optimisation starts from constant init_grid^3 grids and
    doubles the resolution at the `upsample` fractions of the run (experiments.py:99-133).  supergrid = the
    reference's `use_majorant_supergrid` (experiments.py:114,228): majorant supergrid + DDA free-flight sampling;
    views > 1 = the reference's `batch` sensor with that many cameras around the medium (film = views x `film`)."""
    scene = mi.load_dict(_scene_dict(method, grid, film, kind, views=views))
    v = init_grid
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return {
        "scene": scene,
        "variables": {
            "object.interior_medium.sigma_t.data": {"init": torch.full((v, v, v, 1), 0.04, device=dev), "clamp": (0.0, 250.0)},
            "object.interior_medium.albedo.data": {"init": torch.full((v, v, v, 3), 0.6, device=dev), "clamp": (0.0, 1.0),
                                                   "lr_factor": 2.0},
        },
        "use_majorant_supergrid": bool(supergrid), "recomp_freq": 50, "upsample": list(upsample), "schedule": [0.75, 0.85, 0.95],
        "benchmark": True, "n_steps": n_steps, "method": method, "lr": 5e-3, "spp": spp, "spp_grad": spp,
        "spp_ref": max(64, spp), "ref_passes": 2, "spp_inf": 64, "output_dir": output_dir,
        "integrator": mi.load_dict({"type": "twostateprbvolpath" if method == "cv_ps" else "prbvolpath", "max_depth": 64, "rr_depth": 64}),
        "adjoint_integrator": mi.load_dict({"type": "volpathsimple", "max_depth": 64, "rr_depth": 64}),
    }


def scene_synthetic_c1(method, output_dir):
    return scene_synthetic(method, output_dir, grid=64, film=(128, 128), spp=16)


def scene_synthetic_c2(method, output_dir):
    return scene_synthetic(method, output_dir, grid=256, init_grid=32, film=(512, 512), spp=64, upsample=(0.04, 0.16, 0.36))


def scene_synthetic_multiview(method, output_dir):
    # the janga set-up (experiments.py:20-133): 8 views of 256x256 on one film, supergrid, coarse-to-fine
    return scene_synthetic(method, output_dir, grid=128, init_grid=16, film=(256, 256), spp=16, kind="plume",
                           upsample=(0.04, 0.16, 0.36), supergrid=True, views=8)


def scene_synthetic_plume(method, output_dir):
    # like the reference's volume scenes (dust devil / janga: experiments.py:114,228) with the majorant supergrid
    return scene_synthetic(method, output_dir, grid=256, init_grid=32, film=(1024, 1024), spp=256, kind="plume",
                           upsample=(0.04, 0.16, 0.36), supergrid=True)


AVAILABLE_SCENES = [name[len("scene_"):] for name in list(globals()) if name.startswith("scene_")]
