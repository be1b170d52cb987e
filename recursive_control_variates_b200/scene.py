"""Host-side scene description -> struct rcv_scene_desc.

Covers the subset of `mi.load_dict` the reference's volumetric experiments use
(experiments.py:20-133): a perspective sensor (or a batch of views) + hdrfilm (box or
gaussian filter), a box shape with a `null` BSDF containing a heterogeneous grid medium, a `constant`
environment emitter.  All derived quantities (camera frame, tan(fov/2), scene
bounding sphere) are computed here in float64 and rounded once to float32, so
the CUDA kernels and the test oracle receive identical numbers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from ._abi import SceneDesc, RCV_MAX_SENSORS, RCV_PHASE_ISOTROPIC, RCV_PHASE_HG
import ctypes as C

RAY_EPSILON = 1500.0 * 2.0 ** -24  # Mitsuba RayEpsilon (float32)


def look_at(origin: Sequence[float], target: Sequence[float], up: Sequence[float]) -> np.ndarray:
    """Mitsuba's Transform4f.look_at: columns (left, up, forward, origin)."""
    o = np.asarray(origin, dtype=np.float64)
    t = np.asarray(target, dtype=np.float64)
    u = np.asarray(up, dtype=np.float64)
    d = t - o
    d /= np.linalg.norm(d)
    left = np.cross(u, d)
    left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = left, new_up, d, o
    return m


@dataclass
class SceneConfig:
    """Plain-Python description of the supported scene."""
    width: int = 128
    height: int = 128
    fov: float = 45.0                      # degrees, along x (Mitsuba default fov_axis)
    to_world: np.ndarray = field(default_factory=lambda: look_at([0.5, 0.5, -2.0], [0.5, 0.5, 0.5], [0, 1, 0]))
    near_clip: float = 1e-2
    bbox_min: Sequence[float] = (0.0, 0.0, 0.0)
    bbox_max: Sequence[float] = (1.0, 1.0, 1.0)
    sigma_t_scale: float = 20.0            # experiments.py:66
    phase: str = "isotropic"
    phase_g: float = 0.0
    sample_emitters: bool = True
    env_radiance: Sequence[float] = (1.0, 1.0, 1.0)
    max_depth: int = 64                    # experiments.py:130
    rr_depth: int = 64
    hide_emitters: bool = False
    use_nee: bool = True
    rfilter: str = "box"                    # "box" or "gaussian" (Mitsuba's hdrfilm default)
    rfilter_stddev: float = 0.5

    def to_desc(self) -> SceneDesc:
        d = SceneDesc()
        d.struct_size = C.sizeof(SceneDesc)
        d.film_width, d.film_height = int(self.width), int(self.height)
        views = np.asarray(self.to_world, dtype=np.float64)
        batch = views.ndim == 3                       # `batch` sensor: a list of to_world matrices, side by side
        m = views[0] if batch else views
        d.cam_origin[:] = [float(np.float32(v)) for v in m[:3, 3]]
        axes = m[:3, :3].astype(np.float32)          # columns = left, up, forward
        d.cam_axes[:] = [float(v) for v in axes.reshape(-1)]
        n_views = len(views) if batch else 1
        if batch:
            if n_views > RCV_MAX_SENSORS or self.width % n_views:
                raise ValueError("batch sensor: at most 16 views and a film width divisible by their number")
            d.n_sensors = n_views
            for k, mk in enumerate(views):
                d.batch_origin[k][:] = [float(np.float32(v)) for v in mk[:3, 3]]
                d.batch_axes[k][:] = [float(v) for v in mk[:3, :3].astype(np.float32).reshape(-1)]
        tan_x = math.tan(math.radians(self.fov) * 0.5)
        aspect = (self.width / n_views) / self.height
        d.cam_tan_x = float(np.float32(tan_x))
        d.cam_tan_y = float(np.float32(tan_x / aspect))
        d.cam_near_clip = float(np.float32(self.near_clip))
        d.bbox_min[:] = [float(np.float32(v)) for v in self.bbox_min]
        d.bbox_max[:] = [float(np.float32(v)) for v in self.bbox_max]
        d.sigma_t_scale = float(np.float32(self.sigma_t_scale))
        d.phase_type = RCV_PHASE_HG if self.phase == "hg" else RCV_PHASE_ISOTROPIC
        d.phase_g = float(np.float32(self.phase_g))
        d.sample_emitters = int(self.sample_emitters)
        d.env_radiance[:] = [float(np.float32(v)) for v in self.env_radiance]
        # constant.cpp: dist = 2 * bsphere.radius, bsphere of the scene bbox
        # expanded by (1 + RayEpsilon)
        ext = np.asarray(self.bbox_max, dtype=np.float64) - np.asarray(self.bbox_min, dtype=np.float64)
        radius = 0.5 * float(np.linalg.norm(ext))
        radius = max(RAY_EPSILON, radius * (1.0 + RAY_EPSILON))
        d.env_dist = float(np.float32(2.0 * radius))
        d.max_depth, d.rr_depth = int(self.max_depth), int(self.rr_depth)
        d.hide_emitters, d.use_nee = int(self.hide_emitters), int(self.use_nee)
        if self.rfilter not in ("box", "gaussian"):
            raise ValueError("rfilter must be 'box' or 'gaussian'")
        d.rfilter_type = 1 if self.rfilter == "gaussian" else 0
        d.rfilter_stddev = float(np.float32(self.rfilter_stddev))
        return d
