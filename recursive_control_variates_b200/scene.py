"""Host-side scene description -> struct rcv_scene_desc.

Covers the subset of `mi.load_dict` the reference's volumetric experiments use
(experiments.py:20-133): a perspective sensor (or a batch of views) + hdrfilm (box or
gaussian filter), a box shape with a `null` BSDF containing a heterogeneous grid medium, one emitter
(`constant` environment, `envmap`, or an `area` light on a rectangle) and up to 8 diffuse rectangles.  All derived quantities (camera frame, tan(fov/2), scene
bounding sphere) are computed here in float64 and rounded once to float32, so
the CUDA kernels and the test oracle receive identical numbers.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from ._abi import (SceneDesc, RCV_MAX_SENSORS, RCV_MAX_QUADS, RCV_PHASE_ISOTROPIC, RCV_PHASE_HG, RCV_EMITTER_CONSTANT,
                   RCV_EMITTER_ENVMAP, RCV_EMITTER_AREA)
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
    # `envmap` emitter (experiments.py:81-85): (H, W, 3) lat-long texels, row 0 = +y; replaces the constant environment
    envmap: Optional[np.ndarray] = None
    env_scale: float = 1.0
    env_to_world: Optional[np.ndarray] = None          # 3x3 rotation, emitter frame -> world
    # rectangles outside the medium box: dicts with p0, eu, ev (3-vectors), reflectance (rgb) and optionally radiance (rgb):
    # the rectangle with a radiance carries the scene's `area` light (then there is no environment emitter)
    quads: Sequence[dict] = ()

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
        lo, hi = np.asarray(self.bbox_min, dtype=np.float64), np.asarray(self.bbox_max, dtype=np.float64)
        for q in self.quads:                          # the scene's bounding box covers the rectangles as well
            p0, eu, ev = (np.asarray(q[k], dtype=np.float64) for k in ("p0", "eu", "ev"))
            for c in (p0, p0 + eu, p0 + ev, p0 + eu + ev):
                lo, hi = np.minimum(lo, c), np.maximum(hi, c)
        ext = hi - lo
        radius = 0.5 * float(np.linalg.norm(ext))
        radius = max(RAY_EPSILON, radius * (1.0 + RAY_EPSILON))
        d.env_dist = float(np.float32(2.0 * radius))
        d.max_depth, d.rr_depth = int(self.max_depth), int(self.rr_depth)
        d.hide_emitters, d.use_nee = int(self.hide_emitters), int(self.use_nee)
        if self.rfilter not in ("box", "gaussian"):
            raise ValueError("rfilter must be 'box' or 'gaussian'")
        d.rfilter_type = 1 if self.rfilter == "gaussian" else 0
        d.rfilter_stddev = float(np.float32(self.rfilter_stddev))
        # ---- emitter and surfaces (ABI 4)
        d.emitter_type = RCV_EMITTER_CONSTANT
        d.env_scale = float(np.float32(self.env_scale))
        R = np.eye(3) if self.env_to_world is None else np.asarray(self.env_to_world, dtype=np.float64)[:3, :3]
        d.env_to_world[:] = [float(v) for v in R.astype(np.float32).reshape(-1)]
        if self.envmap is not None:
            tex = np.ascontiguousarray(self.envmap, dtype=np.float32)
            if tex.ndim != 3 or tex.shape[2] != 3 or tex.shape[0] < 2 or tex.shape[1] < 1:
                raise ValueError("envmap: expected (H >= 2, W, 3) texels")
            d.emitter_type = RCV_EMITTER_ENVMAP
            d.envmap_rgb = tex.ctypes.data
            d.env_height, d.env_width = int(tex.shape[0]), int(tex.shape[1])
            d._envmap_keepalive = tex                 # the desc points into this array
        if len(self.quads) > RCV_MAX_QUADS:
            raise ValueError(f"at most {RCV_MAX_QUADS} rectangles")
        d.n_quads = len(self.quads)
        d.area_quad = -1
        for k, q in enumerate(self.quads):
            for key in ("p0", "eu", "ev"):
                getattr(d.quads[k], key)[:] = [float(np.float32(v)) for v in q[key]]
            d.quads[k].reflectance[:] = [float(np.float32(v)) for v in q.get("reflectance", (0.5, 0.5, 0.5))]
            rad = q.get("radiance")
            d.quads[k].radiance[:] = [float(np.float32(v)) for v in (rad if rad is not None else (0.0, 0.0, 0.0))]
            if rad is not None and any(float(v) != 0.0 for v in rad):
                if d.emitter_type != RCV_EMITTER_CONSTANT or d.area_quad >= 0:
                    raise ValueError("one emitter per scene: an envmap or one area light")
                d.area_quad = k
        if d.area_quad >= 0:
            d.emitter_type = RCV_EMITTER_AREA
            d.env_radiance[:] = [0.0, 0.0, 0.0]
        return d
