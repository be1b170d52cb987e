"""Closed-form synthetic media of BASELINE.json's configs.

No scene assets of the reference are shipped (`scenes/` is an external
download, README.md:114-123), so every workload is generated from formulas:
a heterogeneous density cube, a perturbed "new" state one Adam step away, a
smoke-plume mask and a dense variant.  Grids use the reference's gridvolume
layout (Z, Y, X, C), float32.
"""
from __future__ import annotations

import numpy as np


def make_medium(n: int, kind: str = "cube", perturb: float = 0.0):
    """Return (sigma_t (n,n,n,1), albedo (n,n,n,3)) float32.

    kind: "cube" (C1/C2), "plume" (C3/C5), "dense" (C4).
    perturb: amplitude of the Adam-step-sized perturbation that turns the
    "old" state into the "new" one (5e-3 = janga's lr, experiments.py:122).
    Evaluated in float64 at the voxel centres (i + 1/2) / n and rounded once; the per-axis factors are computed on 1-D
    arrays and broadcast (the values are those of the full 3-D evaluation, 512^3 takes 13 s instead of 32 s).
    """
    c = (np.arange(n, dtype=np.float64) + 0.5) / n
    x, y, z = c[None, None, :], c[None, :, None], c[:, None, None]        # (Z, Y, X) layout, x fastest
    r2 = (x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2
    s = 0.5 * (1.0 + np.sin(6 * np.pi * x) * np.sin(4 * np.pi * y) * np.sin(10 * np.pi * z)) * np.exp(-8.0 * r2)
    if kind == "plume":
        s = s * np.exp(-((x - 0.5) ** 2 + (z - 0.5) ** 2) / (0.02 + 0.1 * y))
    elif kind == "dense":
        s = s + 0.25
    elif kind != "cube":
        raise ValueError(f"unknown synthetic medium {kind!r}")
    shape = (n, n, n)
    a = np.stack([np.broadcast_to(0.6 + 0.3 * np.sin(2 * np.pi * x), shape),
                  np.broadcast_to(0.6 + 0.3 * np.sin(2 * np.pi * y), shape),
                  np.broadcast_to(0.6 + 0.3 * np.sin(2 * np.pi * z), shape)], axis=-1)
    if perturb != 0.0:
        d = perturb * np.sin(2 * np.pi * (3 * x + 5 * y + 7 * z))
        s = s + d
        a = a + d[..., None]
    s = np.clip(s, 0.0, 1.0)
    a = np.clip(a, 0.0, 1.0)
    return (np.ascontiguousarray(s[..., None], dtype=np.float32),
            np.ascontiguousarray(a, dtype=np.float32))


def boxed_quads(lo=(-1.0, -1.0, -2.5), hi=(2.0, 2.0, 2.0), reflectance=0.5, light_size=1.0, light_radiance=10.0,
                light_center=(0.5, 0.5)):
    """BASELINE.json configs[0..1]: "cube in a box with an area light"  — six diffuse rectangles with
    inward normals enclosing the medium's unit cube and the camera, and a square `area` light just below the ceiling,
    facing down.  Returns the `quads` of scene.SceneConfig."""
    x0, y0, z0 = (float(v) for v in lo)
    x1, y1, z1 = (float(v) for v in hi)
    lx, ly, lz = x1 - x0, y1 - y0, z1 - z0
    r = (reflectance,) * 3 if np.ndim(reflectance) == 0 else tuple(reflectance)
    q = lambda p0, eu, ev, **kw: dict(p0=p0, eu=eu, ev=ev, reflectance=r, **kw)
    s = float(light_size)
    cx, cz = light_center
    return (
        q((x0, y0, z0), (0, 0, lz), (lx, 0, 0)),          # floor, normal +y
        q((x0, y1, z0), (lx, 0, 0), (0, 0, lz)),          # ceiling, normal -y
        q((x0, y0, z0), (0, ly, 0), (0, 0, lz)),          # x = x0, normal +x
        q((x1, y0, z0), (0, 0, lz), (0, ly, 0)),          # x = x1, normal -x
        q((x0, y0, z0), (lx, 0, 0), (0, ly, 0)),          # z = z0, normal +z
        q((x0, y0, z1), (0, ly, 0), (lx, 0, 0)),          # z = z1, normal -z
        dict(p0=(cx - s / 2, y1 - 0.01, cz - s / 2), eu=(s, 0, 0), ev=(0, 0, s), reflectance=(0.0, 0.0, 0.0),
             radiance=(light_radiance,) * 3 if np.ndim(light_radiance) == 0 else tuple(light_radiance)),   # light, normal -y
    )
