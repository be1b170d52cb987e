"""Closed-form synthetic media of BASELINE.json's configs.

No scene assets of the reference are shipped (`scenes/` is an external
download, README.md:114-123), so every workload is generated from formulas:
a heterogeneous density cube, a perturbed "new" state one Adam step away, a
smoke-plume mask and a dense variant.  Grids use the reference's gridvolume
layout (Z, Y, X, C), float32.
"""
from __future__ import annotations

import numpy as np


def _coords(n: int):
    c = (np.arange(n, dtype=np.float64) + 0.5) / n
    z, y, x = np.meshgrid(c, c, c, indexing="ij")
    return x, y, z


def make_medium(n: int, kind: str = "cube", perturb: float = 0.0):
    """Return (sigma_t (n,n,n,1), albedo (n,n,n,3)) float32.

    kind: "cube" (C1/C2), "plume" (C3/C5), "dense" (C4).
    perturb: amplitude of the Adam-step-sized perturbation that turns the
    "old" state into the "new" one (5e-3 = janga's lr, experiments.py:122).
    """
    x, y, z = _coords(n)
    r2 = (x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2
    s = 0.5 * (1.0 + np.sin(6 * np.pi * x) * np.sin(4 * np.pi * y) * np.sin(10 * np.pi * z)) * np.exp(-8.0 * r2)
    if kind == "plume":
        s = s * np.exp(-((x - 0.5) ** 2 + (z - 0.5) ** 2) / (0.02 + 0.1 * y))
    elif kind == "dense":
        s = s + 0.25
    elif kind != "cube":
        raise ValueError(f"unknown synthetic medium {kind!r}")
    a = np.stack([0.6 + 0.3 * np.sin(2 * np.pi * x),
                  0.6 + 0.3 * np.sin(2 * np.pi * y),
                  0.6 + 0.3 * np.sin(2 * np.pi * z)], axis=-1)
    if perturb != 0.0:
        d = perturb * np.sin(2 * np.pi * (3 * x + 5 * y + 7 * z))
        s = s + d
        a = a + d[..., None]
    s = np.clip(s, 0.0, 1.0)
    a = np.clip(a, 0.0, 1.0)
    return (np.ascontiguousarray(s[..., None], dtype=np.float32),
            np.ascontiguousarray(a, dtype=np.float32))
