"""Shared builders for the parity tests: the synthetic scene of SURVEY.md §8d
at small sizes, as inputs for both the oracle (numpy) and the CUDA path (torch)."""
import numpy as np

from recursive_control_variates_b200.scene import SceneConfig
from recursive_control_variates_b200.synthetic import make_medium


def small_scene(width=48, height=40, **kw):
    return SceneConfig(width=width, height=height, **kw)


def media(n=24, kind="cube", perturb=5e-3):
    so, ao = make_medium(n, kind, 0.0)
    sn, an = make_medium(n, kind, perturb)
    return (so, ao), (sn, an)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d


def to_cuda(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def gpu_set_states(ctx, old, new):
    """old/new = (sigma_t, albedo) numpy; leaves ctx with theta_{i-1}=old, theta_i=new."""
    ctx.rotate(to_cuda(old[0]), to_cuda(old[1]))
    ctx.rotate(to_cuda(new[0]), to_cuda(new[1]))


def analytic_transmittance(cfg, sigma, sub=8):
    """Closed form for a homogeneous, purely absorbing medium in the unit box seen by the default camera: per pixel the
    box-filtered averages of exp(-sigma l) and l exp(-sigma l), l = chord of the camera ray through the box (an
    independent numpy computation: pinhole at cfg.to_world's origin looking down +z; the box and the expressions are
    symmetric under the x / y flips of the camera convention)."""
    import math
    W, H = cfg.width, cfg.height
    tan_x = math.tan(math.radians(cfg.fov) * 0.5)
    tan_y = tan_x / (W / H)
    o = np.asarray(cfg.to_world, np.float64)[:3, 3]
    u = (np.arange(W * sub) + 0.5) / (W * sub)
    v = (np.arange(H * sub) + 0.5) / (H * sub)
    dx, dy = np.meshgrid((2 * u - 1) * tan_x, (2 * v - 1) * tan_y)
    d = np.stack([dx, dy, np.ones_like(dx)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    lo, hi = np.asarray(cfg.bbox_min, np.float64), np.asarray(cfg.bbox_max, np.float64)
    with np.errstate(divide="ignore"):
        t0, t1 = (lo - o) / d, (hi - o) / d
    tn, tf = np.minimum(t0, t1).max(-1), np.maximum(t0, t1).min(-1)
    l = np.clip(tf - np.maximum(tn, 0.0), 0.0, None)
    T = np.exp(-sigma * l)
    box = lambda a: a.reshape(H, sub, W, sub).mean((1, 3))
    return box(T), box(l * T)
