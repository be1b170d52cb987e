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
