"""Generates tests/golden/oracle_r2.npz: outputs of the oracle for the round-2 features on tiny fixed configurations — the
`envmap` emitter, and the cube in a diffuse box with an area light (surface transport, the inherited prbvolpath adjoint).
Like oracle_small.npz these come from the restatement itself (the reference cannot be imported here): a regression pin for
the oracle and the numbers the CUDA path must reproduce on the GPU box.

    python tests/golden/make_golden_r2.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def envmap_texels():
    th = np.linspace(0, np.pi, 9)[:, None]; ph = np.linspace(0, 2 * np.pi, 16, endpoint=False)[None, :]
    base = 0.3 + 0.7 * np.cos(th / 2) ** 2 * (1 + 0.3 * np.sin(3 * ph))
    sun = np.exp(-((th - 0.9) ** 2 + (ph - 2.0) ** 2) / 0.05)
    return (np.stack([base, 0.9 * base, 0.8 * base + 0.1], -1) + 20.0 * sun[..., None] * np.array([1.0, 0.9, 0.7])).astype(np.float32)


def configs():
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import boxed_quads, make_medium
    rot = np.array([[np.cos(0.4), 0, np.sin(0.4)], [0, 1, 0], [-np.sin(0.4), 0, np.cos(0.4)]])
    env = SceneConfig(width=16, height=12, envmap=envmap_texels(), env_scale=0.5, env_to_world=rot, phase="hg", phase_g=0.5)
    box = SceneConfig(width=16, height=12, quads=boxed_quads(), max_depth=12, rr_depth=6)
    return env, box, make_medium(8, "cube", 0.0), make_medium(8, "cube", 5e-3)


def compute(O):
    env, box, old, new = configs()
    gimg = (np.arange(16 * 12 * 3, dtype=np.float32).reshape(12, 16, 3) % 7 - 3) / (16 * 12 * 3)
    out = {}
    d = env.to_desc()
    mo, mn = O.OracleMedium(d, *old), O.OracleMedium(d, *new)
    out["env_acc_old"], out["env_acc_new"] = O.render_twostates_accum(d, mo, mn, seed=7, spp=1, threads=1)
    out["env_vps_acc"] = O.vps_render_accum(d, mn, seed=8, spp=1, threads=1)
    gs, ga = O.vps_backward(d, mn, 0x5D8714D7, 2, gimg, threads=1)
    out["env_grad_sigma_t"], out["env_grad_albedo"] = gs.astype(np.float32), ga.astype(np.float32)
    d = box.to_desc()
    mo, mn = O.OracleMedium(d, *old), O.OracleMedium(d, *new)
    out["box_acc_old"], out["box_acc_new"] = O.render_twostates_accum(d, mo, mn, seed=7, spp=1, threads=1)
    gs, ga = O.prb_backward(d, mn, 0x5D8714D7, 2, gimg, threads=1)
    out["box_grad_sigma_t"], out["box_grad_albedo"] = gs.astype(np.float32), ga.astype(np.float32)
    return out, gimg


if __name__ == "__main__":
    import oracle as O
    O.build()
    out, _ = compute(O)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_r2.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})
