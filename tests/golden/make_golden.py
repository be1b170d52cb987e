"""Generates tests/golden/oracle_small.npz: outputs of the oracle on a tiny fixed
configuration.  The reference cannot be imported here (no Mitsuba 3 / Dr.Jit), so these
vectors come from the restatement itself: they pin it against regressions (and the CUDA
path against the same numbers on the GPU box), not against the reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def config():
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import make_medium
    desc = SceneConfig(width=16, height=12).to_desc()
    return desc, make_medium(8, "cube", 0.0), make_medium(8, "cube", 5e-3)


def compute(O):
    desc, old, new = config()
    mo, mn = O.OracleMedium(desc, *old), O.OracleMedium(desc, *new)
    acc_o, acc_n = O.render_twostates_accum(desc, mo, mn, seed=7, spp=1, threads=1)
    vps = O.vps_render_accum(desc, mn, seed=8, spp=1, threads=1)
    gimg = (np.arange(16 * 12 * 3, dtype=np.float32).reshape(12, 16, 3) % 7 - 3) / (16 * 12 * 3)
    gs, ga = O.vps_backward(desc, mn, 0x5D8714D7, 2, gimg, threads=1)
    st = {k: np.zeros((12, 16, 3), np.float32) for k in ("img", "v_n", "w_s", "mean_x", "mean_y", "var_x", "var_y", "cov")}
    H, F = O.develop(acc_o), O.develop(acc_n)
    n = 0
    for it in range(5):
        O.cv_update(H * (1 + 0.01 * it), F * (1 + 0.012 * it), st, it, 1, n, 0.9, 0.999)
        n += it > 0
    # the same with the majorant supergrid (factor 4: a 2^3 supergrid) and DDA free-flight sampling
    so, sn = O.OracleMedium(desc, *old, majorant_factor=4), O.OracleMedium(desc, *new, majorant_factor=4)
    sg_o, sg_n = O.render_twostates_accum(desc, so, sn, seed=7, spp=1, threads=1)
    sg_vps = O.vps_render_accum(desc, sn, seed=8, spp=1, threads=1)
    sg_gs, sg_ga = O.vps_backward(desc, sn, 0x5D8714D7, 2, gimg, threads=1)
    sg = {"sg_acc_old": sg_o, "sg_acc_new": sg_n, "sg_vps_acc": sg_vps, "sg_grad_sigma_t": sg_gs.astype(np.float32),
          "sg_grad_albedo": sg_ga.astype(np.float32), "sg_majorant_grid_old": so.majorant_grid, "sg_majorant_grid_new": sn.majorant_grid}
    # the adjoint with the DRT in-scattering term: sub-sampling + MIS (the reference default), and neither
    d_gs, d_ga = O.vps_backward(desc, mn, 0x5D8714D7, 2, gimg, threads=1, use_drt=True)
    e_gs, e_ga = O.vps_backward(desc, mn, 0x5D8714D7, 2, gimg, threads=1, use_drt=True, use_drt_subsampling=False, use_drt_mis=False)
    sg.update({"drt_grad_sigma_t": d_gs.astype(np.float32), "drt_grad_albedo": d_ga.astype(np.float32),
               "drt_plain_grad_sigma_t": e_gs.astype(np.float32), "drt_plain_grad_albedo": e_ga.astype(np.float32)})
    return {**sg, "acc_old": acc_o, "acc_new": acc_n, "vps_acc": vps, "grad_sigma_t": gs.astype(np.float32),
            "grad_albedo": ga.astype(np.float32), "cv_img": st["img"], "cv_w_s": st["w_s"],
            "alt_seed": np.array([O.alt_seed(0x5D8714D7)], np.uint32), "majorants": np.array([mo.majorant, mn.majorant], np.float32)}


if __name__ == "__main__":
    import oracle as O
    O.build()
    out = compute(O)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"), **out)
    print({k: v.shape for k, v in out.items()})
