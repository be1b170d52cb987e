"""GPU parity at BASELINE.json configs[0] EXACTLY (C1: 64^3 cube, 128x128, 16 spp, cv_ps): the CUDA path through the
C-ABI against the oracle on the full configuration — every lane of the film bit for bit, images < 1e-4 and gradients
< 1e-3 relative L2 (north_star), and the recursive CV image over twelve iterations of a moving medium."""
import numpy as np
import pytest

from helpers import gpu_set_states, rel_l2, to_cuda
from recursive_control_variates_b200.scene import SceneConfig
from recursive_control_variates_b200.synthetic import make_medium

pytestmark = pytest.mark.gpu

GRID, RES, SPP = 64, 128, 16
IMG_TOL, GRAD_TOL = 1e-4, 1e-3
CV_KEYS = ("img", "v_n", "w_s", "mean_x", "mean_y", "var_x", "var_y", "cov")


@pytest.fixture(scope="module")
def c1(oracle, gpu_ctx):
    desc = SceneConfig(width=RES, height=RES).to_desc()
    old, new = make_medium(GRID, "cube", 0.0), make_medium(GRID, "cube", 5e-3)
    gpu_ctx.set_scene(desc)
    gpu_ctx.set_tracer(3)
    gpu_ctx.set_drt(False)
    gpu_ctx.set_majorant_factor(0)
    gpu_set_states(gpu_ctx, old, new)
    return dict(desc=desc, old=old, new=new, mo=oracle.OracleMedium(desc, *old), mn=oracle.OracleMedium(desc, *new))


def test_c1_every_lane_bit_exact(oracle, gpu_ctx, c1):
    """Samples [0, 1) of the 16 spp: one lane per pixel under the box filter, 16 384 lanes compared bit for bit; then
    every sample index on its own (16 x 16 384 lanes = the whole C1 wavefront)."""
    d = c1["desc"]
    for k in range(SPP):
        ao, an = oracle.render_twostates_accum(d, c1["mo"], c1["mn"], 17, SPP, k, k + 1)
        go, gn = gpu_ctx.render_twostates_accum(17, SPP, k, k + 1)
        assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an), f"sample {k} differs"


def test_c1_images(oracle, gpu_ctx, c1):
    d = c1["desc"]
    H, F = oracle.render_twostates(d, c1["mo"], c1["mn"], 17, SPP)
    gH, gF = gpu_ctx.render_twostates(17, SPP)
    eh, ef = rel_l2(gH.cpu().numpy(), H), rel_l2(gF.cpu().numpy(), F)
    print(f"C1 images rel-L2: old {eh:.2e} new {ef:.2e}")
    assert eh < IMG_TOL and ef < IMG_TOL


@pytest.mark.parametrize("use_drt", [False, True])
def test_c1_gradients(oracle, gpu_ctx, c1, use_drt):
    """The adjoint of the L1 loss of the C1 render against a target, 16 spp: d sigma_t and d albedo < 1e-3."""
    d = c1["desc"]
    tgt = make_medium(GRID, "cube", 5e-2)
    ref = oracle.develop(oracle.render_twostates_accum(d, oracle.OracleMedium(d, *tgt), oracle.OracleMedium(d, *tgt), 17843, SPP)[1])
    _, F = oracle.render_twostates(d, c1["mo"], c1["mn"], 17, SPP)
    _, g = oracle.loss_grad(0, F, ref)
    from recursive_control_variates_b200.runtime import tea32
    seed_grad = tea32(17, 1)[0]
    ogs, oga = oracle.vps_backward(d, c1["mn"], seed_grad, SPP, g, use_drt=use_drt)
    gpu_ctx.set_drt(use_drt)
    gpu_ctx.grad_zero()
    gpu_ctx.render_backward(seed_grad, SPP, to_cuda(g))
    gs, ga = gpu_ctx.grad_read()
    gpu_ctx.set_drt(False)
    es, ea = rel_l2(gs.cpu().numpy(), ogs), rel_l2(ga.cpu().numpy(), oga)
    print(f"C1 gradients rel-L2 (use_drt={use_drt}): sigma_t {es:.2e} albedo {ea:.2e}")
    assert es < GRAD_TOL and ea < GRAD_TOL


def test_c1_cv_image_over_iterations(oracle, gpu_ctx, c1):
    """The recursive CV image, end to end, over twelve iterations of a moving medium: each side renders its OWN H and F
    and runs its OWN recursion (cv_integrator.py:58-77, welford.py:59-78).

    The first statistics after the warm-up are rounding residues (welford.py:68 evaluates x - mean / (1 - beta1) with
    mean == (1 - beta1) x), and the first weights are ratios of them (cv_integrator.py:65): ANY float-level difference
    between two renderers' H / F — summation order of the film atomics is enough — is amplified to O(1) there and
    decays only with the beta2 = 0.999 memory of the statistics (measured: ~2e-2 relative L2 on the CV image, also
    between two runs of the oracle with 1e-7 noise on H / F).  No pair of implementations can meet 1e-4 through that
    iteration; what CAN be pinned is that it is the only source: with the residue iterations (it < 3) fed identical
    H / F, the two independent pipelines agree to < 1e-4 on every CV image from iteration 4 on (measured ~2e-7)."""
    d = c1["desc"]
    states = [make_medium(GRID, "cube", 2e-3 * i) for i in range(13)]
    runs = {}
    for shared_warmup in (True, False):
        st_g = gpu_ctx.new_cv_state()
        st_o = {k: np.zeros((RES, RES, 3), np.float32) for k in CV_KEYS}
        n, errs = 0, []
        gpu_set_states(gpu_ctx, states[0], states[0])
        for i in range(1, 13):
            it = i - 1
            gpu_ctx.rotate(to_cuda(states[i][0]), to_cuda(states[i][1]))
            mo, mn = oracle.OracleMedium(d, *states[i - 1]), oracle.OracleMedium(d, *states[i])
            H, F = oracle.render_twostates(d, mo, mn, i, SPP)
            gH, gF = gpu_ctx.render_twostates(i, SPP)
            assert rel_l2(gH.cpu().numpy(), H) < IMG_TOL and rel_l2(gF.cpu().numpy(), F) < IMG_TOL
            if shared_warmup and it < 3:
                gH, gF = to_cuda(H), to_cuda(F)
            gpu_ctx.cv_update(gH, gF, st_g, it, 1, n, 0.9, 0.999)
            oracle.cv_update(H, F, st_o, it, 1, n, 0.9, 0.999)
            n += it > 0
            errs.append(rel_l2(st_g["img"].cpu().numpy(), st_o["img"]))
        runs[shared_warmup] = errs
    print("CV image rel-L2 per iteration, identical H/F for it < 3:", " ".join(f"{e:.1e}" for e in runs[True]))
    print("CV image rel-L2 per iteration, independent throughout:  ", " ".join(f"{e:.1e}" for e in runs[False]))
    assert max(runs[True][4:]) < IMG_TOL, runs[True]
    assert max(runs[False][:2]) < IMG_TOL          # warm-up iterations: img = F
    assert max(runs[False]) < 0.1                  # the residue iteration's amplification stays bounded (measured ~2e-2)
    gpu_set_states(gpu_ctx, c1["old"], c1["new"])
