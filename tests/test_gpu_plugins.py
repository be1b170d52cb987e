"""GPU tests of the reference-facing plugin layer: the reference's class / plugin names,
properties, the render -> loss -> backward -> Adam loop of optimize.py, checked against
the oracle where it computes something."""
import numpy as np
import pytest

from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mi():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from recursive_control_variates_b200 import mitsuba_compat, plugins  # noqa: F401
    return mitsuba_compat


def _params(method, **kw):
    from recursive_control_variates_b200 import experiments
    return experiments.scene_synthetic(method, None, grid=16, init_grid=8, film=(32, 24), spp=4, n_steps=6, upsample=(0.5,), **kw)


def test_plugin_names_and_errors(mi):
    from recursive_control_variates_b200 import plugins
    for name in ("volpathsimple", "meta", "cv", "twostateprbvolpath", "prbvolpath", "twostatevolpath", "cv_ps"):
        assert name in mi._INTEGRATORS
    with pytest.raises(ValueError):                       # cv_integrator.py:19-20
        mi.load_dict({"type": "cv", "integrator": {"type": "volpathsimple"}})
    inner = mi.load_dict({"type": "twostateprbvolpath", "max_depth": 8, "rr_depth": 8})
    assert isinstance(inner, plugins.TwoStatePRBVolpathIntegrator) and inner.max_depth == 8
    p = _params("baseline")                               # plain heterogeneous medium: no TwoStateMedium
    with pytest.raises(RuntimeError):                     # twostatevolpath.py:77-78
        inner.render_twostates(p["scene"], 0, 0, 1)


def test_cv_integrator_matches_oracle(mi, oracle):
    """`cv` around `twostateprbvolpath`: H, F and the recursive CV image over a few iterations of
    changing parameters, against oracle two-state renders + oracle cv_update."""
    import torch
    from recursive_control_variates_b200.synthetic import make_medium
    p = _params("cv_ps")
    scene = p["scene"]
    integ = mi.load_dict({"type": "cv", "integrator": {"type": "twostateprbvolpath", "max_depth": 64, "rr_depth": 64},
                          "adjoint_integrator": {"type": "volpathsimple", "max_depth": 64, "rr_depth": 64}})
    sp = mi.traverse(scene)
    ks, ka = "object.interior_medium.sigma_t.data", "object.interior_medium.albedo.data"
    states = [make_medium(12, "cube", a) for a in (0.0, 0.0, 4e-3, 8e-3, 1.2e-2)]
    desc = None
    st_o = None
    n = 0
    prev = None
    for it, (sig, alb) in enumerate(states):
        sp[ks] = torch.from_numpy(sig).cuda(); sp[ka] = torch.from_numpy(alb).cuda()
        sp.update()
        if it == 0:
            prev = (sig, alb)
            continue                                       # first update only fills `new`; second makes old == new
        img = integ.render(scene, 0, seed=it, spp=4)
        ctx = integ.integrator._ctx(scene)
        desc = ctx.desc
        mo, mn = oracle.OracleMedium(desc, *prev), oracle.OracleMedium(desc, sig, alb)
        H, F = oracle.render_twostates(desc, mo, mn, it, 4)
        assert rel_l2(integ.H.cpu().numpy(), H) < 1e-4 and rel_l2(integ.F.cpu().numpy(), F) < 1e-4
        # The recursion itself is compared on identical inputs (the GPU's H and F): the first
        # weights after the warm-up are ratios of rounding residues (welford.py:68: x - mean/(1-beta)
        # with mean == (1-beta) x), so float-level differences between two renderers' H/F are
        # amplified to O(1) weight differences there — in the reference as much as here.
        gH, gF = integ.H.cpu().numpy(), integ.F.cpu().numpy()
        if st_o is None:
            st_o = {k: np.zeros_like(gH) for k in ("img", "v_n", "w_s", "mean_x", "mean_y", "var_x", "var_y", "cov")}
        oracle.cv_update(gH, gF, st_o, it - 1, 1, n, 0.9, 0.999)
        n += (it - 1) > 0
        assert np.array_equal(img.cpu().numpy(), st_o["img"], equal_nan=True)
        assert np.array_equal(integ.w_s.cpu().numpy(), st_o["w_s"], equal_nan=True)
        prev = (sig, alb)
    assert integ.it == 4 and integ.stats.n == 3


@pytest.mark.parametrize("method", ["cv_ps", "baseline"])
def test_run_opt_loop(mi, method):
    """optimize.run_opt end to end (render, L1 loss, backward through volpathsimple, Adam, clamp,
    upsample with CV reset): finite, parameters move, loss does not blow up."""
    import torch
    from recursive_control_variates_b200.optimize import run_opt
    p = _params(method)
    init_s = p["variables"]["object.interior_medium.sigma_t.data"]["init"].clone()
    res = run_opt(p)
    opt = res["opt"]
    s = opt["object.interior_medium.sigma_t.data"]
    a = opt["object.interior_medium.albedo.data"]
    assert s.shape == (16, 16, 16, 1) and a.shape == (16, 16, 16, 3)          # one x2 upsample happened
    assert torch.isfinite(s).all() and torch.isfinite(a).all()
    assert float(s.min()) >= 0.0 and float(a.min()) >= 0.0 and float(a.max()) <= 1.0
    assert float((s - 0.04).abs().max()) > 0, "sigma_t did not move"
    assert np.isfinite(res["loss"]).all()


@pytest.mark.parametrize("use_drt", [True, False])
def test_gradient_through_render_op(mi, oracle, use_drt):
    """mi.render(...).backward() == volpathsimple adjoint of d loss / d img (RenderOp boundary), with the
    reference's default use_drt = True (volpathsimple.py:39-47) and with the plain path-replay estimator."""
    import torch
    from recursive_control_variates_b200.utils import l1_loss
    p = _params("cv_ps")
    scene = p["scene"]
    assert p["adjoint_integrator"].use_drt and p["adjoint_integrator"].use_drt_subsampling and p["adjoint_integrator"].use_drt_mis
    p["adjoint_integrator"].use_drt = use_drt
    integ = mi.load_dict({"type": "meta", "method": "cv_ps", "integrator": p["integrator"],
                          "adjoint_integrator": p["adjoint_integrator"]})
    opt = mi.Adam(lr=5e-3)
    sp = mi.traverse(scene)
    from recursive_control_variates_b200.synthetic import make_medium
    sig, alb = make_medium(10, "cube", 0.0)
    ks, ka = "object.interior_medium.sigma_t.data", "object.interior_medium.albedo.data"
    for _ in range(2):
        opt[ks] = torch.from_numpy(sig); opt[ka] = torch.from_numpy(alb)
        sp.update(opt)
    img = mi.render(scene, sp, integrator=integ, seed=3, spp=4, spp_grad=2)
    ref = torch.zeros_like(img)
    loss = l1_loss(img, ref)
    loss.backward()
    gs, ga = opt[ks].grad.cpu().numpy(), opt[ka].grad.cpu().numpy()
    desc = integ.integrator._ctx(scene).desc
    from recursive_control_variates_b200.runtime import tea32
    _, gimg = oracle.loss_grad(0, img.detach().cpu().numpy(), ref.cpu().numpy())
    m = oracle.OracleMedium(desc, sig, alb)
    ogs, oga = oracle.vps_backward(desc, m, tea32(3, 1)[0], 2, gimg, use_drt=use_drt)
    assert rel_l2(gs, ogs) < 1e-3 and rel_l2(ga, oga) < 1e-3
    if use_drt:        # and it is a different estimator from the path-replay one
        pgs, _ = oracle.vps_backward(desc, m, tea32(3, 1)[0], 2, gimg)
        assert rel_l2(gs, pgs) > 1e-2


def test_run_opt_loop_with_majorant_supergrid(mi):
    """`use_majorant_supergrid` (experiments.py:114, utils.py:27-43,75-76,187-188): the factor follows the grid
    resolution through the coarse-to-fine schedule and the whole loop runs on the DDA kernels."""
    import torch
    from recursive_control_variates_b200.optimize import run_opt
    p = _params("cv_ps", supergrid=True)
    res = run_opt(p)
    scene = p["scene"]
    s = res["opt"]["object.interior_medium.sigma_t.data"]
    assert s.shape == (16, 16, 16, 1)
    assert scene.medium().majorant_resolution_factor() == 4           # 8^3 -> factor 2, 16^3 -> factor 4
    o, n = scene._ctx.majorant_grid()
    assert o.shape == (4, 4, 4) and float(n.max()) > 0
    assert torch.isfinite(s).all() and np.isfinite(res["loss"]).all()


def test_cv_pss_and_force_baseline(mi, oracle):
    """(spp = 1: one sample per pixel, so renders are reproducible bit for bit.)  metaintegrator.py:90-91,117-128: the primary-sample-space control variate re-renders the previous state with
    the next seed; force_baseline returns the noisy render while the CV state still advances."""
    import torch
    from recursive_control_variates_b200.optimize import run_opt
    p = _params("cv_pss")
    scene = p["scene"]
    meta = mi.load_dict({"type": "meta", "method": "cv_pss", "integrator": p["integrator"],
                         "adjoint_integrator": p["adjoint_integrator"]})
    inner = p["integrator"]
    img0 = meta.render(scene, 0, seed=5, spp=1)
    F0 = inner.render(scene, 0, 5, 1)
    assert torch.equal(meta.F, F0) and torch.equal(img0, F0)            # it = 0: w = 0, img = F
    assert torch.equal(meta.H, inner.render(scene, 0, 6, 1))            # H = current state at the NEXT seed (:117-119)
    H_prev = meta.H.clone()
    meta.render(scene, 0, seed=6, spp=1)
    # unchanged parameters and the same seed: the control (previous state, seed 6) equals the new render exactly
    assert torch.equal(meta.F, H_prev)
    # against the oracle's recursion on the same H / F
    st = {k: np.zeros(tuple(F0.shape), np.float32) for k in ("img", "v_n", "w_s", "mean_x", "mean_y", "var_x", "var_y", "cov")}
    z = np.zeros(tuple(F0.shape), np.float32)
    oracle.cv_update(z, F0.cpu().numpy(), st, 0, 1, 0, 0.9, 0.999)
    oracle.cv_update(H_prev.cpu().numpy(), meta.F.cpu().numpy(), st, 1, 1, 0, 0.9, 0.999)
    assert np.array_equal(meta.img.cpu().numpy(), st["img"])
    # force_baseline: same statistics, noisy image returned
    fb = mi.load_dict({"type": "meta", "method": "cv_ps", "force_baseline": True,
                       "integrator": _params("cv_ps")["integrator"], "adjoint_integrator": p["adjoint_integrator"]})
    p2 = _params("cv_ps")
    out = fb.render(p2["scene"], 0, seed=3, spp=1)
    assert torch.equal(out, fb.integrator.render(p2["scene"], 0, 3, 1)) and fb.it == 1
    with pytest.raises(NotImplementedError):
        mi.load_dict({"type": "meta", "method": "cv_ps", "denoise": True, "integrator": p2["integrator"]})
    res = run_opt(_params("cv_pss"))
    assert np.isfinite(res["loss"]).all()


def test_run_opt_loop_multiview(mi):
    """The reference's sensor set-up: a `batch` of views on one film (experiments.py:28-50) through the whole loop."""
    import torch
    from recursive_control_variates_b200.optimize import run_opt
    p = _params("cv_ps", views=3)
    scene = p["scene"]
    assert scene.sensors()[0].film().size() == (96, 24) and scene.sensors()[0].to_world.shape == (3, 4, 4)
    res = run_opt(p)
    s = res["opt"]["object.interior_medium.sigma_t.data"]
    assert torch.isfinite(s).all() and np.isfinite(res["loss"]).all() and float((s - 0.04).abs().max()) > 0


def test_run_opt_writes_the_reference_outputs(mi, tmp_path):
    """optimize.py:145-155: img_final.exr at `final_res`, the optimised grids as Mitsuba .vol files."""
    from recursive_control_variates_b200.exrio import read_exr
    from recursive_control_variates_b200.optimize import run_opt
    from recursive_control_variates_b200.volio import read_vol
    p = _params("cv_ps")
    p["output_dir"] = str(tmp_path / "out")
    p["final_res"] = (48, 36)
    p["save"] = True; p["recomp_freq"] = 3; p["bias_steps"] = [2]
    res = run_opt(p)
    out = tmp_path / "out"
    for name in ("img_inf/0000.exr", "img/0001.exr", "weights/0001.exr", "img_H/0001.exr", "img_F/0001.exr", "bias_gt/0002.exr",
                 "bias_img/0002.exr"):
        assert (out / name).is_file(), name
    img, names = read_exr(str(tmp_path / "out" / "img_final.exr"))
    assert img.shape == (36, 48, 3) and names == ["B", "G", "R"] and np.isfinite(img).all() and img.mean() > 0
    s, _, _ = read_vol(str(tmp_path / "out" / "object_interior_medium_sigma_t_data_final.vol"))
    assert np.array_equal(s, res["opt"]["object.interior_medium.sigma_t.data"].detach().cpu().numpy())
    assert p["scene"].sensors()[0].film().size() == (32, 24)               # the training film is restored
