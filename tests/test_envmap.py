"""`envmap` emitter (experiments.py:81-85): the oracle's restatement against closed forms and the
constant emitter, the host-emulated lane machines against the oracle (CPU), the CUDA path against the oracle (GPU).
Upstream envmap.cpp is not readable here: the emitter is restated and, like the rest, unpinned."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import media, rel_l2, small_scene


def sky(h=24, w=48, seed=0):
    """A smooth sky with a bright sun: strongly non-uniform importance."""
    rng = np.random.default_rng(seed)
    th = np.linspace(0, np.pi, h)[:, None]; ph = np.linspace(0, 2 * np.pi, w, endpoint=False)[None, :]
    base = 0.3 + 0.7 * np.cos(th / 2) ** 2 * (1 + 0.3 * np.sin(3 * ph))
    tex = np.stack([base, 0.9 * base, 0.8 * base + 0.1], -1) + 0.05 * rng.random((h, w, 3))
    sun = np.exp(-((th - 0.9) ** 2 + (ph - 2.0) ** 2) / 0.02)
    return (tex + 40.0 * sun[..., None] * np.array([1.0, 0.9, 0.7])).astype(np.float32)


def rot_y(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def test_atan2_matches_libm(oracle):
    rng = np.random.default_rng(1)
    for y, x in np.concatenate([rng.standard_normal((2000, 2)), [[0, 1], [0, -1], [1, 0], [-1, 0], [1e-8, -1], [-1e-8, -1]]]):
        assert abs(oracle.atan2f(float(y), float(x)) - math.atan2(np.float32(y), np.float32(x))) < 4e-7


def test_density_normalised_and_proportional(oracle):
    """pdf integrates to one over the sphere and the importance-sampled estimate of the integral of the radiance equals
    its quadrature; a sampled direction evaluates to the pdf / value the sample reports."""
    cfg = small_scene(8, 8, envmap=sky(), env_scale=0.5, env_to_world=rot_y(0.7))
    d = cfg.to_desc()
    n_t, n_p = 360, 720
    th = (np.arange(n_t) + 0.5) * np.pi / n_t; ph = (np.arange(n_p) + 0.5) * 2 * np.pi / n_p
    int_pdf, int_L = 0.0, np.zeros(3)
    for t in th[::3]:
        for p in ph[::3]:
            v = np.array([math.sin(t) * math.cos(p), math.cos(t), math.sin(t) * math.sin(p)], np.float32)
            rgb, pdf = oracle.env_eval(d, v)
            dw = math.sin(t) * (3 * np.pi / n_t) * (3 * 2 * np.pi / n_p)
            int_pdf += pdf * dw; int_L += rgb * dw
    assert abs(int_pdf - 1.0) < 5e-3
    rng = np.random.default_rng(2)
    est = np.zeros(3)
    n = 4000
    for u1, u2 in rng.random((n, 2)):
        v, pdf, val = oracle.env_sample(d, u1, u2)
        assert abs(np.linalg.norm(v) - 1) < 1e-5 and pdf > 0
        rgb, pdf2 = oracle.env_eval(d, v)
        assert pdf2 == pdf and np.array_equal((rgb / np.float32(pdf)).astype(np.float32), val)
        est += val
    est /= n
    assert np.all(np.abs(est - int_L) < 0.05 * int_L), (est, int_L)      # value / pdf is nearly constant: low variance


def test_constant_map_equals_constant_emitter(oracle):
    """White-furnace style check: an envmap with the same radiance everywhere renders what the `constant` emitter
    renders (same paths up to the emitter-sampling directions; compared in expectation), also when rotated."""
    old, new = media(12)
    tex = np.full((9, 16, 3), 1.0, np.float32) * np.array([0.8, 1.0, 1.2], np.float32)
    ce = small_scene(20, 16, env_radiance=(0.4, 0.5, 0.6)).to_desc()
    ee = small_scene(20, 16, envmap=tex, env_scale=0.5, env_to_world=rot_y(1.1)).to_desc()
    mc, me = oracle.OracleMedium(ce, *new), oracle.OracleMedium(ee, *new)
    a = oracle.develop(oracle.vps_render_accum(ce, mc, 3, 512))
    b = oracle.develop(oracle.vps_render_accum(ee, me, 3, 512))
    assert rel_l2(b.mean((0, 1)), a.mean((0, 1))) < 1e-2
    H1, F1 = oracle.render_twostates(ce, oracle.OracleMedium(ce, *old), mc, 5, 512)
    H2, F2 = oracle.render_twostates(ee, oracle.OracleMedium(ee, *old), me, 5, 512)
    assert rel_l2(F2.mean((0, 1)), F1.mean((0, 1))) < 1e-2 and rel_l2(H2.mean((0, 1)), H1.mean((0, 1))) < 1e-2


def test_two_estimators_agree_under_the_envmap(oracle):
    """volpathsimple and the two-state estimator (old == new) render the same expectation under a sun-and-sky map: emitter
    sampling, emitter hits and their MIS weights are consistent in both."""
    _, new = media(12)
    d = small_scene(16, 12, envmap=sky(), env_scale=0.5).to_desc()
    m = oracle.OracleMedium(d, *new)
    a = oracle.develop(oracle.vps_render_accum(d, m, 11, 1024))
    _, F = oracle.render_twostates(d, m, m, 13, 1024)
    assert rel_l2(F.mean((0, 1)), a.mean((0, 1))) < 2e-2
    # without emitter sampling the same expectation, by emitter hits alone
    c = small_scene(16, 12, envmap=sky(), env_scale=0.5, use_nee=False).to_desc()
    b = oracle.develop(oracle.vps_render_accum(c, oracle.OracleMedium(c, *new), 17, 4096))
    assert rel_l2(b.mean((0, 1)), a.mean((0, 1))) < 6e-2


# ---- the lane event machines compiled for the host (tests/host_emu) -----------------------------------------------
@pytest.fixture(scope="module")
def emu():
    from test_host_emu import _load_emu
    return _load_emu("librcv_emu.so")


@pytest.fixture(scope="module")
def env_setup(oracle):
    cfg = small_scene(28, 24, envmap=sky(12, 20), env_scale=0.5, env_to_world=rot_y(0.4), phase="hg", phase_g=0.5)
    desc = cfg.to_desc()
    old, new = media(14)
    return dict(desc=desc, old=old, new=new, mo=oracle.OracleMedium(desc, *old), mn=oracle.OracleMedium(desc, *new))


def test_emu_twostates_envmap(oracle, emu, env_setup):
    d = env_setup["desc"]
    (so, ao), (sn, an) = env_setup["old"], env_setup["new"]
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_twostates(C.addressof(d), so, ao, sn, an, *so.shape[:3], 21, 2, 0, acc_o, acc_n, cnt)
    oracle.counters(True)
    ref_o, ref_n = oracle.render_twostates_accum(d, env_setup["mo"], env_setup["mn"], 21, 2)
    oc = oracle.counters(True)
    assert np.array_equal(acc_o, ref_o) and np.array_equal(acc_n, ref_n)
    assert int(cnt[0]) == int(oc[0]) and int(cnt[1]) == int(oc[1])


@pytest.mark.parametrize("drt", [0, 7])
def test_emu_vps_envmap(oracle, emu, env_setup, drt):
    d = env_setup["desc"]
    sn, an = env_setup["new"]
    H, W = d.film_height, d.film_width
    shape = sn.shape[:3]
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*shape, 1), np.float32); g4 = np.zeros((*shape, 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    if not drt:
        emu.emu_vps(C.addressof(d), sn, an, *shape, 33, 1, 0, 0, None, None, acc, g1, g4, cnt)
        assert np.array_equal(acc, oracle.vps_render_accum(d, env_setup["mn"], 33, 1))
    rng = np.random.default_rng(3)
    grad_img = (rng.standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    seed_grad, spp = 0xC09848F2, 2
    film_w = oracle.film_weights(d, seed_grad, spp)
    g1[:] = 0; g4[:] = 0
    if drt:
        emu.emu_vps_drt_stages(C.addressof(d), sn, an, *shape, seed_grad, spp, oracle.alt_seed(seed_grad), grad_img.ctypes.data,
                               film_w.ctypes.data, g1, g4, cnt, 0, 1)
    else:
        emu.emu_vps(C.addressof(d), sn, an, *shape, seed_grad, spp, 1, oracle.alt_seed(seed_grad), grad_img.ctypes.data,
                    film_w.ctypes.data, acc, g1, g4, cnt)
    gs, ga = oracle.vps_backward(d, env_setup["mn"], seed_grad, spp, grad_img, film_w=film_w, use_drt=bool(drt))
    assert rel_l2(g1, gs) < 1e-5 and rel_l2(g4[..., :3], ga) < 1e-5


# ---- CUDA ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("tracer", [3, 0, 1])
def test_gpu_envmap_parity(oracle, gpu_ctx, env_setup, tracer):
    from helpers import gpu_set_states, to_cuda
    d = env_setup["desc"]
    gpu_ctx.set_scene(d)
    gpu_ctx.set_majorant_factor(0)
    gpu_ctx.set_tracer(tracer)
    gpu_set_states(gpu_ctx, env_setup["old"], env_setup["new"])
    try:
        ao, an = oracle.render_twostates_accum(d, env_setup["mo"], env_setup["mn"], 3, 1)
        go, gn = gpu_ctx.render_twostates_accum(seed=3, spp=1)
        assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
        ref = oracle.vps_render_accum(d, env_setup["mn"], 9, 1)
        assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=9, spp=1).cpu().numpy(), ref)
        H, F = oracle.render_twostates(d, env_setup["mo"], env_setup["mn"], 7, 8)
        gH, gF = gpu_ctx.render_twostates(seed=7, spp=8)
        assert rel_l2(gH.cpu().numpy(), H) < 1e-6 and rel_l2(gF.cpu().numpy(), F) < 1e-6
        rng = np.random.default_rng(5)
        Hh, Ww = d.film_height, d.film_width
        grad_img = (rng.standard_normal((Hh, Ww, 3)) / (Hh * Ww * 3)).astype(np.float32)
        for use_drt in (False, True):
            if use_drt and tracer == 1:
                continue
            gpu_ctx.set_drt(use_drt)
            gs, ga = oracle.vps_backward(d, env_setup["mn"], 0xBEEF, 4, grad_img, use_drt=use_drt)
            gpu_ctx.grad_zero()
            gpu_ctx.render_backward(0xBEEF, 4, to_cuda(grad_img))
            g_s, g_a = gpu_ctx.grad_read()
            assert rel_l2(g_s.cpu().numpy(), gs) < 1e-4 and rel_l2(g_a.cpu().numpy(), ga) < 1e-4
    finally:
        gpu_ctx.set_drt(False)
        gpu_ctx.set_tracer(3)


def test_load_dict_envmap_from_exr(tmp_path):
    """experiments.py:81-85: {'type': 'envmap', 'filename': ..., 'scale': 0.5} — the bitmap is read with exrio, reaches the
    C-ABI description as the emitter, and a scene with a second emitter is refused."""
    import torch
    from recursive_control_variates_b200 import _abi
    from recursive_control_variates_b200 import mitsuba_compat as mi
    from recursive_control_variates_b200 import plugins  # noqa: F401
    from recursive_control_variates_b200.exrio import write_exr
    tex = sky(10, 16)
    path = str(tmp_path / "sky.exr")
    write_exr(path, tex)
    _, new = media(6)
    medium = {"type": "heterogeneous", "sigma_t": {"type": "gridvolume", "data": torch.from_numpy(new[0])},
              "albedo": {"type": "gridvolume", "data": torch.from_numpy(new[1])}, "scale": 20.0}
    sd = {"type": "scene", "object": {"type": "obj", "bsdf": {"type": "null"}, "interior": medium},
          "envmap": {"type": "envmap", "filename": path, "scale": 0.5},
          "sensor": {"type": "perspective", "film": {"type": "hdrfilm", "width": 16, "height": 12}}}
    scene = mi.load_dict(sd)
    d = scene.scene_config(8, 8).to_desc()
    assert d.emitter_type == _abi.RCV_EMITTER_ENVMAP and (d.env_height, d.env_width) == (10, 16) and abs(d.env_scale - 0.5) < 1e-7
    got = np.ctypeslib.as_array((C.c_float * (10 * 16 * 3)).from_address(d.envmap_rgb)).reshape(10, 16, 3)
    assert np.allclose(got, tex, rtol=1e-6)
    sd["extra"] = {"type": "constant", "radiance": 1.0}
    with pytest.raises(ValueError):
        mi.load_dict(sd)


@pytest.mark.gpu
def test_gpu_envmap_through_the_plugins(oracle):
    """The reference's scene layout with an `envmap` emitter (experiments.py:81-85) through load_dict / `cv` / mi.render and
    backward with volpathsimple as adjoint_integrator: F and the gradients against the oracle."""
    import torch
    from recursive_control_variates_b200 import mitsuba_compat as mi
    from recursive_control_variates_b200 import plugins  # noqa: F401
    _, new = media(10)
    medium = {"type": "heterogeneous", "sigma_t": {"type": "gridvolume", "data": torch.from_numpy(new[0])},
              "albedo": {"type": "gridvolume", "data": torch.from_numpy(new[1])}, "scale": 20.0}
    sd = {"type": "scene",
          "object": {"type": "obj", "bsdf": {"type": "null"},
                     "interior": {"type": "twostatemedium", "old": medium, "new": medium, "incoming": medium}},
          "envmap": {"type": "envmap", "bitmap": sky(10, 16), "scale": 0.5},
          "sensor": {"type": "perspective", "film": {"type": "hdrfilm", "width": 24, "height": 20, "rfilter": {"type": "box"}}}}
    scene = mi.load_dict(sd)
    integ = mi.load_dict({"type": "cv", "integrator": {"type": "twostateprbvolpath", "max_depth": 8, "rr_depth": 8},
                          "adjoint_integrator": {"type": "volpathsimple", "max_depth": 8, "rr_depth": 8, "use_drt": False}})
    opt = mi.Adam(lr=5e-3)
    sp = mi.traverse(scene)
    ks, ka = "object.interior_medium.sigma_t.data", "object.interior_medium.albedo.data"
    for _ in range(2):
        opt[ks] = torch.from_numpy(new[0]); opt[ka] = torch.from_numpy(new[1])
        sp.update(opt)
    img = mi.render(scene, sp, integrator=integ, seed=1, spp=4, seed_grad=2, spp_grad=4)
    d = integ.integrator._ctx(scene).desc
    m = oracle.OracleMedium(d, *new)
    _, F = oracle.render_twostates(d, m, m, 1, 4)
    assert rel_l2(integ.F.cpu().numpy(), F) < 1e-6 and F.max() > 0
    img.sum().backward()
    gs, ga = oracle.vps_backward(d, m, 2, 4, np.ones((20, 24, 3), np.float32))
    assert rel_l2(opt[ks].grad.cpu().numpy(), gs) < 1e-4 and rel_l2(opt[ka].grad.cpu().numpy(), ga) < 1e-4
