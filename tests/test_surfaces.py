"""Surfaces + area light + the inherited PRB adjoint (BASELINE.json configs[0..1] "cube in a
box with an area light"; twostatevolpath.py:283-379, :455-477 and upstream prbvolpath): the oracle's restatement against
closed forms, estimator identities and finite differences; the host-emulated lane machines and the CUDA path against
the oracle."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import media, rel_l2, small_scene
from recursive_control_variates_b200.synthetic import boxed_quads, make_medium


def boxed(width=24, height=20, **kw):
    quads = kw.pop("quads", None) or boxed_quads(light_size=kw.pop("light_size", 1.0))
    return small_scene(width, height, quads=quads, **kw)


def test_direct_illumination_matches_quadrature(oracle):
    """Empty medium, max_depth = 2: a pixel sees rho / pi * L * integral over the light of cos cos' / r^2 at the wall
    point behind it.  Pins the rectangle intersection, the area light's sampling density and emission side, the diffuse
    BSDF and the MIS between emitter sampling and BSDF sampling against an independent numpy quadrature."""
    # reflectance 0.05: the reference counts emitter HITS of a path whose depth reached max_depth (the surface block of
    # twostatevolpath.py:283-308 runs once more after :253 switched the lane off), a rho^2 term kept below the tolerance
    cfg = boxed(16, 16, max_depth=2, rr_depth=64, quads=boxed_quads(reflectance=0.05, light_size=1.0))
    d = cfg.to_desc()
    sig = np.zeros((4, 4, 4, 1), np.float32); alb = np.zeros((4, 4, 4, 3), np.float32)
    m = oracle.OracleMedium(d, sig, alb)
    spp = 1024
    _, img = oracle.render_twostates(d, m, m, 3, spp)
    # numpy: camera rays through the pixel centres, nearest wall, quadrature over the light
    W = H = 16
    o = np.array([0.5, 0.5, -2.0]); tan = math.tan(math.radians(45.0) / 2)
    lo, hi = np.array([-1.0, -1.0, -2.5]), np.array([2.0, 2.0, 2.0])
    ls = (np.arange(32) + 0.5) / 32
    lx, lz = np.meshgrid(ls, ls)
    lp = np.stack([lx, np.full_like(lx, 1.99), lz], -1).reshape(-1, 3); dA = 1.0 / lp.shape[0]
    ref = np.zeros((H, W))
    for j in range(H):
        for i in range(W):
            sx, sy = (i + 0.5) / W, (j + 0.5) / H
            dr = np.array([(1 - 2 * sx) * tan, (1 - 2 * sy) * tan, 1.0])      # camera x = left, y = up
            dr /= np.linalg.norm(dr)
            with np.errstate(divide="ignore"):
                t = np.where(dr > 0, (hi - o) / dr, (lo - o) / dr)
            a = int(np.argmin(t)); x = o + t[a] * dr
            n = np.zeros(3); n[a] = -1.0 if dr[a] > 0 else 1.0
            v = lp - x; r2 = (v * v).sum(-1); v /= np.sqrt(r2)[:, None]
            cos_x = np.clip(v @ n, 0, None); cos_l = np.clip(v[:, 1], 0, None)  # the light faces -y: cos' = v.y
            ref[j, i] = 0.05 / math.pi * 10.0 * float((cos_x * cos_l / r2).sum()) * dA
    inner = (slice(1, -1), slice(1, -1))      # pixels on the box edges straddle two walls: centre sample vs. box filter
    assert np.abs(img[..., 0] - ref)[inner].max() < 0.06 * ref.max(), np.abs(img[..., 0] - ref)[inner].max()
    assert abs(img[..., 0][inner].mean() - ref[inner].mean()) < 0.015 * ref[inner].mean()


def test_emitter_sampling_and_emitter_hits_agree(oracle):
    """With and without next event estimation the same expectation: area-light pdfs, BSDF pdfs and MIS weights are
    consistent, through a scattering medium (shadow rays cross the null boundary twice)."""
    _, new = media(8)
    kw = dict(max_depth=6, rr_depth=64, light_size=2.5)
    a = boxed(8, 8, **kw).to_desc(); b = boxed(8, 8, use_nee=False, **kw).to_desc()
    ma, mb = oracle.OracleMedium(a, *new), oracle.OracleMedium(b, *new)
    _, Fa = oracle.render_twostates(a, ma, ma, 3, 2048)
    _, Fb = oracle.render_twostates(b, mb, mb, 4, 8192)
    assert rel_l2(Fb.mean((0, 1)), Fa.mean((0, 1))) < 3e-2


def test_two_state_estimator_in_the_box(oracle):
    """old == new: H and F bit for bit; old != new: E[H] is the render of the old state."""
    old, new = media(8)
    d = boxed(12, 10).to_desc()
    mo, mn = oracle.OracleMedium(d, *old), oracle.OracleMedium(d, *new)
    H, F = oracle.render_twostates(d, mn, mn, 5, 2)
    assert np.array_equal(H, F) and F.max() > 0
    big_new = (np.clip(new[0] * 1.5, 0, 1).astype(np.float32), new[1])
    mb = oracle.OracleMedium(d, *big_new)
    H, F = oracle.render_twostates(d, mo, mb, 7, 1024)
    Ho, _ = oracle.render_twostates(d, mo, mo, 8, 1024)
    _, Fb = oracle.render_twostates(d, mb, mb, 9, 1024)
    assert rel_l2(H.mean((0, 1)), Ho.mean((0, 1))) < 2e-2 and rel_l2(F.mean((0, 1)), Fb.mean((0, 1))) < 2e-2
    assert rel_l2(Ho.mean((0, 1)), Fb.mean((0, 1))) > 3e-2          # the two states do differ


def test_prb_adjoint_matches_finite_differences(oracle):
    """Directional derivatives of sum(grad_img * image) along theta -> theta (1 +- eps) against the PRB adjoint, in the
    boxed scene (the light reaches the medium directly and by wall bounces; shadow rays and the main path both carry
    gradients)."""
    d = boxed(8, 8, max_depth=8, rr_depth=64, light_size=2.0).to_desc()
    sig, alb = make_medium(4, "cube", 0.0)
    sig = (sig * 0.4 + 0.15).astype(np.float32)
    gimg = (np.random.default_rng(0).random((8, 8, 3)) / 192).astype(np.float32)
    spp = 4096

    def loss(s, a, seed):
        m = oracle.OracleMedium(d, s, a)
        return float((oracle.render_twostates(d, m, m, seed, spp)[1].astype(np.float64) * gimg).sum())

    gs, ga = oracle.prb_backward(d, oracle.OracleMedium(d, sig, alb), 1234, spp, gimg)
    eps = 0.03
    fd_s = np.mean([(loss(sig * (1 + eps), alb, s) - loss(sig * (1 - eps), alb, s)) / (2 * eps) for s in (1, 2)])
    fd_a = np.mean([(loss(sig, alb * (1 + eps), s) - loss(sig, alb * (1 - eps), s)) / (2 * eps) for s in (1, 2)])
    print("d sigma_t: adjoint %.5g fd %.5g   d albedo: adjoint %.5g fd %.5g" % ((gs * sig).sum(), fd_s, (ga * alb).sum(), fd_a))
    assert abs(float((gs * sig).sum()) - fd_s) < 0.12 * abs(fd_s)     # at 16384 spp: -0.05230 vs -0.05151, 0.14377 vs 0.14373
    assert abs(float((ga * alb).sum()) - fd_a) < 0.05 * abs(fd_a)


def test_prb_adjoint_agrees_with_volpathsimple_under_the_environment(oracle):
    """No surfaces, constant environment: both adjoints estimate the same gradient (different estimators: ratio-tracking
    derivative at every null collision vs. the four re-sampled points of volpathsimple.py:597-620)."""
    d = small_scene(8, 8, max_depth=8, rr_depth=64).to_desc()
    sig, alb = make_medium(4, "cube", 0.0)
    sig = (sig * 0.4 + 0.15).astype(np.float32)
    m = oracle.OracleMedium(d, sig, alb)
    gimg = (np.random.default_rng(1).random((8, 8, 3)) / 192).astype(np.float32)
    g1 = oracle.prb_backward(d, m, 77, 8192, gimg)
    g2 = oracle.vps_backward(d, m, 78, 8192, gimg)
    assert abs(float((g1[0] * sig).sum()) - float((g2[0] * sig).sum())) < 0.1 * abs(float((g2[0] * sig).sum()))
    assert abs(float((g1[1] * alb).sum()) - float((g2[1] * alb).sum())) < 0.05 * abs(float((g2[1] * alb).sum()))


# ---- the surface machine (csrc/rcv_surface_wf.cuh) compiled for the host -------------------------------------------
@pytest.fixture(scope="module")
def emu():
    import subprocess, os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emu")
    subprocess.run(["make", "-j4", "-C", here, "all"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    lib = C.CDLL(os.path.join(here, "librcv_emu.so"))
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    lib.emu_surface.argtypes = [C.c_void_p, f32, f32, f32, f32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, f32, f32, f32, f32, u64, C.c_int]
    return lib


SCENES = {
    "box": dict(),
    "box_hg_rr": dict(phase="hg", phase_g=0.6, rr_depth=2, max_depth=12),
    "box_gaussian_hidden": dict(rfilter="gaussian", hide_emitters=True, max_depth=5),
    "box_no_nee": dict(use_nee=False, max_depth=6),
    "walls_under_envmap": dict(quads=boxed_quads()[:1] + boxed_quads()[2:3], envmap="sky"),      # a floor and one wall, no area light
    "big_light": dict(light_size=2.8, max_depth=3),
}


def surf_setup(oracle, name, n=10, res=(20, 16)):
    kw = dict(SCENES[name])
    if kw.get("envmap") == "sky":
        from test_envmap import sky
        kw["envmap"] = sky(10, 16); kw["env_scale"] = 0.5
    cfg = boxed(*res, **kw)
    desc = cfg.to_desc()
    old, new = media(n)
    return dict(desc=desc, old=old, new=new, mo=oracle.OracleMedium(desc, *old), mn=oracle.OracleMedium(desc, *new))


def run_emu(emu, d, old, new, seed, spp, mode=0, adj=0, grad_img=None, film_w=None, factor=0):
    (so, ao), (sn, an) = old, new
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32)
    g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_surface(C.addressof(d), so, ao, sn, an, *sn.shape[:3], seed, spp, mode, adj,
                    grad_img.ctypes.data if grad_img is not None else None, film_w.ctypes.data if film_w is not None else None,
                    acc_o, acc_n, g1, g4, cnt, factor)
    return acc_o, acc_n, g1, g4, cnt


@pytest.mark.parametrize("name", sorted(SCENES))
def test_emu_surface_machine_matches_oracle(oracle, emu, name):
    s = surf_setup(oracle, name)
    d = s["desc"]
    acc_o, acc_n, _, _, cnt = run_emu(emu, d, s["old"], s["new"], 21, 2)
    oracle.counters(True)
    ref_o, ref_n = oracle.render_twostates_accum(d, s["mo"], s["mn"], 21, 2)
    oc = oracle.counters(True)
    assert ref_n[..., :3].max() > 0
    assert np.array_equal(acc_o, ref_o) and np.array_equal(acc_n, ref_n)
    assert int(cnt[0]) == int(oc[0]) and int(cnt[1]) == int(oc[1])
    # single-state (mode 2) = the primal of the inherited adjoint
    _, acc_n, _, _, _ = run_emu(emu, d, s["old"], s["new"], 22, 1, mode=2)
    _, ref_n = oracle.render_twostates_accum(d, s["mn"], s["mn"], 22, 1)
    assert np.array_equal(acc_n, ref_n)


@pytest.mark.parametrize("name", ["box", "box_hg_rr", "box_gaussian_hidden", "walls_under_envmap"])
def test_emu_prb_adjoint_matches_oracle(oracle, emu, name):
    s = surf_setup(oracle, name)
    d = s["desc"]
    H, W = d.film_height, d.film_width
    grad_img = (np.random.default_rng(3).standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    seed_grad, spp = 0xC09848F2, 2
    film_w = oracle.film_weights(d, seed_grad, spp)
    _, _, g1, g4, cnt = run_emu(emu, d, s["new"], s["new"], seed_grad, spp, mode=2, adj=1, grad_img=grad_img, film_w=film_w)
    oracle.counters(True)
    gs, ga = oracle.prb_backward(d, s["mn"], seed_grad, spp, grad_img, film_w=film_w)
    oc = oracle.counters(True)
    assert np.abs(gs).max() > 0 and np.abs(ga).max() > 0
    assert rel_l2(g1, gs) < 1e-5 and rel_l2(g4[..., :3], ga) < 1e-5
    assert int(cnt[4]) == int(oc[4]) and int(cnt[5]) == int(oc[5]) and int(cnt[2]) == int(oc[0])


def test_emu_surface_supergrid(oracle, emu):
    s = surf_setup(oracle, "box", n=12)
    d = s["desc"]
    mo, mn = oracle.OracleMedium(d, *s["old"], majorant_factor=4), oracle.OracleMedium(d, *s["new"], majorant_factor=4)
    acc_o, acc_n, _, _, _ = run_emu(emu, d, s["old"], s["new"], 5, 1, factor=4)
    ref_o, ref_n = oracle.render_twostates_accum(d, mo, mn, 5, 1)
    assert np.array_equal(acc_o, ref_o) and np.array_equal(acc_n, ref_n)


# ---- CUDA, through the C-ABI ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_gpu_surface_parity(oracle, gpu_ctx, name):
    from helpers import gpu_set_states, to_cuda
    from recursive_control_variates_b200 import _abi
    s = surf_setup(oracle, name, n=12, res=(28, 24))
    d = s["desc"]
    gpu_ctx.set_scene(d)
    gpu_ctx.set_majorant_factor(0)
    gpu_ctx.set_drt(False)
    gpu_set_states(gpu_ctx, s["old"], s["new"])
    # per-lane radiance, bit for bit (spp = 1; the gaussian filter splats a lane to several pixels: images instead)
    ao, an = oracle.render_twostates_accum(d, s["mo"], s["mn"], 3, 1)
    go, gn = gpu_ctx.render_twostates_accum(seed=3, spp=1)
    if d.rfilter_type == 0:
        assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
    else:
        assert rel_l2(go.cpu().numpy(), ao) < 1e-6 and rel_l2(gn.cpu().numpy(), an) < 1e-6
    H, F = oracle.render_twostates(d, s["mo"], s["mn"], 7, 8)
    gH, gF = gpu_ctx.render_twostates(seed=7, spp=8)
    assert rel_l2(gH.cpu().numpy(), H) < 1e-6 and rel_l2(gF.cpu().numpy(), F) < 1e-6          # tolerance of north_star: 1e-4
    a1, _ = oracle.render_twostates_accum(d, s["mn"], s["mn"], 5, 1)
    assert rel_l2(gpu_ctx.render_primal_accum(1, seed=5, spp=1).cpu().numpy(), a1) < 1e-6
    # the inherited adjoint
    Hh, Ww = d.film_height, d.film_width
    grad_img = (np.random.default_rng(5).standard_normal((Hh, Ww, 3)) / (Hh * Ww * 3)).astype(np.float32)
    gs, ga = oracle.prb_backward(d, s["mn"], 0xBEEF, 4, grad_img)
    gpu_ctx.grad_zero()
    gpu_ctx.render_backward(0xBEEF, 4, to_cuda(grad_img), adjoint=_abi.RCV_ADJOINT_PRBVOLPATH)
    g_s, g_a = gpu_ctx.grad_read()
    assert rel_l2(g_s.cpu().numpy(), gs) < 1e-4 and rel_l2(g_a.cpu().numpy(), ga) < 1e-4     # north_star: 1e-3
    # spp shards add up
    gpu_ctx.grad_zero()
    fw = gpu_ctx.film_weights(0xBEEF, 4)
    for b, e in ((0, 1), (1, 4)):
        gpu_ctx.render_backward(0xBEEF, 4, to_cuda(grad_img), b, e, film_w=fw, adjoint=_abi.RCV_ADJOINT_PRBVOLPATH)
    g_s2, _ = gpu_ctx.grad_read()
    assert rel_l2(g_s2.cpu().numpy(), gs) < 1e-4
    # volpathsimple refuses surfaces, as the reference does (volpathsimple.py:24-30)
    with pytest.raises(_abi.RcvError):
        gpu_ctx.render_backward(1, 1, to_cuda(grad_img))


@pytest.mark.gpu
def test_gpu_prb_adjoint_without_surfaces(oracle, gpu_ctx):
    """The inherited adjoint in the environment-lit scene (supergrid on): what `cv` runs when no adjoint_integrator is given."""
    from helpers import gpu_set_states, to_cuda
    from recursive_control_variates_b200 import _abi
    d = small_scene(24, 20).to_desc()
    old, new = media(12)
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, old, new)
    grad_img = (np.random.default_rng(6).standard_normal((20, 24, 3)) / (20 * 24 * 3)).astype(np.float32)
    for f in (0, 4):
        gpu_ctx.set_majorant_factor(f)
        gs, ga = oracle.prb_backward(d, oracle.OracleMedium(d, *new, majorant_factor=f), 99, 4, grad_img)
        gpu_ctx.grad_zero()
        gpu_ctx.render_backward(99, 4, to_cuda(grad_img), adjoint=_abi.RCV_ADJOINT_PRBVOLPATH)
        g_s, g_a = gpu_ctx.grad_read()
        assert rel_l2(g_s.cpu().numpy(), gs) < 1e-4 and rel_l2(g_a.cpu().numpy(), ga) < 1e-4
    gpu_ctx.set_majorant_factor(0)


def _boxed_scene_dict(new, film=(24, 20)):
    import torch
    from recursive_control_variates_b200.scene import look_at

    def rect(q):
        # a to_world that maps Mitsuba's rectangle [-1,1]^2 x {0} onto p0 + s eu + t ev
        p0, eu, ev = (np.asarray(q[k], np.float64) for k in ("p0", "eu", "ev"))
        m = np.eye(4); m[:3, 0] = eu / 2; m[:3, 1] = ev / 2
        m[:3, 2] = np.cross(eu, ev) / np.linalg.norm(np.cross(eu, ev)); m[:3, 3] = p0 + eu / 2 + ev / 2
        out = {"type": "rectangle", "to_world": m,
               "bsdf": {"type": "diffuse", "reflectance": {"type": "rgb", "value": list(q["reflectance"])}}}
        if "radiance" in q:
            out["emitter"] = {"type": "area", "radiance": {"type": "rgb", "value": list(q["radiance"])}}
        return out

    medium = {"type": "heterogeneous", "sigma_t": {"type": "gridvolume", "data": torch.from_numpy(new[0])},
              "albedo": {"type": "gridvolume", "data": torch.from_numpy(new[1])}, "scale": 20.0}
    d = {"type": "scene",
         "object": {"type": "obj", "bsdf": {"type": "null"},
                    "interior": {"type": "twostatemedium", "old": medium, "new": medium, "incoming": medium}},
         "sensor": {"type": "perspective", "fov": 45, "to_world": look_at([0.5, 0.5, -2.0], [0.5, 0.5, 0.5], [0, 1, 0]),
                    "film": {"type": "hdrfilm", "width": film[0], "height": film[1], "rfilter": {"type": "box"}}}}
    for k, q in enumerate(boxed_quads()):
        d[f"wall_{k}"] = rect(q)
    return d


def test_load_dict_rectangles_and_area_light():
    """The Mitsuba-style scene dictionary (rectangle shapes with to_world, diffuse BSDFs, an area emitter) reaches the
    C-ABI description unchanged; a second emitter is refused."""
    from recursive_control_variates_b200 import mitsuba_compat as mi
    from recursive_control_variates_b200 import plugins  # noqa: F401
    from recursive_control_variates_b200 import _abi
    _, new = media(6)
    sd = _boxed_scene_dict(new)
    scene = mi.load_dict(sd)
    d = scene.scene_config(8, 8).to_desc()
    ref = boxed(24, 20, max_depth=8, rr_depth=8).to_desc()
    assert d.n_quads == 7 and d.area_quad == 6 and d.emitter_type == _abi.RCV_EMITTER_AREA
    for k in range(7):
        for f in ("p0", "eu", "ev", "reflectance", "radiance"):
            assert np.allclose(list(getattr(d.quads[k], f)), list(getattr(ref.quads[k], f)), atol=1e-6), (k, f)
    assert abs(d.env_dist - ref.env_dist) < 1e-6
    sd["sky"] = {"type": "constant", "radiance": 1.0}
    with pytest.raises(ValueError):
        mi.load_dict(sd)


@pytest.mark.gpu
def test_gpu_boxed_scene_through_the_plugins(oracle):
    """load_dict of a boxed scene (rectangles + area light), `cv` around twostateprbvolpath with its INHERITED adjoint
    (no adjoint_integrator given, cv_integrator.py:21): F against the oracle, the gradient of sum(img) against the
    oracle's prbvolpath adjoint."""
    import torch
    from recursive_control_variates_b200 import mitsuba_compat as mi
    from recursive_control_variates_b200 import plugins  # noqa: F401  (registers the integrators)
    _, new = media(10)
    scene = mi.load_dict(_boxed_scene_dict(new))
    integ = mi.load_dict({"type": "cv", "integrator": {"type": "twostateprbvolpath", "max_depth": 8, "rr_depth": 8}})
    assert integ.adjoint_integrator is integ.integrator
    opt = mi.Adam(lr=5e-3)
    sp = mi.traverse(scene)
    ks, ka = "object.interior_medium.sigma_t.data", "object.interior_medium.albedo.data"
    for _ in range(2):
        opt[ks] = torch.from_numpy(new[0]); opt[ka] = torch.from_numpy(new[1])
        sp.update(opt)
    img = mi.render(scene, sp, integrator=integ, seed=1, spp=4, seed_grad=2, spp_grad=4)
    d = integ.integrator._ctx(scene).desc
    m = oracle.OracleMedium(d, *new)
    H, F = oracle.render_twostates(d, m, m, 1, 4)
    assert rel_l2(integ.F.cpu().numpy(), F) < 1e-6 and F.max() > 0
    img.sum().backward()
    gs, ga = oracle.prb_backward(d, m, 2, 4, np.ones((20, 24, 3), np.float32))
    assert rel_l2(opt[ks].grad.cpu().numpy(), gs) < 1e-4 and rel_l2(opt[ka].grad.cpu().numpy(), ga) < 1e-4


# ---- committed golden vectors of the round-2 features (tests/golden/oracle_r2.npz) ---------------------------------------
def _golden_r2():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_r2.npz"))


def test_golden_r2_oracle(oracle):
    from golden.make_golden_r2 import compute
    ref = _golden_r2()
    out, _ = compute(oracle)
    assert set(out) == set(ref.files)
    for k, v in out.items():
        if "grad" in k:
            assert rel_l2(v, ref[k]) < 1e-6, k          # accumulated in float64, rounded once
        else:
            assert np.array_equal(v, ref[k]), k


@pytest.mark.gpu
def test_golden_r2_on_gpu(gpu_ctx):
    """The CUDA path reproduces the committed vectors (no oracle in the loop)."""
    from golden.make_golden_r2 import configs
    from helpers import gpu_set_states, to_cuda
    from recursive_control_variates_b200 import _abi
    ref = _golden_r2()
    env, box, old, new = configs()
    gimg = (np.arange(16 * 12 * 3, dtype=np.float32).reshape(12, 16, 3) % 7 - 3) / (16 * 12 * 3)
    gpu_ctx.set_majorant_factor(0); gpu_ctx.set_drt(False)
    gpu_ctx.set_scene(env.to_desc())
    gpu_set_states(gpu_ctx, old, new)
    go, gn = gpu_ctx.render_twostates_accum(seed=7, spp=1)
    assert np.array_equal(go.cpu().numpy(), ref["env_acc_old"]) and np.array_equal(gn.cpu().numpy(), ref["env_acc_new"])
    assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=8, spp=1).cpu().numpy(), ref["env_vps_acc"])
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(0x5D8714D7, 2, to_cuda(gimg))
    gs, ga = gpu_ctx.grad_read()
    assert rel_l2(gs.cpu().numpy(), ref["env_grad_sigma_t"]) < 1e-4 and rel_l2(ga.cpu().numpy(), ref["env_grad_albedo"]) < 1e-4
    gpu_ctx.set_scene(box.to_desc())
    gpu_set_states(gpu_ctx, old, new)
    go, gn = gpu_ctx.render_twostates_accum(seed=7, spp=1)
    assert np.array_equal(go.cpu().numpy(), ref["box_acc_old"]) and np.array_equal(gn.cpu().numpy(), ref["box_acc_new"])
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(0x5D8714D7, 2, to_cuda(gimg), adjoint=_abi.RCV_ADJOINT_PRBVOLPATH)
    gs, ga = gpu_ctx.grad_read()
    assert rel_l2(gs.cpu().numpy(), ref["box_grad_sigma_t"]) < 1e-4 and rel_l2(ga.cpu().numpy(), ref["box_grad_albedo"]) < 1e-4
