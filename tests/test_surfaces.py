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
    H, F = oracle.render_twostates(d, mo, mb, 7, 4096)
    Ho, _ = oracle.render_twostates(d, mo, mo, 8, 4096)
    _, Fb = oracle.render_twostates(d, mb, mb, 9, 4096)
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
    spp = 16384

    def loss(s, a, seed):
        m = oracle.OracleMedium(d, s, a)
        return float((oracle.render_twostates(d, m, m, seed, spp)[1].astype(np.float64) * gimg).sum())

    gs, ga = oracle.prb_backward(d, oracle.OracleMedium(d, sig, alb), 1234, spp, gimg)
    eps = 0.03
    fd_s = np.mean([(loss(sig * (1 + eps), alb, s) - loss(sig * (1 - eps), alb, s)) / (2 * eps) for s in (1, 2)])
    fd_a = np.mean([(loss(sig, alb * (1 + eps), s) - loss(sig, alb * (1 - eps), s)) / (2 * eps) for s in (1, 2)])
    print("d sigma_t: adjoint %.5g fd %.5g   d albedo: adjoint %.5g fd %.5g" % ((gs * sig).sum(), fd_s, (ga * alb).sum(), fd_a))
    assert abs(float((gs * sig).sum()) - fd_s) < 0.1 * abs(fd_s)
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
