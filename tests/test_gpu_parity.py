"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle
on identical seeded inputs.  Per-lane radiance is bit-exact by construction
(DESIGN.md §2), so spp=1 accumulators must be *equal*; images with several
samples per pixel and gradients differ only by float summation order
(tolerances of BASELINE.json: 1e-4 rel. L2 for images, 1e-3 for gradients)."""
import numpy as np
import pytest

from helpers import gpu_set_states, media, rel_l2, small_scene, to_cuda

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4    # BASELINE.json north_star: primal and CV images
GRAD_TOL = 1e-3   # BASELINE.json north_star: grid gradient


@pytest.fixture(scope="module")
def setup(oracle, gpu_ctx):
    cfg = small_scene()
    desc = cfg.to_desc()
    old, new = media(24)
    mo, mn = oracle.OracleMedium(desc, *old), oracle.OracleMedium(desc, *new)
    gpu_ctx.set_scene(desc)
    gpu_set_states(gpu_ctx, old, new)
    return dict(cfg=cfg, desc=desc, old=old, new=new, mo=mo, mn=mn)


@pytest.mark.parametrize("tracer", [0, 1, 2, 3])
def test_twostates_per_lane_bit_exact(oracle, gpu_ctx, setup, tracer):
    gpu_ctx.set_tracer(tracer)
    d = setup["desc"]
    oracle.counters(True); gpu_ctx.counters(True)
    ao, an = oracle.render_twostates_accum(d, setup["mo"], setup["mn"], seed=3, spp=1)
    go, gn = gpu_ctx.render_twostates_accum(seed=3, spp=1)
    oc, gc = oracle.counters(True), gpu_ctx.counters(True)
    assert np.array_equal(go.cpu().numpy(), ao), "old-state radiance differs from the oracle"
    assert np.array_equal(gn.cpu().numpy(), an), "new-state radiance differs from the oracle"
    assert int(oc[0]) == int(gc[0]) and int(oc[1]) == int(gc[1]), (oc, gc)
    assert int(gc[6]) == d.film_width * d.film_height
    gpu_ctx.set_tracer(3)


def test_majorants(oracle, gpu_ctx, setup):
    m_old, m_new = gpu_ctx.majorants()
    assert m_old == setup["mo"].majorant and m_new == setup["mn"].majorant


def test_twostates_image(oracle, gpu_ctx, setup):
    d = setup["desc"]
    H, F = oracle.render_twostates(d, setup["mo"], setup["mn"], seed=7, spp=16)
    gH, gF = gpu_ctx.render_twostates(seed=7, spp=16)
    assert rel_l2(gH.cpu().numpy(), H) < IMG_TOL
    assert rel_l2(gF.cpu().numpy(), F) < IMG_TOL
    assert rel_l2(gH.cpu().numpy(), H) < 1e-6   # only summation order differs


def test_identity_H_equals_F(gpu_ctx, setup):
    gpu_set_states(gpu_ctx, setup["new"], setup["new"])
    H, F = gpu_ctx.render_twostates(seed=11, spp=4)
    # identical states give identical per-lane radiance; the two film accumulators receive their atomic adds in
    # (possibly) different orders, so the images agree to rounding, not bit for bit
    assert rel_l2(H.cpu().numpy(), F.cpu().numpy()) < 1e-6
    gpu_set_states(gpu_ctx, setup["old"], setup["new"])


def test_single_state_primal(oracle, gpu_ctx, setup):
    d = setup["desc"]
    for state, m in ((0, setup["mo"]), (1, setup["mn"])):
        a, _ = oracle.render_twostates_accum(d, m, m, seed=5, spp=1)
        g = gpu_ctx.render_primal_accum(state, seed=5, spp=1)
        assert np.array_equal(g.cpu().numpy(), a)


def test_spp_sharding_invariance(gpu_ctx, setup):
    full_o, full_n = gpu_ctx.render_twostates_accum(seed=9, spp=8)
    acc = None
    for b, e in ((0, 3), (3, 8)):
        acc = gpu_ctx.render_twostates_accum(seed=9, spp=8, spp_begin=b, spp_end=e, out=acc)
    assert rel_l2(acc[0].cpu().numpy(), full_o.cpu().numpy()) < 1e-6
    assert rel_l2(acc[1].cpu().numpy(), full_n.cpu().numpy()) < 1e-6
    assert np.array_equal(acc[0][..., 3].cpu().numpy(), full_o[..., 3].cpu().numpy())


@pytest.mark.parametrize("tracer", [0, 1, 2, 3])
def test_volpathsimple_primal_bit_exact(oracle, gpu_ctx, setup, tracer):
    gpu_ctx.set_tracer(tracer)
    d = setup["desc"]
    oracle.counters(True); gpu_ctx.counters(True)
    a = oracle.vps_render_accum(d, setup["mn"], seed=13, spp=1)
    g = gpu_ctx.render_adjoint_primal_accum(seed=13, spp=1)
    oc, gc = oracle.counters(True), gpu_ctx.counters(True)
    assert np.array_equal(g.cpu().numpy(), a)
    assert int(oc[2]) == int(gc[2]) and int(oc[3]) == int(gc[3]), (oc, gc)
    gpu_ctx.set_tracer(3)


@pytest.mark.parametrize("tracer", [0, 1, 3])
def test_volpathsimple_backward(oracle, gpu_ctx, setup, tracer):
    gpu_ctx.set_tracer(tracer)
    d = setup["desc"]
    rng = np.random.default_rng(0)
    grad_img = (rng.standard_normal((d.film_height, d.film_width, 3)) / (d.film_height * d.film_width * 3)).astype(np.float32)
    seed_grad, spp_grad = 0x5D8714D7, 4
    oracle.counters(True); gpu_ctx.counters(True)
    gs, ga = oracle.vps_backward(d, setup["mn"], seed_grad, spp_grad, grad_img, threads=1)
    gpu_ctx.grad_zero()
    gpu_ctx.render_backward(seed_grad, spp_grad, to_cuda(grad_img))
    ggs, gga = gpu_ctx.grad_read()
    oc, gc = oracle.counters(True), gpu_ctx.counters(True)
    assert [int(v) for v in oc[3:6]] == [int(v) for v in gc[3:6]], (oc, gc)
    if tracer in (1, 2):
        assert int(oc[2]) == int(gc[2])            # straight replay: every collision is repeated
    else:
        assert int(gc[2]) < 0.7 * int(oc[2])       # vertex-list adjoint: only the shadow rays are traced again
    assert rel_l2(ggs.cpu().numpy(), gs) < GRAD_TOL
    assert rel_l2(gga.cpu().numpy(), ga) < GRAD_TOL
    assert rel_l2(ggs.cpu().numpy(), gs) < 1e-4    # float atomics vs double accumulation
    gpu_ctx.set_tracer(3)


@pytest.mark.parametrize("flags,tracer", [(7, 3), (3, 3), (5, 3), (1, 3), (7, 2), (7, 0), (7, 1)])
def test_volpathsimple_backward_drt(oracle, gpu_ctx, setup, flags, tracer):
    """use_drt = True (volpathsimple.py:39-47, :520-594, :623-778) through rcv_adjoint_set_drt: bit 1 of `flags` is
    use_drt_subsampling, bit 2 use_drt_mis.  With sub-sampling the pool tracer runs the three DRT stages on the event
    machines; every other combination runs the one-thread-per-lane kernel.  Same lanes, same draws and gradient sites
    as the oracle's restatement; the gradients differ in summation order only.  Sharding the spp changes nothing."""
    d = setup["desc"]
    rng = np.random.default_rng(1)
    grad_img = (rng.standard_normal((d.film_height, d.film_width, 3)) / (d.film_height * d.film_width * 3)).astype(np.float32)
    seed_grad, spp_grad = 0x2468ACE1, 4
    kw = dict(use_drt=True, use_drt_subsampling=bool(flags & 2), use_drt_mis=bool(flags & 4))
    oracle.counters(True); gpu_ctx.counters(True)
    gs, ga = oracle.vps_backward(d, setup["mn"], seed_grad, spp_grad, grad_img, threads=1, **kw)
    staged = tracer == 3 and (flags & 2)
    try:
        gpu_ctx.set_tracer(tracer)
        gpu_ctx.set_drt(True, bool(flags & 2), bool(flags & 4))
        gpu_ctx.grad_zero()
        gpu_ctx.render_backward(seed_grad, spp_grad, to_cuda(grad_img))
        ggs, gga = gpu_ctx.grad_read()
        oc, gc = oracle.counters(True), gpu_ctx.counters(True)
        assert [int(v) for v in oc[3:7]] == [int(v) for v in gc[3:7]], (oc, gc)
        if staged:
            assert int(gc[2]) < int(oc[2])         # vertex-list adjoint: the main path's null collisions are not repeated
        else:
            assert int(gc[2]) == int(oc[2])
        assert rel_l2(ggs.cpu().numpy(), gs) < 1e-4 and rel_l2(gga.cpu().numpy(), ga) < 1e-4
        if flags == 7 and tracer == 3:
            gpu_ctx.grad_zero()
            gpu_ctx.render_backward(seed_grad, spp_grad, to_cuda(grad_img), 0, 1)
            gpu_ctx.render_backward(seed_grad, spp_grad, to_cuda(grad_img), 1, 4)
            hs, ha = gpu_ctx.grad_read()
            assert rel_l2(hs.cpu().numpy(), ggs.cpu().numpy()) < 1e-5 and rel_l2(ha.cpu().numpy(), gga.cpu().numpy()) < 1e-5
    finally:
        gpu_ctx.set_drt(False)
        gpu_ctx.set_tracer(3)
    # a different estimator of the same gradient: far from the path-replay one sample by sample
    ps, _ = oracle.vps_backward(d, setup["mn"], seed_grad, spp_grad, grad_img, threads=1)
    assert rel_l2(gs, ps) > 1e-2


def test_drt_stages_at_bench_depth(oracle, gpu_ctx):
    """The DRT stages at the depth of the experiments (max_depth = rr_depth = 64, experiments.py:130-131) on a denser
    medium and a film larger than one pool refill, with the gaussian reconstruction filter."""
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.runtime import Context
    d = SceneConfig(width=64, height=48, rfilter="gaussian").to_desc()
    _, (sn, an) = media(20)
    mn = oracle.OracleMedium(d, sn, an)
    gimg = (np.random.default_rng(4).standard_normal((48, 64, 3)) / (48 * 64 * 3)).astype(np.float32)
    gs, ga = oracle.vps_backward(d, mn, 123, 4, gimg, use_drt=True)
    ctx = Context(0)
    ctx.set_scene(d)
    gpu_set_states(ctx, (sn, an), (sn, an))
    ctx.set_drt(True)
    ctx.grad_zero()
    ctx.render_backward(123, 4, to_cuda(gimg))
    ggs, gga = ctx.grad_read()
    assert rel_l2(ggs.cpu().numpy(), gs) < 1e-4 and rel_l2(gga.cpu().numpy(), ga) < 1e-4


def test_drt_with_majorant_supergrid(oracle, gpu_ctx):
    """DRT on the DDA kernels: the main path samples free flights through the supergrid, the DRT distance sampler
    keeps the volume's majorant."""
    from recursive_control_variates_b200.scene import SceneConfig
    d = SceneConfig(width=24, height=20, max_depth=16, rr_depth=8).to_desc()
    _, (sn, an) = media(16)
    mn = oracle.OracleMedium(d, sn, an, majorant_factor=4)
    gimg = (np.random.default_rng(2).standard_normal((20, 24, 3)) / 1440).astype(np.float32)
    gs, ga = oracle.vps_backward(d, mn, 91, 3, gimg, threads=1, use_drt=True)
    from recursive_control_variates_b200.runtime import Context
    ctx = Context(0)
    ctx.set_scene(d)
    ctx.set_majorant_factor(4)
    gpu_set_states(ctx, (sn, an), (sn, an))
    ctx.set_drt(True)
    ctx.grad_zero()
    ctx.render_backward(91, 3, to_cuda(gimg))
    ggs, gga = ctx.grad_read()
    assert rel_l2(ggs.cpu().numpy(), gs) < 1e-4 and rel_l2(gga.cpu().numpy(), ga) < 1e-4


def test_analytic_pins_on_gpu():
    """The closed forms of tests/test_oracle.py (Beer-Lambert image and gradient of a purely absorbing medium, white
    furnace) against the CUDA path directly — no oracle in the loop."""
    from helpers import analytic_transmittance
    from recursive_control_variates_b200.runtime import Context
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import make_medium
    cfg = SceneConfig(width=16, height=16, sigma_t_scale=4.0, env_radiance=(0.8, 1.0, 1.2))
    sig = np.full((6, 6, 6, 1), 0.5, np.float32); alb = np.zeros((6, 6, 6, 3), np.float32)
    ctx = Context(0)
    ctx.set_scene(cfg.to_desc())
    gpu_set_states(ctx, (sig, alb), (sig, alb))
    T, lT = analytic_transmittance(cfg, 4.0 * 0.5)
    env = np.array([0.8, 1.0, 1.2])
    spp = 512
    H_, F_ = ctx.render_twostates(seed=6, spp=spp)
    imgs = [ctx.develop(ctx.render_adjoint_primal_accum(seed=5, spp=spp)).cpu().numpy(), H_.cpu().numpy(), F_.cpu().numpy()]
    for img in imgs:
        err = img - T[..., None] * env
        assert np.abs(err).max() < 5 * 1.2 * np.sqrt(0.25 / spp) and abs(err.mean()) < 4e-3
    gimg = np.full((16, 16, 3), 1.0 / (16 * 16 * 3), np.float32)
    analytic = float((-(2.0 * lT)[..., None] * env * gimg).sum())
    for drt in (False, True):
        ctx.set_drt(drt)
        ctx.grad_zero(); ctx.render_backward(9, spp, to_cuda(gimg))
        gs, ga = ctx.grad_read()
        assert abs(float((gs.cpu().numpy().astype(np.float64) * sig).sum()) - analytic) < 0.03 * abs(analytic), drt
    # white furnace: albedo 1, uniform environment
    sig, _ = make_medium(8, "dense", 0.0)
    alb = np.ones((8, 8, 8, 3), np.float32)
    kw = dict(width=12, height=12, sigma_t_scale=6.0, env_radiance=(1, 1, 1), max_depth=256, rr_depth=256, phase="hg", phase_g=0.6)
    for nee in (True, False):
        ctx = Context(0)
        ctx.set_scene(SceneConfig(use_nee=nee, sample_emitters=nee, **kw).to_desc())
        gpu_set_states(ctx, (sig, alb), (sig, alb))
        H_, F_ = ctx.render_twostates(seed=4, spp=512 if nee else 64)
        imgs = [ctx.develop(ctx.render_adjoint_primal_accum(seed=3, spp=512 if nee else 64)).cpu().numpy(), H_.cpu().numpy(), F_.cpu().numpy()]
        for img in imgs:
            if nee:
                assert abs(float(img.mean()) - 1.0) < 4e-3 and 0.9 < img.min() and img.max() < 1.1
            else:
                assert np.abs(img - 1.0).max() < 1e-5


def test_backward_sharding(gpu_ctx, setup):
    d = setup["desc"]
    g = to_cuda(np.full((d.film_height, d.film_width, 3), 1e-4, np.float32))
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(77, 4, g)
    full_s, full_a = gpu_ctx.grad_read()
    gpu_ctx.grad_zero()
    gpu_ctx.render_backward(77, 4, g, 0, 1); gpu_ctx.render_backward(77, 4, g, 1, 4)
    s, a = gpu_ctx.grad_read()
    assert rel_l2(s.cpu().numpy(), full_s.cpu().numpy()) < 1e-4
    assert rel_l2(a.cpu().numpy(), full_a.cpu().numpy()) < 1e-4


def test_cv_update_matches_oracle(oracle, gpu_ctx):
    rng = np.random.default_rng(1)
    shape = (20, 28, 3)
    st_o = {k: np.zeros(shape, np.float32) for k in ("img", "v_n", "w_s", "mean_x", "mean_y", "var_x", "var_y", "cov")}
    st_g = {k: to_cuda(v) for k, v in st_o.items()}
    stats_n = 0
    for pre_update in (False, True):
        for it in range(8):
            base = rng.random(shape).astype(np.float32)
            H = (base + 0.05 * rng.standard_normal(shape)).astype(np.float32)
            F = (base + 0.05 * rng.standard_normal(shape)).astype(np.float32)
            oracle.cv_update(H, F, st_o, it, 1, stats_n, 0.9, 0.999, pre_update)
            gpu_ctx.cv_update(to_cuda(H), to_cuda(F), st_g, it, 1, stats_n, 0.9, 0.999, pre_update)
            if it > 0:
                stats_n += 1
            for k in st_o:
                assert np.array_equal(st_g[k].cpu().numpy(), st_o[k], equal_nan=True), (k, it, pre_update)


def test_loss_adam_clamp_upsample(oracle, gpu_ctx):
    rng = np.random.default_rng(2)
    img = rng.random((16, 16, 3)).astype(np.float32); ref = rng.random((16, 16, 3)).astype(np.float32)
    for kind in (0, 1):
        lo, go = oracle.loss_grad(kind, img, ref)
        lg, gg = gpu_ctx.loss_grad(kind, to_cuda(img), to_cuda(ref))
        assert abs(float(lg.item()) - lo) < 1e-6 * max(1.0, abs(lo))
        assert np.array_equal(gg.cpu().numpy(), go)
    p = rng.random(1000).astype(np.float32); g = rng.standard_normal(1000).astype(np.float32)
    m = np.zeros(1000, np.float32); v = np.zeros(1000, np.float32)
    gp, gm, gv = to_cuda(p), to_cuda(m), to_cuda(v)
    for t in range(1, 4):
        oracle.adam(p, g, m, v, t, 5e-3)
        gpu_ctx.adam_flat(gp, to_cuda(g), gm, gv, t, 5e-3)
        assert np.array_equal(gp.cpu().numpy(), p) and np.array_equal(gm.cpu().numpy(), m) and np.array_equal(gv.cpu().numpy(), v)
    oracle.clamp(p, 0.2, 0.7); gpu_ctx.clamp(gp, 0.2, 0.7)
    assert np.array_equal(gp.cpu().numpy(), p)
    a = rng.random((3, 4, 5, 3)).astype(np.float32)
    assert np.array_equal(gpu_ctx.upsample2x(to_cuda(a)).cpu().numpy(), oracle.upsample2x(a))


def test_rotation_semantics(gpu_ctx, setup):
    s_old, a_old = gpu_ctx.medium_read(0)
    s_new, a_new = gpu_ctx.medium_read(1)
    assert np.array_equal(s_old.cpu().numpy(), setup["old"][0]) and np.array_equal(a_new.cpu().numpy(), setup["new"][1])
    assert np.array_equal(s_new.cpu().numpy(), setup["new"][0]) and np.array_equal(a_old.cpu().numpy(), setup["old"][1])


def test_golden_fixture_on_gpu(oracle, gpu_ctx):
    """The committed golden vectors (tests/golden/oracle_small.npz) reproduced by the CUDA path."""
    import os
    from golden.make_golden import config
    ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small.npz"))
    desc, old, new = config()
    gpu_ctx.set_scene(desc)
    gpu_set_states(gpu_ctx, old, new)
    go, gn = gpu_ctx.render_twostates_accum(seed=7, spp=1)
    assert np.array_equal(go.cpu().numpy(), ref["acc_old"]) and np.array_equal(gn.cpu().numpy(), ref["acc_new"])
    assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=8, spp=1).cpu().numpy(), ref["vps_acc"])
    gimg = (np.arange(16 * 12 * 3, dtype=np.float32).reshape(12, 16, 3) % 7 - 3) / (16 * 12 * 3)
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(0x5D8714D7, 2, to_cuda(gimg))
    gs, ga = gpu_ctx.grad_read()
    assert rel_l2(gs.cpu().numpy(), ref["grad_sigma_t"]) < GRAD_TOL and rel_l2(ga.cpu().numpy(), ref["grad_albedo"]) < GRAD_TOL
    m = gpu_ctx.majorants()
    assert np.array_equal(np.array(m, np.float32), ref["majorants"])
    # the adjoint with the DRT term: sub-sampling + MIS (three stages on the pool machines), and neither (per-lane kernel)
    try:
        for sub_mis, key in ((True, "drt"), (False, "drt_plain")):
            gpu_ctx.set_drt(True, sub_mis, sub_mis)
            gpu_ctx.grad_zero(); gpu_ctx.render_backward(0x5D8714D7, 2, to_cuda(gimg))
            gs, ga = gpu_ctx.grad_read()
            assert rel_l2(gs.cpu().numpy(), ref[key + "_grad_sigma_t"]) < GRAD_TOL
            assert rel_l2(ga.cpu().numpy(), ref[key + "_grad_albedo"]) < GRAD_TOL
    finally:
        gpu_ctx.set_drt(False)
    # majorant supergrid (factor 4) fixtures
    gpu_ctx.set_majorant_factor(4)
    try:
        mg_o, mg_n = gpu_ctx.majorant_grid()
        assert np.array_equal(mg_o, ref["sg_majorant_grid_old"]) and np.array_equal(mg_n, ref["sg_majorant_grid_new"])
        go, gn = gpu_ctx.render_twostates_accum(seed=7, spp=1)
        assert np.array_equal(go.cpu().numpy(), ref["sg_acc_old"]) and np.array_equal(gn.cpu().numpy(), ref["sg_acc_new"])
        assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=8, spp=1).cpu().numpy(), ref["sg_vps_acc"])
        gpu_ctx.grad_zero(); gpu_ctx.render_backward(0x5D8714D7, 2, to_cuda(gimg))
        gs, ga = gpu_ctx.grad_read()
        assert rel_l2(gs.cpu().numpy(), ref["sg_grad_sigma_t"]) < GRAD_TOL and rel_l2(ga.cpu().numpy(), ref["sg_grad_albedo"]) < GRAD_TOL
    finally:
        gpu_ctx.set_majorant_factor(0)


@pytest.mark.parametrize("tracer", [0, 1, 2, 3])
def test_hg_phase_rr_and_scaled_box(oracle, gpu_ctx, tracer):
    """HG phase function, active Russian roulette, hide_emitters and a non-unit medium box
    (janga scales the cube by [1, 1.2, 1.5], experiments.py:41,58): all three schedules vs the oracle."""
    from recursive_control_variates_b200.scene import SceneConfig, look_at
    cfg = SceneConfig(width=40, height=28, phase="hg", phase_g=0.877, max_depth=6, rr_depth=2, hide_emitters=True,
                      bbox_min=(0.0, 0.0, 0.0), bbox_max=(1.0, 1.2, 1.5), env_radiance=(0.9, 1.0, 1.1),
                      to_world=look_at([2.2, 0.3, 2.4], [0.5, 0.6, 0.75], [0, 1, 0]))
    d = cfg.to_desc()
    old, new = media(14)
    mo, mn = oracle.OracleMedium(d, *old), oracle.OracleMedium(d, *new)
    gpu_ctx.set_tracer(tracer)
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, old, new)
    ao, an = oracle.render_twostates_accum(d, mo, mn, seed=21, spp=1)
    go, gn = gpu_ctx.render_twostates_accum(seed=21, spp=1)
    assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
    assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=22, spp=1).cpu().numpy(), oracle.vps_render_accum(d, mn, 22, 1))
    gimg = np.full((d.film_height, d.film_width, 3), 1e-3, np.float32)
    gs, ga = oracle.vps_backward(d, mn, 23, 3, gimg, threads=1)
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(23, 3, to_cuda(gimg))
    ggs, gga = gpu_ctx.grad_read()
    assert rel_l2(ggs.cpu().numpy(), gs) < GRAD_TOL and rel_l2(gga.cpu().numpy(), ga) < GRAD_TOL
    gpu_ctx.set_tracer(3)


def test_large_config_properties(gpu_ctx):
    """Size-independent properties at a BASELINE-sized grid (128^3, 256x256, 16 spp): identity H == F,
    shard invariance, film weights == spp, finite gradients (the oracle would take minutes here)."""
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import make_medium
    d = SceneConfig(width=256, height=256).to_desc()
    gpu_ctx.set_scene(d)
    new = make_medium(128, "plume", 5e-3)
    gpu_set_states(gpu_ctx, new, new)
    H, F = gpu_ctx.render_twostates(seed=1, spp=16)
    # identical states give identical per-lane radiance; the two film accumulators receive their atomic adds in
    # (possibly) different orders, so the images agree to rounding, not bit for bit
    assert rel_l2(H.cpu().numpy(), F.cpu().numpy()) < 1e-6
    full = gpu_ctx.render_twostates_accum(seed=2, spp=16)
    part = gpu_ctx.render_twostates_accum(seed=2, spp=16, spp_begin=0, spp_end=5)
    part = gpu_ctx.render_twostates_accum(seed=2, spp=16, spp_begin=5, spp_end=16, out=part)
    assert rel_l2(part[1].cpu().numpy(), full[1].cpu().numpy()) < 1e-6
    w = full[0][..., 3].cpu().numpy()
    assert abs(w.sum() - 256 * 256 * 16) < 1 and w.min() >= 14
    g = to_cuda(np.full((256, 256, 3), 1.0 / (256 * 256 * 3), np.float32))
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(5, 16, g)
    gs, ga = gpu_ctx.grad_read()
    assert np.isfinite(gs.cpu().numpy()).all() and np.isfinite(ga.cpu().numpy()).all() and float(gs.abs().sum()) > 0
    c = gpu_ctx.counters(True)
    assert int(c[6]) > 0


def _ragged_medium(shape, kind):
    """(sigma_t (Z,Y,X,1), albedo (Z,Y,X,3)) for a non-cubic grid and degenerate contents."""
    rng = np.random.default_rng(sum(shape) + len(kind))
    nz, ny, nx = shape
    if kind == "empty":            # majorant 0: every free-flight sample leaves the medium
        s = np.zeros((nz, ny, nx, 1), np.float32)
    elif kind == "opaque_black":   # dense medium, zero albedo: throughput dies at the first real collision
        s = np.ones((nz, ny, nx, 1), np.float32)
    else:
        s = rng.random((nz, ny, nx, 1), dtype=np.float32)
        s[rng.random((nz, ny, nx, 1)) < 0.3] = 0.0            # holes
    a = rng.random((nz, ny, nx, 3), dtype=np.float32)
    if kind == "opaque_black":
        a[:] = 0.0
    if kind == "white":
        a[:] = 1.0
    return s, a


@pytest.mark.parametrize("shape,film,kind", [
    ((5, 9, 13), (37, 23), "holes"),        # ragged grid, odd film
    ((1, 1, 1), (8, 8), "holes"),           # a single voxel: every lookup clamps
    ((2, 3, 4), (1, 1), "white"),           # one pixel, albedo 1
    ((6, 6, 6), (16, 12), "empty"),         # majorant 0
    ((6, 6, 6), (16, 12), "opaque_black"),  # zero albedo
])
def test_ragged_and_degenerate_inputs(oracle, gpu_ctx, shape, film, kind):
    """Edge cases of the reference's tests: non-cubic and one-voxel grids, odd and one-pixel films,
    empty / black / white media, an empty sample shard.  Per-lane radiance stays bit-exact, gradients within 1e-3."""
    cfg = small_scene(width=film[0], height=film[1], max_depth=8, rr_depth=3)
    d = cfg.to_desc()
    old = _ragged_medium(shape, kind)
    new = (np.clip(old[0] * 1.01, 0, 1).astype(np.float32), np.clip(old[1] * 0.99, 0, 1).astype(np.float32))
    mo, mn = oracle.OracleMedium(d, *old), oracle.OracleMedium(d, *new)
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, old, new)
    assert gpu_ctx.majorants() == (mo.majorant, mn.majorant)
    ao, an = oracle.render_twostates_accum(d, mo, mn, seed=5, spp=1)
    go, gn = gpu_ctx.render_twostates_accum(seed=5, spp=1)
    assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
    # an empty shard renders nothing and leaves the accumulators untouched
    e = gpu_ctx.render_twostates_accum(seed=5, spp=4, spp_begin=2, spp_end=2)
    assert float(e[0].abs().sum()) == 0.0 and float(e[1].abs().sum()) == 0.0
    assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=6, spp=1).cpu().numpy(), oracle.vps_render_accum(d, mn, 6, 1))
    gimg = np.full((d.film_height, d.film_width, 3), 1e-2, np.float32)
    gs, ga = oracle.vps_backward(d, mn, 7, 2, gimg, threads=1)
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(7, 2, to_cuda(gimg))
    ggs, gga = gpu_ctx.grad_read()
    assert np.isfinite(ggs.cpu().numpy()).all() and np.isfinite(gga.cpu().numpy()).all()
    assert rel_l2(ggs.cpu().numpy(), gs) < GRAD_TOL and rel_l2(gga.cpu().numpy(), ga) < GRAD_TOL


def test_camera_inside_medium_and_no_emitter_sampling(oracle, gpu_ctx):
    """Camera inside the box (paths start in the medium's bounding box) and sample_emitters = False
    (phase sampling only, emitter hits counted directly: twostatevolpath.py:315-325)."""
    from recursive_control_variates_b200.scene import SceneConfig, look_at
    for kw in (dict(to_world=look_at([0.5, 0.5, 0.45], [0.5, 0.5, 1.5], [0, 1, 0])), dict(sample_emitters=False)):
        cfg = SceneConfig(width=24, height=20, max_depth=12, rr_depth=4, **kw)
        d = cfg.to_desc()
        old, new = media(12)
        mo, mn = oracle.OracleMedium(d, *old), oracle.OracleMedium(d, *new)
        gpu_ctx.set_scene(d)
        gpu_set_states(gpu_ctx, old, new)
        ao, an = oracle.render_twostates_accum(d, mo, mn, seed=31, spp=1)
        go, gn = gpu_ctx.render_twostates_accum(seed=31, spp=1)
        assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
        assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=32, spp=1).cpu().numpy(),
                              oracle.vps_render_accum(d, mn, 32, 1))


@pytest.mark.parametrize("tracer", [0, 1, 2, 3])
@pytest.mark.parametrize("factor", [2, 4])
def test_majorant_supergrid_dda(oracle, gpu_ctx, tracer, factor):
    """Majorant supergrid + DDA free-flight sampling (plugins/twostatemedium.py:25-70):
    the supergrid itself, per-lane radiance of the two-state and the single-state tracer (bit-exact) and the
    adjoint's gradients, for every schedule."""
    cfg = small_scene(width=40, height=32, max_depth=16, rr_depth=5)
    d = cfg.to_desc()
    old, new = media(16, "plume")
    mo = oracle.OracleMedium(d, *old, majorant_factor=factor)
    mn = oracle.OracleMedium(d, *new, majorant_factor=factor)
    gpu_ctx.set_tracer(tracer)
    gpu_ctx.set_scene(d)
    gpu_ctx.set_majorant_factor(factor)
    try:
        gpu_set_states(gpu_ctx, old, new)
        g_old, g_new = gpu_ctx.majorant_grid()
        assert np.array_equal(g_old, mo.majorant_grid) and np.array_equal(g_new, mn.majorant_grid)
        oracle.counters(True); gpu_ctx.counters(True)
        ao, an = oracle.render_twostates_accum(d, mo, mn, seed=41, spp=1)
        go, gn = gpu_ctx.render_twostates_accum(seed=41, spp=1)
        oc, gc = oracle.counters(True), gpu_ctx.counters(True)
        assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
        assert int(oc[0]) == int(gc[0]) and int(oc[1]) == int(gc[1])
        assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=42, spp=1).cpu().numpy(),
                              oracle.vps_render_accum(d, mn, 42, 1))
        gimg = np.full((d.film_height, d.film_width, 3), 1e-3, np.float32)
        gs, ga = oracle.vps_backward(d, mn, 43, 3, gimg, threads=1)
        gpu_ctx.grad_zero(); gpu_ctx.render_backward(43, 3, to_cuda(gimg))
        ggs, gga = gpu_ctx.grad_read()
        assert rel_l2(ggs.cpu().numpy(), gs) < GRAD_TOL and rel_l2(gga.cpu().numpy(), ga) < GRAD_TOL
        # fewer tentative collisions than with one majorant for the volume, same image in expectation
        m0o, m0n = oracle.OracleMedium(d, *old), oracle.OracleMedium(d, *new)
        H1, F1 = gpu_ctx.render_twostates(seed=44, spp=64)
        c1 = gpu_ctx.counters(True)
        gpu_ctx.set_majorant_factor(0)
        assert gpu_ctx.majorant_grid() is None
        H0, F0 = gpu_ctx.render_twostates(seed=44, spp=64)
        c0 = gpu_ctx.counters(True)
        assert int(c1[0]) < int(c0[0])
        assert abs(float(F1.mean()) - float(F0.mean())) < 5e-3 * float(F0.mean())
    finally:
        gpu_ctx.set_majorant_factor(0)
        gpu_ctx.set_tracer(3)


def test_majorant_factor_errors(gpu_ctx):
    from recursive_control_variates_b200._abi import RcvError
    d = small_scene(width=8, height=8).to_desc()
    gpu_ctx.set_scene(d)
    old, new = media(12)
    gpu_set_states(gpu_ctx, old, new)
    gpu_ctx.set_majorant_factor(5)               # 12 % 5 != 0: accepted (a rotation may follow) but nothing renders
    assert gpu_ctx.majorant_grid() is None
    with pytest.raises(RcvError):
        gpu_ctx.render_twostates_accum(seed=1, spp=1)
    with pytest.raises(RcvError):
        gpu_ctx.set_majorant_factor(-1)
    gpu_ctx.set_majorant_factor(1)               # <= 1 disables (utils.py:36-37)
    assert gpu_ctx.majorant_grid() is None
    gpu_ctx.set_majorant_factor(4)
    o, n = gpu_ctx.majorant_grid()
    assert o.shape == (3, 3, 3)
    with pytest.raises(RcvError):                # a rotation to a grid the factor does not divide is refused
        gpu_set_states(gpu_ctx, media(10)[0], media(10)[1])
    gpu_ctx.set_majorant_factor(0)


@pytest.mark.parametrize("tracer", [1, 3])
def test_batch_sensor(oracle, gpu_ctx, tracer):
    """`batch` sensor (experiments.py:28-50): views side by side on one film, per-lane bit-exact against the oracle
    (two-state render, single-state primal, adjoint)."""
    import math
    from recursive_control_variates_b200.scene import SceneConfig, look_at
    from recursive_control_variates_b200._abi import RcvError
    views = [look_at([0.5 + 2.5 * math.cos(t), 0.6, 0.5 + 2.5 * math.sin(t)], [0.5, 0.5, 0.5], [0, 1, 0]) for t in (0.1, 1.7, 3.9, 5.0)]
    cfg = SceneConfig(width=4 * 20, height=18, to_world=np.stack(views), max_depth=8, rr_depth=4)
    d = cfg.to_desc()
    old, new = media(12)
    mo, mn = oracle.OracleMedium(d, *old), oracle.OracleMedium(d, *new)
    gpu_ctx.set_tracer(tracer)
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, old, new)
    ao, an = oracle.render_twostates_accum(d, mo, mn, seed=51, spp=1)
    go, gn = gpu_ctx.render_twostates_accum(seed=51, spp=1)
    assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
    # the four sub-films are different views: their images differ
    sub = [an[:, 20 * k:20 * (k + 1), :3] for k in range(4)]
    assert not np.array_equal(sub[0], sub[1]) and not np.array_equal(sub[2], sub[3])
    assert np.array_equal(gpu_ctx.render_adjoint_primal_accum(seed=52, spp=1).cpu().numpy(), oracle.vps_render_accum(d, mn, 52, 1))
    gimg = np.full((d.film_height, d.film_width, 3), 1e-3, np.float32)
    gs, ga = oracle.vps_backward(d, mn, 53, 2, gimg, threads=1)
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(53, 2, to_cuda(gimg))
    ggs, gga = gpu_ctx.grad_read()
    assert rel_l2(ggs.cpu().numpy(), gs) < GRAD_TOL and rel_l2(gga.cpu().numpy(), ga) < GRAD_TOL
    bad = cfg.to_desc(); bad.film_width = 81                  # not a multiple of the number of views
    with pytest.raises(RcvError):
        gpu_ctx.set_scene(bad)
    gpu_ctx.set_tracer(3)


def test_sharded_film_weights(gpu_ctx, setup):
    """rcv_film_weights / rcv_render_backward_weighted: the weights of disjoint sample ranges add up to the weights of
    all samples, and the adjoint with supplied weights equals the adjoint that computes them itself."""
    import torch
    d = setup["desc"]
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, setup["old"], setup["new"])
    seed, spp = 77, 6
    w_all = gpu_ctx.film_weights(seed, spp)
    w_parts = gpu_ctx.film_weights(seed, spp, 0, 2)
    gpu_ctx.film_weights(seed, spp, 2, 6, out=w_parts)
    assert torch.equal(w_all, w_parts) and abs(float(w_all.sum()) - d.film_width * d.film_height * spp) < 1
    gimg = to_cuda(np.full((d.film_height, d.film_width, 3), 1e-3, np.float32))
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(seed, spp, gimg, 1, 4)
    a_s, a_a = gpu_ctx.grad_read()
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(seed, spp, gimg, 1, 4, film_w=w_all)
    b_s, b_a = gpu_ctx.grad_read()
    assert rel_l2(b_s.cpu().numpy(), a_s.cpu().numpy()) < 1e-5 and rel_l2(b_a.cpu().numpy(), a_a.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("tracer", [0, 1, 3])
def test_gaussian_reconstruction_filter(oracle, gpu_ctx, tracer):
    """hdrfilm's default gaussian filter: every sample is splatted to up to 4 x 4 pixels, so even at spp = 1 a pixel
    sums several contributions in an order that differs between the oracle and the atomics: images to 1e-6."""
    cfg = small_scene(width=44, height=36, max_depth=12, rr_depth=4, rfilter="gaussian")
    d = cfg.to_desc()
    old, new = media(14)
    mo, mn = oracle.OracleMedium(d, *old), oracle.OracleMedium(d, *new)
    gpu_ctx.set_tracer(tracer)
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, old, new)
    ao, an = oracle.render_twostates_accum(d, mo, mn, seed=61, spp=3)
    go, gn = gpu_ctx.render_twostates_accum(seed=61, spp=3)
    assert rel_l2(go.cpu().numpy(), ao) < 1e-6 and rel_l2(gn.cpu().numpy(), an) < 1e-6
    H, F = oracle.render_twostates(d, mo, mn, seed=61, spp=3)
    gH, gF = gpu_ctx.render_twostates(seed=61, spp=3)
    assert rel_l2(gH.cpu().numpy(), H) < IMG_TOL and rel_l2(gF.cpu().numpy(), F) < IMG_TOL
    assert rel_l2(gpu_ctx.render_adjoint_primal_accum(seed=62, spp=2).cpu().numpy(), oracle.vps_render_accum(d, mn, 62, 2)) < 1e-6
    fw = gpu_ctx.film_weights(63, 3).cpu().numpy()
    assert rel_l2(fw, oracle.film_weights(d, 63, 3)) < 1e-6
    rng = np.random.default_rng(7)
    gimg = (rng.standard_normal((d.film_height, d.film_width, 3)) / (d.film_height * d.film_width * 3)).astype(np.float32)
    gs, ga = oracle.vps_backward(d, mn, 63, 3, gimg, threads=1)
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(63, 3, to_cuda(gimg))
    ggs, gga = gpu_ctx.grad_read()
    assert rel_l2(ggs.cpu().numpy(), gs) < GRAD_TOL and rel_l2(gga.cpu().numpy(), ga) < GRAD_TOL
    from recursive_control_variates_b200._abi import RcvError
    bad = cfg.to_desc(); bad.rfilter_stddev = 0.8
    with pytest.raises(RcvError):
        gpu_ctx.set_scene(bad)
    gpu_ctx.set_tracer(3)


@pytest.mark.parametrize("tracer", [0, 3])
def test_long_shadow_rays_fall_back_to_the_replay(oracle, gpu_ctx, tracer):
    """Dense medium (scale 150): shadow rays with more than 64 tentative collisions overflow their record buffer and are
    replayed in pass 2 instead (volpathsimple.py:406-414); gradients still within tolerance."""
    d = small_scene(width=20, height=16, max_depth=6, rr_depth=3, sigma_t_scale=150.0).to_desc()
    old, new = media(10, "dense")
    mn = oracle.OracleMedium(d, *new)
    gpu_ctx.set_tracer(tracer)
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, old, new)
    gimg = np.full((d.film_height, d.film_width, 3), 1e-3, np.float32)
    gs, ga = oracle.vps_backward(d, mn, 71, 2, gimg, threads=1)
    gpu_ctx.grad_zero(); gpu_ctx.render_backward(71, 2, to_cuda(gimg))
    ggs, gga = gpu_ctx.grad_read()
    assert rel_l2(ggs.cpu().numpy(), gs) < GRAD_TOL and rel_l2(gga.cpu().numpy(), ga) < GRAD_TOL
    gpu_ctx.set_tracer(3)


def test_out_of_memory_is_an_error_not_a_leak(oracle, gpu_ctx, setup):
    """A grid that cannot be allocated (1500^3: 0.3 TB of packed state) is refused with a CUDA error; nothing leaks, the
    context keeps its medium and renders as before."""
    import torch
    from recursive_control_variates_b200._abi import RcvError, check
    d = setup["desc"]
    gpu_ctx.set_scene(d)
    gpu_set_states(gpu_ctx, setup["old"], setup["new"])
    free_before = torch.cuda.mem_get_info()[0]
    dummy = torch.zeros(16, device="cuda")
    with pytest.raises(RcvError):
        check(gpu_ctx.lib.rcv_medium_rotate(gpu_ctx._h, dummy.data_ptr(), dummy.data_ptr(), 1500, 1500, 1500))
    with pytest.raises(RcvError):           # 2^32 or more interpolation cells: refused before any allocation
        check(gpu_ctx.lib.rcv_medium_rotate(gpu_ctx._h, dummy.data_ptr(), dummy.data_ptr(), 4096, 4096, 4096))
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] >= free_before - (64 << 20)
    ao, an = oracle.render_twostates_accum(d, setup["mo"], setup["mn"], seed=3, spp=1)
    go, gn = gpu_ctx.render_twostates_accum(seed=3, spp=1)
    assert np.array_equal(go.cpu().numpy(), ao) and np.array_equal(gn.cpu().numpy(), an)
