"""CPU check of the CUDA lane event machines (csrc/rcv_*_wf.cuh) compiled for the
host (tests/host_emu): every lane is single-stepped through its EV_INIT /
EV_COLL / EV_SCAT handlers and must reproduce the oracle bit for bit (spp = 1
accumulators, event counters) — the same comparison the GPU tests make, but
runnable without a GPU.  The harness is test infrastructure, not a CPU path of
the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import media, rel_l2, small_scene

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "host_emu")


def _load_emu(name):
    subprocess.run(["make", "-j4", "-C", EMU_DIR, "all"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    lib = C.CDLL(os.path.join(EMU_DIR, name))
    f32 = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
    u64 = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
    lib.emu_twostates.argtypes = [C.c_void_p, f32, f32, f32, f32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                  f32, f32, u64]
    lib.emu_vps.argtypes = [C.c_void_p, f32, f32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32,
                            C.c_void_p, C.c_void_p, f32, f32, f32, u64]
    lib.emu_twostates_sg.argtypes = [C.c_void_p, f32, f32, f32, f32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int,
                                     C.c_int, f32, f32, u64]
    lib.emu_vps_sg.argtypes = [C.c_void_p, f32, f32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32,
                               C.c_void_p, C.c_void_p, f32, f32, f32, u64, C.c_int]
    lib.emu_vps_backward_lane.argtypes = [C.c_void_p, f32, f32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.c_void_p, C.c_void_p, f32, f32, u64, C.c_int, C.c_int]
    lib.emu_set_flat.argtypes = [C.c_int]
    lib.emu_concat_owner.argtypes = [np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS"), C.c_int]
    lib.emu_concat_owner.restype = C.c_int
    lib.emu_vps_drt_stages.argtypes = [C.c_void_p, f32, f32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_void_p, C.c_void_p, f32, f32, u64, C.c_int, C.c_int]
    return lib


@pytest.fixture(scope="module")
def emu():
    return _load_emu("librcv_emu.so")


@pytest.fixture(scope="module")
def emu_voxelreds():
    """The same machines built with RCV_GRAD_BRICKS=0 (d sigma_t reductions straight into the voxel grid)."""
    return _load_emu("librcv_emu_voxelreds.so")


@pytest.fixture(scope="module")
def setup(oracle):
    cfg = small_scene(40, 36)
    desc = cfg.to_desc()
    old, new = media(20)
    return dict(desc=desc, old=old, new=new, mo=oracle.OracleMedium(desc, *old), mn=oracle.OracleMedium(desc, *new))


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_twostate_event_machine_matches_oracle(oracle, emu, setup, mode):
    d = setup["desc"]
    (so, ao), (sn, an) = setup["old"], setup["new"]
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_twostates(C.addressof(d), so, ao, sn, an, *so.shape[:3], 21, 1, mode, acc_o, acc_n, cnt)
    oracle.counters(True)
    mo, mn = setup["mo"], setup["mn"]
    if mode == 1:
        mn = mo
    elif mode == 2:
        mo = mn
    ref_o, ref_n = oracle.render_twostates_accum(d, mo, mn, 21, 1)
    oc = oracle.counters(True)
    if mode != 2:
        assert np.array_equal(acc_o, ref_o)
    if mode != 1:
        assert np.array_equal(acc_n, ref_n)
    assert int(cnt[0]) == int(oc[0]) and int(cnt[1]) == int(oc[1])
    assert int(cnt[6]) == H * W


def test_twostate_event_machine_multi_spp(oracle, emu, setup):
    d = setup["desc"]
    (so, ao), (sn, an) = setup["old"], setup["new"]
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_twostates(C.addressof(d), so, ao, sn, an, *so.shape[:3], 5, 4, 0, acc_o, acc_n, cnt)
    ref_o, ref_n = oracle.render_twostates_accum(d, setup["mo"], setup["mn"], 5, 4)
    # the harness and the oracle both splat in lane order: identical sums
    assert np.array_equal(acc_o, ref_o) and np.array_equal(acc_n, ref_n)


def test_volpathsimple_event_machine_primal(oracle, emu, setup):
    d = setup["desc"]
    sn, an = setup["new"]
    H, W = d.film_height, d.film_width
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 33, 1, 0, 0, None, None, acc, g1, g4, cnt)
    oracle.counters(True)
    ref = oracle.vps_render_accum(d, setup["mn"], 33, 1)
    oc = oracle.counters(True)
    assert np.array_equal(acc, ref)
    assert [int(v) for v in cnt[2:4]] == [int(v) for v in oc[2:4]]


def test_volpathsimple_event_machine_backward(oracle, emu, setup):
    d = setup["desc"]
    sn, an = setup["new"]
    H, W = d.film_height, d.film_width
    rng = np.random.default_rng(3)
    grad_img = (rng.standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    seed_grad, spp = 0xC09848F2, 2
    film_w = oracle.film_weights(d, seed_grad, spp)
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], seed_grad, spp, 1, oracle.alt_seed(seed_grad),
                grad_img.ctypes.data, film_w.ctypes.data, acc, g1, g4, cnt)
    oracle.counters(True)
    gs, ga = oracle.vps_backward(d, setup["mn"], seed_grad, spp, grad_img, threads=1, film_w=film_w)
    oc = oracle.counters(True)
    # vertex-list adjoint: same vertices, same gradient sites, same number of passes — but pass 2 does not
    # repeat the null collisions of the main path nor the primal shadow rays
    assert [int(v) for v in cnt[3:7]] == [int(v) for v in oc[3:7]]
    assert int(cnt[2]) < 0.7 * int(oc[2])
    assert rel_l2(g1, gs) < 1e-5       # float vs double accumulation
    assert rel_l2(g4[..., 0:3], ga) < 1e-5


def test_hg_phase_and_shallow_depth(oracle, emu):
    """Henyey-Greenstein phase function (experiments.py:179-182, g = 0.877) and a small
    max_depth / rr_depth (Russian roulette active): event machines vs oracle, bit for bit."""
    from recursive_control_variates_b200.scene import SceneConfig
    d = SceneConfig(width=24, height=20, phase="hg", phase_g=0.877, max_depth=6, rr_depth=2).to_desc()
    (so, ao), (sn, an) = media(12)
    mo, mn = oracle.OracleMedium(d, so, ao), oracle.OracleMedium(d, sn, an)
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_twostates(C.addressof(d), so, ao, sn, an, *so.shape[:3], 4, 2, 0, acc_o, acc_n, cnt)
    ref_o, ref_n = oracle.render_twostates_accum(d, mo, mn, 4, 2)
    assert np.array_equal(acc_o, ref_o) and np.array_equal(acc_n, ref_n)
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 6, 2, 0, 0, None, None, acc, g1, g4, cnt)
    assert np.array_equal(acc, oracle.vps_render_accum(d, mn, 6, 2))
    gimg = np.full((H, W, 3), 1e-3, np.float32)
    film_w = oracle.film_weights(d, 9, 2)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 9, 2, 1, oracle.alt_seed(9), gimg.ctypes.data, film_w.ctypes.data, acc, g1, g4, cnt)
    gs, ga = oracle.vps_backward(d, mn, 9, 2, gimg, threads=1, film_w=film_w)
    assert rel_l2(g1, gs) < 1e-5 and rel_l2(g4[..., 0:3], ga) < 1e-5


@pytest.mark.parametrize("factor", [2, 4])
def test_majorant_supergrid_event_machines(oracle, emu, factor):
    """Majorant supergrid + DDA free-flight sampling (plugins/twostatemedium.py:25-70) in the GridKS instantiation of
    the event machines: two-state render, single-state primal and the adjoint against the oracle."""
    d = small_scene(32, 28, max_depth=12, rr_depth=4).to_desc()
    (so, ao), (sn, an) = media(16, "plume")
    mo, mn = oracle.OracleMedium(d, so, ao, majorant_factor=factor), oracle.OracleMedium(d, sn, an, majorant_factor=factor)
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_twostates_sg(C.addressof(d), so, ao, sn, an, *so.shape[:3], 13, 2, 0, factor, acc_o, acc_n, cnt)
    oracle.counters(True)
    ref_o, ref_n = oracle.render_twostates_accum(d, mo, mn, 13, 2)
    oc = oracle.counters(True)
    assert np.array_equal(acc_o, ref_o) and np.array_equal(acc_n, ref_n)
    assert int(cnt[0]) == int(oc[0]) and int(cnt[1]) == int(oc[1])
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    emu.emu_vps_sg(C.addressof(d), sn, an, *sn.shape[:3], 14, 1, 0, 0, None, None, acc, g1, g4, cnt, factor)
    assert np.array_equal(acc, oracle.vps_render_accum(d, mn, 14, 1))
    gimg = np.full((H, W, 3), 1e-3, np.float32)
    film_w = oracle.film_weights(d, 15, 2)
    acc[:] = 0
    emu.emu_vps_sg(C.addressof(d), sn, an, *sn.shape[:3], 15, 2, 1, oracle.alt_seed(15), gimg.ctypes.data, film_w.ctypes.data,
                   acc, g1, g4, cnt, factor)
    gs, ga = oracle.vps_backward(d, mn, 15, 2, gimg, threads=1, film_w=film_w)
    assert rel_l2(g1, gs) < 1e-5 and rel_l2(g4[..., 0:3], ga) < 1e-5


def test_gaussian_filter_event_machines(oracle, emu):
    """hdrfilm's default gaussian reconstruction filter: footprint splats of the two-state render and of the
    single-state primal, and the adjoint's dL gathered over the footprint (sums differ only in their order)."""
    d = small_scene(28, 24, max_depth=10, rr_depth=4, rfilter="gaussian").to_desc()
    (so, ao), (sn, an) = media(12)
    mo, mn = oracle.OracleMedium(d, so, ao), oracle.OracleMedium(d, sn, an)
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_twostates(C.addressof(d), so, ao, sn, an, *so.shape[:3], 31, 2, 0, acc_o, acc_n, cnt)
    ref_o, ref_n = oracle.render_twostates_accum(d, mo, mn, 31, 2)
    assert rel_l2(acc_o, ref_o) < 1e-6 and rel_l2(acc_n, ref_n) < 1e-6
    assert abs(float(acc_n[..., 3].sum()) - float(ref_n[..., 3].sum())) < 1e-3 * float(ref_n[..., 3].sum())
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 32, 2, 0, 0, None, None, acc, g1, g4, cnt)
    assert rel_l2(acc, oracle.vps_render_accum(d, mn, 32, 2)) < 1e-6
    rng = np.random.default_rng(5)
    gimg = (rng.standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    film_w = oracle.film_weights(d, 33, 2)
    acc[:] = 0
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 33, 2, 1, oracle.alt_seed(33), gimg.ctypes.data, film_w.ctypes.data, acc, g1, g4, cnt)
    gs, ga = oracle.vps_backward(d, mn, 33, 2, gimg, threads=1, film_w=film_w)
    assert rel_l2(g1, gs) < 1e-5 and rel_l2(g4[..., 0:3], ga) < 1e-5


def test_long_shadow_rays_fall_back_to_the_replay(oracle, emu):
    """A dense medium (scale 150): shadow rays with more tentative collisions than the 64 records a slot holds take the
    fall-back — the replay of the ray in pass 2 with the parked sampler (volpathsimple.py:406-414) — and the gradients
    still match the oracle."""
    d = small_scene(16, 12, max_depth=6, rr_depth=3, sigma_t_scale=150.0).to_desc()
    (so, ao), (sn, an) = media(10, "dense")
    mn = oracle.OracleMedium(d, sn, an)
    H, W = d.film_height, d.film_width
    gimg = np.full((H, W, 3), 1e-3, np.float32)
    film_w = oracle.film_weights(d, 71, 2)
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 71, 2, 1, oracle.alt_seed(71), gimg.ctypes.data, film_w.ctypes.data, acc, g1, g4, cnt)
    oracle.counters(True)
    gs, ga = oracle.vps_backward(d, mn, 71, 2, gimg, threads=1, film_w=film_w)
    oc = oracle.counters(True)
    assert [int(v) for v in cnt[3:7]] == [int(v) for v in oc[3:7]]
    assert int(oc[2]) / max(1, int(oc[3])) > 40        # dense enough: tens of tentative collisions per vertex
    assert rel_l2(g1, gs) < 1e-5 and rel_l2(g4[..., 0:3], ga) < 1e-5


@pytest.mark.parametrize("flags,factor", [(0, 0), (7, 0), (5, 0), (3, 0), (1, 0), (7, 5)])
def test_per_lane_adjoint_with_drt(oracle, emu, flags, factor):
    """vps_backward_lane (csrc/rcv_volpathsimple.cuh), the one-thread-per-lane adjoint behind use_drt = True
    (volpathsimple.py:520-594, :623-778): flags 0 is the path-replay estimator, else bit 0 use_drt, bit 1
    use_drt_subsampling, bit 2 use_drt_mis.  Same lanes, same draws, same gradient sites as the oracle."""
    from recursive_control_variates_b200.scene import SceneConfig
    d = SceneConfig(width=20, height=16, phase="hg", phase_g=0.5, max_depth=12, rr_depth=4).to_desc()
    _, (sn, an) = media(10)
    mn = oracle.OracleMedium(d, sn, an, majorant_factor=factor)
    H, W = d.film_height, d.film_width
    grad_img = (np.random.default_rng(5).standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    seed_grad, spp = 77, 3
    film_w = oracle.film_weights(d, seed_grad, spp)
    g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_vps_backward_lane(C.addressof(d), sn, an, *sn.shape[:3], seed_grad, spp, oracle.alt_seed(seed_grad),
                              grad_img.ctypes.data, film_w.ctypes.data, g1, g4, cnt, factor, flags)
    oracle.counters(True)
    gs, ga = oracle.vps_backward(d, mn, seed_grad, spp, grad_img, threads=1, film_w=film_w, use_drt=bool(flags & 1),
                                 use_drt_subsampling=bool(flags & 2), use_drt_mis=bool(flags & 4))
    oc = oracle.counters(True)
    assert [int(v) for v in cnt[2:7]] == [int(v) for v in oc[2:7]]
    assert rel_l2(g1, gs) < 1e-5       # float vs double accumulation
    assert rel_l2(g4[..., 0:3], ga) < 1e-5


@pytest.mark.parametrize("use_mis,factor,depths", [(1, 0, (12, 4)), (0, 0, (12, 4)), (1, 5, (12, 4)), (1, 0, (3, 1)), (1, 0, (64, 64))])
def test_drt_stages_of_the_event_machines(oracle, emu, use_mis, factor, depths):
    """use_drt + use_drt_subsampling on the event machines: stage 1 (adjoint, reservoir kept during pass 1, job records),
    the job step (distance sampling, shadow ray, phase sample) and stage 2 (the primal machine resumed from the job's
    PathState, its result scattered into the gradient grids) against the oracle's straight-line restatement: the
    secondary sampler's draws are reordered across the stages but each is the same number."""
    from recursive_control_variates_b200.scene import SceneConfig
    d = SceneConfig(width=20, height=16, phase="hg", phase_g=0.5, max_depth=depths[0], rr_depth=depths[1]).to_desc()
    _, (sn, an) = media(10)
    mn = oracle.OracleMedium(d, sn, an, majorant_factor=factor)
    H, W = d.film_height, d.film_width
    grad_img = (np.random.default_rng(6).standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    seed_grad, spp = 78, 3
    film_w = oracle.film_weights(d, seed_grad, spp)
    g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    cnt = np.zeros(8, np.uint64)
    emu.emu_vps_drt_stages(C.addressof(d), sn, an, *sn.shape[:3], seed_grad, spp, oracle.alt_seed(seed_grad),
                           grad_img.ctypes.data, film_w.ctypes.data, g1, g4, cnt, factor, use_mis)
    oracle.counters(True)
    gs, ga = oracle.vps_backward(d, mn, seed_grad, spp, grad_img, threads=1, film_w=film_w, use_drt=True, use_drt_mis=bool(use_mis))
    oc = oracle.counters(True)
    assert [int(v) for v in cnt[3:7]] == [int(v) for v in oc[3:7]]
    assert int(cnt[2]) < int(oc[2])          # vertex-list adjoint: the main path's null collisions are not repeated
    assert rel_l2(g1, gs) < 1e-5             # float vs double accumulation
    assert rel_l2(g4[..., 0:3], ga) < 1e-5


@pytest.fixture(scope="module")
def emu_vertex_albedo():
    """The same machines built with RCV_VERTEX_ALBEDO=0 (pass 2 gathers the albedo again; the default build reads it
    from the vertex record)."""
    return _load_emu("librcv_emu_vertex_albedo.so")


def test_vertex_record_albedo_is_the_same_adjoint(oracle, emu, emu_vertex_albedo):
    """RCV_VERTEX_ALBEDO: the albedo pass 2 needs at a scattering vertex is the one pass 1 gathered; carried in the
    vertex record it gives bit-identical gradients — also when shadow rays overflow their record buffer and the replay
    takes the record's words back (dense medium), and for the DRT stages."""
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import make_medium
    for kind, scale, depth in (("cube", 20.0, 12), ("dense", 160.0, 6)):
        d = SceneConfig(width=20, height=16, max_depth=depth, rr_depth=4, sigma_t_scale=scale).to_desc()
        sn, an = make_medium(10, kind, 5e-3)
        H, W = d.film_height, d.film_width
        grad_img = (np.random.default_rng(9).standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
        film_w = oracle.film_weights(d, 31, 2)
        res = []
        for lib in (emu, emu_vertex_albedo):
            acc = np.zeros((H, W, 4), np.float32)
            g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32); cnt = np.zeros(8, np.uint64)
            lib.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 31, 2, 1, oracle.alt_seed(31), grad_img.ctypes.data,
                        film_w.ctypes.data, acc, g1, g4, cnt)
            d1 = np.zeros_like(g1); d4 = np.zeros_like(g4); dcnt = np.zeros(8, np.uint64)
            lib.emu_vps_drt_stages(C.addressof(d), sn, an, *sn.shape[:3], 31, 2, oracle.alt_seed(31), grad_img.ctypes.data,
                                   film_w.ctypes.data, d1, d4, dcnt, 0, 1)
            res.append((g1, g4, cnt, d1, d4, dcnt))
        for a, b in zip(*res):
            assert np.array_equal(a, b)
        assert res[0][0].any() and res[0][3].any()


@pytest.fixture(scope="module")
def emu_spec():
    """The same machines with the other setting of RCV_SPEC_COLLIDE than the default build's."""
    return _load_emu("librcv_emu_spec.so")


def _every_output(lib, oracle, factor):
    """Two-state accumulators (modes 0 and 2), adjoint gradients and DRT-stage gradients with all counters, of one build
    of the machines on a fixed small scene (Russian roulette active, dense enough for long shadow rays)."""
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import make_medium
    d = SceneConfig(width=24, height=20, max_depth=16, rr_depth=2, sigma_t_scale=30.0).to_desc()
    so, ao = make_medium(12, "cube", 0.0)
    sn, an = make_medium(12, "cube", 8e-3)
    H, W = d.film_height, d.film_width
    grad_img = (np.random.default_rng(3).standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    film_w = oracle.film_weights(d, 77, 3)
    out = []
    for mode in (0, 2):
        a_old = np.zeros((H, W, 4), np.float32); a_new = np.zeros_like(a_old); cnt = np.zeros(8, np.uint64)
        if factor:
            lib.emu_twostates_sg(C.addressof(d), so, ao, sn, an, *sn.shape[:3], 5, 3, mode, factor, a_old, a_new, cnt)
        else:
            lib.emu_twostates(C.addressof(d), so, ao, sn, an, *sn.shape[:3], 5, 3, mode, a_old, a_new, cnt)
        out += [a_old, a_new, cnt]
    acc = np.zeros((H, W, 4), np.float32)
    g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32); cnt = np.zeros(8, np.uint64)
    args = (C.addressof(d), sn, an, *sn.shape[:3], 77, 3, 1, oracle.alt_seed(77), grad_img.ctypes.data, film_w.ctypes.data,
            acc, g1, g4, cnt)
    if factor:
        lib.emu_vps_sg(*args, factor)
    else:
        lib.emu_vps(*args)
    out += [g1, g4, cnt]
    d1 = np.zeros_like(g1); d4 = np.zeros_like(g4); dcnt = np.zeros(8, np.uint64)
    lib.emu_vps_drt_stages(C.addressof(d), sn, an, *sn.shape[:3], 77, 3, oracle.alt_seed(77), grad_img.ctypes.data,
                           film_w.ctypes.data, d1, d4, dcnt, factor, 1)
    out += [d1, d4, dcnt]
    return out


@pytest.mark.parametrize("factor", [0, 4])
def test_speculative_collisions_change_nothing(oracle, emu, emu_spec, factor):
    """RCV_SPEC_COLLIDE: the second tentative collision of a call is sampled on a copy of the lane's stream as if the
    first were a null collision, and dropped otherwise.  Film accumulators, gradients and every counter are
    bit-identical with and without it — main path and shadow rays, Russian roulette active (rr_depth 2: no
    speculation on the two-state main path then), with and without the majorant supergrid."""
    res = [_every_output(lib, oracle, factor) for lib in (emu, emu_spec)]
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert res[0][0].any() and res[0][6].any() and res[0][9].any()


@pytest.fixture(scope="module")
def emu_walks():
    """The same machines with the non-default walks: RCV_DDA_LIN_TWOSTATE=0, RCV_DDA_LIN_VPS=1, RCV_DDA_AHEAD=3,
    RCV_DRT_AHEAD=3."""
    return _load_emu("librcv_emu_walks.so")


@pytest.mark.parametrize("factor", [2, 4, 0])
def test_walk_variants_change_nothing(oracle, emu, emu_walks, factor):
    """The supergrid walk exists in three spellings — the loop of the specification (dda_sample, K = 1), K cells at a
    time with their majorants in flight together (RCV_DDA_AHEAD), and with a linear cell index, steps-left counters and
    select-based stepping (dda_sample_lin) — and the DRT walk in two (RCV_DRT_AHEAD).  Which machine uses which is a
    build option; every spelling gives the same accumulators, gradients and counters bit for bit (factor 0: only the
    DRT walk differs)."""
    res = [_every_output(lib, oracle, factor) for lib in (emu, emu_walks)]
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert res[0][0].any() and res[0][6].any() and res[0][9].any()


def test_owner_search_of_the_warp_wide_record_scatter(emu):
    """vp_nee_records_warp (rcv_trace.cu) lays the shadow-ray record lists of a slow step end to end and gives every
    thread each 32nd record; concat_owner finds whose list a position falls into from the 32 exclusive prefix sums —
    the same function runs here over an array instead of shuffles.  Empty lists, one long list, all lists full."""
    rng = np.random.default_rng(11)
    cases = [rng.integers(0, 65, 32) * (rng.random(32) < p) for p in (0.1, 0.5, 1.0) for _ in range(20)]
    cases += [np.zeros(32, int), np.full(32, 64), np.eye(32, dtype=int)[0] * 64, np.eye(32, dtype=int)[31] * 7]
    for n in cases:
        n = np.asarray(n, np.int64)
        excl = np.ascontiguousarray(np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int32))
        total = int(n.sum())
        owner_of = np.repeat(np.arange(32), n)
        for j in range(total):
            assert emu.emu_concat_owner(excl, j) == owner_of[j]
        if total:
            assert emu.emu_concat_owner(excl, total + 5) == 31         # past the end: any valid thread (its record is not read)


def test_gradient_bricks_and_voxel_reductions_agree(oracle, emu, emu_voxelreds, setup):
    """RCV_GRAD_BRICKS: the d sigma_t contributions of a site go to the site's 32-byte gradient brick and are folded into
    the voxel grid after the launch (default), or straight into the voxel grid (0).  Same sites, same values: the two
    builds differ in float summation order only, for the path-replay adjoint and for the DRT stages."""
    d = setup["desc"]
    sn, an = setup["new"]
    H, W = d.film_height, d.film_width
    grad_img = (np.random.default_rng(8).standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    seed_grad, spp = 4242, 2
    film_w = oracle.film_weights(d, seed_grad, spp)
    out = {}
    for name, lib in (("bricks", emu), ("voxels", emu_voxelreds)):
        acc = np.zeros((H, W, 4), np.float32)
        g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32); cnt = np.zeros(8, np.uint64)
        lib.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], seed_grad, spp, 1, oracle.alt_seed(seed_grad),
                    grad_img.ctypes.data, film_w.ctypes.data, acc, g1, g4, cnt)
        d1 = np.zeros_like(g1); d4 = np.zeros_like(g4); dcnt = np.zeros(8, np.uint64)
        lib.emu_vps_drt_stages(C.addressof(d), sn, an, *sn.shape[:3], seed_grad, spp, oracle.alt_seed(seed_grad),
                               grad_img.ctypes.data, film_w.ctypes.data, d1, d4, dcnt, 0, 1)
        out[name] = (g1, g4, cnt, d1, d4, dcnt)
    b, v = out["bricks"], out["voxels"]
    assert np.array_equal(b[2], v[2]) and np.array_equal(b[5], v[5])            # same events
    assert np.array_equal(b[1], v[1]) and np.array_equal(b[4], v[4])            # the albedo gradient does not use bricks
    assert rel_l2(b[0], v[0]) < 1e-6 and rel_l2(b[3], v[3]) < 1e-6
    gs, _ = oracle.vps_backward(d, setup["mn"], seed_grad, spp, grad_img, threads=1, film_w=film_w)
    assert rel_l2(b[0], gs) < 1e-5 and rel_l2(v[0], gs) < 1e-5


@pytest.mark.parametrize("case", ["absorbing", "furnace"])
def test_limit_media_on_the_event_machines(oracle, emu, case):
    """The media of the analytic pins (tests/test_oracle.py): albedo 0 (the throughput dies at the first real collision)
    and albedo 1 with max_depth = 256 — lane for lane against the oracle, both tracers and the adjoint."""
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import make_medium
    if case == "absorbing":
        d = SceneConfig(width=16, height=16, sigma_t_scale=4.0, env_radiance=(0.8, 1.0, 1.2)).to_desc()
        sn = np.full((6, 6, 6, 1), 0.5, np.float32); an = np.zeros((6, 6, 6, 3), np.float32)
    else:
        d = SceneConfig(width=12, height=12, sigma_t_scale=6.0, env_radiance=(1, 1, 1), max_depth=256, rr_depth=256,
                        phase="hg", phase_g=0.6).to_desc()
        sn, _ = make_medium(8, "dense", 0.0); an = np.ones((8, 8, 8, 3), np.float32)
    m = oracle.OracleMedium(d, sn, an)
    H, W = d.film_height, d.film_width
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32); cnt = np.zeros(8, np.uint64)
    emu.emu_twostates(C.addressof(d), sn, an, sn, an, *sn.shape[:3], 21, 1, 0, acc_o, acc_n, cnt)
    ref_o, ref_n = oracle.render_twostates_accum(d, m, m, 21, 1)
    assert np.array_equal(acc_o, ref_o) and np.array_equal(acc_n, ref_n)
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 22, 1, 0, 0, None, None, acc, g1, g4, cnt)
    assert np.array_equal(acc, oracle.vps_render_accum(d, m, 22, 1))
    gimg = np.full((H, W, 3), 1.0 / (H * W * 3), np.float32)
    film_w = oracle.film_weights(d, 23, 2)
    emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], 23, 2, 1, oracle.alt_seed(23), gimg.ctypes.data, film_w.ctypes.data, acc, g1, g4, cnt)
    gs, ga = oracle.vps_backward(d, m, 23, 2, gimg, threads=1, film_w=film_w)
    assert rel_l2(g1, gs) < 1e-5
    assert (not ga.any() and not g4.any()) if case == "absorbing" else rel_l2(g4[..., 0:3], ga) < 1e-5
    d1 = np.zeros_like(g1); d4 = np.zeros_like(g4)
    emu.emu_vps_drt_stages(C.addressof(d), sn, an, *sn.shape[:3], 23, 2, oracle.alt_seed(23), gimg.ctypes.data, film_w.ctypes.data,
                           d1, d4, cnt, 0, 1)
    ds, da = oracle.vps_backward(d, m, 23, 2, gimg, threads=1, film_w=film_w, use_drt=True)
    assert rel_l2(d1, ds) < 1e-5
    # (with albedo 0 the path-replay term has no albedo gradient, its throughput being 0 past the vertex; the DRT
    # term does: d(sigma_t * albedo * L_i) / d albedo = sigma_t * L_i)
    assert da.any() and rel_l2(d4[..., 0:3], da) < 1e-5


@pytest.mark.parametrize("case", range(24))
def test_randomised_scenes_lane_for_lane(oracle, emu, case):
    """Seeded random scenes — ragged grids, cameras outside and inside the box, scaled boxes, both phase functions,
    Russian roulette, box / gaussian film, multi-view films, majorant supergrids — through both tracers, the adjoint and
    the DRT stages of the event machines: spp = 1 accumulators bit for bit, gradients to float-summation order."""
    from recursive_control_variates_b200.scene import SceneConfig, look_at
    rng = np.random.default_rng(1000 + case)
    nz, ny, nx = (int(v) for v in rng.integers(1, 9, 3))
    factor = 0
    if case % 3 == 2:                                     # supergrid: dims must be multiples of the factor
        factor = int(rng.integers(2, 4))
        nz, ny, nx = (factor * int(v) for v in rng.integers(1, 4, 3))
    bmax = [1.0, float(rng.uniform(0.6, 1.6)), float(rng.uniform(0.6, 1.6))]
    inside = case % 4 == 3
    eye = [float(rng.uniform(0.2, 0.8)) * b for b in bmax] if inside else \
        [float(v) for v in (rng.uniform(1.5, 3.0, 3) * rng.choice([-1, 1], 3) + 0.5)]
    target = [float(rng.uniform(0.3, 0.7)) * b for b in bmax]
    n_views = 2 if case % 5 == 4 else 1
    W = int(rng.integers(3, 14)) * n_views; H = int(rng.integers(3, 12))
    views = look_at(eye, target, [0, 1, 0])
    if n_views == 2:
        views = np.stack([views, look_at([eye[2], eye[0], eye[1]], target, [0, 0, 1])])
    hg = bool(rng.integers(0, 2))
    d = SceneConfig(width=W, height=H, fov=float(rng.uniform(25, 70)), to_world=views, bbox_max=bmax,
                    sigma_t_scale=float(rng.uniform(2, 40)), phase="hg" if hg else "isotropic",
                    phase_g=float(rng.uniform(-0.8, 0.9)) if hg else 0.0, max_depth=int(rng.integers(1, 20)),
                    rr_depth=int(rng.integers(0, 6)), hide_emitters=bool(rng.integers(0, 2)),
                    env_radiance=tuple(float(v) for v in rng.uniform(0.2, 1.5, 3)),
                    rfilter="gaussian" if case % 2 else "box", rfilter_stddev=float(rng.uniform(0.25, 0.5))).to_desc()
    so = rng.random((nz, ny, nx, 1), dtype=np.float32); so[rng.random(so.shape) < 0.25] = 0.0
    ao = rng.random((nz, ny, nx, 3), dtype=np.float32)
    sn = np.clip(so + rng.normal(0, 0.02, so.shape).astype(np.float32), 0, 1).astype(np.float32)
    an = np.clip(ao + rng.normal(0, 0.02, ao.shape).astype(np.float32), 0, 1).astype(np.float32)
    mo = oracle.OracleMedium(d, so, ao, majorant_factor=factor); mn = oracle.OracleMedium(d, sn, an, majorant_factor=factor)
    acc_o = np.zeros((H, W, 4), np.float32); acc_n = np.zeros((H, W, 4), np.float32); cnt = np.zeros(8, np.uint64)
    emu.emu_twostates_sg(C.addressof(d), so, ao, sn, an, nz, ny, nx, 5, 1, 0, factor, acc_o, acc_n, cnt)
    ref_o, ref_n = oracle.render_twostates_accum(d, mo, mn, 5, 1, threads=1)
    # box filter: one sample per pixel, bit for bit; gaussian: footprints overlap, the splats sum in another order
    same = (lambda a, b: np.array_equal(a, b, equal_nan=True)) if case % 2 == 0 else (lambda a, b: rel_l2(a, b) < 1e-6)
    assert same(acc_o, ref_o) and same(acc_n, ref_n)
    # the film is hit; the medium is entered unless the camera sits inside the box (reach_medium, volpathsimple.py:305-332,
    # expects a sensor outside: from inside, the ray re-spawned at the boundary leaves without meeting the medium)
    assert ref_n[..., 3].any() and (int(cnt[0]) > 0) == (not inside)
    acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((nz, ny, nx, 1), np.float32); g4 = np.zeros((nz, ny, nx, 4), np.float32)
    emu.emu_vps_sg(C.addressof(d), sn, an, nz, ny, nx, 6, 1, 0, 0, None, None, acc, g1, g4, cnt, factor)
    assert same(acc, oracle.vps_render_accum(d, mn, 6, 1, threads=1))
    gimg = (rng.standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    film_w = oracle.film_weights(d, 7, 2)
    emu.emu_vps_sg(C.addressof(d), sn, an, nz, ny, nx, 7, 2, 1, oracle.alt_seed(7), gimg.ctypes.data, film_w.ctypes.data,
                   acc, g1, g4, cnt, factor)
    gs, ga = oracle.vps_backward(d, mn, 7, 2, gimg, threads=1, film_w=film_w)
    tol = lambda a, b: rel_l2(a, b) < 2e-5 if np.any(b) else not np.any(a)
    assert tol(g1, gs) and tol(g4[..., 0:3], ga)
    d1 = np.zeros_like(g1); d4 = np.zeros_like(g4)
    emu.emu_vps_drt_stages(C.addressof(d), sn, an, nz, ny, nx, 7, 2, oracle.alt_seed(7), gimg.ctypes.data, film_w.ctypes.data,
                           d1, d4, cnt, factor, 1)
    ds, da = oracle.vps_backward(d, mn, 7, 2, gimg, threads=1, film_w=film_w, use_drt=True)
    assert tol(d1, ds) and tol(d4[..., 0:3], da)


def test_flattened_pass2_equals_the_list_walk(oracle, emu):
    """RCV_FLAT_PASS2 (off by default in the library): pass 2 of the volpathsimple adjoint taken vertex by vertex outside the
    lane machine — running result rebuilt from the records, secondary sampler skipped ahead with pcg_advance — scatters
    exactly what the list walk scatters (same sites, same values, same order on the host)."""
    d = small_scene(40, 36, max_depth=12, rr_depth=4).to_desc()
    _, (sn, an) = media(20)
    H, W = d.film_height, d.film_width
    grad_img = (np.random.default_rng(3).standard_normal((H, W, 3)) / (H * W * 3)).astype(np.float32)
    seed_grad, spp = 0xC09848F2, 3
    film_w = oracle.film_weights(d, seed_grad, spp)
    out = []
    try:
        for flat in (0, 1):
            emu.emu_set_flat(flat)
            acc = np.zeros((H, W, 4), np.float32); g1 = np.zeros((*sn.shape[:3], 1), np.float32); g4 = np.zeros((*sn.shape[:3], 4), np.float32)
            cnt = np.zeros(8, np.uint64)
            emu.emu_vps(C.addressof(d), sn, an, *sn.shape[:3], seed_grad, spp, 1, oracle.alt_seed(seed_grad), grad_img.ctypes.data,
                        film_w.ctypes.data, acc, g1, g4, cnt)
            out.append((g1, g4, cnt))
    finally:
        emu.emu_set_flat(0)
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert np.abs(out[1][0]).max() > 0
