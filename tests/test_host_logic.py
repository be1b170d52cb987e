"""Host-side logic that needs no GPU: scene description, sharding, the Mitsuba-compat shim."""
import ctypes as C

import numpy as np
import pytest

from recursive_control_variates_b200 import mitsuba_compat as mi
from recursive_control_variates_b200 import parallel, plugins
from recursive_control_variates_b200.scene import SceneConfig, look_at
from recursive_control_variates_b200.synthetic import make_medium


def test_look_at_is_orthonormal():
    m = look_at([0.5, 0.5, -2.0], [0.5, 0.5, 0.5], [0, 1, 0])
    R = m[:3, :3]
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-12)
    assert np.allclose(R[:, 2], [0, 0, 1]) and np.allclose(m[:3, 3], [0.5, 0.5, -2.0])
    assert np.allclose(R[:, 0], [1, 0, 0])       # Mitsuba: camera x axis = cross(up, dir)


def test_scene_desc_fields():
    d = SceneConfig(width=64, height=32, fov=60.0).to_desc()
    assert d.struct_size == C.sizeof(d) and (d.film_width, d.film_height) == (64, 32)
    assert abs(d.cam_tan_x - np.tan(np.radians(30))) < 1e-6 and abs(d.cam_tan_y - d.cam_tan_x / 2) < 1e-6
    assert abs(d.env_dist - np.sqrt(3.0) * (1 + 1500 * 2.0 ** -24)) < 1e-5
    assert d.max_depth == 64 and d.rr_depth == 64 and d.use_nee == 1


def test_synthetic_media():
    s, a = make_medium(8, "cube", 0.0)
    s2, a2 = make_medium(8, "cube", 5e-3)
    assert s.shape == (8, 8, 8, 1) and a.shape == (8, 8, 8, 3) and s.dtype == np.float32
    assert 0 <= s.min() and s.max() <= 1 and 0 <= a.min() and a.max() <= 1
    assert 0 < np.abs(s2 - s).max() <= 5e-3 + 1e-6
    for kind in ("plume", "dense"):
        assert make_medium(4, kind)[0].shape == (4, 4, 4, 1)
    with pytest.raises(ValueError):
        make_medium(4, "nope")


def test_shard_spp_partitions():
    for spp in (1, 7, 64, 100):
        for world in (1, 2, 3, 8):
            ranges = [parallel.shard_spp(spp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_spp(4, 2, 2)
    assert parallel.world() == (0, 1)


def _scene(method="cv_ps", n=4):
    from recursive_control_variates_b200 import experiments
    return experiments.scene_synthetic(method, None, grid=n, init_grid=2, film=(8, 6), spp=2, n_steps=2)


def test_registered_plugin_names():
    # plugins/__init__.py:11-17 of the reference + the north_star aliases
    for name in ("volpathsimple", "meta", "cv", "twostateprbvolpath", "prbvolpath", "twostatevolpath", "cv_ps"):
        assert name in mi._INTEGRATORS
    assert "twostatemedium" in mi._MEDIA


def test_load_dict_and_properties():
    p = _scene("cv_ps")
    scene = p["scene"]
    assert isinstance(scene.medium(), plugins.TwoStateMedium)
    assert isinstance(p["integrator"], plugins.TwoStatePRBVolpathIntegrator) and p["integrator"].max_depth == 64
    assert isinstance(p["adjoint_integrator"], plugins.VolpathSimpleIntegrator) and p["adjoint_integrator"].use_nee
    meta = mi.load_dict({"type": "meta", "method": "cv_ps", "integrator": p["integrator"],
                         "adjoint_integrator": p["adjoint_integrator"], "beta1": 0.8})
    assert meta.method == "cv_ps" and meta.beta1 == 0.8 and meta.warmup == 1 and meta.adjoint_integrator is p["adjoint_integrator"]
    cv = mi.load_dict({"type": "cv", "integrator": {"type": "twostateprbvolpath"}})
    assert cv.adjoint_integrator is cv.integrator and (cv.warmup, cv.beta1, cv.beta2) == (1, 0.9, 0.999)
    assert isinstance(_scene("baseline")["scene"].medium(), mi.HeterogeneousMedium)


def test_error_behaviour_matches_reference():
    with pytest.raises(ValueError):          # cv_integrator.py:19-20
        mi.load_dict({"type": "cv", "integrator": {"type": "volpathsimple"}})
    vps = mi.load_dict({"type": "volpathsimple"})                      # volpathsimple.py:39-47: DRT is on by default
    assert vps.use_drt and vps.use_drt_subsampling and vps.use_drt_mis
    vps = mi.load_dict({"type": "volpathsimple", "use_drt": False, "use_drt_mis": False})
    assert not vps.use_drt and vps.use_drt_subsampling and not vps.use_drt_mis
    with pytest.raises(NotImplementedError):  # OptiX denoiser + AOVs: outside the hot path
        mi.load_dict({"type": "meta", "method": "cv_pss", "denoise": True, "integrator": {"type": "prbvolpath"}})
    with pytest.raises(NotImplementedError):
        mi.load_dict({"type": "meta", "method": "cv_xyz", "integrator": {"type": "prbvolpath"}})
    pss = mi.load_dict({"type": "meta", "method": "cv_pss", "force_baseline": True, "integrator": {"type": "prbvolpath"}})
    assert pss.method == "cv_pss" and pss.force_baseline
    with pytest.raises(ValueError):
        mi.load_dict({"type": "scene", "object": {"type": "cube", "bsdf": {"type": "diffuse"}, "interior": {}}})
    integ = mi.load_dict({"type": "twostateprbvolpath"})
    with pytest.raises(RuntimeError):        # twostatevolpath.py:77-78: no TwoStateMedium in the scene
        integ.check_scene(_scene("baseline")["scene"])


def test_container_to_world_and_grid_resolution():
    """experiments.py:53-80 of the reference: T.scale([1,1.2,1.5]) on the container and on both volumes."""
    import torch
    T = np.diag([1.0, 1.2, 1.5, 1.0])
    vol = lambda c, n=4, **kw: {"type": "gridvolume", "data": torch.zeros(n, n, n, c), **kw}
    med = lambda **kw: {"type": "heterogeneous", "sigma_t": vol(1, **kw), "albedo": vol(3, **kw)}
    cam = {"type": "perspective", "film": {"width": 8, "height": 8}}
    sc = lambda obj: mi.load_dict({"type": "scene", "object": obj, "e": {"type": "constant"}, "sensor": cam})
    s = sc({"type": "obj", "bsdf": {"type": "null"}, "to_world": T, "interior": med(to_world=T)})
    assert s.shapes()[0].bbox_min == (0.0, 0.0, 0.0) and s.shapes()[0].bbox_max == (1.0, 1.2, 1.5)
    # Mitsuba's `cube` is [-1,1]^3 before to_world
    T2 = np.array([[0.5, 0, 0, 0.5], [0, 0.5, 0, 0.5], [0, 0, 0.5, 0.5], [0, 0, 0, 1.0]])
    s = sc({"type": "cube", "bsdf": {"type": "null"}, "to_world": T2, "interior": med()})
    assert s.shapes()[0].bbox_min == (0.0, 0.0, 0.0) and s.shapes()[0].bbox_max == (1.0, 1.0, 1.0)
    with pytest.raises(ValueError):          # a volume that does not fill the container
        sc({"type": "obj", "bsdf": {"type": "null"}, "to_world": T, "interior": med(to_world=np.eye(4))})
    R = np.eye(4); R[:2, :2] = [[0, -1], [1, 0]]
    with pytest.raises(ValueError):          # rotations are not supported: say so instead of rendering the wrong box
        sc({"type": "obj", "bsdf": {"type": "null"}, "to_world": R, "interior": med()})
    with pytest.raises(ValueError):          # janga's 264x136x136 sigma_t with a 256x128x128 albedo
        sc({"type": "obj", "bsdf": {"type": "null"},
            "interior": {"type": "heterogeneous", "sigma_t": vol(1, 6), "albedo": vol(3, 4)}})


def test_scene_parameters_and_rotation_bookkeeping():
    import torch
    scene = _scene("cv_ps")["scene"]
    sp = mi.traverse(scene)
    ks = "object.interior_medium.sigma_t.data"
    assert sorted(sp.keys()) == ["object.interior_medium.albedo.data", ks]
    with pytest.raises(KeyError):
        sp["nope"] = 1
    med = scene.medium()
    new0 = med.new.sigma_t.data
    t = torch.full((4, 4, 4, 1), 0.5)
    sp[ks] = t
    assert sp.update() == [ks]
    # twostatemedium.py:195-197: old <- new, new <- incoming
    assert med.old.sigma_t.data is new0 and torch.equal(med.new.sigma_t.data, t)
    # rotations requested before a device context exists are queued (two from the scene file + this one)
    assert len(scene._pending) == 3


def test_adam_bookkeeping():
    import torch
    opt = mi.Adam(lr=0.01)
    opt["a"] = torch.zeros(2, 2, 2, 1)
    assert opt["a"].requires_grad and opt.t["a"] == 0 and opt.learning_rate("a") == 0.01
    opt.set_learning_rate({"a": 0.02})
    assert opt.learning_rate("a") == 0.02
    opt.t["a"] = 5
    opt["a"] = torch.zeros(4, 4, 4, 1)               # shape change (upsample) resets the state
    assert opt.t["a"] == 0 and opt.state["a"][0].shape == (4, 4, 4, 1)
    opt["b"] = torch.zeros(2)
    opt.set_learning_rate(0.005)                     # float form: the default only, per-key rates are kept
    assert opt.learning_rate("a") == 0.02 and opt.learning_rate("b") == 0.005


def test_vol_files_round_trip(tmp_path):
    """Mitsuba .vol grids (utils.py:45-60 save_final_state; experiments.py:56 gridvolume filename)."""
    import struct
    import torch
    from recursive_control_variates_b200.volio import read_vol, write_vol
    from recursive_control_variates_b200.utils import save_final_state
    rng = np.random.default_rng(0)
    g = rng.random((3, 4, 5, 3), dtype=np.float32)
    path = str(tmp_path / "a.vol")
    write_vol(path, g, (0, 0, 0), (1.0, 1.2, 1.5))
    raw = open(path, "rb").read()
    assert raw[:4] == b"VOL\x03" and struct.unpack("<iiiii", raw[4:24]) == (1, 5, 4, 3, 3)       # encoding, x, y, z, channels
    assert struct.unpack("<6f", raw[24:48])[3:] == pytest.approx((1.0, 1.2, 1.5))
    assert struct.unpack("<f", raw[48:52])[0] == g[0, 0, 0, 0] and len(raw) == 48 + g.size * 4
    back, bmin, bmax = read_vol(path)
    assert np.array_equal(back, g) and bmin == (0.0, 0.0, 0.0) and bmax == pytest.approx((1.0, 1.2, 1.5))
    vol = mi.GridVolume(mi.Properties({"filename": path}), 3)
    assert torch.equal(vol.data, torch.from_numpy(g))
    with pytest.raises(ValueError):
        mi.GridVolume(mi.Properties({"filename": path}), 1)              # channel mismatch
    open(str(tmp_path / "bad.vol"), "wb").write(raw[:100])
    with pytest.raises(ValueError):
        read_vol(str(tmp_path / "bad.vol"))                              # truncated
    opt = {"m.sigma_t.data": torch.from_numpy(g[..., :1].copy()), "m.scalar": torch.tensor([0.5])}
    save_final_state({"variables": {k: {} for k in opt}}, opt, str(tmp_path))
    s, _, _ = read_vol(str(tmp_path / "m_sigma_t_data_final.vol"))
    assert np.array_equal(s, g[..., :1])
    import json
    assert json.load(open(str(tmp_path / "final_state.json"))) == {"m.scalar": [0.5]}


def test_exr_files_round_trip(tmp_path):
    """Uncompressed float32 scanline OpenEXR (utils.save_img, utils.py:8-14): header fields and pixel round trip."""
    import struct
    from recursive_control_variates_b200.exrio import read_exr, write_exr
    from recursive_control_variates_b200.utils import save_img
    rng = np.random.default_rng(1)
    img = rng.random((5, 7, 3), dtype=np.float32)
    path = save_img(img, str(tmp_path), "img", 12)
    assert path.endswith("img/0012.exr")
    raw = open(path, "rb").read()
    assert struct.unpack_from("<ii", raw, 0) == (20000630, 2) and b"channels\0chlist\0" in raw and b"dataWindow\0box2i\0" in raw
    back, names = read_exr(path)
    assert names == ["B", "G", "R"] and np.array_equal(back[..., ::-1], img)          # stored alphabetically
    write_exr(str(tmp_path / "y.exr"), img[..., 0])
    y, n1 = read_exr(str(tmp_path / "y.exr"))
    assert n1 == ["Y"] and np.array_equal(y[..., 0], img[..., 0])
    # header length + offset table + scanlines
    assert len(raw) == raw.index(b"screenWindowWidth") + len("screenWindowWidth\0float\0") + 4 + 4 + 1 + 8 * 5 + 5 * (8 + 4 * 7 * 3)
