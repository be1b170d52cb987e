"""The C-ABI library loads without a GPU, exports every symbol include/rcv_b200.h declares,
and its compute entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rcv_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from recursive_control_variates_b200 import _abi
    return _abi.load_library()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rcv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from recursive_control_variates_b200 import _abi
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rcv_b200.h but not exported"
    assert set(names) == set(_abi.EXPORTED_SYMBOLS), set(names) ^ set(_abi.EXPORTED_SYMBOLS)


def test_struct_layouts_match_header(lib):
    from recursive_control_variates_b200 import _abi
    # 4 + 2*4 + 3*4 + 9*4 + 3*4 + 6*4 + 4 + 4 + 4 + 4 + 3*4 + 4 + 4*4 = 144, + n_sensors 4 + 16 * (3 + 9) * 4 = 916, + rfilter 8 = 924
    # ABI 4: + emitter_type 4, envmap pointer 8, env_width/height 8, env_scale 4, env_to_world 36, n_quads 4, area_quad 4, 8 quads x 60
    assert C.sizeof(_abi.SceneDesc) == 924 + 4 + 8 + 8 + 4 + 36 + 4 + 4 + 8 * 60 == 1472
    assert lib.rcv_abi_version() == 4


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.rcv_device_count() <= 0
    h = C.c_void_p()
    rc = lib.rcv_ctx_create(0, None, C.byref(h))
    assert rc == 2 and b"no CPU fallback" in lib.rcv_last_error()          # RCV_ERR_CUDA
    from recursive_control_variates_b200 import _abi
    from recursive_control_variates_b200.runtime import Context
    with pytest.raises(_abi.RcvError):
        Context()


def test_null_context_is_an_argument_error(lib):
    """Setters reject a null context before touching the device (no compute call, runs without a GPU)."""
    for fn, args in ((lib.rcv_ctx_set_tracer, (None, 3)), (lib.rcv_adjoint_set_drt, (None, 1, 1, 1)),
                     (lib.rcv_ctx_set_stream, (None, None))):
        assert fn(*args) == 1 and b"null" in lib.rcv_last_error()            # RCV_ERR_ARG


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/."""
    pkg = os.path.join(ROOT, "recursive_control_variates_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "rcv_oracle" not in text, f


def test_scene_validation_without_a_gpu(lib):
    """rcv_scene_set validates the description before it touches the device: bad rectangles, a missing area light and an
    envmap without texels are argument errors (no compute call, runs without a GPU)."""
    import numpy as np
    from recursive_control_variates_b200 import _abi
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import boxed_quads

    def status(desc):
        h = C.c_void_p()
        rc = lib.rcv_ctx_create(0, None, C.byref(h))
        if rc != 0:                  # no device (the CPU suite): only the host-side checks below run
            return None
        try:
            return lib.rcv_scene_set(h, C.byref(desc))
        finally:
            lib.rcv_ctx_destroy(h)

    good = SceneConfig(width=8, height=8, quads=boxed_quads()).to_desc()
    assert good.n_quads == 7 and good.area_quad == 6 and good.emitter_type == _abi.RCV_EMITTER_AREA
    # host-side construction errors (scene.py) need no library at all
    with pytest.raises(ValueError):
        SceneConfig(width=8, height=8, quads=boxed_quads() * 2).to_desc()                     # more than RCV_MAX_QUADS
    with pytest.raises(ValueError):
        SceneConfig(width=8, height=8, quads=boxed_quads(), envmap=np.ones((4, 8, 3), np.float32)).to_desc()   # two emitters
    with pytest.raises(ValueError):
        SceneConfig(width=8, height=8, envmap=np.ones((1, 8, 3), np.float32)).to_desc()       # a single row of texels
    if status(good) is None:
        return
    assert status(good) == 0
    bad = SceneConfig(width=8, height=8, quads=boxed_quads()).to_desc()
    bad.quads[0].ev[:] = list(bad.quads[0].eu)                                                # edges not perpendicular
    assert status(bad) == 1 and b"perpendicular" in lib.rcv_last_error()
    bad = SceneConfig(width=8, height=8, quads=boxed_quads()).to_desc()
    bad.area_quad = 9
    assert status(bad) == 1 and b"area_quad" in lib.rcv_last_error()
    bad = SceneConfig(width=8, height=8).to_desc()
    bad.emitter_type = _abi.RCV_EMITTER_ENVMAP
    assert status(bad) == 1 and b"envmap" in lib.rcv_last_error()
