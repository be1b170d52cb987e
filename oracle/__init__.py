"""Python (ctypes + numpy) front end of the CPU oracle.

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never by the product package.
PARITY UNPINNED: see oracle/rcv_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librcv_oracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


class Medium(C.Structure):
    _fields_ = [("sigma_t", C.c_void_p), ("albedo", C.c_void_p),
                ("nz", C.c_int32), ("ny", C.c_int32), ("nx", C.c_int32),
                ("majorant", C.c_float),
                ("majorant_grid", C.c_void_p), ("mz", C.c_int32), ("my", C.c_int32), ("mx", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile librcv_oracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "rcv_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "librcv_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        SP, MP = C.c_void_p, C.POINTER(Medium)
        u32, u64, i32, f = C.c_uint32, C.c_uint64, C.c_int, C.c_float
        L.orc_tea32.argtypes = [u32, u32, C.POINTER(u32), C.POINTER(u32)]
        L.orc_pcg32_kat.argtypes = [u64, u64, _u32p, i32]
        L.orc_u32_to_f32.argtypes = [u32]; L.orc_u32_to_f32.restype = f
        L.orc_logf.argtypes = [f]; L.orc_logf.restype = f
        L.orc_sincosf.argtypes = [f, C.POINTER(f), C.POINTER(f)]
        L.orc_majorant.argtypes = [SP, _f32p, C.c_size_t]; L.orc_majorant.restype = f
        L.orc_majorant_grid.argtypes = [SP, _f32p, i32, i32, i32, i32, _f32p]; L.orc_majorant_grid.restype = i32
        L.orc_alt_seed.argtypes = [u32]; L.orc_alt_seed.restype = u32
        L.orc_atan2f.argtypes = [f, f]; L.orc_atan2f.restype = f
        L.orc_env_eval.argtypes = [SP, _f32p, _f32p, C.POINTER(f)]
        L.orc_env_sample.argtypes = [SP, f, f, _f32p, C.POINTER(f), _f32p]
        L.orc_dda_sample.argtypes = [SP, C.c_void_p, MP, _f32p, _f32p, f, f, f]; L.orc_dda_sample.restype = f
        L.orc_counters.argtypes = [_u64p, i32]
        L.orc_twostates_lanes.argtypes = [SP, MP, MP, u32, u32, u64, u64, _f32p, _f32p, _i32p, i32]
        L.orc_render_twostates.argtypes = [SP, MP, MP, u32, u32, u32, u32, _f32p, _f32p, i32]
        L.orc_develop.argtypes = [_f32p, _f32p, u64]
        L.orc_film_weights.argtypes = [SP, u32, u32, _f32p]
        L.orc_film_weights_range.argtypes = [SP, u32, u32, u32, u32, _f32p]
        L.orc_cv_update.argtypes = [_f32p] * 10 + [u64, i32, i32, i32, f, f, i32]
        L.orc_loss_grad.argtypes = [i32, _f32p, _f32p, u64, C.c_void_p]; L.orc_loss_grad.restype = C.c_double
        L.orc_vps_lanes.argtypes = [SP, MP, u32, u32, u64, u64, _f32p, _i32p, i32]
        L.orc_vps_render.argtypes = [SP, MP, u32, u32, u32, u32, _f32p, i32]
        L.orc_vps_backward.argtypes = [SP, MP, u32, u32, u32, u32, _f32p, _f32p, _f64p, _f64p, i32]
        L.orc_prb_backward.argtypes = [SP, MP, u32, u32, u32, u32, _f32p, _f32p, _f64p, _f64p, i32]
        L.orc_vps_backward_drt.argtypes = [SP, MP, u32, u32, u32, u32, _f32p, _f32p, _f64p, _f64p, i32, i32, i32]
        L.orc_phase_sample.argtypes = [SP, _f32p, f, f, _f32p]; L.orc_phase_sample.restype = f
        L.orc_phase_eval.argtypes = [SP, _f32p, _f32p]; L.orc_phase_eval.restype = f
        L.orc_adam_lr_t.argtypes = [f, f, f, i32]; L.orc_adam_lr_t.restype = f
        L.orc_adam.argtypes = [_f32p, _f32p, _f32p, _f32p, u64, i32, f, f, f, f]
        L.orc_clamp.argtypes = [_f32p, u64, f, f]
        L.orc_upsample2x.argtypes = [_f32p, _f32p, i32, i32, i32, i32]
        _lib = L
    return _lib


def default_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


# --------------------------------------------------------------------------
def tea32(v0: int, v1: int):
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_tea32(v0, v1, C.byref(a), C.byref(b))
    return a.value, b.value


def pcg32_kat(initstate: int, initseq: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_pcg32_kat(initstate, initseq, out, n)
    return out


def logf(x: float) -> float:
    return lib().orc_logf(x)


def sincosf(x: float):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincosf(x, C.byref(s), C.byref(c))
    return s.value, c.value


def dda_sample(desc, m_old, m_new, o, d, mint, maxt, tau) -> float:
    """Free-flight distance through the majorant supergrid(s) (test hook); m_old may be None."""
    return float(lib().orc_dda_sample(C.addressof(desc), C.addressof(m_old.c) if m_old is not None else None, C.byref(m_new.c),
                                      np.asarray(o, np.float32), np.asarray(d, np.float32), mint, maxt, tau))


def atan2f(y: float, x: float) -> float:
    return float(lib().orc_atan2f(y, x))


def env_eval(desc, d):
    """(radiance rgb, solid-angle pdf) of the scene's environment emitter for a world direction d."""
    rgb = np.zeros(3, np.float32); pdf = C.c_float()
    lib().orc_env_eval(C.addressof(desc), np.ascontiguousarray(d, np.float32), rgb, C.byref(pdf))
    return rgb, float(pdf.value)


def env_sample(desc, u1: float, u2: float):
    """(direction, pdf, radiance / pdf) of an environment sample."""
    d = np.zeros(3, np.float32); val = np.zeros(3, np.float32); pdf = C.c_float()
    lib().orc_env_sample(C.addressof(desc), float(u1), float(u2), d, C.byref(pdf), val)
    return d, float(pdf.value), val


def alt_seed(seed_grad: int) -> int:
    return lib().orc_alt_seed(seed_grad)


def counters(reset: bool = True) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_counters(out, int(reset))
    return out


class OracleMedium:
    """One medium state; keeps the numpy arrays alive for the C struct."""

    def __init__(self, desc, sigma_t: np.ndarray, albedo: np.ndarray, majorant_factor: int = 0):
        """majorant_factor > 0: majorant supergrid with cells of that many voxels (heterogeneous
        `majorant_resolution_factor`); 0: one majorant for the whole volume."""
        self.sigma_t = np.ascontiguousarray(sigma_t, dtype=np.float32)
        self.albedo = np.ascontiguousarray(albedo, dtype=np.float32)
        nz, ny, nx = self.sigma_t.shape[:3]
        assert self.albedo.shape[:3] == (nz, ny, nx) and self.albedo.shape[-1] == 3
        self.shape = (nz, ny, nx)
        self.c = Medium()
        self.c.sigma_t = self.sigma_t.ctypes.data
        self.c.albedo = self.albedo.ctypes.data
        self.c.nz, self.c.ny, self.c.nx = nz, ny, nx
        self.c.majorant = lib().orc_majorant(C.addressof(desc), self.sigma_t.reshape(-1), self.sigma_t.size)
        self.majorant_grid = None
        self.c.majorant_grid = None
        if majorant_factor > 0:
            f = int(majorant_factor)
            if nz % f or ny % f or nx % f:
                raise ValueError(f"grid resolution {(nz, ny, nx)} must be divisible by the majorant factor {f}")
            self.majorant_grid = np.zeros((nz // f, ny // f, nx // f), np.float32)
            rc = lib().orc_majorant_grid(C.addressof(desc), self.sigma_t.reshape(-1), nz, ny, nx, f, self.majorant_grid.reshape(-1))
            assert rc == 0
            self.c.majorant_grid = self.majorant_grid.ctypes.data
            self.c.mz, self.c.my, self.c.mx = self.majorant_grid.shape

    @property
    def majorant(self) -> float:
        return float(self.c.majorant)


def twostates_lanes(desc, m_old: OracleMedium, m_new: OracleMedium, seed, spp, lane_begin, lane_end, threads=None):
    n = lane_end - lane_begin
    Lo = np.zeros((n, 3), np.float32); Ln = np.zeros((n, 3), np.float32); pix = np.zeros(n, np.int32)
    lib().orc_twostates_lanes(C.addressof(desc), C.byref(m_old.c), C.byref(m_new.c), seed, spp, lane_begin, lane_end,
                              Lo, Ln, pix, threads or default_threads())
    return Lo, Ln, pix


def render_twostates_accum(desc, m_old, m_new, seed, spp, spp_begin=0, spp_end=None, threads=None):
    spp_end = spp if spp_end is None else spp_end
    H, W = desc.film_height, desc.film_width
    ao = np.zeros((H, W, 4), np.float32); an = np.zeros((H, W, 4), np.float32)
    rc = lib().orc_render_twostates(C.addressof(desc), C.byref(m_old.c), C.byref(m_new.c), seed, spp, spp_begin, spp_end,
                                    ao, an, threads or default_threads())
    assert rc == 0
    return ao, an


def develop(accum: np.ndarray) -> np.ndarray:
    H, W = accum.shape[:2]
    img = np.zeros((H, W, 3), np.float32)
    lib().orc_develop(np.ascontiguousarray(accum), img, H * W)
    return img


def render_twostates(desc, m_old, m_new, seed, spp, threads=None):
    ao, an = render_twostates_accum(desc, m_old, m_new, seed, spp, threads=threads)
    return develop(ao), develop(an)


def film_weights(desc, seed, spp, spp_begin=0, spp_end=None) -> np.ndarray:
    """Film weights of samples [spp_begin, spp_end) of every pixel (all samples by default)."""
    Wt = np.zeros((desc.film_height, desc.film_width), np.float32)
    lib().orc_film_weights_range(C.addressof(desc), seed, spp, spp_begin, spp if spp_end is None else spp_end, Wt)
    return Wt


def cv_update(H, F, state: dict, it: int, warmup: int, stats_n: int, beta1: float, beta2: float, pre_update=False):
    """In-place update of state{img,v_n,w_s,mean_x,mean_y,var_x,var_y,cov} (float32 arrays)."""
    n = H.size
    lib().orc_cv_update(H.reshape(-1), F.reshape(-1), *[state[k].reshape(-1) for k in
                        ("img", "v_n", "w_s", "mean_x", "mean_y", "var_x", "var_y", "cov")],
                        n, it, warmup, stats_n, beta1, beta2, int(pre_update))


def loss_grad(kind: int, img, ref, want_grad=True):
    g = np.zeros(img.shape, np.float32) if want_grad else None
    v = lib().orc_loss_grad(kind, np.ascontiguousarray(img).reshape(-1), np.ascontiguousarray(ref).reshape(-1), img.size,
                            g.ctypes.data if g is not None else None)
    return v, g


def vps_lanes(desc, m: OracleMedium, seed, spp, lane_begin, lane_end, threads=None):
    n = lane_end - lane_begin
    L = np.zeros((n, 3), np.float32); pix = np.zeros(n, np.int32)
    lib().orc_vps_lanes(C.addressof(desc), C.byref(m.c), seed, spp, lane_begin, lane_end, L, pix, threads or default_threads())
    return L, pix


def vps_render_accum(desc, m, seed, spp, spp_begin=0, spp_end=None, threads=None):
    spp_end = spp if spp_end is None else spp_end
    a = np.zeros((desc.film_height, desc.film_width, 4), np.float32)
    rc = lib().orc_vps_render(C.addressof(desc), C.byref(m.c), seed, spp, spp_begin, spp_end, a, threads or default_threads())
    assert rc == 0
    return a


def vps_backward(desc, m: OracleMedium, seed_grad, spp_grad, grad_img, spp_begin=0, spp_end=None, threads=None, film_w=None,
                 use_drt=False, use_drt_subsampling=True, use_drt_mis=True):
    """render_backward of volpathsimple.  use_drt=True (volpathsimple.py:33-36) exists in this oracle only."""
    spp_end = spp_grad if spp_end is None else spp_end
    nz, ny, nx = m.shape
    gs = np.zeros((nz, ny, nx, 1), np.float64); ga = np.zeros((nz, ny, nx, 3), np.float64)
    if film_w is None:
        film_w = film_weights(desc, seed_grad, spp_grad)
    args = (C.addressof(desc), C.byref(m.c), seed_grad, spp_grad, spp_begin, spp_end,
            np.ascontiguousarray(grad_img, np.float32).reshape(-1), film_w.reshape(-1),
            gs.reshape(-1), ga.reshape(-1), threads or default_threads())
    if use_drt:
        rc = lib().orc_vps_backward_drt(*args, int(use_drt_subsampling), int(use_drt_mis))
    else:
        rc = lib().orc_vps_backward(*args)
    assert rc == 0
    return gs, ga


def prb_backward(desc, m: OracleMedium, seed_grad, spp_grad, grad_img, spp_begin=0, spp_end=None, threads=None, film_w=None):
    """render_backward of upstream prbvolpath (the adjoint twostateprbvolpath inherits) on state m."""
    spp_end = spp_grad if spp_end is None else spp_end
    nz, ny, nx = m.shape
    gs = np.zeros((nz, ny, nx, 1), np.float64); ga = np.zeros((nz, ny, nx, 3), np.float64)
    if film_w is None:
        film_w = film_weights(desc, seed_grad, spp_grad)
    rc = lib().orc_prb_backward(C.addressof(desc), C.byref(m.c), seed_grad, spp_grad, spp_begin, spp_end,
                                np.ascontiguousarray(grad_img, np.float32).reshape(-1), film_w.reshape(-1),
                                gs.reshape(-1), ga.reshape(-1), threads or default_threads())
    assert rc == 0
    return gs, ga


def phase_sample(desc, wi, s2x, s2y):
    """(wo, pdf) of the scene's phase function for incident direction wi (pointing away from the vertex)."""
    wo = np.zeros(3, np.float32)
    pdf = lib().orc_phase_sample(C.addressof(desc), np.ascontiguousarray(wi, np.float32), float(s2x), float(s2y), wo)
    return wo, float(pdf)


def phase_eval(desc, wi, wo):
    return float(lib().orc_phase_eval(C.addressof(desc), np.ascontiguousarray(wi, np.float32), np.ascontiguousarray(wo, np.float32)))


def adam(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    lib().orc_adam(p.reshape(-1), g.reshape(-1), m.reshape(-1), v.reshape(-1), p.size, t, lr, beta1, beta2, eps)


def clamp(p, lo, hi):
    lib().orc_clamp(p.reshape(-1), p.size, lo, hi)


def upsample2x(a: np.ndarray) -> np.ndarray:
    nz, ny, nx, c = a.shape
    out = np.zeros((2 * nz, 2 * ny, 2 * nx, c), np.float32)
    lib().orc_upsample2x(np.ascontiguousarray(a), out, nz, ny, nx, c)
    return out
