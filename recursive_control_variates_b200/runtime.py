"""Thin Python front end of the C-ABI (include/rcv_b200.h) over torch tensors.

torch is used for device memory, streams and (elsewhere) torch.distributed —
plumbing only.  Every method enqueues CUDA work through librcv_b200.so on the
current torch stream; nothing here computes on the CPU and there is no
fallback path when the library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _abi
from ._abi import SceneDesc, check, load_library
from .scene import SceneConfig

CV_STATE_KEYS = ("img", "v_n", "w_s", "mean_x", "mean_y", "var_x", "var_y", "cov")


def tea32(v0: int, v1: int, rounds: int = 4) -> Tuple[int, int]:
    """mi.sample_tea_32 (host side; used for seed_grad, plugins/twostatepath.py:87-89)."""
    m = 0xFFFFFFFF
    v0 &= m; v1 &= m
    s = 0
    for _ in range(rounds):
        s = (s + 0x9E3779B9) & m
        v0 = (v0 + ((((v1 << 4) & m) + 0xA341316C) ^ ((v1 + s) & m) ^ ((v1 >> 5) + 0xC8013EA4))) & m
        v1 = (v1 + ((((v0 << 4) & m) + 0xAD90777D) ^ ((v0 + s) & m) ^ ((v0 >> 5) + 0x7E95761E))) & m
    return v0, v1


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "expected a contiguous float32 CUDA tensor"
    return t.data_ptr()


class Context:
    """Owns one rcv_ctx (packed two-state medium, gradient grid, counters) on one GPU."""

    def __init__(self, device: Optional[int] = None):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise _abi.RcvError("no CUDA device: librcv_b200 has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.torch_device = torch.device("cuda", self.device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.rcv_ctx_create(self.device, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.byref(h)))
        self._h = h
        self.desc: Optional[SceneDesc] = None
        self.grid_shape: Optional[Tuple[int, int, int]] = None
        self.majorant_factor = 0

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rcv_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def _sync_stream(self):
        check(self.lib.rcv_ctx_set_stream(self._h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))

    def set_tracer(self, kind: int):
        check(self.lib.rcv_ctx_set_tracer(self._h, kind))

    def set_drt(self, use_drt: bool, use_drt_subsampling: bool = True, use_drt_mis: bool = True):
        """volpathsimple's use_drt / use_drt_subsampling / use_drt_mis (volpathsimple.py:39-47) for render_backward."""
        check(self.lib.rcv_adjoint_set_drt(self._h, int(bool(use_drt)), int(bool(use_drt_subsampling)), int(bool(use_drt_mis))))

    def set_scene(self, scene):
        desc = scene.to_desc() if isinstance(scene, SceneConfig) else scene
        check(self.lib.rcv_scene_set(self._h, C.byref(desc)))
        self.desc = desc

    @property
    def film_shape(self) -> Tuple[int, int]:
        return int(self.desc.film_height), int(self.desc.film_width)

    def _zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float32, device=self.torch_device)

    # ---- medium state ----------------------------------------------------
    def rotate(self, sigma_t: torch.Tensor, albedo: torch.Tensor):
        """TwoStateMedium.parameters_changed: old <- new, new <- incoming."""
        assert sigma_t.dim() == 4 and sigma_t.shape[-1] == 1 and albedo.shape == (*sigma_t.shape[:3], 3)
        self._sync_stream()
        nz, ny, nx = (int(v) for v in sigma_t.shape[:3])
        check(self.lib.rcv_medium_rotate(self._h, _ptr(sigma_t), _ptr(albedo), nz, ny, nx))
        self.grid_shape = (nz, ny, nx)

    def majorants(self) -> Tuple[float, float]:
        self._sync_stream()
        out = (C.c_float * 2)()
        check(self.lib.rcv_medium_majorants_host(self._h, C.byref(out)))
        return float(out[0]), float(out[1])

    def set_majorant_factor(self, factor: int) -> None:
        """TwoStateMedium.set_majorant_resolution_factor (plugins/twostatemedium.py:150-154): factor > 1 switches every
        free-flight sample to DDA traversal of a majorant supergrid with factor^3-voxel cells; 0 / 1 switches back."""
        self._sync_stream()
        check(self.lib.rcv_medium_set_majorant_factor(self._h, int(factor)))
        self.majorant_factor = 0 if factor <= 1 else int(factor)

    def majorant_grid(self):
        """(old, new) local majorants as numpy (mz,my,mx) arrays, or None while the supergrid is disabled."""
        import numpy as np
        self._sync_stream()
        dims = (C.c_int * 3)()
        check(self.lib.rcv_medium_majorant_grid_host(self._h, None, None, C.byref(dims)))
        if dims[0] == 0:
            return None
        o = np.zeros((dims[0], dims[1], dims[2]), np.float32); n = np.zeros_like(o)
        check(self.lib.rcv_medium_majorant_grid_host(self._h, o.ctypes.data, n.ctypes.data, C.byref(dims)))
        return o, n

    def medium_read(self, state: int):
        nz, ny, nx = self.grid_shape
        s, a = self._zeros(nz, ny, nx, 1), self._zeros(nz, ny, nx, 3)
        self._sync_stream()
        check(self.lib.rcv_medium_read(self._h, state, _ptr(s), _ptr(a)))
        return s, a

    # ---- primal ------------------------------------------------------------
    def render_twostates_accum(self, seed: int, spp: int, spp_begin: int = 0, spp_end: Optional[int] = None, out=None):
        H, W = self.film_shape
        spp_end = spp if spp_end is None else spp_end
        acc_old, acc_new = out if out is not None else (self._zeros(H, W, 4), self._zeros(H, W, 4))
        self._sync_stream()
        check(self.lib.rcv_render_twostates(self._h, seed & 0xFFFFFFFF, spp, spp_begin, spp_end, _ptr(acc_old), _ptr(acc_new)))
        return acc_old, acc_new

    def render_primal_accum(self, state: int, seed: int, spp: int, spp_begin: int = 0, spp_end: Optional[int] = None, out=None):
        H, W = self.film_shape
        spp_end = spp if spp_end is None else spp_end
        acc = out if out is not None else self._zeros(H, W, 4)
        self._sync_stream()
        check(self.lib.rcv_render_primal(self._h, state, seed & 0xFFFFFFFF, spp, spp_begin, spp_end, _ptr(acc)))
        return acc

    def render_adjoint_primal_accum(self, seed: int, spp: int, spp_begin: int = 0, spp_end: Optional[int] = None, out=None,
                                    adjoint: int = _abi.RCV_ADJOINT_VOLPATHSIMPLE):
        H, W = self.film_shape
        spp_end = spp if spp_end is None else spp_end
        acc = out if out is not None else self._zeros(H, W, 4)
        self._sync_stream()
        check(self.lib.rcv_render_adjoint_primal(self._h, adjoint, seed & 0xFFFFFFFF, spp, spp_begin,
                                                 spp_end, _ptr(acc)))
        return acc

    def develop(self, accum: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        H, W = self.film_shape
        img = out if out is not None else self._zeros(H, W, 3)
        self._sync_stream()
        check(self.lib.rcv_develop(self._h, _ptr(accum), _ptr(img)))
        return img

    def render_twostates(self, seed: int, spp: int):
        acc_old, acc_new = self.render_twostates_accum(seed, spp)
        return self.develop(acc_old), self.develop(acc_new)

    # ---- recursive control variate -------------------------------------------
    def new_cv_state(self) -> dict:
        H, W = self.film_shape
        return {k: self._zeros(H, W, 3) for k in CV_STATE_KEYS}

    def cv_update(self, H: torch.Tensor, F: torch.Tensor, state: dict, it: int, warmup: int, stats_n: int,
                  beta1: float, beta2: float, pre_update: bool = False):
        self._sync_stream()
        check(self.lib.rcv_cv_update(self._h, _ptr(H), _ptr(F), *[_ptr(state[k]) for k in CV_STATE_KEYS], H.numel(),
                                     it, warmup, stats_n, beta1, beta2, int(pre_update)))

    def loss_grad(self, kind: int, img: torch.Tensor, ref: torch.Tensor, want_grad: bool = True):
        loss = self._zeros(1)
        grad = torch.empty_like(img) if want_grad else None
        self._sync_stream()
        check(self.lib.rcv_loss_grad(self._h, kind, _ptr(img), _ptr(ref), img.numel(), _ptr(loss), _ptr(grad)))
        return loss, grad

    # ---- adjoint -----------------------------------------------------------------
    def grad_zero(self):
        self._sync_stream()
        check(self.lib.rcv_grad_zero(self._h))

    def render_backward(self, seed_grad: int, spp_grad: int, grad_img: torch.Tensor, spp_begin: int = 0,
                        spp_end: Optional[int] = None, film_w: Optional[torch.Tensor] = None,
                        adjoint: int = _abi.RCV_ADJOINT_VOLPATHSIMPLE):
        """film_w: (H,W) film weights of ALL spp_grad samples (see film_weights); None = computed inside the call.
        adjoint: RCV_ADJOINT_VOLPATHSIMPLE (plugins/volpathsimple.py) or RCV_ADJOINT_PRBVOLPATH (the adjoint
        twostateprbvolpath inherits from upstream prbvolpath; the one that handles surfaces and area lights)."""
        spp_end = spp_grad if spp_end is None else spp_end
        self._sync_stream()
        check(self.lib.rcv_render_backward_weighted(self._h, adjoint, seed_grad & 0xFFFFFFFF, spp_grad,
                                                    spp_begin, spp_end, _ptr(grad_img), _ptr(film_w)))

    def film_weights(self, seed: int, spp: int, spp_begin: int = 0, spp_end: Optional[int] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Adds the film weights of samples [spp_begin, spp_end) to `out` (a new zeroed (H,W) buffer by default): the
        per-rank piece of the weights the adjoint needs; all-reduce it across an spp-sharded job."""
        spp_end = spp if spp_end is None else spp_end
        H, W = self.film_shape
        if out is None:
            out = self._zeros(H, W)
        self._sync_stream()
        check(self.lib.rcv_film_weights(self._h, seed & 0xFFFFFFFF, spp, spp_begin, spp_end, _ptr(out)))
        return out

    def grad_packed(self) -> torch.Tensor:
        """Flat view of the library's gradient grids (d sigma_t, d albedo) for an in-place all-reduce."""
        p, n = C.c_void_p(), C.c_uint64()
        check(self.lib.rcv_grad_buffer(self._h, C.byref(p), C.byref(n)))
        return _as_tensor(p.value, int(n.value), self.torch_device)

    def grad_read(self):
        nz, ny, nx = self.grid_shape
        gs, ga = self._zeros(nz, ny, nx, 1), self._zeros(nz, ny, nx, 3)
        self._sync_stream()
        check(self.lib.rcv_grad_read(self._h, _ptr(gs), _ptr(ga)))
        return gs, ga

    # ---- optimiser -----------------------------------------------------------------
    def adam_step(self, sigma_t, albedo, m_s, v_s, m_a, v_a, t, lr_sigma_t, lr_albedo, beta1=0.9, beta2=0.999, eps=1e-8,
                  clamp_sigma=(0.0, 250.0), clamp_albedo=(0.0, 1.0)):
        self._sync_stream()
        check(self.lib.rcv_adam_step(self._h, _ptr(sigma_t), _ptr(albedo), _ptr(m_s), _ptr(v_s), _ptr(m_a), _ptr(v_a), t,
                                     lr_sigma_t, lr_albedo, beta1, beta2, eps, clamp_sigma[0], clamp_sigma[1],
                                     clamp_albedo[0], clamp_albedo[1]))

    def adam_flat(self, p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self._sync_stream()
        check(self.lib.rcv_adam_flat(self._h, _ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), t, lr, beta1, beta2, eps))

    def clamp(self, p, lo, hi):
        self._sync_stream()
        check(self.lib.rcv_clamp(self._h, _ptr(p), p.numel(), lo, hi))

    def upsample2x(self, a: torch.Tensor) -> torch.Tensor:
        nz, ny, nx, c = (int(v) for v in a.shape)
        out = torch.empty(2 * nz, 2 * ny, 2 * nx, c, dtype=torch.float32, device=self.torch_device)
        self._sync_stream()
        check(self.lib.rcv_upsample2x(self._h, _ptr(a), _ptr(out), nz, ny, nx, c))
        return out

    def counters(self, reset: bool = True) -> np.ndarray:
        self._sync_stream()
        out = (C.c_uint64 * 8)()
        check(self.lib.rcv_counters_host(self._h, C.byref(out), int(reset)))
        return np.array(list(out), dtype=np.uint64)

    # ---- end to end with host buffers ----------------------------------------------------
    def cv_iteration_host(self, sigma_t: torch.Tensor, albedo: torch.Tensor, ref: torch.Tensor, seed: int, spp: int,
                          seed_grad: int, spp_grad: int, img_out: torch.Tensor, loss_out: torch.Tensor,
                          grad_sigma_t_out: Optional[torch.Tensor] = None, grad_albedo_out: Optional[torch.Tensor] = None,
                          loss_kind: int = _abi.RCV_LOSS_L1, beta1: float = 0.9, beta2: float = 0.999, warmup: int = 1,
                          reset: bool = False, adjoint: int = _abi.RCV_ADJOINT_VOLPATHSIMPLE):
        """One optimize.py iteration with every user-visible buffer in (pinned) host memory."""
        for t in (sigma_t, albedo, ref, img_out, loss_out, grad_sigma_t_out, grad_albedo_out):
            assert t is None or (not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous())
        a = _abi.IterHostArgs()
        a.struct_size = C.sizeof(_abi.IterHostArgs)
        a.sigma_t_in, a.albedo_in, a.ref = sigma_t.data_ptr(), albedo.data_ptr(), ref.data_ptr()
        a.nz, a.ny, a.nx = (int(v) for v in sigma_t.shape[:3])
        a.seed, a.spp, a.seed_grad, a.spp_grad = seed & 0xFFFFFFFF, spp, seed_grad & 0xFFFFFFFF, spp_grad
        a.loss_kind, a.beta1, a.beta2, a.warmup, a.reset = loss_kind, beta1, beta2, warmup, int(reset)
        a.adjoint_kind = int(adjoint)
        a.img_out, a.loss_out = img_out.data_ptr(), loss_out.data_ptr()
        a.grad_sigma_t_out = grad_sigma_t_out.data_ptr() if grad_sigma_t_out is not None else None
        a.grad_albedo_out = grad_albedo_out.data_ptr() if grad_albedo_out is not None else None
        self._sync_stream()
        check(self.lib.rcv_cv_iteration_host(self._h, C.byref(a)))
        self.grid_shape = (a.nz, a.ny, a.nx)


def _as_tensor(ptr: int, n_floats: int, device: torch.device) -> torch.Tensor:
    """Wrap a device pointer owned by the library as a float32 torch tensor (no copy)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(h, device=device)
