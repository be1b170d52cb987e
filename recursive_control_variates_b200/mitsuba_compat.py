"""Minimal stand-in for the parts of the Mitsuba 3 Python API that the reference's
hot path is driven through (`mi.load_dict`, `mi.traverse`, `mi.render`,
`mi.register_integrator`, `mi.ad.Adam`), so that the plugin classes of
`plugins/` keep the reference's names, constructor properties and call shapes
 without a Mitsuba install.

Only the scene class of the volumetric experiments is understood
(experiments.py:20-133): a perspective sensor or a `batch` of them with an hdrfilm
(box or gaussian reconstruction filter), one box shape with a `null` BSDF whose interior is a `heterogeneous` grid medium
(optionally wrapped in a `twostatemedium`), one emitter (`constant`, `envmap`, or an `area` light on a `rectangle`) and
diffuse `rectangle` shapes around the medium.

Tensors are torch CUDA tensors; differentiation goes through torch autograd:
`render()` returns an image whose backward calls the integrator's
`render_backward`, exactly like Mitsuba's `_RenderOp`.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional

import numpy as np
import torch

from .runtime import Context, tea32
from .scene import SceneConfig, look_at

_INTEGRATORS: Dict[str, Callable] = {}
_MEDIA: Dict[str, Callable] = {}


def register_integrator(name: str, ctor: Callable) -> None:
    """mi.register_integrator (plugins/__init__.py:11-15)."""
    _INTEGRATORS[name] = ctor


def register_medium(name: str, ctor: Callable) -> None:
    """mi.register_medium (plugins/__init__.py:17)."""
    _MEDIA[name] = ctor


class Properties(dict):
    """mi.Properties: `props.get(name, default)` / `props[name]`."""

    def get(self, key, default=None):  # noqa: A003 - mirrors Mitsuba
        return super().get(key, default)


def _as_tensor(data, channels: int) -> torch.Tensor:
    t = data if isinstance(data, torch.Tensor) else torch.as_tensor(np.asarray(data, dtype=np.float32))
    t = t.detach().to(dtype=torch.float32)
    assert t.dim() == 4 and t.shape[-1] == channels, f"expected a (Z,Y,X,{channels}) grid, got {tuple(t.shape)}"
    return t.contiguous()


class GridVolume:
    """`gridvolume` from an in-memory tensor (Z, Y, X, C) (experiments.py:104-112), a constant, or a `.vol` file."""

    def __init__(self, props: Properties, channels: int):
        self.box = _axis_aligned_box(props["to_world"], 0.0, 1.0, "gridvolume") if "to_world" in props else None
        if "data" in props:
            self.data = _as_tensor(props["data"], channels)
        elif "value" in props:
            res = props.get("resolution", (16, 16, 16))
            self.data = torch.full((*res, channels), float(props["value"]), dtype=torch.float32)
        elif "filename" in props:                       # Mitsuba .vol grid (experiments.py:56,62)
            from .volio import read_vol
            data, _, _ = read_vol(props["filename"])
            if data.shape[-1] != channels:
                raise ValueError(f"{props['filename']}: {data.shape[-1]} channels, expected {channels}")
            self.data = _as_tensor(data, channels)
        else:
            raise ValueError("gridvolume needs `data` (tensor), `value` or `filename` (.vol)")


class HeterogeneousMedium:
    """`heterogeneous` medium (experiments.py:52-71): sigma_t = scale * grid, albedo grid, phase."""

    def __init__(self, props: Properties):
        self.sigma_t = GridVolume(Properties(props["sigma_t"]), 1)
        self.albedo = GridVolume(Properties(props["albedo"]), 3)
        self._check_resolution()
        self.scale = float(props.get("scale", 1.0))
        self.sample_emitters = bool(props.get("sample_emitters", True))
        phase = props.get("phase", {"type": "isotropic"})
        self.phase = phase.get("type", "isotropic")
        self.phase_g = float(phase.get("g", 0.0))
        if props.get("has_spectral_extinction", False):
            raise ValueError("spectrally varying extinction is not supported (utils.py:66)")
        # heterogeneous `majorant_resolution_factor` (experiments.py:70,187,306): > 1 = majorant supergrid + DDA
        self._majorant_factor = int(props.get("majorant_resolution_factor", 0))
        if self._majorant_factor <= 1:
            self._majorant_factor = 0

    def is_twostate(self) -> bool:
        return False

    def volume_boxes(self):
        return [self.sigma_t.box, self.albedo.box]

    def majorant_resolution_factor(self) -> int:
        return self._majorant_factor

    def set_majorant_resolution_factor(self, factor: int):
        self._majorant_factor = int(factor) if int(factor) > 1 else 0

    def has_majorant_grid(self) -> bool:
        return self._majorant_factor > 0

    def parameters(self) -> Dict[str, torch.Tensor]:
        return {"sigma_t.data": self.sigma_t.data, "albedo.data": self.albedo.data}

    def _check_resolution(self):
        # the packed two-state grid keeps sigma_t and albedo voxel for voxel; the reference lets them differ
        # (janga: 264x136x136 sigma_t, 256x128x128 albedo, experiments.py:53-63) — refuse that here, not in the kernels
        if tuple(self.sigma_t.data.shape[:3]) != tuple(self.albedo.data.shape[:3]):
            raise ValueError(f"sigma_t {tuple(self.sigma_t.data.shape[:3])} and albedo {tuple(self.albedo.data.shape[:3])} "
                             "grids must share one resolution (resample one of them before loading)")

    def set_parameter(self, key: str, value: torch.Tensor):
        vol = self.sigma_t if key.startswith("sigma_t") else self.albedo
        vol.data = _as_tensor(value, vol.data.shape[-1])

    # what the renderer reads after a parameter update
    def incoming_medium(self) -> "HeterogeneousMedium":
        return self

    def parameters_changed(self, keys=None):
        pass


def _make_medium(d: dict):
    kind = d.get("type")
    if kind == "heterogeneous":
        return HeterogeneousMedium(Properties(d))
    if kind in _MEDIA:
        props = Properties({k: (_make_medium(v) if isinstance(v, dict) and "type" in v else v) for k, v in d.items()})
        return _MEDIA[kind](props)
    raise ValueError(f"unsupported medium type {kind!r}")


class Film:
    def __init__(self, width: int, height: int):
        self.width, self.height = int(width), int(height)

    def set_size(self, size):
        """scene_params['sensor.film.size'] = size (utils.py:117-118,125-126)."""
        self.width, self.height = int(size[0]), int(size[1])

    def size(self):
        return (self.width, self.height)


class Sensor:
    """`perspective`, or `batch` of perspective views side by side on one film (experiments.py:28-50)."""

    def __init__(self, d: dict):
        views = None
        if d.get("type") == "batch":
            subs = [v for k, v in sorted(d.items()) if isinstance(v, dict) and v.get("type") == "perspective"]
            if not subs:
                raise ValueError("batch sensor without perspective sub-sensors")
            if len({float(v.get("fov", 45.0)) for v in subs}) != 1:
                raise ValueError("the views of a batch sensor must share one field of view")
            views = [np.asarray(v["to_world"], dtype=np.float64) for v in subs]
            d = {**subs[0], "film": d.get("film", {})}
        elif d.get("type") != "perspective":
            raise ValueError("only `perspective` and `batch` sensors are supported")
        film = d.get("film", {})
        # hdrfilm: Mitsuba's default reconstruction filter is `gaussian` (stddev 0.5); `box` must be asked for
        rf = film.get("rfilter", {"type": "gaussian"})
        self.rfilter = rf.get("type", "gaussian")
        self.rfilter_stddev = float(rf.get("stddev", 0.5))
        if self.rfilter not in ("box", "gaussian"):
            raise ValueError("only the `box` and `gaussian` reconstruction filters are supported")
        self._film = Film(film.get("width", 768), film.get("height", 576))
        self.fov = float(d.get("fov", 45.0))
        self.to_world = np.asarray(d.get("to_world", look_at([0.5, 0.5, -2.0], [0.5, 0.5, 0.5], [0, 1, 0])), dtype=np.float64)
        if views is not None:
            self.to_world = np.stack(views)           # (n, 4, 4): SceneConfig lays the views out side by side
        self.near_clip = float(d.get("near_clip", 1e-2))

    def film(self) -> Film:
        return self._film


def _axis_aligned_box(to_world, lo, hi, what):
    """Image of the box [lo, hi]^3 under an axis-aligned `to_world` (scale + translation); anything else is refused."""
    m = np.asarray(to_world, dtype=np.float64)
    if m.shape != (4, 4) or np.abs(m[:3, :3] - np.diag(np.diag(m[:3, :3]))).max() > 1e-12 or np.any(np.diag(m[:3, :3]) == 0):
        raise ValueError(f"{what}: only axis-aligned `to_world` transforms (scale + translation) are supported")
    a, b = m[:3, :3] @ np.full(3, lo) + m[:3, 3], m[:3, :3] @ np.full(3, hi) + m[:3, 3]
    return tuple(float(v) for v in np.minimum(a, b)), tuple(float(v) for v in np.maximum(a, b))


class Shape:
    """The medium container.  Mitsuba's `cube` spans [-1,1]^3 and the reference's `cube_unit.obj` [0,1]^3 before
    `to_world` (experiments.py:73-80); without a `to_world` this shim keeps the unit cube [0,1]^3 of the synthetic
    scenes for every type (or the explicit `bbox_min` / `bbox_max` keys, which are not Mitsuba's)."""

    def __init__(self, d: dict):
        if d.get("type") not in ("cube", "obj", "box"):
            raise ValueError("the medium container must be a box (`cube`)")
        if d.get("bsdf", {"type": "null"}).get("type") != "null":
            raise ValueError("the medium boundary must use a `null` BSDF (volpathsimple.py:24-30)")
        if "to_world" in d:
            if "bbox_min" in d or "bbox_max" in d:
                raise ValueError("give the container either `to_world` or `bbox_min`/`bbox_max`, not both")
            lo, hi = (-1.0, 1.0) if d.get("type") == "cube" else (0.0, 1.0)
            self.bbox_min, self.bbox_max = _axis_aligned_box(d["to_world"], lo, hi, "medium container")
        else:
            self.bbox_min = tuple(float(v) for v in d.get("bbox_min", (0.0, 0.0, 0.0)))
            self.bbox_max = tuple(float(v) for v in d.get("bbox_max", (1.0, 1.0, 1.0)))
        self._interior = _make_medium(d["interior"]) if "interior" in d else None
        if self._interior is not None:                 # the grids must fill the container (use_grid_bbox = False)
            for vol_box in self._interior.volume_boxes():
                if vol_box is not None and not np.allclose(vol_box, (self.bbox_min, self.bbox_max), atol=1e-6):
                    raise ValueError("a gridvolume's `to_world` must map [0,1]^3 onto the medium container")

    def interior_medium(self):
        return self._interior

    def exterior_medium(self):
        return None


class Scene:
    """The result of `load_dict({'type': 'scene', ...})`."""

    def __init__(self, d: dict):
        self._shapes, self._sensors, self.env_radiance = [], [], None
        self.shape_key = None
        self.envmap, self.env_scale, self.env_to_world, self.quads = None, 1.0, None, []
        n_emitters = 0
        for key, v in d.items():
            if not isinstance(v, dict):
                continue
            t = v.get("type")
            if t in ("cube", "obj", "box"):
                self._shapes.append(Shape(v)); self.shape_key = key
            elif t in ("perspective", "batch"):
                self._sensors.append(Sensor(v))
            elif t == "constant":
                n_emitters += 1
                rad = v.get("radiance", 1.0)
                if isinstance(rad, dict):
                    rad = rad.get("value", 1.0)
                self.env_radiance = tuple(float(x) for x in (rad if np.ndim(rad) else (rad, rad, rad)))
            elif t == "envmap":                       # experiments.py:81-85
                n_emitters += 1
                if "bitmap" in v:                     # in-memory (H, W, 3) texels (this shim's extension)
                    tex = np.asarray(v["bitmap"], dtype=np.float32)
                else:
                    from .exrio import read_exr
                    img, names = read_exr(v["filename"])          # channels in file order (EXR sorts them: B, G, R)
                    if not all(c in names for c in "RGB"):
                        raise ValueError(f"{v['filename']}: an RGB image is expected")
                    tex = np.stack([img[..., names.index(c)] for c in "RGB"], axis=-1).astype(np.float32)
                self.envmap = np.ascontiguousarray(tex)
                self.env_scale = float(v.get("scale", 1.0))
                tw = np.asarray(v.get("to_world", np.eye(4)), dtype=np.float64)
                self.env_to_world = tw[:3, :3]
                if np.abs(self.env_to_world @ self.env_to_world.T - np.eye(3)).max() > 1e-6:
                    raise ValueError("envmap: `to_world` must be a rotation")
                self.env_radiance = (0.0, 0.0, 0.0)
            elif t == "rectangle":                    # Mitsuba's rectangle: [-1,1]^2 in the xy plane, normal +z, under to_world
                tw = np.asarray(v.get("to_world", np.eye(4)), dtype=np.float64)
                p0 = tw[:3, :3] @ np.array([-1.0, -1.0, 0.0]) + tw[:3, 3]
                eu, ev = tw[:3, :3] @ np.array([2.0, 0.0, 0.0]), tw[:3, :3] @ np.array([0.0, 2.0, 0.0])
                if abs(float(eu @ ev)) > 1e-9 * float(np.linalg.norm(eu) * np.linalg.norm(ev)):
                    raise ValueError("rectangle: `to_world` must keep the edges perpendicular")
                bsdf = v.get("bsdf", {"type": "diffuse"})
                if bsdf.get("type") != "diffuse":
                    raise ValueError("rectangles carry a `diffuse` BSDF")
                refl = bsdf.get("reflectance", 0.5)
                refl = refl.get("value", 0.5) if isinstance(refl, dict) else refl
                q = dict(p0=p0, eu=eu, ev=ev, reflectance=tuple(float(x) for x in (refl if np.ndim(refl) else (refl,) * 3)))
                em = v.get("emitter")
                if em is not None:
                    if em.get("type") != "area":
                        raise ValueError("a shape's emitter must be an `area` light")
                    n_emitters += 1
                    rad = em.get("radiance", 1.0)
                    rad = rad.get("value", 1.0) if isinstance(rad, dict) else rad
                    q["radiance"] = tuple(float(x) for x in (rad if np.ndim(rad) else (rad,) * 3))
                    if self.env_radiance is None:
                        self.env_radiance = (0.0, 0.0, 0.0)
                self.quads.append(q)
            elif t == "area":
                raise ValueError("an `area` emitter belongs to a `rectangle` shape (its `emitter` entry)")
        if n_emitters > 1:
            raise ValueError("one emitter per scene: `constant`, `envmap` or one `area` light")
        if len(self._shapes) != 1 or self._shapes[0].interior_medium() is None:
            raise ValueError("expected a single shape with an interior medium (volpathsimple.py:16-19)")
        if len(self._sensors) != 1:
            raise ValueError("only scenes with a single sensor are supported (optimize.py:36)")
        if self.env_radiance is None:
            raise ValueError("the scene needs an emitter: `constant`, `envmap` or an `area` light on a rectangle")
        self._ctx: Optional[Context] = None
        self._ctx_key = None
        # rotations requested before the device context exists; the scene file loads the same
        # volume into old, new and incoming (experiments.py:89-95), i.e. two rotations of it
        m = self.medium().incoming_medium()
        self._pending = [(m.sigma_t.data, m.albedo.data)] * 2

    def shapes(self):
        return self._shapes

    def sensors(self):
        return self._sensors

    def medium(self):
        return self._shapes[0].interior_medium()

    # ---- device state -------------------------------------------------------
    def scene_config(self, max_depth: int, rr_depth: int, hide_emitters: bool = False, use_nee: bool = True) -> SceneConfig:
        s, sh, m = self._sensors[0], self._shapes[0], self.medium().incoming_medium()
        return SceneConfig(width=s.film().width, height=s.film().height, fov=s.fov, to_world=s.to_world, near_clip=s.near_clip,
                           bbox_min=sh.bbox_min, bbox_max=sh.bbox_max, sigma_t_scale=m.scale, phase=m.phase, phase_g=m.phase_g,
                           sample_emitters=m.sample_emitters, env_radiance=self.env_radiance, max_depth=max_depth,
                           rr_depth=rr_depth, hide_emitters=hide_emitters, use_nee=use_nee, rfilter=s.rfilter,
                           rfilter_stddev=s.rfilter_stddev, envmap=self.envmap, env_scale=self.env_scale,
                           env_to_world=self.env_to_world, quads=tuple(self.quads))

    def context(self, max_depth: int, rr_depth: int, hide_emitters: bool = False, use_nee: bool = True) -> Context:
        """The rcv_ctx of this scene, configured for the calling integrator."""
        if self._ctx is None:
            self._ctx = Context()
        s = self._sensors[0]
        key = (max_depth, rr_depth, hide_emitters, use_nee, s.film().width, s.film().height)
        if key != self._ctx_key:
            self._ctx.set_scene(self.scene_config(max_depth, rr_depth, hide_emitters, use_nee))
            self._ctx_key = key
        if self._pending:
            pending, self._pending = self._pending[-2:], []      # only the last two states can matter
            for sig, alb in pending:
                self._rotate(sig, alb)
        self._sync_majorant_factor()
        return self._ctx

    def _sync_majorant_factor(self):
        """medium.set_majorant_resolution_factor (utils.py:41-43) takes effect on the device context here."""
        f = int(self.medium().majorant_resolution_factor())
        if f != getattr(self._ctx, "majorant_factor", 0):
            self._ctx.set_majorant_factor(f)

    def _rotate(self, sig: torch.Tensor, alb: torch.Tensor):
        self._sync_majorant_factor()
        dev = self._ctx.torch_device
        self._ctx.rotate(sig.detach().to(dev).contiguous(), alb.detach().to(dev).contiguous())

    def push_medium(self):
        """The incoming parameters become the new state: TwoStateMedium.parameters_changed
        (old <- new, new <- incoming) is one rcv_medium_rotate."""
        m = self.medium().incoming_medium()
        if self._ctx is None or self._ctx_key is None:
            self._pending.append((m.sigma_t.data, m.albedo.data))
            return
        self._rotate(m.sigma_t.data, m.albedo.data)


def load_dict(d: dict):
    """mi.load_dict for the supported plugin types."""
    t = d.get("type")
    if t == "scene":
        return Scene(d)
    if t in _INTEGRATORS:
        props = Properties({k: (load_dict(v) if isinstance(v, dict) and v.get("type") in _INTEGRATORS else v) for k, v in d.items()})
        return _INTEGRATORS[t](props)
    if t in ("heterogeneous",) or t in _MEDIA:
        return _make_medium(d)
    raise ValueError(f"load_dict: unsupported plugin type {t!r}")


class SceneParameters(dict):
    """mi.traverse(scene): differentiable parameters by key, e.g.
    'object.interior_medium.sigma_t.data' (experiments.py:104-112)."""

    def __init__(self, scene: Scene):
        super().__init__()
        self.scene = scene
        self.prefix = f"{scene.shape_key}.interior_medium."
        for k, v in scene.medium().incoming_medium().parameters().items():
            super().__setitem__(self.prefix + k, v)
        self._dirty = set()

    def __setitem__(self, key, value):
        if key not in self:
            raise KeyError(key)
        super().__setitem__(key, value)
        self._dirty.add(key)

    def update(self, values: Optional[dict] = None):
        """SceneParameters.update: assign, then notify the medium (plugins/twostatemedium.py:189-200)."""
        if values is not None:
            for k, v in values.items():
                if k in self:
                    self[k] = v
        med = self.scene.medium()
        inc = med.incoming_medium()
        for k in sorted(self._dirty):
            inc.set_parameter(k[len(self.prefix):], self[k])
        keys, self._dirty = list(self._dirty), set()
        med.parameters_changed(keys)
        self.scene.push_medium()
        return keys


def traverse(obj) -> SceneParameters:
    if isinstance(obj, Scene):
        return SceneParameters(obj)
    raise TypeError("traverse() expects a Scene")


# ----------------------------------------------------------------------------
# mi.render with the _RenderOp AD boundary
# ----------------------------------------------------------------------------
class _RenderOp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scene, integrator, params, keys, sensor, seed, spp, seed_grad, spp_grad, *tensors):
        ctx.args = (scene, integrator, params, keys, sensor, seed_grad, spp_grad)
        with torch.no_grad():
            # a fresh tensor per call, like the reference: the CV integrators update their `img` buffer in place
            return integrator.render(scene, sensor, seed, spp).clone()

    @staticmethod
    def backward(ctx, grad_out):
        scene, integrator, params, keys, sensor, seed_grad, spp_grad = ctx.args
        grads = integrator.render_backward(scene, params, grad_out.contiguous(), sensor, seed_grad, spp_grad)
        out = [None] * 9
        for k in keys:
            out.append(grads.get(k) if grads else None)
        return tuple(out)


def render(scene: Scene, params: Optional[SceneParameters] = None, sensor: int = 0, integrator=None, seed: int = 0,
           seed_grad: int = 0, spp: int = 0, spp_grad: int = 0) -> torch.Tensor:
    """mi.render: primal image now, adjoint on `.backward()` (optimize.py:108,122).
    seed_grad defaults to sample_tea_32(seed, 1)[0], spp_grad to spp."""
    assert integrator is not None, "pass integrator=..."
    if spp_grad == 0:
        spp_grad = spp
    if seed_grad == 0:
        seed_grad = tea32(seed, 1)[0]
    keys, tensors = [], []
    if params is not None:
        for k, v in params.items():
            if isinstance(v, torch.Tensor) and v.requires_grad:
                keys.append(k); tensors.append(v)
    if not tensors:
        with torch.no_grad():
            return integrator.render(scene, sensor, seed, spp).clone()
    return _RenderOp.apply(scene, integrator, params, keys, sensor, seed, spp, seed_grad, spp_grad, *tensors)


# ----------------------------------------------------------------------------
# mi.ad.Adam
# ----------------------------------------------------------------------------
class Adam:
    """mi.ad.Adam (optimize.py:90,127): dict of variables, per-key learning rates,
    state reset when a variable changes shape (after utils.upsample)."""

    def __init__(self, lr: float, beta_1: float = 0.9, beta_2: float = 0.999, epsilon: float = 1e-8):
        self.lr_default, self.beta_1, self.beta_2, self.epsilon = float(lr), beta_1, beta_2, epsilon
        self.lr: Dict[str, float] = {}
        self.variables: Dict[str, torch.Tensor] = {}
        self.state: Dict[str, Any] = {}
        self.t: Dict[str, int] = {}

    def __contains__(self, key):
        return key in self.variables

    def __getitem__(self, key) -> torch.Tensor:
        return self.variables[key]

    def __setitem__(self, key, value: torch.Tensor):
        v = value.detach().clone().contiguous().float()
        if torch.cuda.is_available():
            v = v.cuda()
        v.requires_grad_(True)
        if key not in self.variables or self.variables[key].shape != v.shape:
            self.reset(key, v)
        self.variables[key] = v

    def keys(self):
        return self.variables.keys()

    def items(self):
        return self.variables.items()

    def reset(self, key, like=None):
        like = self.variables[key] if like is None else like
        self.state[key] = (torch.zeros_like(like, requires_grad=False), torch.zeros_like(like, requires_grad=False))
        self.t[key] = 0

    def set_learning_rate(self, lr):
        if isinstance(lr, dict):
            self.lr.update({k: float(v) for k, v in lr.items()})
        else:
            # mi.ad.Optimizer.set_learning_rate(float) replaces the default only: per-key rates set through the
            # dict form (e.g. experiments.py:111-116 albedo at lr_factor 2) survive optimize.py's schedule steps
            self.lr_default = float(lr)

    def learning_rate(self, key) -> float:
        return self.lr.get(key, self.lr_default)

    def step(self, ctx: Optional[Context] = None):
        for k, p in self.variables.items():
            g = p.grad
            if g is None:
                continue
            self.t[k] += 1
            m, v = self.state[k]
            c = ctx or _any_context()
            with torch.no_grad():
                c.adam_flat(p.data, g.contiguous(), m, v, self.t[k], self.learning_rate(k), self.beta_1, self.beta_2, self.epsilon)
            p.grad = None


_CTX_FALLBACK: Optional[Context] = None


def _any_context() -> Context:
    global _CTX_FALLBACK
    if _CTX_FALLBACK is None:
        _CTX_FALLBACK = Context()
    return _CTX_FALLBACK
