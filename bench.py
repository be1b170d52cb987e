#!/usr/bin/env python
"""bench.py — headline benchmark of the recursive-CV hot path on B200.

One "step" = one optimisation iteration of optimize.py:104-132 for method
cv_ps on synthetic data: state rotation + majorant, correlated two-state
render (theta_{i-1}, theta_i), film develop, recursive CV / Welford update,
L1 loss + image gradient, volpathsimple path-replay adjoint (pass 1 + pass 2)
scatter-adding into the sigma_t/albedo gradient grid, [all-reduce at N>1],
Adam + clamp.  Metric: Mpaths/s (paths = lanes traced by the three passes)
and ms per iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--scaling weak|strong]

Configurations (BASELINE.json `configs`):
    c2          (default, the headline) 256^3 cube, 512x512, 64 spp per GPU, box filter, one majorant, use_drt = False
    c1          64^3 cube, 128x128, 16 spp — the L2-resident configuration (roofline against the measured L2 gather peak)
    c1box/c2box the same cubes inside a diffuse box with an area light, as BASELINE.json words configs[0..1]: surface
                transport and the inherited prbvolpath adjoint (the environment-lit c1 / c2 use volpathsimple)
    refdefault  what the reference's volume experiments run (experiments.py:20-133): use_drt + sub-sampling + MIS,
                gaussian filter, majorant supergrid (factor 8), batch sensor of 8 views, HG g = 0.877 — C2-sized
    c3          plume 256^3, 1024x1024, 256 spp      (strong scaling: the spp of ONE image split over the GPUs)
    c4          dense 512^3, 1024x1024, 512 spp      (strong scaling)
N > 1 is launched by torchrun (one rank per GPU, NCCL): every rank renders a
disjoint range of the samples of each pixel (global PCG32 lane indices), the
film accumulators and the gradient grid are all-reduced before the CV update
and the Adam step.  --scaling weak (default for c1/c2/refdefault): `spp` per GPU;
--scaling strong (default for c3/c4): `spp` in total.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

METRIC = "Mpaths/s (two-state CV render + adjoint)"
UNIT = "Mpaths/s"
SEED0 = 17   # the reference discards iterations 0..16 when benchmarking (optimize.py:105)

CONFIGS = {
    # BASELINE.json configs[1]: 256^3 grid, 512x512, 64 spp, cv_ps (environment-lit variant)
    "c2": dict(grid=256, res=512, spp=64, kind="cube", scaling="weak"),
    "c1": dict(grid=64, res=128, spp=16, kind="cube", scaling="weak"),
    "refdefault": dict(grid=256, res=256, view_width=128, views=8, spp=64, kind="cube", scaling="weak", rfilter="gaussian",
                       hg=0.877, majorant_factor=8, use_drt=True),
    # BASELINE.json configs[0..1] AS WORDED: the cube in a diffuse box with an area light (surface transport, the
    # adjoint the two-state integrator inherits from upstream prbvolpath: RCV_ADJOINT_PRBVOLPATH)
    "c1box": dict(grid=64, res=128, spp=16, kind="cube", scaling="weak", box=True),
    "c2box": dict(grid=256, res=512, spp=64, kind="cube", scaling="weak", box=True),
    "c3": dict(grid=256, res=1024, spp=256, kind="plume", scaling="strong"),
    "c4": dict(grid=512, res=1024, spp=512, kind="dense", scaling="strong"),
    # BASELINE.json configs[4]: the full optimisation loop (optimize.run_opt through the plugin layer: mi.render, the
    # autograd boundary, recursive CV + Welford state, mi.ad.Adam on the device); --steps = iterations (default 200)
    "c5": dict(grid=256, res=1024, spp=128, kind="plume", scaling="strong", loop=True),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default 5; c5: iterations of the loop, default 200)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="weak: --spp samples per pixel per GPU; strong: --spp in total, split over the GPUs")
    ap.add_argument("--grid", type=int, default=None)
    ap.add_argument("--res", type=int, default=None, help="film height (and width of one view)")
    ap.add_argument("--view-width", type=int, default=None, help="width of one view (default: --res)")
    ap.add_argument("--views", type=int, default=None, help="batch sensor: views side by side on one film")
    ap.add_argument("--spp", type=int, default=None, help="samples per pixel (per GPU when --scaling weak)")
    ap.add_argument("--kind", default=None)
    ap.add_argument("--box", action="store_true", default=None, help="diffuse box + area light around the medium (prbvolpath adjoint)")
    ap.add_argument("--hg", type=float, default=None, help="Henyey-Greenstein g (0 = isotropic)")
    ap.add_argument("--majorant-factor", type=int, default=None, help="majorant supergrid factor of the measured step (0 = one majorant)")
    ap.add_argument("--tracer", type=int, default=3, help="3 pool (default), 0 wavefront, 1 simple, 2 batch")
    ap.add_argument("--use-drt", action="store_true", default=None,
                    help="adjoint with volpathsimple's use_drt/subsampling/mis = True (reference default); the headline is "
                         "use_drt=False, the configuration SURVEY.md 8(d) measures")
    ap.add_argument("--no-drt", action="store_true", help="skip the informational use_drt=True timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="N > 1, weak scaling: skip the strong-scaling leg")
    ap.add_argument("--no-baseline-method", action="store_true", help="skip the method=baseline timing")
    ap.add_argument("--rfilter", default=None, choices=["box", "gaussian"], help="reconstruction filter of the film (headline: box)")
    ap.add_argument("--supergrid", type=int, nargs="*", default=None,
                    help="also time the same iteration with majorant supergrids of these factors (none = skip); reported under "
                         "`supergrid` (8 = what the reference's utils.adjust_majorant_res_factor picks, 32 = what suits these kernels)")
    ap.add_argument("--cpu-sample-spp", type=int, default=0, help="spp range of the CPU sample (0 = auto: ~15 s of CPU work)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of one CPU sample")
    a = ap.parse_args()
    cfg = dict(rfilter="box", hg=0.0, views=1, view_width=None, majorant_factor=0, use_drt=False, box=False, loop=False)
    cfg.update(CONFIGS[a.config])
    for k, v in cfg.items():
        if getattr(a, k, None) is None:
            setattr(a, k, v)
    a.loop = bool(CONFIGS[a.config].get("loop", False))
    if a.steps is None:
        a.steps = 200 if a.loop else 5
    if a.view_width is None:
        a.view_width = a.res
    if a.supergrid is None:
        a.supergrid = [8, 32] if a.config == "c2" else []
    if a.box:
        a.use_drt = False           # volpathsimple's DRT term does not exist in the inherited adjoint
    return a


def spp_layout(a, n_gpus):
    """(spp_total, [(begin, end) per rank])."""
    from recursive_control_variates_b200.parallel import shard_spp
    spp_total = a.spp * n_gpus if a.scaling == "weak" else a.spp
    return spp_total, [shard_spp(spp_total, r, n_gpus) for r in range(n_gpus)]


def workload_name(a, n_gpus):
    spp_total, ranges = spp_layout(a, n_gpus)
    per = ranges[0][1] - ranges[0][0]
    film = f"{a.view_width * a.views}x{a.res}" + (f" ({a.views} views)" if a.views > 1 else "")
    phase = "isotropic" if a.hg == 0.0 else f"HG g={a.hg}"
    maj = "one majorant" if not a.majorant_factor else f"majorant supergrid factor {a.majorant_factor}"
    light = "in a diffuse box [-1,2]x[-1,2]x[-2.5,2] (reflectance 0.5) with a 1x1 area light (radiance 10)" if a.box else \
        "constant environment emitter"
    adj = "adjoint prbvolpath (inherited)" if a.box else f"adjoint volpathsimple use_drt={bool(a.use_drt)}"
    return (f"synthetic {a.kind} {a.grid}^3 sigma_t+albedo two-state grid, {film}, {spp_total} spp "
            f"({per}/GPU), cv_ps, {light}, {a.rfilter} filter, {phase}, {maj}, max_depth=rr_depth=64, {adj}")


def config_keys(a, n_gpus):
    """The keys both arms print under `config` (the driver compares them)."""
    spp_total, ranges = spp_layout(a, n_gpus)
    W, H = a.view_width * a.views, a.res
    return {"workload": workload_name(a, n_gpus), "name": a.config, "grid": a.grid, "film": [W, H], "spp_total": spp_total,
            "spp_per_gpu": ranges[0][1] - ranges[0][0], "paths_per_step": W * H * spp_total * 3,
            "parallelism": f"spp-shard x{n_gpus}", "rfilter": a.rfilter, "use_drt": bool(a.use_drt),
            "majorant_factor": a.majorant_factor, "views": a.views, "phase_g": a.hg, "box": bool(a.box)}


# ----------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md): SM clock + throttle reasons during the timed region
# ----------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def l2_gather_peak():
    """Random 32-byte-sector gathers over an L2-resident array (scripts/microbench/gather_bw.cu, committed result)."""
    p = os.path.join(ROOT, "profiles", "l2_gather_peak.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["l2_gather_gbs"]), j.get("source", "profiles/l2_gather_peak.json")
    return None, None


def scene_config(a):
    """The bench scene: SURVEY.md 8(d)'s camera / medium box; the configuration decides filter, phase function and views."""
    from recursive_control_variates_b200.scene import SceneConfig, look_at
    kw = dict(width=a.view_width * a.views, height=a.res, rfilter=a.rfilter)
    if a.box:
        from recursive_control_variates_b200.synthetic import boxed_quads
        kw["quads"] = boxed_quads()
    if a.hg != 0.0:
        kw.update(phase="hg", phase_g=a.hg)
    if a.views > 1:           # the reference's batch sensor (experiments.py:28-50): cameras on a circle around the medium
        import math
        kw["to_world"] = np.stack([look_at([0.5 + 2 * math.cos(2 * math.pi / a.views * i + 0.1), 0.5,
                                            0.5 + 2 * math.sin(2 * math.pi / a.views * i + 0.1)], [0.5, 0.5, 0.5], [0, 1, 0])
                                   for i in range(a.views)])
    return SceneConfig(**kw)


def adjoint_kind(a):
    """RCV_ADJOINT_*: volpathsimple for the environment-lit configurations (experiments.py:131), the inherited prbvolpath
    adjoint where the scene has surfaces / an area light (volpathsimple.py:24-30 refuses those)."""
    return 1 if a.box else 0


def rel_l2(x, y):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


# ----------------------------------------------------------------------------
# CPU arm: the oracle (CPU restatement of the reference path) on the host cores
# ----------------------------------------------------------------------------
def cpu_iteration_sample(a, n_gpus, sample_spp, threads, keep=None, use_drt=None):
    """Times one bounded sample of the step on the CPU: samples [0, sample_spp) of every pixel
    for the two-state render and for both adjoint passes, on the same grids/film/seeds.
    keep: a dict that receives the oracle's outputs for that sample range (the parity block compares the GPU with them)."""
    import oracle as O
    from recursive_control_variates_b200.runtime import tea32
    from recursive_control_variates_b200.synthetic import make_medium
    O.build()
    desc = scene_config(a).to_desc()
    old, new = make_medium(a.grid, a.kind, 0.0), make_medium(a.grid, a.kind, 5e-3)
    mo = O.OracleMedium(desc, *old, majorant_factor=a.majorant_factor)
    mn = O.OracleMedium(desc, *new, majorant_factor=a.majorant_factor)
    spp, _ = spp_layout(a, n_gpus)
    W, H = a.view_width * a.views, a.res
    npix = W * H
    seed = SEED0
    seed_grad = tea32(seed, 1)[0]
    grad_img = np.full((H, W, 3), 1.0 / (npix * 3), np.float32)
    film_w = O.film_weights(desc, seed_grad, spp) if a.rfilter != "box" else np.full((H, W), float(spp), np.float32)
    drt = bool(a.use_drt) if use_drt is None else use_drt
    t0 = time.perf_counter()
    acc = O.render_twostates_accum(desc, mo, mn, seed, spp, 0, sample_spp, threads=threads)
    t1 = time.perf_counter()
    if a.box:
        grads = O.prb_backward(desc, mn, seed_grad, spp, grad_img, 0, sample_spp, threads=threads, film_w=film_w)
    else:
        grads = O.vps_backward(desc, mn, seed_grad, spp, grad_img, 0, sample_spp, threads=threads, film_w=film_w, use_drt=drt)
    t2 = time.perf_counter()
    if keep is not None:
        keep.update(acc_old=acc[0], acc_new=acc[1], g_sigma=grads[0], g_albedo=grads[1], grad_img=grad_img, film_w=film_w,
                    sample_spp=sample_spp, seed=seed, seed_grad=seed_grad, old=old, new=new, use_drt=drt)
    paths = npix * sample_spp * 3
    return paths, t2 - t0, t1 - t0, t2 - t1


def calibrated_sample_spp(a, n_gpus, threads, budget_s, keep=None):
    """Size the CPU sample: time samples [0,1) of every pixel, then pick the spp range that takes ~budget_s."""
    spp_total, _ = spp_layout(a, n_gpus)
    if a.cpu_sample_spp:
        return min(a.cpu_sample_spp, spp_total)
    _, t1, _, _ = cpu_iteration_sample(a, n_gpus, 1, threads, keep=keep)
    return int(max(1, min(spp_total, round(budget_s / max(t1, 1e-3)))))


def run_reference(a):
    """--impl reference: the reference cannot be installed (no Mitsuba 3 / Dr.Jit wheel);
    its CPU restatement (oracle/, kind 'port') is timed on all host cores instead."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_gpus = a.gpus
    import oracle as O
    threads = O.default_threads()
    # the whole --steps/--warmup run should end within a few minutes
    budget = min(a.cpu_seconds, 150.0 / max(1, a.steps + min(a.warmup, 1)))
    sample_spp = calibrated_sample_spp(a, n_gpus, threads, budget)      # doubles as the warm-up pass
    times, paths = [], 0
    for _ in range(a.steps):
        paths, t, _, _ = cpu_iteration_sample(a, n_gpus, sample_spp, threads)
        times.append(t)
    t = float(np.mean(times))
    v = paths / t / 1e6
    spp_total, _ = spp_layout(a, n_gpus)
    sample = f"samples [0,{sample_spp}) of {spp_total} spp of every pixel, all three passes, same grids/film/seeds"
    cfg = config_keys(a, n_gpus)
    out = {"impl": "reference", "note": "CPU restatement of the reference (oracle/), bounded sample per step; ms_per_step is for the sample", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def parity_block(a, ctx, dev, keep1, keepN, keep_drt, spp_total):
    """GPU against the oracle at the MEASURED configuration: the same grids, film and seeds, sample range [0, n) of the
    spp the bench runs (n = what the CPU leg could afford).  Images and gradients as relative L2; samples [0, 1) — one
    lane per pixel under the box filter — compared bit for bit."""
    import torch
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    k = keepN
    ctx.set_majorant_factor(a.majorant_factor)
    ctx.set_drt(bool(k["use_drt"]))
    ctx.rotate(cu(k["old"][0]), cu(k["old"][1])); ctx.rotate(cu(k["new"][0]), cu(k["new"][1]))
    out = {"config": f"{a.config}: the measured grids / film / seeds, samples [0,{k['sample_spp']}) of {spp_total} spp of every pixel "
                     f"({int(np.prod(k['acc_old'].shape[:2])) * k['sample_spp'] * 3} paths), use_drt={bool(k['use_drt'])}",
           "tolerance": "images 1e-4, gradients 1e-3 relative L2 (north_star); lanes bit-exact"}
    dev_img = lambda acc: (acc[..., :3] / np.where(acc[..., 3:] == 0, 1, acc[..., 3:]))
    if keep1 is not None:
        g_old, g_new = ctx.render_twostates_accum(k["seed"], spp_total, 0, 1)
        eq = np.array_equal(g_old.cpu().numpy(), keep1["acc_old"]) and np.array_equal(g_new.cpu().numpy(), keep1["acc_new"])
        out["lanes_bit_exact"] = bool(eq) if a.rfilter == "box" else None
        out["lanes_compared"] = int(np.prod(keep1["acc_old"].shape[:2]))
    g_old, g_new = ctx.render_twostates_accum(k["seed"], spp_total, 0, k["sample_spp"])
    out["img_rel_l2_old"] = rel_l2(dev_img(g_old.cpu().numpy()), dev_img(k["acc_old"]))
    out["img_rel_l2_new"] = rel_l2(dev_img(g_new.cpu().numpy()), dev_img(k["acc_new"]))
    gi, fw = cu(k["grad_img"]), cu(k["film_w"])
    ctx.grad_zero()
    ctx.render_backward(k["seed_grad"], spp_total, gi, 0, k["sample_spp"], film_w=fw, adjoint=adjoint_kind(a))
    gs, ga = ctx.grad_read()
    out["grad_sigma_rel_l2"] = rel_l2(gs.cpu().numpy(), k["g_sigma"])
    out["grad_albedo_rel_l2"] = rel_l2(ga.cpu().numpy(), k["g_albedo"])
    ok = (out["img_rel_l2_old"] < 1e-4 and out["img_rel_l2_new"] < 1e-4 and out["grad_sigma_rel_l2"] < 1e-3 and
          out["grad_albedo_rel_l2"] < 1e-3 and out.get("lanes_bit_exact") in (True, None))
    if keep_drt:
        ctx.set_drt(bool(keep_drt["use_drt"]))
        ctx.grad_zero()
        ctx.render_backward(keep_drt["seed_grad"], spp_total, gi, 0, keep_drt["sample_spp"], film_w=fw)
        gs, ga = ctx.grad_read()
        key = "drt" if keep_drt["use_drt"] else "no_drt"
        out[key] = {"samples": f"[0,{keep_drt['sample_spp']})", "use_drt": bool(keep_drt["use_drt"]),
                    "grad_sigma_rel_l2": rel_l2(gs.cpu().numpy(), keep_drt["g_sigma"]),
                    "grad_albedo_rel_l2": rel_l2(ga.cpu().numpy(), keep_drt["g_albedo"]),
                    "note": "use_drt / sub-sampling / MIS on; Medium::sample_interaction_drt is restated from the paper (unpinned)"
                    if keep_drt["use_drt"] else "the same adjoint without the DRT term"}
        ok = ok and out[key]["grad_sigma_rel_l2"] < 1e-3 and out[key]["grad_albedo_rel_l2"] < 1e-3
    ctx.set_drt(bool(a.use_drt))
    out["pass"] = bool(ok)
    return out


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from recursive_control_variates_b200 import _abi
    from recursive_control_variates_b200.runtime import Context, tea32
    from recursive_control_variates_b200.synthetic import make_medium

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    dev = torch.device("cuda", local)

    W, H = a.view_width * a.views, a.res
    spp_total, ranges = spp_layout(a, n_gpus)
    spp_b, spp_e = ranges[rank]
    spp_local = spp_e - spp_b
    desc = scene_config(a).to_desc()
    ctx = Context(local)
    ctx.set_scene(desc)
    ctx.set_tracer(a.tracer)
    ctx.set_drt(bool(a.use_drt))

    # synthetic theta_{i-1}, theta_i and a target (generated on the host once; resident in HBM before timing)
    old, new = make_medium(a.grid, a.kind, 0.0), make_medium(a.grid, a.kind, 5e-3)
    tgt = make_medium(a.grid, a.kind, 5e-2)
    sigma = torch.from_numpy(new[0]).to(dev); albedo = torch.from_numpy(new[1]).to(dev)
    ctx.rotate(torch.from_numpy(tgt[0]).to(dev), torch.from_numpy(tgt[1]).to(dev))
    ref = ctx.develop(ctx.render_primal_accum(1, 17843, min(spp_total, 64)))   # utils.py:98-111 reference render
    ctx.rotate(torch.from_numpy(old[0]).to(dev), torch.from_numpy(old[1]).to(dev))
    ctx.set_majorant_factor(a.majorant_factor)
    m_s, v_s = torch.zeros_like(sigma), torch.zeros_like(sigma)
    m_a, v_a = torch.zeros_like(albedo), torch.zeros_like(albedo)
    acc = (torch.zeros(H, W, 4, device=dev), torch.zeros(H, W, 4, device=dev))
    Himg, Fimg = torch.zeros(H, W, 3, device=dev), torch.zeros(H, W, 3, device=dev)
    cv = ctx.new_cv_state()
    film_w = torch.zeros(H, W, device=dev)
    state = {"it": 0, "stats_n": 0, "t": 0}
    npix = W * H
    paths_per_step = npix * spp_total * 3          # primal + adjoint pass 1 + adjoint pass 2, all ranks
    side = torch.cuda.Stream(device=dev) if world > 1 else None

    lay = {"total": spp_total, "b": spp_b, "e": spp_e}       # the sample range this rank renders (the strong-scaling leg swaps it)

    def step(i, marks=None, method="cv_ps"):
        spp_total, spp_b, spp_e = lay["total"], lay["b"], lay["e"]

        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
        seed = SEED0 + i
        seed_grad = tea32(seed, 1)[0]
        mark("start")
        ctx.rotate(sigma, albedo)                                            # scene_params.update(opt)
        mark("rotate")
        fw = None
        if world > 1:
            # film weights of the adjoint's samples (a pure function of seed_grad): own samples on a side stream while the
            # two-state render runs, one 1 MB all-reduce
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                film_w.zero_()
                fw = ctx.film_weights(seed_grad, spp_total, spp_b, spp_e, out=film_w)
                dist.all_reduce(fw)
        if method == "cv_ps":
            acc[0].zero_(); acc[1].zero_()
            ctx.render_twostates_accum(seed, spp_total, spp_b, spp_e, out=acc)   # mi.render -> render_twostates
            mark("twostate_render")
            if world > 1:
                dist.all_reduce(acc[0]); dist.all_reduce(acc[1])
                mark("allreduce_film")
            ctx.develop(acc[0], Himg); ctx.develop(acc[1], Fimg)
            ctx.cv_update(Himg, Fimg, cv, state["it"], 1, state["stats_n"], 0.9, 0.999)
            state["stats_n"] += state["it"] > 0
            state["it"] += 1
            img = cv["img"]
        else:                                                                # method = baseline: prbvolpath primal of theta_i
            acc[1].zero_()
            ctx.render_primal_accum(1, seed, spp_total, spp_b, spp_e, out=acc[1])
            mark("primal_render")
            if world > 1:
                dist.all_reduce(acc[1])
                mark("allreduce_film")
            img = ctx.develop(acc[1], Fimg)
        loss, g = ctx.loss_grad(_abi.RCV_LOSS_L1, img, ref)                  # l1_loss + dr.backward seed
        ctx.grad_zero()
        mark("develop_cv_loss")
        if world > 1:
            torch.cuda.current_stream().wait_stream(side)
        ctx.render_backward(seed_grad, spp_total, g, spp_b, spp_e, film_w=fw, adjoint=adjoint_kind(a))   # adjoint_integrator.render_backward
        mark("adjoint")
        if world > 1:
            dist.all_reduce(ctx.grad_packed())
            mark("allreduce_grad")
        state["t"] += 1
        ctx.adam_step(sigma, albedo, m_s, v_s, m_a, v_a, state["t"], 5e-3, 1e-2)   # opt.step() + clamp
        mark("adam_clamp")
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, first, **kw):
        """n steps starting at iteration `first`: (ms per step (max over ranks), counters per step)."""
        barrier(); ctx.counters(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n):
            step(first + i, **kw)
        s1.record(); barrier()
        c = ctx.counters(True).astype(np.float64) / n
        t_ = torch.tensor([s0.elapsed_time(s1) / n], device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item()), c

    for i in range(a.warmup):
        step(i)
    barrier()
    ctx.counters(True)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(a.steps):
        loss = step(a.warmup + i)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    cnt = ctx.counters(True)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / a.steps
    value = paths_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- where a step spends its time on this rank (CUDA events between the phases of two extra steps)
    phases = {}
    for r in range(2):
        marks = []
        step(a.warmup + a.steps + 100 + r, marks)
        torch.cuda.synchronize()
        for (_, e_prev), (name, e_cur) in zip(marks[:-1], marks[1:]):
            phases[name] = phases.get(name, 0.0) + e_prev.elapsed_time(e_cur) / 2
    barrier()

    # ---- per-kernel timing (CUDA events on the launching stream) + roofline of the dominant kernel
    peak, peak_src = measured_peaks()

    def time_kernel(fn, reps=3):
        fn(); torch.cuda.synchronize(); ctx.counters(True)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            fn()
        a1.record(); torch.cuda.synchronize()
        c = ctx.counters(True).astype(np.float64) / reps
        return a0.elapsed_time(a1) / reps, c

    g_img = torch.full((H, W, 3), 1.0 / (npix * 3), device=dev)
    fw_all = ctx.film_weights(tea32(SEED0, 1)[0], spp_total)               # weights of ALL samples, computed once outside the timing
    k1_ms, k1_c = time_kernel(lambda: ctx.render_twostates_accum(SEED0, spp_total, spp_b, spp_e, out=acc))
    k6_ms, k6_c = time_kernel(lambda: ctx.render_backward(tea32(SEED0, 1)[0], spp_total, g_img, spp_b, spp_e, film_w=fw_all,
                                                          adjoint=adjoint_kind(a)))
    lanes = npix * spp_local
    # algorithmic bytes (SURVEY.md §8d / DESIGN.md §5)
    k1_bytes = k1_c[0] * 64 + k1_c[1] * 192 + lanes * (16 + 32)
    k6_bytes = k6_c[2] * 32 + k6_c[3] * (96 + 32) + k6_c[4] * 32 + k6_c[5] * 96 + lanes * 2 * 16
    kernels = {
        "twostate_primal": {"ms": k1_ms, "algorithmic_gb": k1_bytes / 1e9, "gbs": k1_bytes / k1_ms / 1e6,
                            "collisions": k1_c[0], "scatters": k1_c[1], "mpaths_s": lanes / k1_ms / 1e3},
        ("prbvolpath_adjoint" if a.box else "volpathsimple_adjoint"): {"ms": k6_ms, "algorithmic_gb": k6_bytes / 1e9, "gbs": k6_bytes / k6_ms / 1e6,
                                  "collisions": k6_c[2], "scatters": k6_c[3], "grad_sites": k6_c[4] + k6_c[5],
                                  "mpaths_s": 2 * lanes / k6_ms / 1e3},
    }
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this workload (per launch)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and a.config == "c2" and (a.grid, a.res, a.spp, a.kind) == (256, 512, 64, "cube") and n_gpus == 1:
        with open(tpath) as f:
            tj = json.load(f)
        if dom in tj:
            traffic, traffic_src = tj[dom]["dram_bytes"] / 1e9, tj[dom]["capture"]
    # the grids of the L2-resident configurations fit the 126 MB L2: their roofline is the measured L2 gather peak
    working_set_mb = (a.grid + 1) ** 3 * (64 + 32 + 32) / 2 ** 20 + a.grid ** 3 * (32 + 20) / 2 ** 20
    l2_peak, l2_src = l2_gather_peak()
    l2_bound = working_set_mb < 100 and l2_peak is not None
    for kv in kernels.values():
        kv["frac_of_hbm_peak"] = kv["gbs"] / peak
        if l2_peak:
            kv["frac_of_l2_gather_peak"] = kv["gbs"] / l2_peak
    roofline = {"bound": "l2" if l2_bound else "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"],
                "peak": l2_peak if l2_bound else peak, "unit": "GB/s",
                "frac": kernels[dom]["gbs"] / (l2_peak if l2_bound else peak), "traffic": traffic, "traffic_unit": "GB per launch",
                "traffic_source": traffic_src, "algorithmic_gb": kernels[dom]["algorithmic_gb"],
                "peak_source": l2_src if l2_bound else peak_src, "share_of_step": kernels[dom]["ms"] / ms_per_step,
                "working_set_mb": working_set_mb}

    # ---- method = baseline beside cv_ps (BASELINE.json configs[1]: "cv_ps vs baseline method"): the same iteration with a
    # single-state primal render of theta_i and no CV update
    baseline_method = None
    if not a.no_baseline_method:
        step(a.warmup + a.steps, method="baseline")
        b_ms, _ = timed(max(2, min(a.steps, 3)), a.warmup + a.steps + 1, method="baseline")
        baseline_method = {"ms_per_step": b_ms, "value": paths_per_step / (b_ms * 1e-3) / 1e6, "unit": UNIT,
                           "cv_ps_over_baseline_time": ms_per_step / b_ms,
                           "equal_variance_note": "cv_ps lowers the variance of the image the optimiser sees 4.5x at equal spp "
                                                  "(profiles/r38_convergence.json, scripts/demo_convergence.py, C1 recovery): the "
                                                  "baseline needs ~4.5x the samples for the same variance"}

    # ---- N > 1: the same image with the spp of ONE job split over the GPUs (north_star's sharding; the headline above is
    # weak scaling: `spp` per GPU)
    strong = None
    if world > 1 and a.scaling == "weak" and not a.no_strong:
        s_total = a.spp
        s_b, s_e = spp_layout(argparse.Namespace(spp=a.spp, scaling="strong"), n_gpus)[1][rank]
        lay.update(total=s_total, b=s_b, e=s_e)
        for i in range(2):
            step(a.warmup + a.steps + i)
        s_ms, _ = timed(max(3, min(a.steps, 5)), a.warmup + a.steps + 2)
        s_marks = []
        step(a.warmup + a.steps + 50, s_marks)
        torch.cuda.synchronize()
        s_ph = {name: e_prev.elapsed_time(e_cur) for (_, e_prev), (name, e_cur) in zip(s_marks[:-1], s_marks[1:])}
        strong = {"spp_total": s_total, "spp_per_gpu": s_e - s_b, "ms_per_step": s_ms,
                  "value": npix * s_total * 3 / (s_ms * 1e-3) / 1e6, "unit": UNIT, "phases_ms_rank0": s_ph,
                  "note": "strong scaling: the %d spp of one %s image split over %d GPUs (same step, same kernels)" % (s_total, a.config, n_gpus)}
        lay.update(total=spp_total, b=spp_b, e=spp_e)

    # ---- the same iteration with the majorant supergrid + DDA (reference `use_majorant_supergrid`), informational
    supergrid = []
    for sg_f in [f for f in a.supergrid if f > 1 and a.grid % f == 0 and f != a.majorant_factor]:
        ctx.set_majorant_factor(sg_f)
        step(a.warmup + a.steps)
        sg_ms, sg_cnt = timed(max(2, min(a.steps, 3)), a.warmup + a.steps + 1)
        supergrid.append({"factor": sg_f, "ms_per_step": sg_ms, "value": paths_per_step / (sg_ms * 1e-3) / 1e6, "unit": UNIT,
                          "twostate_collisions": sg_cnt[0], "adjoint_collisions": sg_cnt[2],
                          "note": "same workload, free flights sampled by DDA through the majorant supergrid (not the headline config)"})
    ctx.set_majorant_factor(a.majorant_factor)

    # ---- the same iteration with the adjoint's DRT term (volpathsimple.py:39-47: use_drt / use_drt_subsampling /
    # use_drt_mis all True, the reference's default), informational unless the configuration measures it
    drt_info = None
    if not a.use_drt and not a.no_drt and not a.box:
        ctx.set_drt(True)
        for i in range(2):
            step(a.warmup + a.steps + i)
        d_ms, d_cnt = timed(max(2, min(a.steps, 3)), a.warmup + a.steps + 1)
        drt_info = {"ms_per_step": d_ms, "value": paths_per_step / (d_ms * 1e-3) / 1e6, "unit": UNIT,
                    "adjoint_collisions": d_cnt[2], "gpu_launches": d_cnt[7],
                    "note": "same workload with the adjoint's differential-ratio-tracking term (use_drt, sub-sampling and MIS "
                            "on: the reference default); paths_per_step does not count the DRT estimate's recursive path"}
        ctx.set_drt(False)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    e2e, e2e_dev = None, None
    pin = lambda t_: t_.detach().cpu().contiguous().pin_memory()
    if not a.no_e2e and world > 1:
        # N > 1: the same iteration through the Python API with pinned HOST buffers on every rank — parameters and target
        # up, shard renders + NCCL all-reduces on the device, image / loss / gradients down
        hs, ha, href = pin(sigma), pin(albedo), pin(ref)
        himg, hloss = torch.zeros(H, W, 3).pin_memory(), torch.zeros(1).pin_memory()
        hgs, hga = torch.zeros_like(hs).pin_memory(), torch.zeros_like(ha).pin_memory()
        d_s, d_a, d_ref = torch.empty_like(sigma), torch.empty_like(albedo), torch.empty_like(ref)
        cv2 = ctx.new_cv_state()
        st2 = {"it": 0, "n": 0}

        def e2e_step(i):
            seed = SEED0 + i
            # the host-side optimiser lives on rank 0: its buffers go up once and reach the other ranks over NVLink
            if rank == 0:
                d_s.copy_(hs, non_blocking=True); d_a.copy_(ha, non_blocking=True); d_ref.copy_(href, non_blocking=True)
            dist.broadcast(d_s, 0); dist.broadcast(d_a, 0); dist.broadcast(d_ref, 0)
            ctx.rotate(d_s, d_a)
            acc[0].zero_(); acc[1].zero_()
            ctx.render_twostates_accum(seed, spp_total, spp_b, spp_e, out=acc)
            dist.all_reduce(acc[0]); dist.all_reduce(acc[1])
            ctx.develop(acc[0], Himg); ctx.develop(acc[1], Fimg)
            ctx.cv_update(Himg, Fimg, cv2, st2["it"], 1, st2["n"], 0.9, 0.999)
            st2["n"] += st2["it"] > 0; st2["it"] += 1
            l_, g_ = ctx.loss_grad(_abi.RCV_LOSS_L1, cv2["img"], d_ref)
            ctx.grad_zero()
            film_w.zero_()
            fw_ = ctx.film_weights(tea32(seed, 1)[0], spp_total, spp_b, spp_e, out=film_w)
            dist.all_reduce(fw_)
            ctx.render_backward(tea32(seed, 1)[0], spp_total, g_, spp_b, spp_e, film_w=fw_, adjoint=adjoint_kind(a))
            dist.all_reduce(ctx.grad_packed())
            if rank == 0:
                gs_, ga_ = ctx.grad_read()
                hgs.copy_(gs_, non_blocking=True); hga.copy_(ga_, non_blocking=True)
                himg.copy_(cv2["img"], non_blocking=True); hloss.copy_(l_.reshape(1), non_blocking=True)
            torch.cuda.synchronize()

        e2e_step(0); e2e_step(1)
        barrier()
        n_e2e = max(2, min(a.steps, 3))
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_step(2 + i)
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        h2d = (hs.numel() + ha.numel() + href.numel()) * 4
        d2h = (himg.numel() + 1 + hgs.numel() + hga.numel()) * 4
        e2e = {"value": paths_per_step / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "api": "Context API with pinned host buffers on rank 0 (parameters + target up, NCCL broadcast to the "
               "other ranks; image, loss and all-reduced gradients down from rank 0)"}
    if not a.no_e2e and world == 1:
        hs, ha, href = pin(sigma), pin(albedo), pin(ref)
        himg, hloss = torch.zeros(H, W, 3).pin_memory(), torch.zeros(1).pin_memory()
        hgs, hga = torch.zeros_like(hs).pin_memory(), torch.zeros_like(ha).pin_memory()
        ctx.set_majorant_factor(a.majorant_factor)

        def e2e_step(i, reset=False):
            seed = SEED0 + i
            ctx.cv_iteration_host(hs, ha, href, seed, spp_total, tea32(seed, 1)[0], spp_total, himg, hloss, hgs, hga, reset=reset,
                                  adjoint=adjoint_kind(a))

        e2e_step(0, reset=True); e2e_step(1)
        torch.cuda.synchronize()
        n_e2e = max(2, min(a.steps, 3))
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_step(2 + i)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        h2d = (hs.numel() + ha.numel() + href.numel()) * 4
        d2h = (himg.numel() + 1 + hgs.numel() + hga.numel()) * 4
        e2e = {"value": paths_per_step / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "api": "rcv_cv_iteration_host (render_twostates + cv update + loss + adjoint; "
               "host-side optimiser: parameters and target up, image, loss and gradients down, every step)"}
    if not a.no_e2e:
        # the plugin-API shape of the loop (optimize.py:104-132): the optimiser state lives on the device, as in the
        # reference; per step the host sends the target image and reads the image and the loss back
        href2 = pin(ref)
        himg2, hloss2 = torch.zeros(H, W, 3).pin_memory(), torch.zeros(1).pin_memory()

        def dev_step(i):
            ref.copy_(href2, non_blocking=True)
            l_ = step(1000 + i)
            himg2.copy_(cv["img"], non_blocking=True); hloss2.copy_(l_.reshape(1), non_blocking=True)
            torch.cuda.synchronize()

        dev_step(0); dev_step(1)
        barrier()
        n_e2e = max(2, min(a.steps, 3))
        t0 = time.perf_counter()
        for i in range(n_e2e):
            dev_step(2 + i)
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e_dev = {"value": paths_per_step / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3,
                   "h2d_bytes_per_step": href2.numel() * 4, "d2h_bytes_per_step": (himg2.numel() + 1) * 4,
                   "api": "Context API, Adam + clamp on the device (rcv_adam_step) as the reference keeps its optimiser; the host "
                          "sends the target image and reads image + loss every step, wall clock with a device sync per step"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the host cores, bounded sample; its outputs pin the GPU's
    cpu, parity = None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        import oracle as O
        threads = O.default_threads()
        keep1, keepN, keepD = {}, {}, {}
        sample_spp = calibrated_sample_spp(a, n_gpus, threads, a.cpu_seconds, keep=keep1)
        paths, tt, t_primal, t_adj = cpu_iteration_sample(a, n_gpus, sample_spp, threads, keep=keepN)
        cpu = {"value": paths / tt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"samples [0,{sample_spp}) of {spp_total} spp of every pixel ({paths} paths): two-state render "
                         f"{t_primal:.2f}s + volpathsimple adjoint {t_adj:.2f}s",
               "seconds": tt}
        # the adjoint with the other setting of use_drt on a third of that range
        if not a.box:
            cpu_iteration_sample(a, n_gpus, max(1, sample_spp // 3), threads, keep=keepD, use_drt=not a.use_drt)
        parity = parity_block(a, ctx, dev, keep1 if keep1 else None, keepN, keepD if keepD else None, spp_total)

    if rank == 0:
        cfg = config_keys(a, n_gpus)          # the same keys and values as the reference arm prints
        run_info = {"l2": "inputs larger than L2 (two-state grids + bricks %d MiB): no flush needed" % working_set_mb if not l2_bound else
                    "L2-resident working set (%d MiB): the roofline is the measured L2 gather peak" % working_set_mb,
                    "tracer": {0: "wavefront", 1: "simple", 2: "batch", 3: "pool"}[a.tracer], "loss_last": float(loss.item())}
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": cfg, "run": run_info,
               "clocks": clocks, "gpu_launches": int(cnt[7]), "roofline": roofline, "kernels": kernels,
               "phases_ms_rank0": phases, "e2e": e2e, "e2e_device_optimizer": e2e_dev, "cpu_baseline": cpu, "parity": parity,
               "baseline_method": baseline_method, "strong": strong, "supergrid": supergrid, "drt": drt_info}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_loop(a):
    """--config c5: `--steps` iterations of optimize.run_opt (optimize.py:104-143 of the reference, unchanged in
    structure) through the plugin layer, spp-sharded over the ranks: twostateprbvolpath inside `meta` (method cv_ps),
    volpathsimple as adjoint_integrator, L1 loss, mi.ad.Adam + clamp on the device.  The whole loop is timed (wall clock,
    device synchronised, max over ranks); the reference's own per-iteration stopwatch (benchmark = True: iterations
    > 16) is reported beside it."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from recursive_control_variates_b200 import experiments
    from recursive_control_variates_b200 import mitsuba_compat as mi
    from recursive_control_variates_b200.optimize import run_opt
    n_steps = a.steps
    spp_total, ranges = spp_layout(a, world)
    p = experiments.scene_synthetic("cv_ps", None, grid=a.grid, init_grid=a.grid, film=(a.res, a.res), spp=spp_total, kind=a.kind,
                                    n_steps=n_steps, upsample=())
    p["recomp_freq"] = max(n_steps, 1)            # one checkpoint render (iteration 0), as a run with recomp_freq = n_steps
    p["spp_ref"], p["ref_passes"], p["spp_inf"] = spp_total, 1, min(spp_total, 64)
    p["adjoint_integrator"] = mi.load_dict({"type": "volpathsimple", "max_depth": 64, "rr_depth": 64, "use_drt": bool(a.use_drt)})
    sampler = ClockSampler(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    barrier()
    sampler.start()
    t0 = time.perf_counter()
    res = run_opt(p)
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    te = torch.tensor([dt], device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    dt = float(te.item())
    if rank == 0:
        W = H = a.res
        paths = W * H * spp_total * 3
        rt = res["runtime"][17:] if n_steps > 17 else res["runtime"][:0]
        cfg = config_keys(a, world)
        cfg["workload"] = (f"full optimisation loop: {n_steps} iterations of optimize.run_opt (cv_ps, L1, Adam lr 5e-3, clamp), "
                           f"{a.kind} {a.grid}^3 grids from a constant start, {W}x{H}, {spp_total} spp ({ranges[0][1] - ranges[0][0]}/GPU), "
                           f"adjoint volpathsimple use_drt={bool(a.use_drt)}; includes scene set-up, the reference render and one checkpoint render")
        out = {"metric": METRIC, "value": paths * n_steps / dt / 1e6, "unit": UNIT, "n_gpus": world, "steps": n_steps, "warmup": 0,
               "ms_per_step": dt / max(n_steps, 1) * 1e3, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
               "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": clocks, "loop_seconds": dt,
               "stopwatch_ms": {"render": float(rt[:, 0].mean()) if len(rt) else None,
                                "backward_step": float(rt[:, 1].mean()) if len(rt) else None,
                                "note": "optimize.py's own Stopwatch (iterations > 16): mi.render + loss, then backward + Adam + clamp"},
               "loss_at_iteration_0": float(res["loss"][0]),
               "gpu_launches": int(mi._any_context().counters(False)[7])}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.loop:
        run_loop(args)
    else:
        run_ours(args)
