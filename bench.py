#!/usr/bin/env python
"""bench.py — headline benchmark of the recursive-CV hot path on B200.

One "step" = one optimisation iteration of optimize.py:104-132 for method
cv_ps on synthetic data: state rotation + majorant, correlated two-state
render (theta_{i-1}, theta_i), film develop, recursive CV / Welford update,
L1 loss + image gradient, volpathsimple path-replay adjoint (pass 1 + pass 2)
scatter-adding into the sigma_t/albedo gradient grid, [all-reduce at N>1],
Adam + clamp.  Metric: Mpaths/s (paths = lanes traced by the three passes)
and ms per iteration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU, NCCL): every rank renders a
disjoint range of the samples of each pixel (global PCG32 lane indices), the
film accumulators and the gradient grid are all-reduced before the CV update
and the Adam step.  Weak scaling: 64 spp per GPU.
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
# BASELINE.json configs[1]: 256^3 grid, 512x512, 64 spp, cv_ps (environment-lit variant)
WORKLOAD = dict(grid=256, width=512, height=512, spp_per_gpu=64, kind="cube")
SEED0 = 17   # the reference discards iterations 0..16 when benchmarking (optimize.py:105)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=WORKLOAD["grid"])
    ap.add_argument("--res", type=int, default=WORKLOAD["width"])
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp_per_gpu"], help="samples per pixel per GPU")
    ap.add_argument("--kind", default=WORKLOAD["kind"])
    ap.add_argument("--tracer", type=int, default=3, help="3 pool (default), 0 wavefront, 1 simple, 2 batch")
    ap.add_argument("--use-drt", action="store_true",
                    help="adjoint with volpathsimple's use_drt/subsampling/mis = True (reference default); the headline is "
                         "use_drt=False, the configuration SURVEY.md 8(d) measures")
    ap.add_argument("--no-drt", action="store_true", help="skip the informational use_drt=True timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--rfilter", default="box", choices=["box", "gaussian"], help="reconstruction filter of the film (headline: box)")
    ap.add_argument("--supergrid", type=int, nargs="*", default=[8, 32],
                    help="also time the same iteration with majorant supergrids of these factors (none = skip); reported under "
                         "`supergrid` (8 = what the reference's utils.adjust_majorant_res_factor picks, 32 = what suits these kernels)")
    ap.add_argument("--cpu-sample-spp", type=int, default=0, help="spp range of the CPU sample (0 = auto: ~15 s of CPU work)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target duration of one CPU sample")
    return ap.parse_args()


def workload_name(a, n_gpus):
    return (f"synthetic {a.kind} {a.grid}^3 sigma_t+albedo two-state grid, {a.res}x{a.res}, {a.spp * n_gpus} spp "
            f"({a.spp}/GPU), cv_ps, constant environment emitter, box filter, max_depth=rr_depth=64, "
            f"adjoint volpathsimple use_drt={bool(getattr(a, 'use_drt', False))}")


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


# ----------------------------------------------------------------------------
# CPU arm: the oracle (CPU restatement of the reference path) on the host cores
# ----------------------------------------------------------------------------
def cpu_iteration_sample(a, n_gpus, sample_spp, threads, keep=None, use_drt=False):
    """Times one bounded sample of the step on the CPU: samples [0, sample_spp) of every pixel
    for the two-state render and for both adjoint passes, on the same grids/film/seeds.
    keep: a dict that receives the oracle's outputs for that sample range (the parity block compares the GPU with them)."""
    import oracle as O
    from recursive_control_variates_b200.runtime import tea32
    from recursive_control_variates_b200.scene import SceneConfig
    from recursive_control_variates_b200.synthetic import make_medium
    O.build()
    desc = scene_config(a).to_desc()
    old, new = make_medium(a.grid, a.kind, 0.0), make_medium(a.grid, a.kind, 5e-3)
    mo, mn = O.OracleMedium(desc, *old, majorant_factor=a.majorant_factor), O.OracleMedium(desc, *new, majorant_factor=a.majorant_factor)
    spp = a.spp * n_gpus
    npix = a.res * a.res * a.views
    seed = SEED0
    seed_grad = tea32(seed, 1)[0]
    grad_img = np.full((a.res, a.res * a.views, 3), 1.0 / (npix * 3), np.float32)
    film_w = O.film_weights(desc, seed_grad, spp) if a.rfilter != "box" else np.full((a.res, a.res * a.views), float(spp), np.float32)
    t0 = time.perf_counter()
    acc = O.render_twostates_accum(desc, mo, mn, seed, spp, 0, sample_spp, threads=threads)
    t1 = time.perf_counter()
    grads = O.vps_backward(desc, mn, seed_grad, spp, grad_img, 0, sample_spp, threads=threads, film_w=film_w, use_drt=use_drt)
    t2 = time.perf_counter()
    if keep is not None:
        keep.update(acc_old=acc[0], acc_new=acc[1], g_sigma=grads[0], g_albedo=grads[1], grad_img=grad_img, film_w=film_w,
                    sample_spp=sample_spp, seed=seed, seed_grad=seed_grad, old=old, new=new)
    paths = npix * sample_spp * 3
    return paths, t2 - t0, t1 - t0, t2 - t1


def scene_config(a):
    """The bench scene: SURVEY.md 8(d)'s camera / medium box, `--config` decides filter, phase function and views."""
    from recursive_control_variates_b200.scene import SceneConfig, look_at
    kw = dict(width=a.res * a.views, height=a.res, rfilter=a.rfilter)
    if a.hg != 0.0:
        kw.update(phase="hg", phase_g=a.hg)
    if a.views > 1:           # the reference's batch sensor (experiments.py:28-50): cameras on a circle around the medium
        import math
        kw["to_world"] = np.stack([look_at([0.5 + 2 * math.cos(2 * math.pi / a.views * i + 0.1), 0.5,
                                            0.5 + 2 * math.sin(2 * math.pi / a.views * i + 0.1)], [0.5, 0.5, 0.5], [0, 1, 0])
                                   for i in range(a.views)])
    return SceneConfig(**kw)


def rel_l2(x, y):
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    return float(np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300))


def parity_block(a, ctx, dev, keep1, keepN, keep_drt, spp_total):
    """GPU against the oracle at the MEASURED configuration: the same grids, film and seeds, sample range [0, n) of the
    spp the bench runs (n = what the CPU leg could afford).  Images and gradients as relative L2; samples [0, 1) — one
    lane per pixel under the box filter — compared bit for bit."""
    import torch
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    k = keepN
    ctx.set_drt(False)
    ctx.rotate(cu(k["old"][0]), cu(k["old"][1])); ctx.rotate(cu(k["new"][0]), cu(k["new"][1]))
    out = {"config": f"{a.config}: the measured grids / film / seeds, samples [0,{k['sample_spp']}) of {spp_total} spp of every pixel",
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
    ctx.render_backward(k["seed_grad"], spp_total, gi, 0, k["sample_spp"], film_w=fw)
    gs, ga = ctx.grad_read()
    out["grad_sigma_rel_l2"] = rel_l2(gs.cpu().numpy(), k["g_sigma"])
    out["grad_albedo_rel_l2"] = rel_l2(ga.cpu().numpy(), k["g_albedo"])
    if keep_drt is not None:
        ctx.set_drt(True)
        ctx.grad_zero()
        ctx.render_backward(keep_drt["seed_grad"], spp_total, gi, 0, keep_drt["sample_spp"], film_w=fw)
        gs, ga = ctx.grad_read()
        out["drt"] = {"samples": f"[0,{keep_drt['sample_spp']})", "grad_sigma_rel_l2": rel_l2(gs.cpu().numpy(), keep_drt["g_sigma"]),
                      "grad_albedo_rel_l2": rel_l2(ga.cpu().numpy(), keep_drt["g_albedo"]),
                      "note": "use_drt / sub-sampling / MIS on; Medium::sample_interaction_drt is restated from the paper (unpinned)"}
        ctx.set_drt(a.use_drt)
    out["pass"] = bool(out["img_rel_l2_old"] < 1e-4 and out["img_rel_l2_new"] < 1e-4 and out["grad_sigma_rel_l2"] < 1e-3 and
                       out["grad_albedo_rel_l2"] < 1e-3 and out.get("lanes_bit_exact") in (True, None) and
                       (keep_drt is None or (out["drt"]["grad_sigma_rel_l2"] < 1e-3 and out["drt"]["grad_albedo_rel_l2"] < 1e-3)))
    return out


def calibrated_sample_spp(a, n_gpus, threads, budget_s):
    """Size the CPU sample: time samples [0,1) of every pixel, then pick the spp range that takes ~budget_s."""
    if a.cpu_sample_spp:
        return a.cpu_sample_spp
    _, t1, _, _ = cpu_iteration_sample(a, n_gpus, 1, threads)
    return int(max(1, min(a.spp * n_gpus, round(budget_s / max(t1, 1e-3)))))


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
    sample = f"samples [0,{sample_spp}) of {a.spp * n_gpus} spp of every pixel, all three passes, same grids/film/seeds"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(a, n_gpus), "note": "CPU restatement of the reference (oracle/), "
                      "bounded sample per step; ms_per_step is for the sample"},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    from recursive_control_variates_b200 import _abi
    from recursive_control_variates_b200.runtime import Context, tea32
    from recursive_control_variates_b200.scene import SceneConfig
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

    W = H = a.res
    spp_total = a.spp * n_gpus
    spp_b, spp_e = rank * a.spp, (rank + 1) * a.spp
    desc = SceneConfig(width=W, height=H, rfilter=a.rfilter).to_desc()
    ctx = Context(local)
    ctx.set_scene(desc)
    ctx.set_tracer(a.tracer)
    ctx.set_drt(a.use_drt)

    # synthetic theta_{i-1}, theta_i and a target (generated on the host once; resident in HBM before timing)
    old, new = make_medium(a.grid, a.kind, 0.0), make_medium(a.grid, a.kind, 5e-3)
    tgt = make_medium(a.grid, a.kind, 5e-2)
    sigma = torch.from_numpy(new[0]).to(dev); albedo = torch.from_numpy(new[1]).to(dev)
    ctx.rotate(torch.from_numpy(tgt[0]).to(dev), torch.from_numpy(tgt[1]).to(dev))
    ref = ctx.develop(ctx.render_primal_accum(1, 17843, min(spp_total, 64)))   # utils.py:98-111 reference render
    ctx.rotate(torch.from_numpy(old[0]).to(dev), torch.from_numpy(old[1]).to(dev))
    m_s, v_s = torch.zeros_like(sigma), torch.zeros_like(sigma)
    m_a, v_a = torch.zeros_like(albedo), torch.zeros_like(albedo)
    acc = (torch.zeros(H, W, 4, device=dev), torch.zeros(H, W, 4, device=dev))
    Himg, Fimg = torch.zeros(H, W, 3, device=dev), torch.zeros(H, W, 3, device=dev)
    cv = ctx.new_cv_state()
    film_w = torch.zeros(H, W, device=dev)
    state = {"it": 0, "stats_n": 0, "t": 0}
    npix = W * H
    paths_per_step = npix * a.spp * 3 * n_gpus     # primal + adjoint pass 1 + adjoint pass 2, all ranks

    def step(i, marks=None):
        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
        seed = SEED0 + i
        seed_grad = tea32(seed, 1)[0]
        mark("start")
        ctx.rotate(sigma, albedo)                                            # scene_params.update(opt)
        mark("rotate")
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
        loss, g = ctx.loss_grad(_abi.RCV_LOSS_L1, cv["img"], ref)           # l1_loss + dr.backward seed
        ctx.grad_zero()
        mark("develop_cv_loss")
        fw = None
        if world > 1:                                                        # film weights: own samples, then all-reduce
            film_w.zero_()
            fw = ctx.film_weights(seed_grad, spp_total, spp_b, spp_e, out=film_w)
            dist.all_reduce(fw)
            mark("film_weights")
        ctx.render_backward(seed_grad, spp_total, g, spp_b, spp_e, film_w=fw)   # adjoint_integrator.render_backward
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
    k1_ms, k1_c = time_kernel(lambda: ctx.render_twostates_accum(SEED0, spp_total, spp_b, spp_e, out=acc))
    k6_ms, k6_c = time_kernel(lambda: ctx.render_backward(tea32(SEED0, 1)[0], spp_total, g_img, spp_b, spp_e))
    lanes = npix * a.spp
    # algorithmic bytes (SURVEY.md §8d / DESIGN.md §5)
    k1_bytes = k1_c[0] * 64 + k1_c[1] * 192 + lanes * (16 + 32)
    k6_bytes = k6_c[2] * 32 + k6_c[3] * (96 + 32) + k6_c[4] * 32 + k6_c[5] * 96 + lanes * 2 * 16
    kernels = {
        "twostate_primal": {"ms": k1_ms, "algorithmic_gb": k1_bytes / 1e9, "gbs": k1_bytes / k1_ms / 1e6,
                            "collisions": k1_c[0], "scatters": k1_c[1], "mpaths_s": lanes / k1_ms / 1e3},
        "volpathsimple_adjoint": {"ms": k6_ms, "algorithmic_gb": k6_bytes / 1e9, "gbs": k6_bytes / k6_ms / 1e6,
                                  "collisions": k6_c[2], "scatters": k6_c[3], "grad_sites": k6_c[4] + k6_c[5],
                                  "mpaths_s": 2 * lanes / k6_ms / 1e3},
    }
    dom = max(kernels, key=lambda k: kernels[k]["ms"])
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture of this workload (per launch)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and (a.grid, a.res, a.spp, a.kind) == (256, 512, 64, "cube") and n_gpus == 1:
        with open(tpath) as f:
            tj = json.load(f)
        if dom in tj:
            traffic, traffic_src = tj[dom]["dram_bytes"] / 1e9, tj[dom]["capture"]
    # informational: the random-gather ceiling of this GPU (scripts/microbench/gather_bw.cu, 64-byte granules, 1 GiB array)
    gather_peak = None
    gpath = os.path.join(ROOT, "profiles", "r24_gather_bw_1GiB.jsonl")
    if os.path.exists(gpath):
        with open(gpath) as f:
            rows = [json.loads(l) for l in f if l.strip().startswith("{")]
        g64 = [r["GBps"] for r in rows if r.get("granule_bytes") == 64]
        gather_peak = max(g64) if g64 else None
    for kv in kernels.values():
        kv["frac_of_hbm_peak"] = kv["gbs"] / peak
        kv["frac_of_gather_peak"] = (kv["gbs"] / gather_peak) if gather_peak else None
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["gbs"] / peak, "traffic": traffic, "traffic_unit": "GB per launch",
                "traffic_source": traffic_src, "algorithmic_gb": kernels[dom]["algorithmic_gb"], "peak_source": peak_src,
                "share_of_step": kernels[dom]["ms"] / ms_per_step,
                "gather_peak": gather_peak, "gather_peak_source": "profiles/r24_gather_bw_1GiB.jsonl (random 64-B granules, useful GB/s)",
                "frac_of_gather_peak": (kernels[dom]["gbs"] / gather_peak) if gather_peak else None}

    # ---- the same iteration with the majorant supergrid + DDA (reference `use_majorant_supergrid`), informational
    supergrid = []
    for sg_f in [f for f in a.supergrid if f > 1 and a.grid % f == 0]:
        ctx.set_majorant_factor(sg_f)
        step(a.warmup + a.steps); barrier(); ctx.counters(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_sg = max(2, min(a.steps, 3))
        s0.record()
        for i in range(n_sg):
            step(a.warmup + a.steps + 1 + i)
        s1.record(); barrier()
        sg_cnt = ctx.counters(True).astype(np.float64) / n_sg
        tsg = torch.tensor([s0.elapsed_time(s1) / n_sg], device=dev)
        if world > 1:
            dist.all_reduce(tsg, op=dist.ReduceOp.MAX)
        sg_ms = float(tsg.item())
        supergrid.append({"factor": sg_f, "ms_per_step": sg_ms, "value": paths_per_step / (sg_ms * 1e-3) / 1e6, "unit": UNIT,
                          "twostate_collisions": sg_cnt[0], "adjoint_collisions": sg_cnt[2],
                          "note": "same workload, free flights sampled by DDA through the majorant supergrid (not the headline config)"})
    ctx.set_majorant_factor(0)

    # ---- the same iteration with the adjoint's DRT term (volpathsimple.py:39-47: use_drt / use_drt_subsampling /
    # use_drt_mis all True, the reference's default), informational unless --use-drt makes it the measured configuration
    drt_info = None
    if not a.use_drt and not a.no_drt:
        ctx.set_drt(True)
        for i in range(2):
            step(a.warmup + a.steps + i)
        barrier(); ctx.counters(True)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_d = max(2, min(a.steps, 3))
        s0.record()
        for i in range(n_d):
            step(a.warmup + a.steps + 1 + i)
        s1.record(); barrier()
        d_cnt = ctx.counters(True).astype(np.float64) / n_d
        td = torch.tensor([s0.elapsed_time(s1) / n_d], device=dev)
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        d_ms = float(td.item())
        drt_info = {"ms_per_step": d_ms, "value": paths_per_step / (d_ms * 1e-3) / 1e6, "unit": UNIT,
                    "adjoint_collisions": d_cnt[2], "gpu_launches": d_cnt[7],
                    "note": "same workload with the adjoint's differential-ratio-tracking term (use_drt, sub-sampling and MIS "
                            "on: the reference default); paths_per_step does not count the DRT estimate's recursive path"}
        ctx.set_drt(False)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    e2e = None
    if not a.no_e2e and world > 1:
        # N > 1: the same iteration through the Python API with pinned HOST buffers on every rank — parameters and target
        # up, shard renders + NCCL all-reduces on the device, image / loss / gradients down
        pin = lambda t_: t_.detach().cpu().contiguous().pin_memory()
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
            ctx.render_backward(tea32(seed, 1)[0], spp_total, g_, spp_b, spp_e, film_w=fw_)
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
        pin = lambda t_: t_.detach().cpu().contiguous().pin_memory()
        hs, ha, href = pin(sigma), pin(albedo), pin(ref)
        himg, hloss = torch.zeros(H, W, 3).pin_memory(), torch.zeros(1).pin_memory()
        hgs, hga = torch.zeros_like(hs).pin_memory(), torch.zeros_like(ha).pin_memory()

        def e2e_step(i, reset=False):
            seed = SEED0 + i
            ctx.cv_iteration_host(hs, ha, href, seed, spp_total, tea32(seed, 1)[0], spp_total, himg, hloss, hgs, hga, reset=reset)

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
        e2e = {"value": npix * a.spp * 3 / dt / 1e6, "unit": UNIT, "ms_per_step": dt * 1e3, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "api": "rcv_cv_iteration_host (render_twostates + cv update + loss + adjoint; "
               "host-side optimiser)"}

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        import oracle as O
        threads = O.default_threads()
        sample_spp = calibrated_sample_spp(a, n_gpus, threads, a.cpu_seconds)
        paths, tt, t_primal, t_adj = cpu_iteration_sample(a, n_gpus, sample_spp, threads)
        cpu = {"value": paths / tt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"samples [0,{sample_spp}) of {spp_total} spp of every pixel ({paths} paths): two-state render "
                         f"{t_primal:.2f}s + volpathsimple adjoint {t_adj:.2f}s",
               "seconds": tt}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload_name(a, n_gpus), "grid": a.grid, "film": [W, H], "spp_total": spp_total,
                          "spp_per_gpu": a.spp, "paths_per_step": paths_per_step, "parallelism": f"spp-shard x{n_gpus}",
                          "l2": "inputs larger than L2 (two-state grids %d MiB)" % (a.grid ** 3 * 40 // 2 ** 20),
                          "tracer": {0: "wavefront", 1: "simple", 2: "batch", 3: "pool"}[a.tracer], "rfilter": a.rfilter, "loss_last": float(loss.item())},
               "clocks": clocks, "gpu_launches": int(cnt[7]), "roofline": roofline, "kernels": kernels,
               "phases_ms_rank0": phases, "e2e": e2e, "cpu_baseline": cpu, "supergrid": supergrid, "drt": drt_info}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
