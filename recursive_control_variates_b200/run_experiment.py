"""run_experiment.py of the reference: CLI around run_opt, one (experiment, method) job
after the other; results go to <output>/<experiment>/<method>/result.npz."""
from __future__ import annotations

import argparse
import os

import numpy as np

from . import experiments
from .optimize import run_opt

AVAILABLE_METHODS = ["baseline", "cv_ps", "cv_pss"]


def run_experiment(config):
    base_dir = os.path.join(experiments.OUTPUT_DIR, config.get("output", config["experiment"]))
    output_dir = os.path.join(base_dir, config["method"])
    params = getattr(experiments, f"scene_{config['experiment']}")(config["method"], output_dir)
    for key, value in config.items():
        if key not in ("device", "experiment", "method", "output"):
            params[key] = value
    result = run_opt(params)
    result.pop("opt", None)
    os.makedirs(output_dir, exist_ok=True)
    np.savez(os.path.join(output_dir, "result.npz"), **result)
    return result


def main(argv=None):
    parser = argparse.ArgumentParser(description="Run an experiment on the B200 recursive-CV path")
    parser.add_argument("experiment", type=str, nargs="*", choices=experiments.AVAILABLE_SCENES)
    parser.add_argument("--method", type=str, choices=AVAILABLE_METHODS)
    parser.add_argument("--output", type=str)
    for name, typ in (("n_steps", int), ("lr", float), ("spp", int), ("spp_grad", int), ("beta1", float), ("beta2", float)):
        parser.add_argument(f"--{name}", type=typ)
    parser.add_argument("--loss", type=str, choices=["L1", "L2"])
    parser.add_argument("--pre_update", action="store_true", default=None)
    parser.add_argument("--force_baseline", action="store_true", default=None)
    parser.add_argument("--all", action="store_true", help="also run cv_pss (run_experiment.py:63: not run by default)")
    args = parser.parse_args(argv)
    print(f"{'Experiment':^20} | {'Method':^20}")
    for experiment in args.experiment:
        default_methods = AVAILABLE_METHODS if args.all else [m for m in AVAILABLE_METHODS if m != "cv_pss"]
        for method in (default_methods if args.method is None else [args.method]):
            cfg = {"experiment": experiment, "method": method,
                   **{k: v for k, v in vars(args).items() if v is not None and k not in ("experiment", "method", "all")}}
            res = run_experiment(cfg)
            print(f"{experiment:^20} | {method:^20} | final loss {res['loss'][-1]:.5f}")


if __name__ == "__main__":
    main()
