"""Runs the synthetic C1 recovery (64^3 grids, 128x128, 16 spp) with the three methods of the reference's
run_experiment.py and prints what optimize.py's checkpoint() records: L1 loss of a 64-spp render against the target and
the variance of the image the optimiser saw (`var` = mean((img_64spp - img)^2)), plus the parameter error at the end.
    python scripts/demo_convergence.py [n_steps] [out.json]        (needs a GPU)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recursive_control_variates_b200 import experiments  # noqa: E402
from recursive_control_variates_b200.optimize import run_opt  # noqa: E402
from recursive_control_variates_b200.synthetic import make_medium  # noqa: E402

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
out = {}
tgt_s, tgt_a = make_medium(64, "cube", 0.0)
for method in ("baseline", "cv_ps", "cv_pss"):
    p = experiments.scene_synthetic_c1(method, None)
    p["n_steps"] = n_steps
    p["recomp_freq"] = 10
    res = run_opt(p)
    s = res["opt"]["object.interior_medium.sigma_t.data"].detach().cpu().numpy()
    n = n_steps // 10
    out[method] = {"loss": [float(v) for v in res["loss"][:n]], "var": [float(v) for v in res["var"][:n]],
                   "sigma_t_rmse": float(np.sqrt(np.mean((s - tgt_s) ** 2))),
                   "ms_per_iteration": float(res["runtime"][17:].sum(axis=1).mean())}
    print(method, "loss", ["%.4f" % v for v in out[method]["loss"][-5:]], "var", ["%.2e" % v for v in out[method]["var"][-5:]],
          "sigma_t rmse %.4f" % out[method]["sigma_t_rmse"], "ms/it %.2f" % out[method]["ms_per_iteration"])
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
