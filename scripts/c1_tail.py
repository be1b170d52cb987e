#!/usr/bin/env python
"""Critical path of the C1 two-state render: tentative collisions (dependent grid gathers) per lane, from the oracle.
A lane's collisions are serial — each free-flight sample starts where the previous collision was — so the longest lane
bounds the kernel time from below whatever the scheduling: n_max x (one dependent L2 round trip + the step's arithmetic)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from recursive_control_variates_b200.scene import SceneConfig
from recursive_control_variates_b200.synthetic import make_medium

grid, res, spp = 64, 128, 16
desc = SceneConfig(width=res, height=res).to_desc()
old, new = make_medium(grid, "cube", 0.0), make_medium(grid, "cube", 5e-3)
mo, mn = O.OracleMedium(desc, *old), O.OracleMedium(desc, *new)
n = res * res * spp
step = int(sys.argv[1]) if len(sys.argv) > 1 else 1
per = []
for i in range(0, n, step):
    O.counters(True)
    O.twostates_lanes(desc, mo, mn, 17, spp, i, i + 1, threads=1)
    c = O.counters(True)
    per.append((int(c[0]), int(c[1])))
per = np.array(per)
coll = per[:, 0]
out = {"lanes": int(len(per)), "collisions_total": int(coll.sum()), "collisions_mean": float(coll.mean()),
       "collisions_p50": float(np.percentile(coll, 50)), "collisions_p99": float(np.percentile(coll, 99)),
       "collisions_p9999": float(np.percentile(coll, 99.99)), "collisions_max": int(coll.max()),
       "scatters_max": int(per[:, 1].max()), "lanes_with_zero": int((coll == 0).sum())}
print(json.dumps(out))
