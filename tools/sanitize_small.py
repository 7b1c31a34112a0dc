"""Small all-paths run for compute-sanitizer (memcheck / racecheck): highway branch, quadruped prox, robust chain."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
from _bmpc import abi, batch, scenarios  # noqa: E402

B = 8
x0, z0, xref, pp = scenarios.highway_batch(B, seed=3)
for ctrl, mode in ((abi.CTRL_BRANCH, abi.SLAB_SHARED), (abi.CTRL_BRANCH, abi.SLAB_SPLIT), (abi.CTRL_BRANCH, abi.SLAB_GLOBAL),
                   (abi.CTRL_ROBUST, abi.SLAB_SHARED)):
    cfg = scenarios.highway_config(batch_capacity=B, max_iter=40)
    cfg.controller, cfg.slab_mode = ctrl, mode
    mpc = batch.BatchedBranchMPC(cfg)
    x, z = x0.copy(), z0.copy()
    for s in range(2):
        r = mpc.solve_host(x, z, xref, pp)
        x = scenarios.euler_highway(x, r["u0"])
    print("highway ctrl", ctrl, "mode", mpc.launch_info()["slab_mode"], "status", r["status"].tolist(), flush=True)
    mpc.close()
q0, qz, qr = scenarios.quadruped_batch(4, seed=2)
mpc = batch.BatchedBranchMPC(scenarios.quadruped_config(batch_capacity=4, max_iter=30))
r = mpc.solve_host(q0, qz, qr)
print("quadruped prox status", r["status"].tolist(), mpc.launch_info()["slab_mode"], flush=True)
mpc.close()
