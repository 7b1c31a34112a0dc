"""Development aid: SHARED vs SPLIT slab placement for trees that leave one or two teams per SM."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch  # noqa: E402
from _bmpc import abi, batch, scenarios  # noqa: E402

dev = torch.device("cuda", 0)


def run(name, cfg, arrays, steps=6):
    mpc = batch.BatchedBranchMPC(cfg)
    t = [None if a is None else torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in arrays]
    ms = []
    for s in range(steps + 1):
        out = mpc.solve(t[0], t[1], t[2], t[3], outputs=("u0", "status"))
        torch.cuda.synchronize()
        ms.append(mpc.last_kernel_ms())
        mpc.plant_step(t[0], out["u0"], t[1], 0, t[3])
    st = np.bincount(out["status"].cpu().numpy(), minlength=4).tolist()
    print("%-28s %s warm %.0f solves/s  status %s" % (name, mpc.launch_info(), arrays[0].shape[0] / (np.mean(ms[1:]) * 1e-3), st), flush=True)
    mpc.close()


names = ["maintain", "brake", "lc", "trackv"]
for mode in (abi.SLAB_SHARED, abi.SLAB_SPLIT):
    for m, NB, B in ((3, 3, 7281), (4, 2, 7281), (4, 3, 7281)):
        x0, z0, xref, pp3 = scenarios.highway_batch(B, seed=1239 + 10 * m + NB)
        pp = np.zeros((B, m, 4))
        pp[:, 2, :] = pp3[:, 2, :]
        if m >= 4:
            pp[:, 3, 0] = 20.0
        run("m%d NB%d mode %d" % (m, NB, mode), scenarios.highway_config(policies=names[:m], NB=NB, batch_capacity=B, slab_mode=mode),
            (x0, z0, xref, pp))
    x0, z0, xref = scenarios.quadruped_batch(8192, seed=1238)
    run("quadruped mode %d" % mode, scenarios.quadruped_config(batch_capacity=8192, slab_mode=mode), (x0, z0, xref, None))
