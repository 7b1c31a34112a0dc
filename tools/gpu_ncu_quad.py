"""Development aid: a few closed-loop steps of one bench workload (for ncu captures of the other kernel instances)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch
import bench
which = sys.argv[1] if len(sys.argv) > 1 else "4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_BATCH.get(which, 0)
if which == "m3nb3":
    # the deepest three-policy tree of the sweep (40 branches, 313 input nodes): one team of 256 lanes per SM
    import numpy as np
    from _bmpc import batch, scenarios
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 7281
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=1239 + 33)
    mpc = batch.BatchedBranchMPC(scenarios.highway_config(NB=3, batch_capacity=B))
    t = [torch.as_tensor(np.ascontiguousarray(a), device="cuda") for a in (x0, z0, xref, pp)]
    parts = [{"mpc": mpc, "x": t[0], "z": t[1], "r": t[2], "p": t[3]}]
else:
    parts = bench.make_workload(which, B, 0, 1, 0)
for s in range(int(os.environ.get("STEPS", "5"))):
    for pt in parts:
        if "S" in pt:
            out = pt["mpc"].solve_transformed(pt["x"], pt["z"], pt["r"], pt["S"], pt["bd"], pt["p"])
        else:
            out = pt["mpc"].solve(pt["x"], pt["z"], pt["r"], pt["p"])
        pt["mpc"].plant_step(pt["x"], out["u0"], pt["z"], 0, pt["p"])
    torch.cuda.synchronize()
    print("step", s, [round(pt["mpc"].last_kernel_ms(), 3) for pt in parts], flush=True)
