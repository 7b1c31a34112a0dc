"""Development aid: a few closed-loop steps of one bench workload (for ncu captures of the other kernel instances)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch
import bench
which = sys.argv[1] if len(sys.argv) > 1 else "4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_BATCH[which]
parts = bench.make_workload(which, B, 0, 1, 0)
for s in range(int(os.environ.get("STEPS", "5"))):
    for pt in parts:
        out = pt["mpc"].solve(pt["x"], pt["z"], pt["r"], pt["p"])
        pt["mpc"].plant_step(pt["x"], out["u0"], pt["z"], 0, pt["p"])
    torch.cuda.synchronize()
    print("step", s, [round(pt["mpc"].last_kernel_ms(), 3) for pt in parts], flush=True)
