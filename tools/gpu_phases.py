"""Development aid: where does a solve spend its time?  Runs the same closed loop once per phase with reserved[7] = k, so
that the `cycles` output reports the time each problem spent in phase k (Solver::prof_begin), and prints the shares."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch  # noqa: E402
from _bmpc import batch, scenarios  # noqa: E402

NAMES = {0: "whole solve", 1: "interior point", 2: "tree expansion", 3: "rho selection / cache", 4: "factorisations",
         5: "backward + forward sweeps", 6: "node-parallel polish passes", 7: "adjoint sweep", 8: "ADMM row phase",
         9: "final pass + caches", 10: "  expansion: rollouts", 11: "  expansion: input staging", 12: "  expansion: node set-up",
         13: "  expansion: safety values, probabilities"}
KS = [int(v) for v in os.environ["PHASES"].split(",")] if os.environ.get("PHASES") else list(range(10))
kind = sys.argv[1] if len(sys.argv) > 1 else "hw"
B = int(os.environ.get("B", "16384" if kind == "hw" else "8192"))
steps = int(os.environ.get("STEPS", "5"))
dev = torch.device("cuda", 0)
rows = {}
for k in KS:
    if kind == "hw":
        cfg = scenarios.highway_config(batch_capacity=B)
        x0, z0, xref, pp = scenarios.highway_batch(B)
    else:
        cfg = scenarios.quadruped_config(batch_capacity=B)
        x0, z0, xref = scenarios.quadruped_batch(B)
        pp = None
    cfg.reserved[7] = k
    mpc = batch.BatchedBranchMPC(cfg)
    tx, tz, tr = [torch.as_tensor(a, device=dev) for a in (x0, z0, xref)]
    tp = None if pp is None else torch.as_tensor(pp, device=dev)
    per_step = []
    for s in range(steps):
        out = mpc.solve(tx, tz, tr, tp, outputs=("u0", "status", "cycles"))
        torch.cuda.synchronize()
        per_step.append((out["cycles"].double().mean().item() / 1.965e6, mpc.last_kernel_ms()))
        mpc.plant_step(tx, out["u0"], tz, 0, tp)
    rows[k] = per_step
    mpc.close()
warps = 592
for s in (0, steps - 1):
    tot = rows[0][s][0]
    print("step %d (%s): kernel %.2f ms, mean problem time %.3f ms" % (s, "cold" if s == 0 else "warm", rows[0][s][1], tot))
    acc = 0.0
    for k in KS[1:]:
        v = rows[k][s][0]
        acc += v if k < 10 else 0.0
        print("   %-30s %.4f ms  %5.1f %%" % (NAMES[k], v, 100 * v / tot))
    print("   %-30s %.4f ms  %5.1f %%" % ("(unaccounted)", tot - acc, 100 * (tot - acc) / tot))
