"""Development aid: solver work per step in the device closed loop (bmpc_env_step) vs the fixed-reference loop of bench.py."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch
from _bmpc import batch, scenarios, env as benv
B = 16384
x0, z0, _, _ = scenarios.highway_batch(B, seed=1240)
cfg = scenarios.highway_config(batch_capacity=B)
for k in range(8):
    if os.environ.get("BMPC_R%d" % k):
        cfg.reserved[k] = int(os.environ["BMPC_R%d" % k])
mpc = batch.BatchedBranchMPC(cfg)
e = benv.BatchedHighwayEnv(mpc, x0, z0, 4)
for t in range(16):
    out = e.step(outputs=("u0", "status", "iters", "nfact", "nsolve", "cycles"))
    torch.cuda.synchronize()
    it = out["iters"].double(); nf = out["nfact"].double(); ns = out["nsolve"].double(); cy = out["cycles"].double()
    print("t %2d kernel %.1f ms | iters mean %.1f (==0: %.2f) nfact %.2f nsolve %.1f | status %s | xref v mean %.1f | problem ms mean %.3f p99 %.2f max %.2f" % (
        t, mpc.last_kernel_ms(), it.mean().item(), (it == 0).double().mean().item(), nf.mean().item(), ns.mean().item(),
        torch.bincount(out["status"], minlength=4).tolist(), e.xref[:, 2].mean().item(), cy.mean().item() / 1.965e6,
        torch.quantile(cy, 0.99).item() / 1.965e6, cy.max().item() / 1.965e6), flush=True)
