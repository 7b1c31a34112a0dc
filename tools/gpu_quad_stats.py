"""Development aid: how quadruped (BranchMPCProx) closed-loop solves split between the warm polish, the ADMM path and the
interior-point fallback, and what each costs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch  # noqa: E402
from _bmpc import batch, scenarios  # noqa: E402

B = int(os.environ.get("B", "8192"))
STEPS = int(os.environ.get("STEPS", "12"))


def make(mode):
    cfg = scenarios.quadruped_config(batch_capacity=B)
    for k in range(8):
        if os.environ.get("BMPC_R%d" % k):
            cfg.reserved[k] = int(os.environ["BMPC_R%d" % k])
    for knob in ("polish_passes", "warm_polish", "polish_careful", "polish_first", "polish_every", "rho_refresh"):
        if os.environ.get("BMPC_" + knob):
            setattr(cfg, knob, int(os.environ["BMPC_" + knob]))
    cfg.reserved[7] = mode
    return batch.BatchedBranchMPC(cfg)


x0, z0, xref = scenarios.quadruped_batch(B)
outs = ("u0", "status", "iters", "nfact", "nsolve", "cycles")
runs = {}
for mode in (0, 1):      # whole-solve cycles, then cycles inside the interior point
    mpc = make(mode)
    t = [torch.as_tensor(a.copy(), device="cuda") for a in (x0, z0, xref)]
    rows = []
    for s in range(STEPS):
        out = mpc.solve(t[0], t[1], t[2], None, outputs=outs)
        torch.cuda.synchronize()
        rows.append((mpc.last_kernel_ms(), {k: out[k].cpu().numpy() for k in outs[1:]}))
        mpc.plant_step(t[0], out["u0"], t[1], 0, None)
    runs[mode] = rows
    mpc.close()
for s in range(STEPS):
    ms, o = runs[0][s]
    ipm_cy = runs[1][s][1]["cycles"]
    it, nf, ns, cy = o["iters"], o["nfact"], o["nsolve"], o["cycles"].astype(float)
    st = np.bincount(o["status"], minlength=4).tolist()
    ipm = ipm_cy > 0
    warm = (it == 0) & ~ipm
    admm = (it > 0) & ~ipm
    def grp(m):
        return "%.1f %% (nfact %.1f nsolve %.1f iters %.1f, %.0f kcycles, %.0f %% of time)" % (
            100 * m.mean(), nf[m].mean() if m.any() else 0, ns[m].mean() if m.any() else 0, it[m].mean() if m.any() else 0,
            cy[m].mean() / 1e3 if m.any() else 0, 100 * cy[m].sum() / cy.sum())
    print("step %d  %.2f ms  status %s | warm-only %s | ADMM %s | interior point %s" % (s, ms, st, grp(warm), grp(admm), grp(ipm)), flush=True)
