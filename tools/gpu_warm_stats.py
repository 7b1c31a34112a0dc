"""Development aid: how warm highway solves split between the warm polish and the ADMM path, and what each costs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch  # noqa: E402
from _bmpc import batch, scenarios  # noqa: E402

B = 16384
cfg = scenarios.highway_config(batch_capacity=B)
for k in range(8):
    if os.environ.get("BMPC_R%d" % k):
        cfg.reserved[k] = int(os.environ["BMPC_R%d" % k])
for knob in ("polish_passes", "warm_polish", "polish_careful", "polish_first", "polish_every", "rho_refresh"):
    if os.environ.get("BMPC_" + knob):
        setattr(cfg, knob, int(os.environ["BMPC_" + knob]))
mpc = batch.BatchedBranchMPC(cfg)
x0, z0, xref, pp = scenarios.highway_batch(B)
t = [torch.as_tensor(a, device="cuda") for a in (x0, z0, xref, pp)]
outs = ("u0", "status", "iters", "nfact", "nsolve", "cycles")
prev_admm = None
for s in range(int(os.environ.get("STEPS", "8"))):
    out = mpc.solve(*t, outputs=outs)
    torch.cuda.synchronize()
    ms = mpc.last_kernel_ms()
    it, nf, ns, cy = (out[k].cpu().numpy() for k in ("iters", "nfact", "nsolve", "cycles"))
    st = np.bincount(out["status"].cpu().numpy(), minlength=4).tolist()
    admm = it > 0
    if prev_admm is not None and os.environ.get("CORR"):
        print("   P(ADMM path | ADMM path last step) = %.2f, P(ADMM path | warm-only last step) = %.2f; warm passes wasted by the first group: nfact of its warm successes %.2f"
              % (admm[prev_admm].mean(), admm[~prev_admm].mean(), nf[prev_admm & ~admm].mean()))
    prev_admm = admm
    if os.environ.get("BRIEF"):
        print("step %d  %.2f ms  status %s  ADMM path %.1f %%  iters %.1f nfact %.2f" % (s, ms, st, 100 * (it > 0).mean(), it.mean(), nf.mean()), flush=True)
        mpc.plant_step(t[0], out["u0"], t[1], 0, t[3])
        continue
    warm = it == 0
    print("step %d  %.2f ms | warm-polish only %.1f %% (nfact %.2f nsolve %.1f, %.0f kcycles) | ADMM path %.1f %% (iters %.1f nfact %.2f nsolve %.1f, %.0f kcycles) | share of time on the ADMM path %.1f %%"
          % (s, ms, 100 * warm.mean(), nf[warm].mean() if warm.any() else 0, ns[warm].mean() if warm.any() else 0,
             cy[warm].mean() / 1e3 if warm.any() else 0, 100 * (~warm).mean(), it[~warm].mean(), nf[~warm].mean(), ns[~warm].mean(),
             cy[~warm].mean() / 1e3, 100 * cy[~warm].sum() / cy.sum()), flush=True)
    if s == int(os.environ.get("STEPS", "8")) - 1:
        print("  nfact histogram (warm-only):", np.bincount(nf[warm], minlength=6)[:8].tolist(), " (ADMM path):", np.bincount(nf[~warm], minlength=8)[:10].tolist())
        print("  iters histogram (ADMM path, bins of 5):", np.bincount(it[~warm] // 5, minlength=10)[:12].tolist())
    mpc.plant_step(t[0], out["u0"], t[1], 0, t[3])
mpc.close()
