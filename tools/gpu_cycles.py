"""Development aid: per-problem cycle accounting of the solve kernel (where does a step's time go?)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch  # noqa: E402
from _bmpc import batch, scenarios  # noqa: E402

B = int(os.environ.get("B", "16384"))
knobs = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    knobs[k] = float(v) if ("." in v or "e" in v) else int(v)
occ = knobs.pop("occ", 0)
cfg = scenarios.highway_config(batch_capacity=B, **knobs)
cfg.reserved[1] = occ
for _k in range(8):
    if os.environ.get('BMPC_R%d' % _k):
        cfg.reserved[_k] = int(os.environ['BMPC_R%d' % _k])
mpc = batch.BatchedBranchMPC(cfg)
info = mpc.launch_info()
print("launch:", info, flush=True)
x0, z0, xref, pp = scenarios.highway_batch(B)
dev = torch.device("cuda", 0)
tx, tz, tr, tp = [torch.as_tensor(a, device=dev) for a in (x0, z0, xref, pp)]
outs = ("u0", "status", "iters", "nfact", "nsolve", "cycles")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if os.environ.get('FLUSH') else None
for s in range(int(os.environ.get('STEPS', '4'))):
    if flush is not None:
        flush.zero_()
    out = mpc.solve(tx, tz, tr, tp, outputs=outs)
    torch.cuda.synchronize()
    ms = mpc.last_kernel_ms()
    cyc = out["cycles"].cpu().numpy().astype(float)
    ns = out["nsolve"].cpu().numpy().astype(float)
    nf = out["nfact"].cpu().numpy().astype(float)
    it = out["iters"].cpu().numpy()
    st = out["status"].cpu().numpy()
    # least squares: cycles ~ a*nsolve + b*nfact + c
    A = np.column_stack([ns, nf, np.ones(B)])
    coef, *_ = np.linalg.lstsq(A, cyc, rcond=None)
    warps = info["warps"]
    print("step %d kernel %.2f ms | status %s | sum cycles/warps = %.2f ms @1.965GHz, max problem %.2f ms, p99 %.3f ms, mean %.3f ms"
          % (s, ms, np.bincount(st, minlength=4).tolist(), cyc.sum() / warps / 1.965e6, cyc.max() / 1.965e6,
             np.percentile(cyc, 99) / 1.965e6, cyc.mean() / 1.965e6))
    print("   cycles ~ %.0f*nsolve + %.0f*nfact + %.0f ; mean nsolve %.1f nfact %.2f iters %.1f" % (coef[0], coef[1], coef[2], ns.mean(), nf.mean(), it.mean()))
    if cfg.reserved[7] == 1:
        used = cyc > 0
        print("   interior point: used by %.3f of the problems, %.3f ms each, share of all problem time %.3f (x warps: %.2f ms of the kernel)"
              % (used.mean(), cyc[used].mean() / 1.965e6 if used.any() else 0, 0, cyc.sum() / warps / 1.965e6))
    mpc.plant_step(tx, out["u0"], tz, 0, tp)
mpc.close()
