"""Quick GPU sanity run (development aid): golden fixtures through the C ABI + a cold/warm throughput probe."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch  # noqa: E402
from _bmpc import abi, batch, scenarios  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def fixture(name, **kw):
    g = np.load(os.path.join(G, name + ".npz"))
    cfg = scenarios.highway_config(policies=[str(p) for p in g["meta_policies"]], NB=int(g["meta_NB"]), N=int(g["meta_N"]),
                                   lc_target=tuple(g["meta_lc_target"]), **kw)
    mpc = batch.BatchedBranchMPC(cfg)
    assert (mpc.topology() == g["s0_tree"]).all(), "topology"
    for k in range(int(g["meta_steps"])):
        pre = "s%d_" % k
        r = mpc.solve_host(g[pre + "x0"], g[pre + "z0"], g[pre + "xref"])
        print(name, k, "status", r["status"][0], "iters", r["iters"][0], "nfact", r["nfact"][0],
              "du0 %.2e" % np.abs(r["u0"][0] - g[pre + "uPred"][0]).max(),
              "dU %.2e" % np.abs(r["uPred"][0] - g[pre + "uPred"]).max(),
              "dX %.2e" % np.abs(r["xPred"][0] - g[pre + "xPred"]).max(),
              "relJ %.2e" % (abs(r["objective"][0] - float(g[pre + "objective"])) / abs(float(g[pre + "objective"]))),
              "dw %.1e" % np.abs(r["branch_w"][0] - g[pre + "w"]).max(), flush=True)
    mpc.close()


def throughput(B=16384, steps=6):
    cfg = scenarios.highway_config(batch_capacity=B)
    mpc = batch.BatchedBranchMPC(cfg)
    x0, z0, xref, pp = scenarios.highway_batch(B)
    dev = torch.device("cuda", 0)
    tx, tz, tr, tp = [torch.as_tensor(a, device=dev) for a in (x0, z0, xref, pp)]
    for s in range(steps):
        torch.cuda.synchronize()
        t = time.time()
        out = mpc.solve(tx, tz, tr, tp)
        torch.cuda.synchronize()
        dt = time.time() - t
        st = out["status"].cpu().numpy()
        it = out["iters"].cpu().numpy()
        nf = out["nfact"].cpu().numpy()
        print("step", s, "B", B, "wall %.1f ms" % (dt * 1e3), "kernel %.1f ms" % mpc.last_kernel_ms(),
              "solves/s %.3e" % (B / dt), "status counts", np.bincount(st, minlength=4).tolist(),
              "iters mean %.1f max %d" % (it.mean(), it.max()), "nfact mean %.1f" % nf.mean(), flush=True)
        u0 = out["u0"].cpu().numpy()
        x0 = scenarios.euler_highway(x0, u0)
        z0 = scenarios.euler_highway(z0, np.column_stack([np.zeros(B), -0.1 * z0[:, 3]]))
        tx, tz = torch.as_tensor(x0, device=dev), torch.as_tensor(z0, device=dev)
    mpc.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "fp64 peak TF/s", abi.load_library().bmpc_measure_fp64_peak(0, 4096))
    for name in ("highway_branch_default", "highway_branch_close", "highway_branch_m2_nb3", "highway_branch_m3_nb1"):
        fixture(name)
    throughput(int(os.environ.get("B", "16384")))
