"""Development aid: every bench workload once, errors reported per part (which tree shape / controller fails to launch?)."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch
import bench
for which in ("2", "4", "5", "cvar"):
    B = bench.DEFAULT_BATCH[which] if which != "cvar" else 4096
    try:
        parts = bench.make_workload(which, B, 0, 1, 0)
    except Exception:
        print("config", which, "create failed"); traceback.print_exc(); continue
    for pt in parts:
        info = pt["mpc"].launch_info()
        try:
            out = pt["mpc"].solve(pt["x"], pt["z"], pt["r"], pt["p"])
            torch.cuda.synchronize()
            print("config", which, pt["sizes"]["m"], pt["sizes"]["NB"], "B", pt["B"], info, "ok", torch.bincount(out["status"], minlength=4).tolist(), "%.2f ms" % pt["mpc"].last_kernel_ms())
        except Exception as e:
            print("config", which, pt["sizes"]["m"], pt["sizes"]["NB"], "B", pt["B"], info, "FAILED", e)
        pt["mpc"].close()
