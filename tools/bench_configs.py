"""Throughput of every BASELINE.json config on one GPU (development/reporting aid; bench.py is the contract line).

  cfg2  robustMPC chain, 4096 episodes          cfg3  highway BranchMPC m=3 NB=2, 16384 episodes (cold + warm)
  cfg4  quadruped BranchMPCProx, 8192 episodes  cfg5  highway tree sweep m in {2,3,4} x NB in {1,2,3}, 65536 episodes total
Prints one JSON line per config: cold and warm solves/s from CUDA events, status counts, mean work.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import torch  # noqa: E402
from _bmpc import abi, batch, scenarios  # noqa: E402

dev = torch.device("cuda", 0)
OUTS = ("u0", "status", "iters", "nfact", "nsolve")


def apply_env(cfg):
    """experiment overrides: BMPC_R<k>=v sets cfg.reserved[k]; BMPC_<knob>=v sets a named solver knob"""
    for k in range(8):
        if os.environ.get("BMPC_R%d" % k):
            cfg.reserved[k] = int(os.environ["BMPC_R%d" % k])
    for knob in ("polish_careful", "polish_passes", "polish_first", "polish_every", "warm_polish", "rho_refresh", "max_iter"):
        if os.environ.get("BMPC_" + knob):
            setattr(cfg, knob, int(os.environ["BMPC_" + knob]))
    for knob in ("polish_mult", "polish_big", "theta", "theta_u", "alpha"):
        if os.environ.get("BMPC_" + knob):
            setattr(cfg, knob, float(os.environ["BMPC_" + knob]))
    return cfg


def run(name, cfg, x0, z0, xref, pp, plant_policy, warm_steps=6):
    cfg = apply_env(cfg)
    mpc = batch.BatchedBranchMPC(cfg)
    B = x0.shape[0]
    tx, tz, tr = [torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in (x0, z0, xref)]
    tp = None if pp is None else torch.as_tensor(np.ascontiguousarray(pp), device=dev)
    times, stats = [], []
    for s in range(1 + warm_steps):
        out = mpc.solve(tx, tz, tr, tp, outputs=OUTS)
        torch.cuda.synchronize()
        times.append(mpc.last_kernel_ms())
        st = out["status"].cpu().numpy()
        stats.append((np.bincount(st, minlength=4).tolist(), float(out["iters"].double().mean()),
                      float(out["nfact"].double().mean()), float(out["nsolve"].double().mean())))
        mpc.plant_step(tx, out["u0"], tz, plant_policy, tp)
    info = mpc.launch_info()
    warm = times[2:]
    line = {"config": name, "episodes": B, "cold_ms": round(times[0], 3), "cold_solves_per_s": round(B / times[0] * 1e3),
            "warm_ms_mean": round(float(np.mean(warm)), 3), "warm_solves_per_s": round(B / float(np.mean(warm)) * 1e3),
            "status_cold": stats[0][0], "status_warm_last": stats[-1][0],
            "mean_iters_cold_warm": [round(stats[0][1], 1), round(stats[-1][1], 1)],
            "mean_nfact_cold_warm": [round(stats[0][2], 2), round(stats[-1][2], 2)],
            "mean_nsolve_cold_warm": [round(stats[0][3], 1), round(stats[-1][3], 1)],
            "nodes": [mpc.totalx, mpc.totalu], "launch": info}
    print(json.dumps(line), flush=True)
    mpc.close()
    return line


def run_env(name, B=16384, steps=30):
    """Closed loop entirely on the device (bmpc_env_step: obstacle policy + lanes + xRef + solve + plants per episode)."""
    from _bmpc import env as benv
    x0, z0, _, _ = scenarios.highway_batch(B, seed=1240)
    mpc = batch.BatchedBranchMPC(apply_env(scenarios.highway_config(batch_capacity=B)))
    e = benv.BatchedHighwayEnv(mpc, x0, z0, 4)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    stat = []
    for t in range(steps):
        out = e.step(outputs=("u0", "status", "iters"))
        ev[t + 1].record()
        stat.append(out["status"].clone())
    torch.cuda.synchronize()
    ms = [ev[t].elapsed_time(ev[t + 1]) for t in range(steps)]
    warm = ms[5:]
    h = e.host()
    line = {"config": name, "episodes": B, "steps": steps, "cold_step_ms": round(ms[0], 3), "warm_step_ms_mean": round(float(np.mean(warm)), 3),
            "episode_steps_per_s": round(B / float(np.mean(warm)) * 1e3), "status_all_steps": torch.bincount(torch.cat(stat), minlength=4).tolist(),
            "collided": int(h["collided"].sum()), "obstacle_policy_hist_last": np.bincount(h["obs_policy"], minlength=3).tolist(),
            "launches_per_step": 4}
    print(json.dumps(line), flush=True)
    mpc.close()


def main():
    which = sys.argv[1:] or ["cfg3", "cfg2", "cfg4", "cfg5", "env"]
    if "env" in which:
        run_env("closed loop on the device: Highway_env.step x 16384 episodes (row f1)")
    sweep = [(m, NB) for m in (2, 3, 4) for NB in (1, 2, 3)]
    if os.environ.get("BMPC_SWEEP"):      # e.g. BMPC_SWEEP=3x3,4x2
        sweep = [tuple(int(v) for v in t.split("x")) for t in os.environ["BMPC_SWEEP"].split(",")]
    if "cfg3" in which:
        B = 16384
        x0, z0, xref, pp = scenarios.highway_batch(B, seed=1237)
        run("cfg3 highway BranchMPC m3 NB2", scenarios.highway_config(batch_capacity=B), x0, z0, xref, pp, 0)
    if "cfg2" in which:
        B = 4096
        x0, z0, xref, pp = scenarios.highway_batch(B, seed=1236)
        cfg = scenarios.highway_config(batch_capacity=B)
        cfg.controller = abi.CTRL_ROBUST
        run("cfg2 robustMPC chain (MPC_nobranch semantics)", cfg, x0, z0, xref, pp, 0)
    if "cfg4" in which:
        B = 8192
        x0, z0, xref = scenarios.quadruped_batch(B, seed=1238)
        run("cfg4 quadruped BranchMPCProx m2 NB2 N25", scenarios.quadruped_config(batch_capacity=B), x0, z0, xref, None, 0,
            warm_steps=3)
    if "cfg5" in which:
        names = ["maintain", "brake", "lc", "trackv"]
        per = 65536 // 9
        for m, NB in sweep:
            if True:
                x0, z0, xref, pp3 = scenarios.highway_batch(per, seed=1239 + 10 * m + NB)
                pp = np.zeros((per, m, 4))
                if m >= 3:
                    pp[:, 2, :] = pp3[:, 2, :]
                if m >= 4:
                    pp[:, 3, 0] = 20.0
                run("cfg5 highway sweep m%d NB%d" % (m, NB), scenarios.highway_config(policies=names[:m], NB=NB, batch_capacity=per),
                    x0, z0, xref, pp, 0, warm_steps=3)


if __name__ == "__main__":
    main()
