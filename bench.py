#!/usr/bin/env python
"""Branch-MPC throughput benchmark (BASELINE.json metric: highway Branch-MPC solves/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--config {2,3,4,5,cvar,merge}] [--impl reference]

Workload (N=1): BASELINE.json configs[2] - highway Branch MPC (m=3 policies, NB=2, N=8; 13 branches, 106 state nodes,
97 input nodes), 16384 independent episodes per GPU in closed loop: a *step* is one MPC solve of every episode
(tree update + linearisation + tree QP to the parity tolerances) followed by the Euler plant step that produces the
next step's initial states (both are this library's kernels).  The first (cold, inittree) solve of every episode is a
warm-up step; the timed steps are the updatetree path the reference spends 99 of its 100 solves per episode in.

  value   device-resident throughput: B*K / sum of per-step CUDA-event times (L2 flushed between steps, not timed)
  e2e     same metric through the host API (pinned host buffers -> bmpc_solve_host -> host results), copies included
  e2e_full  same metric through the drop-in `MPC_branch.BranchMPC.solve((B, 4) arrays)`: every result array of the reference
            interface (uPred, xPred, xLin, zPred, weights ...) comes back to the host, ~11.7 KB per solve
  --impl reference   the CPU restatement of the reference path (oracle/, float64: tree update, linearisation, QP assembly with
            the dense->CSC conversions, OSQP-style ADMM with default settings + polish, cold, fresh setup) on the host cores
  --config  other BASELINE configs on the same harness (2 robustMPC chain 4096, 4 quadruped 8192, 5 tree sweep 65536 episodes
            sharded over the ranks, cvar = BranchMPC_CVaR 16384); the default line carries short runs of them in
            `extra.other_configs`, so a driver-run number exists for every config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "belief-planning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "highway Branch-MPC solves/s"
UNIT = "solves/s"
N_TREE = dict(m=3, NB=2, N=8, n=4, d=2, totalu=97, totalx=106, nbranch=13)
# algorithmic flops per input node (n=4, d=2, c=5 one-sided rows)
F_FACT_NODE, F_ITER_NODE = 707.0, 340.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def tree_sizes(m, NB, N, n=4, d=2):
    nbranch = sum(m ** k for k in range(NB + 1))
    totalu = 1 + N * (nbranch - 1)
    return dict(m=m, NB=NB, N=N, n=n, d=d, totalu=totalu, totalx=totalu + m ** NB, nbranch=nbranch)


def node_flops(n, d, c):
    """flops per input node of one Riccati factorisation / one KKT solve + row update."""
    fact = 4 * n ** 3 + 6 * n * n * d + 2 * n * d * d + 2 * d * d * n + d ** 3 / 3.0 + 2 * n * n + 2 * c * n * n
    it = 6 * n * n + 8 * n * d + 2 * d * d + 4 * c * n + 10 * (c + d) + 3 * n + 5 * d
    return fact, it


def algorithmic_bytes_per_solve(t=N_TREE, rows=3, refresh=8):
    """HBM bytes one warm solve has to move (float64): inputs, per-episode policy parameters, the warm-start state read and
    written back (uLin, arg-max children, OldInput, started), the solver's own per-episode caches (curvature-matched rho: read
    on refresh-1 of refresh solves and rewritten on the other; active-set codes of the last optimum: read and rewritten) and
    the light outputs the bench requests."""
    n, d, m = t["n"], t["d"], t["m"]
    inputs = 3 * n * 8 + m * 4 * 8
    state = 2 * ((t["totalu"] + 1) * d * 8 + t["nbranch"] * 4 + d * 8 + 4)
    caches = t["totalu"] * (rows + d) * 8 + 2 * t["totalu"] * 8 + 2 * 8 + 4
    outputs = d * 8 + 8 + 4 * 4
    return inputs + state + caches + outputs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the restated reference path (oracle/) on the host cores
# ------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, count, warm_steps, style = args
    from _bmpc import scenarios
    from oracle import osqp_style, params
    solver = osqp_style.qp_solver if style == "osqp" else None
    x0, z0, xref, pp = scenarios.highway_batch(count, seed=seed)
    t_solve, n_solve = 0.0, 0
    for i in range(count):
        mpc = params.highway_branch_mpc(lc_target=pp[i, 2])
        x, z = x0[i], z0[i]
        for s in range(1 + warm_steps):
            t = time.perf_counter()
            u = mpc.solve(x, z, xref[i], qp_solver=solver).copy()
            dt = time.perf_counter() - t
            if s > 0 or warm_steps == 0:
                t_solve += dt
                n_solve += 1
            x = scenarios.euler_highway(x[None], u[None])[0]
            z = scenarios.euler_highway(z[None], np.array([[0.0, -0.1 * z[3]]]))[0]
    return t_solve, n_solve


def cpu_reference_rate(problems_per_core=2, warm_steps=1, cores=None, seed=4242, style="osqp"):
    """solves/s of the restated reference path with one worker process per host core; every worker runs `problems_per_core`
    episodes for 1 cold + `warm_steps` closed-loop solves and the later solves are timed (the updatetree path the GPU arm
    times; the reference itself never warm-starts its solver, so each of them is a cold OSQP-style solve with a fresh setup).
    style "osqp": tree update + linearisation + assembly + dense->CSC + OSQP-default ADMM + polish (oracle/osqp_style.py, what
    the reference does); style "exact": the parity oracle (exact optimum by interior point + active-set polish)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(seed + 17 * c, problems_per_core, warm_steps, style) for c in range(cores)])
    wall = time.perf_counter() - t0
    n = sum(r[1] for r in res)
    busy = sum(r[0] for r in res)
    # all workers run concurrently: aggregate rate = solves / (mean busy time per worker)
    rate = n / (busy / cores) if busy > 0 else 0.0
    what = ("restated reference path: oracle.BranchMPCOracle + oracle.osqp_style (OSQP-default ADMM eps 1e-3 + polish, cold, fresh "
            "setup, dense->CSC)" if style == "osqp" else "parity oracle: oracle.BranchMPCOracle with the exact QP solve")
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "problems": n,
            "sample": "%d problems = %d episodes x %d timed closed-loop solves per core, %d cores, float64; %s; per-solve mean "
                      "%.3f s; wall %.1f s" % (n, problems_per_core, max(warm_steps, 1), cores, what, busy / max(n, 1), wall)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates = []
    t_all = time.perf_counter()
    last = None
    for s in range(args.warmup + args.steps):
        last = cpu_reference_rate(problems_per_core=1, warm_steps=1, seed=9000 + s, style="osqp")
        if s >= args.warmup:
            rates.append(last["value"])
    value = float(np.mean(rates)) if rates else 0.0
    B = (os.cpu_count() or 1)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (1e3 * B / value) if value else None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample": "bounded sample of that workload: each step = one closed-loop solve of one episode on every host "
                                 "core by the restated reference path (oracle/ + oracle/osqp_style.py: OSQP-default ADMM + "
                                 "polish, cold, fresh setup; casadi/osqp are not installable here)"},
            "cpu_baseline": dict(last, value=value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------------------
# workloads: one entry per BASELINE.json config (and BranchMPC_CVaR); a workload is a list of (handle, x, z, xref, policy
# parameters) parts that one step solves back to back (cfg 5: one part per tree shape)
# ------------------------------------------------------------------------------------------------------------
SWEEP = [(m, NB) for m in (2, 3, 4) for NB in (1, 2, 3)]
POLICY_NAMES = ["maintain", "brake", "lc", "trackv"]
CONFIG_NAMES = {
    "2": "robustMPC chain (BASELINE configs[1], MPC_nobranch semantics): Nx=18/Nu=17, 97 collision rows; %d episodes per GPU",
    "3": None,
    "4": "quadruped BranchMPCProx (BASELINE configs[3]): m=2, NB=2, N=25 -> 7 branches / 155 state nodes / 151 input nodes; %d episodes per GPU",
    "5": "highway tree sweep (BASELINE configs[4]): m in {2,3,4} x NB in {1,2,3}, N=8; %d episodes in total, sharded over the ranks",
    "cvar": "highway BranchMPC_CVaR (what main_branch.py:48 builds, ralpha=0.9): m=3, NB=2, N=8; %d episodes per GPU",
    "merge": "merge scenario (main_branch.sim_merge): PredictiveModel_merge, BranchMPC_CVaR ralpha=0.1, m=2, NB=1, N=40 -> 81 input "
             "nodes, per-episode state transform S and bounds; %d episodes per GPU",
}
DEFAULT_BATCH = {"2": 4096, "3": 16384, "4": 8192, "5": 65536, "cvar": 16384, "merge": 4096}


def make_workload(which, B, rank, world, local):
    """Returns (parts, label, flop/byte model).  B = episodes per GPU (cfg 5: total episodes over all ranks and shapes)."""
    import torch
    from _bmpc import abi, batch, scenarios, shard
    dev = torch.device("cuda", local)
    parts = []

    def part(cfg, arrays, sizes, rows):
        mpc = batch.BatchedBranchMPC(cfg)
        t = [None if a is None else torch.as_tensor(np.ascontiguousarray(a), device=dev) for a in arrays]
        parts.append({"mpc": mpc, "x": t[0], "z": t[1], "r": t[2], "p": t[3], "B": arrays[0].shape[0], "sizes": sizes, "c": rows})
        if len(t) > 4:
            parts[-1]["S"], parts[-1]["bd"] = t[4], t[5]

    if which in ("3", "cvar", "2"):
        x0, z0, xref, pp = scenarios.highway_batch(B, seed={"3": 1237, "cvar": 1241, "2": 1236}[which] + 1000 * rank)
        cfg = scenarios.highway_config(batch_capacity=B, device=local)
        if which == "cvar":
            cfg.controller, cfg.cvar_alpha = abi.CTRL_CVAR, 0.9
        if which == "2":
            cfg.controller = abi.CTRL_ROBUST
        sizes = dict(N_TREE) if which != "2" else dict(m=1, NB=1, N=17, n=4, d=2, totalu=17, totalx=18, nbranch=2)
        part(cfg, (x0, z0, xref, pp), sizes, 5)
    elif which == "4":
        x0, z0, xref = scenarios.quadruped_batch(B, seed=1238 + 1000 * rank)
        part(scenarios.quadruped_config(batch_capacity=B, device=local), (x0, z0, xref, None), tree_sizes(2, 2, 25, 3, 3), 1)
    elif which == "merge":
        x0, z0, xref, S, bd = scenarios.merge_batch(B, seed=1242 + 1000 * rank)
        part(scenarios.merge_config(batch_capacity=B, device=local), (x0, z0, xref, None, S, bd), tree_sizes(2, 1, 40), 3)
    elif which == "5":
        per_shape = B // len(SWEEP)
        for m, NB in SWEEP:
            lo, hi = shard.shard_bounds(per_shape, world, rank)
            x0, z0, xref, pp3 = scenarios.highway_batch(per_shape, seed=1239 + 10 * m + NB)
            pp = np.zeros((per_shape, m, 4))
            if m >= 3:
                pp[:, 2, :] = pp3[:, 2, :]
            if m >= 4:
                pp[:, 3, 0] = 20.0
            cfg = scenarios.highway_config(policies=POLICY_NAMES[:m], NB=NB, batch_capacity=hi - lo, device=local)
            part(cfg, (x0[lo:hi], z0[lo:hi], xref[lo:hi], pp[lo:hi]), tree_sizes(m, NB, 8), 5)
    else:
        raise SystemExit("unknown --config %r" % which)
    return parts


def timed_steps(parts, K, W, flush, barrier, outputs=("u0", "status", "iters", "nfact", "nsolve")):
    """W untimed + K timed closed-loop steps; one step = solve + plant step of every part.  Returns per-step event times (ms),
    per-part stats of the timed steps and the kernel launches counted by the library."""
    import torch
    names = ("iters", "nfact", "nsolve", "status")
    stats = [{k: [] for k in names} for _ in parts]
    # per-step statistics are copied into rows of buffers allocated BEFORE the timed region (row K = scratch of the untimed
    # steps): a tensor allocated inside it can make the caching allocator call cudaMalloc, which drains the queue and shows
    # up as a 20-30 ms step (seen once in ~20 steps before this)
    bufs = [None] * len(parts)
    timed_count = [0]

    def one_step(timed):
        if flush is not None:
            flush.zero_()                  # evict L2 between steps (not inside the event pair)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i, pt in enumerate(parts):
            if "S" in pt:
                out = pt["mpc"].solve_transformed(pt["x"], pt["z"], pt["r"], pt["S"], pt["bd"], pt["p"], outputs=outputs)
            else:
                out = pt["mpc"].solve(pt["x"], pt["z"], pt["r"], pt["p"], outputs=outputs)
            pt["mpc"].plant_step(pt["x"], out["u0"], pt["z"], 0, pt["p"])
            if bufs[i] is None:
                bufs[i] = {k: torch.empty((K + 1,) + tuple(out[k].shape), dtype=out[k].dtype, device=out[k].device) for k in names}
            row = timed_count[0] if timed else K
            for k in names:                                   # warm-up steps do exactly what timed steps do
                bufs[i][k][row].copy_(out[k])
            pt["last"] = out
        e1.record()
        if timed:
            timed_count[0] += 1
        return e0, e1

    for _ in range(W):
        one_step(False)
    # settle: the first steps after start-up (allocator growth, clock ramp after the idle second the clock sampler needs to
    # start) have shown one-off 70-85 ms hiccups; keep stepping untimed until two consecutive steps agree within 25 %
    import torch as _t
    prev = None
    for _ in range(12):
        e0, e1 = one_step(False)
        _t.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if prev is not None and abs(ms - prev) <= 0.25 * min(ms, prev):
            break
        prev = ms
    barrier()
    l0 = sum(pt["mpc"].launch_count() for pt in parts)
    barrier()
    t_wall = time.perf_counter()
    events = [one_step(True) for _ in range(K)]
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = sum(pt["mpc"].launch_count() for pt in parts) - l0
    for i in range(len(parts)):
        for k in names:
            stats[i][k] = [bufs[i][k][s] for s in range(K)]
    return [e0.elapsed_time(e1) for e0, e1 in events], stats, launches, t_wall


def summarize(parts, stats):
    import torch
    st = torch.cat([torch.cat(s["status"]) for s in stats])
    flops = 0.0
    solves = 0
    byts = 0
    for pt, s in zip(parts, stats):
        t = pt["sizes"]
        f_fact, f_it = node_flops(t["n"], t["d"], pt["c"])
        nf, ns = float(torch.cat(s["nfact"]).double().mean()), float(torch.cat(s["nsolve"]).double().mean())
        flops += pt["B"] * t["totalu"] * (nf * f_fact + ns * f_it)
        byts += pt["B"] * algorithmic_bytes_per_solve(t)
        solves += pt["B"]
    it = torch.cat([torch.cat(s["iters"]) for s in stats]).double()
    nf = torch.cat([torch.cat(s["nfact"]) for s in stats]).double()
    ns = torch.cat([torch.cat(s["nsolve"]) for s in stats]).double()
    return {"status_counts": torch.bincount(st, minlength=4).tolist(), "mean_admm_iters": float(it.mean()),
            "mean_factorizations": float(nf.mean()), "mean_kkt_solves": float(ns.mean()),
            "flops_per_step": flops, "bytes_per_step": byts, "solves_per_step": solves}


def spot_check(parts, count=8):
    """The timed numbers are only worth something if the timed solves are right: re-solve `count` episodes of the state the
    LAST timed step started from with the parity oracle (host, exact QP) and compare first input and objective."""
    import torch
    from _bmpc import scenarios
    from oracle import params
    pt = parts[0]
    x, z, r, p = [None if t is None else t[:count].cpu().numpy() for t in (pt["x"], pt["z"], pt["r"], pt["p"])]
    st = pt["mpc"].get_state(count)
    res = pt["mpc"].solve(pt["x"][:count].clone(), pt["z"][:count].clone(), pt["r"][:count].clone(),
                          None if pt["p"] is None else pt["p"][:count].clone(),
                          outputs=("u0", "objective", "status", "iters", "nfact", "nsolve"))
    torch.cuda.synchronize()
    u0, obj = res["u0"].cpu().numpy(), res["objective"].cpu().numpy()
    du, dj = 0.0, 0.0
    for i in range(count):
        ora = params.highway_branch_mpc(lc_target=p[i, 2])
        if st["started"][i]:
            # the oracle linearises about the same warm-start state the device holds: shifted inputs + arg-max children
            ora.uLin = st["uLin"][i].copy()
            ora.p = np.zeros((ora.topo.nbranch, ora.m))
            ora.p[np.arange(ora.topo.nbranch), st["pbest"][i]] = 1.0
            ora.OldInput = st["old_input"][i].copy()
        u = ora.solve(x[i], z[i], r[i])
        du = max(du, float(np.abs(u - u0[i]).max()))
        dj = max(dj, abs(ora.objective - obj[i]) / abs(ora.objective))
    pt["mpc"].set_state(st)
    return {"episodes": count, "max_abs_u0_diff": du, "max_rel_objective_diff": dj, "bars": [1e-3, 1e-4],
            "ok": bool(du < 1e-3 and dj < 1e-4),
            "what": "episodes 0..%d of the closed-loop state after the last timed step, device solve vs oracle.BranchMPCOracle "
                    "(exact QP) from the same warm-start state" % (count - 1)}


def run_gpu(args):
    # stdout carries exactly one JSON line: the image sets NCCL_DEBUG=VERSION, which makes NCCL printf a version banner to
    # stdout (NCCL_DEBUG_FILE does not catch it); must be changed before the library is loaded
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "NONE"
    import torch
    import torch.distributed as dist
    from _bmpc import batch, scenarios, shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    which = args.config
    B = args.batch or DEFAULT_BATCH[which]
    K, W = args.steps, max(args.warmup, 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    parts = make_workload(which, B, rank, world, local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(1.0)      # nvidia-smi's NVML start-up stalls CUDA API calls for tens of ms: let it finish before any step
    step_ms, stats, launches, t_wall = timed_steps(parts, K, W, flush, barrier)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(step_ms))
    # the only communication of the run: max over ranks of the timed region (NCCL all_reduce MAX of one scalar)
    total_ms_max = shard.reduce_stats({"max_total_ms": total_ms}, device=dev)["max_total_ms"]
    summ = summarize(parts, stats)
    solves_all = shard.reduce_stats({"solves": summ["solves_per_step"]}, device=dev)["solves"]
    value = solves_all * K / (total_ms_max * 1e-3)
    check = spot_check(parts) if (rank == 0 and which == "3" and not args.no_cpu) else None

    # ---- end to end through the host API: pinned host inputs, bmpc_solve_host, host results ----
    e2e = e2e_full = lat = None
    if which == "3":
        mpc = parts[0]["mpc"]
        x0, z0, xref, pp = scenarios.highway_batch(B, seed=1237 + 1000 * rank)
        hx = torch.as_tensor(x0).pin_memory().numpy()
        hz = torch.as_tensor(z0).pin_memory().numpy()
        hr = torch.as_tensor(xref).pin_memory().numpy()
        hp = torch.as_tensor(pp).pin_memory().numpy()
        mpc.reset()
        outs = ("u0", "objective", "status")
        e2e_t = 0.0
        for s in range(W + K):
            if s == W:
                barrier()
            t0 = time.perf_counter()
            r = mpc.solve_host(hx, hz, hr, hp, outputs=outs)
            dt = time.perf_counter() - t0
            if s >= W:
                e2e_t += dt
            hx[:] = scenarios.euler_highway(hx, r["u0"])
            hz[:] = scenarios.euler_highway(hz, np.column_stack([np.zeros(B), -0.1 * hz[:, 3]]))
        e2e_max = shard.reduce_stats({"max_e2e_s": e2e_t}, device=dev)["max_e2e_s"]
        e2e = {"value": world * B * K / e2e_max, "unit": UNIT,
               "h2d_bytes_per_step": int(B * (3 * 4 + 3 * 4) * 8), "d2h_bytes_per_step": int(B * (2 * 8 + 8 + 4)),
               "what": "bmpc_solve_host: pinned host arrays in (one DMA each), first input + objective + status out (written by the "
                       "kernel into pinned host memory, then copied into the caller's arrays); all of it inside the timed region"}
        # ---- the same through the reference-facing drop-in class with EVERY result array of the reference interface ----
        import Init_MPC
        import MPC_branch
        from highway_branch_dyn import PredictiveModel, backup_brake, backup_lc, backup_maintain
        from utils import Branch_constants
        cons = Branch_constants(s1=2, s2=3, c2=0.5, tran_diag=0.3, alpha=1, R=1.2, am=6.0, rm=0.3, J_c=20, s_c=1, ylb=0.,
                                yub=7.2, L=4, W=2.5, col_alpha=5, Kpsi=0.1)
        lc = np.array([0.5, 1.8, 15, 0])
        model = PredictiveModel(4, 2, 8, [lambda x: backup_maintain(x, cons), lambda x: backup_brake(x, cons),
                                          lambda x: backup_lc(x, lc)], 0.1, cons)
        par = Init_MPC.initBranchMPC(4, 2, 8, 2, lc, 6.0, 0.3, 4, cons.W)
        ctl = MPC_branch.BranchMPC(par, model)
        ctl.predictiveModel.policy_params = lambda: pp        # per-episode lane-change targets, as update_backup would set
        fx, fz = x0.copy(), z0.copy()
        full_t, per_solve_bytes = 0.0, None
        Kf = max(3, K // 4)
        for s in range(W + Kf):
            if s == W:
                barrier()
            t0 = time.perf_counter()
            ctl.solve(fx, fz, xref)
            dt = time.perf_counter() - t0
            if s >= W:
                full_t += dt
            fx = scenarios.euler_highway(fx, ctl.uPred[:, 0])
            fz = scenarios.euler_highway(fz, np.column_stack([np.zeros(B), -0.1 * fz[:, 3]]))
        full_max = shard.reduce_stats({"max_full_s": full_t}, device=dev)["max_full_s"]
        sz = N_TREE
        per_solve_bytes = (2 + sz["totalu"] * 2 + sz["totalx"] * 4 + 2 * sz["totalu"] * 4 + sz["nbranch"] * 4 + 1) * 8 + 4 * 4 + 8
        e2e_full = {"value": world * B * Kf / full_max, "unit": UNIT, "steps": Kf,
                    "h2d_bytes_per_step": int(B * (3 * 4 + 3 * 4) * 8), "d2h_bytes_per_step": int(B * per_solve_bytes),
                    "what": "drop-in MPC_branch.BranchMPC.solve(x (B,4), z (B,4), xRef (B,4)) -> uPred, xPred, xLin, zPred, branch "
                            "weights/probabilities, objective, status on the host (inputs in one DMA through pinned staging, results "
                            "written by the kernel into the library's pinned host block, which the controller's arrays view)"}
        ctl._solver.close()
    if rank == 0 and which == "3":
        # p50 latency of a single warm solve through the host API (batch of one)
        one = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=1, device=local))
        x0, z0, xref, pp = scenarios.highway_batch(B, seed=1237)
        x1, z1, r1, p1 = x0[:1].copy(), z0[:1].copy(), xref[:1].copy(), pp[:1].copy()
        ts = []
        for s in range(60):
            t0 = time.perf_counter()
            r = one.solve_host(x1, z1, r1, p1, outputs=("u0", "status"))
            ts.append(time.perf_counter() - t0)
            x1 = scenarios.euler_highway(x1, r["u0"])
            z1 = scenarios.euler_highway(z1, np.array([[0.0, -0.1 * z1[0, 3]]]))
        lat = {"p50_ms": float(np.percentile(ts[10:], 50) * 1e3), "p99_ms": float(np.percentile(ts[10:], 99) * 1e3),
               "what": "one warm highway solve through bmpc_solve_host, batch of 1, host buffers"}
        one.close()
    launch = parts[0]["mpc"].launch_info()
    for pt in parts:
        pt["mpc"].close()

    # ---- the other BASELINE configs, short runs on the same harness (default line only) ----
    others = None
    if which == "3" and not args.no_others:
        others = {}
        for oc in ("2", "4", "5", "cvar", "merge"):
            ob = DEFAULT_BATCH[oc] if oc != "cvar" else 4096
            op = make_workload(oc, ob, rank, world, local)
            ms, st, _, _ = timed_steps(op, 3, 3, flush, barrier)
            sm = summarize(op, st)
            tmax = shard.reduce_stats({"max_ms": float(sum(ms))}, device=dev)["max_ms"]
            nall = shard.reduce_stats({"solves": sm["solves_per_step"]}, device=dev)["solves"]
            others["cfg" + oc] = {"workload": (CONFIG_NAMES[oc] % ob), "solves_per_s": nall * 3 / (tmax * 1e-3),
                                  "ms_per_step": tmax / 3, "steps": 3, "warmup": 3,
                                  "status_counts[polished,converged,maxiter,numeric]": sm["status_counts"],
                                  "mean_factorizations": sm["mean_factorizations"], "mean_kkt_solves": sm["mean_kkt_solves"]}
            for pt in op:
                pt["mpc"].close()

    if rank == 0:
        peaks, peak_src = load_peaks()
        solve_ms = total_ms_max / K
        hbm_achieved = summ["bytes_per_step"] / (solve_ms * 1e-3) / 1e9
        fp64_peak = float(batch.abi.load_library().bmpc_measure_fp64_peak(local, 4096))
        fp64_achieved = summ["flops_per_step"] / (solve_ms * 1e-3) / 1e12
        cpu = cpu_exact = None
        if not args.no_cpu:
            cpu = cpu_reference_rate(problems_per_core=args.cpu_problems, warm_steps=1, style="osqp")
            cpu_exact = cpu_reference_rate(problems_per_core=2, warm_steps=1, style="exact")
            cpu["exact_qp_oracle"] = {"value": cpu_exact["value"], "unit": UNIT, "sample": cpu_exact["sample"]}
        workload = WORKLOAD if (which == "3" and B == 16384) else (
            "highway Branch MPC m=3 NB=2 N=8, closed-loop warm solves, %d episodes per GPU" % B if which == "3"
            else CONFIG_NAMES[which] % B)
        traffic = NCU_DRAM_BYTES_PER_LAUNCH.get((which, B))
        line = {
            "metric": METRIC if which in ("3", "5") else METRIC.replace("highway Branch-MPC", "Branch-MPC (config %s)" % which),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": solve_ms, "higher_is_better": True, "scaling": "strong" if which == "5" else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "baseline_config": which,
                       "batch_per_gpu": summ["solves_per_step"],
                       "parallelism": "episodes sharded over %d GPU(s), no collective on the solve path" % world,
                       "l2": "flushed between timed steps (256 MiB write, outside the event pairs)",
                       "status_counts[polished,converged,maxiter,numeric]": summ["status_counts"],
                       "mean_admm_iters": summ["mean_admm_iters"], "mean_factorizations": summ["mean_factorizations"],
                       "mean_kkt_solves": summ["mean_kkt_solves"]},
            "e2e": e2e, "e2e_full": e2e_full, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": hbm_achieved / peaks["hbm_gbs"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch from this round's ncu --set full capture
                         # (profiles/, see NCU_DRAM_BYTES_PER_LAUNCH); a plain run cannot read DRAM counters, so other sizes are null
                         "traffic": traffic, "traffic_source": NCU_SOURCE if traffic else None, "peak_source": peak_src,
                         "note": "latency/FP64-pipe bound, not HBM bound (arithmetic intensity >> ridge): see fp64",
                         "algorithmic_bytes_per_step": summ["bytes_per_step"],
                         "fp64": {"achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": fp64_achieved / fp64_peak if fp64_peak > 0 else None,
                                  "flops_per_solve": summ["flops_per_step"] / summ["solves_per_step"],
                                  "peak_source": "bmpc_measure_fp64_peak (DFMA loop, this GPU, this run)"}},
            "cpu_baseline": cpu, "latency": lat, "spot_check": check, "wall_s_timed_region": t_wall,
            "step_ms": [round(v, 3) for v in step_ms], "launch": launch,
            "extra": {"other_configs": others} if others else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


WORKLOAD = ("highway Branch MPC (BASELINE configs[2]): m=3 policies [maintain, brake, lane-change], NB=2, N=8 -> 13 branches / "
            "106 state nodes / 97 input nodes; closed-loop warm solves (updatetree path), 16384 episodes per GPU")
# dram__bytes_read.sum + dram__bytes_write.sum per launch of bmpc_solve_kernel<HighwayModel,3,...>, keyed by (config, episodes)
NCU_DRAM_BYTES_PER_LAUNCH = {("3", 16384): 105128704 + 20563456}
NCU_SOURCE = ("profiles/r02_solve_kernel_v22_ncu_raw.csv: ncu --set full of one warm launch over 16384 episodes (tools/gpu_ncu.sh), "
              "dram__bytes_read.sum 105.1 MB + dram__bytes_write.sum 20.6 MB; the launch's dirty lines are still in the 126 MB L2 "
              "when it ends, so the written share is below the algorithmic figure")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="episodes per GPU (config 5: in total); 0 = the config's BASELINE size")
    ap.add_argument("--config", default="3", choices=["2", "3", "4", "5", "cvar", "merge"], help="BASELINE.json config to run")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other configs")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-problems", type=int, default=16, help="episodes per host core in the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
