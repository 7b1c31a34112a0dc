#!/usr/bin/env python
"""Branch-MPC throughput benchmark (BASELINE.json metric: highway Branch-MPC solves/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Workload (N=1): BASELINE.json configs[2] - highway Branch MPC (m=3 policies, NB=2, N=8; 13 branches, 106 state nodes,
97 input nodes), 16384 independent episodes per GPU in closed loop: a *step* is one MPC solve of every episode
(tree update + linearisation + tree QP to the parity tolerances) followed by the Euler plant step that produces the
next step's initial states (both are this library's kernels).  The first (cold, inittree) solve of every episode is a
warm-up step; the timed steps are the updatetree path the reference spends 99 of its 100 solves per episode in.

  value   device-resident throughput: B*K / sum of per-step CUDA-event times (L2 flushed between steps, not timed)
  e2e     same metric through the host API (pinned host buffers -> bmpc_solve_host -> host results), copies included
  --impl reference   the CPU restatement of the reference path (oracle/, float64) on the host cores, bounded sample
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


def algorithmic_bytes_per_solve(t=N_TREE):
    """HBM bytes one warm solve has to move (float64): inputs, per-episode policy parameters, warm-start state read and
    written back, and the light outputs the bench requests."""
    n, d, m = t["n"], t["d"], t["m"]
    inputs = 3 * n * 8 + m * 4 * 8
    state = 2 * ((t["totalu"] + 1) * d * 8 + t["nbranch"] * 4 + d * 8 + 4)
    outputs = d * 8 + 8 + 4 * 4
    return inputs + state + outputs


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
    seed, count, warm_steps = args
    from _bmpc import scenarios
    from oracle import params
    x0, z0, xref, pp = scenarios.highway_batch(count, seed=seed)
    t_solve, n_solve = 0.0, 0
    for i in range(count):
        mpc = params.highway_branch_mpc(lc_target=pp[i, 2])
        x, z = x0[i], z0[i]
        for s in range(1 + warm_steps):
            t = time.perf_counter()
            u = mpc.solve(x, z, xref[i]).copy()
            dt = time.perf_counter() - t
            if s > 0 or warm_steps == 0:
                t_solve += dt
                n_solve += 1
            x = scenarios.euler_highway(x[None], u[None])[0]
            z = scenarios.euler_highway(z[None], np.array([[0.0, -0.1 * z[3]]]))[0]
    return t_solve, n_solve


def cpu_reference_rate(problems_per_core=2, warm_steps=1, cores=None, seed=4242):
    """solves/s of the oracle (float64 restatement of MPC_branch.BranchMPC.solve with an exact QP solve) with one worker
    process per host core; every worker runs `problems_per_core` episodes for 1 cold + `warm_steps` warm solves and the
    warm solves are timed (the same updatetree path the GPU arm times)."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(seed + 17 * c, problems_per_core, warm_steps) for c in range(cores)])
    wall = time.perf_counter() - t0
    n = sum(r[1] for r in res)
    busy = sum(r[0] for r in res)
    # all workers run concurrently: aggregate rate = solves / (mean busy time per worker)
    rate = n / (busy / cores) if busy > 0 else 0.0
    return {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d episodes x %d timed warm solves per core, %d cores, oracle.BranchMPCOracle (float64, exact QP); "
                      "per-solve mean %.3f s; wall %.1f s" % (problems_per_core, max(warm_steps, 1), cores,
                                                               busy / max(n, 1), wall)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rates = []
    t_all = time.perf_counter()
    last = None
    for s in range(args.warmup + args.steps):
        last = cpu_reference_rate(problems_per_core=1, warm_steps=1, seed=9000 + s)
        if s >= args.warmup:
            rates.append(last["value"])
    value = float(np.mean(rates)) if rates else 0.0
    B = (os.cpu_count() or 1)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (1e3 * B / value) if value else None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample": "bounded sample of that workload: each step = one warm closed-loop solve of one episode on every "
                                 "host core (CPU restatement of the reference path, oracle/; casadi/osqp are not "
                                 "installable here)"},
            "cpu_baseline": dict(last, value=value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
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

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    mpc = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B, device=local))
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=1237 + 1000 * rank)
    tx, tz, tr, tp = [torch.as_tensor(a, device=dev) for a in (x0, z0, xref, pp)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    stats = {"iters": [], "nfact": [], "nsolve": [], "status": []}

    def one_step(timed):
        flush.zero_()                      # evict L2 between steps (not inside the event pair)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        out = mpc.solve(tx, tz, tr, tp)
        mpc.plant_step(tx, out["u0"], tz, 0, tp)
        e1.record()
        if timed:
            for k in stats:
                stats[k].append(out[k].clone())
        return e0, e1

    for _ in range(W):
        one_step(False)
    barrier()
    launches0 = mpc.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t_wall = time.perf_counter()
    events = [one_step(True) for _ in range(K)]
    kernel_ms = []
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if rank == 0 else None
    launches = mpc.launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in events]
    total_ms = float(sum(step_ms))
    # the only communication of the run: max over ranks of the timed region (NCCL all_reduce MAX of one scalar)
    total_ms_max = shard.reduce_stats({"max_total_ms": total_ms}, device=dev)["max_total_ms"]
    value = world * B * K / (total_ms_max * 1e-3)

    it = torch.cat(stats["iters"]).double()
    nf = torch.cat(stats["nfact"]).double()
    ns = torch.cat(stats["nsolve"]).double()
    st = torch.cat(stats["status"])
    mean_nf, mean_ns = float(nf.mean()), float(ns.mean())
    status_counts = torch.bincount(st, minlength=4).tolist()

    # ---- end to end through the host API: pinned host inputs, bmpc_solve_host, host results ----
    e2e = None
    lat = None
    if rank == 0 or world > 1:
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
               "h2d_bytes_per_step": int(B * (3 * 4 + 3 * 4) * 8), "d2h_bytes_per_step": int(B * (2 * 8 + 8 + 4))}
    if rank == 0:
        # p50 latency of a single warm solve through the host API (batch of one)
        one = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=1, device=local))
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

    if rank == 0:
        peaks, peak_src = load_peaks()
        solve_ms = total_ms_max / K
        bytes_per_launch = algorithmic_bytes_per_solve() * B
        hbm_achieved = bytes_per_launch / (solve_ms * 1e-3) / 1e9
        flops_per_solve = N_TREE["totalu"] * (mean_nf * F_FACT_NODE + mean_ns * F_ITER_NODE)
        fp64_peak = float(batch.abi.load_library().bmpc_measure_fp64_peak(local, 4096))
        fp64_achieved = flops_per_solve * B / (solve_ms * 1e-3) / 1e12
        cpu = cpu_reference_rate(problems_per_core=args.cpu_problems, warm_steps=1) if not args.no_cpu else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": solve_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD if B == 16384 else
                                   "highway Branch MPC m=3 NB=2 N=8, closed-loop warm solves, %d episodes per GPU" % B,
                       "batch_per_gpu": B, "parallelism": "episodes sharded over %d GPU(s), no collective on the solve path" % world,
                       "l2": "flushed between timed steps (256 MiB write, outside the event pairs)",
                       "status_counts[polished,converged,maxiter,numeric]": status_counts,
                       "mean_admm_iters": float(it.mean()), "mean_factorizations": mean_nf, "mean_kkt_solves": mean_ns},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": hbm_achieved / peaks["hbm_gbs"],
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch over 16384 episodes from the ncu --set full
                         # capture profiles/r01_solve_kernel_v12_ncu_raw.csv (110.0 MB + 14.2 MB); not re-measured per run
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH_16384 if B == 16384 else None, "peak_source": peak_src,
                         "note": "latency/FP64-pipe bound, not HBM bound (arithmetic intensity >> ridge): see fp64",
                         "algorithmic_bytes_per_solve": algorithmic_bytes_per_solve(),
                         "fp64": {"achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": fp64_achieved / fp64_peak if fp64_peak > 0 else None,
                                  "flops_per_solve": flops_per_solve,
                                  "peak_source": "bmpc_measure_fp64_peak (DFMA loop, this GPU, this run)"}},
            "cpu_baseline": cpu, "latency": lat, "wall_s_timed_region": t_wall,
            "step_ms": [round(v, 3) for v in step_ms], "launch": mpc.launch_info(),
        }
        print(json.dumps(line), flush=True)
    mpc.close()
    if world > 1:
        dist.destroy_process_group()


WORKLOAD = ("highway Branch MPC (BASELINE configs[2]): m=3 policies [maintain, brake, lane-change], NB=2, N=8 -> 13 branches / "
            "106 state nodes / 97 input nodes; closed-loop warm solves (updatetree path), 16384 episodes per GPU")
NCU_DRAM_BYTES_PER_LAUNCH_16384 = 109988864 + 14200832


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16384, help="episodes per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-problems", type=int, default=4, help="episodes per host core in the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
