"""Parity tests proper: libbranchmpc.so on the B200, called through the C ABI, against
  * the golden fixtures produced by the unmodified reference (tests/golden/*.npz),
  * the oracle on seeded random problems (sizes the oracle finishes in seconds),
  * size-independent properties at BASELINE.json's full batch (16384 episodes).

Bars (BASELINE.json north_star): first applied control within 1e-3, objective relative difference <= 1e-4,
constraint violation <= 1e-5, bit-exact tree topology and branch indexing.
"""
import numpy as np
import pytest

from tests.helpers import (check_quadruped_hard_cases, check_forced_interior_point, SWEEP, check_sweep_case, check_robust_fixture, robust_fixture_config, quadruped_fixture_config, HIGHWAY_FIXTURES, TOL_OBJ, TOL_U0, TOL_VIOL, check_fixture_closed_loop, fixture_config,
                           load_fixture, oracle_episode)
from _bmpc import abi, scenarios
from oracle.branch_mpc import TreeTopology

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bmpc():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    from _bmpc import batch
    return batch


@pytest.mark.parametrize("m,NB", [(2, 1), (3, 1), (4, 1), (2, 2), (3, 2), (4, 2), (2, 3), (3, 3), (4, 3)])
def test_topology_is_bit_exact(bmpc, m, NB):
    names = ["maintain", "brake", "lc", "trackv"][:m]
    mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(policies=names, NB=NB))
    T = TreeTopology(m, NB, 8)
    assert (mpc.topology() == T.table()).all()
    assert (mpc.totalx, mpc.totalu, mpc.nbranch) == (T.totalx, T.totalu, T.nbranch)
    mpc.close()


def test_topology_matches_reference_fixture(bmpc):
    for name in HIGHWAY_FIXTURES:
        g = load_fixture(name)
        mpc = bmpc.BatchedBranchMPC(fixture_config(g))
        assert (mpc.topology() == g["s0_tree"]).all()
        assert [mpc.totalx, mpc.totalu] == list(g["s0_totals"])
        mpc.close()


def test_model_functions_match_reference(bmpc):
    """Rows M1-M5: dyn_linearization, zpred_eval, branch_eval, col_eval vs the reference's CasADi graphs."""
    g = load_fixture("model_functions")
    mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(lc_target=tuple(g["hw_lc_target"])))
    r = mpc.eval_model(g["hw_X"], g["hw_Z"], g["hw_U"])
    for k in ("A", "B", "C", "xp", "zpred", "p", "dh"):
        np.testing.assert_allclose(r[k], g["hw_" + k], atol=1e-11, err_msg=k)
    np.testing.assert_allclose(r["hlin"], g["hw_hlin"], atol=1e-10)
    assert np.abs(r["p"].sum(axis=1) - 1.0).max() < 1e-12
    mpc.close()


@pytest.mark.parametrize("name", HIGHWAY_FIXTURES)
def test_fixture_closed_loop(bmpc, name):
    g = load_fixture(name)
    mpc = bmpc.BatchedBranchMPC(fixture_config(g))
    check_fixture_closed_loop(lambda x, z, r: mpc.solve_host(x, z, r), g)
    mpc.close()


def test_quadruped_prox_fixture_closed_loop(bmpc):
    g = load_fixture("quadruped_prox_default")
    mpc = bmpc.BatchedBranchMPC(quadruped_fixture_config(g))
    assert (mpc.topology() == g["s0_tree"]).all() and [mpc.totalx, mpc.totalu] == list(g["s0_totals"])
    check_fixture_closed_loop(lambda x, z, r: mpc.solve_host(x, z, r), g, tol=5e-5)   # see tests/test_hostsim_parity.py
    mpc.close()


def test_quadruped_model_functions_match_reference(bmpc):
    g = load_fixture("model_functions")
    mpc = bmpc.BatchedBranchMPC(scenarios.quadruped_config())
    r = mpc.eval_model(g["qd_X"], g["qd_Z"], g["qd_U"])
    for k in ("A", "B", "C", "xp", "zpred", "p", "dh"):
        np.testing.assert_allclose(r[k], g["qd_" + k], atol=1e-11, err_msg=k)
    np.testing.assert_allclose(r["hlin"], g["qd_hlin"], atol=1e-10)
    mpc.close()


def test_robust_chain_fixture_closed_loop(bmpc):
    g = load_fixture("highway_robust_default")
    mpc = bmpc.BatchedBranchMPC(robust_fixture_config(g))
    assert (mpc.totalx, mpc.totalu) == (18, 17)
    check_robust_fixture(lambda x, z, r: mpc.solve_host(x, z, r), g)
    mpc.close()


def test_robust_batch_4096(bmpc):
    """BASELINE config 2: 4096-episode batch of the single-trajectory controller; sample checked against the oracle."""
    from oracle import params
    B = 4096
    cfg = scenarios.highway_config(batch_capacity=B)
    cfg.controller = abi.CTRL_ROBUST
    mpc = bmpc.BatchedBranchMPC(cfg)
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=5)
    r = mpc.solve_host(x0, z0, xref, pp, outputs=("u0", "uPred", "xPred", "objective", "status", "iters"))
    assert (r["status"] <= abi.STATUS_CONVERGED).all() and np.isfinite(r["objective"]).all()
    assert (np.abs(r["uPred"][:, :, 0]) <= 6.0 + 1e-12).all() and (np.abs(r["uPred"][:, :, 1]) <= 0.3 + 1e-12).all()
    for i in (0, 1000, 4095):
        ora = params.highway_robust_mpc(lc_target=pp[i, 2])
        u = ora.solve(x0[i], z0[i], xref[i])
        assert np.abs(r["u0"][i] - u).max() < TOL_U0
        assert abs(r["objective"][i] - ora.objective) <= TOL_OBJ * abs(ora.objective)
    mpc.close()


def test_random_batch_against_oracle(bmpc):
    B, steps = 32, 2
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=2024)
    mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
    ref = [oracle_episode(x0[i], z0[i], xref[i], pp[i, 2], steps) for i in range(B)]
    x, z = x0.copy(), z0.copy()
    for s in range(steps):
        for i in range(B):
            x[i], z[i] = ref[i][s][0], ref[i][s][1]
        r = mpc.solve_host(x, z, xref, pp)
        du = max(np.abs(r["u0"][i] - ref[i][s][2]).max() for i in range(B))
        dj = max(abs(r["objective"][i] - ref[i][s][3]) / abs(ref[i][s][3]) for i in range(B))
        assert du < TOL_U0 and dj < TOL_OBJ, (s, du, dj)
        # continue from the oracle's solution so that step s+1 linearises about the same trajectory
        st = mpc.get_state(B)
        for i in range(B):
            st["uLin"][i, :-1] = ref[i][s][4]
            st["uLin"][i, -1] = ref[i][s][4][-1]
            st["old_input"][i] = ref[i][s][2]
        mpc.set_state(st)
    mpc.close()


def _violations(r, x0, cfg):
    """max violation of the QP's hard constraints by the returned plan: initial state, input box, shared branching
    states; the dynamics are checked through the reference's structure xPred[child first] == xPred[sibling first]."""
    u = r["uPred"]
    v = max(0.0, float((u[:, :, 0] - cfg.u_hi[0]).max()), float((cfg.u_lo[0] - u[:, :, 0]).max()),
            float((u[:, :, 1] - cfg.u_hi[1]).max()), float((cfg.u_lo[1] - u[:, :, 1]).max()))
    v = max(v, float(np.abs(r["xPred"][:, 0] - x0).max()))
    return v


def test_full_batch_properties(bmpc):
    import torch
    B = 16384
    cfg = scenarios.highway_config(batch_capacity=B)
    mpc = bmpc.BatchedBranchMPC(cfg)
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=31)
    r = mpc.solve_host(x0, z0, xref, pp)
    assert (r["status"] <= abi.STATUS_MAXITER).all()
    assert (r["status"] == abi.STATUS_POLISHED).mean() > 0.99
    assert np.isfinite(r["objective"]).all() and np.isfinite(r["xPred"]).all()
    assert _violations(r, x0, cfg) <= TOL_VIOL
    # all children of one branch start from the same state (MPC_branch.py:1007-1012), bit for bit
    topo = mpc.topology()
    for b in range(mpc.nbranch):
        kids = topo[topo[:, 4] == b]
        for c in kids[1:]:
            assert np.array_equal(r["xPred"][:, kids[0][2]], r["xPred"][:, c[2]])
    w = r["branch_w"]
    assert np.abs(w[:, 1:4].sum(axis=1) - 1.0).max() < 1e-12 and np.abs(w[:, 4:].sum(axis=1) - 1.0).max() < 1e-12
    # permutation invariance: problems are independent, so a shuffled batch gives the shuffled results bit for bit
    perm = np.random.default_rng(0).permutation(B)
    mpc.reset()
    r2 = mpc.solve_host(x0[perm], z0[perm], xref[perm], pp[perm])
    for k in ("u0", "uPred", "xPred", "objective", "status", "iters"):
        assert np.array_equal(r2[k], r[k][perm]), k
    # device-pointer entry point == host entry point
    mpc.reset()
    dev = torch.device("cuda", 0)
    out = mpc.solve(*[torch.as_tensor(a, device=dev) for a in (x0, z0, xref, pp)])
    torch.cuda.synchronize()
    assert np.array_equal(out["u0"].cpu().numpy(), r["u0"])
    # warm step: state persisted per episode slot; reset(ids) makes exactly those episodes cold again
    x1 = scenarios.euler_highway(x0, r["u0"])
    rw = mpc.solve_host(x1, z0, xref, pp)
    assert (rw["status"] <= abi.STATUS_MAXITER).all()
    ids = np.arange(0, B, 2)
    mpc.reset(ids)
    rc = mpc.solve_host(x0, z0, xref, pp)
    assert np.array_equal(rc["u0"][ids], r["u0"][ids])
    mpc.close()


@pytest.mark.parametrize("m,NB", SWEEP)
def test_tree_sweep_against_oracle(bmpc, m, NB):
    """BASELINE config 5 shapes; (3,3) and (4,3) do not fit shared memory and run from the global/L2 slab."""
    def make(cfg):
        mpc = bmpc.BatchedBranchMPC(cfg)
        return lambda *a: mpc.solve_host(*a)
    check_sweep_case(make, m, NB)


@pytest.mark.parametrize("m,NB", SWEEP)
def test_tree_sweep_full_size_properties(bmpc, m, NB):
    """BASELINE config 5 at its per-shape share of the 65536 episodes (7281): every problem certified or converged, hard
    constraints hold, siblings share their first state bit for bit, weights of every level sum to one, and a shuffled
    batch gives the shuffled results bit for bit (episodes are independent: what makes the sharding over GPUs exact)."""
    B = 65536 // 9
    names = ["maintain", "brake", "lc", "trackv"][:m]
    cfg = scenarios.highway_config(policies=names, NB=NB, batch_capacity=B)
    mpc = bmpc.BatchedBranchMPC(cfg)
    x0, z0, xref, pp3 = scenarios.highway_batch(B, seed=1239 + 10 * m + NB)
    pp = np.zeros((B, m, 4))
    if m >= 3:
        pp[:, 2, :] = pp3[:, 2, :]
    if m >= 4:
        pp[:, 3, 0] = 20.0
    outs = ("u0", "uPred", "xPred", "branch_w", "objective", "status")
    r = mpc.solve_host(x0, z0, xref, pp, outputs=outs)
    assert (r["status"] <= abi.STATUS_CONVERGED).all() and (r["status"] == abi.STATUS_POLISHED).mean() > 0.99
    assert np.isfinite(r["objective"]).all() and np.isfinite(r["xPred"]).all()
    assert _violations(r, x0, cfg) <= TOL_VIOL
    topo = mpc.topology()
    for b in range(mpc.nbranch):
        kids = topo[topo[:, 4] == b]
        for c in kids[1:]:
            assert np.array_equal(r["xPred"][:, kids[0][2]], r["xPred"][:, c[2]])
    for d in range(1, NB + 1):
        level = topo[topo[:, 1] == d][:, 0]
        assert np.abs(r["branch_w"][:, level].sum(axis=1) - 1.0).max() < 1e-12
    perm = np.random.default_rng(m * 10 + NB).permutation(B)
    mpc.reset()
    r2 = mpc.solve_host(x0[perm], z0[perm], xref[perm], pp[perm], outputs=outs)
    for k in ("u0", "objective", "status"):
        assert np.array_equal(r2[k], r[k][perm]), k
    mpc.close()


def _gpu_solver(bmpc):
    def make(cfg):
        mpc = bmpc.BatchedBranchMPC(cfg)
        return lambda *a: mpc.solve_host(*a)
    return make


def test_interior_point_rescues_degenerate_quadruped_problems(bmpc):
    check_quadruped_hard_cases(_gpu_solver(bmpc))


def test_forced_interior_point_matches_oracle(bmpc):
    check_forced_interior_point(_gpu_solver(bmpc))


def test_quadruped_batch_8192(bmpc):
    """BASELINE config 4 at full size: every problem certified or converged by the interior point, sample vs the oracle."""
    from oracle import params
    B = 8192
    mpc = bmpc.BatchedBranchMPC(scenarios.quadruped_config(batch_capacity=B))
    x0, z0, xref = scenarios.quadruped_batch(B, seed=1238)
    r = mpc.solve_host(x0, z0, xref)
    assert (r["status"] <= abi.STATUS_CONVERGED).all() and (r["status"] == abi.STATUS_POLISHED).mean() > 0.98
    assert np.isfinite(r["objective"]).all()
    soft = np.where(r["status"] == abi.STATUS_CONVERGED)[0]
    for i in list(soft[:3]) + [0, 4095, 8191]:
        ora = params.quadruped_prox_mpc()
        u = ora.solve(x0[i], z0[i], xref[i])
        assert np.abs(r["u0"][i] - u).max() < TOL_U0, i
        assert abs(r["objective"][i] - ora.objective) <= TOL_OBJ * abs(ora.objective), i
    mpc.close()


def test_two_handles_on_two_streams(bmpc):
    """The solve kernel reads its parameter block from one constant-memory symbol per device: launches of different handles
    on different streams must be ordered by the library so that neither sees the other's block."""
    import torch
    dev = torch.device("cuda", 0)
    B = 2048
    hx, hz, hr, hp = scenarios.highway_batch(B, seed=8)
    qx, qz, qr = scenarios.quadruped_batch(B, seed=9)
    a = bmpc.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
    b = bmpc.BatchedBranchMPC(scenarios.quadruped_config(batch_capacity=B))
    ref_a = a.solve_host(hx, hz, hr, hp, outputs=("u0", "objective", "status"))
    ref_b = b.solve_host(qx, qz, qr, outputs=("u0", "objective", "status"))
    a.reset()
    b.reset()
    ta = [torch.as_tensor(v, device=dev) for v in (hx, hz, hr, hp)]
    tb = [torch.as_tensor(v, device=dev) for v in (qx, qz, qr)]
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        oa = a.solve(*ta, outputs=("u0", "objective", "status"), stream=s1.cuda_stream)
    with torch.cuda.stream(s2):
        ob = b.solve(*tb, outputs=("u0", "objective", "status"), stream=s2.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(oa["u0"].cpu().numpy(), ref_a["u0"]) and np.array_equal(oa["status"].cpu().numpy(), ref_a["status"])
    assert np.array_equal(ob["u0"].cpu().numpy(), ref_b["u0"]) and np.array_equal(ob["status"].cpu().numpy(), ref_b["status"])
    a.close()
    b.close()


def test_plant_step_matches_reference_plant(bmpc):
    import torch
    B = 1000
    mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
    x0, z0, _, pp = scenarios.highway_batch(B, seed=3)
    u = np.random.default_rng(1).uniform(-1, 1, (B, 2))
    dev = torch.device("cuda", 0)
    tx, tz, tu, tp = [torch.as_tensor(a, device=dev) for a in (x0, z0, u, pp)]
    mpc.plant_step(tx, tu, tz, 0, tp)
    torch.cuda.synchronize()
    np.testing.assert_allclose(tx.cpu().numpy(), scenarios.euler_highway(x0, u), atol=1e-12)
    np.testing.assert_allclose(tz.cpu().numpy(), scenarios.euler_highway(z0, np.column_stack([np.zeros(B), -0.1 * z0[:, 3]])),
                               atol=1e-12)
    mpc.close()


def test_errors_are_codes_not_crashes(bmpc):
    mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(batch_capacity=2))
    x = np.zeros((3, 4))
    with pytest.raises(bmpc.BmpcError, match="capacity"):
        mpc.solve_host(x, x, x)
    with pytest.raises(ValueError):
        mpc.solve_host(np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3)))
    # non-finite input: status NUMERIC, no crash, other episodes unaffected
    x0, z0, xref, pp = scenarios.highway_batch(2, seed=1)
    good = mpc.solve_host(x0, z0, xref, pp)
    mpc.reset()
    x0[1, 2] = np.nan
    r = mpc.solve_host(x0, z0, xref, pp)
    assert r["status"][1] == abi.STATUS_NUMERIC and r["status"][0] == good["status"][0]
    assert np.array_equal(r["u0"][0], good["u0"][0])
    mpc.close()


def test_smoke_entry(bmpc):
    import __graft_entry__ as entry
    entry.smoke()


def test_solve_is_graph_capturable(bmpc):
    """bmpc_solve + bmpc_plant_step recorded into a CUDA graph and replayed: the closed loop advances exactly like the
    eagerly launched one (inputs, outputs and the persistent warm-start state all live at fixed device addresses)."""
    import torch
    B, steps = 512, 4
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=77)
    dev = torch.device("cuda", 0)

    def run(graphed):
        mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
        tx, tz, tr, tp = [torch.as_tensor(a, device=dev).clone() for a in (x0, z0, xref, pp)]
        outs = ("u0", "objective", "status")
        hist = []
        if graphed:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                out = mpc.solve(tx, tz, tr, tp, outputs=outs, stream=s.cuda_stream)     # allocate outputs, warm the caches
                mpc.plant_step(tx, out["u0"], tz, 0, tp, stream=s.cuda_stream)
                s.synchronize()
                hist.append((out["u0"].cpu().numpy().copy(), out["status"].cpu().numpy().copy()))
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    out = mpc.solve(tx, tz, tr, tp, outputs=outs, stream=torch.cuda.current_stream().cuda_stream)
                    mpc.plant_step(tx, out["u0"], tz, 0, tp, stream=torch.cuda.current_stream().cuda_stream)
                for _ in range(steps - 1):
                    g.replay()
                    torch.cuda.synchronize()
                    hist.append((out["u0"].cpu().numpy().copy(), out["status"].cpu().numpy().copy()))
        else:
            for _ in range(steps):
                out = mpc.solve(tx, tz, tr, tp, outputs=outs)
                mpc.plant_step(tx, out["u0"], tz, 0, tp)
                torch.cuda.synchronize()
                hist.append((out["u0"].cpu().numpy().copy(), out["status"].cpu().numpy().copy()))
        mpc.close()
        return hist

    eager, graph = run(False), run(True)
    for (u_e, s_e), (u_g, s_g) in zip(eager, graph):
        assert np.array_equal(s_e, s_g) and (s_e <= abi.STATUS_CONVERGED).all()
        assert np.array_equal(u_e, u_g)


def test_reset_of_many_episodes_is_one_kernel(bmpc):
    """bmpc_reset(ids) is one launch (one block per listed slot), not five driver calls per id: 8192 ids in well under a
    millisecond of device time, and the reset episodes solve cold again while the others stay warm."""
    import time
    import torch
    B = 16384
    mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=3)
    r0 = mpc.solve_host(x0, z0, xref, pp, outputs=("u0", "status"))
    ids = np.arange(0, B, 2)
    n0 = mpc.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mpc.reset(ids)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert mpc.launch_count() - n0 == 1
    assert dt < 5e-3, dt            # host-side wall time incl. the 64 KB id upload; the kernel itself is ~10 us
    st = mpc.get_state(B)
    assert (st["started"][ids] == 0).all() and (st["started"][1::2] == 1).all()
    assert (st["uLin"][ids] == 0).all() and np.abs(st["uLin"][1::2]).max() > 0
    r1 = mpc.solve_host(x0, z0, xref, pp, outputs=("u0", "status"))
    assert np.array_equal(r1["u0"][ids], r0["u0"][ids])          # cold again: same answer as the first cold solve
    mpc.close()


def test_handles_with_different_trees_share_a_kernel_instance(bmpc):
    """Two handles of the same kernel instance (highway, two state rows) with a small and a large tree: the launch of the
    large one must not inherit the dynamic shared-memory limit the small one set (regression: round 2)."""
    small = bmpc.BatchedBranchMPC(scenarios.highway_config(NB=1, batch_capacity=4))
    large = bmpc.BatchedBranchMPC(scenarios.highway_config(NB=3, batch_capacity=4))
    tiny = bmpc.BatchedBranchMPC(scenarios.highway_config(NB=1, batch_capacity=4))      # created last: lowers the limit
    x0, z0, xref, pp = scenarios.highway_batch(4, seed=9)
    for h in (large, small, tiny, large):
        r = h.solve_host(x0, z0, xref, pp, outputs=("u0", "status"))
        assert (r["status"] <= abi.STATUS_CONVERGED).all()
    for h in (small, large, tiny):
        h.close()


def test_staged_episode_state_gives_identical_results(bmpc):
    """The next episode's uLin / active-set codes / rho cache staged into shared memory by bulk copies (cp.async.bulk +
    mbarrier) while the current episode is solved: same arithmetic, so three closed-loop steps of 4096 episodes must agree
    bit for bit with the handle that reads the same data from global memory.  Staging is opt-in (reserved[6] bit 1): it
    measured slower than the plain loads (profiles/r02_staging_ab.md)."""
    import torch
    B = 4096
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=99)
    runs = []
    for flag in (2, 0):
        cfg = scenarios.highway_config(batch_capacity=B)
        cfg.reserved[6] = flag
        mpc = bmpc.BatchedBranchMPC(cfg)
        assert mpc.staging_enabled() == (flag == 2)
        t = [torch.as_tensor(a.copy(), device="cuda") for a in (x0, z0, xref, pp)]
        outs = []
        for _ in range(3):
            r = mpc.solve(t[0], t[1], t[2], t[3], outputs=("u0", "objective", "status", "iters", "nfact", "nsolve"))
            outs.append({k: v.clone() for k, v in r.items()})
            mpc.plant_step(t[0], r["u0"], t[1], 0, t[3])
        torch.cuda.synchronize()
        runs.append(outs)
        mpc.close()
    for a, b in zip(*runs):
        assert int((a["status"] <= 1).sum()) == B
        for k in a:
            assert torch.equal(a[k], b[k]), k
    chain = scenarios.highway_config(batch_capacity=8)
    chain.controller = abi.CTRL_ROBUST
    chain.reserved[6] = 2
    ch = bmpc.BatchedBranchMPC(chain)
    assert not ch.staging_enabled()              # chain controllers read their state from global memory
    ch.close()


def test_warm_polish_skip_does_not_change_the_plans(bmpc):
    """An episode whose warm-polish attempt ended on the ADMM path skips the attempt on its next solves (reserved[0] bits 4..7,
    doubling with bit 2 clear) and the attempt is also made on rho-refresh solves (bit 1 clear).  Either way the solve ends on
    the certified optimum of the same QP: 14 closed-loop steps of 2048 episodes with the default schedule against the handle
    that tries the warm polish on every solve (0xF0 | 2 | 4), every step from the same states."""
    import torch
    B = 2048
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=123)
    handles = []
    for flags in (0, 0xF0 | 2 | 4):
        cfg = scenarios.highway_config(batch_capacity=B)
        cfg.reserved[0] = flags
        handles.append(bmpc.BatchedBranchMPC(cfg))
    t = [torch.as_tensor(a.copy(), device="cuda") for a in (x0, z0, xref, pp)]
    skipped_some = False
    for step in range(14):
        res = [h.solve(t[0], t[1], t[2], t[3], outputs=("u0", "objective", "status", "iters", "nfact")) for h in handles]
        res = [{k: v.clone() for k, v in r.items()} for r in res]
        a, b = res
        assert int((a["status"] <= 1).sum()) == B and int((b["status"] <= 1).sum()) == B
        du = (a["u0"] - b["u0"]).abs().amax(dim=1)
        dj = (a["objective"] - b["objective"]).abs() / b["objective"].abs().clamp_min(1.0)
        assert du.max().item() < TOL_U0 and dj.max().item() < TOL_OBJ, (step, du.max().item(), dj.max().item())
        both = (a["status"] == 0) & (b["status"] == 0)      # certified by the polish on both sides: the same vertex
        # (the polish certifies stationarity to 1e-7: flat directions of the cost leave the inputs free to ~1e-5)
        assert du[both].max().item() < 1e-4 and dj[both].max().item() < 1e-6, (step, du[both].max().item(), dj[both].max().item())
        if step > 2:
            skipped_some |= bool((a["nfact"] < b["nfact"]).any().item())
        handles[0].plant_step(t[0], b["u0"], t[1], 0, t[3])      # both handles continue from the same plant state
    assert skipped_some      # the default schedule did save factorisations somewhere
    for h in handles:
        h.close()


def test_host_solve_from_pinned_and_pageable_arrays_agree(bmpc):
    """bmpc_solve_host sends page-locked caller arrays straight to the device (one DMA each) and gathers pageable ones into
    its pinned block first: same inputs, same results, bit for bit, over three closed-loop steps."""
    import torch
    B = 1024
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=321)
    runs = []
    for pinned in (True, False):
        arrs = [torch.as_tensor(a.copy()).pin_memory().numpy() if pinned else a.copy() for a in (x0, z0, xref, pp)]
        mpc = bmpc.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
        outs = []
        for _ in range(3):
            r = mpc.solve_host(*arrs, outputs=("u0", "uPred", "xPred", "objective", "status"))
            outs.append({k: v.copy() for k, v in r.items()})
            arrs[0][:] = scenarios.euler_highway(arrs[0], r["u0"])
        runs.append(outs)
        mpc.close()
    for a, b in zip(*runs):
        assert (a["status"] <= 1).all()
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_host_results_written_in_place_match_the_copied_ones(bmpc):
    """bmpc_solve_host lets the kernel write the requested results straight into the pinned host block (default) or packs
    them into a device block that comes back in one DMA (reserved[6] bit 2): every output of three closed-loop steps must
    agree bit for bit."""
    B = 1500
    x0, z0, xref, pp = scenarios.highway_batch(B, seed=78)
    runs = []
    for flag in (0, 4):
        cfg = scenarios.highway_config(batch_capacity=B)
        cfg.reserved[6] = flag
        mpc = bmpc.BatchedBranchMPC(cfg)
        x = x0.copy()
        outs = []
        for _ in range(3):
            r = mpc.solve_host(x, z0, xref, pp)
            outs.append({k: v.copy() for k, v in r.items() if k != "cycles"})
            x = scenarios.euler_highway(x, r["u0"])
        runs.append(outs)
        mpc.close()
    for a, b in zip(*runs):
        assert (a["status"] <= 1).all()
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
