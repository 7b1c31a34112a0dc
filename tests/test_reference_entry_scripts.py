"""The reference's UNMODIFIED entry scripts and UNMODIFIED closed-loop environments driving this package.

`runpy` executes /root/reference/main_branch.py and main_quadruped.py as `__main__`.  Their imports resolve to
  * this package for the modules it replaces: MPC_branch, Init_MPC, highway_branch_dyn, quadruped_branch_dyn, utils;
  * the reference's own files, through symlinks made at test time, for everything else: Highway_env_branch.py and
    quadruped_env.py (vehicle plants, Highway_env.step, Highway_sim, the collision check);
  * tests/golden/shims for the third-party packages the reference imports but this image lacks (matplotlib, osqp, ...).
Reference-side edits: none to any file; the only patched name is `animate_scenario` (matplotlib animation of the finished
run, Highway_env_branch.py:566-709), replaced by a recorder of its arguments.

The scripts need /root/reference, which exists in the build container only, and the build container has no GPU: the
drop-in classes are therefore backed by the single-lane host build of the kernel's solver text (tests/hostsim) instead of
libbranchmpc.so.  Everything above the C ABI - the classes main_branch.py constructs and the attributes the reference
environment reads and writes every step - is the shipped code.  tests/test_gpu_dropin.py runs the same construction
sequence on the B200.
"""
import os
import runpy
import sys

import numpy as np
import pytest

from tests.helpers import PKG, ROOT

REFERENCE = os.environ.get("BMPC_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REFERENCE, "main_branch.py")),
                                reason="the reference tree is only present in the build container")


@pytest.fixture
def reference_world(tmp_path, monkeypatch):
    """sys.path: symlinks to the reference's env modules, then this package, then the third-party shims."""
    from _bmpc import batch
    from tests.hostsim.backend import HostBackend
    for name in ("Highway_env_branch.py", "quadruped_env.py"):
        os.symlink(os.path.join(REFERENCE, name), tmp_path / name)
    shims = os.path.join(ROOT, "tests", "golden", "shims")
    monkeypatch.setattr(sys, "path", [str(tmp_path), PKG, shims] + [p for p in sys.path if p not in (PKG, shims)])
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    monkeypatch.setattr(batch, "BatchedBranchMPC", HostBackend)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("Highway_env_branch", "quadruped_env", "MPC_branch", "Init_MPC", "highway_branch_dyn",
                      "quadruped_branch_dyn", "utils", "MPC_nobranch")}
    yield
    for k in ("Highway_env_branch", "quadruped_env", "MPC_branch", "Init_MPC", "highway_branch_dyn", "quadruped_branch_dyn",
              "utils"):
        sys.modules.pop(k, None)
    sys.modules.update(saved)


def test_main_branch_runs_unmodified(reference_world):
    """main_branch.py:20-51 -> BranchMPC_CVaR(ralpha=0.9) -> Highway_env_branch.sim_overtake: 100 closed-loop steps of the
    reference's own environment around this package's controller; the ego overtakes without a collision."""
    import Highway_env_branch as henv
    assert henv.__file__.startswith(os.path.dirname(str(henv.__file__))) and "reference" in os.path.realpath(henv.__file__)
    rec = {}

    def record(env, state_rec, backup_rec, backup_choice_rec, xPred_rec, zPred_rec, lm, *a, **k):
        rec.update(env=env, state=np.array(state_rec), choice=np.array(backup_choice_rec))

    henv.animate_scenario = record
    sim = henv.Highway_sim
    out = {}

    def spy(env, T):
        r = sim(env, T)
        out["collision"] = r[-1]
        out["input"] = np.array(r[1])
        return r

    henv.Highway_sim = spy
    np.random.seed(3)
    runpy.run_path(os.path.join(REFERENCE, "main_branch.py"), run_name="__main__")
    import MPC_branch
    mpc = rec["env"].mpc
    assert os.path.realpath(MPC_branch.__file__).startswith(os.path.realpath(PKG))
    assert type(mpc) is MPC_branch.BranchMPC_CVaR and mpc.ralpha == 0.9
    assert mpc.timeStep == 100 and mpc.feasible == 1
    assert mpc.uPred.shape == (97, 2) and mpc.xPred.shape == (106, 4)
    state = rec["state"]                      # (NV, steps, 4): Highway_sim's state_rec (Highway_env_branch.py:403)
    assert state.shape == (2, 100, 4) and np.isfinite(state).all()
    assert not out["collision"]
    assert state[0, -1, 0] > state[1, -1, 0], "the ego ends ahead of the obstacle (overtake)"
    u_ego = out["input"][0]
    assert u_ego.shape == (100, 2)
    assert (np.abs(u_ego[:, 0]) <= 6.0 + 1e-9).all() and (np.abs(u_ego[:, 1]) <= 0.3 + 1e-9).all()


def test_main_quadruped_runs_as_the_reference_does(reference_world):
    """main_quadruped.py:10-43 -> BranchMPCProx -> quadruped_env.sim.  As shipped the reference stops right after its first
    solve: Quad_env.step unpacks three values from BT2array(), which returns four (quadruped_env.py:120 vs
    MPC_branch.py:459).  The drop-in reproduces exactly that: the first solve completes, then the reference's own unpack raises."""
    import quadruped_env as qenv
    seen = {}
    import MPC_branch
    solve = MPC_branch.BranchMPCProx.solve

    def spy(self, *a, **k):
        r = solve(self, *a, **k)
        seen["mpc"] = self
        return r

    MPC_branch.BranchMPCProx.solve = spy
    try:
        with pytest.raises(ValueError, match="too many values to unpack"):
            runpy.run_path(os.path.join(REFERENCE, "main_quadruped.py"), run_name="__main__")
    finally:
        MPC_branch.BranchMPCProx.solve = solve
    mpc = seen["mpc"]
    assert mpc.timeStep == 1 and mpc.feasible == 1
    assert mpc.uPred.shape == (151, 3) and mpc.xPred.shape == (155, 3)
    assert len(mpc.BT2array()) == 4
    assert "reference" in os.path.realpath(qenv.__file__)


def _run_sim_merge(tmp_path, henv):
    """`sim_merge()` of the unmodified main_branch.py (:53-88; the script's __main__ block has the call commented out) with
    the animation replaced by a no-op; returns (Highway_sim records, the controller)."""
    os.symlink(os.path.join(REFERENCE, "main_branch.py"), tmp_path / "main_branch.py")
    sys.modules.pop("main_branch", None)
    import main_branch
    import MPC_branch
    got = {}
    sim = henv.Highway_sim

    def spy(env, T):
        got["rec"] = sim(env, T)
        got["env"] = env
        return got["rec"]

    henv.Highway_sim = spy
    if hasattr(henv, "animate_scenario"):
        henv.animate_scenario = lambda *a, **k: None
    try:
        main_branch.sim_merge()
    finally:
        henv.Highway_sim = sim
        sys.modules.pop("main_branch", None)
    mpc = got["env"].mpc
    assert os.path.realpath(MPC_branch.__file__).startswith(os.path.realpath(PKG))
    assert type(mpc) is MPC_branch.BranchMPC_CVaR and mpc.ralpha == 0.1 and mpc.N == 40 and mpc.NB == 1
    assert type(mpc.predictiveModel).__name__ == "PredictiveModel_merge"
    return got["rec"], mpc, got["env"]


def _check_merge_run(rec, mpc, env):
    state, inputs, collision = np.array(rec[0]), np.array(rec[1]), rec[-1]
    assert state.shape == (2, 60, 4) and np.isfinite(state).all() and inputs.shape == (2, 60, 2)
    assert mpc.timeStep == 60 and mpc.feasible == 1
    assert mpc.uPred.shape == (81, 2) and mpc.xPred.shape == (83, 4)
    assert (np.abs(inputs[0, :, 0]) <= 7.0 + 1e-9).all() and (np.abs(inputs[0, :, 1]) <= 0.3 + 1e-9).all()
    assert not np.any(collision)
    assert env.laneID == [0, 0], "the ego has left the ramp"
    # the ego came down from the ramp (y = 13) into the highway's lanes and drives along them
    assert 1.25 - 0.5 < state[0, -1, 1] < 7.2 and abs(state[0, -1, 3]) < 0.3
    # the first controller call has no warm start, so it is the reference's call exactly: the recorded first input
    g = np.load(os.path.join(ROOT, "tests", "golden", "highway_merge_default.npz"))
    np.testing.assert_allclose(inputs[0, 0], g["input_rec"][0, 0], atol=1e-3)
    np.testing.assert_allclose(state[:, 0], g["state_rec"][:, 0], atol=1e-4)


def test_sim_merge_runs_unmodified_on_the_reference_environment(reference_world, tmp_path):
    """main_branch.sim_merge -> PredictiveModel_merge x 2 (one with lookup-table policies), BranchMPC_CVaR(ralpha=0.1),
    the reference's own Highway_env_merge.step calling mpc.solve(x, z, xRef, S, Fx=None, bx=bx) for 60 steps."""
    import Highway_env_branch as henv
    assert "reference" in os.path.realpath(henv.__file__)
    rec, mpc, env = _run_sim_merge(tmp_path, henv)
    _check_merge_run(rec, mpc, env)


def test_sim_merge_runs_unmodified_on_the_dropin_environment(reference_world, tmp_path):
    """The same script with this package's Highway_env_branch (merge_geometry, Highway_env_merge, sim_merge) on the path."""
    os.remove(tmp_path / "Highway_env_branch.py")
    sys.modules.pop("Highway_env_branch", None)
    import Highway_env_branch as henv
    assert os.path.realpath(henv.__file__).startswith(os.path.realpath(PKG))
    rec, mpc, env = _run_sim_merge(tmp_path, henv)
    _check_merge_run(rec, mpc, env)
