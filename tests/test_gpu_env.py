"""Closed-loop environment on the device (row f1): bmpc_env_step against the traces recorded from the reference's own
Highway_env (tests/golden/highway_env_*.npz) and against the environment oracle on seeded batches."""
import numpy as np
import pytest

from tests.helpers import load_fixture, TOL_U0
from _bmpc import scenarios

pytestmark = pytest.mark.gpu

ENV_FIXTURES = ["highway_env_default", "highway_env_overtake"]


@pytest.fixture(scope="module")
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    from _bmpc import batch, env
    return batch, env


@pytest.mark.parametrize("name", ENV_FIXTURES)
def test_env_replays_reference_trace(mods, name):
    """Every recorded step of the reference environment: obstacle policy, lanes, lane-change target, xRef, both inputs,
    both states, collision flag.  The closed loop feeds on its own outputs, so tolerances are those of the solver."""
    batch, env = mods
    g = load_fixture(name)
    mpc = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=1))
    e = env.BatchedHighwayEnv(mpc, g["x_init"], g["z_init"], int(g["meta_N_lane"]))
    for t in range(int(g["meta_steps"])):
        out = e.step()
        h = e.host()
        assert int(h["obs_policy"][0]) == int(g["backupidx"][t]), t
        assert list(h["lane"][0]) == list(g["lane"][t]), t
        np.testing.assert_allclose(h["xref"][0], g["xref"][t], atol=1e-6, err_msg="xref %d" % t)
        np.testing.assert_allclose(h["policy_params"][0, 2, 1:3], g["lc_target"][t][1:3], atol=1e-9)
        np.testing.assert_allclose(h["u_obs"][0], g["u_obs"][t], atol=1e-9, err_msg="obstacle input %d" % t)
        assert np.abs(out["u0"].cpu().numpy()[0] - g["u_ego"][t]).max() < 1e-5, t
        np.testing.assert_allclose(h["x"][0], g["x"][t], atol=1e-5)
        np.testing.assert_allclose(h["z"][0], g["z"][t], atol=1e-8)
        assert bool(h["collided"][0]) == bool(g["collision"][t])
        assert int(out["status"].cpu().numpy()[0]) <= 1
    mpc.close()


def test_env_batch_matches_oracle_env(mods):
    """Seeded batch, 4 closed-loop steps: a sample of episodes stepped by the oracle environment + oracle controller."""
    from oracle import params
    from oracle.env import HighwayEnvOracle
    batch, env = mods
    B, steps = 512, 4
    x0, z0, _, _ = scenarios.highway_batch(B, seed=4242)
    mpc = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=B))
    e = env.BatchedHighwayEnv(mpc, x0, z0, 4)
    sample = [0, 17, 255, 511]
    oracles = [HighwayEnvOracle(params.highway_branch_mpc(N_lane=4), 4, x0[i], z0[i]) for i in sample]
    for t in range(steps):
        out = e.step()
        h = e.host()
        assert (out["status"].cpu().numpy() <= 1).all()
        for k, i in enumerate(sample):
            u, u_obs = oracles[k].step(t)
            assert int(h["obs_policy"][i]) == oracles[k].backupidx, (t, i)
            assert list(h["lane"][i]) == oracles[k].lane, (t, i)
            np.testing.assert_allclose(h["xref"][i], oracles[k].xref, atol=1e-6)
            assert np.abs(out["u0"].cpu().numpy()[i] - u).max() < TOL_U0, (t, i)
            np.testing.assert_allclose(h["x"][i], oracles[k].x, atol=1e-5)
            np.testing.assert_allclose(h["z"][i], oracles[k].z, atol=1e-9)
    mpc.close()


def test_quadruped_env_matches_oracle_env(mods):
    from oracle import params
    from oracle.env import QuadEnvOracle
    batch, env = mods
    B, steps = 64, 3
    x0, z0, goal = scenarios.quadruped_batch(B, seed=99)
    mpc = batch.BatchedBranchMPC(scenarios.quadruped_config(batch_capacity=B))
    e = env.BatchedQuadEnv(mpc, x0, z0, goal)
    sample = [0, 31, 63]
    oracles = [QuadEnvOracle(params.quadruped_prox_mpc(), goal[i], x0[i], z0[i]) for i in sample]
    for t in range(steps):
        out = e.step()
        h = e.host()
        assert (out["status"].cpu().numpy() <= 1).all()
        for k, i in enumerate(sample):
            u, u_obs = oracles[k].step(t)
            assert int(h["obs_policy"][i]) == oracles[k].backupidx, (t, i)
            np.testing.assert_allclose(h["xref"][i], oracles[k].xref, atol=1e-6)
            np.testing.assert_allclose(h["u_obs"][i], u_obs, atol=1e-12)
            assert np.abs(out["u0"].cpu().numpy()[i] - u).max() < TOL_U0, (t, i)
            np.testing.assert_allclose(h["x"][i], oracles[k].x, atol=1e-4)
            np.testing.assert_allclose(h["z"][i], oracles[k].z, atol=1e-9)
    mpc.close()


def test_default_scenario_full_length_against_oracle(mods):
    """BASELINE config 1: the reference's own workload (main_branch.py sim_overtake: ego [0,1.8,20,0], obstacle [5,5.4,20,0],
    100 closed-loop steps) on the device against the oracle environment + oracle controller, step by step.  The loop feeds on
    its own outputs, so the comparison tolerance grows with the step count; discrete decisions must agree exactly."""
    from oracle import params
    from oracle.env import HighwayEnvOracle
    batch, env = mods
    steps = 100
    mpc = batch.BatchedBranchMPC(scenarios.highway_config(batch_capacity=1))
    e = env.BatchedHighwayEnv(mpc, [0., 1.8, 20., 0.], [5., 5.4, 20., 0.], 4)
    ora = HighwayEnvOracle(params.highway_branch_mpc(N_lane=4), 4)
    worst = 0.0
    for t in range(steps):
        out = e.step()
        u, _ = ora.step(t)
        h = e.host()
        assert int(out["status"].cpu().numpy()[0]) <= 1, t
        assert int(h["obs_policy"][0]) == ora.backupidx and list(h["lane"][0]) == ora.lane, t
        worst = max(worst, float(np.abs(out["u0"].cpu().numpy()[0] - u).max()))
        assert worst < TOL_U0, (t, worst)
        np.testing.assert_allclose(h["x"][0], ora.x, atol=1e-4)
        np.testing.assert_allclose(h["z"][0], ora.z, atol=1e-9)
    assert bool(h["collided"][0]) == bool(ora.collision)
    mpc.close()
