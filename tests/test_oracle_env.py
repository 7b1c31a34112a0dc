"""The closed-loop environment restatement (oracle/env.py) against traces recorded from the reference's own
Highway_env_branch.Highway_env around the reference BranchMPC (tests/golden/highway_env_*.npz)."""
import numpy as np
import pytest

from tests.helpers import load_fixture
from oracle import params
from oracle.env import HighwayEnvOracle

ENV_FIXTURES = ["highway_env_default", "highway_env_overtake"]


@pytest.mark.parametrize("name", ENV_FIXTURES)
def test_env_oracle_replays_reference_trace(name):
    g = load_fixture(name)
    steps = min(int(g["meta_steps"]), 12)          # the exact QP solve dominates (0.3 s per step)
    mpc = params.highway_branch_mpc(N_lane=int(g["meta_N_lane"]))
    env = HighwayEnvOracle(mpc, int(g["meta_N_lane"]), g["x_init"], g["z_init"])
    for t in range(steps):
        u, u_obs = env.step(t)
        assert env.backupidx == int(g["backupidx"][t]), t
        assert env.lane == list(g["lane"][t]), t
        np.testing.assert_allclose(env.xref, g["xref"][t], atol=1e-7, err_msg="xref %d" % t)
        np.testing.assert_allclose(env.lc_target()[1:3], g["lc_target"][t][1:3], atol=1e-9, err_msg="lc target %d" % t)
        np.testing.assert_allclose(u_obs, g["u_obs"][t], atol=1e-9, err_msg="obstacle input %d" % t)
        np.testing.assert_allclose(u, g["u_ego"][t], atol=1e-6, err_msg="ego input %d" % t)
        np.testing.assert_allclose(env.x, g["x"][t], atol=1e-6)
        np.testing.assert_allclose(env.z, g["z"][t], atol=1e-9)
        assert bool(env.collision) == bool(g["collision"][t])
