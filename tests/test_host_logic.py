"""Host-side translation of the reference's parameter objects into bmpc_config, and the synthetic scenario generator."""
import numpy as np
import pytest

from tests.helpers import params
from _bmpc import abi, config, scenarios


def test_state_rows_are_paired_like_the_reference_constraints():
    par = params.highway_mpc_params()
    rows = config.pair_state_rows(par["Fx"], par["bx"], 4)
    assert len(rows) == 2
    (f0, lo0, hi0), (f1, lo1, hi1) = rows
    assert list(f0) == [0, 1, 0, 0] and lo0 == pytest.approx(1.25) and hi0 == pytest.approx(4 * 3.6 - 1.25)
    assert list(f1) == [0, 0, 0, 1] and (lo1, hi1) == (-0.25, 0.25)
    # one-sided row stays one-sided
    rows = config.pair_state_rows(np.array([[1., 0, 0, 0]]), np.array([3.0]), 4)
    assert rows[0][1] == -np.inf and rows[0][2] == 3.0
    # quadruped: no state rows (Init_MPC.py:75)
    assert config.pair_state_rows(np.empty((0, 3)), np.empty(0), 3) == []


def test_input_box_from_Fu():
    par = params.highway_mpc_params()
    lo, hi = config.input_box(par["Fu"], par["bu"], 2)
    assert list(lo) == [-6.0, -0.3] and list(hi) == [6.0, 0.3]
    q = params.quadruped_mpc_params()
    lo, hi = config.input_box(q["Fu"], q["bu"], 3)
    assert list(lo) == [0.0, -0.1, -0.5] and list(hi) == [0.2, 0.1, 0.5]     # vx in [0, vxm] (Init_MPC.py:79-84)
    with pytest.raises(ValueError):
        config.input_box(np.array([[1., 1.]]), np.array([1.0]), 2)
    with pytest.raises(ValueError):
        config.input_box(np.array([[1., 0.]]), np.array([1.0]), 2)


def test_highway_config_fields():
    c = scenarios.highway_config(batch_capacity=7)
    assert (c.model, c.controller, c.n, c.d, c.N, c.NB, c.m) == (abi.MODEL_HIGHWAY, abi.CTRL_BRANCH, 4, 2, 8, 2, 3)
    assert list(c.policy_kind)[:3] == [abi.POLICY_MAINTAIN, abi.POLICY_BRAKE, abi.POLICY_LC]
    assert list(c.policy_param[2]) == [0.5, 1.8, 15.0, 0.0]
    assert np.allclose(np.array(c.Q).reshape(4, 4), np.diag([0, 3, 3, 10]))
    assert list(c.Qslack) == [0.0, 300.0]
    assert c.lane_lo == 1.25 and c.lane_hi == pytest.approx(3 * 3.6 - 1.25)   # model lane boundary uses N_lane=3
    assert c.batch_capacity == 7
    with pytest.raises(TypeError):
        scenarios.highway_config(no_such_knob=1)


def test_scenario_batch_is_seeded_and_in_range():
    a = scenarios.highway_batch(256, seed=5)
    b = scenarios.highway_batch(256, seed=5)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)
    x0, z0, xref, pp = a
    assert x0.shape == (256, 4) and pp.shape == (256, 3, 4)
    assert (x0[:, 2] >= 15).all() and (x0[:, 2] <= 25).all() and (np.abs(x0[:, 3]) <= 0.2).all()
    assert ((z0[:, 0] >= -15) & (z0[:, 0] <= 25)).all()
    lanes = (pp[:, 2, 1] - 1.8) / 3.6
    assert np.allclose(lanes, np.round(lanes)) and lanes.min() >= 0 and lanes.max() <= 3


def test_reciprocal_integer_division_is_exact():
    """bmpc_idiv (csrc/bmpc_params.h): floor(q / d) as int(float32(q) * nextafter(float32(1/d), 2)) for every q < 65536 and
    every divisor the tree numbering can meet (N, m <= 64)."""
    q = np.arange(65536)
    for d in range(1, 65):
        inv = np.nextafter(np.float32(1.0) / np.float32(d), np.float32(2.0))
        got = (q.astype(np.float32) * inv).astype(np.int64)
        assert np.array_equal(got, q // d), d
