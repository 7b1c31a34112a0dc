import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
import numpy as np
from _bmpc import batch, scenarios
from oracle.belief_mpc import BeliefModelOracle
np.set_printoptions(precision=5, suppress=True, linewidth=220)
mpc = batch.BatchedBranchMPC(scenarios.belief_config(N=10, M=2, m=2))
xb = np.array([0, 1.8, 20, 0.02, .6, .4, .3, .7]); xbackup = np.array([[10, 1.8, 18, 0], [9, 1.8, 15, 0], [-5, 5.4, 22, 0], [-6, 5.4, 20, 0.]])
u = np.array([0.5, 0.01])
r = mpc.eval_belief(xb, xbackup[None], u)
A, B, C, h0, Jh, xbp = BeliefModelOracle(2, 2, 0.1).linearize(xb, xbackup, u)
print("device A\n", r["A"][0]); print("oracle A\n", A)
print("dC", np.abs(r["C"][0] - C).max(), "dh0", np.abs(r["h0"][0] - h0).max(), "dJh", np.abs(r["Jh"][0] - Jh).max(), "dxbp", np.abs(r["xbp"][0] - xbp).max())
