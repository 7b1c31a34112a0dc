"""Drop-in for the reference's `MPC_nobranch.py` (BASELINE configs[1] names it): single-trajectory MPC against every
obstacle node of the scenario tree.  The reference file is a draft copy of `MPC_branch.robustMPC` that cannot run
(MPC_nobranch.py:140-222: undefined names, arity mismatches); its working equivalent is
`MPC_branch.robustMPC` (:1275-1595), which is what this module exports under the reference's names."""
from MPC_branch import BranchMPCParams, BranchTree, robustMPC  # noqa: F401

__all__ = ["robustMPC", "BranchMPCParams", "BranchTree"]
