"""Import-only stub (ECOS is used by BranchMPC_CVaR only; outside the golden-fixture scope)."""


def solve(*a, **k):
    raise NotImplementedError("ecos stub")
