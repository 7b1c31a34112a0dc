"""Backup policies as data.

The reference passes policies around as Python closures over CasADi-or-numpy functions
(`main_branch.py:39`: `lambda x: backup_lc(x, xRef)`).  A kernel cannot call a closure, so the drop-in policy functions
recognise a probe object: called with a `PolicyProbe` they return a `PolicyDescriptor` (kind + parameters) instead of a
control; called with numbers they evaluate the reference's numeric branch.  `describe(callables)` turns a
`backupcons` list into the policy table of bmpc_config.
"""
import numpy as np

from . import abi


class PolicyProbe:
    """Stand-in state handed to a policy closure to find out which policy it is."""

    def __getitem__(self, k):
        raise TypeError("policy closures must call one of the library's backup_* functions directly on the state")


class PolicyDescriptor:
    def __init__(self, kind, params=(), consts=None, table=None):
        self.kind = int(kind)
        self.params = [float(v) for v in params] + [0.0] * (4 - len(params))
        self.consts = dict(consts or {})
        self.table = table        # lookup table psiref(x) of the merge scenario's ramp policies, or None

    def entry(self):
        return self.kind, self.params


def describe(backupcons):
    out = []
    for k, fn in enumerate(backupcons):
        d = fn(PolicyProbe())
        if not isinstance(d, PolicyDescriptor):
            raise TypeError("backup policy %d is not built from the library's backup_* functions" % k)
        out.append(d)
    return out


def table(descriptors):
    return [d.entry() for d in descriptors]


def param_array(descriptors):
    return np.array([d.params for d in descriptors], dtype=np.float64)
