"""Set up an import environment in which the UNMODIFIED reference modules load.

TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py in the build container;
`/root/reference` does not exist on the GPU box).

Two things stand between the Python-3.8-era reference and this image:
  * casadi / osqp / ecos / cvxopt / matplotlib are not installed  -> shims in this directory;
  * `utils.py:72` and `MPC_branch.py:36` declare `field(default=np.array((n, n)))`, which
    Python >= 3.11 rejects (unhashable dataclass default).  `dataclasses.field` is wrapped
    for the duration of the import so such defaults are viewed as a hashable ndarray subclass
    (the class attribute must survive: `PythonMsg.__setattr__` checks `hasattr`).
Neither changes any arithmetic of the reference.
"""
import contextlib
import dataclasses
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("BMPC_REFERENCE", "/root/reference")


def _hashable(arr):
    import numpy as np

    class _HashableArray(np.ndarray):
        __hash__ = object.__hash__

    return np.asarray(arr).view(_HashableArray)


@contextlib.contextmanager
def reference_imports():
    sys.dont_write_bytecode = True          # /root/reference is read-only
    orig_field = dataclasses.field

    def field(*args, **kw):
        d = kw.get("default", dataclasses.MISSING)
        if d is not dataclasses.MISSING and d.__class__.__hash__ is None:
            kw["default"] = _hashable(d)
        return orig_field(*args, **kw)

    dataclasses.field = field
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, HERE)
    try:
        yield
    finally:
        dataclasses.field = orig_field
        sys.path.remove(HERE)
        sys.path.remove(REFERENCE)
