"""Drop-in for the reference's `utils.py` (parameter containers; /root/reference/utils.py:14-90).

Same class names, field names, defaults and the "no new fields" guard, so that `main_branch.py` /
`main_quadruped.py` construct them unchanged.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class PythonMsg:
    """Attribute guard of the reference (utils.py:14-20): assigning an unknown field is an error."""

    def __setattr__(self, key, value):
        if not hasattr(self, key):
            raise TypeError('Cannot add new field "%s" to frozen class %s' % (key, self))
        object.__setattr__(self, key, value)


def _f():
    return field(default=None)


@dataclass
class Branch_constants:
    """Highway scenario constants (utils.py:25-42)."""
    s1: float = _f()
    s2: float = _f()
    c2: float = _f()
    tran_diag: float = _f()
    alpha: float = _f()
    R: float = _f()
    am: float = _f()
    rm: float = _f()
    J_c: float = _f()
    s_c: float = _f()
    ylb: float = _f()
    yub: float = _f()
    W: float = _f()
    L: float = _f()
    col_alpha: float = _f()
    Kpsi: float = _f()


@dataclass
class Quad_constants:
    """Quadruped scenario constants (utils.py:44-59)."""
    s1: float = _f()
    s2: float = _f()
    c2: float = _f()
    alpha: float = _f()
    R: float = _f()
    vxm: float = _f()
    vym: float = _f()
    rm: float = _f()
    W1: float = _f()
    L1: float = _f()
    W2: float = _f()
    L2: float = _f()
    col_tol: float = _f()
    col_alpha: float = _f()


@dataclass
class HMM_constants:
    """Constants of the belief-state model.  The reference's HMM_backup_dyn.py imports this name from utils (:5) but its
    utils.py does not define it (the module does not import as shipped); the fields are the ones the module reads."""
    s1: float = _f()
    s2: float = _f()
    c2: float = _f()
    tran_diag: float = _f()
    alpha: float = _f()
    R: float = _f()
    am: float = _f()
    rm: float = _f()
    J_c: float = _f()
    s_c: float = _f()
    ylb: float = _f()
    yub: float = _f()
    W: float = _f()
    L: float = _f()
    col_alpha: float = _f()
    Kpsi: float = _f()


@dataclass
class MPCParams(PythonMsg):
    """Belief-state MPC parameters (utils.py:61-90)."""
    n: int = _f()
    d: int = _f()
    N: int = _f()
    M: int = _f()
    m: int = _f()
    A: np.ndarray = _f()
    B: np.ndarray = _f()
    Q: np.ndarray = _f()
    R: np.ndarray = _f()
    Qf: np.ndarray = _f()
    dR: np.ndarray = _f()
    Qslack: float = _f()
    Fx: np.ndarray = _f()
    bx: np.ndarray = _f()
    Fu: np.ndarray = _f()
    bu: np.ndarray = _f()
    xRef: np.ndarray = _f()
    slacks: bool = field(default=True)
    timeVarying: bool = field(default=False)

    def __post_init__(self):
        if self.Qf is None:
            self.Qf = np.zeros((self.n, self.n))
        if self.dR is None:
            self.dR = np.zeros(self.d)
        if self.xRef is None:
            self.xRef = np.zeros(self.n)
