"""Import-only stub: the reference does `from cvxopt import spmatrix, matrix, solvers` but never calls them."""
from . import solvers  # noqa: F401


def spmatrix(*a, **k):
    raise NotImplementedError("cvxopt stub")


def matrix(*a, **k):
    raise NotImplementedError("cvxopt stub")
