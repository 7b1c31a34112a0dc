"""Minimal stand-in for the `casadi` API surface used by the reference's model files.

TEST INFRASTRUCTURE ONLY.  The real CasADi wheel is not installable in this
environment (no network), so the golden-fixture generator
(`tests/golden/make_golden.py`) puts this directory on `sys.path` and imports
the UNMODIFIED reference modules (`/root/reference/highway_branch_dyn.py`,
`quadruped_branch_dyn.py`, `MPC_branch.py`) on top of it.  The reference's own
expression-building code (`calc_xp_expr`, `veh_col`, `softmin`, ...) then runs
as written; this shim only supplies what CasADi supplies: a scalar expression
graph (`SX`), exact derivatives of it (`jacobian`, forward-mode AD here) and
numeric evaluation (`Function`).

Only the calls that appear in the reference are implemented:
SX(r[,c]) / SX.sym / indexing / arithmetic / .T / .shape, cos sin exp fabs,
vertcat, sum1, norm_1, jacobian, Function, DM-like results supporting
`A@x`, `xp-A@x-B@u`, `np.array(.)`.
"""
import sys as _sys
import math as _math
import numpy
import numpy as _np

casadi = _sys.modules[__name__]

# ----------------------------------------------------------------------------
# scalar expression nodes
# ----------------------------------------------------------------------------
_OPS = ("const", "sym", "add", "sub", "mul", "div", "neg", "cos", "sin", "exp", "fabs", "lut")


class _Node:
    __slots__ = ("op", "a", "b", "val", "idx")
    _counter = 0

    def __init__(self, op, a=None, b=None, val=0.0):
        self.op = op
        self.a = a
        self.b = b
        self.val = val
        _Node._counter += 1
        self.idx = _Node._counter          # creation order == a topological order


def _n(v):
    if isinstance(v, _Node):
        return v
    if isinstance(v, SX):
        if v.numel() != 1:
            raise ValueError("expected a scalar SX")
        return v._e.flat[0]
    return _Node("const", val=float(v))


def _is_zero(nd):
    return nd.op == "const" and nd.val == 0.0


def _bin(op, a, b):
    a = _n(a)
    b = _n(b)
    if a.op == "const" and b.op == "const":
        x, y = a.val, b.val
        return _Node("const", val={"add": x + y, "sub": x - y, "mul": x * y,
                                   "div": (x / y) if op == "div" else 0.0}[op])
    # structural zeros keep the graph small; numerically exact (x+0, 0*x)
    if op == "add":
        if _is_zero(a):
            return b
        if _is_zero(b):
            return a
    if op == "sub" and _is_zero(b):
        return a
    if op == "mul" and (_is_zero(a) or _is_zero(b)):
        return _Node("const", val=0.0)
    return _Node(op, a, b)


def _un(op, a):
    a = _n(a)
    if a.op == "const":
        f = {"neg": lambda v: -v, "cos": _math.cos, "sin": _math.sin,
             "exp": _math.exp, "fabs": abs}[op]
        return _Node("const", val=f(a.val))
    return _Node(op, a)


# ----------------------------------------------------------------------------
# SX: dense matrix of scalar nodes
# ----------------------------------------------------------------------------
class SX:
    __array_priority__ = 1000

    def __init__(self, *args):
        if len(args) == 0:
            self._e = _np.empty((0, 0), dtype=object)
        elif len(args) == 1 and isinstance(args[0], SX):
            self._e = args[0]._e.copy()
        elif len(args) == 1 and isinstance(args[0], (int, _np.integer)) and not isinstance(args[0], bool):
            self._e = self._zeros(int(args[0]), 1)
        elif len(args) == 2 and all(isinstance(a, (int, _np.integer)) for a in args):
            self._e = self._zeros(int(args[0]), int(args[1]))
        elif len(args) == 1:
            arr = _np.atleast_1d(_np.asarray(args[0], dtype=float))
            if arr.ndim == 1:
                arr = arr.reshape(-1, 1)
            self._e = _np.empty(arr.shape, dtype=object)
            for i in range(arr.shape[0]):
                for j in range(arr.shape[1]):
                    self._e[i, j] = _Node("const", val=float(arr[i, j]))
        else:
            raise TypeError("unsupported SX constructor")

    @staticmethod
    def _zeros(r, c):
        e = _np.empty((r, c), dtype=object)
        for i in range(r):
            for j in range(c):
                e[i, j] = _Node("const", val=0.0)
        return e

    @staticmethod
    def _wrap(e):
        s = SX.__new__(SX)
        s._e = e
        return s

    @staticmethod
    def sym(name, r=1, c=1):
        e = _np.empty((r, c), dtype=object)
        for j in range(c):
            for i in range(r):
                e[i, j] = _Node("sym")
        return SX._wrap(e)

    @staticmethod
    def ones(r, c=1):
        return SX(_np.ones((r, c)))

    # shape ------------------------------------------------------------------
    @property
    def shape(self):
        return self._e.shape

    def numel(self):
        return self._e.size

    @property
    def T(self):
        return SX._wrap(self._e.T.copy())

    # indexing ---------------------------------------------------------------
    def _key(self, key):
        if not isinstance(key, tuple):
            # single index on a vector (column or row): linear (column-major) index
            if self._e.shape[1] == 1:
                key = (key, 0)
            elif self._e.shape[0] == 1:
                key = (0, key)
            else:
                raise IndexError("linear indexing only supported on vectors")
        return key

    def __getitem__(self, key):
        key = self._key(key)
        r = self._e[key]
        if isinstance(r, _Node):
            e = _np.empty((1, 1), dtype=object)
            e[0, 0] = r
            return SX._wrap(e)
        if r.ndim == 1:
            # keep orientation: row slice -> 1 x k, column slice -> k x 1
            if isinstance(key[0], (int, _np.integer)):
                r = r.reshape(1, -1)
            else:
                r = r.reshape(-1, 1)
        return SX._wrap(r.copy())

    def __setitem__(self, key, value):
        key = self._key(key)
        target = self._e[key]
        if isinstance(target, _Node):
            self._e[key] = _n(value)
            return
        if isinstance(value, SX):
            src = value._e
            if src.size != target.size:
                raise ValueError("size mismatch in SX assignment")
            # CasADi accepts a column into a row slice (same numel): column-major flatten
            self._e[key] = src.reshape(-1, order="F").reshape(target.shape, order="F")
        else:
            arr = _np.asarray(value, dtype=float)
            if arr.ndim == 0:
                flat = [float(arr)] * target.size
            else:
                flat = [float(v) for v in arr.reshape(-1, order="F")]
            out = _np.empty(target.size, dtype=object)
            for k, v in enumerate(flat):
                out[k] = _Node("const", val=v)
            self._e[key] = out.reshape(target.shape, order="F")

    # arithmetic (elementwise, with scalar broadcasting) ----------------------
    def _ew(self, other, op, swap=False):
        if isinstance(other, SX):
            o = other._e
        else:
            arr = _np.asarray(other, dtype=float)
            if arr.ndim == 0:
                o = None
                sc = float(arr)
            else:
                if arr.ndim == 1:
                    arr = arr.reshape(-1, 1)
                o = SX(arr)._e
        a = self._e
        if isinstance(other, SX) or o is not None:
            if a.shape != o.shape:
                if a.size == 1:
                    a = _np.broadcast_to(a, o.shape)
                elif o.size == 1:
                    o = _np.broadcast_to(o, a.shape)
                else:
                    raise ValueError("shape mismatch %s vs %s" % (a.shape, o.shape))
            out = _np.empty(a.shape, dtype=object)
            for idx in _np.ndindex(a.shape):
                x, y = (o[idx], a[idx]) if swap else (a[idx], o[idx])
                out[idx] = _bin(op, x, y)
            return SX._wrap(out)
        out = _np.empty(a.shape, dtype=object)
        for idx in _np.ndindex(a.shape):
            x, y = (sc, a[idx]) if swap else (a[idx], sc)
            out[idx] = _bin(op, x, y)
        return SX._wrap(out)

    def __add__(self, o): return self._ew(o, "add")
    def __radd__(self, o): return self._ew(o, "add", True)
    def __sub__(self, o): return self._ew(o, "sub")
    def __rsub__(self, o): return self._ew(o, "sub", True)
    def __mul__(self, o): return self._ew(o, "mul")
    def __rmul__(self, o): return self._ew(o, "mul", True)
    def __truediv__(self, o): return self._ew(o, "div")
    def __rtruediv__(self, o): return self._ew(o, "div", True)

    def __neg__(self):
        return self._map("neg")

    def _map(self, op):
        out = _np.empty(self._e.shape, dtype=object)
        for idx in _np.ndindex(self._e.shape):
            out[idx] = _un(op, self._e[idx])
        return SX._wrap(out)

    def __matmul__(self, other):
        o = other._e if isinstance(other, SX) else SX(other)._e
        a = self._e
        out = _np.empty((a.shape[0], o.shape[1]), dtype=object)
        for i in range(a.shape[0]):
            for j in range(o.shape[1]):
                acc = _Node("const", val=0.0)
                for k in range(a.shape[1]):
                    acc = _bin("add", acc, _bin("mul", a[i, k], o[k, j]))
                out[i, j] = acc
        return SX._wrap(out)

    # numpy ufuncs applied to an SX (np.exp(alpha*dx) in veh_col / softsat)
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        name = ufunc.__name__
        if name in ("exp", "cos", "sin"):
            return inputs[0]._map(name)
        if name in ("absolute", "fabs"):
            return inputs[0]._map("fabs")
        table = {"add": "add", "subtract": "sub", "multiply": "mul",
                 "divide": "div", "true_divide": "div"}
        if name in table:
            a, b = inputs
            if isinstance(a, SX):
                return a._ew(b, table[name])
            return b._ew(a, table[name], True)
        if name == "negative":
            return inputs[0]._map("neg")
        return NotImplemented

    def exp(self): return self._map("exp")
    def cos(self): return self._map("cos")
    def sin(self): return self._map("sin")


MX = SX   # the SX-only reference paths never build MX graphs; isinstance checks still work


def _lift(x):
    return x if isinstance(x, SX) else SX(_np.atleast_1d(_np.asarray(x, dtype=float)))


def cos(x):
    return x._map("cos") if isinstance(x, SX) else _np.cos(x)


def sin(x):
    return x._map("sin") if isinstance(x, SX) else _np.sin(x)


def exp(x):
    return x._map("exp") if isinstance(x, SX) else _np.exp(x)


def fabs(x):
    return x._map("fabs") if isinstance(x, SX) else _np.fabs(x)


def vertcat(*args):
    if not any(isinstance(a, SX) for a in args):
        # numeric use in the reference: softmax(vertcat(-5,-x[2]),3) on floats
        return _np.array([float(a) for a in args])
    cols = [_lift(a)._e for a in args]
    return SX._wrap(_np.vstack(cols))


def sum1(x):
    if not isinstance(x, SX):
        return _np.sum(x, axis=0)
    out = _np.empty((1, x._e.shape[1]), dtype=object)
    for j in range(x._e.shape[1]):
        acc = _Node("const", val=0.0)
        for i in range(x._e.shape[0]):
            acc = _bin("add", acc, x._e[i, j])
        out[0, j] = acc
    return SX._wrap(out)


def norm_1(x):
    acc = _Node("const", val=0.0)
    for idx in _np.ndindex(x._e.shape):
        acc = _bin("add", acc, _un("fabs", x._e[idx]))
    e = _np.empty((1, 1), dtype=object)
    e[0, 0] = acc
    return SX._wrap(e)


def reshape(x, *shape):
    """casadi.reshape is column-major (numeric: HMM_backup_dyn.py:213; symbolic: :244, :258 of the belief-state model)."""
    if len(shape) == 1:
        shape = tuple(shape[0])
    if isinstance(x, SX):
        return SX._wrap(x._e.reshape(-1, order="F").reshape(shape, order="F").copy())
    return _np.reshape(_np.asarray(x, dtype=float), shape, order="F")


def sumsqr(x):
    e = _lift(x)._e.reshape(-1)
    acc = _bin("mul", e[0], e[0])
    for nd in e[1:]:
        acc = _bin("add", acc, _bin("mul", nd, nd))
    out = _np.empty((1, 1), dtype=object)
    out[0, 0] = acc
    return SX._wrap(out)


def kron(a, b):
    a = _lift(a)._e
    b = _lift(b)._e
    out = _np.empty((a.shape[0] * b.shape[0], a.shape[1] * b.shape[1]), dtype=object)
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            for k in range(b.shape[0]):
                for l in range(b.shape[1]):
                    out[i * b.shape[0] + k, j * b.shape[1] + l] = _bin("mul", a[i, j], b[k, l])
    return SX._wrap(out)


class interpolant:
    """1-D piecewise-linear lookup table `interpolant(name, 'linear', [grid], values)` (main_branch.py:78-79): callable on
    numbers (returns a DM) and on symbolic scalars (a graph node whose derivative is the slope of the active segment).
    Outside the grid the end segments are continued, as CasADi's 'linear' plugin does."""

    def __init__(self, name, kind, grid, values, *args, **kwargs):
        if kind != "linear" or len(grid) != 1:
            raise NotImplementedError("only 1-D linear lookup tables")
        self.name = name
        self.xs = _np.asarray(grid[0], dtype=float).reshape(-1)
        self.ys = _np.asarray(values, dtype=float).reshape(-1)
        if self.xs.size != self.ys.size or self.xs.size < 2 or _np.any(_np.diff(self.xs) <= 0):
            raise ValueError("lookup table needs a strictly increasing grid")

    def segment(self, x):
        k = int(_np.searchsorted(self.xs, x, side="right")) - 1
        k = min(max(k, 0), self.xs.size - 2)
        slope = (self.ys[k + 1] - self.ys[k]) / (self.xs[k + 1] - self.xs[k])
        return self.ys[k] + slope * (x - self.xs[k]), slope

    def __call__(self, x):
        if isinstance(x, SX):
            nd = _n(x)
            if nd.op == "const":
                nd = _Node("const", val=self.segment(nd.val)[0])
            else:
                nd = _Node("lut", nd, val=self)
            e = _np.empty((1, 1), dtype=object)
            e[0, 0] = nd
            return SX._wrap(e)
        return _LookupValue(_np.array([[self.segment(float(_np.asarray(x).reshape(-1)[0]))[0]]]))


# ----------------------------------------------------------------------------
# jacobian: a deferred object evaluated by forward-mode AD inside Function
# ----------------------------------------------------------------------------
class _Jacobian:
    def __init__(self, expr, wrt):
        self.expr = _lift(expr)
        self.wrt = wrt
        # CasADi returns numel(expr) x numel(wrt)
        self.shape = (self.expr.numel(), wrt.numel())


def jacobian(expr, wrt):
    return _Jacobian(expr, wrt)


class DM:
    """Numeric result with CasADi's matrix semantics (1-D numpy operands are columns)."""
    __array_priority__ = 2000

    def __init__(self, a):
        a = _np.asarray(a, dtype=float)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        self._a = a

    @property
    def shape(self):
        return self._a.shape

    def __array__(self, dtype=None, copy=None):
        return self._a.astype(dtype) if dtype is not None else self._a.copy()

    @staticmethod
    def _col(o):
        if isinstance(o, DM):
            return o._a
        o = _np.asarray(o, dtype=float)
        if o.ndim == 1:
            o = o.reshape(-1, 1)
        return o

    def __matmul__(self, o): return type(self)(self._a @ self._col(o))
    def __rmatmul__(self, o): return type(self)(self._col(o) @ self._a)
    def __add__(self, o): return type(self)(self._a + self._col(o))
    def __radd__(self, o): return type(self)(self._col(o) + self._a)
    def __sub__(self, o): return type(self)(self._a - self._col(o))
    def __rsub__(self, o): return type(self)(self._col(o) - self._a)
    def __mul__(self, o): return type(self)(self._a * self._col(o))
    def __rmul__(self, o): return type(self)(self._col(o) * self._a)
    def __neg__(self): return type(self)(-self._a)
    def __float__(self): return float(self._a.reshape(-1)[0])

    def full(self):
        return self._a.copy()


class _LookupValue(DM):
    """Value of a lookup table at a number: a 1x1 DM that numpy sees as a scalar, so that the reference's
    `np.array([a, psiref(x[0]) - K * x[3]])` (highway_branch_dyn.py:96) builds a length-2 vector."""

    def __array__(self, dtype=None, copy=None):
        return _np.asarray(self._a.reshape(-1)[0], dtype=dtype or float)


class Function:
    def __init__(self, name, inputs, outputs):
        self.name = name
        self._in = inputs
        self._out = outputs
        self._sym_index = {}
        k = 0
        for inp in inputs:
            for nd in inp._e.reshape(-1, order="F"):
                self._sym_index[id(nd)] = k
                k += 1
        self._nin = k
        # collect reachable nodes once, in creation (= topological) order
        roots = []
        for o in outputs:
            e = o.expr._e if isinstance(o, _Jacobian) else _lift(o)._e
            roots.extend(e.reshape(-1))
        seen = {}
        stack = list(roots)
        while stack:
            nd = stack.pop()
            if id(nd) in seen:
                continue
            seen[id(nd)] = nd
            if nd.a is not None:
                stack.append(nd.a)
            if nd.b is not None:
                stack.append(nd.b)
        self._order = sorted(seen.values(), key=lambda nd: nd.idx)
        self._need_grad = any(isinstance(o, _Jacobian) for o in outputs)

    def __call__(self, *args):
        vals_in = []
        for a, inp in zip(args, self._in):
            arr = _np.asarray(a.full() if isinstance(a, DM) else a, dtype=float)
            vals_in.extend(arr.reshape(-1, order="F").tolist())
        if len(vals_in) != self._nin:
            raise ValueError("Function %s: wrong number of input elements" % self.name)
        val = {}
        grad = {} if self._need_grad else None
        nin = self._nin
        zero = _np.zeros(nin)
        for nd in self._order:
            op = nd.op
            key = id(nd)
            if op == "const":
                v = nd.val
                g = zero
            elif op == "sym":
                k = self._sym_index.get(key)
                if k is None:
                    raise ValueError("free symbol in Function %s" % self.name)
                v = vals_in[k]
                if grad is not None:
                    g = _np.zeros(nin)
                    g[k] = 1.0
            else:
                va = val[id(nd.a)]
                ga = grad[id(nd.a)] if grad is not None else None
                if nd.b is not None:
                    vb = val[id(nd.b)]
                    gb = grad[id(nd.b)] if grad is not None else None
                if op == "add":
                    v = va + vb
                    g = ga + gb if grad is not None else None
                elif op == "sub":
                    v = va - vb
                    g = ga - gb if grad is not None else None
                elif op == "mul":
                    v = va * vb
                    g = ga * vb + va * gb if grad is not None else None
                elif op == "div":
                    v = va / vb
                    g = (ga - v * gb) / vb if grad is not None else None
                elif op == "neg":
                    v = -va
                    g = -ga if grad is not None else None
                elif op == "cos":
                    v = _math.cos(va)
                    g = -_math.sin(va) * ga if grad is not None else None
                elif op == "sin":
                    v = _math.sin(va)
                    g = _math.cos(va) * ga if grad is not None else None
                elif op == "exp":
                    v = _math.exp(va)
                    g = v * ga if grad is not None else None
                elif op == "lut":
                    v, slope = nd.val.segment(va)
                    g = slope * ga if grad is not None else None
                elif op == "fabs":
                    v = abs(va)
                    sgn = 1.0 if va > 0 else (-1.0 if va < 0 else 0.0)
                    g = sgn * ga if grad is not None else None
                else:
                    raise ValueError(op)
            val[key] = v
            if grad is not None:
                grad[key] = g
        results = []
        for o in self._out:
            if isinstance(o, _Jacobian):
                cols = [self._sym_index[id(nd)] for nd in o.wrt._e.reshape(-1, order="F")]
                rows = o.expr._e.reshape(-1, order="F")
                J = _np.zeros((len(rows), len(cols)))
                for i, nd in enumerate(rows):
                    J[i, :] = grad[id(nd)][cols]
                results.append(DM(J))
            else:
                e = _lift(o)._e
                out = _np.zeros(e.shape)
                for idx in _np.ndindex(e.shape):
                    out[idx] = val[id(e[idx])]
                results.append(DM(out))
        return results[0] if len(results) == 1 else tuple(results)
