"""Constraint plugin surface: proximal operators ``f(X, step) -> X'`` attached to parameters.

Mirrors scarlet/constraint.py (``Constraint`` 10-55, ``ConstraintChain`` 58-80, ``PositivityConstraint`` 83-92,
``NormalizationConstraint`` 95-114, ``MonotonicityConstraint`` 183-234, ``SymmetryConstraint`` 262-273,
``CenterOnConstraint`` 276-287).  Every built-in constraint knows its device op-code (``to_ops``); calling one
directly runs the same CUDA chain kernel the fitting loop uses (``sb_prox_chain_*``).  A constraint without a
device op-code cannot be fitted: the plan builder raises ``TypeError`` naming the class -- nothing is ever
evaluated on the host behind the user's back.
"""
import ctypes

import numpy as np

from . import _native as nat
from . import operator
from .cache import Cache


class MonoTables:
    """Collects the distinct monotonic-operator tables a set of chains refers to."""

    def __init__(self):
        self.keys = {}
        self.tables = []

    def index(self, shape, center, neighbor_weight):
        key = (tuple(shape), tuple(center), neighbor_weight)
        if key not in self.keys:
            name = "operator.monotonic_tables"
            try:
                t = Cache.check(name, key)
            except KeyError:
                t = operator.monotonic_tables(shape, neighbor_weight, center)
                Cache.set(name, key, t)
            self.keys[key] = len(self.tables)
            self.tables.append(t)
        return self.keys[key]

    def descs(self):
        arr = (nat.sb_mono_desc * max(len(self.tables), 1))()
        for i, (w, off, idx) in enumerate(self.tables):
            arr[i] = nat.sb_mono_desc(int(w.shape[1]), int(off.size), int(idx.size), 0, nat.ptr(w), nat.ptr(off), nat.ptr(idx))
        return arr


def chain_desc(ops, repeat=1):
    if len(ops) > nat.SB_MAX_CHAIN_OPS:
        raise ValueError("constraint chain longer than %d operators" % nat.SB_MAX_CHAIN_OPS)
    d = nat.sb_chain_desc()
    d.n_ops, d.repeat = len(ops), int(repeat)
    for i, (code, iarg, farg) in enumerate(ops):
        d.ops[i] = nat.sb_op(int(code), int(iarg), float(farg))
    return d


def constraint_ops(constraint, shape, tables):
    """Flatten a Constraint / ConstraintChain into ``([(code, iarg, farg), ...], repeat)`` for ``shape``."""
    if isinstance(constraint, ConstraintChain):
        ops = []
        for c in constraint.constraints:
            sub, rep = constraint_ops(c, shape, tables)
            if rep != 1:
                raise TypeError("nested ConstraintChain with repeat != 1 has no device op-code")
            ops += sub
        return ops, constraint.repeat
    if isinstance(constraint, Constraint) and type(constraint).to_ops is not Constraint.to_ops:
        return constraint.to_ops(shape, tables), 1
    raise TypeError("constraint %s has no device op-code (supported: Positivity, Normalization, Monotonicity, "
                    "Symmetry, CenterOn and chains thereof); scarlet_b200 has no host fallback"
                    % type(constraint).__name__)


def run_on_device(constraint, X):
    """Apply ``constraint`` to one image (2-D) or spectrum (1-D) with the CUDA chain kernel; returns a new array."""
    X = np.asarray(X)
    dtype = np.float32 if X.dtype == np.float32 else np.float64
    shape = X.shape if X.ndim == 2 else (1, X.size)
    tables = MonoTables()
    ops, repeat = constraint_ops(constraint, shape, tables)
    out = np.array(X, dtype=dtype, order="C", copy=True).reshape(shape)
    fn = nat.lib().sb_prox_chain_f32 if dtype == np.float32 else nat.lib().sb_prox_chain_f64
    desc = chain_desc(ops, repeat)
    monos = tables.descs()
    nat.check(fn(nat.ptr(out), int(shape[0]), int(shape[1]), 1, ctypes.byref(desc), ctypes.addressof(monos),
                 len(tables.tables), nat.default_device()))
    return out.reshape(X.shape)


class Constraint:
    """Base class.  ``Constraint(f)`` wraps an arbitrary host callable -- usable directly, but not fittable."""

    def __init__(self, f=None):
        self.f = f

    def __call__(self, X, step):
        if self.f is not None:
            return self.f(X, step)
        return X

    def to_ops(self, shape, tables):
        raise TypeError("%s has no device op-code" % type(self).__name__)


class ConstraintChain:
    def __init__(self, *constraints, repeat=1):
        assert isinstance(repeat, int) and repeat >= 1
        self.constraints = constraints
        self.repeat = repeat

    def __call__(self, X, step):
        out = run_on_device(self, X)
        if isinstance(X, np.ndarray) and X.dtype == out.dtype:
            X[...] = out  # the reference's chains end in in-place operators
            return X
        return out


class PositivityConstraint(Constraint):
    def __init__(self, zero=0):
        self.zero = zero

    def to_ops(self, shape, tables):
        return [(nat.OP_POSITIVITY, 0, self.zero)]

    def __call__(self, X, step):
        return run_on_device(self, X)


class NormalizationConstraint(Constraint):
    def __init__(self, type="sum"):
        type = type.lower()
        assert type in ["sum", "max"]
        self.type = type

    def to_ops(self, shape, tables):
        return [(nat.OP_NORMALIZE, 1 if self.type == "max" else 0, 0.0)]

    def __call__(self, X, step):
        X[...] = run_on_device(self, X)
        return X


class MonotonicityConstraint(Constraint):
    def __init__(self, neighbor_weight="flat", min_gradient=0.1, use_mask=False, fit_center_radius=0):
        if use_mask or fit_center_radius > 0:
            raise NotImplementedError("use_mask / fit_center_radius are outside the device path (SURVEY 8b)")
        self.neighbor_weight = neighbor_weight
        self.min_gradient = min_gradient
        self.use_mask = False
        self.fit_center = False
        self.fit_center_radius = 0

    def to_ops(self, shape, tables):
        center = (shape[0] // 2, shape[1] // 2)
        return [(nat.OP_MONOTONIC, tables.index(shape, center, self.neighbor_weight), self.min_gradient)]

    def __call__(self, morph, step):
        morph[...] = run_on_device(self, morph)
        return morph


class SymmetryConstraint(Constraint):
    def __init__(self, strength=1):
        self.strength = strength

    def to_ops(self, shape, tables):
        return [(nat.OP_SYMMETRY, 0, self.strength)]

    def __call__(self, morph, step):
        return run_on_device(self, morph)


class CenterOnConstraint(Constraint):
    def __init__(self, tiny=1e-6):
        self.tiny = tiny

    def to_ops(self, shape, tables):
        return [(nat.OP_CENTER_ON, 0, self.tiny)]

    def __call__(self, morph, step):
        morph[...] = run_on_device(self, morph)
        return morph
