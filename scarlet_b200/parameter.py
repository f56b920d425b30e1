"""Optimisation parameters: ``numpy.ndarray`` subclass that carries optimiser state.

Mirrors scarlet/parameter.py (``Parameter`` 9-113, ``relative_step`` 126-129).  The device loop reads the value
and ``m/v/vhat`` on entry (warm start, blend.py:154-163) and writes them -- and ``std`` -- back after the fit, so
sources keep holding the same array objects.

The value is updated in place in the caller's array.  The optimiser state of all parameters of a plan comes back
from the device as a few packed host arrays; ``m``, ``v``, ``vhat`` are views into those arrays, materialised
when read (``_StateLink``), and ``std = 1/sqrt(masked v)`` (blend.py:189-192) is evaluated when read.  Assigning to
any of the four replaces it, exactly like setting the attribute on the reference's class.
"""
import numpy as np
import numpy.ma as ma

from .prior import Prior


class _StateLink:
    """Where the optimiser state of one parameter lives inside the packed host arrays of a device plan."""
    __slots__ = ("store", "group", "start", "stop", "shape")

    def __init__(self, store, group, start, stop, shape):
        self.store, self.group, self.start, self.stop, self.shape = store, group, start, stop, shape

    def view(self, key):
        if not self.store.valid:
            return None
        return self.store.arrays[key][self.group].reshape(-1)[self.start:self.stop].reshape(self.shape)


def _state_property(key):
    slot = "_" + key

    def get(self):
        val = self.__dict__.get(slot)
        if val is None:
            link = self.__dict__.get("_link")
            if link is not None:
                return link.view(key)
        return val

    def put(self, value):
        self.__dict__[slot] = value

    return property(get, put)


class Parameter(np.ndarray):
    def __new__(cls, array, name="unnamed", prior=None, constraint=None, step=0, std=None, m=None, v=None, vhat=None,
                fixed=False):
        array = np.asarray(array)
        obj = np.asarray(array, dtype=array.dtype).view(cls)
        obj.name = name
        if prior is not None and not isinstance(prior, Prior):
            raise TypeError("prior must be a scarlet_b200.Prior")
        obj.prior = prior
        if constraint is not None:
            from .constraint import Constraint, ConstraintChain
            if not isinstance(constraint, (Constraint, ConstraintChain)):
                raise TypeError("constraint must be a Constraint or ConstraintChain")
        obj.constraint = constraint
        obj.step = step
        obj.std = std
        obj.m, obj.v, obj.vhat = m, v, vhat
        obj.fixed = fixed
        return obj

    m = _state_property("m")
    v = _state_property("v")
    vhat = _state_property("vhat")

    @property
    def std(self):
        d = self.__dict__
        val = d.get("_std")
        if val is None:
            v, link = None, d.get("_link")
            if link is not None and link.store.valid:
                v = link.view("v")
            elif d.get("_std_from_v"):
                v = d.get("_v")
            if v is not None:
                return 1 / np.sqrt(ma.masked_equal(v, 0))
        return val

    @std.setter
    def std(self, value):
        self.__dict__["_std"] = value
        self.__dict__["_std_from_v"] = False

    _ATTRS = (("name", "unnamed"), ("prior", None), ("constraint", None), ("step", 0), ("std", None), ("m", None),
              ("v", None), ("vhat", None), ("fixed", False))

    def __array_finalize__(self, obj):
        if obj is None:
            return
        src = getattr(obj, "__dict__", None)
        if src:
            self.__dict__.update(src)  # attributes travel by reference (lazy state links included), as in the reference
        else:
            for key, default in self._ATTRS:
                setattr(self, key, default)

    def __reduce__(self):
        base = super().__reduce__()
        state = {key: getattr(self, key) for key, _ in self._ATTRS}
        for key in ("m", "v", "vhat", "std"):
            if state[key] is not None:
                state[key] = state[key].copy()
        return (base[0], base[1], base[2] + (state,))

    def __setstate__(self, state):
        super().__setstate__(state[:-1])
        for key, value in state[-1].items():
            setattr(self, key, value)

    @property
    def _data(self):
        return self.view(np.ndarray)

    @property
    def is_finite(self):
        return bool(np.isfinite(self._data).all())


def prepare_param(X, name, fixed=True, step=None):
    if isinstance(X, Parameter):
        assert X.name == name
        return X
    if np.isscalar(X):
        X = (X,)
    return Parameter(np.array(X, dtype="float"), name=name, fixed=fixed, step=step)


def relative_step(X, it, factor=0.1, minimum=0, axis=None):
    """Step = ``factor`` x mean(X), floored at ``minimum``.  The device evaluates this rule itself every
    iteration (csrc/kernels.cuh: sed_update); the plan builder recognises this function by identity."""
    return np.maximum(minimum, factor * X.mean(axis=axis))
