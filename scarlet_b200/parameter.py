"""Optimisation parameters: ``numpy.ndarray`` subclass that carries optimiser state.

Mirrors scarlet/parameter.py (``Parameter`` 9-113, ``relative_step`` 126-129).  The device loop reads the value
and ``m/v/vhat`` on entry (warm start, blend.py:154-163) and writes them -- and ``std`` -- back in place after
the fit, so sources keep holding the same array objects.
"""
import numpy as np

from .prior import Prior


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

    _ATTRS = (("name", "unnamed"), ("prior", None), ("constraint", None), ("step", 0), ("std", None), ("m", None),
              ("v", None), ("vhat", None), ("fixed", False))

    def __array_finalize__(self, obj):
        if obj is None:
            return
        for key, default in self._ATTRS:
            setattr(self, key, getattr(obj, key, default))

    def __reduce__(self):
        base = super().__reduce__()
        return (base[0], base[1], base[2] + (self.__dict__,))

    def __setstate__(self, state):
        self.__dict__.update(state[-1])
        super().__setstate__(state[:-1])

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
