"""Model tree bookkeeping.  Mirrors scarlet/model.py (parameters 51-54, get_parameter 71-110,
check_parameters 153-165)."""
from abc import ABC, abstractmethod

from .parameter import Parameter


class UpdateException(Exception):
    pass


class Model(ABC):
    def __init__(self, *parameters, children=None):
        for p in parameters:
            if not isinstance(p, Parameter):
                raise TypeError("parameters must be Parameter instances")
        self._parameters = tuple(parameters)
        if children is None:
            children = ()
        elif isinstance(children, Model):
            children = (children,)
        for c in children:
            if not isinstance(c, Model):
                raise TypeError("children must be Model instances")
        self._children = children
        self.check_parameters()

    @property
    def parameters(self):
        """Own parameters first, then the children's, depth first."""
        return self._parameters + tuple(p for c in self.children for p in c.parameters)

    @property
    def children(self):
        return self._children

    def __getitem__(self, i):
        return self._children[i]

    def __iter__(self):
        return iter(self._children)

    def get_parameter(self, i, *parameters):
        pool = parameters if parameters else self.parameters
        if isinstance(i, (int, slice)):
            return pool[i]
        if isinstance(i, str):
            match = tuple(p for p in pool if isinstance(p, Parameter) and p.name == i)
            if not match:
                return None
            return match[0] if len(match) == 1 else match
        return None

    @abstractmethod
    def get_model(self, *parameters, **kwargs):
        pass

    def get_models_of_children(self, *parameters, **kwargs):
        models = []
        if parameters:
            i = len(self._parameters)
            for c in self._children:
                j = len(c.parameters)
                models.append(c.get_model(*parameters[i:i + j], **kwargs))
                i += j
        else:
            models = [c.get_model(**kwargs) for c in self._children]
        return models

    def check_parameters(self):
        for p in self.parameters:
            if not p.is_finite:
                raise ArithmeticError("Model {}, Parameter '{}' is not finite:\n{}".format(type(self).__name__, p.name, p))

    def update(self):
        pass
