"""Prior interface (scarlet/prior.py:3-19).  The device path has no prior support: a Parameter that carries a
prior makes ``Blend.fit`` raise instead of silently running the prior on the host."""
from abc import ABC, abstractmethod


class Prior(ABC):
    @abstractmethod
    def __call__(self, x):
        pass

    @abstractmethod
    def grad(self, x):
        pass
