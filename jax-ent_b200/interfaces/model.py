"""Model_Parameters base (reference jaxent/src/interfaces/model.py:15-196): slots-based parameter
containers with `static_params`, arithmetic and `update_parameters`."""
from __future__ import annotations

from typing import Any, ClassVar


class Model_Parameters:
    key: Any
    static_params: ClassVar[set] = {"key"}

    @classmethod
    def _get_ordered_slots(cls) -> tuple:
        slots = []
        for c in cls.__mro__:
            if hasattr(c, "__slots__"):
                slots.extend(c.__slots__)
        return tuple(dict.fromkeys(slots))

    @classmethod
    def _get_grouped_slots(cls):
        dyn = tuple(s for s in cls._get_ordered_slots() if s not in cls.static_params)
        sta = tuple(s for s in cls._get_ordered_slots() if s in cls.static_params)
        return dyn, sta

    def tree_flatten(self):
        dyn, sta = self._get_grouped_slots()
        return tuple(getattr(self, s) for s in dyn), tuple(getattr(self, s) for s in sta)

    @classmethod
    def tree_unflatten(cls, static_data, arrays):
        dyn, sta = cls._get_grouped_slots()
        obj = object.__new__(cls)
        for k, v in list(zip(dyn, arrays)) + list(zip(sta, static_data)):
            object.__setattr__(obj, k, v)
        return obj

    def _map(self, fn):
        kw = {}
        for s in self._get_ordered_slots():
            v = getattr(self, s)
            kw[s] = v if s in self.static_params else fn(s, v)
        return self.__class__(**kw)

    def update_parameters(self, new_params):
        return self._map(lambda s, v: getattr(new_params, s))

    def __add__(self, other):
        return self._map(lambda s, v: v + getattr(other, s))

    def __sub__(self, other):
        return self._map(lambda s, v: v - getattr(other, s))

    def __mul__(self, scalar):
        return self._map(lambda s, v: v * scalar)

    __rmul__ = __mul__

    def __truediv__(self, scalar):
        return self._map(lambda s, v: v / scalar)

    def __neg__(self):
        return self._map(lambda s, v: -v)
