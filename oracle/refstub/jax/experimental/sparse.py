"""jax.experimental.sparse.BCOO as used by the reference (data/splitting/sparse_map.py): built from a dense matrix,
applied as `todense() @ x`."""
import torch

from ..numpy import _b


class BCOO:
    def __init__(self, dense):
        self._dense = _b(dense)
        self.shape = tuple(self._dense.shape)
        self.dtype = self._dense.dtype

    @classmethod
    def fromdense(cls, dense, **kw):
        return cls(dense)

    def todense(self):
        return self._dense

    def __matmul__(self, other):
        return self._dense @ _b(other)

    @property
    def nse(self):
        return int((self._dense != 0).sum())

    @property
    def T(self):
        return BCOO(self._dense.T)
