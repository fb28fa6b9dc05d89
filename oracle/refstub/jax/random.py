import torch

from . import _state


def PRNGKey(seed):
    return torch.Generator().manual_seed(int(seed))


key = PRNGKey


def split(key, num=2):
    base = int(key.initial_seed())
    return [torch.Generator().manual_seed(base * 1000003 + i + 1) for i in range(num)]


def normal(key, shape=(), dtype=None):
    return torch.randn(tuple(shape), generator=key, dtype=dtype or _state.default_float())


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return torch.rand(tuple(shape), generator=key, dtype=dtype or _state.default_float()) * (maxval - minval) + minval
