"""jax.numpy stand-in on torch CPU tensors (oracle/refstub/README.md).  Default float is fp32 -- jax without x64 --
unless jax.config.update("jax_enable_x64", True) was called; 64-bit inputs are then kept."""
from __future__ import annotations

import builtins
import math

import numpy as _np
import torch

from .. import _state

ndarray = torch.Tensor
float32, float64, float16, bfloat16 = torch.float32, torch.float64, torch.float16, torch.bfloat16
int32, int64, int8, uint8, bool_ = torch.int32, torch.int64, torch.int8, torch.uint8, torch.bool
inf, pi, nan, newaxis, e = math.inf, math.pi, math.nan, None, math.e

_NP2T = {_np.dtype("float32"): torch.float32, _np.dtype("float64"): torch.float64, _np.dtype("int32"): torch.int32,
         _np.dtype("int64"): torch.int64, _np.dtype("bool"): torch.bool, _np.dtype("float16"): torch.float16,
         _np.dtype("int8"): torch.int8, _np.dtype("uint8"): torch.uint8}


def _dt(d):
    if d is None or isinstance(d, torch.dtype):
        return _canon(d)
    if d is float:
        return _state.default_float()
    if d is int:
        return _state.default_int()
    if d is bool:
        return torch.bool
    return _canon(_NP2T[_np.dtype(d)])


def _canon(d):
    """jax truncates 64-bit types unless x64 is enabled"""
    if d is None or _state.x64:
        return d
    return {torch.float64: torch.float32, torch.int64: torch.int32}.get(d, d)


def asarray(x, dtype=None):
    d = _dt(dtype)
    if isinstance(x, torch.Tensor):
        t = x
    elif isinstance(x, _np.ndarray):
        t = torch.from_numpy(_np.ascontiguousarray(x))
    elif isinstance(x, (bool, _np.bool_)):
        t = torch.tensor(bool(x))
    elif isinstance(x, (int, _np.integer)):
        t = torch.tensor(int(x), dtype=_state.default_int())
    elif isinstance(x, (float, _np.floating)):
        t = torch.tensor(float(x), dtype=_state.default_float())
    elif isinstance(x, (list, tuple)):
        if len(x) and builtins.any(isinstance(v, torch.Tensor) for v in x):
            t = torch.stack([asarray(v) for v in x])
        else:
            a = _np.asarray(x)
            t = asarray(a) if a.dtype != object else torch.stack([asarray(v) for v in x])
    else:
        t = torch.from_numpy(_np.asarray(x))
    want = d if d is not None else _canon(t.dtype)
    return t if t.dtype == want else t.to(want)


array = asarray


def _b(x):
    return x if isinstance(x, torch.Tensor) else asarray(x)


def _pair(a, b):
    """binary-op operands with jax's weak typing: a Python scalar adopts the array's dtype"""
    ta, tb = isinstance(a, torch.Tensor), isinstance(b, torch.Tensor)
    if ta and not tb and isinstance(b, (int, float, bool)):
        return a, b
    if tb and not ta and isinstance(a, (int, float, bool)):
        return a, b
    return _b(a), _b(b)


def add(a, b): a, b = _pair(a, b); return a + b
def subtract(a, b): a, b = _pair(a, b); return a - b
def multiply(a, b): a, b = _pair(a, b); return a * b
def divide(a, b): a, b = _pair(a, b); return a / b
def power(a, b): a, b = _pair(a, b); return a ** b
def maximum(a, b): a, b = _pair(a, b); return torch.maximum(_b(a).to(_res(a, b)), _b(b).to(_res(a, b)))
def minimum(a, b): a, b = _pair(a, b); return torch.minimum(_b(a).to(_res(a, b)), _b(b).to(_res(a, b)))


def _res(a, b):
    if isinstance(a, torch.Tensor) and not isinstance(b, torch.Tensor):
        return a.dtype if a.dtype.is_floating_point or not isinstance(b, float) else _state.default_float()
    if isinstance(b, torch.Tensor) and not isinstance(a, torch.Tensor):
        return b.dtype if b.dtype.is_floating_point or not isinstance(a, float) else _state.default_float()
    return torch.result_type(a, b)


def _f(x):
    """float view of x for transcendental functions"""
    x = _b(x)
    return x if x.dtype.is_floating_point else x.to(_state.default_float())


def exp(x): return torch.exp(_f(x))
def log(x): return torch.log(_f(x))
def log1p(x): return torch.log1p(_f(x))
def sqrt(x): return torch.sqrt(_f(x))
def square(x): return _b(x) * _b(x)
def abs(x): return torch.abs(_b(x))
absolute = abs
def sign(x): return torch.sign(_b(x))
def tanh(x): return torch.tanh(_f(x))
def round(x, decimals=0): return torch.round(_b(x), decimals=decimals)
def isnan(x): return torch.isnan(_b(x))
def isinf(x): return torch.isinf(_b(x))
def isfinite(x): return torch.isfinite(_b(x))
def logical_and(a, b): return torch.logical_and(_b(a), _b(b))
def logical_or(a, b): return torch.logical_or(_b(a), _b(b))
def logical_not(a): return torch.logical_not(_b(a))


def _ax(axis):
    return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def sum(x, axis=None, keepdims=False, dtype=None):
    x = _b(x)
    if x.dtype == torch.bool:
        x = x.to(_state.default_int())
    return torch.sum(x, dtype=_dt(dtype)) if axis is None and not keepdims else torch.sum(x, dim=_ax(axis) if axis is not None else tuple(range(x.dim())), keepdim=keepdims, dtype=_dt(dtype))


def mean(x, axis=None, keepdims=False):
    x = _f(x)
    return torch.mean(x) if axis is None and not keepdims else torch.mean(x, dim=_ax(axis) if axis is not None else tuple(range(x.dim())), keepdim=keepdims)


def max(x, axis=None, keepdims=False):
    x = _b(x)
    return torch.max(x) if axis is None else torch.amax(x, dim=_ax(axis), keepdim=keepdims)


def min(x, axis=None, keepdims=False):
    x = _b(x)
    return torch.min(x) if axis is None else torch.amin(x, dim=_ax(axis), keepdim=keepdims)


def any(x, axis=None): return torch.any(_b(x)) if axis is None else torch.any(_b(x), dim=axis)
def all(x, axis=None): return torch.all(_b(x)) if axis is None else torch.all(_b(x), dim=axis)
def argmin(x, axis=None): return torch.argmin(_b(x)) if axis is None else torch.argmin(_b(x), dim=axis)
def argmax(x, axis=None): return torch.argmax(_b(x)) if axis is None else torch.argmax(_b(x), dim=axis)
def cumsum(x, axis=0): return torch.cumsum(_b(x), dim=axis)
def var(x, axis=None): return torch.var(_f(x), unbiased=False) if axis is None else torch.var(_f(x), dim=axis, unbiased=False)
def std(x, axis=None): return torch.std(_f(x), unbiased=False) if axis is None else torch.std(_f(x), dim=axis, unbiased=False)


def clip(x, min=None, max=None, a_min=None, a_max=None):
    lo = min if min is not None else a_min
    hi = max if max is not None else a_max
    return torch.clamp(_b(x), lo, hi)


def where(c, a, b):
    a, b = _pair(a, b)
    return torch.where(_b(c).bool(), a if isinstance(a, torch.Tensor) else torch.tensor(a, dtype=_res(a, b)), b if isinstance(b, torch.Tensor) else torch.tensor(b, dtype=_res(a, b)))


def dot(a, b):
    a, b = _b(a), _b(b)
    if a.dim() == 1 and b.dim() == 1:
        return torch.dot(a, b)
    return torch.matmul(a, b)


def vdot(a, b): return torch.sum(_b(a).reshape(-1) * _b(b).reshape(-1))
def matmul(a, b): return torch.matmul(_b(a), _b(b))
def outer(a, b): return torch.outer(_b(a).reshape(-1), _b(b).reshape(-1))
def trace(x): return torch.trace(_b(x))
def diag(x): return torch.diag(_b(x))
def transpose(x, axes=None): x = _b(x); return x.permute(*axes) if axes is not None else x.permute(*reversed(range(x.dim())))
def squeeze(x, axis=None): return torch.squeeze(_b(x)) if axis is None else torch.squeeze(_b(x), dim=axis)
def expand_dims(x, axis): return torch.unsqueeze(_b(x), axis)
def reshape(x, shape): return torch.reshape(_b(x), shape if isinstance(shape, (tuple, list)) else (shape,))
def ravel(x): return _b(x).reshape(-1)
def concatenate(xs, axis=0): return torch.cat([_b(v) for v in xs], dim=axis)
def stack(xs, axis=0): return torch.stack([_b(v) for v in xs], dim=axis)
def repeat(x, n, axis=None): x = _b(x); return torch.repeat_interleave(x.reshape(-1) if axis is None else x, n, dim=0 if axis is None else axis)
def tile(x, reps): return torch.tile(_b(x), reps if isinstance(reps, (tuple, list)) else (reps,))
def sort(x, axis=-1): return torch.sort(_b(x), dim=axis).values
def argsort(x, axis=-1): return torch.argsort(_b(x), dim=axis)
def flip(x, axis=None): x = _b(x); return torch.flip(x, dims=tuple(range(x.dim())) if axis is None else (axis,))
def take(x, idx, axis=None): return _b(x).reshape(-1)[_b(idx)] if axis is None else torch.index_select(_b(x), axis, _b(idx))
def broadcast_to(x, shape): return torch.broadcast_to(_b(x), shape)


def _shape(s):
    return tuple(s) if isinstance(s, (tuple, list, torch.Size)) else (s,)


def zeros(shape, dtype=None): return torch.zeros(_shape(shape), dtype=_dt(dtype) or _state.default_float())
def ones(shape, dtype=None): return torch.ones(_shape(shape), dtype=_dt(dtype) or _state.default_float())
def empty(shape, dtype=None): return zeros(shape, dtype)
def full(shape, v, dtype=None): return torch.full(_shape(shape), float(v) if not isinstance(v, (bool, int)) else v, dtype=_dt(dtype) or (asarray(v).dtype))
def zeros_like(x, dtype=None): return torch.zeros_like(_b(x), dtype=_dt(dtype))
def ones_like(x, dtype=None): return torch.ones_like(_b(x), dtype=_dt(dtype))
def full_like(x, v, dtype=None): return torch.full_like(_b(x), v, dtype=_dt(dtype))
def eye(n, m=None, dtype=None): return torch.eye(n, n if m is None else m, dtype=_dt(dtype) or _state.default_float())
def arange(*a, dtype=None):
    isf = builtins.any(isinstance(v, float) for v in a)
    return torch.arange(*a, dtype=_dt(dtype) or (_state.default_float() if isf else _state.default_int()))
def linspace(a, b, n, dtype=None): return torch.linspace(a, b, n, dtype=_dt(dtype) or _state.default_float())
def triu_indices(n, k=0, m=None): i = torch.triu_indices(n, n if m is None else m, offset=k); return i[0], i[1]
def ix_(*a): return torch.meshgrid(*[_b(v) for v in a], indexing="ij")
def allclose(a, b, rtol=1e-5, atol=1e-8): return bool(torch.allclose(_f(a), _f(b).to(_f(a).dtype), rtol=rtol, atol=atol))
def array_equal(a, b): return bool(torch.equal(_b(a), _b(b)))
def isscalar(x): return isinstance(x, (int, float, bool))
def shape(x): return tuple(_b(x).shape)
def ndim(x): return _b(x).dim()
def size(x): return _b(x).numel()


def savez(file, *args, **kw):
    _np.savez(file, *[_np.asarray(v) for v in args], **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else _np.asarray(v)) for k, v in kw.items()})


def load(file, *a, **kw):
    return _np.load(file, *a, **kw)


class _Linalg:
    @staticmethod
    def norm(x, ord=None, axis=None):
        return torch.linalg.norm(_f(x)) if axis is None and ord is None else torch.linalg.norm(_f(x), ord=ord, dim=axis)

    @staticmethod
    def inv(x): return torch.linalg.inv(_f(x))

    @staticmethod
    def pinv(x): return torch.linalg.pinv(_f(x))


linalg = _Linalg()

integer, floating, number, generic = _np.integer, _np.floating, _np.number, _np.generic
dtype = _np.dtype
