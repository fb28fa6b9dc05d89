"""A torch-backed stand-in for the slice of `jax` the reference's optimise path uses (oracle/refstub/README.md).
Test infrastructure only: lets /root/reference/jaxent/src run unmodified in an image without jax."""
from __future__ import annotations

import functools

import torch

from . import _state
from . import numpy  # noqa: F401  (jax.numpy)
from . import tree_util  # noqa: F401
from . import core, experimental, lax, nn, random  # noqa: F401
from .numpy import _dt

Array = torch.Tensor
__version__ = "0.4.35+refstub"
torch.set_grad_enabled(True)
torch.set_num_threads(max(1, torch.get_num_threads()))


# ---- array methods jax arrays have and torch tensors lack
class _AtIdx:
    def __init__(self, t, idx):
        self.t, self.idx = t, idx

    def set(self, v):
        out = self.t.clone()
        out[self.idx] = v if not isinstance(v, torch.Tensor) else v.to(out.dtype)
        return out

    def add(self, v):
        out = self.t.clone()
        out[self.idx] += v
        return out


class _At:
    def __init__(self, t):
        self.t = t

    def __getitem__(self, idx):
        return _AtIdx(self.t, idx)


torch.Tensor.astype = lambda self, d: self.to(_dt(d))
torch.Tensor.at = property(lambda self: _At(self))
torch.Tensor.block_until_ready = lambda self: self
torch.Tensor.copy = lambda self: self.clone()
if not hasattr(torch.Tensor, "_refstub_array"):
    _orig_array = torch.Tensor.__array__

    def _array(self, dtype=None, *a, **k):  # np.asarray(x) on a tensor that still carries autograd history
        t = self.detach()
        return _orig_array(t) if dtype is None else _orig_array(t, dtype)

    torch.Tensor.__array__ = _array
    torch.Tensor._refstub_array = True


if not hasattr(torch.Tensor, "_refstub_getitem"):
    _orig_getitem = torch.Tensor.__getitem__

    def _has_neg_step(idx):
        items = idx if isinstance(idx, tuple) else (idx,)
        return any(isinstance(i, slice) and i.step is not None and i.step < 0 for i in items)

    def _getitem(self, idx):  # numpy-style negative-step slices (x[::-1]), which torch rejects
        if not _has_neg_step(idx):
            return _orig_getitem(self, idx)
        items = list(idx) if isinstance(idx, tuple) else [idx]
        out, dim = self, 0
        for it in items:
            if isinstance(it, slice) and it.step is not None and it.step < 0:
                sel = torch.tensor(list(range(*it.indices(out.shape[dim]))), dtype=torch.long)
                out = torch.index_select(out, dim, sel)
                dim += 1
            elif it is None:
                out = out.unsqueeze(dim)
                dim += 1
            elif isinstance(it, slice):
                out = _orig_getitem(out, (slice(None),) * dim + (it,))
                dim += 1
            else:
                out = _orig_getitem(out, (slice(None),) * dim + (it,))
        return out

    torch.Tensor.__getitem__ = _getitem
    torch.Tensor._refstub_getitem = True


class _Config:
    def update(self, key, value):
        if key == "jax_enable_x64":
            _state.x64 = bool(value)

    def __getattr__(self, k):
        return None


config = _Config()


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def _is_float(t):
    return isinstance(t, torch.Tensor) and t.dtype.is_floating_point


def value_and_grad(fn, argnums=0, has_aux=False, allow_int=False, **_):
    """jax.value_and_grad through torch.autograd: gradients for the float leaves of args[argnums], zeros elsewhere"""
    if not isinstance(argnums, int):
        raise NotImplementedError("refstub: tuple argnums")

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        leaves, td = tree_util.tree_flatten(args[argnums])
        live = []
        for v in leaves:
            if _is_float(v):
                live.append(v.detach().clone().requires_grad_(True))
            else:
                live.append(v)
        args2 = list(args)
        args2[argnums] = tree_util.tree_unflatten(td, live)
        with torch.enable_grad():
            out = fn(*args2, **kwargs)
            val, aux = (out if has_aux else (out, None))
            wrt = [v for v in live if _is_float(v)]
            gs = torch.autograd.grad(val, wrt, allow_unused=True) if wrt else []
        it = iter(gs)
        grads = []
        for v in live:
            if _is_float(v):
                g = next(it)
                grads.append(torch.zeros_like(v) if g is None else g.detach())
            elif isinstance(v, torch.Tensor):
                grads.append(torch.zeros_like(v))
            else:
                grads.append(v)
        det = lambda t: t.detach() if isinstance(t, torch.Tensor) else t
        g_tree = tree_util.tree_unflatten(td, grads)
        val = det(val)
        if has_aux:
            return (val, tree_util.tree_map(det, aux)), g_tree
        return val, g_tree

    return wrapped


def grad(fn, argnums=0, has_aux=False, **kw):
    vg = value_and_grad(fn, argnums=argnums, has_aux=has_aux, **kw)

    def wrapped(*a, **k):
        v, g = vg(*a, **k)
        return (g, v[1]) if has_aux else g

    return wrapped


def vmap(fn, in_axes=0, out_axes=0):
    """sequential stand-in: slices every mapped argument's leaves along axis 0, stacks the outputs' leaves"""

    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            for leaf in tree_util.tree_leaves(a):
                n = leaf.shape[0]
                break
            if n is not None:
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree_util.tree_map(lambda v: v[i], a) for a, ax in zip(args, axes)]
            outs.append(fn(*sl))
        return tree_util.tree_map(lambda *vs: numpy.stack([numpy.asarray(v) for v in vs]), outs[0], *outs[1:])

    return wrapped


def clear_caches():
    pass


def devices(*a, **k):
    return []


def default_backend():
    return "cpu"


def device_get(x):
    return x


def block_until_ready(x):
    return x
