"""jax.lax control flow executed eagerly in Python"""
import torch

from . import tree_util
from .numpy import _b, asarray


def select(pred, on_true, on_false):
    p = bool(pred) if not (isinstance(pred, torch.Tensor) and pred.dim() > 0) else None
    if p is None:
        return torch.where(pred, _b(on_true), _b(on_false))
    return _b(on_true) if p else _b(on_false)


def cond(pred, true_fn, false_fn, *operands):
    return true_fn(*operands) if bool(pred) else false_fn(*operands)


def while_loop(cond_fn, body_fn, init):
    c = init
    while bool(cond_fn(c)):
        c = body_fn(c)
    return c


def fori_loop(lo, hi, body_fn, init):
    c = init
    for i in range(int(lo), int(hi)):
        c = body_fn(asarray(i), c)
    return c


def scan(f, init, xs, length=None):
    n = length if xs is None else tree_util.tree_leaves(xs)[0].shape[0]
    c, ys = init, []
    for i in range(n):
        c, y = f(c, None if xs is None else tree_util.tree_map(lambda v: v[i], xs))
        ys.append(y)
    return c, (tree_util.tree_map(lambda *v: torch.stack(v), ys[0], *ys[1:]) if ys and ys[0] is not None else None)


def map(f, xs):
    n = tree_util.tree_leaves(xs)[0].shape[0]
    ys = [f(tree_util.tree_map(lambda v: v[i], xs)) for i in range(n)]
    return tree_util.tree_map(lambda *v: torch.stack([_b(u) for u in v]), ys[0], *ys[1:])


def dynamic_update_slice(operand, update, start_indices):
    out = operand.clone()
    idx = tuple(slice(int(s), int(s) + d) for s, d in zip(start_indices, update.shape))
    out[idx] = update.to(out.dtype)
    return out


def dynamic_slice(operand, start_indices, slice_sizes):
    return operand[tuple(slice(int(s), int(s) + d) for s, d in zip(start_indices, slice_sizes))]


def stop_gradient(x):
    return tree_util.tree_map(lambda v: v.detach() if isinstance(v, torch.Tensor) else v, x)
