"""optax 0.2.4 stand-in (oracle/refstub/README.md): the transformations the reference's OptaxOptimizer builds
(opt/optimiser.py:103-148), restated from optax's documented update rules.  optax itself is not installable here, so these
restatements are NOT pinned against optax (DESIGN.md says so); they only let the reference's own step code execute."""
from __future__ import annotations

from typing import Any, Callable, NamedTuple

import torch

import jax
import jax.numpy as jnp
from jax import tree_util as tu

from . import losses, projections  # noqa: F401

OptState = Any
Params = Any
Updates = Any


class GradientTransformation(NamedTuple):
    init: Callable
    update: Callable


GradientTransformationExtraArgs = GradientTransformation


class EmptyState(NamedTuple):
    pass


def identity():
    return GradientTransformation(lambda p: EmptyState(), lambda u, s, p=None, **k: (u, s))


def clip(max_delta):
    """optax.clip: elementwise clip of the updates to [-max_delta, max_delta]"""

    def update(updates, state, params=None, **_):
        return tu.tree_map(lambda g: jnp.clip(g, -max_delta, max_delta), updates), state

    return GradientTransformation(lambda p: EmptyState(), update)


class ScaleByAdamState(NamedTuple):
    count: Any
    mu: Any
    nu: Any


def scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    """optax.scale_by_adam: mu = b1 mu + (1-b1) g, nu = b2 nu + (1-b2) g^2, count += 1,
    mu_hat = mu / (1 - b1^count), nu_hat = nu / (1 - b2^count), u = mu_hat / (sqrt(nu_hat + eps_root) + eps)"""

    def init(params):
        z = lambda p: jnp.zeros_like(p)
        return ScaleByAdamState(jnp.zeros((), dtype=jnp.int32), tu.tree_map(z, params), tu.tree_map(z, params))

    def update(updates, state, params=None, **_):
        mu = tu.tree_map(lambda g, t: (1 - b1) * g + b1 * t, updates, state.mu)
        nu = tu.tree_map(lambda g, t: (1 - b2) * (g * g) + b2 * t, updates, state.nu)
        count = state.count + jnp.asarray(1, dtype=jnp.int32)
        # optax: `1 - decay**count` with a Python-float decay and an int32 count: the default float type, i.e. fp32 unless
        # x64 is enabled (fl32(0.999) makes 1 - b2^1 differ from 1e-3 by 1.3e-5 relative -- part of the reference's arithmetic)
        from jax import _state

        fd = _state.default_float()
        cf = count.to(fd)
        bc1 = 1 - jnp.asarray(b1, dtype=fd) ** cf
        bc2 = 1 - jnp.asarray(b2, dtype=fd) ** cf
        mu_hat = tu.tree_map(lambda t: t / bc1.to(t.dtype), mu)
        nu_hat = tu.tree_map(lambda t: t / bc2.to(t.dtype), nu)
        out = tu.tree_map(lambda m, v: m / (jnp.sqrt(v + eps_root) + eps), mu_hat, nu_hat)
        return out, ScaleByAdamState(count, mu, nu)

    return GradientTransformation(init, update)


def scale(step_size):
    return GradientTransformation(lambda p: EmptyState(), lambda u, s, p=None, **k: (tu.tree_map(lambda g: step_size * g, u), s))


def chain(*transforms):
    def init(params):
        return tuple(t.init(params) for t in transforms)

    def update(updates, state, params=None, **kw):
        new = []
        for t, s in zip(transforms, state):
            updates, s2 = t.update(updates, s, params, **kw)
            new.append(s2)
        return updates, tuple(new)

    return GradientTransformation(init, update)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0, **_):
    return chain(scale_by_adam(b1, b2, eps, eps_root), scale(-learning_rate))


def sgd(learning_rate, momentum=None, nesterov=False, **_):
    if momentum is not None:
        raise NotImplementedError("refstub: sgd momentum")
    return chain(scale(-learning_rate))


def _unsupported(name):
    def f(*a, **k):
        raise NotImplementedError(f"refstub: optax.{name} is not restated (not on the BV MaxEnt path)")

    return f


adagrad, adamw, rmsprop, lbfgs = (_unsupported(n) for n in ("adagrad", "adamw", "rmsprop", "lbfgs"))


class InjectHyperparamsState(NamedTuple):
    count: Any
    hyperparams: dict
    inner_state: Any


def inject_hyperparams(inner_factory, static_args=(), hyperparam_dtype=None):
    """optax.inject_hyperparams: numeric hyper-parameters live in the state as arrays and are fed to the factory each update"""

    def wrapped(*args, **kwargs):
        if args:
            raise NotImplementedError("refstub: inject_hyperparams with positional arguments")
        numeric = {k: jnp.asarray(v, dtype=jnp.float32) for k, v in kwargs.items() if isinstance(v, (int, float)) and not isinstance(v, bool)}
        other = {k: v for k, v in kwargs.items() if k not in numeric}

        def init(params):
            return InjectHyperparamsState(jnp.zeros((), dtype=jnp.int32), dict(numeric), inner_factory(**other, **numeric).init(params))

        def update(updates, state, params=None, **kw):
            hp = state.hyperparams
            upd, inner = inner_factory(**other, **hp).update(updates, state.inner_state, params, **kw)
            return upd, InjectHyperparamsState(state.count + jnp.asarray(1, dtype=jnp.int32), hp, inner)

        return GradientTransformation(init, update)

    return wrapped


def keep_params_nonnegative():
    """optax.keep_params_nonnegative: where params + updates < 0 the update becomes -params"""

    def update(updates, state, params=None, **_):
        if params is None:
            raise ValueError("keep_params_nonnegative needs params")
        return tu.tree_map(lambda p, u: jnp.where((p + u) < 0.0, -p, u), params, updates), state

    return GradientTransformation(lambda p: EmptyState(), update)


class MultiTransformState(NamedTuple):
    inner_states: dict


def multi_transform(transforms, param_labels, mask_compatible_extra_args=False):
    """optax.multi_transform: every leaf is updated by the transformation its label names; labels may be a prefix tree"""

    def labels_of(params):
        lab = param_labels(params) if callable(param_labels) else param_labels
        lab_leaves, lab_def = tu.tree_flatten(lab, is_leaf=lambda x: isinstance(x, str))
        subtrees = []
        tu._flatten_up_to(lab_def, params, subtrees)
        out = []
        for name, sub in zip(lab_leaves, subtrees):
            out.extend([name] * len(tu.tree_leaves(sub)))
        return out

    def split(tree, labs):
        leaves, td = tu.tree_flatten(tree)
        if len(leaves) != len(labs):
            raise ValueError("refstub multi_transform: label / leaf mismatch")
        groups = {k: [v for v, l in zip(leaves, labs) if l == k] for k in transforms}
        return leaves, td, groups

    def init(params):
        labs = labels_of(params)
        _, _, groups = split(params, labs)
        return MultiTransformState({k: t.init(groups[k]) for k, t in transforms.items()})

    def update(updates, state, params=None, **kw):
        labs = labels_of(updates)
        leaves, td, ugroups = split(updates, labs)
        pgroups = split(params, labs)[2] if params is not None else {k: None for k in transforms}
        new_states, outs = {}, {}
        for k, t in transforms.items():
            outs[k], new_states[k] = t.update(ugroups[k], state.inner_states[k], pgroups[k])
        its = {k: iter(v) for k, v in outs.items()}
        merged = [next(its[l]) for l in labs]
        return tu.tree_unflatten(td, merged), MultiTransformState(new_states)

    return GradientTransformation(init, update)


def apply_updates(params, updates):
    return tu.tree_map(lambda p, u: p if u is None else (p + u).to(p.dtype) if isinstance(p, torch.Tensor) else p + u, params, updates)


def global_norm(updates):
    return jnp.sqrt(sum(jnp.sum(jnp.square(x)) for x in tu.tree_leaves(updates)))
