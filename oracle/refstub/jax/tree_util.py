"""Pytrees for the torch-backed jax stand-in (oracle/refstub/README.md): list / tuple / dict / namedtuple / None and
registered classes, with jax's flattening order (dict keys sorted)."""
from __future__ import annotations

import functools

_REGISTRY = {}


class _Leaf:
    def __repr__(self):
        return "*"


LEAF = _Leaf()


class TreeDef:
    def __init__(self, kind, meta, children):
        self.kind, self.meta, self.children = kind, meta, children

    @property
    def num_leaves(self):
        return 1 if self.kind is LEAF else sum(c.num_leaves for c in self.children)

    def __eq__(self, other):
        return isinstance(other, TreeDef) and self.kind == other.kind and self.meta == other.meta and self.children == other.children

    def __hash__(self):
        return hash((str(self.kind), len(self.children)))

    def __repr__(self):
        return f"TreeDef({self.kind}, {self.children})"


def register_pytree_node(cls, flatten, unflatten):
    _REGISTRY[cls] = (flatten, unflatten)


def register_pytree_node_class(cls):
    register_pytree_node(cls, lambda x: x.tree_flatten(), lambda aux, ch: cls.tree_unflatten(aux, ch))
    return cls


def register_dataclass(cls, data_fields=None, meta_fields=None, **_):
    import dataclasses

    if data_fields is None:
        data_fields = [f.name for f in dataclasses.fields(cls)]
        meta_fields = []
    data_fields, meta_fields = list(data_fields), list(meta_fields or [])

    def fl(x):
        return [getattr(x, n) for n in data_fields], tuple(getattr(x, n) for n in meta_fields)

    def un(meta, ch):
        obj = object.__new__(cls)
        for n, v in zip(data_fields, ch):
            object.__setattr__(obj, n, v)
        for n, v in zip(meta_fields, meta):
            object.__setattr__(obj, n, v)
        return obj

    register_pytree_node(cls, fl, un)
    return cls


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def _flatten(tree, is_leaf, leaves):
    if is_leaf is not None and is_leaf(tree):
        leaves.append(tree)
        return TreeDef(LEAF, None, [])
    if tree is None:
        return TreeDef(type(None), None, [])
    t = type(tree)
    if t in _REGISTRY:
        ch, aux = _REGISTRY[t][0](tree)
        ch = list(ch)
        return TreeDef(t, aux, [_flatten(c, is_leaf, leaves) for c in ch])
    if _is_namedtuple(tree):
        return TreeDef(t, "namedtuple", [_flatten(c, is_leaf, leaves) for c in tree])
    if isinstance(tree, (list, tuple)):
        return TreeDef(t, None, [_flatten(c, is_leaf, leaves) for c in tree])
    if isinstance(tree, dict):
        keys = sorted(tree.keys())
        return TreeDef(dict, tuple(keys), [_flatten(tree[k], is_leaf, leaves) for k in keys])
    leaves.append(tree)
    return TreeDef(LEAF, None, [])


def tree_flatten(tree, is_leaf=None):
    leaves = []
    td = _flatten(tree, is_leaf, leaves)
    return leaves, td


def _unflatten(td, it):
    if td.kind is LEAF:
        return next(it)
    if td.kind is type(None):
        return None
    ch = [_unflatten(c, it) for c in td.children]
    if td.kind in _REGISTRY:
        return _REGISTRY[td.kind][1](td.meta, ch)
    if td.meta == "namedtuple":
        return td.kind(*ch)
    if td.kind is dict:
        return dict(zip(td.meta, ch))
    return td.kind(ch)


def tree_unflatten(treedef, leaves):
    return _unflatten(treedef, iter(list(leaves)))


def tree_leaves(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[0]


def tree_structure(tree, is_leaf=None):
    return tree_flatten(tree, is_leaf)[1]


def _flatten_up_to(td, tree, out):
    """leaves of `tree` down to the structure of `td` (subtrees at td's leaf positions are kept whole)"""
    if td.kind is LEAF:
        out.append(tree)
        return
    if td.kind is type(None):
        return
    t = td.kind
    if t in _REGISTRY:
        ch = list(_REGISTRY[t][0](tree)[0])
    elif td.kind is dict:
        ch = [tree[k] for k in td.meta]
    else:
        ch = list(tree)
    if len(ch) != len(td.children):
        raise ValueError("tree structures do not match")
    for c, sub in zip(td.children, ch):
        _flatten_up_to(c, sub, out)


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, td = tree_flatten(tree, is_leaf)
    others = []
    for r in rest:
        o = []
        _flatten_up_to(td, r, o)
        others.append(o)
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


_NO = object()


def tree_reduce(f, tree, initializer=_NO, is_leaf=None):
    leaves = tree_leaves(tree, is_leaf)
    return functools.reduce(f, leaves) if initializer is _NO else functools.reduce(f, leaves, initializer)


def tree_all(tree):
    return all(tree_leaves(tree))


class Partial(functools.partial):
    pass
