"""Makes `import jaxent.src...` resolve to the UNMODIFIED reference sources under /root/reference with the stand-ins of
this directory for jax / optax / chex / MDAnalysis (oracle/refstub/README.md).  The two package __init__ files that pull
in the runtime configuration and the beartype import hook (jaxent/__init__.py, jaxent/src/__init__.py) are bypassed by
pre-seeding plain namespace modules; every other file is the reference's own, executed as is."""
from __future__ import annotations

import os
import sys
import types

import importlib.abc
import importlib.machinery
import importlib.util

HERE = os.path.dirname(os.path.abspath(__file__))
# third-party packages the reference imports at module scope for code that is NOT on the optimise path (trajectory
# handling, plotting, HDF5, intrinsic-rate tables): phantom modules whose every attribute is an inert placeholder class
PHANTOM = ("MDAnalysis", "hdxrate", "icecream", "matplotlib", "h5py", "networkx", "seaborn", "pymol", "mdtraj", "Bio")


class _PhantomMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Phantom

    def __or__(cls, other):  # `X | None` in annotations evaluated at import time
        return cls

    __ror__ = __or__


class _Phantom(metaclass=_PhantomMeta):
    def __init__(self, *a, **k):
        raise RuntimeError("refstub: a phantom third-party object was instantiated -- this code path is outside the BV optimise step")

    def __class_getitem__(cls, item):
        return cls


class _PhantomModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Phantom


class _PhantomFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in PHANTOM and importlib.machinery.PathFinder.find_spec(root) is None:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _PhantomModule(spec.name)

    def exec_module(self, module):
        pass
REFERENCE = os.environ.get("JAXENT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "jaxent", "src"))


def load(x64: bool = False):
    if not available():
        raise RuntimeError(f"reference sources not found under {REFERENCE}")
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    if not any(isinstance(f, _PhantomFinder) for f in sys.meta_path):
        sys.meta_path.append(_PhantomFinder())
    import jax  # noqa: F401  (the stand-in)

    if not getattr(jax, "__version__", "").endswith("refstub"):
        raise RuntimeError("a real jax is importable: use it instead of the stand-in")
    jax.config.update("jax_enable_x64", x64)
    for name, path in (("jaxent", os.path.join(REFERENCE, "jaxent")), ("jaxent.src", os.path.join(REFERENCE, "jaxent", "src"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [path]
            sys.modules[name] = m
    sys.modules["jaxent"].src = sys.modules["jaxent.src"]
    return jax
