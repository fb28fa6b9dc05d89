"""jaxent_b200 -- B200-native (sm_100a) BV MaxEnt optimise step behind JAX-ENT's API.

The directory is named ``jax-ent_b200`` (the framework name); ``import jaxent_b200`` resolves
to it through the shim package next to it.  Only the hot path of SURVEY.md section 8 lives here.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
