"""Inert stand-in for jax (test infrastructure; see oracle/refshim/README.md)."""
import functools

from . import numpy  # noqa: F401
from . import scipy  # noqa: F401


class Array:  # isinstance(x, jax.Array) is always False for NumPy data
    pass


class _Config:
    def update(self, *a, **k):
        return None


config = _Config()


def jit(fn=None, **kwargs):
    if fn is None:
        return lambda f: f
    return fn


def devices(*a, **k):
    return []


def default_backend():
    return "cpu"


class _Lax:
    def __getattr__(self, name):
        raise NotImplementedError(f"jax.lax.{name} is not available in the shim")


lax = _Lax()
partial = functools.partial
