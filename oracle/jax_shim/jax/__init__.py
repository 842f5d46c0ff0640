"""Minimal stand-in for the ``jax`` package.  TEST INFRASTRUCTURE ONLY.

The reference (erinhay/GausSN) does its math through ``jax.numpy`` with
``jax_enable_x64`` switched on; JAX is not installed in the build container and
cannot be (no network).  This shim lets ``tests/golden/make_golden.py`` import
the reference's *own, unmodified* source files from ``/root/reference`` and run
them on numpy float64, so the golden vectors come from the reference's code and
not from a restatement.  Only the names the reference uses are provided.

Behavioural differences from real JAX that matter on this path are restated in
``jax/numpy/__init__.py`` (``linalg.cholesky``: symmetrise the input and return
NaN instead of raising).  Elementary functions come from numpy's libm rather
than XLA's, which can differ from real JAX in the last ulp.
"""
from . import numpy  # noqa: F401
from . import scipy  # noqa: F401


def jit(fun, *args, **kwargs):
    """Identity: the shim executes eagerly."""
    return fun


class _Config:
    def update(self, name, value):
        if name == "jax_enable_x64" and not value:
            raise ValueError("the shim is float64-only")


config = _Config()
