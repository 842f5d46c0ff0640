"""Batch adapters between third-party samplers and the batched likelihood.

The reference hands ``GP.jointprobability`` to dynesty / emcee / zeus / scipy, which call it with ONE
parameter vector at a time (gaussn/gausSN.py:321-323, :335, :352-359).  These adapters give the same
samplers whole batches to evaluate, which is what the GPU engine is built for:

* ``VectorizedPosterior`` -- callable on ``[n, ndim]`` (emcee / zeus ``vectorize=True``) or on ``[ndim]``;
* ``BatchPool``           -- a pool-like object whose ``map`` recognises the posterior and evaluates the whole
                             queue in one launch (dynesty ``pool=``, ``queue_size=``, ``sample='unif'`` and the
                             initial live points); any other function is mapped serially;
* ``fd_value_and_grad``   -- objective value and forward-difference gradient from one batch of P + 1 points,
                             for ``scipy.optimize.minimize(..., jac=True)`` (the reference lets SciPy difference
                             the scalar function: P + 1 separate calls per gradient).
"""
from __future__ import annotations

import numpy as np


class VectorizedPosterior:
    """``invert * (logL + log prior)`` for one or many parameter vectors (gaussn/gausSN.py:162-206)."""

    def __init__(self, gp, logprior=None, fix_kernel_params=False, fix_mean_params=False, fix_lensing_params=False, invert=1):
        self.gp = gp
        self.logprior = logprior
        self.fix = (fix_kernel_params, fix_mean_params, fix_lensing_params)
        self.invert = invert
        self.n_calls = 0          # batched launches
        self.n_evals = 0          # parameter vectors evaluated

    def __call__(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        single = theta.ndim == 1
        batch = theta.reshape(1, -1) if single else theta
        out = np.asarray(self.gp.jointprobability_batch(batch, self.logprior, *self.fix, invert=self.invert))
        self.n_calls += 1
        self.n_evals += len(batch)
        return float(out[0]) if single else out


class BatchPool:
    """Pool-like adapter: ``map(f, points)`` with ``f`` a ``VectorizedPosterior`` (or a function that carries one as
    ``f.vectorized``) evaluates all points in one batched launch.  ``size`` is what dynesty reads as the default
    ``queue_size``."""

    def __init__(self, size=512):
        self.size = int(size)

    def map(self, func, iterable, chunksize=None):
        items = list(iterable)
        vec = func if isinstance(func, VectorizedPosterior) else getattr(func, "vectorized", None)
        if vec is not None and items and all(np.ndim(it) == 1 for it in items):
            return list(np.asarray(vec(np.stack([np.asarray(it, dtype=np.float64) for it in items]))))
        return [func(it) for it in items]

    def close(self):
        pass

    def join(self):
        pass


def fd_value_and_grad(posterior: VectorizedPosterior, theta, rel_step=None, bounds=None):
    """f(theta) and a forward-difference gradient from ONE batch of P + 1 evaluations.

    Steps follow SciPy's 2-point rule, h_i = sqrt(eps) * max(1, |theta_i|) with the sign of theta_i, flipped
    where the step would leave ``bounds`` (a sequence of (lo, hi))."""
    theta = np.asarray(theta, dtype=np.float64)
    P = len(theta)
    rel = np.sqrt(np.finfo(np.float64).eps) if rel_step is None else rel_step
    h = rel * np.where(theta >= 0, 1.0, -1.0) * np.maximum(1.0, np.abs(theta))
    if bounds is not None:
        for i, (lo, hi) in enumerate(bounds):
            if hi is not None and theta[i] + h[i] > hi:
                h[i] = -abs(h[i])
            if lo is not None and theta[i] + h[i] < lo:
                h[i] = abs(h[i])
    pts = np.repeat(theta[None, :], P + 1, axis=0)
    pts[np.arange(1, P + 1), np.arange(P)] += h
    h = pts[np.arange(1, P + 1), np.arange(P)] - theta            # the step actually taken
    vals = posterior(pts)
    return float(vals[0]), (vals[1:] - vals[0]) / h
