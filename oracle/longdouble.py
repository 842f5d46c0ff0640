"""80-bit long-double arbiter for the log-likelihood.  TEST INFRASTRUCTURE ONLY.

When the float64 oracle (LAPACK) and the CUDA engine disagree at the 1e-9..1e-7
level on an ill-conditioned covariance, neither is "the truth": both carry
O(cond * eps) error.  This module evaluates the same formula
(gaussn/gausSN.py:135-160, ExpSquared / Matern32 / OU kernels, UniformMean,
ConstantMagnification or NoLensing) entirely in numpy.longdouble (x87 80-bit on
x86-64: 64-bit mantissa), so tests can check that the engine is at least as close
to the arbiter as the float64 oracle is.
"""
import numpy as np

LD = np.longdouble


def _chol_lower(a):
    """Plain column Cholesky in long double; returns None if a pivot is not positive."""
    a = a.copy()
    n = a.shape[0]
    for j in range(n):
        d = a[j, j] - np.dot(a[j, :j], a[j, :j])
        if not d > 0:
            return None
        d = np.sqrt(d)
        a[j, j] = d
        if j + 1 < n:
            a[j + 1:, j] = (a[j + 1:, j] - a[j + 1:, :j] @ a[j, :j]) / d
    return np.tril(a)


def loglikelihood_ld(x, y, yerr, kernel, kernel_params, mean_c, lens_params=None, n_bands=1, n_images=1, indices=None):
    x = np.asarray(x, dtype=LD); y = np.asarray(y, dtype=LD); yerr = np.asarray(yerr, dtype=LD)
    n = len(x)
    if lens_params is None:
        xs, b = x, np.ones(n, dtype=LD)
        mask = np.ones((n, n), dtype=LD)
    else:
        p = [LD(v) for v in lens_params]
        deltas = np.array([LD(0)] + p[0::2], dtype=LD)
        betas = np.array([LD(1)] + p[1::2], dtype=LD)
        rep = np.diff(np.asarray(indices))
        xs = x - np.repeat(np.tile(deltas, n_bands), rep)
        b = np.repeat(np.tile(betas, n_bands), rep)
        mask = np.zeros((n, n), dtype=LD)
        for pb in range(n_bands):
            s, e = int(indices[n_images * pb]), int(indices[n_images * (pb + 1)])
            mask[s:e, s:e] = 1
    d = xs[:, None] - xs[None, :]
    kp = [LD(v) for v in kernel_params]
    if kernel == "ExpSquaredKernel":
        K = kp[0] ** 2 * np.exp(-d ** 2 / (2 * kp[1] ** 2))
    elif kernel == "Matern32Kernel":
        s = np.sqrt(3 * d ** 2) / kp[1]
        K = kp[0] * ((1 + s) * np.exp(-s))
    elif kernel == "OUKernel":
        K = kp[0] * np.exp(-np.sqrt(d ** 2) / kp[1])
    else:
        raise KeyError(kernel)
    cov = mask * (np.outer(b, b) * K) + np.diag(yerr ** 2)
    L = _chol_lower(cov)
    if L is None:
        return float("nan")
    r = b * LD(mean_c) - y
    z = np.zeros(n, dtype=LD)
    for i in range(n):
        z[i] = (r[i] - np.dot(L[i, :i], z[:i])) / L[i, i]
    a = n * np.log(2 * LD(np.pi)) + 2 * np.sum(np.log(np.diag(L)))
    return float(-(a + np.dot(z, z)) / 2)
