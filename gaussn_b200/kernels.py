"""Covariance-function plugins with the reference's class names and signatures.

Drop-in for ``gaussn/kernels.py``: same constructors (``params`` list in the same
order), same ``.params`` / ``_reset(params)`` / ``covariance(x, x_prime=None,
params=None)`` protocol and the same named attributes.  Each class carries the
``kind`` enum of include/gsn.h, which is how ``GP`` lowers the plugin to the CUDA
engine; ``covariance`` itself evaluates on the GPU through ``gsn_covariance``.
The arithmetic lives in csrc/gsn_device.cuh.
"""
from __future__ import annotations

from . import _engine


class _Kernel:
    kind = -1
    _names = ()

    def __init__(self, params):
        self._reset(params)

    @property
    def n_params(self):
        return len(self._names)

    def _reset(self, params):
        params = list(params)
        if len(params) < len(self._names):
            raise IndexError(f"{type(self).__name__} needs {len(self._names)} parameters {list(self._names)}")
        for name, value in zip(self._names, params):
            setattr(self, name, value)
        self.params = params

    def covariance(self, x, x_prime=None, params=None):
        if params is not None:
            self._reset(params)
        return _engine.covariance(self.kind, [float(v) for v in self.params[:len(self._names)]], x, x_prime, device=getattr(self, "device", None))

    _covariance = covariance   # the reference exposes both names (kernels.py:20, :30)


class ExpSquaredKernel(_Kernel):
    """``A**2 * exp(-(x - x')**2 / (2 tau**2))``, params ``[A, tau]`` (gaussn/kernels.py:4-56)."""
    kind = _engine.K_EXPSQUARED
    _names = ("A", "tau")


class ExponentialKernel(_Kernel):
    """``A**2 * exp(-|x - x'| / tau)``, params ``[A, tau]`` (gaussn/kernels.py:58-111)."""
    kind = _engine.K_EXPONENTIAL
    _names = ("A", "tau")


class ConstantKernel(_Kernel):
    """``c``, params ``[c]`` (gaussn/kernels.py:113-158)."""
    kind = _engine.K_CONSTANT
    _names = ("c",)


class DotProductKernel(_Kernel):
    """``x_i * x'_j``; takes no parameters and has no ``params`` attribute, so it
    cannot be fitted (gaussn/kernels.py:160-197)."""
    kind = _engine.K_DOTPRODUCT
    _names = ()

    def __init__(self):
        self.c = 0

    def _reset(self):
        self.c = 0

    def covariance(self, x, x_prime=None, params=None):
        return _engine.covariance(self.kind, [], x, x_prime, device=getattr(self, "device", None))

    _covariance = covariance


class Matern32Kernel(_Kernel):
    """``A (1 + sqrt(3 d^2)/l) exp(-sqrt(3 d^2)/l)`` -- A is not squared; params ``[A, l]``
    (gaussn/kernels.py:199-252)."""
    kind = _engine.K_MATERN32
    _names = ("A", "l")


class Matern52Kernel(_Kernel):
    """``A (1 + sqrt(5 d^2)/l + 5 d^2/(3 l^2)) exp(-sqrt(5 d^2)/l)``; params ``[A, l]``
    (gaussn/kernels.py:254-309)."""
    kind = _engine.K_MATERN52
    _names = ("A", "l")


class RationalQuadraticKernel(_Kernel):
    """As the reference writes it: ``A**2 (1 + d^2 / (2 alpha tau^2))`` with no ``**(-alpha)``;
    params ``[A, tau, scale_mixture]`` (gaussn/kernels.py:311-367)."""
    kind = _engine.K_RATQUAD
    _names = ("A", "tau", "scale_mixture")


class GibbsKernel(_Kernel):
    """Gibbs kernel with ``tau(t) = lambda (1 - p N(t | mu, sigma))``; params
    ``[A, lambda, p, mu, sigma]``; the attribute is spelt ``lamdba`` as in the
    reference (gaussn/kernels.py:369-448)."""
    kind = _engine.K_GIBBS
    _names = ("A", "lamdba", "p", "mu", "sigma")


class OUKernel(_Kernel):
    """``A exp(-|d| / l)``; params ``[A, l]`` (gaussn/kernels.py:450-500)."""
    kind = _engine.K_OU
    _names = ("A", "l")


class PeriodicKernel(_Kernel):
    """``A exp(-2 (sin(pi |d| / p) / l)^2)``; params ``[A, l, p]`` (gaussn/kernels.py:502-527)."""
    kind = _engine.K_PERIODIC
    _names = ("A", "l", "p")


class JJKernel(_Kernel):
    """Periodic kernel whose period ``m x + b`` is taken from the column epoch
    (gaussn/kernels.py:529-559); ``covariance`` returns it as written (asymmetric),
    the likelihood symmetrises it as ``jnp.linalg.cholesky`` does."""
    kind = _engine.K_JJ
    _names = ("A", "l", "m", "b")

    def _p(self, x):
        return self.m * x + self.b


BY_KIND = {cls.kind: cls for cls in (ExpSquaredKernel, ExponentialKernel, ConstantKernel, DotProductKernel, Matern32Kernel,
                                     Matern52Kernel, RationalQuadraticKernel, GibbsKernel, OUKernel, PeriodicKernel, JJKernel)}
