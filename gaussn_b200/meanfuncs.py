"""Mean-function plugins with the reference's class names and signatures.

Drop-in for ``gaussn/meanfuncs.py``: ``params`` lists in the same order,
``_reset(params)``, ``mean(y, params=None, bands=None, zp=None, zpsys=None)``.
The closed-form classes evaluate on the GPU (csrc/gsn_device.cuh ``mean_eval``).
``sncosmoMean`` wraps a third-party host model; ``GP`` routes it through the
engine's host-supplied-mean entry point.
"""
from __future__ import annotations

import numpy as np

from . import _engine


class _Mean:
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

    def mean(self, y, params=None, bands=None, zp=None, zpsys=None):
        if params is not None:
            self._reset(params)
        return _engine.mean(self.kind, [float(v) for v in self.params[:len(self._names)]], y, device=getattr(self, "device", None))


class UniformMean(_Mean):
    """``c``; params ``[c]`` (gaussn/meanfuncs.py:4-45)."""
    kind = _engine.M_UNIFORM
    _names = ("c",)


class Sin(_Mean):
    """``A sin(w t + phi)``; params ``[A, w, phi]`` (gaussn/meanfuncs.py:123-172)."""
    kind = _engine.M_SIN
    _names = ("A", "w", "phi")


class Gaussian(_Mean):
    """``A exp(-(t-mu)^2 / (2 sigma^2)) / (sigma sqrt(2 pi))``; params ``[A, mu, sigma]``
    (gaussn/meanfuncs.py:174-224)."""
    kind = _engine.M_GAUSSIAN
    _names = ("A", "mu", "sigma")


class ExpFunction(_Mean):
    """``A exp(t tau)``; params ``[A, tau]`` (gaussn/meanfuncs.py:226-273)."""
    kind = _engine.M_EXPFUNCTION
    _names = ("A", "tau")


class Bazin2009(_Mean):
    """``A exp(-(t-t0)/Tfall) / (1 + exp((t-t0)/Trise)) + beta``; params
    ``[A, beta, t0, Tfall, Trise]`` (gaussn/meanfuncs.py:275-337)."""
    kind = _engine.M_BAZIN2009
    _names = ("A", "beta", "t0", "Tfall", "Trise")


class Karpenka2012(_Mean):
    """Params ``[lnA, lnB, t1, t0, Tfall, Trise, (offset)]``: the documented 7th
    entry is never read, but ``len(params)`` still counts it
    (gaussn/meanfuncs.py:339-408).  ``A`` and ``B`` hold ``exp`` of the first two."""
    kind = _engine.M_KARPENKA2012
    _names = ("lnA", "lnB", "t1", "t0", "Tfall", "Trise")

    def _reset(self, params):
        super()._reset(params)
        self.A = float(np.exp(params[0]))
        self.B = float(np.exp(params[1]))


class Villar2019(_Mean):
    """Piecewise at ``t1``; params ``[A, beta, t1, t0, Tfall, Trise]``
    (gaussn/meanfuncs.py:410-484)."""
    kind = _engine.M_VILLAR2019
    _names = ("A", "beta", "t1", "t0", "Tfall", "Trise")


class sncosmoMean:
    """Mean function from an ``sncosmo.Model`` (gaussn/meanfuncs.py:47-121).

    sncosmo integrates spectral templates on the host, so this class has no
    device formula: ``GP`` evaluates ``mean`` on the host for every theta at the
    device-computed shifted epochs and hands the values to
    ``gsn_logl_batch_hostmean``.  Any object with ``parameters``, ``param_names``,
    ``update(dict)`` and ``bandflux(bands, t, zp=, zpsys=)`` works as ``model``.
    """
    kind = _engine.M_HOST

    def __init__(self, model, fixed=None):
        self.model = model
        n = len(self.model.parameters)
        if fixed is None:
            self.fixed = np.repeat(False, n)
        elif len(fixed) != n:
            raise IndexError("Array 'fixed' is not the correct length for given model. Please check the length of "
                             "'fixed' is equivalent to the length of model.param_names!")
        else:
            self.fixed = np.array(fixed)
        self.params = np.array(self.model.parameters)[~self.fixed]

    @property
    def n_params(self):
        return len(self.params)

    def _reset(self, params):
        self.params = params
        names = np.array(self.model.param_names)[~self.fixed]
        params_dict = {names[k]: params[k] for k in range(len(self.params))}
        if "x0" in params_dict:
            params_dict["x0"] = params_dict["x0"] * 1.e-8   # meanfuncs.py:78
        self.model.update(params_dict)

    def mean(self, x, params=None, bands=None, zp=27.5, zpsys="ab"):
        if params is not None:
            self._reset(params)
        x = np.asarray(x)
        args = np.argsort(x)
        revert_args = np.zeros(len(args), dtype=int)
        revert_args[args] = np.arange(len(args))

        def reorder(v):
            return np.asarray(v)[args] if (np.ndim(v) > 0 and len(v) >= 2) else v

        flux = self.model.bandflux(reorder(bands), x[args], zp=reorder(zp), zpsys=reorder(zpsys))
        return np.asarray(flux)[revert_args]
