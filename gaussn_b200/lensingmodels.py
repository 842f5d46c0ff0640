"""Lensing-model plugins with the reference's class names and signatures.

Drop-in for ``gaussn/lensingmodels.py``: ``params`` interleaved per extra image,
``_reset``, ``lens(x, params=None) -> (shifted_x, b)``, ``import_from_gp(n_bands,
n_images, indices)``, ``mask``, ``repeats``.  ``lens`` evaluates on the GPU
(csrc/gsn_device.cuh ``lens_eval``); the likelihood never materialises ``mask`` --
the kernels test ``band[i] == band[j]`` instead -- so the attribute is built
lazily for callers that read it.
"""
from __future__ import annotations

import numpy as np

from . import _engine


class NoLensing:
    """Identity: ``lens`` returns ``(x, 1)`` and ``mask`` is the scalar 1, so
    bands are NOT decoupled (gaussn/lensingmodels.py:5-14)."""
    kind = _engine.L_NONE
    per_image = 0

    def __init__(self, params=None):
        self.mask = 1

    def lens(self, x, params=None):
        return x, 1

    _lens = lens


class _Magnification:
    kind = -1
    per_image = 0
    _fields = ()      # (attribute name, value for the reference image)

    def __init__(self, params):
        self._reset(params)
        self._mask = None

    def _reset(self, params):
        params = list(params)
        if self.per_image and len(params) % self.per_image:
            raise IndexError(f"{type(self).__name__} takes {self.per_image} parameters per extra image")
        for k, (name, first) in enumerate(self._fields):
            setattr(self, name, np.array([first] + params[k::self.per_image], dtype=float))
        self.params = params

    def import_from_gp(self, n_bands, n_images, indices):
        """gaussn/lensingmodels.py:102-118."""
        self.n_bands = n_bands
        self.n_images = n_images
        self.indices = np.asarray(indices)
        self.repeats = self.indices[1:] - self.indices[:-1]
        self._mask = None

    def _make_mask(self):
        """1 iff both rows are in the same band (gaussn/lensingmodels.py:39-51)."""
        n = int(self.indices[-1])
        mask = np.zeros((n, n))
        for pb in range(self.n_bands):
            start = int(self.indices[self.n_images * pb])
            stop = int(self.indices[self.n_images * (pb + 1)])
            mask[start:stop, start:stop] = 1
        return mask

    @property
    def mask(self):
        if self._mask is None:
            self._mask = self._make_mask()
        return self._mask

    def lens(self, x, params=None):
        if params is not None:
            self._reset(params)
        if not hasattr(self, "indices"):
            raise AttributeError("call import_from_gp(n_bands, n_images, indices) first")
        if len(self.params) != self.per_image * (self.n_images - 1):
            raise IndexError(f"{type(self).__name__}: {self.n_images} images need "
                             f"{self.per_image * (self.n_images - 1)} parameters, got {len(self.params)}")
        return _engine.lens(self.kind, [float(v) for v in self.params], x, self.n_bands, self.n_images, self.indices, device=getattr(self, "device", None))

    _lens = lens


class ConstantMagnification(_Magnification):
    """``[delta, beta]`` per extra image; ``x~ = x - delta``, ``b = beta``
    (gaussn/lensingmodels.py:16-118)."""
    kind = _engine.L_CONSTANT
    per_image = 2
    _fields = (("deltas", 0.0), ("betas", 1.0))


class SigmoidMagnification(_Magnification):
    """``[delta, beta0, beta1, r, t0]`` per extra image;
    ``b = beta0 + beta1 / (1 + exp(-r (x~ - t0)))`` at the shifted epoch
    (gaussn/lensingmodels.py:120-241)."""
    kind = _engine.L_SIGMOID
    per_image = 5
    _fields = (("deltas", 0.0), ("beta0s", 1.0), ("beta1s", 0.0), ("rs", 0.0), ("t0s", 0.0))


class SinusoidalMagnification(_Magnification):
    """``[delta, beta0, beta1, t0, T]`` per extra image;
    ``b = beta0 + beta1 sin((x~ - t0)/T)``.  The reference image has ``T = 0``
    (gaussn/lensingmodels.py:259-263), so its ``b`` is NaN and every likelihood is
    -inf, exactly as in the reference."""
    kind = _engine.L_SINUSOIDAL
    per_image = 5
    _fields = (("deltas", 0.0), ("beta0s", 1.0), ("beta1s", 0.0), ("t0s", 0.0), ("Ts", 0.0))
