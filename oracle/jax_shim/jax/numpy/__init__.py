"""``jax.numpy`` stand-in: numpy float64, plus JAX's Cholesky conventions."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import inf, pi  # noqa: F401


class _Linalg:
    @staticmethod
    def cholesky(a, *, upper=False, symmetrize_input=True):
        # jax.numpy.linalg.cholesky: symmetrize_input=True by default, lower
        # factor, and a matrix of NaN (no exception) when potrf fails.
        a = _np.asarray(a, dtype=_np.float64)
        if symmetrize_input:
            a = (a + _np.swapaxes(a, -1, -2)) / 2
        if not _np.all(_np.isfinite(a)):
            return _np.full_like(a, _np.nan)
        try:
            L = _np.linalg.cholesky(a)
        except _np.linalg.LinAlgError:
            return _np.full_like(a, _np.nan)
        return _np.swapaxes(L, -1, -2) if upper else L

    def __getattr__(self, name):
        return getattr(_np.linalg, name)


linalg = _Linalg()
