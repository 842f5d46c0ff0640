"""``jax.scipy.linalg`` stand-in (only ``solve_triangular`` is used)."""
import numpy as _np
from scipy.linalg import solve_triangular as _st


def solve_triangular(a, b, trans=0, lower=False, unit_diagonal=False, **_):
    a = _np.asarray(a, dtype=_np.float64)
    b = _np.asarray(b, dtype=_np.float64)
    if not _np.all(_np.isfinite(a)):
        return _np.full_like(b, _np.nan)
    return _st(a, b, trans=trans, lower=lower, unit_diagonal=unit_diagonal, check_finite=False)
