"""gaussn_b200 -- B200-native batched GP log-likelihood engine behind GausSN's API.

Use it the way the reference package is used
(``from gaussn import gausSN, kernels, meanfuncs, lensingmodels``):

    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 30.]), meanfuncs.UniformMean([0.]),
                   lensingmodels.ConstantMagnification([40., 0.5]))

The likelihood runs in hand-written sm_100a CUDA (``libgsn.so``, C ABI in
``include/gsn.h``).  There is no CPU implementation in this package.
"""
from . import _engine, kernels, meanfuncs, lensingmodels, gausSN, samplers, ingest, sharding  # noqa: F401
from .gausSN import GP  # noqa: F401

__all__ = ["gausSN", "kernels", "meanfuncs", "lensingmodels", "samplers", "ingest", "sharding", "GP"]
