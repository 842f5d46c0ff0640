#!/usr/bin/env python
"""Generate tests/golden/reference_minimize.json: the REFERENCE's own optimize_parameters(method='minimize')
(gaussn/gausSN.py:208-336, unmodified, ``import jax`` -> oracle/jax_shim) on a small two-image data set.  Pins the whole
drop-in call path -- _prepare_indices, import_from_gp, the theta packing of _get_initial_pos / jointprobability, the
invert = -1 convention -- against the optimum the reference itself finds.  Run in the build container only:

    python tests/golden/make_golden_minimize.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "oracle", "jax_shim"))
sys.path.insert(0, "/root/reference")

from gaussn import gausSN, kernels, meanfuncs, lensingmodels  # noqa: E402  (the reference)


def main():
    base = json.load(open(os.path.join(HERE, "reference_vectors.json")))
    cases = []
    # all groups free: with fix_mean_params != fix_kernel_params the reference's _get_initial_pos (defined with its first
    # two arguments swapped, gausSN.py:103 vs :329) packs the wrong groups and jointprobability raises IndexError
    for dname, fix_mean in (("s1x2", False), ("s2x2", False)):
        ds = base["logl_datasets"][dname]
        k0, m0, l0 = [0.35, 9.0], [0.45], [6.0, 0.9]
        gp = gausSN.GP(kernels.ExpSquaredKernel(list(k0)), meanfuncs.UniformMean(list(m0)), lensingmodels.ConstantMagnification(list(l0)))
        opts = dict(method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-12, maxiter=6000, maxfev=12000))
        res = gp.optimize_parameters(np.array(ds["x"]), np.array(ds["y"]), np.array(ds["yerr"]), band=np.array(ds["band"]),
                                     image=np.array(ds["image"]), method="minimize", fix_mean_params=fix_mean, minimize_kwargs=opts)
        cases.append(dict(dataset=dname, kernel0=k0, mean0=m0, lens0=l0, fix_mean=fix_mean, minimize_kwargs=opts,
                          x=[float(v) for v in res.x], fun=float(res.fun), nfev=int(res.nfev), success=bool(res.success)))
        print(dname, res.x, res.fun, res.nfev, res.success)
    path = os.path.join(HERE, "reference_minimize.json")
    json.dump(dict(_about="optimum found by the reference's optimize_parameters(method='minimize'); see make_golden_minimize.py",
                   cases=cases), open(path, "w"))
    print("wrote", path)


if __name__ == "__main__":
    main()
