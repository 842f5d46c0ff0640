#!/usr/bin/env python
"""Generate tests/golden/reference_predict_bands.json by running the REFERENCE's own code.

The body of the reference's plotting loop (gaussn/utils.py:193-204), executed with the reference's unmodified classes
(``import jax`` resolves to oracle/jax_shim, as in make_golden.py): per posterior sample the plugins are reset, a fresh GP
is built, and for every band the band's observations are lensed with the sample's parameters (``lensingmodel.repeats`` /
``n_bands`` patched exactly as utils.py does), divided by the magnifications and handed to ``GP.predict``.  (utils.py
itself imports matplotlib and dynesty and cannot be imported here; the loop body is transcribed below, its calls are the
reference's.)  Run in the build container only:

    python tests/golden/make_golden_predict.py
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
import jax.numpy as jnp  # noqa: E402  (the shim)


def L(a):
    return [float(v) for v in np.asarray(a, dtype=np.float64).ravel()]


class JaxLike(np.ndarray):
    """numpy array that answers ``arr == None`` the way a jax.Array does (False, not an elementwise array): every kernel
    but ExpSquaredKernel tests its x_prime argument with ``== None`` (gaussn/kernels.py:106, 154, ... 552), which only
    works on JAX arrays."""

    def __eq__(self, other):
        if other is None:
            return False
        return np.ndarray.__eq__(self, other)

    __hash__ = None


def plotting_loop_body(ds, kernel, meanfunc, lensingmodel, sample, predict_times):
    """utils.py:159-204 for one sample (all groups free), returning (expectation, covariance) per band."""
    nk, nm = len(kernel.params), len(meanfunc.params)
    kernel._reset([sample[i] for i in range(nk)])
    meanfunc._reset([sample[i + nk] for i in range(nm)])
    lensingmodel._reset([sample[i + nk + nm] for i in range(len(lensingmodel.params))])
    gp = gausSN.GP(kernel, meanfunc, lensingmodel)
    band, image = np.array(ds["band"]), np.array(ds["image"])
    x, y, yerr = np.array(ds["x"]), np.array(ds["y"]), np.array(ds["yerr"])
    out = []
    for pb_id in np.unique(band):
        sel = band == pb_id
        repeats = np.array([int(np.sum(image[sel] == im_id)) for im_id in np.unique(image)])     # utils.py:200
        lensingmodel.repeats = repeats
        lensingmodel.n_bands = 1
        shifted_time_data, b_vector = lensingmodel._lens(jnp.array(x[sel]))                        # utils.py:203
        exp, cov = gp.predict(np.asarray(predict_times).view(JaxLike), np.asarray(shifted_time_data).view(JaxLike),
                              y[sel] / b_vector, yerr[sel] / b_vector, band=[pb_id])
        out.append((exp, cov))
    return out


def main():
    base = json.load(open(os.path.join(HERE, "reference_vectors.json")))
    cases = []
    rng = np.random.default_rng(20240601)
    specs = [
        ("s2x2", "ExpSquaredKernel", [(0.2, 0.5), (8, 14)], "UniformMean", [(0.3, 0.6)], "ConstantMagnification", 2,
         [(2, 12), (0.6, 1.3)]),
        ("s3x4", "Matern32Kernel", [(0.05, 0.3), (8, 25)], "Sin", [(0.1, 0.3), (0.05, 0.08), (-0.5, 0.5)], "ConstantMagnification", 4,
         [(2, 12), (0.6, 1.3)]),
        ("s3x4", "OUKernel", [(0.05, 0.3), (8, 30)], "UniformMean", [(0.3, 0.6)], "SigmoidMagnification", 4,
         [(2, 12), (0.5, 1.0), (-0.2, 0.2), (0.05, 0.2), (30, 70)]),
    ]
    for dname, kname, kbox, mname, mbox, lname, n_images, lbox in specs:
        ds = base["logl_datasets"][dname]
        for _ in range(3):
            kp = [rng.uniform(*b) for b in kbox]
            mp = [rng.uniform(*b) for b in mbox]
            lp = [rng.uniform(*b) for _i in range(n_images - 1) for b in lbox]
            kernel = getattr(kernels, kname)(list(kp))
            meanfunc = getattr(meanfuncs, mname)(list(mp))
            lensingmodel = getattr(lensingmodels, lname)(list(lp))
            # what optimize_parameters does before sampling (gausSN.py:286-288)
            gp0 = gausSN.GP(kernel, meanfunc, lensingmodel)
            gp0._prepare_indices(np.array(ds["x"]), np.array(ds["band"]), np.array(ds["image"]))
            lensingmodel.import_from_gp(gp0.n_bands, gp0.n_images, gp0.indices)
            predict_times = np.linspace(-3.0, 125.0, 17)
            per_band = plotting_loop_body(ds, kernel, meanfunc, lensingmodel, kp + mp + lp, predict_times)
            cases.append(dict(dataset=dname, kernel=kname, mean=mname, lensing=lname, n_images=n_images, sample=L(kp + mp + lp),
                              x_prime=L(predict_times), bands=[dict(expectation=L(e), variance=L(v)) for e, v in per_band]))
    path = os.path.join(HERE, "reference_predict_bands.json")
    json.dump(dict(_about="outputs of the reference's GP.predict driven as gaussn/utils.py:193-204 drives it; see make_golden_predict.py",
                   cases=cases), open(path, "w"))
    print(f"wrote {path}: {len(cases)} samples x bands, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
