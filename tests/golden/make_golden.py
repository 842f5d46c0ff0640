#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.json by running the REFERENCE's own code.

Run in the build container only (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/make_golden.py

The reference's modules (``/root/reference/gaussn/{gausSN,kernels,meanfuncs,
lensingmodels}.py``) are imported unmodified.  Their ``import jax`` resolves to
``oracle/jax_shim`` because real JAX is not installable here; see that package's
docstring for what the shim does and does not reproduce.  Everything written to
the JSON file is an output of reference code on the inputs stored beside it.

Sections of the file
  datasets   prepared inputs (x, y, yerr, band, image) from the shipped data files,
             prepared the way the notebooks do (sort by band, image, time; divide
             flux and error by max(flux)-min(flux)): ipynb/fitting_SLSN_example.ipynb
             cell 3, ipynb/fitting_glQSO_example.ipynb cell 5.
  kat        GP.loglikelihood on those datasets at the SURVEY section 4.1 parameter points.
  cov/mean/lens/mask   per-plugin outputs on small inputs.
  logl       GP.loglikelihood / GP.jointprobability for plugin combinations on
             synthetic multi-band, multi-image data, including non-PD and NaN cases.
  indices    GP._prepare_indices cases.
  slicing    jointprobability under all fix_* flag combinations.
  predict    GP.predict.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REPO, "oracle", "jax_shim"))
sys.path.insert(0, REF)

import pandas as pd  # noqa: E402
from gaussn import gausSN, kernels, meanfuncs, lensingmodels  # noqa: E402  (the reference)


def L(a):
    return [float(v) for v in np.asarray(a, dtype=np.float64).ravel()]


# ------------------------------------------------------------------ datasets
def load_sim(name):
    df = pd.read_csv(f"{REF}/data/glSN_forGP_sims/{name}.csv").sort_values(["band", "image", "time"], kind="stable")
    factor = df.flux.max() - df.flux.min()
    return dict(x=L(df.time), y=L(df.flux / factor), yerr=L(df.fluxerr / factor),
                band=[str(b) for b in df.band], image=[str(i) for i in df["image"]], rescale=float(factor))


def load_quasar(only_euler):
    raw = pd.read_csv(f"{REF}/data/COSMOGRAIL/DES_J0408-5354.txt", sep="\t")
    rows = []
    for t, tag in enumerate(["A", "B", "D"]):
        flux = 10 ** ((raw[f"mag_{tag}"].values - 27.5) / -2.5)
        err = flux * raw[f"magerr_{tag}"].values * np.log(10) / 2.5
        rows.append(pd.DataFrame(dict(time=raw["mhjd"].values, band=raw["telescope"].values, flux=flux, fluxerr=err,
                                      image=f"image_{t + 1}")))
    df = pd.concat(rows).sort_values(["band", "image", "time"], kind="stable")
    factor = df.flux.max() - df.flux.min()
    df["flux"] /= factor
    df["fluxerr"] /= factor
    if only_euler:
        df = df[df.band != "WFI"]
    return dict(x=L(df.time), y=L(df.flux), yerr=L(df.fluxerr), band=[str(b) for b in df.band],
                image=[str(i) for i in df["image"]], rescale=float(factor))


def make_gp(kernel, meanfunc, lensing, ds):
    """Bring a reference GP to the state optimize_parameters leaves it in
    before sampling (gausSN.py:276-291), without starting a sampler."""
    gp = gausSN.GP(kernel, meanfunc, lensing)
    gp.x = np.array(ds["x"])
    gp.y = np.array(ds["y"])
    gp.yerr = np.array(ds["yerr"])
    band = None if ds.get("band") is None else np.array(ds["band"])
    image = None if ds.get("image") is None else np.array(ds["image"])
    gp.bands, gp.zp, gp.zpsys = band, 27.5, "ab"
    gp._prepare_indices(gp.x, band, image)
    fix_lens = False
    try:
        gp.lensingmodel.import_from_gp(gp.n_bands, gp.n_images, gp.indices)
    except Exception:
        fix_lens = True
    return gp, fix_lens


KCLS = {n: getattr(kernels, n) for n in
        ["ExpSquaredKernel", "ExponentialKernel", "ConstantKernel", "Matern32Kernel", "Matern52Kernel",
         "RationalQuadraticKernel", "GibbsKernel", "OUKernel", "PeriodicKernel", "JJKernel"]}
MCLS = {n: getattr(meanfuncs, n) for n in
        ["UniformMean", "Sin", "Gaussian", "ExpFunction", "Bazin2009", "Karpenka2012", "Villar2019"]}
LCLS = {n: getattr(lensingmodels, n) for n in
        ["ConstantMagnification", "SigmoidMagnification", "SinusoidalMagnification"]}


def ref_logl(ds, kname, kp, mname, mp, lname, lp):
    lens = lensingmodels.NoLensing() if lname == "NoLensing" else LCLS[lname](list(lp))
    gp, fix_lens = make_gp(KCLS[kname](list(kp)), MCLS[mname](list(mp)), lens, ds)
    ll = gp.loglikelihood(gp.x, gp.y, gp.yerr, list(kp), list(mp), None if lname == "NoLensing" else list(lp))
    theta = list(kp) + list(mp) + ([] if lname == "NoLensing" else list(lp))
    jp = gp.jointprobability(theta, gp.logprior, False, False, fix_lens)
    return float(ll), float(jp), [int(i) for i in gp.indices]


def synth(seed, n_bands, n_images, per_seg, span=100.0, jitter=True):
    rng = np.random.default_rng(seed)
    x, band, image = [], [], []
    for b in range(n_bands):
        for i in range(n_images):
            n = per_seg + (int(rng.integers(0, 4)) if jitter else 0)
            x.append(np.sort(rng.uniform(0, span, n)) + 12.0 * i)
            band += [f"b{b}"] * n
            image += [f"image_{i + 1}"] * n
    x = np.concatenate(x)
    y = 0.5 + 0.3 * np.sin(x / 17.0) + 0.05 * rng.standard_normal(len(x))
    yerr = rng.uniform(0.01, 0.05, len(x))
    return dict(x=L(x), y=L(y), yerr=L(yerr), band=band if n_bands > 1 or True else None, image=image)


def main():
    out = {"_about": "outputs of /root/reference/gaussn/*.py executed on numpy float64 through oracle/jax_shim; "
                     "generated by tests/golden/make_golden.py", "numpy": np.__version__}

    # ---- datasets + KATs
    ds = {"simSN_107": load_sim("simSN_107"), "J0408_euler": load_quasar(True), "J0408_all": load_quasar(False)}
    out["datasets"] = ds
    lp_q = [116.9, 1.78, 151.2, 0.632]
    kat_rows = [
        ("simSN_107", "ExpSquaredKernel", [0.31, 30.0], "UniformMean", [0.0], "ConstantMagnification", [39.4, 0.5]),
        ("simSN_107", "ExpSquaredKernel", [0.5, 20.0], "UniformMean", [0.0], "ConstantMagnification", [0.0, 1.0]),
        ("simSN_107", "Matern32Kernel", [0.1, 30.0], "UniformMean", [0.01], "ConstantMagnification", [39.4, 0.5]),
        ("simSN_107", "OUKernel", [0.1, 30.0], "UniformMean", [0.0], "ConstantMagnification", [40.0, 0.5]),
        ("J0408_euler", "OUKernel", [0.0095, 330.0], "UniformMean", [0.69], "ConstantMagnification", lp_q),
        ("J0408_euler", "Matern32Kernel", [0.01, 200.0], "UniformMean", [0.69], "ConstantMagnification", lp_q),
        ("J0408_euler", "ExpSquaredKernel", [0.1, 50.0], "UniformMean", [0.69], "ConstantMagnification", lp_q),
        ("J0408_all", "OUKernel", [0.0095, 330.0], "UniformMean", [0.69], "ConstantMagnification", lp_q),
        ("J0408_all", "Matern32Kernel", [0.01, 200.0], "UniformMean", [0.69], "ConstantMagnification", lp_q),
        ("J0408_all", "ExpSquaredKernel", [0.1, 50.0], "UniformMean", [0.69], "ConstantMagnification", lp_q),
        # a fitted non-trivial mean on the SN data (SLSN-notebook style)
        ("simSN_107", "ExpSquaredKernel", [0.2, 25.0], "Bazin2009", [1.2, 0.02, 20.0, 30.0, 6.0], "ConstantMagnification", [39.4, 0.5]),
        ("simSN_107", "Matern52Kernel", [0.05, 25.0], "Karpenka2012", [-0.5, -6.0, 15.0, 20.0, 30.0, 6.0], "SigmoidMagnification", [39.4, 0.5, 0.05, 0.1, 60.0]),
    ]
    out["kat"] = []
    for name, k, kp, m, mp, l, lp in kat_rows:
        ll, jp, idx = ref_logl(ds[name], k, kp, m, mp, l, lp)
        out["kat"].append(dict(dataset=name, kernel=k, kernel_params=kp, mean=m, mean_params=mp, lensing=l,
                               lensing_params=lp, indices=idx, logl=ll, jointprob=jp))

    # ---- per-plugin vectors
    xs = np.array([0.0, 0.7, 1.9, 3.2, 7.5, 11.0, 11.0, 23.4])
    xq = np.array([-1.0, 0.7, 5.5])
    kparams = {
        "ExpSquaredKernel": [[0.8, 3.0], [1.3, 0.4]], "ExponentialKernel": [[0.8, 3.0]], "ConstantKernel": [[0.37]],
        "Matern32Kernel": [[0.8, 3.0], [2.0, 11.0]], "Matern52Kernel": [[0.8, 3.0]],
        "RationalQuadraticKernel": [[0.8, 3.0, 1.7]], "GibbsKernel": [[0.9, 4.0, 0.6, 5.0, 2.5]],
        "OUKernel": [[0.8, 3.0]], "PeriodicKernel": [[0.8, 1.4, 6.0]], "JJKernel": [[0.8, 1.4, 0.3, 5.0]],
    }
    out["cov"] = []
    for k, plist in kparams.items():
        for p in plist:
            kern = KCLS[k](list(p))
            out["cov"].append(dict(kernel=k, params=p, x=L(xs), K=L(kern._covariance(xs)), shape=[len(xs), len(xs)]))
    # the only kernel whose x_prime branch works on arrays as written (`is None`, kernels.py:53)
    kern = kernels.ExpSquaredKernel([0.8, 3.0])
    out["cov"].append(dict(kernel="ExpSquaredKernel", params=[0.8, 3.0], x=L(xq), x_prime=L(xs),
                           K=L(kern._covariance(xq, x_prime=xs)), shape=[len(xq), len(xs)]))
    dp = kernels.DotProductKernel()
    out["cov"].append(dict(kernel="DotProductKernel", params=[], x=L(xs), K=L(dp._covariance(xs)), shape=[len(xs), len(xs)]))

    mparams = {
        "UniformMean": [[0.4]], "Sin": [[1.2, 0.35, 0.8]], "Gaussian": [[2.0, 6.0, 3.5]], "ExpFunction": [[0.7, -0.08]],
        "Bazin2009": [[1.5, 0.1, 4.0, 12.0, 2.0]], "Karpenka2012": [[0.2, -3.0, 5.0, 4.0, 12.0, 2.0, 99.0]],
        "Villar2019": [[1.1, -0.02, 9.0, 2.0, 15.0, 1.5]],
    }
    out["mean"] = []
    for m, plist in mparams.items():
        for p in plist:
            mf = MCLS[m](list(p))
            out["mean"].append(dict(mean=m, params=p, t=L(xs), m=L(mf.mean(xs, params=list(p)))))

    out["lens"] = []
    lens_ds = synth(5, 2, 3, 3, jitter=True)
    lparams = {
        "ConstantMagnification": [[10.0, 0.5, 25.0, 1.7]],
        "SigmoidMagnification": [[10.0, 0.5, 0.3, 0.2, 40.0, 25.0, 1.7, -0.4, 0.05, 60.0]],
        "SinusoidalMagnification": [[10.0, 0.5, 0.3, 40.0, 9.0, 25.0, 1.7, -0.4, 60.0, 21.0]],
    }
    for l, plist in lparams.items():
        for p in plist:
            gp, _ = make_gp(kernels.ConstantKernel([1.0]), meanfuncs.UniformMean([0.0]), LCLS[l](list(p)), lens_ds)
            sx, b = gp.lensingmodel._lens(gp.x, params=list(p))
            out["lens"].append(dict(lensing=l, params=p, x=lens_ds["x"], band=lens_ds["band"], image=lens_ds["image"],
                                    indices=[int(i) for i in gp.indices], n_bands=int(gp.n_bands), n_images=int(gp.n_images),
                                    shifted_x=L(sx), b=L(b), mask=L(gp.lensingmodel.mask)))

    # ---- logL over plugin combinations on synthetic data
    out["logl_datasets"] = {
        "s1x2": synth(11, 1, 2, 16), "s2x2": synth(12, 2, 2, 9), "s3x4": synth(13, 3, 4, 5), "s1x1": synth(14, 1, 1, 21),
        "s2x3_tiny": synth(15, 2, 3, 1, jitter=False),
    }
    lens_p = {
        ("ConstantMagnification", 2): [11.0, 0.7], ("ConstantMagnification", 3): [11.0, 0.7, 25.0, 1.4],
        ("ConstantMagnification", 4): [11.0, 0.7, 25.0, 1.4, 33.0, 0.45],
        ("SigmoidMagnification", 2): [11.0, 0.7, 0.3, 0.2, 40.0],
        ("SigmoidMagnification", 3): [11.0, 0.7, 0.3, 0.2, 40.0, 25.0, 1.4, -0.2, 0.05, 70.0],
        ("SigmoidMagnification", 4): [11.0, 0.7, 0.3, 0.2, 40.0, 25.0, 1.4, -0.2, 0.05, 70.0, 33.0, 0.45, 0.1, 1.5, 55.0],
        ("SinusoidalMagnification", 2): [11.0, 0.7, 0.3, 40.0, 9.0],
    }
    cases = []
    kern_cases = {
        "ExpSquaredKernel": [0.3, 12.0], "ExponentialKernel": [0.3, 12.0], "ConstantKernel": [0.2],
        "Matern32Kernel": [0.1, 12.0], "Matern52Kernel": [0.1, 12.0], "RationalQuadraticKernel": [0.05, 200.0, 2.0],
        "GibbsKernel": [0.3, 15.0, 0.5, 50.0, 20.0], "OUKernel": [0.1, 12.0], "PeriodicKernel": [0.1, 1.5, 30.0],
        "JJKernel": [0.1, 1.5, 0.2, 30.0],
    }
    mean_cases = {
        "UniformMean": [0.5], "Sin": [0.3, 1 / 17.0, 0.1], "Gaussian": [20.0, 50.0, 25.0], "ExpFunction": [0.4, 0.003],
        "Bazin2009": [1.5, 0.1, 30.0, 40.0, 8.0], "Karpenka2012": [-0.7, -8.0, 35.0, 30.0, 40.0, 8.0],
        "Villar2019": [0.6, -0.002, 60.0, 20.0, 30.0, 5.0],
    }
    for dname, n_img in (("s1x2", 2), ("s2x2", 2), ("s3x4", 4)):
        for k, kp in kern_cases.items():
            cases.append((dname, k, kp, "UniformMean", [0.5], "ConstantMagnification", lens_p[("ConstantMagnification", n_img)]))
        for m, mp in mean_cases.items():
            cases.append((dname, "Matern32Kernel", [0.1, 12.0], m, mp, "SigmoidMagnification", lens_p[("SigmoidMagnification", n_img)]))
    for k, kp in kern_cases.items():
        cases.append(("s1x1", k, kp, "Sin", [0.3, 1 / 17.0, 0.1], "NoLensing", []))
        cases.append(("s2x2", k, kp, "UniformMean", [0.5], "NoLensing", []))   # dense cross-band covariance
    # parameter points where the two kernels that are usually non-PD as written do factorise
    for dname, n_img in (("s1x2", 2), ("s3x4", 4)):
        cases.append((dname, "JJKernel", [0.01, 1.5, 1e-4, 30.0], "UniformMean", [0.5], "ConstantMagnification", lens_p[("ConstantMagnification", n_img)]))
        cases.append((dname, "RationalQuadraticKernel", [0.02, 500.0, 2.0], "UniformMean", [0.5], "ConstantMagnification", lens_p[("ConstantMagnification", n_img)]))
    cases.append(("s2x3_tiny", "ExpSquaredKernel", [0.3, 12.0], "UniformMean", [0.5], "ConstantMagnification", lens_p[("ConstantMagnification", 3)]))
    cases.append(("s1x2", "ExpSquaredKernel", [0.3, 12.0], "UniformMean", [0.5], "SinusoidalMagnification", lens_p[("SinusoidalMagnification", 2)]))
    # failure values: tau = 0 (NaN), negative amplitude (non-PD), huge length scale on tiny errors, beta = 0
    cases.append(("s1x2", "ExpSquaredKernel", [0.3, 0.0], "UniformMean", [0.5], "ConstantMagnification", [11.0, 0.7]))
    cases.append(("s1x2", "OUKernel", [-0.5, 12.0], "UniformMean", [0.5], "ConstantMagnification", [11.0, 0.7]))
    cases.append(("s1x2", "Matern32Kernel", [-0.5, 12.0], "UniformMean", [0.5], "ConstantMagnification", [11.0, 0.7]))
    cases.append(("s1x2", "RationalQuadraticKernel", [0.8, 3.0, 1.7], "UniformMean", [0.5], "ConstantMagnification", [11.0, 0.7]))
    cases.append(("s1x2", "ExpSquaredKernel", [0.3, 12.0], "UniformMean", [0.5], "ConstantMagnification", [11.0, 0.0]))
    cases.append(("s1x2", "ExpSquaredKernel", [0.3, 12.0], "UniformMean", [float("nan")], "ConstantMagnification", [11.0, 0.7]))
    cases.append(("s1x2", "ExpSquaredKernel", [0.3, 12.0], "Bazin2009", [1.5, 0.1, 30.0, 0.0, 8.0], "ConstantMagnification", [11.0, 0.7]))
    out["logl"] = []
    for dname, k, kp, m, mp, l, lp in cases:
        ll, jp, idx = ref_logl(out["logl_datasets"][dname], k, kp, m, mp, l, lp)
        out["logl"].append(dict(dataset=dname, kernel=k, kernel_params=kp, mean=m, mean_params=mp, lensing=l,
                                lensing_params=lp, indices=idx, logl=ll, jointprob=jp))

    # ---- _prepare_indices
    out["indices"] = []
    band = np.array(["g", "g", "g", "r", "r", "z", "z", "z", "z"])
    image = np.array(["a", "a", "b", "a", "b", "a", "a", "b", "b"])
    for bd, im in ((band, image), (band, None), (None, image), (None, None)):
        gp = gausSN.GP(kernels.ConstantKernel([1.0]), meanfuncs.UniformMean([0.0]))
        gp._prepare_indices(np.arange(9.0), bd, im)
        out["indices"].append(dict(band=None if bd is None else list(bd), image=None if im is None else list(im), n=9,
                                   indices=[int(i) for i in gp.indices], n_bands=int(gp.n_bands), n_images=int(gp.n_images),
                                   factor=float(gp.factor)))

    # ---- theta slicing under fix_* flags (gausSN.py:176-194)
    out["slicing"] = []
    dsl = out["logl_datasets"]["s2x2"]
    k0, m0, l0 = [0.2, 9.0], [0.45], [9.0, 0.9]
    k1, m1, l1 = [0.3, 12.0], [0.5], [11.0, 0.7]
    for fk in (False, True):
        for fm in (False, True):
            for fl in (False, True):
                gp, _ = make_gp(kernels.ExpSquaredKernel(list(k0)), meanfuncs.UniformMean(list(m0)),
                                lensingmodels.ConstantMagnification(list(l0)), dsl)
                theta = ([] if fk else k1) + ([] if fm else m1) + ([] if fl else l1)
                for invert in (1, -1):
                    gp.kernel._reset(list(k0)); gp.meanfunc._reset(list(m0)); gp.lensingmodel._reset(list(l0))
                    v = gp.jointprobability(list(theta), lambda p: -0.125, fk, fm, fl, invert)
                    out["slicing"].append(dict(fix_kernel=fk, fix_mean=fm, fix_lens=fl, invert=invert, theta=theta,
                                               kernel0=k0, mean0=m0, lens0=l0, logprior=-0.125, value=float(v)))
    gp, _ = make_gp(kernels.ExpSquaredKernel(list(k0)), meanfuncs.UniformMean(list(m0)),
                    lensingmodels.ConstantMagnification(list(l0)), dsl)
    for invert in (1, -1):
        v = gp.jointprobability(k1 + m1 + l1, lambda p: -np.inf, False, False, False, invert)
        out["slicing"].append(dict(fix_kernel=False, fix_mean=False, fix_lens=False, invert=invert, theta=k1 + m1 + l1,
                                   kernel0=k0, mean0=m0, lens0=l0, logprior=float("-inf"), value=float(v)))

    # ---- predict (gausSN.py:362-403)
    dp1 = out["logl_datasets"]["s1x1"]
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 12.0]), meanfuncs.Sin([0.3, 1 / 17.0, 0.1]))
    xp = np.linspace(-5.0, 110.0, 9)
    e, v = gp.predict(xp, np.array(dp1["x"]), np.array(dp1["y"]), np.array(dp1["yerr"]))
    out["predict"] = [dict(dataset="s1x1", kernel="ExpSquaredKernel", kernel_params=[0.3, 12.0], mean="Sin",
                           mean_params=[0.3, 1 / 17.0, 0.1], x_prime=L(xp), expectation=L(e), variance=L(v))]

    path = os.path.join(HERE, "reference_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f)
    n_nan = sum(1 for c in out["logl"] if not np.isfinite(c["logl"]))
    print(f"wrote {path}: {len(out['kat'])} KATs, {len(out['cov'])} cov, {len(out['mean'])} mean, {len(out['lens'])} lens, "
          f"{len(out['logl'])} logl ({n_nan} non-finite), {len(out['slicing'])} slicing; {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
