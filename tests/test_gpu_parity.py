"""GPU parity: the CUDA engine, called through the reference-shaped API and the
C ABI, against (a) golden vectors produced by the reference's own code and (b) the
numpy/scipy oracle on seeded inputs.  Tolerance: |dlogL| <= max(1e-8, 1e-12 |logL|)."""
import numpy as np
import pytest

from conftest import assert_logl_close
from oracle import gp_oracle

pytestmark = pytest.mark.gpu

KIND = {"ExpSquaredKernel": 0, "ExponentialKernel": 1, "ConstantKernel": 2, "DotProductKernel": 3, "Matern32Kernel": 4,
        "Matern52Kernel": 5, "RationalQuadraticKernel": 6, "GibbsKernel": 7, "OUKernel": 8, "PeriodicKernel": 9, "JJKernel": 10}


def make_gp(ds, kernel, kp, meanf, mp, lens, lp):
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    k = getattr(kernels, kernel)(list(kp))
    m = getattr(meanfuncs, meanf)(list(mp))
    l = lensingmodels.NoLensing() if lens == "NoLensing" else getattr(lensingmodels, lens)(list(lp))
    gp = gausSN.GP(k, m, l)
    gp.x, gp.y, gp.yerr = np.array(ds["x"]), np.array(ds["y"]), np.array(ds["yerr"])
    band = None if ds.get("band") is None else np.array(ds["band"])
    image = None if ds.get("image") is None else np.array(ds["image"])
    gp._prepare_indices(gp.x, band, image)
    if lens != "NoLensing":
        gp.lensingmodel.import_from_gp(gp.n_bands, gp.n_images, gp.indices)
    return gp


def test_kat_against_reference_vectors(libgsn, golden):
    """SURVEY section 4.1 known answers on the shipped data, as computed by the reference's code."""
    got, want = [], []
    for row in golden["kat"]:
        ds = golden["datasets"][row["dataset"]]
        gp = make_gp(ds, row["kernel"], row["kernel_params"], row["mean"], row["mean_params"], row["lensing"], row["lensing_params"])
        assert list(gp.indices) == row["indices"]
        ll = gp.loglikelihood(gp.x, gp.y, gp.yerr, row["kernel_params"], row["mean_params"], row["lensing_params"])
        got.append(ll)
        want.append(row["logl"])
        print(row["dataset"], row["kernel"], row["mean"], f"got {ll!r} want {row['logl']!r} diff {ll - row['logl']:.2e}")
    assert_logl_close(got, want, "KAT")


def test_logl_cases_against_reference_vectors(libgsn, golden):
    """Every plugin combination in the golden file, including non-PD / NaN parameter points."""
    got, want, gotj, wantj = [], [], [], []
    for row in golden["logl"]:
        ds = golden["logl_datasets"][row["dataset"]]
        gp = make_gp(ds, row["kernel"], row["kernel_params"], row["mean"], row["mean_params"], row["lensing"], row["lensing_params"])
        assert list(gp.indices) == row["indices"]
        lp = row["lensing_params"] if row["lensing"] != "NoLensing" else None
        got.append(gp.loglikelihood(gp.x, gp.y, gp.yerr, row["kernel_params"], row["mean_params"], lp))
        want.append(row["logl"])
        theta = row["kernel_params"] + row["mean_params"] + (row["lensing_params"] if lp is not None else [])
        gotj.append(gp.jointprobability(theta, gp.logprior, False, False, lp is None))
        wantj.append(row["jointprob"])
        if not (np.isnan(want[-1]) and np.isnan(got[-1])) and abs(got[-1] - want[-1]) > 1e-9:
            print("LARGE", row["dataset"], row["kernel"], row["mean"], row["lensing"], got[-1], want[-1])
    assert_logl_close(got, want, "logl")
    assert_logl_close(gotj, wantj, "jointprobability")


# --------------------------------------------------------------------------- plugin single calls
def test_plugin_covariance_mean_lens_against_reference_vectors(libgsn, golden):
    from gaussn_b200 import kernels, meanfuncs, lensingmodels
    for row in golden["cov"]:
        k = kernels.DotProductKernel() if row["kernel"] == "DotProductKernel" else getattr(kernels, row["kernel"])(row["params"])
        K = k.covariance(np.array(row["x"]), x_prime=None if row.get("x_prime") is None else np.array(row["x_prime"]))
        assert list(K.shape) == row["shape"]
        # The device multiplies by per-theta reciprocals instead of dividing per element, so an
        # exponent e = ln(K / amplitude) carries ~2 ulp: relative error of the entry ~ (|e| + 3) * 2^-51.
        want = np.array(row["K"])
        nz = want != 0
        depth = np.abs(np.log(np.abs(want[nz]) / np.abs(want).max()))
        rel = np.abs(K.ravel()[nz] - want[nz]) / np.abs(want[nz])
        assert np.all(rel <= (depth + 3.0) * 2.0**-51), (row["kernel"], rel.max())
        assert np.all(np.abs(K.ravel()[~nz]) <= 1e-300)
    for row in golden["mean"]:
        m = getattr(meanfuncs, row["mean"])(row["params"]).mean(np.array(row["t"]))
        np.testing.assert_allclose(m, row["m"], rtol=2e-14, atol=1e-300, err_msg=row["mean"])
    for row in golden["lens"]:
        lm = getattr(lensingmodels, row["lensing"])(row["params"])
        lm.import_from_gp(row["n_bands"], row["n_images"], np.array(row["indices"]))
        xs, b = lm.lens(np.array(row["x"]))
        np.testing.assert_array_equal(xs, row["shifted_x"])
        want_b = np.array(row["b"])
        np.testing.assert_array_equal(np.isnan(b), np.isnan(want_b))
        np.testing.assert_allclose(b[~np.isnan(b)], want_b[~np.isnan(b)], rtol=2e-14, atol=0)
        np.testing.assert_array_equal(lm.mask.ravel(), row["mask"])


def test_theta_slicing_under_fix_flags(libgsn, golden):
    """All eight fix_* combinations and both signs of `invert` (gausSN.py:173-206)."""
    ds = golden["logl_datasets"]["s2x2"]
    for row in golden["slicing"]:
        gp = make_gp(ds, "ExpSquaredKernel", row["kernel0"], "UniformMean", row["mean0"], "ConstantMagnification", row["lens0"])
        got = gp.jointprobability(row["theta"], (lambda v: (lambda p: v))(row["logprior"]), row["fix_kernel"], row["fix_mean"],
                                  row["fix_lens"], row["invert"])
        assert_logl_close([got], [row["value"]], f"slicing {row['fix_kernel']}{row['fix_mean']}{row['fix_lens']}")


def test_predict_against_reference_vectors(libgsn, golden):
    from gaussn_b200 import gausSN, kernels, meanfuncs
    for row in golden["predict"]:
        ds = golden["logl_datasets"][row["dataset"]]
        gp = gausSN.GP(getattr(kernels, row["kernel"])(row["kernel_params"]), getattr(meanfuncs, row["mean"])(row["mean_params"]))
        e, v = gp.predict(np.array(row["x_prime"]), np.array(ds["x"]), np.array(ds["y"]), np.array(ds["yerr"]))
        np.testing.assert_allclose(e, row["expectation"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(v.ravel(), row["variance"], rtol=1e-8, atol=1e-10)


# --------------------------------------------------------------------------- seeded sweeps against the oracle
def _synth(rng, N, n_bands, n_images, span=120.0):
    S = n_bands * n_images
    counts = [N // S] * S
    counts[-1] += N - sum(counts)
    x, band, image = [], [], []
    for b in range(n_bands):
        for i in range(n_images):
            n = counts[b * n_images + i]
            x.append(np.sort(rng.uniform(0, span, n)) + 9.0 * i)
            band += [f"b{b}"] * n
            image += [f"i{i}"] * n
    x = np.concatenate(x) if N else np.zeros(0)
    y = 0.4 + 0.3 * np.sin(x / 15.0) + 0.04 * rng.standard_normal(N)
    yerr = rng.uniform(0.02, 0.06, N)
    return dict(x=x, y=y, yerr=yerr, band=np.array(band), image=np.array(image))


def _theta_box(rng, B, kernel, mean, lens, n_images):
    kbox = {"ExpSquaredKernel": [(0.05, 0.6), (3, 12)], "ExponentialKernel": [(0.05, 0.6), (3, 30)], "ConstantKernel": [(0.01, 0.3)],
            "Matern32Kernel": [(0.01, 0.3), (3, 30)], "Matern52Kernel": [(0.01, 0.3), (3, 20)], "OUKernel": [(0.01, 0.3), (3, 40)],
            "GibbsKernel": [(0.05, 0.5), (5, 15), (0.1, 0.6), (40, 80), (15, 30)], "PeriodicKernel": [(0.01, 0.2), (0.8, 2.0), (20, 50)],
            "RationalQuadraticKernel": [(0.005, 0.02), (400, 600), (1.5, 3.0)], "JJKernel": [(0.005, 0.02), (1.0, 2.0), (1e-5, 1e-4), (25, 40)]}
    mbox = {"UniformMean": [(0.2, 0.7)], "Sin": [(0.1, 0.4), (0.05, 0.08), (-0.5, 0.5)], "Gaussian": [(5, 25), (40, 70), (15, 30)],
            "ExpFunction": [(0.2, 0.6), (-0.004, 0.004)], "Bazin2009": [(0.5, 2), (0, 0.2), (20, 40), (25, 50), (4, 10)],
            "Karpenka2012": [(-1.5, 0), (-9, -7), (20, 50), (20, 40), (25, 50), (4, 10)],
            "Villar2019": [(0.3, 0.8), (-0.004, 0.0), (40, 80), (10, 30), (20, 40), (3, 8)]}
    lbox = {"NoLensing": [], "ConstantMagnification": [(0, 25), (0.4, 1.6)],
            "SigmoidMagnification": [(0, 25), (0.4, 1.2), (-0.3, 0.3), (0.02, 0.3), (30, 90)]}
    cols = [rng.uniform(lo, hi, B) for lo, hi in kbox[kernel] + mbox[mean]]
    for _ in range(n_images - 1):
        cols += [rng.uniform(lo, hi, B) for lo, hi in lbox[lens]]
    return np.column_stack(cols)


def _oracle_batch(ds, theta, kernel, mean, lens, gp):
    return gp_oracle.loglikelihood_batch(theta, ds["x"], ds["y"], ds["yerr"], kernel, mean, lens, gp.n_bands, gp.n_images, gp.indices)


def _gp_for(ds, kernel, mean, lens, n_images, theta_row):
    nk, nm = gp_oracle.KERNEL_NPARAMS[kernel], gp_oracle.MEAN_NPARAMS[mean]
    return make_gp(ds, kernel, theta_row[:nk], mean, theta_row[nk:nk + nm], lens, theta_row[nk + nm:])


@pytest.mark.parametrize("N", [1, 2, 3, 7, 31, 32, 33, 63, 64, 65, 96, 141, 194, 257, 429, 512])
def test_sizes_around_tile_boundaries(libgsn, N):
    """Ragged and tiny inputs: the device pads to a multiple of 32 with identity rows."""
    rng = np.random.default_rng(1000 + N)
    n_images = 2 if N >= 2 else 1
    ds = _synth(rng, N, 1, n_images)
    theta = _theta_box(rng, 24, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", n_images)
    gp = _gp_for(ds, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", n_images, theta[0])
    got = gp.loglikelihood_batch(theta)
    assert_logl_close(got, _oracle_batch(ds, theta, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", gp), f"N={N}")


@pytest.mark.parametrize("kernel", ["ExpSquaredKernel", "ExponentialKernel", "ConstantKernel", "Matern32Kernel", "Matern52Kernel",
                                    "OUKernel", "GibbsKernel", "PeriodicKernel", "RationalQuadraticKernel", "JJKernel"])
def test_every_kernel_multiband_multiimage(libgsn, kernel):
    rng = np.random.default_rng(abs(hash(kernel)) % 2**31)
    ds = _synth(rng, 150, 2, 3)
    theta = _theta_box(rng, 48, kernel, "UniformMean", "ConstantMagnification", 3)
    gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", 3, theta[0])
    want = _oracle_batch(ds, theta, kernel, "UniformMean", "ConstantMagnification", gp)
    assert_logl_close(gp.loglikelihood_batch(theta), want, kernel)
    assert np.isfinite(want).sum() >= 24, f"{kernel}: box too hostile, {np.isfinite(want).sum()} finite"


@pytest.mark.parametrize("mean", ["UniformMean", "Sin", "Gaussian", "ExpFunction", "Bazin2009", "Karpenka2012", "Villar2019"])
@pytest.mark.parametrize("lens", ["NoLensing", "ConstantMagnification", "SigmoidMagnification"])
def test_every_mean_and_lensing_model(libgsn, mean, lens):
    rng = np.random.default_rng(abs(hash(mean + lens)) % 2**31)
    n_images = 1 if lens == "NoLensing" else 2
    ds = _synth(rng, 120, 2, n_images)
    theta = _theta_box(rng, 32, "Matern32Kernel", mean, lens, n_images)
    gp = _gp_for(ds, "Matern32Kernel", mean, lens, n_images, theta[0])
    want = _oracle_batch(ds, theta, "Matern32Kernel", mean, lens, gp)
    assert_logl_close(gp.loglikelihood_batch(theta), want, mean + "/" + lens)


def test_four_images_three_bands_sigmoid_config_c2(libgsn):
    """BASELINE.json config 2 as SURVEY 8d makes it concrete: 3 bands x 4 images, N = 396, Matern-3/2,
    sigmoid microlensing, P = 18."""
    rng = np.random.default_rng(42)
    ds = _synth(rng, 396, 3, 4)
    theta = _theta_box(rng, 64, "Matern32Kernel", "UniformMean", "SigmoidMagnification", 4)
    assert theta.shape[1] == 18
    gp = _gp_for(ds, "Matern32Kernel", "UniformMean", "SigmoidMagnification", 4, theta[0])
    assert_logl_close(gp.loglikelihood_batch(theta), _oracle_batch(ds, theta, "Matern32Kernel", "UniformMean", "SigmoidMagnification", gp), "C2")


def test_failure_values_per_theta_do_not_abort_the_batch(libgsn):
    """Non-PD and NaN parameter points give NaN (loglikelihood) / -inf (jointprobability) for that
    theta only (gausSN.py:203-204); info carries the failing pivot."""
    from gaussn_b200 import _engine
    rng = np.random.default_rng(5)
    ds = _synth(rng, 100, 1, 2)
    theta = _theta_box(rng, 16, "OUKernel", "UniformMean", "ConstantMagnification", 2)
    theta[3, 0] = -0.5            # negative amplitude: not positive definite
    theta[5, 1] = 0.0             # l = 0: NaN covariance
    theta[7, 2] = np.nan          # NaN mean
    theta[9, 4] = np.inf          # infinite magnification
    theta[11, 3] = np.nan         # NaN delay: non-finite shifted epochs
    gp = _gp_for(ds, "OUKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    want = _oracle_batch(ds, theta, "OUKernel", "UniformMean", "ConstantMagnification", gp)
    got = gp.loglikelihood_batch(theta)
    assert_logl_close(got, want, "failures")
    assert set(np.where(np.isnan(got))[0]) == {3, 5, 7, 9, 11}
    out, info = gp._problem.logl_host(theta, _engine.FLAG_NEG_INF, return_info=True)
    assert np.all(np.isneginf(out[[3, 5, 7, 9, 11]])) and np.all(info[[3, 5, 7, 9, 11]] != 0)
    assert np.all(info[np.isfinite(want)] == 0) and info[3] > 0
    jp = gp.jointprobability_batch(theta, logprior=lambda p: 0.0)
    np.testing.assert_array_equal(np.isneginf(jp), np.isnan(want))


def test_sinusoidal_magnification_is_minus_inf_as_in_the_reference(libgsn):
    rng = np.random.default_rng(6)
    ds = _synth(rng, 60, 1, 2)
    gp = make_gp(ds, "ExpSquaredKernel", [0.3, 10.0], "UniformMean", [0.5], "SinusoidalMagnification", [5.0, 0.8, 0.1, 30.0, 9.0])
    assert np.isnan(gp.loglikelihood(gp.x, gp.y, gp.yerr, [0.3, 10.0], [0.5], [5.0, 0.8, 0.1, 30.0, 9.0]))
    assert gp.jointprobability([0.3, 10.0, 0.5, 5.0, 0.8, 0.1, 30.0, 9.0]) == -np.inf


def test_host_buffer_and_device_buffer_paths_are_bit_identical(libgsn):
    import torch
    rng = np.random.default_rng(8)
    ds = _synth(rng, 200, 2, 2)
    theta = _theta_box(rng, 500, "Matern52Kernel", "UniformMean", "ConstantMagnification", 2)
    gp = _gp_for(ds, "Matern52Kernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    a = gp.loglikelihood_batch(theta)
    b = gp.loglikelihood_batch(torch.as_tensor(theta, device="cuda")).cpu().numpy()
    c = gp.loglikelihood_batch(theta[::-1].copy())[::-1]
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(a, c)        # scheduling order does not change any value
    assert gp.loglikelihood_batch(np.zeros((0, 5))).shape == (0,)


def test_no_lensing_multiband_is_dense(libgsn):
    """NoLensing.mask is the scalar 1 (lensingmodels.py:10): bands are NOT decoupled."""
    rng = np.random.default_rng(9)
    ds = _synth(rng, 90, 3, 1)
    theta = _theta_box(rng, 8, "ExpSquaredKernel", "UniformMean", "NoLensing", 1)
    gp = _gp_for(ds, "ExpSquaredKernel", "UniformMean", "NoLensing", 1, theta[0])
    dense = _oracle_batch(ds, theta, "ExpSquaredKernel", "UniformMean", "NoLensing", gp)
    assert_logl_close(gp.loglikelihood_batch(theta), dense, "dense")
    lensed = make_gp(ds, "ExpSquaredKernel", theta[0, :2], "UniformMean", theta[0, 2:3], "ConstantMagnification", [])
    masked = lensed.loglikelihood_batch(theta)
    assert np.max(np.abs(masked - dense)) > 1e-3          # the band mask does change the answer


def test_host_supplied_mean_path(libgsn):
    """sncosmoMean-style host mean (meanfuncs.py:47-121) through gsn_lens_batch + gsn_logl_batch_hostmean,
    with a stand-in model whose bandflux is Bazin2009 so that the oracle can check it."""
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels

    class FakeModel:
        param_names = ["A", "beta", "t0", "Tfall", "Trise"]

        def __init__(self):
            self.parameters = np.array([1.0, 0.1, 30.0, 40.0, 8.0])

        def update(self, d):
            for k, v in d.items():
                self.parameters[self.param_names.index(k)] = v

        def bandflux(self, bands, t, zp=None, zpsys=None):
            return gp_oracle.mean("Bazin2009", list(self.parameters), t)

    rng = np.random.default_rng(10)
    ds = _synth(rng, 80, 2, 2)
    theta = _theta_box(rng, 12, "ExpSquaredKernel", "Bazin2009", "ConstantMagnification", 2)
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 8.0]), meanfuncs.sncosmoMean(FakeModel()), lensingmodels.ConstantMagnification([5.0, 0.9]))
    gp.x, gp.y, gp.yerr, gp.bands = ds["x"], ds["y"], ds["yerr"], ds["band"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    got = gp.loglikelihood_batch(theta)
    want = gp_oracle.loglikelihood_batch(theta, ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", "Bazin2009", "ConstantMagnification",
                                         gp.n_bands, gp.n_images, gp.indices)
    assert_logl_close(got, want, "host mean")


def test_minimize_reaches_the_notebook_optimum_region(libgsn, golden):
    """Soft end-to-end: optimize_parameters(method='minimize') on simSN_107 from the
    SLSN-notebook start climbs above the notebook's loglstar = 368.647."""
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    ds = golden["datasets"]["simSN_107"]
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 28.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([38.0, 0.5]))
    res = gp.optimize_parameters(np.array(ds["x"]), np.array(ds["y"]), np.array(ds["yerr"]), band=np.array(ds["band"]),
                                 image=np.array(ds["image"]), method="minimize", fix_mean_params=True,
                                 minimize_kwargs=dict(method="Nelder-Mead", options=dict(xatol=1e-4, fatol=1e-6, maxiter=600)))
    assert -res.fun >= 368.647, res


@pytest.mark.parametrize("n_bands,per_band", [(4, 64), (3, 32), (2, 96), (5, 40), (4, 96), (2, 256), (3, 130)])
def test_band_block_diagonal_structure(libgsn, n_bands, per_band):
    """With a lensing model the covariance is block diagonal by band (lensingmodels.py:39-51); the engine skips
    the zero blocks, including the case where band boundaries coincide with its 32-row chunks (then diagonal
    blocks of different bands are independent and may be factored concurrently), in both launch shapes."""
    rng = np.random.default_rng(77 + n_bands)
    ds = _synth(rng, n_bands * per_band, n_bands, 2)
    theta = _theta_box(rng, 40, "Matern52Kernel", "UniformMean", "ConstantMagnification", 2)
    gp = _gp_for(ds, "Matern52Kernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    want = _oracle_batch(ds, theta, "Matern52Kernel", "UniformMean", "ConstantMagnification", gp)
    assert_logl_close(gp.loglikelihood_batch(theta), want, f"{n_bands} bands x {per_band}")
    # the same data without a lensing model couples the bands (mask = 1) and must differ
    theta_nl = theta[:, :3]
    gp_nl = _gp_for(ds, "Matern52Kernel", "UniformMean", "NoLensing", 1, theta_nl[0])
    want_nl = gp_oracle.loglikelihood_batch(theta_nl, ds["x"], ds["y"], ds["yerr"], "Matern52Kernel", "UniformMean", "NoLensing",
                                            gp_nl.n_bands, gp_nl.n_images, gp_nl.indices)
    assert_logl_close(gp_nl.loglikelihood_batch(theta_nl), want_nl, "dense")


def test_sampler_batch_adapters(libgsn, golden):
    """Vectorised posterior (emcee / zeus vectorize=True), pool-style queue evaluation (dynesty) and the
    batched finite-difference gradient agree with the scalar jointprobability the reference hands to samplers."""
    from gaussn_b200 import samplers
    ds = golden["logl_datasets"]["s2x2"]
    gp = make_gp(ds, "ExpSquaredKernel", [0.3, 12.0], "UniformMean", [0.5], "ConstantMagnification", [11.0, 0.7])
    prior = lambda p: -0.5 * float(np.sum(np.asarray(p) ** 2)) * 1e-3 if p[1] > 0 else -np.inf
    rng = np.random.default_rng(3)
    pts = np.column_stack([rng.uniform(0.1, 0.6, 20), rng.uniform(-2, 20, 20), rng.uniform(0.3, 0.7, 20), rng.uniform(0, 20, 20),
                           rng.uniform(0.4, 1.2, 20)])
    scalar = np.array([gp.jointprobability(list(p), prior) for p in pts])
    post = samplers.VectorizedPosterior(gp, prior)
    np.testing.assert_array_equal(post(pts), scalar)
    assert post(pts[0]) == scalar[0] and np.isneginf(scalar).any()
    pool = samplers.BatchPool(size=64)
    calls0 = post.n_calls
    np.testing.assert_array_equal(np.array(pool.map(post, list(pts))), scalar)
    assert post.n_calls == calls0 + 1                      # one launch for the whole queue
    assert pool.map(lambda v: v + 1, [1, 2]) == [2, 3]      # anything else is mapped serially
    neg = samplers.VectorizedPosterior(gp, prior, invert=-1)
    th = pts[np.isfinite(scalar)][0]
    f, g = samplers.fd_value_and_grad(neg, th)
    assert f == -gp.jointprobability(list(th), prior)
    h = 1e-6
    for i in range(len(th)):
        e = np.zeros_like(th); e[i] = h
        central = (neg(th + e) - neg(th - e)) / (2 * h)
        assert abs(g[i] - central) <= 1e-4 * max(1.0, abs(central)), (i, g[i], central)


def test_minimize_with_batched_gradient(libgsn, golden):
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    ds = golden["datasets"]["simSN_107"]
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 28.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([38.0, 0.5]))
    res = gp.optimize_parameters(np.array(ds["x"]), np.array(ds["y"]), np.array(ds["yerr"]), band=np.array(ds["band"]),
                                 image=np.array(ds["image"]), method="minimize", fix_mean_params=True,
                                 minimize_kwargs=dict(method="L-BFGS-B", batched_gradient=True,
                                                      bounds=[(0.01, 1.0), (10.0, 40.0), (30.0, 50.0), (0.1, 1.0)]))
    assert -res.fun >= 368.647, res       # the notebook's loglstar for this fit (ipynb/fitting_SLSN_example.ipynb:153)


def test_hypothesis_random_layouts_and_theta(libgsn):
    """Property sweep: random segment layouts, sizes and parameter points, random plugin triple."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    kernels_ok = ["ExpSquaredKernel", "ExponentialKernel", "Matern32Kernel", "Matern52Kernel", "OUKernel", "PeriodicKernel", "GibbsKernel"]
    means_ok = ["UniformMean", "Sin", "Gaussian", "ExpFunction", "Bazin2009", "Karpenka2012", "Villar2019"]

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(seed=st.integers(0, 2**31 - 1), n_bands=st.integers(1, 4), n_images=st.integers(1, 4), per_seg=st.integers(1, 40),
           kernel=st.sampled_from(kernels_ok), mean=st.sampled_from(means_ok), lens=st.sampled_from(["ConstantMagnification", "SigmoidMagnification", "NoLensing"]))
    def run(seed, n_bands, n_images, per_seg, kernel, mean, lens):
        rng = np.random.default_rng(seed)
        if lens == "NoLensing":
            n_images = 1
        N = n_bands * n_images * per_seg + int(rng.integers(0, 5))
        ds = _synth(rng, N, n_bands, n_images)
        theta = _theta_box(rng, 6, kernel, mean, lens, n_images)
        gp = _gp_for(ds, kernel, mean, lens, n_images, theta[0])
        want = _oracle_batch(ds, theta, kernel, mean, lens, gp)
        assert_logl_close(gp.loglikelihood_batch(theta), want, f"{kernel}/{mean}/{lens} N={N} bands={n_bands} images={n_images} seed={seed}")

    run()


@pytest.mark.parametrize("kernel,n_bands,n_images,N", [("ExpSquaredKernel", 1, 2, 141), ("Matern32Kernel", 4, 2, 194), ("OUKernel", 1, 3, 33),
                                                       ("GibbsKernel", 2, 2, 200), ("JJKernel", 1, 2, 96), ("Matern52Kernel", 3, 2, 336),
                                                       ("ExpSquaredKernel", 4, 2, 512), ("Matern32Kernel", 3, 4, 396), ("OUKernel", 6, 2, 900)])
def test_one_warp_per_theta_shape_on_large_batches(libgsn, monkeypatch, kernel, n_bands, n_images, N):
    """Workspace kernels (band blocks above 208 rows; here forced for the smaller ones as well): batches of 512+ theta with
    band blocks of at most 208 rows run the narrow launch shape (one warp per theta) once the on-chip shape is switched off; smaller batches of the same problem run
    the four-warp shape.  Both must agree with the oracle and with each other bit for bit."""
    from gaussn_b200 import _engine
    monkeypatch.setenv("GSN_FORCE_SHAPE", "legacy")
    rng = np.random.default_rng(N * 7 + n_bands)
    ds = _synth(rng, N, n_bands, n_images)
    theta = _theta_box(rng, 640, kernel, "UniformMean", "ConstantMagnification", n_images)
    theta[17, 0] = -abs(theta[17, 0])          # a non-PD point inside the batch
    gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta[0])
    big = gp.loglikelihood_batch(theta)
    assert gp._problem.query(_engine.Q_PATH) == 2
    small = np.concatenate([gp.loglikelihood_batch(theta[i:i + 160]) for i in range(0, 640, 160)])
    assert gp._problem.query(_engine.Q_PATH) == 1
    np.testing.assert_array_equal(big, small)
    want = _oracle_batch(ds, theta[:96], kernel, "UniformMean", "ConstantMagnification", gp)
    assert_logl_close(big[:96], want, f"narrow {kernel} N={N}")


def test_host_supplied_mean_with_fixed_groups(libgsn):
    """The host-mean path under fix_* flags: fixed mean parameters come from the plugin, fixed kernel / lensing
    parameters from the handle's theta map."""
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels

    class FakeModel:
        param_names = ["A", "beta", "t0", "Tfall", "Trise"]

        def __init__(self):
            self.parameters = np.array([1.0, 0.1, 30.0, 40.0, 8.0])

        def update(self, d):
            for k, v in d.items():
                self.parameters[self.param_names.index(k)] = v

        def bandflux(self, bands, t, zp=None, zpsys=None):
            return gp_oracle.mean("Bazin2009", list(self.parameters), t)

    rng = np.random.default_rng(21)
    ds = _synth(rng, 70, 1, 2)
    theta = _theta_box(rng, 9, "ExpSquaredKernel", "Bazin2009", "ConstantMagnification", 2)
    k0, m0, l0 = [0.3, 8.0], [1.2, 0.05, 33.0, 41.0, 7.0], [5.0, 0.9]

    def fresh():
        gp = gausSN.GP(kernels.ExpSquaredKernel(list(k0)), meanfuncs.sncosmoMean(FakeModel()), lensingmodels.ConstantMagnification(list(l0)))
        gp.meanfunc._reset(list(m0))
        gp.x, gp.y, gp.yerr, gp.bands = ds["x"], ds["y"], ds["yerr"], ds["band"]
        gp._prepare_indices(ds["x"], ds["band"], ds["image"])
        return gp

    def oracle(kp, mp, lp):
        return gp_oracle.loglikelihood(ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", kp, "Bazin2009", mp, "ConstantMagnification", lp,
                                       1, 2, fresh().indices)

    # kernel fixed: theta = mean | lensing
    got = fresh().loglikelihood_batch(theta[:, 2:], fix_kernel_params=True)
    assert_logl_close(got, [oracle(k0, list(t[2:7]), list(t[7:])) for t in theta], "host mean, kernel fixed")
    # mean fixed: theta = kernel | lensing
    th = np.column_stack([theta[:, :2], theta[:, 7:]])
    got = fresh().loglikelihood_batch(th, fix_mean_params=True)
    assert_logl_close(got, [oracle(list(t[:2]), m0, list(t[7:])) for t in theta], "host mean, mean fixed")
    # lensing fixed: theta = kernel | mean
    got = fresh().loglikelihood_batch(theta[:, :7], fix_lensing_params=True)
    assert_logl_close(got, [oracle(list(t[:2]), list(t[2:7]), l0) for t in theta], "host mean, lensing fixed")


@pytest.mark.parametrize("kernel,n_bands,n_images,N,force", [("Matern32Kernel", 1, 2, 512, None), ("ExpSquaredKernel", 2, 2, 440, None),
                                                             ("OUKernel", 1, 4, 1100, None), ("PeriodicKernel", 1, 2, 420, None),
                                                             ("Matern32Kernel", 1, 2, 330, None),
                                                             ("Matern52Kernel", 3, 2, 130, "cluster4"), ("ExpSquaredKernel", 1, 2, 70, "cluster8"),
                                                             ("OUKernel", 2, 2, 200, "cluster2")])
def test_cluster_shape_for_few_thetas(libgsn, monkeypatch, kernel, n_bands, n_images, N, force):
    """Calls with a few thetas at N >= 256 spread each theta over a thread-block cluster of 2, 4 or 8 CTAs (control words
    in the rank-0 CTA's shared memory, reached through DSMEM).  Same items, same arithmetic: bit-identical to the one-CTA
    shape, which larger batches of the same problem use."""
    from gaussn_b200 import _engine
    if force:
        monkeypatch.setenv("GSN_FORCE_SHAPE", force)
    rng = np.random.default_rng(N * 3 + n_bands)
    ds = _synth(rng, N, n_bands, n_images, span=4.0 * N / (n_bands * n_images))
    theta = _theta_box(rng, 200, kernel, "UniformMean", "ConstantMagnification", n_images)
    if kernel != "ExpSquaredKernel":
        theta[3, 0] = -abs(theta[3, 0])          # a non-PD point among the few
    gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta[0])
    tail = gp.loglikelihood_batch(theta[8:24])
    mid = gp.loglikelihood_batch(theta[1:8])
    assert gp._problem.query(_engine.Q_PATH) == 3
    few = np.concatenate([gp.loglikelihood_batch(theta[:1]), mid, tail])
    assert gp._problem.query(_engine.Q_PATH) == 3
    monkeypatch.setenv("GSN_FORCE_SHAPE", "wide")
    gp2 = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta[0])
    many = gp2.loglikelihood_batch(theta)
    assert gp2._problem.query(_engine.Q_PATH) == 1
    np.testing.assert_array_equal(few, many[:24])
    if kernel != "ExpSquaredKernel":
        assert np.isnan(few[3])
    want = _oracle_batch(ds, theta[:12], kernel, "UniformMean", "ConstantMagnification", gp)
    assert_logl_close(few[:12], want, f"cluster {kernel} N={N}", rtol=1e-11 if N >= 1024 else 1e-12)
    # scalar entry point, as the reference's samplers call it
    nk = gp_oracle.KERNEL_NPARAMS[kernel]
    monkeypatch.delenv("GSN_FORCE_SHAPE")
    if force:
        monkeypatch.setenv("GSN_FORCE_SHAPE", force)
    got = gp.jointprobability(list(theta[5]))
    assert got == many[5] or (np.isnan(many[5]) and got == -np.inf)


def test_new_observations_of_the_same_shape_reuse_the_handle(libgsn):
    from gaussn_b200 import _engine
    rng = np.random.default_rng(55)
    ds = _synth(rng, 90, 2, 2)
    theta = _theta_box(rng, 12, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2)
    gp = _gp_for(ds, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    gp.loglikelihood_batch(theta)
    handle = gp._problem
    y2 = ds["y"] * 1.3 + 0.05
    e2 = ds["yerr"] * 1.3
    got = gp.loglikelihood_batch(theta, y=y2, yerr=e2)
    assert gp._problem is handle and handle.query(_engine.Q_LAUNCHES) == 2
    want = gp_oracle.loglikelihood_batch(theta, ds["x"], y2, e2, "Matern32Kernel", "UniformMean", "ConstantMagnification",
                                         gp.n_bands, gp.n_images, gp.indices)
    assert_logl_close(got, want, "data replaced in place")


def test_handles_on_different_threads_run_concurrently(libgsn):
    """include/gsn.h: a handle is not thread-safe, but different handles may be driven from different threads at the
    same time (each host-buffer call runs on its handle's own stream).  Results must not depend on what else runs."""
    import threading
    rng = np.random.default_rng(66)
    jobs = []
    for kernel, n_bands, n_images, N, B in (("ExpSquaredKernel", 1, 2, 141, 700), ("Matern32Kernel", 4, 2, 194, 64),
                                            ("OUKernel", 1, 3, 429, 5), ("Matern52Kernel", 2, 2, 520, 150)):
        ds = _synth(rng, N, n_bands, n_images, span=4.0 * N / (n_bands * n_images))
        theta = _theta_box(rng, B, kernel, "UniformMean", "ConstantMagnification", n_images)
        gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta[0])
        jobs.append((gp, theta, gp.loglikelihood_batch(theta)))      # serial reference (also creates the handle)
    errors = []

    def worker(gp, theta, want):
        try:
            for _ in range(25):
                got = gp.loglikelihood_batch(theta)
                if not np.array_equal(got, want):
                    errors.append("mismatch")
                    return
        except Exception as e:   # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=job) for job in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_workspace_grows_with_the_largest_batch(libgsn):
    """A handle used one theta at a time (or only for prediction) holds a workspace for the CTAs it launched, not for a
    full batch; a larger batch grows it, results are unaffected."""
    from gaussn_b200 import _engine
    rng = np.random.default_rng(77)
    ds = _synth(rng, 300, 1, 2)
    theta = _theta_box(rng, 900, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2)
    gp = _gp_for(ds, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    one = gp.loglikelihood_batch(theta[:1])
    w1 = gp._problem.query(_engine.Q_WORKSPACE_BYTES)
    mid = gp.loglikelihood_batch(theta[:200])
    w2 = gp._problem.query(_engine.Q_WORKSPACE_BYTES)
    big = gp.loglikelihood_batch(theta)
    w3 = gp._problem.query(_engine.Q_WORKSPACE_BYTES)
    assert 0 < w1 < w2 < w3
    again = gp.loglikelihood_batch(theta[:1])
    assert gp._problem.query(_engine.Q_WORKSPACE_BYTES) == w3
    assert one[0] == big[0] == again[0] and np.array_equal(mid, big[:200])


def test_minimize_finds_the_reference_optimum(libgsn, golden):
    """optimize_parameters(method='minimize') with the reference's own result as the answer: the optimum that the
    reference's unmodified optimize_parameters finds on the same data with the same Nelder-Mead settings
    (tests/golden/make_golden_minimize.py).  Pins the whole drop-in call: index bookkeeping, theta packing, invert = -1."""
    import json
    import os
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_minimize.json")))
    for case in ref["cases"]:
        ds = golden["logl_datasets"][case["dataset"]]
        gp = gausSN.GP(kernels.ExpSquaredKernel(list(case["kernel0"])), meanfuncs.UniformMean(list(case["mean0"])),
                       lensingmodels.ConstantMagnification(list(case["lens0"])))
        res = gp.optimize_parameters(np.array(ds["x"]), np.array(ds["y"]), np.array(ds["yerr"]), band=np.array(ds["band"]),
                                     image=np.array(ds["image"]), method="minimize", fix_mean_params=case["fix_mean"],
                                     minimize_kwargs=case["minimize_kwargs"])
        assert res.success
        # the simplex walks on function values that agree to ~1e-10, so the two runs may part ways late: same optimum,
        # not the same iterates
        assert abs(res.fun - case["fun"]) <= 1e-8, (res.fun, case["fun"])
        np.testing.assert_allclose(res.x, case["x"], rtol=1e-5, atol=1e-5)


def test_mean_and_cov_of_the_last_loglikelihood_call(libgsn):
    """The reference leaves self.mean and self.cov behind (gausSN.py:142, :147); here they are produced on demand."""
    rng = np.random.default_rng(31)
    ds = _synth(rng, 60, 2, 2)
    kp, mp, lp = [0.3, 9.0], [0.8, 0.05, 30.0, 40.0, 8.0], [7.0, 0.8]
    gp = make_gp(ds, "Matern32Kernel", kp, "Bazin2009", mp, "ConstantMagnification", lp)
    with pytest.raises(AttributeError):
        gp.mean
    ll = gp.loglikelihood(ds["x"], ds["y"], ds["yerr"], kp, mp, lp)
    want_mean, want_cov = gp_oracle.build_mean_cov(ds["x"], ds["yerr"], "Matern32Kernel", kp, "Bazin2009", mp, "ConstantMagnification", lp,
                                                   gp.n_bands, gp.n_images, gp.indices)
    np.testing.assert_allclose(gp.mean, want_mean, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(gp.cov, want_cov, rtol=1e-13, atol=1e-300)
    # and they are what the likelihood was computed from
    import scipy.linalg
    L = np.linalg.cholesky(gp.cov)
    z = scipy.linalg.solve_triangular(L, gp.mean - ds["y"], lower=True)
    assert abs(-0.5 * (len(z) * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(L))) + z @ z) - ll) <= 1e-8
