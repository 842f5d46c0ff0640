"""GPU parity of the prediction path (gsn_predict_batch) against the oracle's restatement of GP.predict
(gaussn/gausSN.py:362-403) driven the way gaussn/utils.py:156-209 drives it: per posterior sample, the observations of
one band are de-lensed and the GP is conditioned on them."""
import numpy as np
import pytest

from oracle import gp_oracle
from test_gpu_parity import _synth, _theta_box, _gp_for

pytestmark = pytest.mark.gpu

# Both sides lose digits to the conditioning of K + diag(yerr^2) (the oracle solves by LU as the reference does, the
# device by Cholesky); with the test data's noise floor the moments agree far inside these bounds.
E_TOL = dict(rtol=1e-9, atol=1e-10)
V_TOL = dict(rtol=1e-7, atol=1e-10)


def _oracle_predict_band(ds, gp, theta_row, kernel, mean, lens, x_prime, band_index):
    """utils.py:193-209 for one sample and one band."""
    nk, nm = gp_oracle.KERNEL_NPARAMS[kernel], gp_oracle.MEAN_NPARAMS[mean]
    kp, mp, lp = list(theta_row[:nk]), list(theta_row[nk:nk + nm]), list(theta_row[nk + nm:])
    shifted, b = gp_oracle.lens(lens, lp, ds["x"], gp.n_bands, gp.n_images, gp.indices)
    b = np.broadcast_to(b, shifted.shape)
    if band_index is None:
        r0, r1 = 0, len(ds["x"])
    else:
        r0, r1 = int(gp.indices[band_index * gp.n_images]), int(gp.indices[(band_index + 1) * gp.n_images])
    return gp_oracle.predict(x_prime, shifted[r0:r1], ds["y"][r0:r1] / b[r0:r1], ds["yerr"][r0:r1] / b[r0:r1], kernel, kp, mean, mp)


def _check(gp, ds, theta, kernel, mean, lens, x_prime, band_index, what):
    e, v = gp.predict_batch(theta, x_prime, band_index=band_index)
    assert e.shape == (len(theta), len(x_prime)) and v.shape == (len(theta), len(x_prime), len(x_prime))
    for i, row in enumerate(theta):
        e0, v0 = _oracle_predict_band(ds, gp, row, kernel, mean, lens, x_prime, band_index)
        np.testing.assert_allclose(e[i], e0, err_msg=f"{what}: expectation, theta {i}", **E_TOL)
        np.testing.assert_allclose(v[i], v0, err_msg=f"{what}: variance, theta {i}", **V_TOL)
        assert np.array_equal(v[i], v[i].T), f"{what}: variance is not exactly symmetric"


def _prepared(ds, kernel, mean, lens, n_images, theta_row):
    gp = _gp_for(ds, kernel, mean, lens, n_images, theta_row)
    gp.bands = ds["band"]
    return gp


@pytest.mark.parametrize("n_rows,M", [(1, 1), (5, 3), (31, 32), (32, 33), (33, 31), (64, 64), (70, 9), (100, 200), (257, 100)])
def test_predict_sizes_around_tile_boundaries(libgsn, n_rows, M):
    """Ragged conditioning sets and prediction grids: both are padded to 32 inside the augmented matrix."""
    rng = np.random.default_rng(7000 + 13 * n_rows + M)
    n_images = 2 if n_rows >= 2 else 1
    ds = _synth(rng, n_rows, 1, n_images)
    theta = _theta_box(rng, 6, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", n_images)
    gp = _prepared(ds, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", n_images, theta[0])
    x_prime = np.sort(rng.uniform(-10, 140, M))
    _check(gp, ds, theta, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", x_prime, 0, f"n_rows={n_rows} M={M}")


@pytest.mark.parametrize("kernel", ["ExpSquaredKernel", "ExponentialKernel", "ConstantKernel", "Matern32Kernel", "Matern52Kernel",
                                    "OUKernel", "GibbsKernel", "PeriodicKernel"])
def test_predict_every_kernel_per_band(libgsn, kernel):
    rng = np.random.default_rng(abs(hash("p" + kernel)) % 2**31)
    ds = _synth(rng, 180, 3, 2)
    theta = _theta_box(rng, 5, kernel, "UniformMean", "ConstantMagnification", 2)
    gp = _prepared(ds, kernel, "UniformMean", "ConstantMagnification", 2, theta[0])
    x_prime = np.linspace(-5, 135, 57)
    for band_index in range(3):
        _check(gp, ds, theta, kernel, "UniformMean", "ConstantMagnification", x_prime, band_index, f"{kernel} band {band_index}")


@pytest.mark.parametrize("mean,lens", [("Sin", "NoLensing"), ("Bazin2009", "ConstantMagnification"), ("Gaussian", "SigmoidMagnification"),
                                       ("Villar2019", "ConstantMagnification"), ("Karpenka2012", "SigmoidMagnification"),
                                       ("ExpFunction", "ConstantMagnification")])
def test_predict_means_and_lensing_models(libgsn, mean, lens):
    rng = np.random.default_rng(abs(hash(mean + lens)) % 2**31)
    n_images = 1 if lens == "NoLensing" else 3
    ds = _synth(rng, 120, 2, n_images)
    theta = _theta_box(rng, 5, "Matern32Kernel", mean, lens, n_images)
    gp = _prepared(ds, "Matern32Kernel", mean, lens, n_images, theta[0])
    x_prime = np.sort(rng.uniform(0, 130, 40))
    _check(gp, ds, theta, "Matern32Kernel", mean, lens, x_prime, 1, f"{mean} {lens}")
    _check(gp, ds, theta, "Matern32Kernel", mean, lens, x_prime, None, f"{mean} {lens} all rows")


def test_predict_band_by_label_and_expectation_only(libgsn):
    rng = np.random.default_rng(71)
    ds = _synth(rng, 90, 2, 2)
    theta = _theta_box(rng, 8, "OUKernel", "UniformMean", "ConstantMagnification", 2)
    gp = _prepared(ds, "OUKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    x_prime = np.linspace(0, 120, 45)
    e_idx, v_idx = gp.predict_batch(theta, x_prime, band_index=1)
    e_lab, v_lab = gp.predict_batch(theta, x_prime, band="b1")
    assert np.array_equal(e_idx, e_lab) and np.array_equal(v_idx, v_lab)
    e_only, none = gp.predict_batch(theta, x_prime, band_index=1, return_cov=False)
    assert none is None and np.array_equal(e_only, e_idx)


def test_predict_fixed_groups(libgsn):
    """Posterior samples that only carry the free groups (utils.py:160-190)."""
    rng = np.random.default_rng(72)
    ds = _synth(rng, 80, 1, 2)
    theta = _theta_box(rng, 6, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", 2)
    x_prime = np.linspace(0, 120, 33)
    gp = _prepared(ds, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    e, v = gp.predict_batch(theta[:, 2:], x_prime, band_index=0, fix_kernel_params=True)
    for i, row in enumerate(theta):
        full = np.concatenate([theta[0, :2], row[2:]])
        e0, v0 = _oracle_predict_band(ds, gp, full, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", x_prime, 0)
        np.testing.assert_allclose(e[i], e0, **E_TOL)
        np.testing.assert_allclose(v[i], v0, **V_TOL)


def test_predict_device_tensors_match_host_buffers_bitwise(libgsn):
    import torch
    rng = np.random.default_rng(73)
    ds = _synth(rng, 150, 1, 2)
    theta = _theta_box(rng, 40, "Matern52Kernel", "UniformMean", "ConstantMagnification", 2)
    gp = _prepared(ds, "Matern52Kernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    x_prime = np.linspace(0, 130, 77)
    e_h, v_h = gp.predict_batch(theta, x_prime, band_index=0)
    e_d, v_d = gp.predict_batch(torch.as_tensor(theta, device="cuda"), x_prime, band_index=0)
    assert e_d.is_cuda and v_d.is_cuda
    assert np.array_equal(e_d.cpu().numpy(), e_h) and np.array_equal(v_d.cpu().numpy(), v_h)


def test_predict_pieces_of_a_long_grid_agree_with_one_call(libgsn, monkeypatch):
    """Grids too long for one factorisation are predicted piecewise, pair by pair."""
    from gaussn_b200 import _engine
    rng = np.random.default_rng(74)
    ds = _synth(rng, 60, 1, 2)
    theta = _theta_box(rng, 4, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", 2)
    gp = _prepared(ds, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    x_prime = np.linspace(0, 130, 100)
    e1, v1 = gp.predict_batch(theta, x_prime, band_index=0)
    monkeypatch.setattr(_engine, "MAX_ORDER", 64 + 70)     # room for 70 epochs: pieces of 35
    e2, v2 = gp.predict_batch(theta, x_prime, band_index=0)
    np.testing.assert_allclose(e2, e1, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(v2, v1, rtol=1e-10, atol=1e-13)
    e3, _ = gp.predict_batch(theta, x_prime, band_index=0, return_cov=False)
    np.testing.assert_allclose(e3, e1, rtol=1e-12, atol=1e-13)


def test_predict_not_positive_definite_gives_nan_and_info(libgsn):
    """RationalQuadratic as written (kernels.py:366) grows with distance: not positive definite for most parameters."""
    rng = np.random.default_rng(75)
    ds = _synth(rng, 50, 1, 1)
    gp = _prepared(ds, "RationalQuadraticKernel", "UniformMean", "NoLensing", 1, np.array([1.0, 5.0, 1.0, 0.5]))
    theta = np.array([[1.0, 5.0, 1.0, 0.5], [0.01, 500.0, 2.0, 0.5]])
    gp._ensure_problem(gp.x, gp.y, gp.yerr)
    gp._set_map(False, False, True)
    e, v, info = gp._problem.predict_host(theta, 0, 50, np.linspace(0, 100, 10))
    assert info[0] > 0 and np.all(np.isnan(e[0])) and np.all(np.isnan(v[0]))
    e0, v0 = _oracle_predict_band(ds, gp, theta[1], "RationalQuadraticKernel", "UniformMean", "NoLensing", np.linspace(0, 100, 10), None)
    assert info[1] == 0
    np.testing.assert_allclose(e[1], e0, rtol=1e-7, atol=1e-9)


def test_predict_with_host_supplied_mean(libgsn):
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

    rng = np.random.default_rng(76)
    ds = _synth(rng, 64, 2, 2)
    theta = _theta_box(rng, 5, "ExpSquaredKernel", "Bazin2009", "ConstantMagnification", 2)
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 8.0]), meanfuncs.sncosmoMean(FakeModel()), lensingmodels.ConstantMagnification([5.0, 0.9]))
    gp.x, gp.y, gp.yerr, gp.bands = ds["x"], ds["y"], ds["yerr"], ds["band"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    x_prime = np.linspace(0, 120, 21)
    e, v = gp.predict_batch(theta, x_prime, band_index=1)
    for i, row in enumerate(theta):
        e0, v0 = _oracle_predict_band(ds, gp, row, "ExpSquaredKernel", "Bazin2009", "ConstantMagnification", x_prime, 1)
        np.testing.assert_allclose(e[i], e0, **E_TOL)
        np.testing.assert_allclose(v[i], v0, **V_TOL)


def test_scalar_predict_matches_oracle_on_user_arrays(libgsn):
    """GP.predict(x_prime, x, y, yerr) with arbitrary arrays and the plugins' current parameters (gausSN.py:362)."""
    from gaussn_b200 import gausSN, kernels, meanfuncs
    rng = np.random.default_rng(77)
    x = np.sort(rng.uniform(0, 100, 83))
    y = np.sin(x / 9.0) + 0.05 * rng.standard_normal(83)
    yerr = rng.uniform(0.03, 0.08, 83)
    x_prime = np.linspace(-5, 110, 64)
    for kname, kp in (("Matern32Kernel", [0.8, 14.0]), ("DotProductKernel", []), ("PeriodicKernel", [0.5, 1.1, 31.0])):
        gp = gausSN.GP(getattr(kernels, kname)(*([kp] if kp else [])), meanfuncs.Sin([0.2, 0.11, 0.4]))
        e, v = gp.predict(x_prime, x, y, yerr)
        e0, v0 = gp_oracle.predict(x_prime, x, y, yerr, kname, kp, "Sin", [0.2, 0.11, 0.4])
        # x x'^T reaches 1e4 against a noise floor of 1e-3: condition number ~1e9, so LU (oracle) and Cholesky (device)
        # can only agree to ~1e9 * 2^-53 relative to the matrix scale
        scale = max(1.0, np.abs(v0).max())
        loose = 100.0 if kname == "DotProductKernel" else 1.0
        np.testing.assert_allclose(e, e0, rtol=1e-8, atol=1e-9 * loose, err_msg=kname)
        np.testing.assert_allclose(v, v0, rtol=1e-7, atol=1e-9 * scale, err_msg=kname)


@pytest.mark.parametrize("force", ["cluster2", "cluster4", "cluster8", None])
def test_predict_on_a_cluster_matches_one_cta_bitwise(libgsn, monkeypatch, force):
    """Few parameter sets on a long light curve: one theta per thread-block cluster (as the likelihood does).  Same
    items, same arithmetic as one CTA per theta."""
    from gaussn_b200 import gausSN, kernels, meanfuncs
    rng = np.random.default_rng(81)
    x = np.sort(rng.uniform(0, 900, 600))
    y = 0.4 * np.sin(x / 40.0) + 0.03 * rng.standard_normal(600)
    yerr = rng.uniform(0.02, 0.05, 600)
    x_prime = np.linspace(-10, 920, 300)

    def run():
        gp = gausSN.GP(kernels.Matern32Kernel([0.3, 35.0]), meanfuncs.Sin([0.1, 0.02, 0.3]))
        return gp.predict(x_prime, x, y, yerr)

    monkeypatch.setenv("GSN_FORCE_SHAPE", "wide")
    e1, v1 = run()
    if force:
        monkeypatch.setenv("GSN_FORCE_SHAPE", force)
    else:
        monkeypatch.delenv("GSN_FORCE_SHAPE")      # the shipped dispatch: eight CTAs for this order
    e2, v2 = run()
    assert np.array_equal(e1, e2) and np.array_equal(v1, v2)
    e0, v0 = gp_oracle.predict(x_prime, x, y, yerr, "Matern32Kernel", [0.3, 35.0], "Sin", [0.1, 0.02, 0.3])
    np.testing.assert_allclose(e2, e0, **E_TOL)
    np.testing.assert_allclose(v2, v0, **V_TOL)


def test_scalar_predict_reuses_handles_across_gp_objects_and_data(libgsn):
    """The reference's plotting loop builds a new GP for every posterior sample and calls predict with freshly shifted
    data (utils.py:193-204): handles are kept per length and the observations replaced in place."""
    from gaussn_b200 import gausSN, kernels, meanfuncs, _engine
    rng = np.random.default_rng(91)
    x_prime = np.linspace(0, 100, 40)
    launches = []
    for i in range(6):
        N = (60, 75)[i % 2]                                  # two bands of different length, alternating
        x = np.sort(rng.uniform(0, 100, N)) + 0.3 * i
        y = np.sin(x / 11.0) + 0.05 * rng.standard_normal(N)
        yerr = rng.uniform(0.03, 0.08, N)
        kp, mp = [0.5 + 0.05 * i, 12.0 + i], [0.1 * i]
        gp = gausSN.GP(kernels.ExpSquaredKernel(kp), meanfuncs.UniformMean(mp))
        e, v = gp.predict(x_prime, x, y, yerr)
        e0, v0 = gp_oracle.predict(x_prime, x, y, yerr, "ExpSquaredKernel", kp, "UniformMean", mp)
        np.testing.assert_allclose(e, e0, **E_TOL)
        np.testing.assert_allclose(v, v0, **V_TOL)
        handle = gausSN._PREDICT_HANDLES[(N, _engine.K_EXPSQUARED, _engine.M_UNIFORM, 1, None)][0]
        launches.append(handle.query(_engine.Q_LAUNCHES))
    assert launches == [1, 1, 2, 2, 3, 3]                    # each length kept its handle


def test_predict_batch_against_reference_plotting_loop_vectors(libgsn):
    """gp.predict_batch against outputs of the reference's own code driven as its plotting loop drives it
    (tests/golden/reference_predict_bands.json, made by tests/golden/make_golden_predict.py)."""
    import json
    import os
    from test_gpu_parity import make_gp
    here = os.path.dirname(__file__)
    ref = json.load(open(os.path.join(here, "golden", "reference_predict_bands.json")))
    base = json.load(open(os.path.join(here, "golden", "reference_vectors.json")))
    for case in ref["cases"]:
        ds = base["logl_datasets"][case["dataset"]]
        nk, nm = gp_oracle.KERNEL_NPARAMS[case["kernel"]], gp_oracle.MEAN_NPARAMS[case["mean"]]
        th = np.array(case["sample"])
        gp = make_gp(ds, case["kernel"], th[:nk], case["mean"], th[nk:nk + nm], case["lensing"], th[nk + nm:])
        gp.bands = np.array(ds["band"])
        for bi, want in enumerate(case["bands"]):
            e, v = gp.predict_batch(th[None, :], np.array(case["x_prime"]), band_index=bi)
            np.testing.assert_allclose(e[0], want["expectation"], **E_TOL)
            np.testing.assert_allclose(v[0].ravel(), want["variance"], **V_TOL)


def test_integer_band_labels_are_labels_not_positions(libgsn):
    """Filter ids 1, 2 as band labels: band=2 is the SECOND band (a label), band_index=1 the same by position; a value that
    is not a label raises instead of silently selecting another band's rows."""
    rng = np.random.default_rng(73)
    ds = _synth(rng, 90, 2, 2)
    ds = dict(ds, band=np.where(ds["band"] == "b0", 1, 2))
    theta = _theta_box(rng, 4, "OUKernel", "UniformMean", "ConstantMagnification", 2)
    gp = _prepared(ds, "OUKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    x_prime = np.linspace(0, 120, 20)
    e2, v2 = gp.predict_batch(theta, x_prime, band=2)
    e_pos, v_pos = gp.predict_batch(theta, x_prime, band_index=1)
    assert np.array_equal(e2, e_pos) and np.array_equal(v2, v_pos)
    e1, _ = gp.predict_batch(theta, x_prime, band=1)
    assert not np.array_equal(e1, e2)
    with pytest.raises(ValueError, match="not one of the labels"):
        gp.predict_batch(theta, x_prime, band=0)
    with pytest.raises(ValueError, match="not both"):
        gp.predict_batch(theta, x_prime, band=1, band_index=0)
