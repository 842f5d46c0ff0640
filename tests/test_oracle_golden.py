"""CPU: the numpy/scipy oracle against golden vectors produced by the reference's own code
(tests/golden/make_golden.py) -- this is what pins the oracle.  Also the long-double arbiter."""
import numpy as np
import pytest

from conftest import assert_logl_close
from oracle import gp_oracle as o
from oracle import longdouble as ld


def _idx(ds):
    return o.prepare_indices(ds["x"], None if ds.get("band") is None else np.array(ds["band"]),
                             None if ds.get("image") is None else np.array(ds["image"]))


def test_kat_rows_match_reference(golden):
    got, want = [], []
    for row in golden["kat"]:
        ds = golden["datasets"][row["dataset"]]
        idx, nb, ni, _ = _idx(ds)
        assert list(idx) == row["indices"]
        got.append(o.loglikelihood(ds["x"], ds["y"], ds["yerr"], row["kernel"], row["kernel_params"], row["mean"],
                                   row["mean_params"], row["lensing"], row["lensing_params"], nb, ni, idx))
        want.append(row["logl"])
    assert_logl_close(got, want, "oracle KAT")


def test_kat_matches_survey_table(golden):
    """the notebook-anchored known answers."""
    survey = [368.662295773483, -1805.710045550164, 349.551587105639, 251.168492377418, 339.064098386586,
              285.748723741819, 286.035489323340, 1104.629740022919, 810.833904200799, 413.645455597859]
    got = [row["logl"] for row in golden["kat"][:10]]
    np.testing.assert_allclose(got, survey, rtol=0, atol=5e-10)


def test_logl_cases_match_reference(golden):
    got, want, gotj, wantj = [], [], [], []
    for row in golden["logl"]:
        ds = golden["logl_datasets"][row["dataset"]]
        idx, nb, ni, _ = _idx(ds)
        assert list(idx) == row["indices"]
        got.append(o.loglikelihood(ds["x"], ds["y"], ds["yerr"], row["kernel"], row["kernel_params"], row["mean"],
                                   row["mean_params"], row["lensing"], row["lensing_params"], nb, ni, idx))
        want.append(row["logl"])
        theta = row["kernel_params"] + row["mean_params"] + row["lensing_params"]
        gotj.append(o.jointprobability(theta, ds["x"], ds["y"], ds["yerr"], row["kernel"], row["kernel_params"], row["mean"],
                                       row["mean_params"], row["lensing"], row["lensing_params"], nb, ni, idx))
        wantj.append(row["jointprob"])
    assert_logl_close(got, want, "oracle logl")
    assert_logl_close(gotj, wantj, "oracle jointprobability")
    assert sum(np.isnan(want)) >= 10 and sum(np.isfinite(want)) >= 50   # both regimes are covered


def test_covariance_vectors(golden):
    for row in golden["cov"]:
        K = o.covariance(row["kernel"], row["params"], row["x"], row.get("x_prime"))
        np.testing.assert_allclose(K.ravel(), row["K"], rtol=1e-15, atol=0, err_msg=row["kernel"])
        assert list(K.shape) == row["shape"]


def test_mean_vectors(golden):
    for row in golden["mean"]:
        np.testing.assert_allclose(o.mean(row["mean"], row["params"], row["t"]), row["m"], rtol=1e-15, atol=0, err_msg=row["mean"])


def test_lens_and_mask_vectors(golden):
    for row in golden["lens"]:
        xs, b = o.lens(row["lensing"], row["params"], row["x"], row["n_bands"], row["n_images"], row["indices"])
        np.testing.assert_array_equal(xs, row["shifted_x"])
        np.testing.assert_array_equal(np.isnan(b), np.isnan(row["b"]))
        fin = ~np.isnan(b)
        np.testing.assert_allclose(b[fin], np.array(row["b"])[fin], rtol=1e-15, atol=0)
        mask = o.band_mask(row["lensing"], row["n_bands"], row["n_images"], row["indices"])
        np.testing.assert_array_equal(mask.ravel(), row["mask"])


def test_prepare_indices(golden):
    for row in golden["indices"]:
        idx, nb, ni, factor = o.prepare_indices(np.arange(row["n"], dtype=float),
                                                None if row["band"] is None else np.array(row["band"]),
                                                None if row["image"] is None else np.array(row["image"]))
        assert list(idx) == row["indices"] and nb == row["n_bands"] and ni == row["n_images"]
        assert factor == pytest.approx(row["factor"], rel=1e-15)


def test_theta_slicing_under_fix_flags(golden):
    ds = golden["logl_datasets"]["s2x2"]
    idx, nb, ni, _ = _idx(ds)
    for row in golden["slicing"]:
        lp = (lambda v: (lambda p: v))(row["logprior"])
        got = o.jointprobability(row["theta"], ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", row["kernel0"], "UniformMean",
                                 row["mean0"], "ConstantMagnification", row["lens0"], nb, ni, idx, lp,
                                 row["fix_kernel"], row["fix_mean"], row["fix_lens"], row["invert"])
        assert_logl_close([got], [row["value"]], f"slicing {row}")


def test_predict(golden):
    for row in golden["predict"]:
        ds = golden["logl_datasets"][row["dataset"]]
        e, v = o.predict(row["x_prime"], ds["x"], ds["y"], ds["yerr"], row["kernel"], row["kernel_params"], row["mean"], row["mean_params"])
        np.testing.assert_allclose(e, row["expectation"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(v.ravel(), row["variance"], rtol=1e-10, atol=1e-12)


def test_longdouble_arbiter_agrees_with_oracle(golden):
    """On the (well-conditioned) real data sets the float64 oracle and the 80-bit arbiter agree to
    ~1e-10, which bounds how much of a GPU-vs-oracle difference can be blamed on the oracle."""
    for row in golden["kat"][:10]:
        ds = golden["datasets"][row["dataset"]]
        idx, nb, ni, _ = _idx(ds)
        ref = ld.loglikelihood_ld(ds["x"], ds["y"], ds["yerr"], row["kernel"], row["kernel_params"], row["mean_params"][0],
                                  row["lensing_params"], nb, ni, idx)
        assert abs(ref - row["logl"]) < 5e-10, (row["dataset"], row["kernel"], ref - row["logl"])


def test_predict_per_band_as_the_plotting_loop_does():
    """The reference's plotting-loop body (utils.py:193-204: lens the band with the sample, divide by the magnifications,
    GP.predict), run with the reference's own classes by tests/golden/make_golden_predict.py."""
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_predict_bands.json")))
    base = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    for case in ref["cases"]:
        ds = base["logl_datasets"][case["dataset"]]
        x, y, yerr = (np.array(ds[k]) for k in ("x", "y", "yerr"))
        idx, nb, ni, _ = _idx(ds)
        nk, nm = o.KERNEL_NPARAMS[case["kernel"]], o.MEAN_NPARAMS[case["mean"]]
        th = case["sample"]
        shifted, b = o.lens(case["lensing"], th[nk + nm:], x, nb, ni, idx)
        for bi, want in enumerate(case["bands"]):
            r0, r1 = int(idx[bi * ni]), int(idx[(bi + 1) * ni])
            e, v = o.predict(case["x_prime"], shifted[r0:r1], y[r0:r1] / b[r0:r1], yerr[r0:r1] / b[r0:r1],
                             case["kernel"], th[:nk], case["mean"], th[nk:nk + nm])
            np.testing.assert_allclose(e, want["expectation"], rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(v.ravel(), want["variance"], rtol=1e-9, atol=1e-12)
