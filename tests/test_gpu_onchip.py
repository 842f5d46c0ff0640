"""The on-chip launch shape (gsn_onchip.cuh): band blocks of at most 224 rows are factored with the Cholesky factor
resident in shared memory, at 8-row granularity, by teams of 1-16 warps per theta.  Results must not depend on the team
size or the batch (bit for bit), must agree with the oracle within |dlogL| <= max(1e-8, 1e-12 |logL|), and with the
workspace kernels (gsn_chol_dmma.cuh: same factorisation, residual row and pivot products accumulated in another order)
ten times closer than that."""
import numpy as np
import pytest

from conftest import assert_logl_close
from oracle import gp_oracle
from test_gpu_parity import _synth, _theta_box, _gp_for, _oracle_batch

pytestmark = pytest.mark.gpu


def _agree(got, legacy, what=""):
    assert_logl_close(got, legacy, what + " (on-chip vs workspace kernels)", atol=1e-9, rtol=1e-13)


def _legacy(monkeypatch, ds, kernel, mean, lens, n_images, theta, **kw):
    """The same problem through the workspace kernels (GSN_FORCE_SHAPE=legacy is read when the handle is created)."""
    from gaussn_b200 import _engine
    monkeypatch.setenv("GSN_FORCE_SHAPE", "legacy")
    monkeypatch.delenv("GSN_ONCHIP_MAX_ROWS", raising=False)
    gp = _gp_for(ds, kernel, mean, lens, n_images, theta[0])
    out = gp.loglikelihood_batch(theta, **kw)
    assert gp._problem.query(_engine.Q_ONCHIP) == 0 and gp._problem.query(_engine.Q_PATH) != _engine.PATH_ONCHIP
    monkeypatch.delenv("GSN_FORCE_SHAPE")
    return out


@pytest.mark.parametrize("N", [1, 2, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 40, 63, 64, 65, 96, 100, 141, 160, 161, 192, 194, 217, 224])
def test_every_order_up_to_224_matches_oracle_and_workspace_kernels(libgsn, monkeypatch, N):
    """The shape is compiled for band blocks of up to 224 rows and dispatched up to 208, while two thetas fit an SM (GSN_ONCHIP_MAX_ROWS moves the limit)."""
    from gaussn_b200 import _engine
    monkeypatch.setenv("GSN_ONCHIP_MAX_ROWS", "224")
    rng = np.random.default_rng(5000 + N)
    n_images = 2 if N >= 2 else 1
    ds = _synth(rng, N, 1, n_images)
    theta = _theta_box(rng, 48, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", n_images)
    theta[5, 0] = -abs(theta[5, 0]) if N > 1 else theta[5, 0]
    gp = _gp_for(ds, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", n_images, theta[0])
    got, info = gp.loglikelihood_batch(theta), None
    assert gp._problem.query(_engine.Q_ONCHIP) == 1 and gp._problem.query(_engine.Q_PATH) == _engine.PATH_ONCHIP
    assert_logl_close(got, _oracle_batch(ds, theta, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", gp), f"on-chip N={N}")
    _agree(got, _legacy(monkeypatch, ds, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", n_images, theta))


@pytest.mark.parametrize("N,n_bands,onchip", [(192, 1, 1), (208, 1, 1), (209, 1, 0), (225, 1, 0), (832, 4, 1), (844, 4, 0)])
def test_dispatch_limit_is_the_largest_band_block(libgsn, N, n_bands, onchip):
    from gaussn_b200 import _engine
    rng = np.random.default_rng(N)
    ds = _synth(rng, N, n_bands, 2)
    theta = _theta_box(rng, 8, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", 2)
    gp = _gp_for(ds, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    got = gp.loglikelihood_batch(theta)
    assert gp._problem.query(_engine.Q_ONCHIP) == onchip
    assert_logl_close(got, _oracle_batch(ds, theta, "ExpSquaredKernel", "UniformMean", "ConstantMagnification", gp), f"N={N}")


@pytest.mark.parametrize("kernel", ["ExpSquaredKernel", "ExponentialKernel", "ConstantKernel", "Matern32Kernel", "Matern52Kernel",
                                    "OUKernel", "GibbsKernel", "PeriodicKernel", "RationalQuadraticKernel", "JJKernel"])
@pytest.mark.parametrize("n_bands,n_images,N", [(1, 3, 141), (4, 2, 194), (3, 4, 396)])
def test_every_kernel_on_the_reference_shapes(libgsn, monkeypatch, kernel, n_bands, n_images, N):
    """The quasar notebook's shape (N = 141, 3 images), the SN notebook's (4 bands x 2 images, N = 194) and BASELINE
    config 2 (3 bands x 4 images, N = 396)."""
    from gaussn_b200 import _engine
    rng = np.random.default_rng(N + 17 * n_bands)
    ds = _synth(rng, N, n_bands, n_images)
    theta = _theta_box(rng, 40, kernel, "UniformMean", "ConstantMagnification", n_images)
    gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta[0])
    got = gp.loglikelihood_batch(theta)
    assert gp._problem.query(_engine.Q_PATH) == _engine.PATH_ONCHIP
    assert_logl_close(got, _oracle_batch(ds, theta, kernel, "UniformMean", "ConstantMagnification", gp), f"{kernel} {n_bands}x{n_images} N={N}")
    _agree(got, _legacy(monkeypatch, ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta))


@pytest.mark.parametrize("n_bands,n_images,N,kernel", [(1, 2, 64, "ExpSquaredKernel"), (1, 3, 141, "OUKernel"), (4, 2, 194, "Matern32Kernel"),
                                                       (3, 4, 396, "Matern32Kernel"), (2, 2, 448, "Matern52Kernel"), (7, 1, 100, "ExpSquaredKernel")])
def test_team_size_and_batch_order_do_not_change_a_bit(libgsn, monkeypatch, n_bands, n_images, N, kernel):
    """1, 2, 4, 8 or 16 warps per theta; 700 thetas (several waves of the persistent grid), 40, 3 and 1 per call; reversed."""
    from gaussn_b200 import _engine
    monkeypatch.setenv("GSN_ONCHIP_MAX_ROWS", "224")
    rng = np.random.default_rng(N * 3 + n_bands)
    ds = _synth(rng, N, n_bands, n_images)
    lens = "ConstantMagnification" if n_images > 1 else "NoLensing"
    theta = _theta_box(rng, 700, kernel, "UniformMean", lens, n_images)
    theta[11, 1] = np.nan                       # a theta that cannot be evaluated, inside the batch
    gp = _gp_for(ds, kernel, "UniformMean", lens, n_images, theta[0])
    ref = gp.loglikelihood_batch(theta)
    assert np.isnan(ref[11]) and np.isfinite(ref).sum() > 600
    seen = {int(gp._problem.query(_engine.Q_ONCHIP_WARPS))}
    for T in (1, 2, 4, 8, 16):
        monkeypatch.setenv("GSN_ONCHIP_T", str(T))
        np.testing.assert_array_equal(gp.loglikelihood_batch(theta), ref, err_msg=f"T={T}")
        assert gp._problem.query(_engine.Q_ONCHIP_WARPS) == T
        np.testing.assert_array_equal(gp.loglikelihood_batch(theta[:3]), ref[:3], err_msg=f"T={T}, 3 thetas")
    monkeypatch.delenv("GSN_ONCHIP_T")
    np.testing.assert_array_equal(gp.loglikelihood_batch(theta[::-1].copy())[::-1], ref)
    np.testing.assert_array_equal(np.concatenate([gp.loglikelihood_batch(theta[i:i + 40]) for i in range(0, 200, 40)]), ref[:200])
    seen.add(int(gp._problem.query(_engine.Q_ONCHIP_WARPS)))
    np.testing.assert_array_equal(np.array([gp.loglikelihood_batch(theta[i:i + 1])[0] for i in range(12)]), ref[:12])
    seen.add(int(gp._problem.query(_engine.Q_ONCHIP_WARPS)))
    print("team sizes chosen by the dispatch for 700 / 40 / 1 thetas:", sorted(seen))
    assert_logl_close(ref[:64], _oracle_batch(ds, theta[:64], kernel, "UniformMean", lens, gp), "team")


@pytest.mark.parametrize("mean", ["UniformMean", "Sin", "Gaussian", "ExpFunction", "Bazin2009", "Karpenka2012", "Villar2019"])
@pytest.mark.parametrize("lens", ["NoLensing", "ConstantMagnification", "SigmoidMagnification"])
def test_every_mean_and_lensing_model_on_chip(libgsn, monkeypatch, mean, lens):
    from gaussn_b200 import _engine
    rng = np.random.default_rng(abs(hash(mean + lens)) % 2**31)
    n_images = 1 if lens == "NoLensing" else 3
    ds = _synth(rng, 170, 2, n_images)
    theta = _theta_box(rng, 32, "Matern32Kernel", mean, lens, n_images)
    gp = _gp_for(ds, "Matern32Kernel", mean, lens, n_images, theta[0])
    got = gp.loglikelihood_batch(theta)
    assert gp._problem.query(_engine.Q_PATH) == _engine.PATH_ONCHIP
    assert_logl_close(got, _oracle_batch(ds, theta, "Matern32Kernel", mean, lens, gp), mean + "/" + lens)
    _agree(got, _legacy(monkeypatch, ds, "Matern32Kernel", mean, lens, n_images, theta))


@pytest.mark.parametrize("n_bands", [1, 3])
def test_failure_info_is_the_callers_row(libgsn, monkeypatch, n_bands):
    """info = 1-based index, in the caller's row order, of a non-positive pivot.  One band: the first such pivot, the same
    in both shapes.  With a band mask the internal layouts differ (bands aligned to 8 rows here, to 32 in the workspace
    kernels, which also factor the bands concurrently and report whichever failure comes first in time): both must name a
    pivot that LAPACK's factorisation of that band also rejects or precedes."""
    import scipy.linalg
    from gaussn_b200 import _engine
    rng = np.random.default_rng(12)
    ds = _synth(rng, 150, n_bands, 2)
    theta = _theta_box(rng, 16, "RationalQuadraticKernel", "UniformMean", "ConstantMagnification", 2)
    theta[:, 0] *= 40.0        # A^2 (1 + d^2 / ...) with a large amplitude: indefinite early on
    gp = _gp_for(ds, "RationalQuadraticKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    gp.loglikelihood_batch(theta)
    out, info = gp._problem.logl_host(theta, 0, return_info=True)
    assert gp._problem.query(_engine.Q_PATH) == _engine.PATH_ONCHIP
    assert np.isnan(out).any() and np.all((info > 0) == np.isnan(out)) and info.max() <= 150
    monkeypatch.setenv("GSN_FORCE_SHAPE", "legacy")
    gp2 = _gp_for(ds, "RationalQuadraticKernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    gp2.loglikelihood_batch(theta)
    out2, info2 = gp2._problem.logl_host(theta, 0, return_info=True)
    _agree(out, out2)
    if n_bands == 1:
        np.testing.assert_array_equal(info, info2)
    else:
        assert np.all((info2 > 0) == np.isnan(out2)) and info2.max() <= 150
    # the first failing pivot of the first failing band, from LAPACK (potrf reports the order of the leading minor)
    per_band = 150 // n_bands
    for t in np.flatnonzero(info > 0)[:6]:
        shifted, b = gp_oracle.lens("ConstantMagnification", list(theta[t, 4:]), ds["x"], n_bands, 2, gp.indices)
        K = np.outer(b, b) * gp_oracle.covariance("RationalQuadraticKernel", list(theta[t, :3]), shifted) + np.diag(ds["yerr"] ** 2)
        first = None
        for band in range(n_bands):
            r0, r1 = band * per_band, (band + 1) * per_band if band < n_bands - 1 else 150
            _, lapack_info = scipy.linalg.lapack.dpotrf(K[r0:r1, r0:r1], lower=1)
            if lapack_info > 0:
                first = r0 + lapack_info
                break
        assert first is not None and abs(int(info[t]) - first) <= 1, (t, info[t], first)   # borderline pivots may flip by rounding


def test_new_observations_reach_the_on_chip_tables(libgsn):
    rng = np.random.default_rng(55)
    ds = _synth(rng, 90, 2, 2)
    theta = _theta_box(rng, 20, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2)
    gp = _gp_for(ds, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2, theta[0])
    gp.loglikelihood_batch(theta)
    handle = gp._problem
    ds2 = dict(ds, y=ds["y"] * 1.1 + 0.01, yerr=ds["yerr"] * 0.9, x=ds["x"] + 0.25)
    got = gp.loglikelihood_batch(theta, x=ds2["x"], y=ds2["y"], yerr=ds2["yerr"])
    assert gp._problem is handle
    assert_logl_close(got, _oracle_batch(ds2, theta, "Matern32Kernel", "UniformMean", "ConstantMagnification", gp), "set_data")


@pytest.mark.parametrize("n_bands,n_images,N,kernel", [(1, 3, 141, "OUKernel"), (4, 2, 194, "ExpSquaredKernel"), (1, 2, 190, "Matern32Kernel")])
def test_repeated_runs_with_every_team_size_are_bit_identical(libgsn, monkeypatch, n_bands, n_images, N, kernel):
    """Race detector of our own (compute-sanitizer is closed on this pool): the dataflow between the warps of a team --
    progress counters, column counters that guard the recycling of tile slots -- is exercised 25 times per team size with
    several waves of thetas per SM; a missing wait shows up as a value that depends on timing."""
    rng = np.random.default_rng(N + n_bands)
    ds = _synth(rng, N, n_bands, n_images)
    theta = _theta_box(rng, 1500, kernel, "UniformMean", "ConstantMagnification", n_images)
    gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta[0])
    ref = gp.loglikelihood_batch(theta)
    assert np.isfinite(ref).sum() > 1400
    for T in (2, 4, 8, 16):
        monkeypatch.setenv("GSN_ONCHIP_T", str(T))
        for rep in range(25):
            got = gp.loglikelihood_batch(theta)
            assert np.array_equal(got, ref, equal_nan=True), f"T={T} repetition {rep}: {int(np.sum(got != ref))} values differ"


@pytest.mark.parametrize("kernel", ["ExpSquaredKernel", "Matern32Kernel", "OUKernel", "GibbsKernel", "PeriodicKernel", "JJKernel"])
@pytest.mark.parametrize("n_bands,n_images,N", [(1, 1, 5), (1, 2, 12), (1, 2, 20), (1, 3, 30), (1, 2, 40), (1, 2, 47), (1, 2, 56),
                                                (1, 2, 64), (4, 2, 194), (7, 1, 100)])
def test_register_shape_is_bit_identical_to_the_shared_memory_shape(libgsn, monkeypatch, kernel, n_bands, n_images, N):
    """Band blocks of at most 64 rows, one warp per theta: the factor lives in registers (reg_factor, every unrolled block
    size 1 ... 8 tile rows is hit here, alone and mixed within one theta).  Same tile arithmetic in the same order as the
    shared-memory sweep: the same bits, failures included; and the oracle's values."""
    from gaussn_b200 import _engine
    rng = np.random.default_rng(9000 + 31 * N + n_bands)
    ds = _synth(rng, N, n_bands, n_images)
    theta = _theta_box(rng, 1200, kernel, "UniformMean", "ConstantMagnification", n_images)   # enough thetas for one-warp teams
    theta[3, 0] = np.nan
    theta[7, 1] = np.inf
    gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", n_images, theta[0])
    gp.loglikelihood_batch(theta[:2])                           # programs the handle and the theta map
    got, info = gp._problem.logl_host(theta, 0, return_info=True)
    assert gp._problem.query(_engine.Q_PATH) == _engine.PATH_ONCHIP and gp._problem.query(_engine.Q_ONCHIP_REG) == 1
    assert gp._problem.query(_engine.Q_ONCHIP_WARPS) == 1
    monkeypatch.setenv("GSN_ONCHIP_REG", "0")
    want, info0 = gp._problem.logl_host(theta, 0, return_info=True)
    assert gp._problem.query(_engine.Q_ONCHIP_REG) == 0
    monkeypatch.delenv("GSN_ONCHIP_REG")
    assert np.array_equal(got, want, equal_nan=True) and np.array_equal(info, info0)
    assert not np.isfinite(got[3]) and np.isfinite(got).sum() >= 1100      # (an infinite length scale can be a valid model)
    sel = np.r_[0:3, 4:40, 1150:1200]
    assert_logl_close(got[sel], _oracle_batch(ds, theta[sel], kernel, "UniformMean", "ConstantMagnification", gp), f"register shape {kernel} N={N}")
