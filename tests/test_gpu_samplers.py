"""GPU: the sampler adapters driven by protocol fakes of dynesty / emcee (tests/fake_samplers.py), and the stream
contract of a handle (device-buffer and host-buffer calls interleaved on one GP)."""
import time

import numpy as np
import pytest

import benchdata
import fake_samplers
from test_gpu_parity import _synth, _theta_box, _gp_for

pytestmark = pytest.mark.gpu


def _quasar_like_gp():
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    ds = benchdata.make_dataset(141, 1, 3)
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]),
                   lensingmodels.ConstantMagnification([20.0, 0.5, 30.0, 0.6]))
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    return gp


def _ptform(u):
    lo = np.array([0.05, 10.0, -0.2, 0.0, 0.1, 0.0, 0.1])
    hi = np.array([1.0, 40.0, 0.2, 60.0, 1.0, 60.0, 1.0])
    return lo + u * (hi - lo)


def test_fake_dynesty_rslice_batches_reach_the_gpu(libgsn):
    """The reference's default sampler configuration (dynesty, sample='rslice', gausSN.py:318-323) through BatchPool: the
    initial live points are one launch, the proposal walks meet in batches of up to queue_size, and the chains are the
    ones the scalar path produces."""
    from gaussn_b200 import samplers, _engine
    gp = _quasar_like_gp()
    runs = {}
    for name, pool in (("pool", samplers.BatchPool(64, lockstep=True)), ("scalar", None)):
        post = samplers.VectorizedPosterior(gp)
        if pool is not None:
            pool.posterior = post
        t0 = time.perf_counter()
        s = fake_samplers.FakeNestedSampler(post, _ptform, 7, nlive=128, pool=pool, queue_size=64, seed=11).run_nested(3)
        runs[name] = (s, post, time.perf_counter() - t0)
        if pool is not None:
            sizes = pool.batch_sizes
            assert sizes[0] == 128 and max(sizes[1:]) == 64 and sizes[1] == 64
            assert post.n_evals == s.ncall and post.n_calls == len(sizes) < s.ncall / 10
            pool.close()
    s, post, dt = runs["pool"]
    s0, post0, dt0 = runs["scalar"]
    np.testing.assert_array_equal(s.live_logl, s0.live_logl)
    assert post0.n_calls == post0.n_evals == s0.ncall == s.ncall
    assert gp._problem.query(_engine.Q_PATH) == _engine.PATH_ONCHIP
    print(f"N=141: {s.ncall} likelihood calls; BatchPool {post.n_calls} launches ({s.ncall / post.n_calls:.1f} theta per launch) "
          f"{s.ncall / dt:.0f} calls/s; scalar path {s0.ncall / dt0:.0f} calls/s")


def test_fake_emcee_vectorize_is_one_launch_per_half_ensemble(libgsn):
    from gaussn_b200 import samplers
    gp = _quasar_like_gp()
    post = samplers.VectorizedPosterior(gp, logprior=lambda p: 0.0 if 0 < p[1] < 60 else -np.inf)
    rng = np.random.default_rng(5)
    p0 = _ptform(rng.uniform(0.3, 0.7, size=(48, 7)))
    es = fake_samplers.FakeEnsembleSampler(48, 7, post, vectorize=True)
    p, lp = es.run_mcmc(p0, 6, rng)
    assert post.n_calls == 1 + 12 and post.n_evals == 48 + 12 * 24
    scalar = np.array([gp.jointprobability(list(r), post.logprior) for r in p])
    np.testing.assert_array_equal(scalar, lp)


def test_device_call_followed_by_host_call_on_one_handle(libgsn):
    """launches on different streams share the handle's counter and workspace; they must be ordered.  A large
    device-buffer batch (asynchronous, torch's stream) immediately followed by host-buffer calls (the handle's own stream)
    and a change of the theta map, repeatedly."""
    import torch
    rng = np.random.default_rng(99)
    for N, bands in ((400, 1), (141, 1)):        # workspace kernels and the on-chip shape
        ds = _synth(rng, N, bands, 2)
        theta = _theta_box(rng, 20000 if N == 141 else 3000, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2)
        gp = _gp_for(ds, "Matern32Kernel", "UniformMean", "ConstantMagnification", 2, theta[0])
        ref_big = gp.loglikelihood_batch(theta)
        ref_small = gp.loglikelihood_batch(theta[:300])
        ref_fixed = gp.loglikelihood_batch(theta[:300, :3], fix_lensing_params=True)
        th_d = torch.as_tensor(theta, device="cuda")
        side = torch.cuda.Stream()
        for rep in range(4):
            with torch.cuda.stream(side if rep % 2 else torch.cuda.current_stream()):
                big = gp.loglikelihood_batch(th_d)                      # returns at once, kernel still running
            small = gp.loglikelihood_batch(theta[:300])                 # host path, other stream, same handle
            fixed = gp.loglikelihood_batch(theta[:300, :3], fix_lensing_params=True)   # rewrites the theta map
            torch.cuda.synchronize()
            np.testing.assert_array_equal(small, ref_small)
            np.testing.assert_array_equal(fixed, ref_fixed)
            np.testing.assert_array_equal(big.cpu().numpy(), ref_big)
