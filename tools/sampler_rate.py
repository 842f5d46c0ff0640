"""Likelihood calls per second under the reference's default sampler protocol (dynesty, sample='rslice', nlive=500,
gausSN.py:318-323) reproduced by tests/fake_samplers.py: scalar calls (what the reference does), BatchPool in lock step,
BatchPool 'auto'.  Usage: sampler_rate.py [N ...]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import benchdata, fake_samplers
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels, samplers

lo = np.array([0.05, 10.0, -0.2, 0.0, 0.1]); hi = np.array([1.0, 40.0, 0.2, 60.0, 1.0])
ptform = lambda u: lo + u * (hi - lo)
for N in [int(a) for a in sys.argv[1:]] or [141, 512, 1024]:
    ds = benchdata.make_dataset(N, 1, 2, span=150.0 * max(1.0, N / 141.0))
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    gp.loglikelihood_batch(ptform(np.random.default_rng(0).uniform(size=(500, 5))))
    row = dict(N=N, nlive=500)
    for name, pool, q in (("scalar", None, 1), ("lockstep_q64", samplers.BatchPool(64, lockstep=True), 64),
                          ("lockstep_q250", samplers.BatchPool(250, lockstep=True), 250)):
        post = samplers.VectorizedPosterior(gp)
        if pool is not None:
            pool.posterior = post
        s = fake_samplers.FakeNestedSampler(post, ptform, 5, nlive=500, pool=pool, queue_size=q, seed=7)
        n0 = post.n_evals
        t0 = time.perf_counter()
        s.run_nested(2 if q > 1 else (60 if N <= 1024 else 12))     # a scalar "round" is one walk
        dt = time.perf_counter() - t0
        n = post.n_evals - n0
        row[name] = dict(calls_per_s=round(n / dt), calls=n, launches=post.n_calls - 1, theta_per_launch=round(n / max(1, post.n_calls - 1), 1))
        if pool is not None:
            pool.close()
    print(json.dumps(row), flush=True)
