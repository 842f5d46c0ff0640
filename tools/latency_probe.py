"""Scalar-call and small-batch latency (the reference's samplers call the likelihood one theta at a time)."""
import sys, os, time, json, gc
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
for N in (141, 194, 429, 512):
    ds = benchdata.make_dataset(N, 1, 3 if N == 141 else 2)
    ni = ds["n_images"]
    theta = benchdata.make_theta(4096, ni)
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5] * (ni - 1)))
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    gp.loglikelihood_batch(theta[:8])
    gc.collect()        # the previous handle's workspace is freed now, not inside a timing loop
    row = dict(N=N, shape=os.environ.get("GSN_FORCE_SHAPE", "auto"))
    t0 = time.perf_counter(); n = 300
    for i in range(n): gp.jointprobability(list(theta[i]))
    row["scalar_calls_per_s"] = round(n / (time.perf_counter() - t0))
    for B in (1, 16, 64, 148, 500, 4096):
        reps = 50 if B <= 500 else 10
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); gp.loglikelihood_batch(theta[:B]); ts.append(time.perf_counter() - t0)
        row[f"B{B}_ms"] = round(float(np.median(ts)) * 1e3, 3)      # median: one-off costs (allocations) do not count
    print(json.dumps(row), flush=True)
