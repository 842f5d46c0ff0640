"""Latency of calls with few thetas: the one-theta-per-cluster shape against the one-CTA shape (GSN_FORCE_SHAPE=wide)."""
import sys, os, time, json, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import benchdata
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels, _engine
    for N in (256, 320, 384, 429, 512, 768, 1024, 2048, 4096):
        ni = 3 if N == 429 else 2
        ds = benchdata.make_dataset(N, 1, ni)
        theta = benchdata.make_theta(128, ni)
        gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5] * (ni - 1)))
        gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
        gp.loglikelihood_batch(theta[:2])
        row = dict(N=N, force=os.environ.get("GSN_FORCE_SHAPE", "auto"))
        for B in (1, 8, 16, 24, 32, 48, 64, 74, 100):
            reps = max(3, min(100, int(2000 / N)))
            gp.loglikelihood_batch(theta[:B])
            t0 = time.perf_counter()
            for _ in range(reps): gp.loglikelihood_batch(theta[:B])
            row[f"B{B}_ms"] = round((time.perf_counter() - t0) / reps * 1e3, 3)
            row[f"B{B}_shape"] = int(gp._problem.query(_engine.Q_PATH))
        print(json.dumps(row), flush=True)
else:
    for force in (None, "wide", "cluster2", "cluster4", "cluster8"):
        env = dict(os.environ)
        if force: env["GSN_FORCE_SHAPE"] = force
        subprocess.run([sys.executable, __file__, "--child"], env=env, check=False)
