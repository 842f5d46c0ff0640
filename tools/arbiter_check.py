"""GPU log-likelihood against the 80-bit arbiter (oracle/longdouble.py) and against LAPACK, on the headline workload
(span 150 d: cond ~ 1e8) and at the shipped data's cadence (span 1500 d).  Usage: arbiter_check.py [n_theta] [N]"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata, bench
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
if __name__ == "__main__":
    pool = bench.CpuPool()
    theta = benchdata.make_theta(n, 2)
    for span in ([float(v) for v in os.environ["SPANS"].split(",")] if os.environ.get("SPANS") else (150.0, 1500.0 * N / 512.0)):
        ds = benchdata.make_dataset(N, 1, 2, span=span)
        gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
        gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
        gp._prepare_indices(ds["x"], ds["band"], ds["image"])
        gpu = gp.loglikelihood_batch(theta)
        cpu, _, _ = pool.logl(ds, theta)
        arb = pool.arbiter(ds, theta)
        f = np.isfinite(arb) & np.isfinite(cpu) & np.isfinite(gpu)
        eg, ec = np.abs(gpu[f] - arb[f]), np.abs(cpu[f] - arb[f])
        tol = np.maximum(1e-8, 1e-12 * np.abs(arb[f]))
        print(json.dumps(dict(N=N, span=span, n=int(f.sum()), gpu_max=float(eg.max()), lapack_max=float(ec.max()), gpu_median=float(np.median(eg)),
                              lapack_median=float(np.median(ec)), gpu_p90=float(np.quantile(eg, 0.9)), lapack_p90=float(np.quantile(ec, 0.9)),
                              gpu_over_tol=int(np.sum(eg > tol)), lapack_over_tol=int(np.sum(ec > tol)),
                              gpu_rel_max=float((eg / np.abs(arb[f])).max()), lapack_rel_max=float((ec / np.abs(arb[f])).max()),
                              shape=int(gp._problem.query(10)))), flush=True)
    pool.close()
