"""Throughput of the BASELINE.json configurations C1-C4, device-resident theta."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels

ONLY = sys.argv[1:]          # optional substrings: run only the configurations whose name contains one of them


def run(name, N, n_bands, n_images, kernel, kp, mean, mp, lens, lp_per, B, kname):
    if ONLY and not any(o in name for o in ONLY):
        return
    ds = benchdata.make_dataset(N, n_bands, n_images, span=150.0 if N < 1000 else 2.0 * N / n_images)
    rng = np.random.default_rng(3)
    cols = [rng.uniform(lo, hi, B) for lo, hi in kp + mp]
    for _ in range(n_images - 1):
        cols += [rng.uniform(lo, hi, B) for lo, hi in lp_per]
    theta = torch.as_tensor(np.column_stack(cols), device="cuda")
    lm = lensingmodels.NoLensing() if lens is None else lens([1.0] * (len(lp_per) * (n_images - 1)))
    gp = gausSN.GP(kernel([1.0] * len(kp)), mean([1.0] * len(mp)), lm)
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    for _ in range(2):
        out = gp.loglikelihood_batch(theta)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = gp.loglikelihood_batch(theta)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = benchdata.algorithmic_flops(N, n_bands, lens is not None, kname)
    print(json.dumps(dict(config=name, N=N, bands=n_bands, images=n_images, P=theta.shape[1], B=B, ms=ms, evals_per_s=B / ms * 1e3,
                          tflops=flops * B / ms / 1e9, frac_of_fp64_peak=flops * B / ms / 1e9 / 37.165,
                          finite=float(torch.isfinite(out).float().mean()), shape=int(gp._problem.query(10)),
                          warps_per_theta=int(gp._problem.query(15)))), flush=True)

A, TAU, C = (0.05, 1.0), (10, 40), (-0.2, 0.2)
CM = [(0, 60), (0.1, 1.0)]
SG = [(0, 60), (0.3, 1.0), (-0.3, 0.3), (0.02, 0.3), (30, 90)]
run("C1 getting-started-like", 64, 1, 2, kernels.ExpSquaredKernel, [A, TAU], meanfuncs.UniformMean, [C], lensingmodels.ConstantMagnification, CM, 262144, "ExpSquaredKernel")
run("C1 literal notebook N=11, NoLensing", 11, 1, 1, kernels.ExpSquaredKernel, [A, TAU], meanfuncs.UniformMean, [C], None, [], 262144, "ExpSquaredKernel")
run("C2 3 bands x 4 images, Matern32 + sigmoid", 396, 3, 4, kernels.Matern32Kernel, [(0.01, 0.3), TAU], meanfuncs.UniformMean, [C], lensingmodels.SigmoidMagnification, SG, 131072, "Matern32Kernel")
run("C3 quasar-like N=2048, OU", 2048, 1, 4, kernels.OUKernel, [(0.01, 0.3), (50, 400)], meanfuncs.UniformMean, [C], lensingmodels.ConstantMagnification, CM, 4096, "OUKernel")
run("C3 real-file shape N=141, OU", 141, 1, 3, kernels.OUKernel, [(0.01, 0.3), (50, 400)], meanfuncs.UniformMean, [C], lensingmodels.ConstantMagnification, CM, 262144, "OUKernel")
run("C4 SLSN-like 4 bands x 2 images N=194, Bazin", 194, 4, 2, kernels.ExpSquaredKernel, [A, TAU], meanfuncs.Bazin2009, [(0.5, 2), (0, 0.2), (20, 40), (25, 50), (4, 10)], lensingmodels.ConstantMagnification, CM, 262144, "ExpSquaredKernel")
run("C4 synthetic N=512, 4 bands x 2 images, Bazin", 512, 4, 2, kernels.ExpSquaredKernel, [A, TAU], meanfuncs.Bazin2009, [(0.5, 2), (0, 0.2), (20, 40), (25, 50), (4, 10)], lensingmodels.ConstantMagnification, CM, 131072, "ExpSquaredKernel")
run("C4 dynesty nlive batch B=500, N=194", 194, 4, 2, kernels.ExpSquaredKernel, [A, TAU], meanfuncs.Bazin2009, [(0.5, 2), (0, 0.2), (20, 40), (25, 50), (4, 10)], lensingmodels.ConstantMagnification, CM, 500, "ExpSquaredKernel")
