"""SURVEY 8(d) grid: N in {64 ... 4096} x B in {1 024, 8 192, 65 536, 1 048 576}, B capped so that one launch stays
under ~2 s.  theta resident in HBM, CUDA events, median of 3 launches after 2 warm-ups.  One JSON line per cell."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels, _engine

peak = 37.165
for N in [64, 128, 256, 512, 1024, 2048, 4096]:
    ds = benchdata.make_dataset(N, 1, 2)
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    flops = benchdata.algorithmic_flops(N, 1, True)
    cap = int(2.0 * 0.55 * peak * 1e12 / flops)          # thetas that take about 2 s at 55 % of peak
    for B_nominal in [1024, 8192, 65536, 1048576]:
        B = min(B_nominal, max(1024, cap))
        theta = torch.as_tensor(benchdata.make_theta(B, 2), device="cuda")
        for _ in range(2):
            out = gp.loglikelihood_batch(theta)
        torch.cuda.synchronize()
        times = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = gp.loglikelihood_batch(theta); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        print(json.dumps(dict(N=N, B_nominal=B_nominal, B=B, capped=B != B_nominal, ms=round(ms, 4), evals_per_s=round(B / ms * 1e3, 1),
                              tflops=round(flops * B / ms / 1e9, 3), frac_of_fp64_peak=round(flops * B / ms / 1e9 / peak, 4),
                              shape=int(gp._problem.query(_engine.Q_PATH)))), flush=True)
        del theta
