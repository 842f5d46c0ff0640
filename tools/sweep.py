"""Throughput sweep over N: device-resident theta, CUDA events.
Prints one JSON line per N; run on the GPU box."""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels, _engine

peak = 37.165
sizes = [int(a) for a in sys.argv[1:]] or [32, 64, 128, 141, 194, 256, 396, 429, 512, 1024, 2048, 4096]
for N in sizes:
    ds = benchdata.make_dataset(N, 1, 2)
    B = int(min(262144, max(2048, 3.0e12 / (N ** 3 / 3 + 40 * N * N + 2e5))))   # ~0.1-0.3 s of work
    B = (B // 444 + 1) * 444
    theta = torch.as_tensor(benchdata.make_theta(B, 2), device="cuda")
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    for _ in range(2):
        out = gp.loglikelihood_batch(theta)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    reps = 3
    ev[0].record()
    for _ in range(reps):
        out = gp.loglikelihood_batch(theta)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    flops = benchdata.algorithmic_flops(N, 1, True)
    print(json.dumps(dict(N=N, B=B, ms=ms, evals_per_s=B / ms * 1e3, tflops=flops * B / ms / 1e9, frac=flops * B / ms / 1e9 / peak,
                          finite=float(torch.isfinite(out).float().mean()))), flush=True)
