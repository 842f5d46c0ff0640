"""Dump GPU logL for a sample of the bench workload (run on the GPU box); analysed offline
against the float64 oracle and the long-double arbiter."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
ds = benchdata.make_dataset(N, 1, 2)
theta = benchdata.make_theta(B, 2)
gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
gp._prepare_indices(ds["x"], ds["band"], ds["image"])
ll = gp.loglikelihood_batch(theta)
os.makedirs("gpurun_out", exist_ok=True)
np.savez(f"gpurun_out/probe_N{N}.npz", theta=theta, logl=ll, x=ds["x"], y=ds["y"], yerr=ds["yerr"], indices=ds["indices"])
print("saved", ll[:4])
