import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
out = {}
ds = benchdata.make_dataset(512, 1, 2, span=1500.0); theta = benchdata.make_theta(96, 2)
gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
out["a_ll"] = gp.loglikelihood_batch(theta); out["a_theta"] = theta
N = 1024; kernel = "Matern32Kernel"
ds = benchdata.make_dataset(N, 1, 4, span=2.0 * (N // 4)); rng = np.random.default_rng(N); B = 12
theta = np.column_stack([rng.uniform(0.01, 0.3, B), rng.uniform(50, 400, B), rng.uniform(-0.1, 0.3, B)] +
                        [f(B) for _ in range(3) for f in (lambda b: rng.uniform(0, 60, b), lambda b: rng.uniform(0.3, 1.5, b))])
gp = gausSN.GP(getattr(kernels, kernel)([0.1, 100.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([10.0, 1.0] * 3))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
out["b_ll"] = gp.loglikelihood_batch(theta); out["b_theta"] = theta
np.savez("gpurun_out/probe2.npz", **out)
