import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from oracle import gp_oracle as o
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
N, nb, ni, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
ds = benchdata.make_dataset(N, nb, ni)
theta = benchdata.make_theta(B, ni)
gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5] * (ni - 1)))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
g = gp.loglikelihood_batch(theta)
w = o.loglikelihood_batch(theta[:8], ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", "UniformMean", "ConstantMagnification", nb, ni, ds["indices"])
print("ok", np.abs(g[:8] - w).max())
