import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from oracle import gp_oracle as o
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
N, nimg, span = 512, 2, 1500.
ds = benchdata.make_dataset(N, 1, nimg, span=span); theta = benchdata.make_theta(96, nimg)
gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
w = o.loglikelihood_batch(theta, ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", "UniformMean", "ConstantMagnification", 1, nimg, ds["indices"])
for trial in range(3):
    g = gp.loglikelihood_batch(theta)
    e = np.abs(g-w)
    print("trial", trial, "n_bad", (e>1e-6).sum(), "max %.3e" % e.max(), "bad idx", np.where(e>1e-6)[0][:20], flush=True)
g1 = gp.loglikelihood_batch(theta[:32]); print("first 32 alone: max %.3e" % np.abs(g1-w[:32]).max())
g2 = gp.loglikelihood_batch(theta[32:]); e2=np.abs(g2-w[32:]); print("rest alone: nbad", (e2>1e-6).sum(), "max %.3e" % e2.max())
bad = np.where(np.abs(gp.loglikelihood_batch(theta)-w)>1e-6)[0]
for i in bad[:6]: print(i, theta[i], w[i], gp.loglikelihood_batch(theta[i:i+1])[0]-w[i])
