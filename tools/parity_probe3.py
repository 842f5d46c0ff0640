import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from oracle import gp_oracle as o
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
for N, nimg, span in [(512,2,150.),(512,2,400.),(512,2,1500.),(256,2,1500.),(128,2,1500.),(64,2,1500.),(512,1,1500.),(1024,4,512.),(1024,2,150.)]:
    ds = benchdata.make_dataset(N, 1, nimg, span=span); theta = benchdata.make_theta(32, nimg)
    lm = lensingmodels.ConstantMagnification([20.0, 0.5]*(nimg-1)) if nimg > 1 else lensingmodels.NoLensing()
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lm)
    gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    g = gp.loglikelihood_batch(theta)
    w = o.loglikelihood_batch(theta, ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", "UniformMean", "ConstantMagnification" if nimg>1 else "NoLensing", 1, nimg, ds["indices"])
    e = np.abs(g-w)
    print(N, nimg, span, "max err %.3e" % e.max(), "at logL", w[np.argmax(e)], "median %.2e" % np.median(e), flush=True)
