"""Cost of the host-supplied mean path (sncosmoMean, meanfuncs.py:47-121) for the SN notebook's SALT configuration
(ipynb/fitting_SLSN_example.ipynb:262-300): B = 500 thetas (dynesty's nlive), N = 194 epochs in 4 bands x 2 images.
sncosmo is not installed, so the model is a stand-in whose bandflux() evaluates a Bazin curve and then burns CPU time to
the cost of a real SALT bandflux call (about 1 ms for 194 epochs: spline evaluation of the template in phase and
wavelength, band integration).  Reported: the engine's own share of a batch (lens on the device, copies, the batched
likelihood) and the share of the host loop that calls the model B times."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels


class StandInModel:
    param_names = ["A", "beta", "t0", "Tfall", "Trise"]

    def __init__(self, cost_s):
        self.parameters = np.array([1.0, 0.1, 30.0, 40.0, 8.0]); self.cost_s = cost_s

    def update(self, d):
        for k, v in d.items():
            self.parameters[self.param_names.index(k)] = v

    def bandflux(self, bands, t, zp=None, zpsys=None):
        A, beta, t0, Tf, Tr = self.parameters
        out = A * np.exp(-(t - t0) / Tf) / (1 + np.exp((t - t0) / Tr)) + beta
        end = time.perf_counter() + self.cost_s
        while time.perf_counter() < end:
            pass
        return out


ds = benchdata.make_dataset(194, 4, 2)
rng = np.random.default_rng(5)
B = 500
theta = np.column_stack([rng.uniform(0.05, 1, B), rng.uniform(10, 40, B), rng.uniform(0.5, 2, B), rng.uniform(0, 0.2, B), rng.uniform(20, 40, B),
                         rng.uniform(25, 50, B), rng.uniform(4, 10, B), rng.uniform(0, 60, B), rng.uniform(0.1, 1, B)])
for cost in (0.0, 1e-3):
    gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.sncosmoMean(StandInModel(cost)), lensingmodels.ConstantMagnification([20.0, 0.5]))
    gp.x, gp.y, gp.yerr, gp.bands = ds["x"], ds["y"], ds["yerr"], ds["band"]
    gp._prepare_indices(ds["x"], ds["band"], ds["image"])
    gp.loglikelihood_batch(theta[:8])
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); out = gp.loglikelihood_batch(theta); ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print(json.dumps(dict(B=B, N=194, model_cost_per_call_ms=cost * 1e3, batch_ms=round(dt * 1e3, 2), per_theta_us=round(dt / B * 1e6, 1),
                          engine_and_python_overhead_ms=round((dt - B * cost) * 1e3, 2), finite=float(np.mean(np.isfinite(out))))), flush=True)
