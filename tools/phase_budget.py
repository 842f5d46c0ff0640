"""Per-phase cycle budget of the four-warp workspace kernel at the headline size (debug build with -DGSN_PHASE_TIMERS:
every warp accumulates clock64 deltas per phase; summed over the launch).  Usage: phase_budget.py [N] [B]"""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels, _engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
ds = benchdata.make_dataset(N, 1, 2)
theta = torch.as_tensor(benchdata.make_theta(B, 2), device="cuda")
gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
gp._prepare_indices(ds["x"], ds["band"], ds["image"])
lib = _engine.load()
buf = (ctypes.c_ulonglong * 16)()
gp.loglikelihood_batch(theta); torch.cuda.synchronize()
lib.gsn_debug_phase_cycles(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); gp.loglikelihood_batch(theta); e1.record(); torch.cuda.synchronize()
lib.gsn_debug_phase_cycles(buf, 1)
c = np.array(buf[:10], dtype=np.float64)
names = ["prologue", "item fetch", "covariance build", "off-diagonal GEMM", "wait for pivot block", "solve + store", "diagonal GEMM",
         "pivot block factorisation", "end of theta (barrier, reduce)", "(flag waits inside the GEMMs)"]
tot = c[:9].sum()
print(json.dumps(dict(N=N, B=B, ms=e0.elapsed_time(e1), path=int(gp._problem.query(10)),
                      share={n: round(float(v / tot), 4) for n, v in zip(names, c)}, warp_cycles_total=float(tot))))
