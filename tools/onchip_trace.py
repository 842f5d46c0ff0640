"""Per-column timeline of the pivot chain of one theta in the on-chip kernel (debug build with -DGSN_ONCHIP_TRACE=<n>:
CTA 0 records clock64 at the stations of pivot_row for its n-th theta).  Usage: onchip_trace.py N [T]"""
import ctypes, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels, _engine
N = int(sys.argv[1])
ds = benchdata.make_dataset(N, 1, 2)
theta = torch.as_tensor(benchdata.make_theta(65536, 2), device="cuda")
gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5]))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]
gp._prepare_indices(ds["x"], ds["band"], ds["image"])
for _ in range(2):
    gp.loglikelihood_batch(theta)
torch.cuda.synchronize()
lib = _engine.load()
buf = (ctypes.c_longlong * 8192)()
assert lib.gsn_debug_onchip_trace(buf, 8192) == 0
t = np.array(buf[:], dtype=np.int64)
t0, nt, t1 = t[0], int(t[1]), t[2]
print(f"N={N} nt={nt} T={gp._problem.query(_engine.Q_ONCHIP_WARPS)} theta total {t1 - t0} cycles")
names = ["enter", "built", "row_ok", "gemm_done", "diag_ok", "pre_chol", "diag_pub", "z_pub"]
print("col " + " ".join(f"{n:>9s}" for n in names) + "   | diag_pub - prev diag_pub")
prev = None
for j in range(-1, nt - 1):
    row = t[(j + 1) * 16:(j + 1) * 16 + 8] - t0
    d = "" if prev is None else f"{row[6] - prev:7d}"
    print(f"{j:3d} " + " ".join(f"{int(v):9d}" if v > -t0 else "        -" for v in row) + "   | " + d)
    prev = row[6]
