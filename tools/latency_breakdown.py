"""Where a single likelihood call spends its time: Python layers, C call, kernel."""
import sys, os, time, json, cProfile, pstats
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
N = int(sys.argv[1]) if len(sys.argv) > 1 else 141
ds = benchdata.make_dataset(N, 1, 3); ni = 3
theta = benchdata.make_theta(512, ni)
gp = gausSN.GP(kernels.ExpSquaredKernel([0.3, 25.0]), meanfuncs.UniformMean([0.0]), lensingmodels.ConstantMagnification([20.0, 0.5] * (ni - 1)))
gp.x, gp.y, gp.yerr = ds["x"], ds["y"], ds["yerr"]; gp._prepare_indices(ds["x"], ds["band"], ds["image"])
gp.loglikelihood_batch(theta[:8])
p = gp._problem
n = 2000
t0 = time.perf_counter()
for i in range(n): gp.jointprobability(list(theta[i % 512]))
t_joint = (time.perf_counter() - t0) / n
t0 = time.perf_counter()
for i in range(n): gp.loglikelihood_batch(theta[i % 512:i % 512 + 1])
t_batch = (time.perf_counter() - t0) / n
one = np.ascontiguousarray(theta[:1])
t0 = time.perf_counter()
for i in range(n): p.logl_host(one)
t_c = (time.perf_counter() - t0) / n
th_d = torch.as_tensor(one, device="cuda"); out = torch.empty(1, dtype=torch.float64, device="cuda")
s = torch.cuda.current_stream().cuda_stream
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(200): p.logl_device(th_d.data_ptr(), 1, 5 + 2 * (ni - 2), out.data_ptr(), 0, 0, s)
e1.record(); torch.cuda.synchronize()
t_kernel = e0.elapsed_time(e1) / 200 * 1e-3
print(json.dumps(dict(N=N, jointprobability_us=t_joint * 1e6, loglikelihood_batch_us=t_batch * 1e6, c_call_host_buffers_us=t_c * 1e6,
                      kernel_back_to_back_us=t_kernel * 1e6)))
pr = cProfile.Profile(); pr.enable()
for i in range(500): gp.jointprobability(list(theta[i % 512]))
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
