"""One small call per launch shape, for compute-sanitizer (memcheck / racecheck / synccheck, one tool per run):
on-chip shape as one warp per theta, as teams of 4 and 16, with several band blocks; workspace kernels as one warp, four
warps and clusters of 2 / 4 / 8 CTAs; prediction (one CTA and cluster); the gather entry point; the host-mean path.
Every result is compared with the oracle, so a run that passes under the sanitizer also computed the right numbers."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import gp_oracle
from test_gpu_parity import _synth, _theta_box, _gp_for, _oracle_batch
from gaussn_b200 import _engine

rng = np.random.default_rng(2)
only = sys.argv[1:]


def case(name, N, bands, images, B, kernel="ExpSquaredKernel", env=None, check=8):
    if only and not any(o in name for o in only):
        return
    for k, v in (env or {}).items():
        os.environ[k] = v
    ds = _synth(rng, N, bands, images)
    theta = _theta_box(rng, B, kernel, "UniformMean", "ConstantMagnification", images)
    theta[min(3, B - 1), 1] = np.nan
    gp = _gp_for(ds, kernel, "UniformMean", "ConstantMagnification", images, theta[0])
    got = gp.loglikelihood_batch(theta)
    want = _oracle_batch(ds, theta[:check], kernel, "UniformMean", "ConstantMagnification", gp)
    ok = np.array_equal(np.isnan(got[:check]), np.isnan(want)) and np.nanmax(np.abs(got[:check] - want)) <= 1e-8 + 1e-12 * np.nanmax(np.abs(want))
    print(f"{name}: path {gp._problem.query(_engine.Q_PATH)}, warps/theta {gp._problem.query(_engine.Q_ONCHIP_WARPS)}, "
          f"{B} theta, N={N}: {'ok' if ok else 'MISMATCH'}", flush=True)
    for k in (env or {}):
        os.environ.pop(k)
    assert ok
    return gp, ds, theta


case("onchip one warp", 40, 1, 2, 64)
case("onchip 4 bands", 194, 4, 2, 48, kernel="Matern32Kernel")
case("onchip team of 4", 141, 1, 3, 16, env={"GSN_ONCHIP_T": "4"})
case("onchip team of 16", 180, 1, 2, 4, env={"GSN_ONCHIP_T": "16"})
case("workspace one warp", 100, 1, 2, 520, env={"GSN_FORCE_SHAPE": "legacy"})
case("workspace four warps", 400, 2, 2, 6)
for c in (2, 4, 8):
    case(f"workspace cluster of {c}", 300, 1, 2, 2, env={"GSN_FORCE_SHAPE": f"cluster{c}"}, check=2)
if not only or any("predict" in o for o in only):
    out = case("predict (setup)", 90, 2, 2, 6)
    gp, ds, theta = out
    gp.bands = ds["band"]
    x_prime = np.linspace(0, 120, 40)
    e, v = gp.predict_batch(theta[:3], x_prime, band_index=1)
    os.environ["GSN_FORCE_SHAPE"] = "cluster2"
    e2, v2 = gp.predict_batch(theta[:3], x_prime, band_index=1)
    os.environ.pop("GSN_FORCE_SHAPE")
    assert np.array_equal(e[:3][np.isfinite(e[:3])], e2[:3][np.isfinite(e2[:3])])
    print("predict: one CTA and cluster of 2 agree", flush=True)
if not only or any("gather" in o for o in only):
    import torch
    out = case("gather (setup)", 141, 1, 2, 32)
    gp, ds, theta = out
    th = torch.as_tensor(theta, device="cuda")
    a = torch.full((64,), -1.0, dtype=torch.float64, device="cuda")
    b = torch.full((64,), -1.0, dtype=torch.float64, device="cuda")
    gp._problem.logl_device_gather(th.data_ptr(), 32, th.shape[1], [a.data_ptr(), b.data_ptr()], 16, 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = gp.loglikelihood_batch(theta)
    assert np.array_equal(a.cpu().numpy()[16:48], ref, equal_nan=True) and np.array_equal(b.cpu().numpy()[16:48], ref, equal_nan=True)
    assert float(a[:16].max()) == -1.0 and float(b[48:].max()) == -1.0
    print("gather: two output buffers at an offset agree", flush=True)
print("all shapes done")
