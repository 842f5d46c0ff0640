"""Throughput of the prediction path (gsn_predict_batch) against the per-sample route it replaces.

The reference's plotting loop (gaussn/utils.py:156-209) calls GP.predict once per posterior sample and band; this
script times `GP.predict_batch` (one launch for all samples) for SLSN-like and quasar-like shapes and, beside it, the
same moments computed sample by sample with torch (cuSOLVER LU, the library route).  Prints one JSON line per shape.

    python tools/predict_bench.py [--samples 200]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=200)
    ap.add_argument("--only", default=None, help="run one shape only (for ncu)")
    ap.add_argument("--no-library-route", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(5)
    shapes = [("slsn_like", 4, 2, 194, 100), ("sweep_512", 1, 2, 512, 256), ("quasar_like", 1, 4, 2048, 512)]
    for name, n_bands, n_images, N, M in shapes:
        if args.only and name != args.only:
            continue
        S = n_bands * n_images
        counts = [N // S] * S
        counts[-1] += N - sum(counts)
        x, band, image = [], [], []
        for b in range(n_bands):
            for i in range(n_images):
                n = counts[b * n_images + i]
                x.append(np.sort(rng.uniform(0, 150, n)) + 10.0 * i)
                band += [f"b{b}"] * n
                image += [f"i{i}"] * n
        x = np.concatenate(x)
        y = 0.5 + 0.3 * np.sin(x / 20.0) + 0.02 * rng.standard_normal(N)
        yerr = rng.uniform(0.01, 0.03, N)
        lp0 = [10.0, 0.9] * (n_images - 1)
        gp = gausSN.GP(kernels.Matern32Kernel([0.3, 25.0]), meanfuncs.UniformMean([0.5]), lensingmodels.ConstantMagnification(lp0))
        gp.x, gp.y, gp.yerr, gp.bands = x, y, yerr, np.array(band)
        gp._prepare_indices(x, np.array(band), np.array(image))
        B = args.samples
        theta = np.column_stack([rng.uniform(0.2, 0.4, B), rng.uniform(15, 35, B), rng.uniform(0.4, 0.6, B)] +
                                [rng.uniform(8, 12, B) if k % 2 == 0 else rng.uniform(0.8, 1.0, B) for k in range(2 * (n_images - 1))])
        th_d = torch.as_tensor(theta, device="cuda")
        x_prime = np.linspace(0, 160, M)
        gp.predict_batch(th_d, x_prime, band_index=0)            # warm-up: kernel image, workspace, item list
        torch.cuda.synchronize()
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            for b in range(n_bands):
                e, v = gp.predict_batch(th_d, x_prime, band_index=b)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        n_rows = int(gp.indices[n_images] - gp.indices[0])
        # algorithmic flops per sample and band: factorisation n^3/3, W = K_UV L^-T n^2 M, W W^T (lower half twice = full
        # symmetric product counted once) n M^2, residual and mean 2 n^2 + 2 n M; Matern32 build 11 flop per element
        flops = n_rows ** 3 / 3 + n_rows ** 2 * M + n_rows * M ** 2 + n_rows ** 2 + 2 * n_rows * M \
            + 11 * ((n_rows + M) * (n_rows + M + 1) / 2)
        tflops = flops * B * n_bands / dt / 1e12
        if args.no_library_route:
            print(json.dumps({"shape": name, "rows_per_band": n_rows, "M": M, "samples": B, "ms_per_figure": round(dt * 1e3, 3),
                              "tflops": round(tflops, 2)}))
            continue
        # the library route, per sample: three covariance blocks + LU solve (what GP.predict did before the fused kernel)
        k = min(B, 20)
        xs_d = torch.as_tensor(x[:n_rows], device="cuda")
        xp_d = torch.as_tensor(x_prime, device="cuda")

        def m32(a, b_, A, l):
            r = torch.sqrt(3.0 * (a[:, None] - b_[None, :]) ** 2) / l
            return A * (1 + r) * torch.exp(-r)

        torch.linalg.solve(torch.eye(4, dtype=torch.float64, device="cuda"), torch.ones(4, 1, dtype=torch.float64, device="cuda"))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(k):
            A, l, c = theta[i, :3]
            Kvv = m32(xs_d, xs_d, A, l) + torch.diag(torch.as_tensor(yerr[:n_rows] ** 2, device="cuda"))
            Kuv = m32(xp_d, xs_d, A, l)
            Kuu = m32(xp_d, xp_d, A, l)
            rhs = torch.cat([(torch.as_tensor(y[:n_rows], device="cuda") - c)[:, None], Kuv.T], dim=1)
            sol = torch.linalg.solve(Kvv, rhs)
            _e = c + Kuv @ sol[:, 0]
            _v = Kuu - Kuv @ sol[:, 1:]
        torch.cuda.synchronize()
        dt_lib = (time.perf_counter() - t0) / k * B * n_bands
        print(json.dumps({"shape": name, "n_bands": n_bands, "n_images": n_images, "N": N, "rows_per_band": n_rows, "M": M,
                          "samples": B, "predict_batch_ms_per_figure": round(dt * 1e3, 3),
                          "samples_x_bands_per_s": round(B * n_bands / dt, 1),
                          "tflops_algorithmic": round(tflops, 2), "frac_of_fp64_peak_37.165": round(tflops / 37.165, 3),
                          "torch_per_sample_ms_per_figure_est": round(dt_lib * 1e3, 1), "speedup": round(dt_lib / dt, 1)}))


if __name__ == "__main__":
    main()
