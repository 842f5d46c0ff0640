"""Synthetic light curves and theta batches for bench.py and the full-size tests.

Follows SURVEY.md section 8(d): S = n_bands * n_images segments of N // S epochs
(remainder on the last), epochs sorted uniform on [0, T] per segment with image
i > 1 offset by its true delay, fluxes drawn once from the model's own GP prior
(random Fourier features) at a "truth" theta and scaled to unit range, errors uniform(0.005, 0.03), and
theta batches uniform in the notebooks' prior boxes
(ipynb/fitting_SLSN_example.ipynb cell "ptform": A in [0,1], tau in [10,40],
delta in [0,60], beta in [0.1,1]).  Seeds: data 20240530, theta 20240531.

This is workload generation (plain numpy), not a likelihood implementation.
"""
from __future__ import annotations

import numpy as np

DATA_SEED = 20240530
THETA_SEED = 20240531

# nominal build cost per lower-triangle element, exp/sqrt/sin/div counted as one flop
BUILD_FLOPS = {"ExpSquaredKernel": 7, "OUKernel": 8, "ExponentialKernel": 8, "Matern32Kernel": 11, "Matern52Kernel": 15,
               "PeriodicKernel": 10, "GibbsKernel": 14}


def algorithmic_flops(N: int, n_bands: int, masked: bool, kernel: str = "ExpSquaredKernel") -> float:
    """F(N) = sum over blocks [n^3/3 + n^2 + c_k n(n+1)/2]; blocks are the bands when a
    band mask is present, otherwise one block of N.  Equal-sized
    bands are assumed for the masked case."""
    ck = BUILD_FLOPS.get(kernel, 7)
    sizes = [N // n_bands + (1 if b < N % n_bands else 0) for b in range(n_bands)] if masked else [N]
    return float(sum(n**3 / 3.0 + n**2 + ck * n * (n + 1) / 2.0 for n in sizes))


def make_dataset(N: int, n_bands: int = 1, n_images: int = 2, span: float = 150.0, seed: int = DATA_SEED,
                 truth_A: float = 0.35, truth_tau: float = 25.0, truth_c: float = 0.0):
    """Returns dict(x, y, yerr, band, image, indices, truth) sorted by (band, image, time)."""
    rng = np.random.default_rng(seed)
    S = n_bands * n_images
    counts = [N // S] * S
    counts[-1] += N - sum(counts)
    deltas = [0.0] + [float(d) for d in rng.uniform(8.0, 45.0, n_images - 1)]
    betas = [1.0] + [float(b) for b in rng.uniform(0.3, 0.9, n_images - 1)]
    x, band, image, xs, bvec = [], [], [], [], []
    for b in range(n_bands):
        for i in range(n_images):
            n = counts[b * n_images + i]
            t = np.sort(rng.uniform(0.0, span, n))
            x.append(t + deltas[i])          # observed epochs: the intrinsic curve arrives delta later
            xs.append(t)
            bvec.append(np.full(n, betas[i]))
            band += [f"band_{b}"] * n
            image += [f"image_{i + 1}"] * n
    x, xs, bvec = np.concatenate(x), np.concatenate(xs), np.concatenate(bvec)
    band, image = np.array(band), np.array(image)
    # One draw from (a random-Fourier-feature approximation of) the ExpSquared GP prior at the truth,
    # independent per band, magnified per image.  Only elementwise operations are used, so the data are
    # reproducible across machines to the last ulp or two (a dense eigen-decomposition of the nearly
    # singular prior covariance is not: its null-space vectors depend on the BLAS build).
    M = 512
    y = np.empty(N)
    for b in range(n_bands):
        sel = band == f"band_{b}"
        omega = rng.standard_normal(M) / truth_tau
        phase = rng.uniform(0.0, 2.0 * np.pi, M)
        amp = truth_A * np.sqrt(2.0 / M)
        f = amp * np.cos(np.outer(xs[sel], omega) + phase[None, :]).sum(axis=1)
        y[sel] = bvec[sel] * (truth_c + f)
    yerr = rng.uniform(0.005, 0.03, N)
    scale = y.max() - y.min()
    y = y / scale + yerr * rng.standard_normal(N)
    indices = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    truth = dict(A=truth_A / scale, tau=truth_tau, c=truth_c / scale, deltas=deltas[1:], betas=betas[1:])
    return dict(x=x, y=y, yerr=yerr, band=band, image=image, indices=indices, n_bands=n_bands, n_images=n_images, truth=truth)


def make_theta(B: int, n_images: int = 2, seed: int = THETA_SEED, with_mean: bool = True):
    """[B, P] uniform in the SLSN-notebook prior box: [A, tau, (c), (delta, beta) x (n_images-1)]."""
    rng = np.random.default_rng(seed)
    cols = [rng.uniform(0.0, 1.0, B), rng.uniform(10.0, 40.0, B)]
    if with_mean:
        cols.append(rng.uniform(-0.2, 0.2, B))
    for _ in range(n_images - 1):
        cols.append(rng.uniform(0.0, 60.0, B))
        cols.append(rng.uniform(0.1, 1.0, B))
    return np.ascontiguousarray(np.column_stack(cols))
