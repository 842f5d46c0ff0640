"""Which deviation from LAPACK's arithmetic costs what on the headline workload (N = 512, cond ~ 1e8)?  CPU only.

Emulates the engine's factorisation in numpy float64 -- 8x8 tiles, left-looking, off-diagonal tiles solved by multiplying
with the explicit inverse of the 8x8 pivot tile -- with switches for each deviation, and measures |logL - 80-bit| for each
variant next to LAPACK's (oracle/longdouble.py is the arbiter).  Usage: accuracy_study.py [n_theta] [span]"""
import os, sys, json
import numpy as np
import scipy.linalg
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from oracle import gp_oracle, longdouble

N = 512
n_theta = int(sys.argv[1]) if len(sys.argv) > 1 else 32
span = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
ds = benchdata.make_dataset(N, 1, 2, span=span)
theta = benchdata.make_theta(n_theta, 2)


def cov_and_resid(t, recip):
    A, tau, c, delta, beta = t
    x = ds["x"].copy(); n1 = int(ds["indices"][1])
    xs = x.copy(); xs[n1:] -= delta
    b = np.ones(N); b[n1:] = beta
    d = xs[:, None] - xs[None, :]
    if recip:
        c1 = 1.0 / (2.0 * (tau * tau))
        K = (A * A) * np.exp(-(d * d) * c1)
    else:
        K = A ** 2 * np.exp(-d ** 2 / (2 * tau ** 2))
    C = np.outer(b, b) * K + np.diag(ds["yerr"] ** 2)
    return C, b * c - ds["y"]


def tiled(C, r, inverse_solve=True, refine=False):
    nt = N // 8
    L = np.zeros_like(C)
    W = [None] * nt
    z = np.zeros(N)
    ld = 0.0
    for j in range(nt):
        J = slice(8 * j, 8 * j + 8)
        S = C[J, J] - L[J, :8 * j] @ L[J, :8 * j].T
        Ljj = np.linalg.cholesky(S)
        ld += 2.0 * np.sum(np.log(np.diag(Ljj)))
        Wj = scipy.linalg.solve_triangular(Ljj, np.eye(8), lower=True)
        if refine:
            Wj = Wj + Wj @ (np.eye(8) - Ljj @ Wj)
        L[J, J] = Ljj
        zr = r[J] - L[J, :8 * j] @ z[:8 * j]
        z[J] = Wj @ zr if inverse_solve else scipy.linalg.solve_triangular(Ljj, zr, lower=True)
        if j + 1 < nt:
            R = slice(8 * j + 8, N)
            S2 = C[R, J] - L[R, :8 * j] @ L[J, :8 * j].T
            L[R, J] = S2 @ Wj.T if inverse_solve else scipy.linalg.solve_triangular(Ljj, S2.T, lower=True).T
    return -0.5 * (N * np.log(2 * np.pi) + ld + z @ z)


rows = {k: [] for k in ("lapack", "tiled_inverse", "tiled_inverse_recip", "tiled_trisolve", "tiled_inverse_refined")}
for t in theta:
    truth = longdouble.loglikelihood_ld(ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", list(t[:2]), t[2], list(t[3:]), 1, 2, ds["indices"])
    if not np.isfinite(truth):
        continue
    C, r = cov_and_resid(t, False)
    Cr, _ = cov_and_resid(t, True)
    Lc = np.linalg.cholesky(C)
    zz = scipy.linalg.solve_triangular(Lc, r, lower=True)
    rows["lapack"].append(abs(-0.5 * (N * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(Lc))) + zz @ zz) - truth))
    rows["tiled_inverse"].append(abs(tiled(C, r) - truth))
    rows["tiled_inverse_recip"].append(abs(tiled(Cr, r) - truth))
    rows["tiled_trisolve"].append(abs(tiled(C, r, inverse_solve=False) - truth))
    rows["tiled_inverse_refined"].append(abs(tiled(C, r, refine=True) - truth))
print(json.dumps(dict(N=N, span=span, n_theta=len(rows["lapack"]),
                      **{k: dict(max=float(np.max(v)), median=float(np.median(v))) for k, v in rows.items()})))
