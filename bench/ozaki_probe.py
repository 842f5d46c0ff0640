"""Ozaki-style error-free splitting for the factorisation's trailing updates -- accuracy side of the spike (CPU, numpy).

S = K - L L^T is what the left-looking factorisation accumulates.  Each row of L is scaled by a power of two and cut into
`s` slices of `bits` bits (int8 operands); slice products are exact in int32 / int64 and the pairs (p, q) with p + q < s
are recombined in float64.  Question: how many slices make the recombined product as accurate as the FP64 product, on the
headline workload (N = 512, cond ~ 1e8), measured where it matters -- in logL against the 80-bit arbiter?"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import benchdata
from oracle import longdouble

N, BITS = 512, 6   # 6 magnitude bits per slice: |slice| <= 63, products of 512-long rows stay far inside int32


def split(A, s):
    """Row-scaled slices: A ~ 2^e[:,None] * sum_p slices[p] * 2^(-BITS (p+1))."""
    e = np.ceil(np.log2(np.maximum(np.abs(A).max(axis=1), 1e-300)))
    R = A / np.exp2(e)[:, None]
    slices = []
    for _ in range(s):
        R = R * (1 << BITS)
        q = np.trunc(R)
        slices.append(q.astype(np.int64))
        R = R - q
    return e, slices


def sliced_aat(A, B, s):
    """A B^T through slice products (exact integer arithmetic), recombined in float64, highest-order terms last."""
    ea, sa = split(A, s)
    eb, sb = split(B, s)
    out = np.zeros((A.shape[0], B.shape[0]))
    for order in range(2 * s - 2, -1, -1):
        if order >= s:
            continue                      # pairs below the target precision are dropped: s (s + 1) / 2 products
        acc = np.zeros(out.shape, dtype=np.int64)
        for p in range(order + 1):
            q = order - p
            if p < s and q < s:
                acc += sa[p] @ sb[q].T
        out += acc.astype(np.float64) * 2.0 ** (-BITS * (order + 2))
    return out * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def chol_logl(C, r, s):
    """Blocked left-looking Cholesky (32-row blocks) whose trailing updates use sliced products (s = 0: plain FP64)."""
    n = len(r); nb = 32
    L = np.zeros_like(C)
    for j in range(0, n, nb):
        J = slice(j, j + nb)
        upd = (L[j:, :j] @ L[J, :j].T) if s == 0 else sliced_aat(L[j:, :j], L[J, :j], s) if j else 0.0
        S = C[j:, J] - upd
        Ljj = np.linalg.cholesky(S[:nb])
        L[J, J] = Ljj
        if j + nb < n:
            import scipy.linalg
            L[j + nb:, J] = scipy.linalg.solve_triangular(Ljj, S[nb:].T, lower=True).T
    import scipy.linalg
    z = scipy.linalg.solve_triangular(L, r, lower=True)
    return -0.5 * (n * np.log(2 * np.pi) + 2 * np.sum(np.log(np.diag(L))) + z @ z)


ds = benchdata.make_dataset(N, 1, 2)
theta = benchdata.make_theta(12, 2)
err = {s: [] for s in (0, 7, 8, 9, 10, 11)}
for t in theta:
    A, tau, c, delta, beta = t
    n1 = int(ds["indices"][1])
    xs = ds["x"].copy(); xs[n1:] -= delta
    b = np.ones(N); b[n1:] = beta
    d = xs[:, None] - xs[None, :]
    C = np.outer(b, b) * (A ** 2 * np.exp(-d ** 2 / (2 * tau ** 2))) + np.diag(ds["yerr"] ** 2)
    r = b * c - ds["y"]
    truth = longdouble.loglikelihood_ld(ds["x"], ds["y"], ds["yerr"], "ExpSquaredKernel", [A, tau], c, [delta, beta], 1, 2, ds["indices"])
    if not np.isfinite(truth):
        continue
    for s in err:
        try:
            err[s].append(abs(chol_logl(C, r, s) - truth))
        except np.linalg.LinAlgError:
            err[s].append(np.inf)
print(json.dumps(dict(N=N, bits_per_slice=BITS, n_theta=len(err[0]),
                      max_abs_err_vs_80bit={("fp64" if s == 0 else f"{s}_slices_{s * (s + 1) // 2}_products"): float(np.max(v)) for s, v in err.items()},
                      median_abs_err_vs_80bit={("fp64" if s == 0 else f"{s}_slices"): float(np.median(v)) for s, v in err.items()})))
