"""Protocol fakes of the third-party samplers the reference drives (gaussn/gausSN.py:316-360).  TEST INFRASTRUCTURE.

dynesty, emcee and zeus are not installed here and cannot be fetched, so the adapters in
gaussn_b200/samplers.py are exercised against minimal stand-ins that reproduce HOW each library calls the user's
likelihood and the user's pool -- not what it computes with the values.  The call shapes are restated from the libraries'
public sources (dynesty 2.1, emcee 3.1, zeus 2.5), quoted by file and function since line numbers cannot be checked offline:

dynesty
  * dynesty/dynesty.py `_function_wrapper.__call__`: the user's loglikelihood / prior_transform are stored as `.func` with
    `.args` / `.kwargs` and called as `func(x, *args, **kwargs)`;
  * dynesty/dynesty.py `NestedSampler.__new__`: the initial live points are evaluated with
    `loglikelihood.map(live_v)`, i.e. (dynesty/utils.py `LogLikelihood.map`) `pool.map(self.loglikelihood, pars)` when a pool
    is given and `use_pool['loglikelihood']` is true (the default);
  * dynesty/sampler.py `Sampler._fill_queue`: builds `queue_size` argument bundles (point, loglstar, axes, scale,
    prior_transform, loglikelihood, seed, kwargs) and runs `self.mapper(self.evolve_point, args)`, where `mapper` is
    `pool.map` when `use_pool['propose_point']` is true (the default);
  * dynesty/sampling.py `sample_rslice` / `sample_slice` / `sample_rwalk`: a sequential walk that calls
    `loglikelihood(prior_transform(u))` one point at a time until it has an accepted point; `sample_unif` calls it once.
emcee / zeus
  * emcee/ensemble.py `EnsembleSampler.compute_log_prob`: with `vectorize=True` the whole (half-)ensemble goes to
    `log_prob_fn(coords)` as one [n, ndim] array; otherwise `pool.map(log_prob_fn, coords)` / `map`;
  * zeus/ensemble.py `EnsembleSampler.compute_log_prob`: the same switch.
"""
import numpy as np


class FunctionWrapper:
    """dynesty/dynesty.py `_function_wrapper`."""

    def __init__(self, func, args=(), kwargs=None, name="input"):
        self.func, self.args, self.kwargs, self.name = func, tuple(args), dict(kwargs or {}), name

    def __call__(self, x):
        return self.func(np.asarray(x).copy(), *self.args, **self.kwargs)


class LogLikelihood:
    """dynesty/utils.py `LogLikelihood`: wraps the wrapped likelihood once more and owns the pool for `.map`."""

    def __init__(self, loglikelihood, ndim, pool=None):
        self.loglikelihood, self.ndim, self.pool = loglikelihood, ndim, pool

    def map(self, pars):
        if self.pool is None:
            return [self.loglikelihood(p) for p in pars]
        return list(self.pool.map(self.loglikelihood, pars))

    def __call__(self, x):
        return self.loglikelihood(x)


def evolve_point_rslice(args):
    """Stand-in for dynesty/sampling.py `sample_rslice`: a random-direction slice walk in the unit cube, a few
    slices per proposal, stepping out and shrinking, the likelihood called one point at a time.  Returns (u, v, logl, ncalls)."""
    u, loglstar, scale, prior_transform, loglikelihood, seed, slices = args
    rng = np.random.default_rng(seed)
    n = len(u)
    nc = 0

    def f(uu):
        nonlocal nc
        if np.any(uu <= 0) or np.any(uu >= 1):
            return -np.inf
        nc += 1
        return loglikelihood(prior_transform(uu))

    logl = None
    for _ in range(slices):
        d = rng.standard_normal(n)
        d *= scale / np.linalg.norm(d)
        r = rng.uniform()
        lo, hi = -r, 1.0 - r
        while f(u + lo * d) > loglstar:          # stepping out
            lo -= 1.0
        while f(u + hi * d) > loglstar:
            hi += 1.0
        while True:                              # shrinking
            t = rng.uniform(lo, hi)
            logl = f(u + t * d)
            if logl > loglstar:
                u = u + t * d
                break
            if t < 0:
                lo = t
            else:
                hi = t
            if hi - lo < 1e-12:
                logl = None
                break
    if logl is None:
        logl = f(u)
    return u, prior_transform(u), logl, nc


class FakeNestedSampler:
    """The part of dynesty's static nested sampler that touches the user's callables: initial live points through
    `LogLikelihood.map`, then rounds of `mapper(evolve_point, queue_size bundles)` whose results replace the worst points."""

    def __init__(self, loglikelihood, prior_transform, ndim, nlive=64, pool=None, queue_size=None, logl_args=(), seed=0, slices=3):
        self.rng = np.random.default_rng(seed)
        self.ptform = FunctionWrapper(prior_transform, name="prior_transform")
        self.loglike = LogLikelihood(FunctionWrapper(loglikelihood, logl_args, name="loglikelihood"), ndim, pool)
        self.pool = pool
        self.mapper = pool.map if pool is not None else (lambda f, it: list(map(f, it)))
        self.queue_size = queue_size or (getattr(pool, "size", 1) if pool is not None else 1)
        self.slices = slices
        self.live_u = self.rng.uniform(size=(nlive, ndim))
        self.live_v = np.array([self.ptform(u) for u in self.live_u])
        self.live_logl = np.array(self.loglike.map(self.live_v), dtype=float)       # NestedSampler.__new__
        self.ncall = nlive

    def run_nested(self, rounds=4):
        for _ in range(rounds):
            worst = np.argsort(self.live_logl)[:self.queue_size]
            loglstar = self.live_logl[worst[-1]]
            starts = self.rng.integers(0, len(self.live_u), size=len(worst))
            args = [(self.live_u[s].copy(), loglstar, 0.2, self.ptform, self.loglike, int(self.rng.integers(1 << 31)), self.slices)
                    for s in starts]
            out = self.mapper(evolve_point_rslice, args)                            # Sampler._fill_queue
            for w, (u, v, logl, nc) in zip(worst, out):
                self.ncall += nc
                if logl > self.live_logl[w]:
                    self.live_u[w], self.live_v[w], self.live_logl[w] = u, v, logl
        return self


class FakeEnsembleSampler:
    """emcee / zeus `EnsembleSampler.compute_log_prob`: vectorize=True hands the ensemble over as one array."""

    def __init__(self, nwalkers, ndim, log_prob_fn, vectorize=False, pool=None):
        self.nwalkers, self.ndim, self.log_prob_fn, self.vectorize, self.pool = nwalkers, ndim, log_prob_fn, vectorize, pool

    def compute_log_prob(self, coords):
        if self.vectorize:
            return np.asarray(self.log_prob_fn(coords))
        mapper = self.pool.map if self.pool is not None else map
        return np.array(list(mapper(self.log_prob_fn, (c for c in coords))))

    def run_mcmc(self, p0, nsteps, rng=None):
        rng = rng or np.random.default_rng(0)
        p, lp = np.array(p0), self.compute_log_prob(p0)
        half = self.nwalkers // 2
        for _ in range(nsteps):
            for first in (True, False):                                             # stretch move on half the ensemble
                S = slice(0, half) if first else slice(half, None)
                C = slice(half, None) if first else slice(0, half)
                z = ((2.0 - 1.0) * rng.uniform(size=p[S].shape[0]) + 1) ** 2 / 2.0
                partner = p[C][rng.integers(0, p[C].shape[0], size=p[S].shape[0])]
                q = partner + z[:, None] * (p[S] - partner)
                lq = self.compute_log_prob(q)
                acc = np.log(rng.uniform(size=len(z))) < (self.ndim - 1) * np.log(z) + lq - lp[S]
                p[S] = np.where(acc[:, None], q, p[S])
                lp[S] = np.where(acc, lq, lp[S])
        return p, lp
