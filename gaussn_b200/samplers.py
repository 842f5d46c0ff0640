"""Batch adapters between third-party samplers and the batched likelihood.

The reference hands ``GP.jointprobability`` to dynesty / emcee / zeus / scipy, which call it with ONE
parameter vector at a time (gaussn/gausSN.py:321-323, :335, :352-359).  These adapters give the same
samplers whole batches to evaluate, which is what the GPU engine is built for:

* ``VectorizedPosterior`` -- callable on ``[n, ndim]`` (emcee / zeus ``vectorize=True``) or on ``[ndim]``;
* ``BatchPool``           -- a pool-like object for dynesty's ``pool=`` / ``queue_size=``.  ``map(f, points)`` with the
                             posterior -- bare, or inside the wrapper objects dynesty builds around it -- evaluates the
                             whole list in one launch (the initial live points).  ``map(evolve_point, args)``, which is
                             how dynesty runs its proposals ('rslice', the reference's default gausSN.py:319, 'slice',
                             'rwalk', 'unif'), runs the ``queue_size`` walks on threads in LOCK STEP: every walk's next
                             likelihood call parks at a rendezvous, and when all live walks have arrived their thetas go
                             to the GPU as one batch.  Each walk still sees a plain scalar function;
* ``fd_value_and_grad``   -- objective value and forward-difference gradient from one batch of P + 1 points,
                             for ``scipy.optimize.minimize(..., jac=True)`` (the reference lets SciPy difference
                             the scalar function: P + 1 separate calls per gradient).
"""
from __future__ import annotations

import numpy as np


class VectorizedPosterior:
    """``invert * (logL + log prior)`` for one or many parameter vectors (gaussn/gausSN.py:162-206)."""

    def __init__(self, gp, logprior=None, fix_kernel_params=False, fix_mean_params=False, fix_lensing_params=False, invert=1):
        self.gp = gp
        self.logprior = logprior
        self.fix = (fix_kernel_params, fix_mean_params, fix_lensing_params)
        self.invert = invert
        self.n_calls = 0          # batched launches
        self.n_evals = 0          # parameter vectors evaluated
        import threading
        self._scalar_hook = threading.local()   # set by BatchPool on its walk threads: scalar calls join a rendezvous
        self._lock = threading.Lock()            # the GP handle takes one batch call at a time

    def __call__(self, theta, *args, **kwargs):
        theta = np.asarray(theta, dtype=np.float64)
        single = theta.ndim == 1
        rdv = getattr(self._scalar_hook, "rdv", None)
        if single and rdv is not None:
            return rdv.call(theta)
        batch = theta.reshape(1, -1) if single else theta
        with self._lock:
            return self._evaluate(batch, single)

    def _evaluate(self, batch, single):
        out = np.asarray(self.gp.jointprobability_batch(batch, self.logprior, *self.fix, invert=self.invert))
        self.n_calls += 1
        self.n_evals += len(batch)
        return float(out[0]) if single else out


class _Rendezvous:
    """Collects the single-theta calls of concurrently running walks into batches.

    ``n_active`` walks are running; a walk that needs a likelihood value parks its theta and sleeps on an event of its own.
    The walk whose arrival (or whose end) leaves no walk computing evaluates everything that is parked in ONE call, hands the
    values back and wakes the others one by one (no condition variable: 64 threads contending for one lock after a
    ``notify_all`` cost 5 ms per round, per-walk events about 1 ms).  The GPU call is a C function: ctypes releases the GIL."""

    def __init__(self, evaluate, n_active):
        import threading
        self.evaluate = evaluate
        self.lock = threading.Lock()
        self.local = threading.local()
        self.n_active = n_active
        self.pending = []          # [theta, value, exception, event]

    def _take_if_complete(self):
        # with the lock held: every live walk is parked -> the batch is ours
        if self.pending and len(self.pending) >= self.n_active:
            batch, self.pending = self.pending, []
            return batch
        return None

    def _flush(self, batch, own=None):
        try:
            vals = self.evaluate(np.stack([e[0] for e in batch]))
            for e, v in zip(batch, vals):
                e[1] = float(v)
        except BaseException as exc:      # noqa: BLE001 -- every parked walk must be released
            for e in batch:
                e[2] = exc
        for e in batch:
            if e is not own:
                e[3].set()

    def call(self, theta):
        import threading
        ev = getattr(self.local, "event", None)
        if ev is None:
            ev = self.local.event = threading.Event()
        ev.clear()
        entry = [np.asarray(theta, dtype=np.float64), None, None, ev]
        with self.lock:
            self.pending.append(entry)
            batch = self._take_if_complete()
        if batch is not None:
            self._flush(batch, own=entry)
        else:
            ev.wait()
        if entry[2] is not None:
            raise entry[2]
        return entry[1]

    def walk_finished(self):
        with self.lock:
            self.n_active -= 1
            batch = self._take_if_complete()
        if batch is not None:
            self._flush(batch)


def _find_posterior(func, depth=0):
    """The VectorizedPosterior inside ``func``: bare, carried as ``.vectorized``, or wrapped the way dynesty wraps the
    user's likelihood (dynesty/dynesty.py: ``_function_wrapper`` keeps it as ``.func`` with ``.args`` / ``.kwargs``; from 2.0
    on ``LogLikelihood`` keeps that wrapper as ``.loglikelihood``).  Wrappers that add arguments are not unwrapped."""
    if isinstance(func, VectorizedPosterior):
        return func
    if depth > 4:
        return None
    vec = getattr(func, "vectorized", None)
    if isinstance(vec, VectorizedPosterior):
        return vec
    for attr in ("loglikelihood", "func"):
        inner = getattr(func, attr, None)
        if inner is not None and inner is not func and not (getattr(func, "args", None) or getattr(func, "kwargs", None)):
            found = _find_posterior(inner, depth + 1)
            if found is not None:
                return found
    return None


class BatchPool:
    """Pool-like adapter for dynesty (``pool=BatchPool(size)``, ``queue_size=size``).

    * ``map(f, points)`` where ``f`` is, or wraps, a ``VectorizedPosterior``: all points in one batched launch.
    * ``map(g, tasks)`` for anything else, with a posterior registered (``BatchPool(size, posterior=...)`` or
      ``pool.posterior = ...``; ``GP.optimize_parameters`` does it): the tasks run on threads and every scalar call they
      make to the posterior is collected with the other tasks' calls into one launch (see ``_Rendezvous``).  This is the
      proposal step of nested sampling -- dynesty/sampler.py ``_fill_queue`` maps ``evolve_point`` over ``queue_size``
      argument bundles, and each ``evolve_point`` (dynesty/sampling.py ``sample_rslice`` / ``sample_slice`` / ``sample_rwalk`` /
      ``sample_unif``) calls the likelihood it finds in its bundle one point at a time.
    * without a posterior: plain serial map.
    ``size`` is what dynesty reads as the default ``queue_size``.  ``batch_sizes`` records the number of thetas of every
    launch made through the pool.

    ``lockstep``: True (default), False (serial map of the proposals) or "auto" (the first proposal round runs serially and
    is timed; lock step from then on if a call took longer than ``lockstep_min_call_s``).  Waking a parked walk costs
    10-20 us of Python, which a round of walks must earn back: measured with the fake rslice driver, nlive = 500
    (profiles/r02_sampler_rate.jsonl), queue_size 250 / 64 against scalar calls: N = 141 24 k / 19 k against 11-14 k calls/s,
    N = 512 23 k / 14 k against 4.5 k, N = 1024 16 k / 8.6 k against 2.0 k, N = 2048 5.6 k / 2.5 k against 0.34 k."""

    def __init__(self, size=512, posterior=None, lockstep=True, lockstep_min_call_s=100e-6):
        self.size = int(size)
        self.posterior = posterior
        self.batch_sizes = []
        self.lockstep = lockstep
        self.lockstep_min_call_s = float(lockstep_min_call_s)
        self._executor = None

    def _evaluate(self, vec, thetas):
        self.batch_sizes.append(len(thetas))
        return np.asarray(vec(thetas))

    def map(self, func, iterable, chunksize=None):
        items = list(iterable)
        if not items:
            return []
        vec = _find_posterior(func)
        if vec is not None and all(np.ndim(it) == 1 for it in items):
            return list(self._evaluate(vec, np.stack([np.asarray(it, dtype=np.float64) for it in items])))
        vec = self.posterior
        if vec is None or len(items) == 1 or self.lockstep is False:
            return [func(it) for it in items]
        if self.lockstep == "auto":
            import time
            n0, t0 = vec.n_evals, time.perf_counter()
            out = [func(it) for it in items]
            dt, n = time.perf_counter() - t0, vec.n_evals - n0
            if n >= 8:
                self.lockstep = (dt / n) > self.lockstep_min_call_s
            return out
        return self._map_lockstep(vec, func, items)

    def _map_lockstep(self, vec, func, items):
        from concurrent.futures import ThreadPoolExecutor
        if self._executor is None or self._executor._max_workers < len(items):
            if self._executor is not None:
                self._executor.shutdown(wait=True)
            self._executor = ThreadPoolExecutor(max_workers=max(len(items), self.size), thread_name_prefix="gsn-walk")
        rdv = _Rendezvous(lambda th: self._evaluate(vec, th), len(items))

        def run(item):
            vec._scalar_hook.rdv = rdv
            try:
                return func(item)
            finally:
                vec._scalar_hook.rdv = None
                rdv.walk_finished()

        return list(self._executor.map(run, items))

    def close(self):
        if self._executor is not None:
            self._executor.shutdown(wait=True)
            self._executor = None

    def join(self):
        pass


def fd_value_and_grad(posterior: VectorizedPosterior, theta, rel_step=None, bounds=None):
    """f(theta) and a forward-difference gradient from ONE batch of P + 1 evaluations.

    Steps follow SciPy's 2-point rule, h_i = sqrt(eps) * max(1, |theta_i|) with the sign of theta_i, flipped
    where the step would leave ``bounds`` (a sequence of (lo, hi))."""
    theta = np.asarray(theta, dtype=np.float64)
    P = len(theta)
    rel = np.sqrt(np.finfo(np.float64).eps) if rel_step is None else rel_step
    h = rel * np.where(theta >= 0, 1.0, -1.0) * np.maximum(1.0, np.abs(theta))
    if bounds is not None:
        for i, (lo, hi) in enumerate(bounds):
            if hi is not None and theta[i] + h[i] > hi:
                h[i] = -abs(h[i])
            if lo is not None and theta[i] + h[i] < lo:
                h[i] = abs(h[i])
    pts = np.repeat(theta[None, :], P + 1, axis=0)
    pts[np.arange(1, P + 1), np.arange(P)] += h
    h = pts[np.arange(1, P + 1), np.arange(P)] - theta            # the step actually taken
    vals = posterior(pts)
    return float(vals[0]), (vals[1:] - vals[0]) / h
