"""``GP`` -- drop-in for ``gaussn/gausSN.py`` backed by the CUDA engine (libgsn.so).

Same constructor and method signatures as the reference class
(``gaussn/gausSN.py:25-403``): ``loglikelihood``, ``jointprobability``,
``optimize_parameters``, ``predict``, ``_prepare_indices`` ...  What is new is the
batch dimension the reference lacks: ``loglikelihood_batch`` and
``jointprobability_batch`` evaluate many parameter vectors in one kernel launch,
and the scalar methods are a batch of one.  There is no CPU implementation of
the likelihood in this package; a missing library or GPU raises.
"""
from __future__ import annotations

import numpy as np

from . import _engine, lensingmodels, samplers

try:  # optional samplers, exactly as the reference treats them (gausSN.py:6-21)
    from scipy.optimize import minimize
except ImportError:  # pragma: no cover
    minimize = None
try:
    import dynesty
except ImportError:
    dynesty = None
try:
    import emcee
except ImportError:
    emcee = None
try:
    import zeus
except ImportError:
    zeus = None


_PREDICT_HANDLES = {}   # (N, kernel kind, mean kind, n_mean, device) -> [Problem, data key]; see GP.predict


def _is_torch_cuda(t) -> bool:
    return type(t).__module__.startswith("torch") and getattr(t, "is_cuda", False)


class GP:
    """Gaussian process for (lensed) light curves; see the reference docstring
    (``gaussn/gausSN.py:25-55``).  ``device`` selects the GPU (default: LOCAL_RANK or 0)."""

    def __init__(self, kernel, meanfunc, lensingmodel=None, device=None):
        self.kernel = kernel
        self.meanfunc = meanfunc
        self.lensingmodel = lensingmodel if lensingmodel is not None else lensingmodels.NoLensing()
        self.device = device
        self._problem = None
        self._problem_key = None
        self._map_key = None
        self.bands, self.zp, self.zpsys = None, 27.5, "ab"
        for plugin, what in ((kernel, "kernel"), (meanfunc, "mean function"), (self.lensingmodel, "lensing model")):
            if not hasattr(plugin, "kind"):
                raise TypeError(f"{type(plugin).__name__} is not a gaussn_b200 {what}: only plugins with a device "
                                "implementation can be used (there is no CPU fallback)")
            if device is not None and getattr(plugin, "device", None) is None:
                plugin.device = device      # the plugins' own covariance / mean / lens calls run on the GP's GPU

    # ------------------------------------------------------------------ bookkeeping
    def _prepare_indices(self, x, band, image):
        """Cumulative per-(band, image) counts, band-major / image-minor in
        ``np.unique`` order; rows must already be sorted (gaussn/gausSN.py:74-101)."""
        self.n_bands = len(np.unique(band)) if band is not None else 1
        self.n_images = len(np.unique(image)) if image is not None else 1
        indices = [0]
        if band is not None:
            band = np.asarray(band)
            image_arr = None if image is None else np.asarray(image)
            for pb_id in np.unique(band):
                sel = band == pb_id
                if image_arr is not None:
                    for im_id in np.unique(image_arr):
                        indices.append(int(np.sum(image_arr[sel] == im_id)) + indices[-1])
                else:
                    indices.append(int(np.sum(sel)) + indices[-1])
        elif image is not None:
            image_arr = np.asarray(image)
            for im_id in np.unique(image_arr):
                indices.append(int(np.sum(image_arr == im_id)) + indices[-1])
        else:
            indices.append(len(x) + indices[-1])
        self.indices = np.asarray(indices, dtype=np.int64)
        self.factor = len(x) * np.log(2 * np.pi)

    def _get_initial_pos(self, fix_kernel_params, fix_mean_params, fix_lensing_params):
        """Initial vector in the packed order kernel | mean | lensing.  (The
        reference defines this helper with its first two arguments swapped,
        gausSN.py:103 vs :329; the order used here is the one ``jointprobability``
        slices by.)"""
        init_pos = []
        if not fix_kernel_params:
            init_pos.extend(self.kernel.params)
        if not fix_mean_params:
            init_pos.extend(self.meanfunc.params)
        if not fix_lensing_params:
            init_pos.extend(self.lensingmodel.params)
        if len(init_pos) < 0.5:
            raise Exception("No parameters to fit. Check fix_mean_params and fix_kernel_params to make sure there are parameters to fit.")
        return init_pos

    def _rescale_data(self, y, yerr):
        """Rescale so that y spans one unit (gaussn/gausSN.py:120-127)."""
        y, yerr = np.asarray(y, dtype=float), np.asarray(yerr, dtype=float)
        factor = np.max(y) - np.min(y)
        return y / factor, yerr / factor

    def logprior(self, params):
        return 0

    # ------------------------------------------------------------------ engine plumbing
    def _n_lens_params(self):
        return len(getattr(self.lensingmodel, "params", []) or [])

    def _ensure_problem(self, x, y, yerr):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).ravel()
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).ravel()
        yerr = np.ascontiguousarray(np.asarray(yerr, dtype=np.float64)).ravel()
        if not hasattr(self, "indices") or int(self.indices[-1]) != len(x):
            self._prepare_indices(x, None, None)
        if self.lensingmodel.kind != _engine.L_NONE and getattr(self.lensingmodel, "n_images", None) != self.n_images:
            self.lensingmodel.import_from_gp(self.n_bands, self.n_images, self.indices)
        n_mean = len(self.meanfunc.params)
        key = (x.tobytes(), y.tobytes(), yerr.tobytes(), self.indices.tobytes(), self.n_bands, self.n_images,
               self.kernel.kind, self.meanfunc.kind, n_mean, self.lensingmodel.kind, self.device)
        if self._problem is not None and key != self._problem_key and key[3:] == self._problem_key[3:]:
            self._problem.set_data(x, y, yerr)      # same structure and plugins, new observations: keep the handle
            self._problem_key = key
        if self._problem is None or key != self._problem_key:
            if self._problem is not None:
                self._problem.close()
            self._problem = _engine.Problem(x, y, yerr, self.n_bands, self.n_images, self.indices, self.kernel.kind,
                                            self.meanfunc.kind, self.lensingmodel.kind, n_mean, self.device)
            self._problem_key = key
            self._map_key = None
            expected = self._problem.n_lens
            if self.lensingmodel.kind != _engine.L_NONE and self._n_lens_params() != expected:
                raise IndexError(f"{type(self.lensingmodel).__name__} has {self._n_lens_params()} parameters but "
                                 f"{self.n_images} images need {expected}")
        return self._problem

    def _current_full_theta(self):
        kp = [float(v) for v in getattr(self.kernel, "params", [])][:self._problem.n_kernel]
        mp = [float(v) for v in self.meanfunc.params]
        lp = [float(v) for v in (getattr(self.lensingmodel, "params", []) or [])]
        return np.array(kp + mp + lp, dtype=np.float64)

    def _set_map(self, fix_kernel, fix_mean, fix_lens):
        """Program the device-side theta map for a fix_* combination
        (gaussn/gausSN.py:173-194): free groups consume the caller's columns in the
        order kernel, mean, lensing; fixed groups read the plugins' current values."""
        p = self._problem
        full = self._current_full_theta()
        cols, pos = [], 0
        for fixed, count in ((fix_kernel, p.n_kernel), (fix_mean, p.n_mean), (fix_lens, p.n_lens)):
            if fixed:
                cols += [-1] * count
            else:
                cols += list(range(pos, pos + count))
                pos += count
        key = (tuple(cols), full.tobytes() if -1 in cols else None)
        if key != self._map_key:
            p.set_theta_map(cols, full)
            self._map_key = key
        self._map_cols = cols
        return pos

    def _evaluate(self, theta, flags=0):
        """theta: [B, ncols] numpy array (host path) or torch CUDA tensor (device path)."""
        p = self._problem
        if self.meanfunc.kind == _engine.M_HOST:
            return self._evaluate_hostmean(theta, flags)
        if _is_torch_cuda(theta):
            import torch
            th = theta.to(torch.float64).contiguous()
            if th.dim() == 1:
                th = th.reshape(1, -1)
            if th.device.index != p.device:
                raise ValueError(f"theta lives on cuda:{th.device.index} but the problem is on cuda:{p.device}")
            out = torch.empty(th.shape[0], dtype=torch.float64, device=th.device)
            p.logl_device(th.data_ptr(), th.shape[0], th.shape[1], out.data_ptr(), 0, flags,
                          torch.cuda.current_stream(th.device).cuda_stream)
            return out
        theta = np.asarray(theta, dtype=np.float64)
        if theta.ndim == 1:
            theta = theta.reshape(1, -1)
        return p.logl_host(theta, flags)

    def _evaluate_hostmean(self, theta, flags):
        """sncosmoMean-style means: lens on the device, mean on the host per theta,
        everything else on the device (gsn_logl_batch_hostmean)."""
        import torch
        p = self._problem
        dev = torch.device("cuda", p.device)
        was_torch = _is_torch_cuda(theta)
        th = (theta if was_torch else torch.as_tensor(np.atleast_2d(np.asarray(theta, dtype=np.float64)), device=dev))
        th = th.to(torch.float64).contiguous()
        B = th.shape[0]
        stream = torch.cuda.current_stream(dev).cuda_stream
        shifted = torch.empty((B, p.N), dtype=torch.float64, device=dev)
        bvec = torch.empty((B, p.N), dtype=torch.float64, device=dev)
        p.lens_device(th.data_ptr(), B, th.shape[1], shifted.data_ptr(), bvec.data_ptr(), stream)
        shifted_h = shifted.cpu().numpy()
        th_h = th.cpu().numpy()
        # which caller columns feed the mean slots under the current map
        cols = self._map_cols
        mean_cols = cols[p.n_kernel:p.n_kernel + p.n_mean]
        m = np.empty((B, p.N))
        for i in range(B):
            mp = None if (len(mean_cols) and mean_cols[0] < 0) else [th_h[i, c] for c in mean_cols]
            m[i] = self.meanfunc.mean(shifted_h[i], params=mp, bands=self.bands, zp=self.zp, zpsys=self.zpsys)
        m_d = torch.as_tensor(m, device=dev)
        out = torch.empty(B, dtype=torch.float64, device=dev)
        p.logl_device_hostmean(th.data_ptr(), B, th.shape[1], m_d.data_ptr(), out.data_ptr(), 0, flags, stream)
        return out if was_torch else out.cpu().numpy()

    # ------------------------------------------------------------------ likelihood
    def loglikelihood(self, x, y, yerr, kernel_params, meanfunc_params, lensing_params):
        """log N(y | b m(x~), mask o (b b^T o K(x~)) + diag(yerr^2)) for one parameter
        set (gaussn/gausSN.py:135-160).  ``None`` keeps a plugin's current
        parameters; given parameters are stored on the plugin, as the reference's
        ``_reset`` does.  A covariance that is not positive definite gives NaN."""
        if kernel_params is not None:
            self.kernel._reset(list(kernel_params))
        if meanfunc_params is not None:
            self.meanfunc._reset(list(meanfunc_params))
        if lensing_params is not None and self.lensingmodel.kind != _engine.L_NONE:
            self.lensingmodel._reset(list(lensing_params))
        self._ensure_problem(x, y, yerr)
        self._set_map(True, True, True)
        self._last_call = (np.asarray(x, dtype=np.float64), np.asarray(yerr, dtype=np.float64))   # for the lazy .mean / .cov
        self._mean_cache = self._cov_cache = None
        return float(self._evaluate(np.zeros((1, 0)))[0])

    # The reference leaves ``self.mean`` and ``self.cov`` of the last ``loglikelihood`` call behind (gausSN.py:142, :147).  The
    # engine never materialises either, so they are produced on demand, from the plugins' current parameters and the data of
    # that call, through the plugins' own device evaluations.
    @property
    def mean(self):
        if getattr(self, "_last_call", None) is None:
            raise AttributeError("mean is available after a loglikelihood call")
        if self._mean_cache is None:
            x, _ = self._last_call
            shifted, b = self.lensingmodel.lens(x)
            self._mean_cache = b * np.asarray(self.meanfunc.mean(shifted, bands=self.bands, zp=self.zp, zpsys=self.zpsys))
        return self._mean_cache

    @property
    def cov(self):
        if getattr(self, "_last_call", None) is None:
            raise AttributeError("cov is available after a loglikelihood call")
        if self._cov_cache is None:
            x, yerr = self._last_call
            shifted, b = self.lensingmodel.lens(x)
            K = np.outer(b, b) * np.asarray(self.kernel.covariance(shifted))
            self._cov_cache = np.multiply(self.lensingmodel.mask, K) + np.diag(yerr ** 2)
        return self._cov_cache

    def loglikelihood_batch(self, theta, x=None, y=None, yerr=None, fix_kernel_params=False, fix_mean_params=False,
                            fix_lensing_params=False, neg_inf=False):
        """``loglikelihood`` for every row of ``theta`` (packed kernel | mean |
        lensing, skipping fixed groups) in one launch.  ``theta`` may be a numpy
        array (host buffers; copies happen inside the C call) or a torch CUDA
        tensor (no copies; result is a CUDA tensor).  Defaults to the data set
        given to ``optimize_parameters`` / the last ``loglikelihood`` call."""
        x = self.x if x is None else x
        y = self.y if y is None else y
        yerr = self.yerr if yerr is None else yerr
        self._ensure_problem(x, y, yerr)
        if self.lensingmodel.kind == _engine.L_NONE:
            fix_lensing_params = True
        ncols = self._set_map(fix_kernel_params, fix_mean_params, fix_lensing_params)
        width = theta.shape[-1] if hasattr(theta, "shape") and len(theta.shape) > 0 else len(theta)
        if width != ncols:
            raise ValueError(f"theta has {width} columns, expected {ncols} (kernel | mean | lensing minus fixed groups)")
        return self._evaluate(theta, _engine.FLAG_NEG_INF if neg_inf else 0)

    def loglikelihood_batch_sharded(self, theta, group=None, fused=False, **kwargs):
        """``loglikelihood_batch`` with the rows of ``theta`` split over the ranks of a
        torch.distributed group (one process per GPU) and the results gathered; every
        rank passes the same ``theta`` and receives the full vector.  ``fused=True`` (CUDA tensors
        only) lets the kernel store each result directly into every rank's gathered buffer over
        NVLink instead of issuing an all-gather afterwards."""
        from . import sharding
        if not fused:
            return sharding.evaluate_sharded(lambda th: self.loglikelihood_batch(th, **kwargs), theta, group)
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B = theta.shape[0]
        start, stop, per = sharding.shard_bounds(B, world, rank)
        self.loglikelihood_batch(theta[:0], **kwargs)          # programs the problem handle and the theta map
        key = (per, theta.device, id(group))
        if getattr(self, "_fused_key", None) != key:
            self._fused = sharding.FusedGather(per, theta.device, group)
            self._fused_key = key
        flags = _engine.FLAG_NEG_INF if kwargs.get("neg_inf") else 0
        out = self._fused.run(self._problem, theta[start:stop].contiguous(), flags)
        return out[:B].clone()    # gathered index rank * per + i is the global row index, also for a ragged last shard

    def jointprobability(self, params, logprior=None, fix_kernel_params=False, fix_mean_params=False,
                         fix_lensing_params=False, invert=1):
        """``invert * (logL + log prior)``; NaN / inf become ``invert * -inf``
        (gaussn/gausSN.py:162-206)."""
        logprior = self.logprior if logprior is None else logprior
        log_prior = logprior(params)
        if np.isinf(log_prior) or np.isnan(log_prior):
            return invert * -np.inf
        theta = np.asarray([float(v) for v in params], dtype=np.float64).reshape(1, -1)
        loglike = float(self.loglikelihood_batch(theta, fix_kernel_params=fix_kernel_params, fix_mean_params=fix_mean_params,
                                                 fix_lensing_params=fix_lensing_params)[0])
        loglike += log_prior
        if np.isinf(loglike) or np.isnan(loglike):
            return invert * -np.inf
        return invert * loglike

    def jointprobability_batch(self, thetas, logprior=None, fix_kernel_params=False, fix_mean_params=False,
                               fix_lensing_params=False, invert=1):
        """Vector version of ``jointprobability``: one launch for all rows.
        ``logprior`` is called once per row on the host (it is user code)."""
        is_dev = _is_torch_cuda(thetas)
        ll = self.loglikelihood_batch(thetas, fix_kernel_params=fix_kernel_params, fix_mean_params=fix_mean_params,
                                      fix_lensing_params=fix_lensing_params, neg_inf=True)
        if logprior is None:
            return invert * ll
        host = thetas.detach().cpu().numpy() if is_dev else np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        lp = np.array([float(logprior(row)) for row in host])
        if is_dev:
            import torch
            lp_t = torch.as_tensor(lp, device=ll.device)
            tot = ll + lp_t
            tot = torch.where(torch.isfinite(tot) & torch.isfinite(lp_t), tot, torch.full_like(tot, -float("inf")))
            return invert * tot
        with np.errstate(invalid="ignore"):
            tot = ll + lp
        tot[~(np.isfinite(tot) & np.isfinite(lp))] = -np.inf
        return invert * tot

    # ------------------------------------------------------------------ fitting
    def optimize_parameters(self, x, y, yerr, band=None, image=None, zp=27.5, zpsys='ab', method='minimize', loglikelihood=None,
                            logprior=None, ptform=None, fix_kernel_params=False, fix_mean_params=False, fix_lensing_params=False,
                            init_scale=1., minimize_kwargs=None, sampler_kwargs=None, run_sampler_kwargs=None, rescale_data=False):
        """Same contract as the reference (gaussn/gausSN.py:208-360): ``method`` is
        'minimize' (scipy), 'dynesty', 'emcee' or 'zeus'; returns the scipy result
        or the sampler.  Every likelihood call goes to the GPU."""
        sampler_kwargs = {} if sampler_kwargs is None else dict(sampler_kwargs)
        run_sampler_kwargs = {} if run_sampler_kwargs is None else dict(run_sampler_kwargs)
        minimize_kwargs = {} if minimize_kwargs is None else dict(minimize_kwargs)
        if loglikelihood is not None:
            raise NotImplementedError("a user-supplied loglikelihood callable would bypass the CUDA engine; "
                                      "gaussn_b200 only evaluates the built-in multivariate-normal likelihood")

        self.x = np.asarray(x, dtype=np.float64)
        if rescale_data:
            self.y, self.yerr = self._rescale_data(y, yerr)
        else:
            self.y, self.yerr = np.asarray(y, dtype=np.float64), np.asarray(yerr, dtype=np.float64)
        self.bands, self.zp, self.zpsys = band, zp, zpsys

        self._prepare_indices(self.x, band, image)
        if self.lensingmodel.kind == _engine.L_NONE:
            fix_lensing_params = True          # gausSN.py:287-291
        else:
            self.lensingmodel.import_from_gp(self.n_bands, self.n_images, self.indices)

        self.ndim = 0
        if not fix_kernel_params:
            self.ndim += len(self.kernel.params)
        if not fix_mean_params:
            self.ndim += len(self.meanfunc.params)
        if not fix_lensing_params:
            self.ndim += len(self.lensingmodel.params)
        if logprior is None:
            logprior = self.logprior
        else:
            self.logprior = logprior
        self._ensure_problem(self.x, self.y, self.yerr)
        args = (logprior, fix_kernel_params, fix_mean_params, fix_lensing_params)

        if method == 'dynesty':
            if dynesty is None:
                raise ImportError("dynesty is not installed")
            nlive = sampler_kwargs.pop('nlive', 500)
            sample = sampler_kwargs.pop('sample', 'rslice')
            if isinstance(sampler_kwargs.get('pool'), samplers.BatchPool):
                # Batched evaluation.  The initial live points are one launch (the pool finds this callable inside dynesty's
                # wrappers); the proposals -- `queue_size` walks of evolve_point per pool.map, each calling the likelihood
                # one point at a time -- run in lock step on threads and meet at a rendezvous, one launch per round.
                post = samplers.VectorizedPosterior(self, *args)
                sampler_kwargs['pool'].posterior = post
                sampler_kwargs.setdefault('queue_size', sampler_kwargs['pool'].size)
                sampler = dynesty.NestedSampler(post, ptform, self.ndim, nlive=nlive, sample=sample, **sampler_kwargs)
                sampler.run_nested(**run_sampler_kwargs)
                return sampler
            sampler = dynesty.NestedSampler(self.jointprobability, ptform, self.ndim, logl_args=args, nlive=nlive,
                                            sample=sample, **sampler_kwargs)
            sampler.run_nested(**run_sampler_kwargs)
            return sampler

        if method in ('emcee', 'zeus', 'minimize'):
            init_pos = self._get_initial_pos(fix_kernel_params, fix_mean_params, fix_lensing_params)
            if type(init_scale) == list and len(init_pos) != len(init_scale):
                raise ValueError("The length of the initial parameter positions does not match the length of the list with the "
                                 "initial parameter scatter. The init_scale parameter should either be a single number (e.g., "
                                 "init_scale = 1.) or a list with the scale values for each parameter being fit.")
            if method == 'minimize':
                if minimize is None:
                    raise ImportError("scipy is not installed")
                if minimize_kwargs.pop('batched_gradient', False):
                    # value and forward-difference gradient from one batch of P + 1 points per iteration
                    post = samplers.VectorizedPosterior(self, logprior, fix_kernel_params, fix_mean_params, fix_lensing_params, -1)
                    bounds = minimize_kwargs.get('bounds')
                    return minimize(lambda th: samplers.fd_value_and_grad(post, th, bounds=bounds), init_pos, jac=True, **minimize_kwargs)
                return minimize(self.jointprobability, init_pos, args=args + (-1,), **minimize_kwargs)

            if np.any(np.isinf(logprior(init_pos))):
                raise Exception("The initial parameters have an infinite log prior; initialise the plugins inside the prior bounds.")
            # the reference reads nwalkers before assigning it (gausSN.py:342 vs :347) and passes invert=False
            # (:352, :355), which zeroes the posterior; both are repaired here rather than mirrored
            nwalkers = sampler_kwargs.pop('nwalkers', 24)
            nsteps = run_sampler_kwargs.pop('nsteps', 1000)
            p0 = np.random.normal(init_pos, init_scale, size=(nwalkers, self.ndim))
            for r, row in enumerate(p0):
                while np.isinf(logprior(row)):
                    p0[r] = row = np.random.normal(init_pos, 0.001)
            # both ensemble samplers accept vectorize=True: the whole half-ensemble is one GPU batch per step
            post = samplers.VectorizedPosterior(self, *args)
            sampler_kwargs.setdefault('vectorize', True)
            logprob = post if sampler_kwargs['vectorize'] else self.jointprobability
            extra = {} if sampler_kwargs['vectorize'] else {'args': args + (1,)}
            if method == 'emcee':
                if emcee is None:
                    raise ImportError("emcee is not installed")
                sampler = emcee.EnsembleSampler(nwalkers, self.ndim, logprob, **extra, **sampler_kwargs)
            else:
                if zeus is None:
                    raise ImportError("zeus is not installed")
                sampler = zeus.EnsembleSampler(nwalkers, self.ndim, logprob, **extra, **sampler_kwargs)
            sampler.run_mcmc(p0, nsteps=nsteps, **run_sampler_kwargs)
            return sampler
        raise ValueError(f"unknown method {method!r}")

    # ------------------------------------------------------------------ prediction
    def predict(self, x_prime, x, y, yerr, band=None):
        """Conditional mean and covariance at ``x_prime`` given observations ``(x, y, yerr)`` of a single band, for the
        plugins' current parameters (gaussn/gausSN.py:362-403).  One launch of the prediction kernel: the
        observations are factorised once and both moments come out of the same pass.  ``C = K + diag(yerr^2)`` must be
        positive definite (the reference's LU solve does not check); otherwise the moments are NaN."""
        x_prime = np.ascontiguousarray(np.asarray(x_prime, dtype=np.float64)).ravel()
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).ravel()
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64)).ravel()
        yerr = np.ascontiguousarray(np.asarray(yerr, dtype=np.float64)).ravel()
        if self.kernel.kind == _engine.K_JJ:
            raise NotImplementedError("JJKernel.covariance(x, x_prime) needs equal lengths (gaussn/kernels.py:555)")
        # One handle per (length, plugin kinds), kept across calls: the reference's plotting loop calls predict with
        # fresh (shifted, de-magnified) data for every posterior sample and band (gaussn/utils.py:193-204); replacing
        # the observations of a handle in place costs microseconds, creating one costs milliseconds.
        n_mean = len(self.meanfunc.params)
        key = (len(x), self.kernel.kind, self.meanfunc.kind, n_mean, self.device)
        data_key = (x.tobytes(), y.tobytes(), yerr.tobytes())
        cache = _PREDICT_HANDLES      # module level: that loop builds a new GP object for every sample (utils.py:193)
        entry = cache.pop(key, None)
        if entry is None:
            entry = [_engine.Problem(x, y, yerr, 1, 1, [0, len(x)], self.kernel.kind, self.meanfunc.kind,
                                     _engine.L_NONE, n_mean, self.device), data_key]
            while len(cache) >= 8:                      # least recently used first
                cache.pop(next(iter(cache)))[0].close()
        elif entry[1] != data_key:
            entry[0].set_data(x, y, yerr)
            entry[1] = data_key
        cache[key] = entry                               # most recently used last
        p = entry[0]
        kp = [float(v) for v in getattr(self.kernel, "params", [])][:p.n_kernel]
        theta = np.array(kp + [float(v) for v in self.meanfunc.params], dtype=np.float64).reshape(1, -1)
        mean_v = mean_u = None
        if self.meanfunc.kind == _engine.M_HOST:
            mean_v = np.asarray(self.meanfunc.mean(x, bands=band), dtype=np.float64).reshape(1, -1)
            mean_u = np.asarray(self.meanfunc.mean(x_prime, bands=band), dtype=np.float64).reshape(1, -1)
        e, v = _predict_chunked(p, theta, 0, len(x), x_prime, mean_v, mean_u, True)
        return e[0], v[0]

    def predict_batch(self, theta, x_prime, band=None, fix_kernel_params=False, fix_mean_params=False,
                      fix_lensing_params=False, return_cov=True, band_index=None):
        """``predict`` for every row of ``theta`` (e.g. equal-weight posterior samples) on the data set given to
        ``optimize_parameters``, the way the reference's plotting loop uses it (gaussn/utils.py:156-209): per sample,
        the observations of ``band`` are shifted by the sample's time delays and divided by its magnifications, and the
        GP is conditioned on them at ``x_prime`` (epochs in the frame of the first image).  ``band`` is a band LABEL (an entry
        of the ``band`` array given to ``optimize_parameters`` -- also when the labels are integers), ``band_index`` its
        position in ``np.unique`` order; neither: all rows.  Returns ``expectation [B, M]`` and
        ``variance [B, M, M]`` (``None`` with ``return_cov=False``) -- numpy for a numpy ``theta``, CUDA tensors for a
        CUDA ``theta``.  The whole batch is one kernel launch."""
        self._ensure_problem(self.x, self.y, self.yerr)
        if self.lensingmodel.kind == _engine.L_NONE:
            fix_lensing_params = True
        ncols = self._set_map(fix_kernel_params, fix_mean_params, fix_lensing_params)
        p = self._problem
        if band is None and band_index is None:
            row0, row1, label = 0, p.N, None
        else:
            labels = list(np.unique(self.bands)) if self.bands is not None else [None]
            if band is not None and band_index is not None:
                raise ValueError("give band (a label) or band_index (a position), not both")
            if band is not None:
                if band not in labels:
                    raise ValueError(f"band {band!r} is not one of the labels {labels}; use band_index= to select by position")
                b = labels.index(band)
            else:
                b = int(band_index)
                if not 0 <= b < len(labels):
                    raise IndexError(f"band_index {b} out of range for {len(labels)} bands")
            label = labels[b]
            row0, row1 = int(self.indices[b * self.n_images]), int(self.indices[(b + 1) * self.n_images])
        x_prime = np.ascontiguousarray(np.asarray(x_prime, dtype=np.float64)).ravel()
        is_dev = _is_torch_cuda(theta)
        if is_dev:
            import torch
            th = theta.to(torch.float64).contiguous().reshape(-1, ncols)
        else:
            th = np.ascontiguousarray(np.asarray(theta, dtype=np.float64)).reshape(-1, ncols)
        mean_v = mean_u = None
        if self.meanfunc.kind == _engine.M_HOST:
            mean_v, mean_u = self._host_means_for_predict(th, row0, row1, x_prime, label)
        e, v = _predict_chunked(p, th, row0, row1 - row0, x_prime, mean_v, mean_u, return_cov)
        return e, v

    def _host_means_for_predict(self, th, row0, row1, x_prime, label):
        """Host mean (sncosmoMean) per theta at the shifted epochs of the band's rows and at x_prime."""
        import torch
        p = self._problem
        dev = torch.device("cuda", p.device)
        th_d = th if _is_torch_cuda(th) else torch.as_tensor(th, device=dev)
        B = th_d.shape[0]
        shifted = torch.empty((B, p.N), dtype=torch.float64, device=dev)
        bvec = torch.empty((B, p.N), dtype=torch.float64, device=dev)
        p.lens_device(th_d.data_ptr(), B, th_d.shape[1], shifted.data_ptr(), bvec.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        shifted_h, th_h = shifted.cpu().numpy()[:, row0:row1], th_d.cpu().numpy()
        mean_cols = self._map_cols[p.n_kernel:p.n_kernel + p.n_mean]
        bands_v = None if self.bands is None else np.asarray(self.bands)[row0:row1]
        bands_u = None if label is None else np.repeat(label, len(x_prime))
        mv, mu = np.empty((B, row1 - row0)), np.empty((B, len(x_prime)))
        for i in range(B):
            mp = None if (len(mean_cols) and mean_cols[0] < 0) else [th_h[i, c] for c in mean_cols]
            mv[i] = self.meanfunc.mean(shifted_h[i], params=mp, bands=bands_v, zp=self.zp, zpsys=self.zpsys)
            mu[i] = self.meanfunc.mean(x_prime, params=None, bands=bands_u, zp=self.zp, zpsys=self.zpsys)
        return mv, mu


def _predict_chunked(p, th, row0, n_rows, x_prime, mean_v, mean_u, want_cov):
    """One gsn_predict_batch call when the conditioning rows and x_prime fit one factorisation (padded n_rows + M <=
    8128); otherwise x_prime is cut into pieces and every pair of pieces is predicted jointly (the cross blocks of the
    covariance need both)."""
    M = len(x_prime)
    room = _engine.MAX_ORDER - (n_rows + 31) // 32 * 32
    if room < 2:
        raise ValueError(f"{n_rows} conditioning epochs leave no room for prediction epochs (order limit {_engine.MAX_ORDER})")
    if M <= room:
        return _predict_call(p, th, row0, n_rows, x_prime, mean_v, mean_u, want_cov)
    step = room if not want_cov else room // 2
    cuts = list(range(0, M, step)) + [M]
    B = th.shape[0]
    is_dev = _is_torch_cuda(th)
    if is_dev:
        import torch
        e = torch.empty((B, M), dtype=torch.float64, device=th.device)
        v = torch.empty((B, M, M), dtype=torch.float64, device=th.device) if want_cov else None
    else:
        e = np.empty((B, M))
        v = np.empty((B, M, M)) if want_cov else None
    pieces = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
    for a, (a0, a1) in enumerate(pieces):
        if not want_cov:
            e[:, a0:a1] = _predict_call(p, th, row0, n_rows, x_prime[a0:a1], mean_v, None if mean_u is None else mean_u[:, a0:a1], False)[0]
            continue
        for (b0, b1) in pieces[a:] if len(pieces) > 1 else [pieces[a]]:
            same = (a0, a1) == (b0, b1)
            sel = np.arange(a0, a1) if same else np.concatenate([np.arange(a0, a1), np.arange(b0, b1)])
            ee, vv = _predict_call(p, th, row0, n_rows, x_prime[sel], mean_v, None if mean_u is None else mean_u[:, sel], True)
            na = a1 - a0
            e[:, a0:a1] = ee[:, :na]
            v[:, a0:a1, a0:a1] = vv[:, :na, :na]
            if not same:
                e[:, b0:b1] = ee[:, na:]
                v[:, b0:b1, b0:b1] = vv[:, na:, na:]
                v[:, a0:a1, b0:b1] = vv[:, :na, na:]
                v[:, b0:b1, a0:a1] = vv[:, na:, :na]
    return e, v


def _predict_call(p, th, row0, n_rows, x_prime, mean_v, mean_u, want_cov):
    if not _is_torch_cuda(th):
        e, v, _ = p.predict_host(th, row0, n_rows, x_prime, mean_v, mean_u, want_cov)
        return e, v
    import torch
    dev = th.device
    if dev.index != p.device:
        raise ValueError(f"theta lives on cuda:{dev.index} but the problem is on cuda:{p.device}")
    B, M = th.shape[0], len(x_prime)
    xp = torch.as_tensor(np.ascontiguousarray(x_prime), device=dev)
    mv = None if mean_v is None else torch.as_tensor(np.ascontiguousarray(mean_v), device=dev)
    mu = None if mean_u is None else torch.as_tensor(np.ascontiguousarray(mean_u), device=dev)
    e = torch.empty((B, M), dtype=torch.float64, device=dev)
    v = torch.empty((B, M, M), dtype=torch.float64, device=dev) if want_cov else None
    p.predict_device(th.data_ptr(), B, th.shape[1], row0, n_rows, xp.data_ptr(), M, e.data_ptr(), v.data_ptr() if want_cov else 0, 0,
                     0 if mv is None else mv.data_ptr(), 0 if mu is None else mu.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.current_stream(dev).synchronize()   # xp / mv / mu are temporaries of this call
    return e, v
