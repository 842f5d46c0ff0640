"""Host logic of the sampler adapters against protocol fakes of dynesty / emcee / zeus (tests/fake_samplers.py): the
batches the likelihood sees, and that lock-step evaluation changes no result.  No GPU: the "GP" is a stub."""
import numpy as np
import pytest

import fake_samplers
from gaussn_b200 import samplers


class StubGP:
    """Stands in for GP.jointprobability_batch: a Gaussian bump, vectorised, recording the batch sizes."""

    def __init__(self):
        self.batches = []

    def jointprobability_batch(self, thetas, logprior=None, fk=False, fm=False, fl=False, invert=1):
        thetas = np.atleast_2d(thetas)
        self.batches.append(len(thetas))
        return invert * (-0.5 * np.sum(((thetas - 0.4) / 0.15) ** 2, axis=1))


def _ptform(u):
    return 2.0 * u - 0.5


def _run(pool, queue_size, seed=3, rounds=5):
    gp = StubGP()
    post = samplers.VectorizedPosterior(gp)
    if pool is not None:
        pool.posterior = post
    s = fake_samplers.FakeNestedSampler(post, _ptform, 4, nlive=128, pool=pool, queue_size=queue_size, seed=seed).run_nested(rounds)
    return s, gp, post


def test_initial_live_points_are_one_launch_through_dynestys_wrappers():
    pool = samplers.BatchPool(64)
    s, gp, post = _run(pool, 64, rounds=0)
    assert gp.batches == [128] and pool.batch_sizes == [128]
    # what dynesty maps is not the posterior itself but LogLikelihood.loglikelihood, a _function_wrapper around it
    assert samplers._find_posterior(s.loglike.loglikelihood) is post and samplers._find_posterior(s.loglike) is post
    # a wrapper that adds arguments is left alone
    assert samplers._find_posterior(fake_samplers.FunctionWrapper(post, args=(1,))) is None


def test_rslice_walks_run_in_lock_step_and_meet_in_batches_of_queue_size():
    pool = samplers.BatchPool(64, lockstep=True)
    s, gp, post = _run(pool, 64)
    prop = gp.batches[1:]                    # after the live points
    # every round starts with all 64 walks at the rendezvous; walks end after different numbers of calls, so the batches
    # shrink towards the end of a round
    print(f"{s.ncall - 128} calls in {len(prop)} launches: mean {np.mean(prop):.1f}, max {max(prop)}, "
          f"{np.mean(np.array(prop) >= 32) * 100:.0f} % of the launches carry >= 32 thetas")
    assert max(prop) == 64 and prop[0] == 64
    assert sum(prop) == s.ncall - 128        # every scalar call of every walk went through a batch
    assert np.mean(prop) >= 16 and len(prop) <= (s.ncall - 128) / 16
    # the same sampler without a pool: identical chains (each walk has its own random stream), one call per launch
    s0, gp0, _ = _run(None, 64)
    np.testing.assert_array_equal(s.live_logl, s0.live_logl)
    np.testing.assert_array_equal(s.live_v, s0.live_v)
    assert s0.ncall == s.ncall and set(gp0.batches[1:]) == {1}
    pool.close()


def test_queue_shorter_than_pool_and_single_task():
    pool = samplers.BatchPool(256, lockstep=True)
    s, gp, _ = _run(pool, 16, rounds=3)
    assert max(gp.batches[1:]) == 16
    post = samplers.VectorizedPosterior(StubGP())
    pool.posterior = post
    assert pool.map(lambda a: post(np.full(4, a)), [0.3]) == [post(np.full(4, 0.3))]   # one task: no threads
    assert pool.map(lambda v: v + 1, []) == []
    pool.close()


def test_exception_in_one_walk_reaches_the_caller_and_releases_the_others():
    pool = samplers.BatchPool(8, lockstep=True)
    post = samplers.VectorizedPosterior(StubGP())
    pool.posterior = post

    def walk(k):
        v = post(np.full(4, 0.1 * k))
        if k == 3:
            raise ValueError("walk 3 fails")
        return v + post(np.full(4, 0.2))

    with pytest.raises(ValueError, match="walk 3"):
        pool.map(walk, list(range(8)))
    assert pool.map(walk, [0, 1, 2]) == [walk(0), walk(1), walk(2)]      # the pool is usable afterwards
    pool.close()


def test_failing_batch_evaluation_is_raised_in_every_walk():
    class Bad(StubGP):
        def jointprobability_batch(self, thetas, *a, **k):
            raise RuntimeError("device lost")

    pool = samplers.BatchPool(4, lockstep=True)
    post = samplers.VectorizedPosterior(Bad())
    pool.posterior = post
    with pytest.raises(RuntimeError, match="device lost"):
        pool.map(lambda k: post(np.full(4, 0.1 * k)), list(range(4)))
    pool.close()


def test_ensemble_samplers_vectorize_hands_over_half_ensembles():
    gp = StubGP()
    post = samplers.VectorizedPosterior(gp)
    s = fake_samplers.FakeEnsembleSampler(24, 4, post, vectorize=True)
    p0 = np.random.default_rng(1).uniform(0.2, 0.6, size=(24, 4))
    s.run_mcmc(p0, 10)
    assert gp.batches == [24] + [12] * 20
    gp2 = StubGP()
    s2 = fake_samplers.FakeEnsembleSampler(24, 4, samplers.VectorizedPosterior(gp2), vectorize=False)
    s2.run_mcmc(p0, 2)
    assert set(gp2.batches) == {1}


def test_auto_mode_times_the_first_round_and_decides():
    import time

    class Slow(StubGP):
        def jointprobability_batch(self, thetas, *a, **k):
            time.sleep(0.001)              # an evaluation that costs more than waking the walks
            return super().jointprobability_batch(thetas, *a, **k)

    for gp_cls, expect in ((StubGP, False), (Slow, True)):
        pool = samplers.BatchPool(16, lockstep="auto")
        gp = gp_cls()
        post = samplers.VectorizedPosterior(gp)
        pool.posterior = post
        fake_samplers.FakeNestedSampler(post, _ptform, 4, nlive=32, pool=pool, queue_size=16, seed=1).run_nested(3)
        assert pool.lockstep is expect
        assert (max(gp.batches[1:]) > 1) == expect
        pool.close()
