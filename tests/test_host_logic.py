"""CPU: host-side mirror of the reference interface -- plugin classes, segment bookkeeping,
theta packing, sharding (gloo, world size 2)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_prepare_indices_matches_reference(golden, libgsn):
    from gaussn_b200 import gausSN, kernels, meanfuncs
    for row in golden["indices"]:
        gp = gausSN.GP(kernels.ConstantKernel([1.0]), meanfuncs.UniformMean([0.0]))
        gp._prepare_indices(np.arange(row["n"], dtype=float), None if row["band"] is None else np.array(row["band"]),
                            None if row["image"] is None else np.array(row["image"]))
        assert list(gp.indices) == row["indices"]
        assert (gp.n_bands, gp.n_images) == (row["n_bands"], row["n_images"])
        assert gp.factor == pytest.approx(row["factor"], rel=1e-15)


def test_plugin_classes_mirror_the_reference_attributes(libgsn):
    from gaussn_b200 import kernels, meanfuncs, lensingmodels
    k = kernels.GibbsKernel([1.0, 2.0, 0.3, 4.0, 5.0])
    assert (k.A, k.lamdba, k.p, k.mu, k.sigma) == (1.0, 2.0, 0.3, 4.0, 5.0) and k.params == [1.0, 2.0, 0.3, 4.0, 5.0]
    k._reset([9.0, 8.0, 0.7, 6.0, 5.0])
    assert k.lamdba == 8.0 and k.params[0] == 9.0
    dp = kernels.DotProductKernel()
    assert not hasattr(dp, "params") and dp.c == 0
    dp._reset()
    m = meanfuncs.Karpenka2012([0.5, -1.0, 3.0, 2.0, 10.0, 2.0, 99.0])
    assert len(m.params) == 7 and m.A == pytest.approx(np.exp(0.5)) and m.B == pytest.approx(np.exp(-1.0))
    assert kernels.JJKernel([1.0, 2.0, 0.5, 3.0])._p(2.0) == 4.0
    lm = lensingmodels.SigmoidMagnification([10.0, 0.5, 0.3, 0.2, 40.0, 25.0, 1.7, -0.4, 0.05, 60.0])
    np.testing.assert_array_equal(lm.deltas, [0.0, 10.0, 25.0])
    np.testing.assert_array_equal(lm.beta0s, [1.0, 0.5, 1.7])
    np.testing.assert_array_equal(lm.t0s, [0.0, 40.0, 60.0])
    sm = lensingmodels.SinusoidalMagnification([10.0, 0.5, 0.3, 40.0, 9.0])
    np.testing.assert_array_equal(sm.Ts, [0.0, 9.0])
    cm = lensingmodels.ConstantMagnification([10.0, 0.5])
    cm.import_from_gp(2, 2, np.array([0, 2, 3, 5, 6]))
    np.testing.assert_array_equal(cm.repeats, [2, 1, 2, 1])
    want = np.zeros((6, 6)); want[:3, :3] = 1; want[3:, 3:] = 1
    np.testing.assert_array_equal(cm.mask, want)
    nl = lensingmodels.NoLensing()
    x = np.arange(3.0)
    assert nl.mask == 1 and nl.lens(x)[1] == 1 and nl.lens(x)[0] is x
    with pytest.raises(IndexError):
        kernels.ExpSquaredKernel([1.0])


def test_gp_rejects_plugins_without_a_device_implementation(libgsn):
    from gaussn_b200 import gausSN, kernels, meanfuncs

    class MyKernel:
        params = [1.0]

    with pytest.raises(TypeError, match="no CPU fallback"):
        gausSN.GP(MyKernel(), meanfuncs.UniformMean([0.0]))
    gp = gausSN.GP(kernels.ExpSquaredKernel([1.0, 2.0]), meanfuncs.UniformMean([0.0]))
    with pytest.raises(NotImplementedError):
        gp.optimize_parameters(np.arange(4.0), np.arange(4.0), np.ones(4), loglikelihood=lambda *a: 0.0)


def test_initial_position_follows_the_theta_packing_order(libgsn):
    from gaussn_b200 import gausSN, kernels, meanfuncs, lensingmodels
    gp = gausSN.GP(kernels.ExpSquaredKernel([1.0, 2.0]), meanfuncs.UniformMean([3.0]), lensingmodels.ConstantMagnification([4.0, 5.0]))
    assert gp._get_initial_pos(False, False, False) == [1.0, 2.0, 3.0, 4.0, 5.0]
    assert gp._get_initial_pos(True, False, False) == [3.0, 4.0, 5.0]
    assert gp._get_initial_pos(False, True, False) == [1.0, 2.0, 4.0, 5.0]
    assert gp._get_initial_pos(True, True, False) == [4.0, 5.0]
    with pytest.raises(Exception, match="No parameters"):
        gp._get_initial_pos(True, True, True)


def test_shard_bounds_cover_every_row_exactly_once():
    from gaussn_b200.sharding import shard_bounds
    for B in (0, 1, 7, 8, 9, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(B, dtype=int)
            for r in range(world):
                a, b, per = shard_bounds(B, world, r)
                assert 0 <= a <= b <= B and b - a <= per
                seen[a:b] += 1
            assert np.all(seen == 1)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _gloo_worker(rank, world, port, B, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from gaussn_b200.sharding import evaluate_sharded, shard_bounds
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    theta = rng.normal(size=(B, 5))
    calls = []

    def evaluate(th):       # stands in for GP.loglikelihood_batch (no GPU here): a row-wise function
        calls.append(len(th))
        return np.sin(th).sum(axis=1)

    full = evaluate_sharded(evaluate, theta)
    a, b, _ = shard_bounds(B, world, rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    assert calls == ([b - a] if b > a else [])
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [10, 11, 1])
def test_sharded_evaluation_world_size_2_gloo(tmp_path, B):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() + B) % 2000
    mp.spawn(_gloo_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.default_rng(1)
    want = np.sin(rng.normal(size=(B, 5))).sum(axis=1)
    for r in range(2):
        got = np.load(tmp_path / f"r{r}.npy")
        np.testing.assert_array_equal(got, want)       # bit-identical to the unsharded evaluation


def test_bench_flop_model():
    import benchdata
    # 44 739 243 + 262 144 + 919 296 flop for N=512, one block, ExpSquared
    assert benchdata.algorithmic_flops(512, 1, True) == pytest.approx(44739242.67 + 262144 + 919296, rel=1e-9)
    # block-diagonal by band: never count the masked-out zero blocks
    assert benchdata.algorithmic_flops(512, 4, True) == pytest.approx(4 * (128**3 / 3 + 128**2 + 7 * 128 * 129 / 2))
    assert benchdata.algorithmic_flops(512, 4, False) == benchdata.algorithmic_flops(512, 1, True)
    ds = benchdata.make_dataset(96, 2, 3)
    assert len(ds["x"]) == 96 and list(ds["indices"]) == [0, 16, 32, 48, 64, 80, 96]
    for s in range(6):
        seg = ds["x"][ds["indices"][s]:ds["indices"][s + 1]]
        assert np.all(np.diff(seg) >= 0)
    th = benchdata.make_theta(100, 3)
    assert th.shape == (100, 7) and th[:, 1].min() >= 10 and th[:, 1].max() <= 40


def _check_item_order(N, band_of_row, n_bands, masked):
    """Every non-zero block exactly once, and every item after everything it waits for (the kernel's
    warps pull items in this order and spin on unfinished predecessors: a non-topological list could deadlock).
    With a band mask the rows are laid out with every band starting at a multiple of 32, so that no chunk couples
    two bands: the factor is then exactly block diagonal at chunk granularity."""
    from gaussn_b200 import _engine
    items = _engine.item_order(N, band_of_row, n_bands, masked)
    band = np.zeros(N, dtype=int) if band_of_row is None else np.asarray(band_of_row)
    use_mask = masked and n_bands > 1
    if use_mask:
        chunk_band = []
        for b in range(n_bands):
            chunk_band += [b] * ((int(np.sum(band == b)) + 31) // 32)
    else:
        chunk_band = [0] * ((N + 31) // 32)
    nblk = len(chunk_band)
    first = [chunk_band.index(chunk_band[c]) if use_mask else 0 for c in range(nblk)]
    want = {(c, j) for c in range(nblk) for j in range(first[c], c + 1)}
    got = [(c, j) for c, j, _ in items]
    assert len(got) == len(set(got)) and set(got) == want
    assert all(k0 == first[c] for c, j, k0 in items)
    pos = {cj: i for i, cj in enumerate(got)}
    for (c, j), i in pos.items():
        deps = []
        if j > first[c]:
            deps += [(c, j - 1), (j, j - 1)] if c != j else [(j, j - 1)]
        if c != j:
            deps.append((j, j))
        assert all(pos[d] < i for d in deps), ((c, j), deps)
    return items


@pytest.mark.parametrize("N", [1, 31, 32, 33, 64, 200, 336, 337, 352, 399, 400, 432, 512, 1000, 2048, 4192, 8192])
def test_item_order_is_complete_and_topological(libgsn, N):
    _check_item_order(N, None, 1, False)


def test_item_order_skips_zero_blocks_of_the_band_mask(libgsn):
    for N, sizes in ((512, [128] * 4), (512, [256, 256]), (396, [132] * 3), (194, [51, 48, 45, 50]), (700, [100, 300, 300])):
        band = np.repeat(np.arange(len(sizes)), sizes)
        dense = _check_item_order(N, band, len(sizes), False)
        masked = _check_item_order(N, band, len(sizes), True)
        assert len(masked) < len(dense)
    # without a mask (NoLensing) the bands stay coupled: nothing is skipped
    nblk = 16
    assert len(_check_item_order(512, np.repeat([0, 1], 256), 2, False)) == nblk * (nblk + 1) // 2


@pytest.mark.parametrize("n_rows,M,want_cov", [(1, 1, True), (48, 100, True), (64, 64, False), (70, 9, True), (512, 256, True),
                                               (2048, 512, True), (2048, 512, False), (33, 4000, True)])
@pytest.mark.parametrize("cluster", [False, True])
def test_predict_item_order_is_complete_and_topological(libgsn, n_rows, M, want_cov, cluster):
    """Prediction = partial factorisation of the augmented matrix: the conditioning block columns are factorised (all
    rows), the trailing blocks among the prediction rows are only accumulated -- and only listed when wanted."""
    from gaussn_b200 import _engine
    items = _engine.predict_item_order(n_rows, M, want_cov, cluster)
    nVb = (n_rows + 31) // 32
    nblk = nVb + (M + 31) // 32
    want = {(c, j, False) for c in range(nblk) for j in range(min(c, nVb - 1) + 1)}
    want |= {(c, j, True) for c in range(nVb, nblk) for j in range(nVb, c + 1) if want_cov or c == j}
    assert len(items) == len(set(items)) and set(items) == want
    pos = {(c, j): i for i, (c, j, _) in enumerate(items)}
    for (c, j, trail), i in zip(items, range(len(items))):
        if trail:    # needs the finished rows c and j of the factorised columns
            deps = [(c, nVb - 1), (j, nVb - 1)]
        else:
            deps = ([(c, j - 1), (j, j - 1)] if c != j else [(j, j - 1)]) if j > 0 else []
            if c != j:
                deps.append((j, j))
        assert all(pos[d] < i for d in deps), ((c, j, trail), deps)


def test_long_prediction_grids_are_assembled_from_pairs_of_pieces(monkeypatch):
    """GP.predict_batch cuts a grid that does not fit one factorisation into pieces and predicts every pair of pieces
    jointly (the cross blocks of the covariance need both).  Host logic only: the engine call is replaced by an exact
    numpy conditional, and the assembled moments must equal the direct ones."""
    from gaussn_b200 import gausSN, _engine
    rng = np.random.default_rng(5)
    xv = np.sort(rng.uniform(0, 50, 40))
    yv = np.sin(xv / 7.0)
    noise = 0.05

    def k(a, b):
        return 0.8 * np.exp(-0.5 * (a[:, None] - b[None, :]) ** 2 / 36.0)

    def fake_call(p, th, row0, n_rows, x_prime, mean_v, mean_u, want_cov):
        C = k(xv, xv) + noise ** 2 * np.eye(len(xv))
        Kuv = k(np.asarray(x_prime), xv)
        e = np.tile(Kuv @ np.linalg.solve(C, yv), (th.shape[0], 1))
        v = np.tile(k(np.asarray(x_prime), np.asarray(x_prime)) - Kuv @ np.linalg.solve(C, Kuv.T), (th.shape[0], 1, 1)) if want_cov else None
        return e, v

    monkeypatch.setattr(gausSN, "_predict_call", fake_call)
    th = np.zeros((3, 2))
    x_prime = np.linspace(-5, 60, 101)
    e_ref, v_ref = fake_call(None, th, 0, 40, x_prime, None, None, True)
    monkeypatch.setattr(_engine, "MAX_ORDER", 64 + 30)      # 40 rows pad to 64: room for 30 epochs, pieces of 15
    e, v = gausSN._predict_chunked(None, th, 0, 40, x_prime, None, None, True)
    np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-13)
    np.testing.assert_allclose(v, v_ref, rtol=0, atol=1e-13)
    e2, v2 = gausSN._predict_chunked(None, th, 0, 40, x_prime, None, None, False)
    assert v2 is None
    np.testing.assert_allclose(e2, e_ref, rtol=0, atol=1e-13)
    monkeypatch.setattr(_engine, "MAX_ORDER", 64 + 1)
    with pytest.raises(ValueError):
        gausSN._predict_chunked(None, th, 0, 40, x_prime, None, None, True)


class _FakeGP:
    """Stands in for GP in the sampler adapters: a quadratic 'posterior' evaluated in batches, counting the launches."""

    def __init__(self):
        self.calls = []

    def jointprobability_batch(self, thetas, logprior=None, fix_kernel_params=False, fix_mean_params=False,
                               fix_lensing_params=False, invert=1):
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        self.calls.append((thetas.shape[0], (fix_kernel_params, fix_mean_params, fix_lensing_params), invert))
        val = -0.5 * np.sum((thetas - np.arange(1, thetas.shape[1] + 1)) ** 2, axis=1)
        if logprior is not None:
            val = val + np.array([logprior(t) for t in thetas])
        return invert * val


def test_vectorized_posterior_and_batch_pool_evaluate_queues_in_one_call():
    """the adapters that turn the samplers' one-theta-at-a-time loops (gausSN.py:321-323, :352-359)
    into batches.  No GPU needed: the likelihood is a stand-in."""
    from gaussn_b200 import samplers
    gp = _FakeGP()
    post = samplers.VectorizedPosterior(gp, lambda t: -0.1 * t[0], False, True, False, invert=1)
    pts = np.random.default_rng(1).normal(size=(7, 3))
    out = post(pts)
    assert out.shape == (7,) and gp.calls[-1] == (7, (False, True, False), 1) and post.n_calls == 1 and post.n_evals == 7
    one = post(pts[2])
    assert isinstance(one, float) and one == out[2] and post.n_calls == 2 and post.n_evals == 8
    # dynesty-style pool: the whole queue in one launch; any other callable falls back to a plain loop
    pool = samplers.BatchPool(size=64)
    assert pool.size == 64
    got = pool.map(post, [p for p in pts])
    assert np.allclose(got, out) and gp.calls[-1][0] == 7 and post.n_calls == 3
    assert pool.map(lambda v: v * 2, [1, 2, 3]) == [2, 4, 6]

    def wrapped(theta):          # a function that carries the posterior, as dynesty's loglikelihood wrapper may
        return post(theta)
    wrapped.vectorized = post
    got2 = pool.map(wrapped, [p for p in pts[:4]])
    assert np.allclose(got2, out[:4]) and gp.calls[-1][0] == 4
    pool.close(); pool.join()


def test_batched_finite_difference_gradient_matches_the_analytic_one():
    """scipy's minimize differences the scalar objective with P + 1 separate calls (gausSN.py:335); fd_value_and_grad
    gets value and gradient from one batch of P + 1 points and keeps the steps inside the bounds."""
    from gaussn_b200 import samplers
    gp = _FakeGP()
    post = samplers.VectorizedPosterior(gp, None, False, False, False, invert=-1)     # minimise -logP
    theta = np.array([0.3, -2.0, 5.0, 40.0])
    f, g = samplers.fd_value_and_grad(post, theta)
    assert gp.calls[-1][0] == len(theta) + 1 and post.n_calls == 1
    want_f = 0.5 * np.sum((theta - np.arange(1, 5)) ** 2)
    np.testing.assert_allclose(f, want_f, rtol=1e-14)
    np.testing.assert_allclose(g, theta - np.arange(1, 5), rtol=1e-5, atol=1e-5)   # forward differences: ~eps f / h
    # a point on its upper bound: the step goes inwards, the gradient is still right
    bounds = [(0.0, 0.3), (-5.0, 5.0), (None, 5.0), (0.0, None)]
    f2, g2 = samplers.fd_value_and_grad(post, theta, bounds=bounds)
    np.testing.assert_allclose(g2, theta - np.arange(1, 5), rtol=1e-5, atol=1e-5)
    used = None

    class Spy(samplers.VectorizedPosterior):
        def __call__(self, th):
            nonlocal used
            used = np.array(th)
            return super().__call__(th)
    spy = Spy(gp, None, False, False, False, invert=-1)
    samplers.fd_value_and_grad(spy, theta, bounds=bounds)
    assert np.all(used[1:, 0] <= 0.3) and np.all(used[:, 2] <= 5.0) and np.all(used[:, 3] >= 0.0)


@pytest.mark.parametrize("lag", [1, 2])
def test_onchip_tile_slots_are_never_handed_over_while_their_tenant_is_live(libgsn, lag):
    """The on-chip shape keeps only the live part of the factor: tile (i, k) sits in a recycled shared-memory slot
    (gsn_onchip.cuh).  A left-looking sweep reads the tiles of row r for the last time in tile column r, and the kernel lets
    a warp store into column p only when every warp has left column p - lag (one-warp teams: lag 1, the warp's own program
    order).  So the table is safe iff, for every slot, a tenant born in phase p replaces a tile of a row <= p - lag -- replayed
    here for every block size the kernel accepts (compute-sanitizer cannot be run on this pool; this is the host-side half
    of its job, the device-side half is the NaN-poisoned build of tests/test_gpu_onchip.py)."""
    from gaussn_b200 import _engine

    def birth(i, k):            # phase (tile column in work) in which tile (i, k) is first written; -1: the first pivot tile
        return k if i > k else i - 1

    for nt in range(1, 29):
        n_slots, tab = _engine.onchip_slot_table(nt, lag)
        tenants = {}
        for i in range(nt):
            for k in range(i + 1):
                slot = int(tab[i * (i + 1) // 2 + k])
                assert 0 <= slot < n_slots
                tenants.setdefault(slot, []).append((birth(i, k), i, k))
        assert len(tenants) == n_slots                                  # no slot is counted but unused
        for slot, ts in tenants.items():
            ts.sort()
            for (p0, i0, k0), (p1, i1, k1) in zip(ts, ts[1:]):
                assert p1 > p0, f"nt={nt}: tiles ({i0},{k0}) and ({i1},{k1}) are born in the same phase into slot {slot}"
                assert i0 <= p1 - lag, f"nt={nt}: tile ({i1},{k1}) takes slot {slot} in phase {p1} while row {i0} is still read"
        live_max = max(sum(1 for i in range(nt) for k in range(i + 1) if birth(i, k) <= p and i > p - lag) for p in range(-1, nt - 1))
        assert n_slots >= live_max                                       # cannot be smaller than the largest live set
    assert _engine.onchip_slot_table(18, 2)[0] == 100 and _engine.onchip_slot_table(18, 1)[0] == 91    # N = 141
    assert _engine.onchip_slot_table(8, 2)[0] == 25 and _engine.onchip_slot_table(8, 1)[0] == 21       # N = 64
