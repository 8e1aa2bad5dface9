"""Size-independent properties at BASELINE.json's full sizes (N = 2^16 particles, device randoms, f32), where the
CPU oracle is too slow to be the checker (one N = 2^16 gradient takes 2.5 s per subsequence on the CPU).

With the same Philox seed / call offset / global item index the genealogy of an item is a deterministic function
of (observations, theta, prior), so:
  * linearity      : the statistic is additive in the per-step weights -> scaling them by c scales the gradient by c
                     and the weighted log-likelihood by c, on an identical particle system;
  * stat invariance: the log-likelihood does not depend on which statistic is carried (score / suff / none);
  * shard invariance: a batch run in one call == the same items run as two calls with item_id_base (what the
                     multi-GPU sharding does), bit for bit;
  * stream invariance: the two-stream pipelining of half batches is bit-identical to the single-stream run;
  * consistency    : f32 and f64 gradients of the same 512 windows agree within the standard error of their paired
                     differences at N = 2^16;
  * N = 2^20       : the largest supported particle count runs and agrees with N = 2^16 on the log-likelihood."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N16 = 1 << 16
THETA = [0.95, np.sqrt(2.0), 2.0 + 1e-16, np.sqrt(2.0), 2.0 + 1e-16]


def _series(T=400, seed=3):
    rs = np.random.RandomState(seed)
    x, y = 0.0, np.zeros(T)
    for t in range(T):
        x = 0.95 * x + np.sqrt(0.5) * rs.normal()
        y[t] = np.sqrt(0.5) * np.exp(0.5 * x) * rs.normal()
    return y


def _items(B, scale=1.0, seed=0, Tb=60):
    import sgmcmc_ssm_b200 as sg
    y = _series()
    rs = np.random.RandomState(seed)
    it = sg.PFItems()
    for _ in range(B):
        s = int(rs.randint(0, 400 - Tb))
        it.add(y[s:s + Tb], THETA, t1=10, tL=Tb - 10, weights=scale * (1.0 + rs.rand(Tb - 20)), prior_mean=0.0, prior_var=10.0)
    return it


def test_linearity_in_the_step_weights_at_full_size():
    import sgmcmc_ssm_b200 as sg
    kw = dict(dtype="f32", rng="philox", seed=7, offset=3)
    a = sg.run_pf("svm", "prior", "poyiadjis_N", _items(24, 1.0), N16, **kw)
    b = sg.run_pf("svm", "prior", "poyiadjis_N", _items(24, 4.0), N16, **kw)      # power of two: exact in f32
    np.testing.assert_allclose(b.grad, 4.0 * a.grad, rtol=2e-5, atol=1e-3)
    np.testing.assert_allclose(b.loglik, 4.0 * a.loglik, rtol=1e-12)


def test_loglikelihood_independent_of_the_carried_statistic():
    import sgmcmc_ssm_b200 as sg
    kw = dict(dtype="f32", rng="philox", seed=8, offset=1)
    it = _items(16)
    ll = [sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, stat_kind=k, **kw).loglik for k in ("score", "suff", "none")]
    np.testing.assert_array_equal(ll[0], ll[1])
    np.testing.assert_array_equal(ll[0], ll[2])


def test_shard_and_stream_invariance_bit_for_bit():
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200 import engine
    it = _items(40, seed=5)
    kw = dict(dtype="f32", rng="philox", seed=9, offset=2)
    whole = sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, **kw)            # 40 items x 32 CTAs: two-stream pipelined
    pk = it.pack()
    lo = sg.run_pf("svm", "prior", "poyiadjis_N", pk.slice(0, 17), N16, item_id_base=0, **kw)
    hi = sg.run_pf("svm", "prior", "poyiadjis_N", pk.slice(17, 40), N16, item_id_base=17, **kw)
    np.testing.assert_array_equal(whole.grad, np.concatenate([lo.grad, hi.grad]))
    np.testing.assert_array_equal(whole.loglik, np.concatenate([lo.loglik, hi.loglik]))
    engine.config.two_streams = False
    try:
        single = sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, **kw)
    finally:
        engine.config.two_streams = True
    np.testing.assert_array_equal(whole.grad, single.grad)


@pytest.mark.parametrize("dtype,resample,N,per", [("f32", "multinomial_sorted", N16, 4), ("f32", "multinomial_sorted", 60000, 4),
                                                  ("f64", "multinomial", 8192, 30), ("f32", "multinomial_sorted", 10000, 25),
                                                  ("f32", "multinomial_sorted", 32768, 3), ("f64", "multinomial_sorted", 20000, 2)])
def test_cooperative_single_launch_equals_per_step_launches_bit_for_bit(dtype, resample, N, per):
    """A small batch of the tile kernels runs the whole time loop in ONE launch (csrc/coop_kernels.cuh, per-CTA header copies):
    N <= 32768 with the CTAs of an item as a thread-block cluster and the cluster barrier instead of the kernel boundary (the
    8192 / 10000 / 20000 / 32768 cases), else cooperatively with a grid barrier when all CTAs are resident (the 60000 / 65536 cases);
    path='steps' takes one header + one step launch per time step.  Same device functions, same counter-based randoms:
    bit-identical."""
    import sgmcmc_ssm_b200 as sg
    it = _items(3 * per, seed=5)
    kw = dict(dtype=dtype, rng="philox", resample=resample, seed=9, offset=2)
    whole = sg.run_pf("svm", "prior", "poyiadjis_N", it, N, path="steps", **kw)
    pk = it.pack()
    parts = [sg.run_pf("svm", "prior", "poyiadjis_N", pk.slice(i, i + per), N, item_id_base=i, path="tiles", **kw)
             for i in range(0, 3 * per, per)]
    assert whole.launches > 100 and all(p.launches == 1 for p in parts), (whole.launches, [p.launches for p in parts])
    np.testing.assert_array_equal(whole.grad, np.concatenate([p.grad for p in parts]))
    np.testing.assert_array_equal(whole.loglik, np.concatenate([p.loglik for p in parts]))


@pytest.mark.parametrize("dtype", ["f32", "f64"])
@pytest.mark.parametrize("N", [200, 512, 1000, 2048])
def test_shared_memory_kernel_and_tile_kernels_have_the_same_law(N, dtype):
    """N <= 2048 runs the shared-memory-resident one-CTA-per-item kernel (small_kernels.cuh: iid multinomial uniforms,
    block-level CDF), `path='tiles'` forces the per-step warp-tile kernels (order-statistics resampling, hierarchical
    CDF).  Different random streams, same estimator: over 1024 repetitions of one window the means agree within 4.5
    standard errors and the spreads within 15 %, for the gradient and the log-likelihood."""
    import sgmcmc_ssm_b200 as sg
    R = 1024
    it = _items(1, seed=21).pack()
    pk = sg.PackedItems(np.tile(it.obs_flat, R), np.tile(it.T_buf, R), np.tile(it.t1, R), np.tile(it.tL, R),
                        np.tile(it.wts_flat, R), np.arange(R, dtype=np.int64) * it.wts_flat.shape[0], it.theta[0], 0.0, 10.0)
    kw = dict(dtype=dtype, rng="philox", offset=1)
    a = sg.run_pf("svm", "prior", "poyiadjis_N", pk, N, seed=4, path="small", **kw)         # shared-memory kernel
    b = sg.run_pf("svm", "prior", "poyiadjis_N", pk, N, seed=5, path="tiles", **kw)         # per-step tile kernels
    assert a.launches == 1 and b.launches > 100
    ga, gb = np.column_stack([a.grad, a.loglik]), np.column_stack([b.grad, b.loglik])
    se = np.sqrt(ga.var(axis=0, ddof=1) / R + gb.var(axis=0, ddof=1) / R)
    assert np.all(np.abs(ga.mean(axis=0) - gb.mean(axis=0)) <= 4.5 * se), (ga.mean(axis=0), gb.mean(axis=0), se)
    assert np.all(np.abs(ga.std(axis=0) / gb.std(axis=0) - 1) < 0.15), (ga.std(axis=0), gb.std(axis=0))


@pytest.mark.parametrize("variates", ["native", "f32"])
def test_f32_and_f64_agree_within_standard_error_at_full_size(variates):
    """N = 2^16, the SAME 512 heterogeneous windows in f32 and in f64 (independent random streams).  Paired differences
    d_i = g32_i - g64_i have mean zero iff the two arithmetic types estimate the same quantity; with 512 items the mean of
    d is resolved to std(d) / sqrt(512), i.e. a bias of ~4 % of ONE item's Monte-Carlo standard deviation -- or of 0.2 % of
    a gradient component's typical size -- fails:   |mean d| <= 4.5 std(d) / sqrt(512)  per gradient component and for the
    log-likelihood, and the spreads of the two samples agree within 15 %."""
    import sgmcmc_ssm_b200 as sg
    B = 512
    it = _items(B, seed=11)
    a = sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, dtype="f32", seed=1, offset=1)
    b = sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, dtype="f64", seed=2, offset=1, variates=variates)
    c = sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, dtype="f32", seed=3, offset=1)      # a second f32 draw: the noise floor
    ga, gb, gc = (np.column_stack([r.grad, r.loglik]) for r in (a, b, c))
    d = ga - gb
    se = d.std(axis=0, ddof=1) / np.sqrt(B)
    assert np.all(np.abs(d.mean(axis=0)) <= 4.5 * se), (d.mean(axis=0), se)
    # the f32-f64 differences are no wider than the differences between two f32 runs (pure Monte-Carlo noise)
    ratio = d.std(axis=0, ddof=1) / (ga - gc).std(axis=0, ddof=1)
    assert np.all(np.abs(ratio - 1) < 0.15), ratio


def test_largest_particle_count_runs_and_agrees():
    import sgmcmc_ssm_b200 as sg
    it = _items(2, seed=13, Tb=24)
    big = sg.run_pf("svm", "prior", "poyiadjis_N", it, 1 << 20, dtype="f32", seed=3, offset=1)
    ref = sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, dtype="f64", seed=4, offset=1)
    assert np.all(big.status == 0) and np.all(np.isfinite(big.grad))
    np.testing.assert_allclose(big.loglik, ref.loglik, rtol=0, atol=0.3)
    with pytest.raises(ValueError):
        sg.run_pf("svm", "prior", "poyiadjis_N", it, (1 << 20) + 1)


def test_cuda_graph_replay_is_bit_identical_to_direct_launches():
    """Launch-bound batches replay a captured CUDA graph (Philox call offset read from device memory): the results of
    two calls with different offsets must equal the direct-launch results, and differ from each other."""
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200 import engine
    it = _items(3, seed=17)
    outs = {}
    for graphs in (True, False):
        engine.config.cuda_graphs = graphs
        try:
            outs[graphs] = [sg.run_pf("svm", "prior", "poyiadjis_N", it, N16, dtype="f32", seed=21, offset=o) for o in (5, 6, 5)]
        finally:
            engine.config.cuda_graphs = True
    for a, b in zip(outs[True], outs[False]):
        np.testing.assert_array_equal(a.grad, b.grad)
        np.testing.assert_array_equal(a.loglik, b.loglik)
    np.testing.assert_array_equal(outs[True][0].grad, outs[True][2].grad)
    assert np.any(outs[True][0].grad != outs[True][1].grad)
