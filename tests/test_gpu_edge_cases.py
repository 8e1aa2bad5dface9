"""Edge cases of the device path against the oracle (f64, injected randoms) and of the error contract:
tiny and tile-boundary particle counts, one-step windows, empty subsequences, degenerate weights
(NaN observation -> ValueError like np.random.choice with NaN probabilities, pf.py:28-29), invalid arguments."""
import numpy as np
import pytest

from oracle import pf_oracle as po

pytestmark = pytest.mark.gpu
THETA = dict(A=0.95, LQinv=np.sqrt(2.0), Qinv=2.0 + 1e-16, LRinv=np.sqrt(2.0), Rinv=2.0 + 1e-16)
TH = [THETA[k] for k in ("A", "LQinv", "Qinv", "LRinv", "Rinv")]


def _run(N, obs, t1, tL, pf="poyiadjis_N", resample="multinomial", seed=3, weights=None, path="auto", **kw):
    import sgmcmc_ssm_b200 as sg
    K = po.make_kernel("svm", "prior", THETA)
    if resample == "multinomial":
        rec = po.LegacyStream(seed, record=True)
        ref = po.buffered_pf(pf, obs, K, N, K.score, K.p, rec, t1=t1, tL=tL, weights=weights, prior_mean=0.0, prior_var=10.0, **kw)
        parts = po.split_events(rec.events, N)
        inj = dict(z0=parts["z0"], u=parts["u"], z=parts["z"], extra=parts["extra"])
    else:
        rs = np.random.RandomState(seed)
        T = len(obs)
        z0, z = rs.normal(size=N), rs.normal(size=(T, N))
        u = np.sort(rs.random_sample((T, N)), axis=1)
        ref = po.buffered_pf(pf, obs, K, N, K.score, K.p, po.InjectedStream(u.ravel(), np.concatenate([z0, z.ravel()])),
                             t1=t1, tL=tL, weights=weights, prior_mean=0.0, prior_var=10.0, **kw)
        inj = dict(z0=z0, u=u, z=z)
    items = sg.PFItems().add(obs, TH, t1=t1, tL=tL, weights=weights, prior_mean=0.0, prior_var=10.0)
    res = sg.run_pf("svm", "prior", pf, items, N, dtype="f64", rng="injected", resample=resample, injected=inj, path=path, **kw)
    expect = ref["statistics"] if pf == "filter" else po.average_statistic(ref)
    return res, expect, ref


@pytest.mark.parametrize("path", ["auto", "small", "tiles", "cluster"])      # shared-memory kernel (N <= 2048), tile kernels, cluster kernel (256 < N <= 16384)
@pytest.mark.parametrize("resample", ["multinomial", "multinomial_sorted"])
@pytest.mark.parametrize("N", [1, 2, 31, 255, 256, 257, 511, 513, 1023, 1024, 1025, 2047, 2048, 2049, 8191, 65537, 70000])   # 65537+: 1024-thread header
def test_tiny_and_tile_boundary_particle_counts(N, resample, path):
    if (path == "small" and N > 2048) or (path == "cluster" and not 256 < N <= 16384) or (path == "tiles" and N > 16384):
        pytest.skip("path does not apply / same kernels as path='auto'")
    obs = np.array([0.3, -1.2, 0.8, 2.0, -0.1, 0.4])
    res, expect, ref = _run(N, obs, 1, 5, resample=resample, weights=np.array([1.0, 2.0, 0.5, 3.0]), path=path)
    np.testing.assert_allclose(res.grad[0], expect, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(res.loglik[0], ref["loglikelihood_estimate"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("pf", ["poyiadjis_N", "poyiadjis_N2", "nemeth", "filter", "paris"])
def test_one_step_window_and_empty_subsequence(pf):
    kw = dict(Ntilde=2) if pf == "paris" else {}
    res, expect, ref = _run(300, np.array([0.7]), 0, 1, pf=pf, **kw)
    np.testing.assert_allclose(res.grad[0], expect, rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(res.loglik[0], ref["loglikelihood_estimate"], rtol=1e-9)
    res, expect, ref = _run(300, np.array([0.7, -0.2, 0.1]), 2, 2, pf=pf, **kw)          # [t1, tL) empty: all buffer
    np.testing.assert_allclose(res.grad[0], 0.0, atol=1e-12)
    assert res.loglik[0] == 0.0 and ref["loglikelihood_estimate"] == 0.0


def test_degenerate_weights_raise_like_the_reference():
    import sgmcmc_ssm_b200 as sg
    obs = np.array([0.3, np.nan, 0.8, 0.1])
    items = sg.PFItems().add(obs, TH, prior_mean=0.0, prior_var=10.0).add(np.array([0.3, 0.2, 0.8, 0.1]), TH, prior_mean=0.0, prior_var=10.0)
    with pytest.raises(ValueError, match="NaN"):
        sg.run_pf("svm", "prior", "poyiadjis_N", items, 3000, dtype="f32")
    res = sg.run_pf("svm", "prior", "poyiadjis_N", items, 3000, dtype="f32", check=False)
    assert res.status[0] != 0 and res.status[1] == 0 and np.all(np.isfinite(res.grad[1]))


def test_injected_uniforms_default_to_the_generic_search_and_sorted_modes_need_ascending_uniforms():
    """Recorded uniforms come in the reference's iid order: with rng='injected' the resampling mode defaults to 'multinomial'
    (per-child search); the streaming modes stage a CDF window from a tile's first and last target only, so unsorted
    uniforms are refused instead of silently clamped (N > 256: more than one tile)."""
    import sgmcmc_ssm_b200 as sg
    N, obs = 700, np.array([0.3, -1.2, 0.8, 2.0])
    K = po.make_kernel("svm", "prior", THETA)
    rec = po.LegacyStream(5, record=True)
    ref = po.buffered_pf("poyiadjis_N", obs, K, N, K.score, K.p, rec, t1=0, tL=4, prior_mean=0.0, prior_var=10.0)
    parts = po.split_events(rec.events, N)
    inj = dict(z0=parts["z0"], u=parts["u"], z=parts["z"])
    items = sg.PFItems().add(obs, TH, prior_mean=0.0, prior_var=10.0)
    for path in ("tiles", "small", "cluster"):
        res = sg.run_pf("svm", "prior", "poyiadjis_N", items, N, dtype="f64", rng="injected", injected=inj, path=path)   # no resample= given
        np.testing.assert_allclose(res.grad[0], po.average_statistic(ref), rtol=1e-8, atol=1e-9)
    with pytest.raises(ValueError, match="ascending"):
        sg.run_pf("svm", "prior", "poyiadjis_N", items, N, dtype="f64", rng="injected", injected=inj, resample="multinomial_sorted")


def test_invalid_arguments():
    import sgmcmc_ssm_b200 as sg
    items = sg.PFItems().add(np.array([0.3, 0.2]), TH)
    with pytest.raises(ValueError):
        sg.run_pf("svm", "prior", "no_such_pf", items, 100)
    with pytest.raises(NotImplementedError):
        sg.run_pf("svm", "optimal", "poyiadjis_N", items, 100)           # svm/helper.py:62
    with pytest.raises(ValueError):
        sg.run_pf("svm", "prior", "paris", items, 100, Ntilde=0)
    with pytest.raises(ValueError):
        sg.run_pf("svm", "prior", "poyiadjis_N", sg.PFItems(), 100)
    with pytest.raises(ValueError):
        sg.PFItems().add(np.array([0.3, 0.2, 0.1]), TH, t1=0, tL=3, weights=np.ones(2))
