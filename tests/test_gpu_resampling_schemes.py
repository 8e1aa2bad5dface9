"""Systematic and stratified resampling (north-star extension; the reference itself is multinomial only, pf.py:27-29).

1. f64 + INJECTED: the scheme's uniforms are fed as the resampling uniforms u_i = (i + v) / N (systematic, one v per step)
   or (i + v_i) / N (stratified) to the ORACLE -- still the reference algorithm: searchsorted(cdf, u_i, 'right') -- and
   to the kernel with resample='systematic' / 'stratified': ancestors bit-exact, gradient 1e-8.
2. Device randoms (Philox): structural properties of the drawn ancestors, from the traced log-weights:
   systematic  -> parent j gets floor(N w_j) or ceil(N w_j) children;
   stratified  -> child i's ancestor owns a CDF interval that meets [i / N, (i + 1) / N);
   both        -> ancestors ascending.
3. Same estimator mean as multinomial resampling (4.5 sigma over 1024 repetitions), spread not larger.
"""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests import _cases as C

pytestmark = pytest.mark.gpu


def _scheme_uniforms(rs, scheme, T, N):
    i = np.arange(N)[None, :]
    v = rs.random_sample((T, 1)) if scheme == "systematic" else rs.random_sample((T, N))
    return (i + v) / N


@pytest.mark.parametrize("scheme", ["systematic", "stratified"])
@pytest.mark.parametrize("name,N", [("k/svm_prior_poyiadjis_N_1000_d", None), ("k/garch_optimal_poyiadjis_N_1000_d", 5000),
                                    ("k/lgssm_optimal_nemeth_200_lambduh0.8", 777)])
def test_injected_scheme_uniforms_f64(name, N, scheme):
    import sgmcmc_ssm_b200 as sg
    c = C.case(name)
    model, kernel, pf = C.parse_kernel_case(name)
    K = po.make_kernel(model, kernel, C.theta_dict(model, c["theta"]))
    N = int(c["N"]) if N is None else N
    T = c["obs"].shape[0]
    rs = np.random.RandomState(int(c["seed"]) + 1)
    z0, z = rs.normal(size=N), rs.normal(size=(T, N))
    u = _scheme_uniforms(rs, scheme, T, N)
    kw = dict(t1=int(c["t1"]), tL=int(c["tL"]), weights=c["weights"], prior_mean=float(c["prior_mean"]),
              prior_var=float(c["prior_var"]))
    opts = C.case_opts(c)
    ref = po.buffered_pf(pf, c["obs"], K, N, K.score, K.p, po.InjectedStream(u.ravel(), np.concatenate([z0, z.ravel()])),
                         save_all=True, **kw, **opts)
    items = sg.PFItems().add(c["obs"], c["theta"], **kw)
    res = sg.run_pf(model, kernel, pf, items, N, dtype="f64", rng="injected", resample=scheme,
                    injected=dict(z0=z0, u=u, z=z), want=("anc",), **opts)
    np.testing.assert_array_equal(res.tensor("anc")[0].cpu().numpy(), np.array(ref["trace"]["ancestors"]))
    np.testing.assert_allclose(res.grad[0], po.average_statistic(ref), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(res.loglik[0], ref["loglikelihood_estimate"], rtol=1e-9, atol=1e-10)


def _svm_items(sg, B):
    c = C.case("k/svm_prior_poyiadjis_N_1000_d")
    obs = c["obs"].reshape(-1)
    return sg.PackedItems(np.tile(obs, B), np.full(B, obs.shape[0]), np.full(B, int(c["t1"])), np.full(B, int(c["tL"])),
                          None, None, c["theta"], 0.0, 10.0), obs.shape[0]


@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("N", [1000, 4096, 70000])
@pytest.mark.parametrize("scheme", ["systematic", "stratified"])
def test_philox_scheme_structure(scheme, N, dtype):
    import sgmcmc_ssm_b200 as sg
    pk, T = _svm_items(sg, 3)
    res = sg.run_pf("svm", "prior", "poyiadjis_N", pk, N, dtype=dtype, rng="philox", resample=scheme, seed=5, offset=3,
                    want=("anc", "trace_lw"))
    anc = res.tensor("anc").cpu().numpy().astype(np.int64)                  # (B, T, N)
    lw = res.tensor("trace_lw").double().cpu().numpy()                      # (B, T + 1, N)
    # slack: CDF rounding of the kernel's arithmetic type, in units of one child's mass 1 / N
    slack = (1e-9 if dtype == "f64" else 3e-5) * N + 1e-9
    for b in range(anc.shape[0]):
        for t in range(T):
            a = anc[b, t]
            assert np.all(np.diff(a) >= 0)
            w = np.exp(lw[b, t] - lw[b, t].max())
            w /= w.sum()
            if scheme == "systematic":
                counts = np.bincount(a, minlength=N)
                assert np.all(np.abs(counts - N * w) < 1.0 + slack), (scheme, N, dtype, b, t)
            else:
                cdf = np.cumsum(w)
                lo = np.concatenate([[0.0], cdf[:-1]])[a] * N                # the ancestor's interval, in child units
                hi = cdf[a] * N
                i = np.arange(N)
                assert np.all(hi >= i - slack) and np.all(lo <= i + 1 + slack), (scheme, N, dtype, b, t)


@pytest.mark.parametrize("scheme", ["systematic", "stratified"])
def test_philox_scheme_estimator_matches_multinomial(scheme):
    import sgmcmc_ssm_b200 as sg
    R = 1024
    pk, _ = _svm_items(sg, R)
    a = sg.run_pf("svm", "prior", "poyiadjis_N", pk, 1000, dtype="f32", rng="philox", resample="multinomial_sorted", seed=8, offset=1).grad
    b = sg.run_pf("svm", "prior", "poyiadjis_N", pk, 1000, dtype="f32", rng="philox", resample=scheme, seed=9, offset=1).grad
    se = np.sqrt(a.var(axis=0, ddof=1) / R + b.var(axis=0, ddof=1) / R)
    assert np.all(np.abs(a.mean(axis=0) - b.mean(axis=0)) <= 4.5 * se), (a.mean(axis=0), b.mean(axis=0), se)
    assert np.all(b.std(axis=0) <= 1.15 * a.std(axis=0))
