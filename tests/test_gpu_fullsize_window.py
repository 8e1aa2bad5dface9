"""Full-window oracle check at the headline size: ONE S = 40 / B = 10 window (60 time steps) at N = 2^16 particles per
model / proposal kernel, f64 arithmetic with the oracle's recorded randoms injected (reference semantics: iid resampling
uniforms, np.random.choice == searchsorted(cdf, u, 'right'); pf.py:7-38, 138-181).  3.9 million ancestor draws per case:
ancestors bit-exact at every step, gradient rtol 1e-8, log-likelihood 1e-9.  The oracle needs ~3 s per case."""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests import _cases as C

pytestmark = pytest.mark.gpu
N = 1 << 16


@pytest.mark.parametrize("model,kernel,src", [("svm", "prior", "k/svm_prior_poyiadjis_N_1000_d"),
                                              ("lgssm", "optimal", "k/lgssm_optimal_poyiadjis_N_1000_d"),
                                              ("lgssm", "prior", "k/lgssm_prior_poyiadjis_N_1000_d"),
                                              ("garch", "optimal", "k/garch_optimal_poyiadjis_N_1000_d"),
                                              ("garch", "prior", "k/garch_prior_poyiadjis_N_1000_d")])
def test_sixty_step_window_at_headline_size_f64_injected(model, kernel, src):
    import sgmcmc_ssm_b200 as sg
    c = C.case(src)
    theta = C.theta_dict(model, c["theta"])
    K = po.make_kernel(model, kernel, theta)
    rs = np.random.RandomState(99)
    base = c["obs"].reshape(-1)
    obs = np.concatenate([base, base[::-1], base])[:60] * (1.0 + 0.1 * rs.normal(size=60))     # 60 steps of model-scale data
    t1, tL = 10, 50
    weights = np.linspace(200.0, 300.0, 40)
    pm, pv = float(c["prior_mean"]), float(c["prior_var"])
    rec = po.LegacyStream(4321, record=True)
    ref = po.buffered_pf("poyiadjis_N", obs, K, N, K.score, K.p, rec, t1=t1, tL=tL, weights=weights, prior_mean=pm,
                         prior_var=pv, save_all=True)
    parts = po.split_events(rec.events, N)
    items = sg.PFItems().add(obs, c["theta"], t1=t1, tL=tL, weights=weights, prior_mean=pm, prior_var=pv)
    res = sg.run_pf(model, kernel, "poyiadjis_N", items, N, dtype="f64", rng="injected", resample="multinomial",
                    injected=dict(z0=parts["z0"], u=parts["u"], z=parts["z"]), want=("anc", "x", "lw"))
    anc = res.tensor("anc")[0].cpu().numpy()
    ref_anc = np.array(ref["trace"]["ancestors"])
    assert anc.shape == ref_anc.shape == (60, N)
    np.testing.assert_array_equal(anc, ref_anc)
    x = res.tensor("x")[0].cpu().numpy()
    np.testing.assert_allclose(x, ref["x_t"].reshape(x.shape), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res.tensor("lw")[0].cpu().numpy(), ref["log_weights"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res.grad[0], po.average_statistic(ref), rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(res.loglik[0], ref["loglikelihood_estimate"], rtol=1e-9, atol=1e-9)
