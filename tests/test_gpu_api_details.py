"""API details of the drop-in surface that the reference offers next to the gradient itself:
  * `save_all=True`: all_x_t / all_log_weights / all_statistics / all_loglikelihood_estimate after every step
    (buffered_smoother.py:128-147), against the UNMODIFIED reference's traces (fixtures k/*_poyiadjis_N_64_d);
  * `pf_latent_var_distr(lag=0, pf='filter')`: filtered marginals, against the oracle;
  * a PFResult left un-waited when the next call reuses the device staging buffer raises instead of returning stale data."""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests import _cases as C
from tests.test_host_logic import MODELS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rng", ["injected", "philox"])
@pytest.mark.parametrize("name", [n for n in C.case_names("k") if n.endswith("poyiadjis_N_64_d")])
def test_save_all_returns_every_trace_of_the_reference(name, rng):
    from sgmcmc_ssm_b200.particle_filters.buffered_smoother import buffered_pf_wrapper
    from sgmcmc_ssm_b200.particle_filters import statistics as S
    import sgmcmc_ssm_b200 as sg
    c = C.case(name)
    model, kernel, pf = C.parse_kernel_case(name)
    make, _, Sampler = MODELS[model]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    K = sampler.message_helper._get_kernel(kernel)
    score = {"svm": S.svm_complete_data_loglike_gradient, "lgssm": S.lgssm_complete_data_loglike_gradient,
             "garch": S.garch_complete_data_loglike_gradient}[model]
    np.random.seed(int(c["seed"]))
    sg.set_seed(3)
    out = buffered_pf_wrapper(pf=pf, observations=c["obs"], parameters=sampler.parameters, N=int(c["N"]), kernel=K,
                              additive_statistic_func=score, statistic_dim=c["statistics"].shape[1], t1=int(c["t1"]),
                              tL=int(c["tL"]), weights=c["weights"], prior_mean=float(c["prior_mean"]),
                              prior_var=float(c["prior_var"]), save_all=True, rng=rng, dtype="f64")
    T = c["obs"].shape[0]
    assert out["all_x_t"].shape == c["all_x_t"].shape and out["all_statistics"].shape == c["all_statistics"].shape
    assert out["all_loglikelihood_estimate"].shape == (T + 1,) and out["all_log_weights"].shape == c["all_log_weights"].shape
    # the traces end in the returned final state
    np.testing.assert_allclose(out["all_statistics"][-1], out["statistics"], rtol=1e-12)
    np.testing.assert_allclose(out["all_loglikelihood_estimate"][-1], out["loglikelihood_estimate"], rtol=1e-10)
    np.testing.assert_allclose(out["all_x_t"][-1], out["x_t"], rtol=1e-12)
    if rng == "injected":
        np.testing.assert_allclose(out["all_x_t"], c["all_x_t"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["all_log_weights"], c["all_log_weights"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["all_statistics"], c["all_statistics"], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(out["all_loglikelihood_estimate"], c["all_loglik"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("model", ["svm", "garch"])
def test_filtered_latent_marginals_lag_zero(model):
    c = C.case("h/latent_{0}_default".format(model))
    make, _, Sampler = MODELS[model]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    theta = C.theta_dict(model, c["theta"])
    N, t1, tL = 400, int(c["t1"]), int(c["tL"])
    kernel = "prior" if model == "svm" else "optimal"
    K = po.make_kernel(model, kernel, theta)
    pm, pv = po.prior_moments(model, theta, None)
    stat = po.garch_sufficient_statistics if model == "garch" else po.gaussian_sufficient_statistics
    ref = po.buffered_pf("filter", c["obs"], K, N, lambda xa, xn, y: stat(xa, xn), 3, po.LegacyStream(77), t1=t1, tL=tL,
                         prior_mean=pm, prior_var=pv, elementwise_statistic=True)
    avg = np.reshape(ref["statistics"], (-1, 3))
    np.random.seed(77)
    mean, cov = sampler.message_helper.pf_latent_var_distr(observations=c["obs"], parameters=sampler.parameters, lag=0,
                                                            subsequence_start=t1, subsequence_end=tL, pf="filter", N=N,
                                                            rng="injected", dtype="f64")
    assert mean.shape == (tL - t1, 1) and cov.shape == (tL - t1, 1, 1)
    np.testing.assert_allclose(mean[:, 0], avg[:, 0], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(cov[:, 0, 0], avg[:, 1] - avg[:, 0] ** 2, rtol=1e-7, atol=1e-9)
    with pytest.raises(ValueError):
        sampler.message_helper.pf_latent_var_distr(observations=c["obs"], parameters=sampler.parameters, lag=0, pf="poyiadjis_N")


def test_unwaited_result_is_not_silently_overwritten():
    import sgmcmc_ssm_b200 as sg
    th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
    it = sg.PFItems().add(np.array([0.3, -0.2, 0.5, 0.1]), th, prior_mean=0.0, prior_var=10.0)
    first = sg.run_pf("svm", "prior", "poyiadjis_N", it, 500, sync=False)
    second = sg.run_pf("svm", "prior", "poyiadjis_N", it, 500)
    assert np.all(np.isfinite(second.grad))
    with pytest.raises(RuntimeError, match="staging buffer"):
        first.wait()
