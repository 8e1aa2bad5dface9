"""GPU parity of the k-step-ahead predictive log-likelihood: pf = 'filter' with the `logsumexp`
statistic (pf.py:40-82), models svm / lgssm / garch (svm/helper.py:187-247, 352-395; lgssm/helper.py:1048-1087,
1281-1336; garch/helper.py:374-412), sampler level sgmcmc_sampler.py:94-126.

f64 + injected randoms (the reference's legacy numpy stream, consumed in its order): the CUDA path must land on the
oracle AND on the golden outputs of the unmodified reference (tests/golden/pred_cases.npz) to rtol 1e-8."""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests import _cases as C
from tests.test_host_logic import MODELS

pytestmark = pytest.mark.gpu


def _names(section):
    z = C.load("pred_cases.npz")
    return sorted({"/".join(k.split("/")[:2]) for k in z.files if k.startswith(section + "/")})


def _case(name):
    z = C.load("pred_cases.npz")
    return {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "/")}


def _theta_vec(model, c):
    th = np.zeros(12)
    th[:len(c["theta"])] = c["theta"]
    if model != "garch":
        th[10], th[11] = float(c["Q"]), float(c["R"])
    return th


@pytest.mark.parametrize("per_horizon", [False, True])
@pytest.mark.parametrize("name", _names("p"))
def test_predictive_kernel_level_f64(name, per_horizon):
    _kernel_level(name, per_horizon)


@pytest.mark.parametrize("K", [8, 10, 14])
@pytest.mark.parametrize("name", ["p/svm_prior_c", "p/lgssm_optimal_c", "p/garch_optimal_c", "p/garch_prior_a"])
def test_predictive_kernel_level_long_horizons(name, K):
    """The reference's default horizon (num_steps_ahead = 10, sgmcmc_sampler.py:61) and the largest supported one
    (SGM_PRED_MAX_STEPS = 14) against the oracle; includes horizons that run past the end of the buffer."""
    if name not in _names("p"):
        pytest.skip("no such stored case")
    _kernel_level(name, False, K=K)
    _kernel_level(name, True, K=K)


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_predictive_default_horizon_and_limit(model):
    """num_steps_ahead left at the reference's default (10) -- what the drivers' predictive metric ends up running --
    against the unmodified reference under the same seed; horizons above SGM_PRED_MAX_STEPS are refused."""
    make_params, _, Sampler = MODELS[model]
    s = Sampler(n=1, m=1, observations=_case("ps/" + model)["obs"], parameters=make_params())
    np.random.seed(6)
    out = s.predictive_loglikelihood(kind="pf", subsequence_length=30, minibatch_size=2, buffer_length=4, N=100,
                                     dtype="f64", rng="injected")
    assert out.shape == (11,)
    np.testing.assert_allclose(out, C.load("pred_cases.npz")["ps10/{0}/out".format(model)], rtol=1e-8, atol=1e-7)
    with pytest.raises(NotImplementedError):
        s.predictive_loglikelihood(kind="pf", num_steps_ahead=15, subsequence_length=30, N=128)


def _kernel_level(name, per_horizon, K=None):
    import sgmcmc_ssm_b200 as sg
    c = _case(name)
    model, kernel = name.split("/")[1].split("_")[:2]
    theta = C.theta_dict(model, c["theta"])
    N, t1, tL = int(c["N"]), int(c["t1"]), int(c["tL"])
    stored = K is None
    K = int(c["K"]) if stored else K
    rec = po.LegacyStream(int(c["seed"]), record=True)
    ref = po.pf_predictive_loglikelihood_estimate(model, c["obs"], theta, float(c["Q"]), float(c["R"]), rec, num_steps_ahead=K,
                                                  subsequence_start=t1, subsequence_end=tL, N=N, kernel=kernel,
                                                  per_horizon=per_horizon)
    if stored and not per_horizon:
        np.testing.assert_allclose(ref, c["out"], rtol=1e-11, atol=1e-11)
    parts = po.split_events_pred(rec.events, N)
    T = c["obs"].shape[0]
    zp = np.zeros((1, T, 16, N))            # SGM_PRED_SLOTS horizon slots
    for t, blk in enumerate(parts["zp"]):
        zp[0, t, :blk.shape[0]] = blk
    pm, pv = po.prior_moments(model, theta, None)
    items = sg.PFItems().add(c["obs"], _theta_vec(model, c), t1=t1, tL=tL, prior_mean=pm, prior_var=pv)
    res = sg.run_pf(model, kernel, "filter", items, N, dtype="f64", rng="injected", resample="multinomial", stat_kind="pred",
                    num_steps_ahead=K, per_horizon=per_horizon, injected=dict(z0=parts["z0"], u=parts["u"], z=parts["z"], zp=zp))
    out = res.grad[0][:K + 1].copy()
    out[0] = res.loglik[0]
    np.testing.assert_allclose(out, ref, rtol=1e-8, atol=1e-8)


@pytest.mark.parametrize("name", _names("p"))
def test_predictive_helper_api_replays_the_reference_stream(name):
    """np.random.seed(s) + the reference-shaped call: the drop-in consumes the legacy stream like the reference."""
    c = _case(name)
    model, kernel = name.split("/")[1].split("_")[:2]
    make_params, _, Sampler = MODELS[model]
    p = make_params()
    s = Sampler(n=1, m=1, observations=np.asarray(c["obs"]).reshape(-1, 1), parameters=p)
    np.random.seed(int(c["seed"]))
    out = s.message_helper.pf_predictive_loglikelihood_estimate(
        np.asarray(c["obs"]).reshape(-1, 1), p, num_steps_ahead=int(c["K"]), subsequence_start=int(c["t1"]),
        subsequence_end=int(c["tL"]), N=int(c["N"]), kernel=kernel, dtype="f64", rng="injected")
    np.testing.assert_allclose(out, c["out"], rtol=1e-8, atol=1e-8)
    with pytest.raises(ValueError):
        s.message_helper.pf_predictive_loglikelihood_estimate(np.asarray(c["obs"]).reshape(-1, 1), p, pf="poyiadjis_N")


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_predictive_sampler_level(model):
    """sampler.predictive_loglikelihood(kind='pf') vs the reference (sgmcmc_sampler.py:94-126), same seed."""
    c = _case("ps/" + model)
    make_params, _, Sampler = MODELS[model]
    s = Sampler(n=1, m=1, observations=c["obs"], parameters=make_params())
    np.random.seed(5)
    out = s.predictive_loglikelihood(kind="pf", num_steps_ahead=4, subsequence_length=20, minibatch_size=3, buffer_length=4,
                                     N=200, dtype="f64", rng="injected")
    np.testing.assert_allclose(out, c["out"], rtol=1e-8, atol=1e-7)


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_predictive_f32_philox_close_to_f64(model):
    """Device randoms, f32: the per-horizon estimate at N = 20000 agrees with an independent f64 run within
    Monte-Carlo error (both are consistent estimators of the same quantity); horizon 0 is the log-likelihood."""
    c = _case("p/{0}_{1}_c".format(model, "prior" if model == "svm" else "optimal"))
    make_params, _, Sampler = MODELS[model]
    p = make_params()
    h = Sampler(n=1, m=1, observations=np.asarray(c["obs"]).reshape(-1, 1), parameters=p).message_helper
    kw = dict(num_steps_ahead=5, subsequence_start=4, subsequence_end=20, N=20000, per_horizon=True)
    a = h.pf_predictive_loglikelihood_estimate(np.asarray(c["obs"]).reshape(-1, 1), p, dtype="f32", rng="philox", seed=1, **kw)
    b = h.pf_predictive_loglikelihood_estimate(np.asarray(c["obs"]).reshape(-1, 1), p, dtype="f64", rng="philox", seed=2, **kw)
    assert np.all(np.isfinite(a)) and np.all(np.isfinite(b))
    assert np.all(np.abs(a - b) <= 0.05 * np.abs(b) + 0.5), (a, b)
