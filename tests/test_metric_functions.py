"""Evaluation-tick metric factories around the PF path (SURVEY 8(f1)/(f2) callers; reference
`sgmcmc_ssm/metric_functions.py:8-261, 362-417`).  CPU part: record protocol and call contract against a stub
sampler (the expected records were produced by the unmodified reference module on the same stub).  GPU part: the
factories drive the real samplers and reproduce the direct sampler calls on the same numpy stream."""
import numpy as np
import pytest

import sgmcmc_ssm_b200  # noqa: F401
from sgmcmc_ssm_b200 import metric_functions as mf


class _Params:
    def __init__(self):
        self.vector = np.array([0.9, 1.3])

    A = property(lambda self: np.array([[self.vector[0]]]))
    LQinv = property(lambda self: np.array([[self.vector[1]]]))


class _Stub:
    def __init__(self):
        self.parameters, self.calls = _Params(), []

    def noisy_logjoint(self, return_loglike=False, **kw):
        self.calls.append(("logjoint", return_loglike, kw))
        return dict(logjoint=-3.0, loglikelihood=-2.0)

    def predictive_loglikelihood(self, **kw):
        self.calls.append(("pred", kw))
        return np.arange(11.0) if kw.get("kind") == "pf" else 7.0

    def foo(self, a=1):
        return 2 * a


def test_logjoint_metric_records_and_single_call():
    s = _Stub()
    out = mf.noisy_logjoint_loglike_metric(metric_name_prefix="x_", kind="pf", N=50)(s)
    assert out == [dict(variable="sampler", metric="x_noisy_logjoint", value=-3.0),
                   dict(variable="sampler", metric="x_noisy_loglikelihood", value=-2.0)]
    assert s.calls == [("logjoint", True, dict(kind="pf", N=50))]


def test_predictive_metric_forwards_the_horizon_as_lag_like_the_reference():
    s = _Stub()
    out = mf.noisy_predictive_logjoint_loglike_metric(3, kind="pf", N=5)(s)
    assert [r["metric"] for r in out] == ["%d_pred_loglikelihood" % k for k in range(4)]
    assert [r["value"] for r in out] == [0.0, 1.0, 2.0, 3.0]
    assert s.calls == [("pred", dict(lag=3, kind="pf", N=5))]
    out = mf.noisy_predictive_logjoint_loglike_metric(3, kind="marginal", metric_name_prefix="p")(s)
    assert out == [dict(variable="sampler", metric="p3_pred_loglikelihood", value=7.0)]
    with pytest.raises(IndexError):                       # 10 horizons are all the pf branch computes by default
        mf.noisy_predictive_logjoint_loglike_metric(11, kind="pf")(s)


def test_parameter_samples_and_metrics():
    s = _Stub()
    rec = mf.sample_function_parameters(["A", "LQinv"], ["a", None])(s)
    assert [r["variable"] for r in rec] == ["a", "LQinv"] and rec[0]["value"][0, 0] == 0.9
    rec[0]["value"][0, 0] = 5.0                           # a copy, not a view of the live parameters
    assert s.parameters.A[0, 0] == 0.9
    one = np.array([[1.0]])
    out = mf.metric_function_parameters(["A", "LQinv", "A", "A"], [one] * 4, ["mse", "logmse", "rmse", "mae"])(s)
    np.testing.assert_allclose([r["value"] for r in out], [0.01, np.log10(0.09), 0.1, 0.1], rtol=1e-12)
    assert [r["metric"] for r in out] == ["mse", "logmse", "rmse", "mae"]
    with pytest.raises(ValueError):
        mf.construct_metric_function("nope")
    with pytest.raises(ValueError):
        mf.metric_function_parameters(["A"], [one, one], ["mse"])
    with pytest.raises(ValueError):
        mf.sample_function_parameters(["A"], ["a", "b"])
    with pytest.raises(NotImplementedError):
        mf.metric_function_parameters(["A"], [one], ["mse"], criteria=[min])


def test_metric_from_sampler_and_running_average():
    s = _Stub()
    assert mf.metric_function_from_sampler("foo", a=4)(s) == dict(variable="sampler", metric="foo", value=8)
    assert mf.metric_function_from_sampler("foo", metric_name="m", return_variable_name="v")(s)["metric"] == "m"
    with pytest.raises(ValueError):
        mf.metric_function_from_sampler("missing")(s)
    avg = mf.average_input_decorator(mf.metric_function_parameter("A", np.array([[1.0]]), "mae"))
    r1 = avg(s)
    s.parameters.vector = np.array([0.5, 1.0])
    r2 = avg(s)
    assert r1["variable"] == r2["variable"] == "avg_A"
    np.testing.assert_allclose([r1["value"], r2["value"]], [0.1, 0.3], rtol=1e-12)      # means 0.9, then 0.7
    np.testing.assert_array_equal(s.parameters.vector, [0.5, 1.0])                      # restored
    lst = mf.average_input_decorator(mf.sample_function_parameters(["A", "LQinv"]))(s)
    assert [r["variable"] for r in lst] == ["avg_A", "avg_LQinv"]


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_metric_factories_drive_the_real_samplers(model):
    from tests.test_host_logic import MODELS
    from sgmcmc_ssm_b200.models.svm import generate_svm_data
    from sgmcmc_ssm_b200.models.lgssm import generate_lgssm_data
    from sgmcmc_ssm_b200.models.garch import generate_garch_data
    truth, _, Sampler = MODELS[model]
    gen = dict(svm=generate_svm_data, lgssm=generate_lgssm_data, garch=generate_garch_data)[model]
    np.random.seed(3)
    p = truth()
    data = gen(T=200, parameters=p)
    s = Sampler(n=1, m=1, observations=data["observations"], parameters=p)
    kw = dict(kind="pf", N=64, subsequence_length=30, buffer_length=5, minibatch_size=2, rng="injected", dtype="f64")
    np.random.seed(11)
    rec = mf.noisy_logjoint_loglike_metric(**kw)(s)
    np.random.seed(11)
    direct = s.noisy_logjoint(return_loglike=True, **kw)
    assert rec[0]["value"] == direct["logjoint"] and rec[1]["value"] == direct["loglikelihood"]
    np.testing.assert_allclose(rec[0]["value"] - rec[1]["value"], s.prior.logprior(s.parameters), rtol=1e-10)
    np.random.seed(12)
    rec = mf.noisy_predictive_logjoint_loglike_metric(2, **kw)(s)
    np.random.seed(12)
    direct = s.predictive_loglikelihood(**kw)              # default 10 horizons, as the reference's metric gets
    assert len(rec) == 3 and direct.shape == (11,)
    np.testing.assert_array_equal([r["value"] for r in rec], direct[:3])
    assert all(np.isfinite(r["value"]) for r in rec)
    one = mf.metric_function_from_sampler("noisy_loglikelihood", **kw)(s)
    assert np.isfinite(one["value"]) and one["metric"] == "noisy_loglikelihood"


def test_average_input_decorator_with_a_real_sampler():
    """The decorator reads / assigns `parameters.vector`: works on the real Parameters containers (not only on a stub)
    and restores the sampler's own parameters afterwards."""
    from sgmcmc_ssm_b200 import metric_functions as mf
    from sgmcmc_ssm_b200.models.svm import SVMSampler, SVMParameters
    p = SVMParameters(A=np.eye(1) * 0.9, LQinv=np.eye(1) * 1.5, LRinv=np.eye(1) * 2.0)
    s = SVMSampler(n=1, m=1, observations=np.zeros((10, 1)), parameters=p)
    avg = mf.average_input_decorator(mf.metric_function_parameter("A", np.array([[1.0]]), "mae"))
    out1 = avg(s)
    s.parameters.A = np.eye(1) * 0.5
    out2 = avg(s)
    rec1 = out1 if isinstance(out1, dict) else out1[0]
    rec2 = out2 if isinstance(out2, dict) else out2[0]
    assert rec1["variable"].startswith("avg_")
    np.testing.assert_allclose(rec1["value"], 0.1, atol=1e-12)
    np.testing.assert_allclose(rec2["value"], 0.3, atol=1e-12)          # running mean of A: (0.9 + 0.5) / 2 = 0.7
    np.testing.assert_allclose(s.parameters.A, [[0.5]])                  # restored
    np.testing.assert_allclose(s.parameters.vector, [0.5, 1.5, 2.0])
