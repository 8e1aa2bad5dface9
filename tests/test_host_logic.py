"""CPU tests of the host-side mirror of the reference API (no GPU): parameter containers, priors,
subsequence sampling, synthetic-data generators, and that the C-ABI library loads and exports every
symbol include/sgmpf.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests import _cases as C
from oracle import pf_oracle as po

import sgmcmc_ssm_b200  # noqa: F401
from sgmcmc_ssm_b200.models.svm import SVMParameters, SVMPrior, SVMSampler, generate_svm_data
from sgmcmc_ssm_b200.models.lgssm import LGSSMParameters, LGSSMPrior, LGSSMSampler, generate_lgssm_data
from sgmcmc_ssm_b200.models.garch import GARCHParameters, GARCHPrior, GARCHSampler, generate_garch_data
from sgmcmc_ssm_b200.sgmcmc_sampler import random_subsequence_and_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def svm_params():
    return SVMParameters(A=np.eye(1) * 0.95, LQinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.5)),
                         LRinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.5)))


def lgssm_params():
    return LGSSMParameters(A=np.eye(1) * 0.9, C=np.eye(1), LQinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.1)),
                           LRinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 1.0)))


def garch_params():
    log_mu, logit_phi, logit_lambduh = GARCHParameters.convert_alpha_beta_gamma(0.1, 0.8, 0.05)
    return GARCHParameters(log_mu=log_mu, logit_phi=logit_phi, logit_lambduh=logit_lambduh,
                           LRinv=np.linalg.cholesky(np.linalg.inv(np.eye(1) * 0.3 ** 2)))


MODELS = dict(svm=(svm_params, SVMPrior, SVMSampler), lgssm=(lgssm_params, LGSSMPrior, LGSSMSampler),
              garch=(garch_params, GARCHPrior, GARCHSampler))


def test_cabi_library_exports_every_declared_symbol():
    from sgmcmc_ssm_b200 import _native as nat
    lib = nat.load()
    header = open(os.path.join(ROOT, "include", "sgmpf.h")).read()
    declared = set(re.findall(r"\b(sgm_[a-z_]+)\s*\(", header))
    assert declared == set(nat.EXPORTS), declared ^ set(nat.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.sgm_version() == 200
    assert lib.sgm_stat_dim(nat.MODEL["svm"], 0) == 3 and lib.sgm_stat_dim(nat.MODEL["garch"], 0) == 4
    assert lib.sgm_state_dim(nat.MODEL["garch"]) == 2
    # descriptor validation happens before any CUDA call
    d = nat.SgmPfDesc()
    assert lib.sgm_pf_workspace_bytes(ctypes.byref(d)) == 0
    assert lib.sgm_pf_run(ctypes.byref(d), None) == -1 and b"size mismatch" in lib.sgm_last_error()


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import sgmcmc_ssm_b200 as sg
    items = sg.PFItems().add(np.zeros(4), [0.9, 1, 1, 1, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sg.run_pf("svm", "prior", "poyiadjis_N", items, 16)


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_prior_matches_reference(model):
    """default prior hyper-parameters, logprior and grad_logprior vs the reference (s/prior_<model>)."""
    c = C.case("s/prior_" + model)
    make, Prior, _ = MODELS[model]
    params = make()
    np.testing.assert_allclose(np.concatenate([np.ravel(params.var_dict[k]) for k in sorted(params.var_dict)]),
                               c["param_values"], rtol=0, atol=1e-15)
    prior = Prior.generate_default_prior(n=1, m=1)
    for k, v in c.items():
        if k.startswith("hyper_"):
            np.testing.assert_allclose(prior.hyperparams[k[6:]], v, rtol=1e-14)
    np.testing.assert_allclose(prior.logprior(params), c["logprior"], rtol=1e-12)
    g = prior.grad_logprior(params)
    assert sorted(g) == [str(k) for k in c["grad_keys"]]
    np.testing.assert_allclose(np.concatenate([np.ravel(g[k]) for k in sorted(g)]), c["grad_values"], rtol=1e-12)


@pytest.mark.parametrize("name", [n for n in C.case_names("s") if n.startswith("s/subseq_")])
def test_random_subsequence_and_weights_matches_reference(name):
    c = C.case(name)
    style = name.rsplit("_", 1)[1]
    np.random.seed(int(c["seed"]))
    s, e, w = random_subsequence_and_weights(S=int(c["S"]), T=int(c["T"]), partition_style=style)
    assert (s, e) == (int(c["start"]), int(c["end"]))
    np.testing.assert_array_equal(w, c["weights"])
    np.testing.assert_array_equal(po.subsequence_weights(int(c["S"]), int(c["T"]), s, style), c["weights"])


def test_data_generators_reproduce_reference_streams():
    """np.random.seed(2024) then svm, lgssm, garch T=200 -- the series stored with the sampler cases."""
    np.random.seed(2024)
    for model, gen in (("svm", generate_svm_data), ("lgssm", generate_lgssm_data), ("garch", generate_garch_data)):
        data = gen(T=200, parameters=MODELS[model][0]())
        np.testing.assert_allclose(data["observations"], C.case("s/sgld_" + model)["obs"], rtol=0, atol=1e-14)


def test_parameter_containers():
    p = svm_params()
    assert list(p.var_dict) == ["A", "LQinv_vec", "LRinv_vec"] and p.dim == {"n": 1, "m": 1}
    np.testing.assert_allclose(p.Qinv, p.LQinv ** 2 + 1e-16)
    q = SVMParameters(**SVMParameters.from_vector_to_dict(p.as_vector(), **p.dim))
    np.testing.assert_array_equal(q.as_vector(), p.as_vector())
    c = p.copy()
    c += {k: np.ones_like(v) for k, v in p.var_dict.items()}
    assert c.A[0, 0] == p.A[0, 0] + 1
    # projection: |A| clipped to 0.9999, negative Cholesky diagonal reflected, LGSSM C fixed to identity
    c.A = np.array([[1.7]])
    c.LQinv_vec = np.array([-2.0])
    c.project_parameters()
    assert abs(c.A[0, 0] - 0.9999) < 1e-12 and abs(c.LQinv_vec[0] - 2.0) < 1e-12
    l = lgssm_params()
    l.C = np.array([[3.0]])
    l.project_parameters()
    assert l.C[0, 0] == 1.0 and list(l.var_dict) == ["A", "C", "LQinv_vec", "LRinv_vec"]
    g = garch_params()
    np.testing.assert_allclose([g.alpha[0], g.beta[0], g.gamma[0]], [0.1, 0.8, 0.05], rtol=1e-12)
    assert list(g.var_dict) == ["log_mu", "logit_phi", "logit_lambduh", "LRinv_vec"]
    with pytest.raises(ValueError):
        GARCHParameters.convert_alpha_beta_gamma(0.1, 0.9, 0.2)


def test_sampler_errors_and_kwargs_contract():
    obs = np.zeros((50, 1))
    s = SVMSampler(n=1, m=1, observations=obs, parameters=svm_params())
    with pytest.raises(ValueError, match="Use SGRLD"):
        s.sample_sgld(epsilon=0.1, preconditioner=object(), kind="pf", N=10)
    with pytest.raises(NotImplementedError):
        s.noisy_gradient(kind="marginal")
    with pytest.raises(ValueError, match="Unrecognized kind"):
        s.noisy_gradient(kind="bogus")
    with pytest.raises(NotImplementedError, match="No Default Preconditioner"):
        s.get_iter_step("SGRLD", epsilon=0.1, subsequence_length=10, buffer_length=2)
    with pytest.raises(NotImplementedError, match="not analytic"):
        s.message_helper._get_kernel("optimal")
    with pytest.raises(ValueError, match="Unrecoginized kernel"):
        s.message_helper._get_kernel("nope")
    with pytest.raises(ValueError, match="Unrecognized pf"):
        s.message_helper.pf_gradient_estimate(obs[:10], s.parameters, pf="nope", N=10)
    bad = svm_params()
    bad.A = np.array([[1.5]])
    with pytest.raises(ValueError, match="AR parameter"):
        s.message_helper.pf_gradient_estimate(obs[:10], bad, N=10)
    names, kws = s.get_iter_step("SGLD", epsilon=0.1, subsequence_length=10, buffer_length=2, kind="pf",
                                 pf_kwargs=dict(pf="paris", N=100, Ntilde=3), steps_per_iteration=2)
    assert names == ["sample_sgld", "project_parameters"] * 2 and kws[0]["Ntilde"] == 3 and kws[0]["N"] == 100
    bd = s._random_subsequence_and_buffers(buffer_length=4, subsequence_length=-1)
    assert (bd["subsequence_start"], bd["subsequence_end"], bd["weights"]) == (0, 50, None)


@pytest.mark.parametrize("style", [None, "uniform", "naive"])
@pytest.mark.parametrize("T,S,B", [(1000, 40, 10), (100, 16, 20), (90, 40, 3), (300, 16, -1)])
def test_vectorised_minibatch_windows_equal_the_sequential_draws(style, T, S, B):
    """random_subsequences_packed == M sequential _random_subsequence_and_buffers + _window calls
    (sgmcmc_sampler.py:259-288, 364-374, 1969-2017), including the numpy stream position afterwards."""
    from sgmcmc_ssm_b200.sgmcmc_sampler import random_subsequences_packed, _window
    from sgmcmc_ssm_b200 import engine
    rs = np.random.RandomState(3)
    obs = rs.normal(size=(T, 1))
    s = SVMSampler(n=1, m=1, observations=obs, parameters=svm_params(), partition_style=style)
    M = 37
    np.random.seed(99)
    windows = [_window(obs, s._random_subsequence_and_buffers(buffer_length=B, subsequence_length=S, T=T)) for _ in range(M)]
    after_seq = np.random.random_sample()
    np.random.seed(99)
    arrays = random_subsequences_packed(obs, S, M, B, style)
    after_vec = np.random.random_sample()
    assert after_seq == after_vec
    items = engine.PFItems()
    for w in windows:
        items.add(w["observations"], [1.0, 2.0], t1=w["subsequence_start"], tL=w["subsequence_end"], weights=w["weights"])
    ref = items.pack()
    pk = engine.PackedItems(theta=[1.0, 2.0], prior_mean=0.0, prior_var=1.0, **arrays)
    for name in ("obs_flat", "T_buf", "t1", "tL", "wts_flat", "wts_off"):
        np.testing.assert_array_equal(getattr(pk, name), getattr(ref, name), err_msg=name)
    np.testing.assert_array_equal(pk.theta, ref.theta)
    # packing only a rank's shard == slicing the full batch (and the numpy stream advances identically)
    np.random.seed(99)
    shard = engine.PackedItems(theta=[1.0, 2.0], prior_mean=0.0, prior_var=1.0, **random_subsequences_packed(obs, S, M, B, style, 5, 21))
    assert np.random.random_sample() == after_seq
    full_slice = pk.slice(5, 21)
    for name in ("obs_flat", "T_buf", "t1", "tL", "wts_flat", "wts_off"):
        np.testing.assert_array_equal(getattr(shard, name), getattr(full_slice, name), err_msg=name)
    # shard slicing keeps every item intact
    lo, hi = 5, 21
    sl = pk.slice(lo, hi)
    off = np.concatenate([[0], np.cumsum(pk.T_buf)])
    np.testing.assert_array_equal(sl.obs_flat, pk.obs_flat[off[lo]:off[hi]])
    np.testing.assert_array_equal(sl.T_buf, pk.T_buf[lo:hi])
    for b in range(hi - lo):
        n = int(sl.tL[b] - sl.t1[b])
        np.testing.assert_array_equal(sl.wts_flat[sl.wts_off[b]:sl.wts_off[b] + n], windows[lo + b]["weights"][:n])


def test_python_descriptor_mirror_has_the_size_the_library_expects():
    """sgm_pf_run checks struct_bytes against its own sizeof(sgm_pf_desc): with the mirror's size filled in the
    validation must get past the size check (and then reject the empty descriptor for another reason)."""
    from sgmcmc_ssm_b200 import _native as nat
    lib = nat.load()
    d = nat.SgmPfDesc()
    d.struct_bytes = ctypes.sizeof(nat.SgmPfDesc)
    assert lib.sgm_pf_run(ctypes.byref(d), None) == -1
    assert b"size mismatch" not in lib.sgm_last_error()
    d.model, d.kernel = nat.MODEL["svm"], nat.KERNEL["optimal"]
    d.n_items = d.n_particles = 1
    assert lib.sgm_pf_run(ctypes.byref(d), None) == -2 and b"optimal" in lib.sgm_last_error()   # NotImplementedError


def test_python_constants_mirror_the_header():
    """Enumerators and #defines of include/sgmpf.h against the Python mirror (a drifted constant would silently select
    another smoother / statistic)."""
    from sgmcmc_ssm_b200 import _native as nat
    text = open(os.path.join(ROOT, "include", "sgmpf.h")).read()
    defines = {k: int(v) for k, v in re.findall(r"#define\s+(SGM_[A-Z0-9_]+)\s+(\d+)", text)}
    assert defines["SGM_PRED_SLOTS"] == nat.PRED_SLOTS and defines["SGM_PRED_MAX_STEPS"] == nat.PRED_MAX_STEPS
    enums = {}
    for body in re.findall(r"enum\s*\{(.*?)\}", text, flags=re.S):
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        for k, v in re.findall(r"(SGM_[A-Z0-9_]+)\s*=\s*(-?\d+)", body):
            enums[k] = int(v)
    expect = {"SGM_STAT_SCORE": nat.STAT["score"], "SGM_STAT_PRED": nat.STAT["pred"],
              "SGM_PF_NEMETH": nat.PF["nemeth"], "SGM_PF_POY_N2": nat.PF["poyiadjis_N2"], "SGM_PF_PARIS": nat.PF["paris"],
              "SGM_PF_FILTER": nat.PF["filter"], "SGM_MODEL_SVM": nat.MODEL["svm"], "SGM_MODEL_LGSSM": nat.MODEL["lgssm"],
              "SGM_RNG_PHILOX": nat.RNG["philox"], "SGM_RNG_INJECTED": nat.RNG["injected"],
              "SGM_RESAMPLE_MULTINOMIAL": nat.RESAMPLE["multinomial"],
              "SGM_RESAMPLE_MULTINOMIAL_SORTED": nat.RESAMPLE["multinomial_sorted"],
              "SGM_N2_TENSOR": nat.N2_MODE["tensor"], "SGM_N2_FP32_PIPE": nat.N2_MODE["fp32_pipe"]}
    for k, v in expect.items():
        assert enums[k] == v, k
    assert nat.PF["poyiadjis_N"] == nat.PF["nemeth"]                  # buffered_smoother.py:175-180: nemeth with lambduh = 1


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the package, the drop-in alias or scripts/ may import it, and in
    bench.py only the CPU legs (cpu_baseline / --impl reference) do."""
    pkg = os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200")
    for base in (pkg, os.path.join(ROOT, "scripts")):
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    assert "oracle" not in open(os.path.join(dirpath, f)).read(), os.path.join(dirpath, f)
    src = open(os.path.join(ROOT, "bench.py")).read()
    gpu_arm = src[src.index("def run_gpu("):src.index("def extras(")]
    assert "import pf_oracle" not in gpu_arm and "po." not in gpu_arm.replace("cpu_rate", "")
    assert src.count("from oracle import pf_oracle") == 1 and "def _cpu_one_gradient" in src


def test_prior_gradient_is_evaluated_while_the_filters_run(monkeypatch):
    """noisy_gradient(kind='pf') hands the helper a `while_running` hook (host work between launch and wait); whether
    or not the device layer calls it, the result is grad log-likelihood + grad log-prior, scaled by 1/T, and the
    prior gradient is evaluated exactly once (sgmcmc_sampler.py:427-464)."""
    from sgmcmc_ssm_b200.models.svm import SVMHelper
    calls = []

    def fake_batch(self, windows, parameters, while_running=None, **kwargs):
        if fake_batch.use_hook and while_running is not None:
            while_running()
            while_running()                       # idempotent: a second device call must not re-evaluate it
        return [dict(LRinv_vec=1.0, LQinv_vec=2.0, A=3.0) for _ in windows], None

    monkeypatch.setattr(SVMHelper, "pf_gradient_estimate_batch", fake_batch)
    np.random.seed(0)
    y = np.random.normal(size=(300, 1))
    s = SVMSampler(n=1, m=1, observations=y, parameters=svm_params())
    prior_grad = s.prior.grad_logprior(parameters=s.parameters)
    orig = type(s.prior).grad_logprior
    monkeypatch.setattr(type(s.prior), "grad_logprior", lambda self, **kw: (calls.append(1), orig(self, **kw))[1])
    out = {}
    for use_hook in (True, False):
        fake_batch.use_hook = use_hook
        del calls[:]
        np.random.seed(1)
        out[use_hook] = s.noisy_gradient(kind="pf", N=16, subsequence_length=20, buffer_length=5, minibatch_size=1)
        assert len(calls) == 1
    for k, like in dict(LRinv_vec=1.0, LQinv_vec=2.0, A=3.0).items():
        expect = (np.ravel(prior_grad[k])[0] + like) / 300.0
        np.testing.assert_allclose(np.ravel(out[True][k])[0], expect, rtol=1e-13)
        np.testing.assert_allclose(np.ravel(out[False][k])[0], expect, rtol=1e-13)
