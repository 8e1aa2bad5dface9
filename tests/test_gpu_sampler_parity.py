"""GPU parity, API level: the drop-in Python surface (Helper / Sampler classes) driven exactly like
the reference -- `np.random.seed(s)` then the same method call -- in parity mode
(rng='injected': the host consumes the global numpy legacy stream in the reference's order and
ships the uniforms / normals to the device; dtype='f64').  Expected values come from the unmodified
reference (tests/golden/ref_cases.npz, sections h/ and s/) and from the reference's own stored
golden gradients (tests/golden/svm_replay.npz).

Tolerance: rtol 1e-8 / atol 1e-9 on gradients and parameters, 1e-9 on log-likelihoods (f64 device
arithmetic vs numpy: libm ulps + FMA contraction + different summation trees).
"""
import numpy as np
import pytest

from tests import _cases as C
from tests.test_host_logic import MODELS

pytestmark = pytest.mark.gpu
PARITY = dict(rng="injected", dtype="f64")


def _fm(c):
    if "fm_precision" in c:
        return dict(log_constant=0.0, precision=c["fm_precision"], mean_precision=c["fm_mean_precision"])
    return None


def _helper(model, fm):
    from sgmcmc_ssm_b200.models.svm import SVMHelper
    from sgmcmc_ssm_b200.models.lgssm import LGSSMHelper
    from sgmcmc_ssm_b200.models.garch import GARCHHelper
    return dict(svm=SVMHelper, lgssm=LGSSMHelper, garch=GARCHHelper)[model](n=1, m=1, forward_message=fm)


@pytest.mark.parametrize("name", [n for n in C.case_names("h") if n.startswith("h/grad_") and "paris" not in n])
def test_helper_pf_gradient_estimate(name):
    c = C.case(name)
    _, model, fmk, pf = name.split("/")[1].split("_", 3)
    fm = _fm(c)
    helper = _helper(model, fm)
    np.random.seed(int(c["seed"]))
    grad = helper.pf_gradient_estimate(observations=c["obs"].reshape(-1, 1), parameters=MODELS[model][0](),
                                       subsequence_start=int(c["t1"]), subsequence_end=int(c["tL"]),
                                       weights=c["weights"], pf=pf, N=int(c["N"]), tqdm=None, unknown_kwarg=3,
                                       **PARITY)
    keys = [str(k) for k in c["keys"]]
    assert sorted(grad) == keys
    np.testing.assert_allclose([grad[k] for k in keys], c["values"], rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("name", [n for n in C.case_names("h") if n.startswith("h/loglik_")])
def test_helper_pf_loglikelihood_estimate(name):
    c = C.case(name)
    model = name.split("/")[1].split("_")[1]
    helper = _helper(model, _fm(c))
    np.random.seed(int(c["seed"]))
    ll = helper.pf_loglikelihood_estimate(observations=c["obs"].reshape(-1, 1), parameters=MODELS[model][0](),
                                          subsequence_start=int(c["t1"]), subsequence_end=int(c["tL"]),
                                          pf="poyiadjis_N", N=int(c["N"]), **PARITY)
    np.testing.assert_allclose(ll, c["value"], rtol=1e-9)


@pytest.mark.parametrize("name", [n for n in C.case_names("h") if n.startswith("h/latent_")])
def test_helper_pf_latent_var_distr(name):
    c = C.case(name)
    model = name.split("/")[1].split("_")[1]
    helper = _helper(model, _fm(c))
    np.random.seed(int(c["seed"]))
    mean, cov = helper.pf_latent_var_distr(observations=c["obs"].reshape(-1, 1), parameters=MODELS[model][0](),
                                           subsequence_start=int(c["t1"]), subsequence_end=int(c["tL"]),
                                           pf="poyiadjis_N", N=int(c["N"]), **PARITY)
    assert mean.shape == c["mean"].shape and cov.shape == c["cov"].shape
    np.testing.assert_allclose(mean, c["mean"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(cov, c["cov"], rtol=1e-7, atol=1e-8)


def _vec(d):
    return np.concatenate([np.ravel(d[k]) for k in sorted(d)])


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
@pytest.mark.parametrize("pf", ["poyiadjis_N", "nemeth"])
def test_sampler_noisy_gradient(model, pf):
    c = C.case("s/noisy_grad_{0}_{1}".format(model, pf))
    make, _, Sampler = MODELS[model]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    np.random.seed(int(c["seed"]))
    g = sampler.noisy_gradient(kind="pf", pf=pf, N=int(c["N"]), subsequence_length=12, buffer_length=4,
                               minibatch_size=int(c["minibatch_size"]), **PARITY)
    assert sorted(g) == [str(k) for k in c["keys"]]
    np.testing.assert_allclose(_vec(g), c["values"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_sampler_sgld_steps(model):
    """three sample_sgld + project_parameters iterations land on the reference's parameters."""
    c = C.case("s/sgld_" + model)
    make, _, Sampler = MODELS[model]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    np.random.seed(int(c["seed"]))
    for _ in range(3):
        sampler.sample_sgld(epsilon=0.01, kind="pf", pf="poyiadjis_N", N=200, subsequence_length=12,
                            buffer_length=4, minibatch_size=1, **PARITY)
        sampler.project_parameters()
    np.testing.assert_allclose(_vec(sampler.parameters.var_dict), c["after"], rtol=1e-8, atol=1e-10)


def test_sampler_sgrld_steps_lgssm():
    from sgmcmc_ssm_b200.models.lgssm import LGSSMPreconditioner
    c = C.case("s/sgrld_lgssm")
    make, _, Sampler = MODELS["lgssm"]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    np.random.seed(int(c["seed"]))
    for _ in range(3):
        sampler.sample_sgrld(epsilon=0.01, preconditioner=LGSSMPreconditioner(), kind="pf", pf="poyiadjis_N",
                             N=200, subsequence_length=12, buffer_length=4, minibatch_size=1, **PARITY)
        sampler.project_parameters()
    np.testing.assert_allclose(_vec(sampler.parameters.var_dict), c["after"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_sampler_noisy_loglikelihood(model):
    c = C.case("s/noisy_loglik_" + model)
    make, _, Sampler = MODELS[model]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    np.random.seed(int(c["seed"]))
    ll = sampler.noisy_loglikelihood(kind="pf", pf="poyiadjis_N", N=250, subsequence_length=20, buffer_length=5,
                                     minibatch_size=2, **PARITY)
    np.testing.assert_allclose(ll, c["value"], rtol=1e-9)


@pytest.mark.parametrize("name,num_sequences", [("s/seq_svm", 2), ("s/seq_svm_all", -1)])
def test_seq_sampler_noisy_gradient(name, num_sequences):
    from sgmcmc_ssm_b200.models.svm import SeqSVMSampler
    c = C.case(name)
    obs = c["obs"]
    seqs = [obs[0:60], obs[60:110], obs[110:200]]
    sampler = SeqSVMSampler(n=1, m=1, observations=seqs, parameters=MODELS["svm"][0]())
    np.random.seed(int(c["seed"]))
    kw = dict(num_sequences=num_sequences) if num_sequences != -1 else {}
    g = sampler.noisy_gradient(kind="pf", pf="poyiadjis_N", N=200, subsequence_length=10, buffer_length=3,
                               minibatch_size=1, **kw, **PARITY)
    np.testing.assert_allclose(_vec(g), c["values"], rtol=1e-8, atol=1e-10)


# ---- the reference's own stored golden vectors, reproduced on the GPU ----------------------------------
def _unpack_state(vec):
    return ("MT19937", vec[:624].astype(np.uint32), int(vec[624]), int(vec[625]), float(vec[626]))


@pytest.mark.parametrize("cell", list(range(18)))
def test_stored_svm_golden_gradients_on_gpu(cell):
    """scratch/svm_grad_compare/(0.95, 0.5, 0.5)/trial/dat0_joblib.gz, rows (rep, buffer_size,
    poyiadjis_{100,1000,10000}): restore the numpy stream state, call the drop-in helper like
    svm_grad_compare.py:97-129 does, compare with the stored values (abs tol 1e-9)."""
    from sgmcmc_ssm_b200.models.svm import SVMHelper
    z = C.load("svm_replay.npz")
    B, t0, L = int(z["cell_B"][cell]), int(z["t0"]), int(z["L"])
    obs = z["observations"]
    fm = dict(log_constant=0.0, precision=z["prior_precision"], mean_precision=z["prior_mean_precision"])
    helper = SVMHelper(forward_message=fm, n=1, m=1)
    params = MODELS["svm"][0]()
    np.random.set_state(_unpack_state(z["cell_state"][cell]))
    for k, N in enumerate((100, 1000, 10000)):
        g = helper.pf_gradient_estimate(observations=obs[t0 - B:t0 + L + B], parameters=params, kernel=None,
                                        subsequence_start=B, subsequence_end=L + B, pf="poyiadjis_N", N=N,
                                        tqdm=None, **PARITY)
        got = [g["A"], g["LQinv_vec"], g["LRinv_vec"]]
        np.testing.assert_allclose(got, z["cell_stored"][cell][k], rtol=0, atol=1e-9)


def test_stored_svm_truth_run_N_1e6_on_gpu():
    """first of the ten N = 1 000 000 'truth' runs of svm_grad_compare.py:64-82 (a 46 s numpy run)."""
    from sgmcmc_ssm_b200.models.svm import SVMHelper
    z = C.load("svm_replay.npz")
    t0, L = int(z["t0"]), int(z["L"])
    obs = z["observations"]
    fm = dict(log_constant=0.0, precision=z["prior_precision"], mean_precision=z["prior_mean_precision"])
    helper = SVMHelper(forward_message=fm, n=1, m=1)
    np.random.set_state(_unpack_state(z["truth_states"][0]))
    g = helper.pf_gradient_estimate(observations=obs[t0 - L:t0 + 2 * L], parameters=MODELS["svm"][0](),
                                    subsequence_start=L, subsequence_end=2 * L, pf="poyiadjis_N", N=1000000,
                                    **PARITY)
    np.testing.assert_allclose([g["A"], g["LQinv_vec"], g["LRinv_vec"]], z["truth_grads"][0], rtol=0, atol=1e-8)
