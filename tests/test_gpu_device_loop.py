"""Device-resident SG-MCMC loop (C-ABI sgm_sgld_run, sgmcmc_ssm_b200/device_loop.py): whole iterations -- window draw,
particle filters, grad log-prior, Langevin noise, update, projection -- on the device.

Parity mode (rng='injected', f64): the host consumes the global numpy stream in the reference's order and ships the
draws; three iterations land on the parameters the UNMODIFIED reference reaches (fixtures s/sgld_*, s/sgrld_lgssm of
tests/golden/ref_cases.npz: `sample_sgld` / `sample_sgrld` + `project_parameters`, sgmcmc_sampler.py:549-567, 613-640,
650-656): rtol 1e-8.  A Seq sampler (sequence choice, T / S rescale, :1249-1283) is compared with this repo's host loop,
whose gradient path is itself reference-pinned (s/seq_svm*).
Device randoms: graph replay == plain enqueue bit for bit; the law of the chain matches the host loop's."""
import numpy as np
import pytest

from tests import _cases as C
from tests.test_host_logic import MODELS, svm_params

pytestmark = pytest.mark.gpu


def _vec(var_dict):
    return np.concatenate([np.ravel(var_dict[k]) for k in sorted(var_dict)])


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
def test_injected_sgld_iterations_land_on_the_reference_parameters(model):
    from sgmcmc_ssm_b200.device_loop import DeviceChains
    c = C.case("s/sgld_" + model)
    make, _, Sampler = MODELS[model]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    np.random.seed(int(c["seed"]))
    chains = DeviceChains([sampler], method="SGLD", epsilon=0.01, pf="poyiadjis_N", N=200, subsequence_length=12,
                          buffer_length=4, minibatch_size=1, rng="injected", dtype="f64", trace_every=1, max_trace_rows=3)
    chains.run(3)
    chains.pull_parameters()
    np.testing.assert_allclose(_vec(sampler.parameters.var_dict), c["after"], rtol=1e-8, atol=1e-10)
    tr = chains.trace()
    assert tr.shape == (4, 1, len(chains.slots))
    np.testing.assert_allclose(sorted(tr[0, 0]), sorted(c["before"]), rtol=0, atol=1e-15)
    np.testing.assert_allclose(sorted(tr[3, 0]), sorted(c["after"]), rtol=1e-8, atol=1e-10)


def test_injected_sgrld_iterations_lgssm():
    from sgmcmc_ssm_b200.device_loop import DeviceChains
    from sgmcmc_ssm_b200.models.lgssm import LGSSMPreconditioner
    c = C.case("s/sgrld_lgssm")
    make, _, Sampler = MODELS["lgssm"]
    sampler = Sampler(n=1, m=1, observations=c["obs"], parameters=make())
    np.random.seed(int(c["seed"]))
    DeviceChains([sampler], method="SGRLD", preconditioner=LGSSMPreconditioner(), epsilon=0.01, pf="poyiadjis_N", N=200,
                 subsequence_length=12, buffer_length=4, minibatch_size=1, rng="injected", dtype="f64").run(3).pull_parameters()
    np.testing.assert_allclose(_vec(sampler.parameters.var_dict), c["after"], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("num_sequences,M", [(2, 1), (None, 2), (1, 3), (1, 1)])       # (1, 1): persistent kernel
def test_injected_seq_sampler_matches_the_host_loop(num_sequences, M):
    from sgmcmc_ssm_b200.device_loop import DeviceChains
    from sgmcmc_ssm_b200.models.svm import SeqSVMSampler
    rs = np.random.RandomState(3)
    seqs = [rs.normal(size=(n, 1)) for n in (60, 9, 50, 90, 14)]          # two sequences shorter than S run whole
    kw = dict(kind="pf", pf="poyiadjis_N", N=300, subsequence_length=16, buffer_length=4, minibatch_size=M)
    if num_sequences is not None:
        kw["num_sequences"] = num_sequences
    host = SeqSVMSampler(n=1, m=1, observations=seqs, parameters=svm_params())
    np.random.seed(77)
    for _ in range(3):
        host.sample_sgld(epsilon=1e-3, rng="injected", dtype="f64", **kw)
        host.project_parameters()
    dev = SeqSVMSampler(n=1, m=1, observations=seqs, parameters=svm_params())
    np.random.seed(77)
    DeviceChains([dev], method="SGLD", epsilon=1e-3, rng="injected", dtype="f64", **kw).run(3).pull_parameters()
    np.testing.assert_allclose(_vec(dev.parameters.var_dict), _vec(host.parameters.var_dict), rtol=1e-9, atol=1e-12)


def _lgssm_sampler(T=400):
    make, _, Sampler = MODELS["lgssm"]
    rs = np.random.RandomState(5)
    x, y = 0.0, np.zeros((T, 1))
    for t in range(T):
        x = 0.9 * x + np.sqrt(0.1) * rs.normal()
        y[t, 0] = x + rs.normal()
    return Sampler(n=1, m=1, observations=y, parameters=make())


def test_graph_replay_equals_plain_enqueue_and_draws_fresh_randoms():
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200.device_loop import DeviceChains
    kw = dict(method="SGLD", epsilon=0.01, pf="poyiadjis_N", N=1000, subsequence_length=20, buffer_length=5, minibatch_size=2,
              trace_every=1, max_trace_rows=40)
    out = []
    for graph in (False, True):
        sg.set_seed(123)
        ch = DeviceChains([_lgssm_sampler() for _ in range(3)], **kw)
        ch.run(33, graph=graph, chunk=8)
        ch.pull_parameters()
        out.append(ch.trace())
    np.testing.assert_array_equal(out[0], out[1])
    steps = np.diff(out[1][:, 0, 0])
    assert len(np.unique(steps)) == len(steps)          # every replayed iteration moved A by a different amount


@pytest.mark.parametrize("model,N,path", [("lgssm", 1000, "auto"), ("svm", 200, "auto"), ("garch", 2048, "auto"), ("lgssm", 1000, "cluster"),
                                          ("svm", 8192, "cluster")])   # auto: one CTA per chain up to N = 2048; path='cluster': a cluster of CTAs per chain
def test_persistent_kernel_equals_one_launch_sequence_per_iteration(model, N, path):
    """One work item per chain and N <= 2048: all iterations run inside ONE persistent kernel.  Iteration k uses the
    Philox call offset (base + k) either way, so the chains must be bit-identical to the launch-per-iteration path."""
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200.device_loop import DeviceChains
    make, _, Sampler = MODELS[model]
    obs = C.case("s/sgld_" + model)["obs"]
    kw = dict(method="SGLD", epsilon=0.005, pf="poyiadjis_N", N=N, subsequence_length=16, buffer_length=4, minibatch_size=1,
              trace_every=1, max_trace_rows=12, path=path)
    out = []
    for persistent in (True, False):
        sg.set_seed(77)
        ch = DeviceChains([Sampler(n=1, m=1, observations=obs, parameters=make()) for _ in range(5)], persistent=persistent, **kw)
        assert ch.persistent == persistent and ch.cluster == (persistent and path == "cluster")
        ch.run(7); ch.run(5)
        ch.pull_parameters()
        assert ch.launches == (2 if persistent else (3 if N <= 16384 else 124) * 5)
        out.append(ch.trace())
    assert out[0].shape == (13, 5, len(ch.slots))
    np.testing.assert_array_equal(out[0], out[1])
    assert len(np.unique(out[0][:, 0, 0])) == 13 and not np.array_equal(out[0][:, 0], out[0][:, 1])


def test_device_loop_has_the_law_of_the_host_loop():
    """256 LGSSM chains from the same start, 12 SGLD iterations each: device loop (Philox windows / noise) vs the host
    ensemble (numpy windows / noise, the reference's update code).  Means within 4.5 standard errors, spreads within 25 %."""
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200.ensemble import ChainEnsemble
    Cn, K = 256, 12
    kw = dict(kind="pf", pf="poyiadjis_N", N=500, subsequence_length=20, buffer_length=5, minibatch_size=1)
    sg.set_seed(9)
    host = ChainEnsemble([_lgssm_sampler() for _ in range(Cn)], seeds=list(range(1000, 1000 + Cn)))
    for _ in range(K):
        host.sample_sgld(epsilon=0.05, **kw)
    a = np.array([_vec(s.parameters.var_dict) for s in host.samplers])
    dev = ChainEnsemble([_lgssm_sampler() for _ in range(Cn)], seeds=list(range(Cn)))
    dev.fit("SGLD", K, epsilon=0.05, **kw)
    b = np.array([_vec(s.parameters.var_dict) for s in dev.samplers])
    keep = a.std(axis=0) > 0                           # C is projected to 1 in both
    se = np.sqrt(a.var(axis=0, ddof=1) / Cn + b.var(axis=0, ddof=1) / Cn)
    assert np.all(np.abs(a.mean(axis=0) - b.mean(axis=0))[keep] <= 4.5 * se[keep]), (a.mean(axis=0), b.mean(axis=0), se)
    assert np.all(np.abs(b.std(axis=0)[keep] / a.std(axis=0)[keep] - 1) < 0.25), (a.std(axis=0), b.std(axis=0))
    np.testing.assert_array_equal(a[:, ~keep], b[:, ~keep])


def test_fit_runs_on_the_device_and_returns_the_reference_shapes():
    import sgmcmc_ssm_b200 as sg
    s = _lgssm_sampler()
    sg.set_seed(4)
    before = _vec(s.parameters.var_dict)
    plist = s.fit("SGLD", 10, output_all=True, epsilon=0.01, subsequence_length=20, buffer_length=5, kind="pf",
                  pf_kwargs=dict(pf="poyiadjis_N", N=1000))
    assert len(plist) == 11 and type(plist[0]) is type(s.parameters)
    np.testing.assert_array_equal(_vec(plist[0].var_dict), before)
    np.testing.assert_array_equal(_vec(plist[-1].var_dict), _vec(s.parameters.var_dict))
    assert plist[-1].A.shape == (1, 1) and plist[-1].LQinv_vec.shape == (1,)
    assert not np.array_equal(_vec(plist[5].var_dict), before) and float(plist[-1].C[0, 0]) == 1.0
    # host loop still available
    s.fit("SGLD", 2, epsilon=0.01, subsequence_length=20, buffer_length=5, kind="pf", device_loop=False,
          pf_kwargs=dict(pf="poyiadjis_N", N=1000))
