"""ChainEnsemble: independent SG-MCMC chains (BASELINE configs[3]: many SGLD chains of a Seq sampler) advanced with
one batched launch per iteration.  Checks: an ensemble of one chain IS the plain sampler (bit for bit, same numpy
stream and same Philox key); inside a bigger ensemble chain 0 is still bit-identical (its items keep their global
ids) and every chain consumes its own numpy stream exactly as a separately seeded process would (same windows, same
SGLD noise), so the other chains differ from their solo runs only by the device randoms (Monte-Carlo level)."""
import numpy as np
import pytest

from tests.test_host_logic import svm_params, lgssm_params

pytestmark = pytest.mark.gpu


def _seq_data(seed=0, n_seq=7):
    rs = np.random.RandomState(seed)
    return [rs.normal(size=(int(rs.randint(30, 90)), 1)) for _ in range(n_seq)]


def _make(kind, c):
    from sgmcmc_ssm_b200.models.svm import SeqSVMSampler
    from sgmcmc_ssm_b200.models.lgssm import LGSSMSampler
    if kind == "seqsvm":
        return SeqSVMSampler(n=1, m=1, observations=_seq_data(), parameters=svm_params())
    rs = np.random.RandomState(100)
    return LGSSMSampler(n=1, m=1, observations=rs.normal(size=(300, 1)), parameters=lgssm_params())


# path='tiles': bit-for-bit reproducibility of an item across batch compositions holds within ONE kernel family -- with
# path='auto' the family (cluster / shared-memory / tile kernels, each with its own random streams) follows the batch size
KW = dict(seqsvm=dict(kind="pf", pf="poyiadjis_N", N=4000, subsequence_length=16, buffer_length=4, minibatch_size=2, num_sequences=2, path="tiles"),
          lgssm=dict(kind="pf", pf="poyiadjis_N", N=4000, subsequence_length=20, buffer_length=5, minibatch_size=3, path="tiles"))


def _vec(p):
    return np.concatenate([np.ravel(v) for v in p.as_dict().values()])


@pytest.mark.parametrize("kind", ["seqsvm", "lgssm"])
def test_ensemble_of_one_is_the_plain_sampler(kind):
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200.ensemble import ChainEnsemble
    solo = _make(kind, 0)
    sg.set_seed(5)
    np.random.seed(31)
    for _ in range(3):
        solo.sample_sgld(epsilon=1e-3, **KW[kind]); solo.project_parameters()
    ens = ChainEnsemble([_make(kind, 0)], seeds=[31])
    sg.set_seed(5)
    for _ in range(3):
        ens.sample_sgld(epsilon=1e-3, **KW[kind])
    np.testing.assert_array_equal(_vec(ens.samplers[0].parameters), _vec(solo.parameters))


@pytest.mark.parametrize("kind", ["seqsvm", "lgssm"])
def test_chains_keep_their_own_numpy_streams(kind):
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200.ensemble import ChainEnsemble
    C, seeds = 5, [31, 32, 33, 34, 35]
    np.random.seed(999)
    outer_before = np.random.get_state()[1].copy()
    ens = ChainEnsemble([_make(kind, c) for c in range(C)], seeds=seeds)
    sg.set_seed(5)
    ens.sample_sgld(epsilon=1e-3, **KW[kind])
    np.testing.assert_array_equal(np.random.get_state()[1], outer_before)        # the caller's stream is untouched
    for c in range(C):
        solo = ChainEnsemble([_make(kind, c)], seeds=[seeds[c]])
        sg.set_seed(5)
        solo.sample_sgld(epsilon=1e-3, **KW[kind])
        a, b = _vec(ens.samplers[c].parameters), _vec(solo.samplers[0].parameters)
        if c == 0:
            np.testing.assert_array_equal(a, b)
        else:       # same windows and same SGLD noise; only the particle filters' device randoms differ
            assert np.all(np.isfinite(a)) and np.max(np.abs(a - b)) < 0.05 * (1 + np.max(np.abs(b))), (a, b)
    with pytest.raises(ValueError):
        ChainEnsemble([_make("seqsvm", 0), _make("lgssm", 1)], seeds=[1, 2])
