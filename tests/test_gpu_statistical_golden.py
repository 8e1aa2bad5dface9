"""Statistical parity of the PRODUCTION path (f32, device Philox randoms, order-statistics multinomial resampling)
against the reference's own stored gradient-error sweep (scratch/svm_grad_compare/.../dat0_joblib.gz, trial 0:
50 repetitions per cell; summarised by tests/golden/make_svm_sweep_stats.py).

For every (buffer size B, N) cell the reference's 50 gradients are a sample of the estimator's distribution; the GPU
draws R = 1024 repetitions of the same cell in one batched launch.  The estimator's law is the same iff
  |mean_gpu - mean_ref| <= 4.5 * sqrt(std_ref^2 / 50 + std_gpu^2 / R)       (4.5 sigma, 81 cells x 3 components)
  0.62 <= std_gpu / std_ref <= 1.6                                           (50-sample std: +-10 % at 1 sigma)
Both the bias (which depends on N and on the buffer B) and the variance are pinned this way."""
import numpy as np
import pytest

from tests import _cases as C

pytestmark = pytest.mark.gpu
R = 1024


def _cell(B, N, dtype="f32", resample="multinomial_sorted", seed=5):
    import sgmcmc_ssm_b200 as sg
    z = C.load("svm_replay.npz")
    obs, t0, L = z["observations"], int(z["t0"]), int(z["L"])
    prior_var = float(np.linalg.inv(z["prior_precision"])[0, 0])
    prior_mean = float(np.linalg.solve(np.linalg.inv(z["prior_precision"]), z["prior_mean_precision"])[0])
    window = obs[t0 - B:t0 + L + B, 0]
    pk = sg.PackedItems(np.tile(window, R), np.full(R, window.shape[0]), np.full(R, B), np.full(R, L + B), None, None,
                        z["theta"], prior_mean, prior_var)
    res = sg.run_pf("svm", "prior", "poyiadjis_N", pk, N, dtype=dtype, rng="philox", resample=resample, seed=seed, offset=B + 1)
    g = res.grad                                   # columns [LRinv_vec, LQinv_vec, A]
    return g[:, [2, 1, 0]]                         # golden order [A, LQinv_vec, LRinv_vec]


@pytest.mark.parametrize("N_idx", [0, 1, 2])
def test_gradient_distribution_matches_the_references_stored_sweep(N_idx):
    gold = C.load("svm_sweep_stats.npz")
    N = int(gold["Ns"][N_idx])
    for i, B in enumerate(gold["buffer_sizes"]):
        g = _cell(int(B), N)
        m_ref, s_ref = gold["mean"][i, N_idx], gold["std"][i, N_idx]
        m, s = g.mean(axis=0), g.std(axis=0, ddof=1)
        tol = 4.5 * np.sqrt(s_ref ** 2 / 50.0 + s ** 2 / R)
        assert np.all(np.abs(m - m_ref) <= tol), (int(B), N, m, m_ref, tol)
        assert np.all(s / s_ref > 0.62) and np.all(s / s_ref < 1.6), (int(B), N, s, s_ref)


def test_reference_semantics_resampling_has_the_same_law_as_the_sorted_sampler():
    """iid-uniform multinomial (the reference's np.random.choice semantics) vs the order-statistics sampler: same
    estimator distribution (means within 4.5 sigma, stds within 12 %) at N = 1000, B = 10."""
    a = _cell(10, 1000, resample="multinomial", seed=11)
    b = _cell(10, 1000, resample="multinomial_sorted", seed=12)
    se = np.sqrt(a.var(axis=0, ddof=1) / R + b.var(axis=0, ddof=1) / R)
    assert np.all(np.abs(a.mean(axis=0) - b.mean(axis=0)) <= 4.5 * se)
    assert np.all(np.abs(a.std(axis=0) / b.std(axis=0) - 1) < 0.12)
