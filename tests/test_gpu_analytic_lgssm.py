"""Known-answer test of the production LGSSM instantiations: for the linear-Gaussian model the buffered subsequence
gradient has a closed form (Kalman forward / backward messages; reference `lgssm/helper.py:312-420`, reached through
`sampler._single_noisy_grad_loglikelihood(kind='marginal')`, `lgssm_grad_compare.py:59-78`).  The fixtures `a/lgssm_*`
of tests/golden/ref_cases.npz hold that analytic gradient for four buffered windows (interior, start of the series, end
of the series, the S = 40 / B = 10 shape), computed by the unmodified reference (make_ref_cases.py:analytic_level).

The particle estimate (Poyiadjis O(N)) converges to it as N grows, so the mean over a batch of independent f32 / Philox
/ order-statistics runs at N = 2^16 -- the fast-mode `pf_step_kernel<float, Lgssm{Prior,Optimal}, true, FM_POY>` -- must
hit the analytic value:   |mean - analytic| <= 5 * standard error + 5e-3 * (1 + |analytic|)
(the second term covers the O(1/N) bias of the estimator and f32 round-off; the components are O(1)-O(10))."""
import numpy as np
import pytest

from tests import _cases as C

pytestmark = pytest.mark.gpu
ORDER = [3, 2, 1, 0]        # device columns [LRinv_vec, LQinv_vec, C, A] -> fixture keys sorted: A, C, LQinv_vec, LRinv_vec


@pytest.mark.parametrize("dtype,N,R", [("f32", 65536, 256), ("f32", 10000, 1024), ("f64", 16384, 256)])
@pytest.mark.parametrize("kernel", ["optimal", "prior"])
@pytest.mark.parametrize("name", C.case_names("a"))
def test_large_N_particle_gradient_hits_the_analytic_kalman_gradient(name, kernel, dtype, N, R):
    import sgmcmc_ssm_b200 as sg
    c = C.case(name)
    assert list(c["keys"]) == ["A", "C", "LQinv_vec", "LRinv_vec"]
    obs = c["obs"].reshape(-1)
    S = int(c["tL"]) - int(c["t1"])
    pk = sg.PackedItems(np.tile(obs, R), np.full(R, obs.shape[0]), np.full(R, int(c["t1"])), np.full(R, int(c["tL"])),
                        np.tile(c["weights"], R), np.arange(R, dtype=np.int64) * S, c["theta"], 0.0, 10.0)
    res = sg.run_pf("lgssm", kernel, "poyiadjis_N", pk, N, dtype=dtype, rng="philox", resample="multinomial_sorted",
                    seed=41, offset=7)
    g = res.grad[:, ORDER]
    mean, se = g.mean(axis=0), g.std(axis=0, ddof=1) / np.sqrt(R)
    tol = 5.0 * se + 5e-3 * (1.0 + np.abs(c["values"]))
    assert np.all(np.abs(mean - c["values"]) <= tol), (name, kernel, dtype, mean, c["values"], tol)
