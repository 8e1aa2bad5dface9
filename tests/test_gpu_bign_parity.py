"""GPU parity of the O(N^2) Poyiadjis and PaRIS back ends at the sizes where they leave their single-tile regime
(N = 1024 / 2048: several 512-parent tiles, split-J grid, tensor-core kernel; PaRIS N = 1024 / 4096: multi-CTA work
queues, thousands of accept-reject entries per round), against outputs of the UNMODIFIED reference
(tests/golden/bign_cases.npz, make_bign_cases.py; reference pf.py:84-136, 183-341) and against the oracle's traces.

f64 + injected randoms (the oracle records the legacy stream the reference consumed under the fixture's seed):
  * ancestors and PaRIS backward indices J: bit-exact
  * particles / log-weights rtol 1e-10, per-particle statistics and gradient rtol 1e-8 (atol 1e-9), log-likelihood 1e-9
f32, O(N^2): both back ends (`n2_mode='tensor'`: TF32 mma.sync, V split hi + lo, P rounded to 10 bits; `'fp32_pipe'`:
CUDA cores) against the ORACLE's recursion evaluated on the run's own traced particle system (see the test's docstring).
"""
import numpy as np
import pytest

from oracle import pf_oracle as po
from tests import _cases as C

pytestmark = pytest.mark.gpu
FILE = "bign_cases.npz"
NAMES = C.case_names("k", FILE)


def _replay(name):
    c = C.case(name, FILE)
    model, kernel, pf = C.parse_kernel_case(name)
    K = po.make_kernel(model, kernel, C.theta_dict(model, c["theta"]))
    N, opts = int(c["N"]), C.case_opts(c)
    rec = po.LegacyStream(int(c["seed"]), record=True)
    kw = dict(t1=int(c["t1"]), tL=int(c["tL"]), weights=c["weights"], prior_mean=float(c["prior_mean"]),
              prior_var=float(c["prior_var"]))
    ref = po.buffered_pf(pf, c["obs"], K, N, K.score, K.p, rec, save_all=True, **kw, **opts)
    parts = po.split_events(rec.events, N)
    return c, model, kernel, pf, K, N, opts, kw, ref, parts


def _run(c, model, kernel, pf, N, opts, kw, parts, dtype, **extra):
    import sgmcmc_ssm_b200 as sg
    items = sg.PFItems().add(c["obs"], c["theta"], **kw)
    want = ("x", "lw", "stats", "anc") + (("J",) if pf == "paris" else ())
    return sg.run_pf(model, kernel, pf, items, N, dtype=dtype, rng="injected", resample="multinomial",
                     injected=dict(z0=parts["z0"], u=parts["u"], z=parts["z"], extra=parts["extra"]), want=want,
                     **opts, **extra)


@pytest.mark.parametrize("name", NAMES)
def test_f64_injected_matches_reference_and_oracle(name):
    c, model, kernel, pf, K, N, opts, kw, ref, parts = _replay(name)
    # the oracle itself reproduces the stored reference output at this size
    np.testing.assert_allclose(ref["statistics"], c["statistics"], rtol=1e-9, atol=1e-10)
    res = _run(c, model, kernel, pf, N, opts, kw, parts, "f64")
    np.testing.assert_array_equal(res.tensor("anc")[0].cpu().numpy(), np.array(ref["trace"]["ancestors"]))
    if pf == "paris":
        np.testing.assert_array_equal(res.tensor("J")[0].cpu().numpy(), np.array(ref["trace"]["J"]))
    x = res.tensor("x")[0].cpu().numpy()
    np.testing.assert_allclose(x, c["x_t"].reshape(x.shape), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res.tensor("lw")[0].cpu().numpy(), c["log_weights"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(res.loglik[0], c["loglik"], rtol=1e-9, atol=1e-10)
    stats = res.tensor("stats")[0].cpu().numpy()[:, :K.p]
    np.testing.assert_allclose(stats, c["statistics"], rtol=1e-8, atol=1e-9)
    avg = po.average_statistic(dict(statistics=c["statistics"], log_weights=c["log_weights"]))
    np.testing.assert_allclose(res.grad[0], avg, rtol=1e-8, atol=1e-9)


def _oracle_backward_recursion(K, tx, tlw, obs, t1, tL, weights):
    """The reference's O(N^2) recursion (pf.py:115-135; oracle poyiadjis_smoother) evaluated in float64 on a GIVEN
    particle system: tx (T + 1, N, n) particles and tlw (T + 1, N) log-weights of every step."""
    T, N = tx.shape[0] - 1, tx.shape[1]
    stats = np.zeros((N, K.p))
    for t in range(T):
        scale = float(weights[t - t1]) if t1 <= t < tL else 0.0
        y = float(obs[t])
        old, new, lw = tx[t], tx[t + 1], tlw[t]
        nxt = np.zeros_like(stats)
        for i in range(N):
            rep = np.outer(np.ones(N), new[i])
            bw = po.log_normalize(lw + K.prior_log_density(old, rep))
            add = K.score(old, rep, y) * scale if scale != 0.0 else 0.0
            nxt[i] = np.einsum("jk,j->k", stats + add, bw)
        stats = nxt
    return stats


@pytest.mark.parametrize("n2_mode,ptol,gtol", [("tensor", 5e-3, 1e-3), ("fp32_pipe", 5e-4, 2e-4)])
@pytest.mark.parametrize("name", [n for n in NAMES if "poyiadjis_N2" in n])
def test_f32_n2_back_ends_match_the_oracle_recursion(name, n2_mode, ptol, gtol):
    """f32 runs differ from the f64 reference run by ancestor flips (a uniform within f32 round-off of a CDF boundary:
    ~0.4 % of the draws), so the backward kernels are checked on the particle system the f32 run itself produced: the
    oracle's recursion (reference formulas, float64, full N x N weights) is evaluated on the GPU's traced particles and
    log-weights and must reproduce the GPU's per-particle statistics:
        |stat_gpu - stat_oracle| <= ptol * (|stat_oracle| + mean |stat|)   for every particle and component
        gradient (weighted mean)  <= gtol * (|grad| + mean |stat|)
    tensor cores (TF32, P rounded to 10 bits, V split hi + lo): 5e-3 / 1e-3;  FP32 pipe: 5e-4 / 2e-4."""
    import sgmcmc_ssm_b200 as sg
    c, model, kernel, pf, K, N, opts, kw, ref, parts = _replay(name)
    items = sg.PFItems().add(c["obs"], c["theta"], **kw)
    res = sg.run_pf(model, kernel, pf, items, N, dtype="f32", rng="injected", resample="multinomial",
                    injected=dict(z0=parts["z0"], u=parts["u"], z=parts["z"]), want=("stats", "lw", "trace_x", "trace_lw"),
                    n2_mode=n2_mode)
    tx = res.tensor("trace_x")[0].double().cpu().numpy()
    tlw = res.tensor("trace_lw")[0].double().cpu().numpy()
    expect = _oracle_backward_recursion(K, tx, tlw, c["obs"].reshape(-1), kw["t1"], kw["tL"], c["weights"])
    stats = res.tensor("stats")[0].double().cpu().numpy()[:, :K.p]
    scale = np.abs(expect) + np.mean(np.abs(expect), axis=0)
    assert np.all(np.abs(stats - expect) <= ptol * scale), (name, n2_mode, np.max(np.abs(stats - expect) / scale))
    avg = po.average_statistic(dict(statistics=expect, log_weights=tlw[-1]))
    gscale = np.abs(avg) + np.mean(np.abs(expect), axis=0)
    assert np.all(np.abs(res.grad[0] - avg) <= gtol * gscale), (name, n2_mode, res.grad[0], avg)
    # ... and the f32 run is the same Monte-Carlo estimate as the reference's f64 run up to the flips
    ref_avg = po.average_statistic(dict(statistics=c["statistics"], log_weights=c["log_weights"]))
    assert np.all(np.abs(res.grad[0] - ref_avg) <= 0.25 * (np.abs(ref_avg) + np.mean(np.abs(c["statistics"]), axis=0)))
