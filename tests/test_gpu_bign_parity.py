"""GPU parity of the O(N^2) Poyiadjis and PaRIS back ends at the sizes where they leave their single-tile regime
(N = 1024 / 2048: several 512-parent tiles, split-J grid, tensor-core kernel; PaRIS N = 1024 / 4096: multi-CTA work
queues, thousands of accept-reject entries per round), against outputs of the UNMODIFIED reference
(tests/golden/bign_cases.npz, make_bign_cases.py; reference pf.py:84-136, 183-341) and against the oracle's traces.

f64 + injected randoms (the oracle records the legacy stream the reference consumed under the fixture's seed):
  * ancestors and PaRIS backward indices J: bit-exact
  * particles / log-weights rtol 1e-10, per-particle statistics and gradient rtol 1e-8 (atol 1e-9), log-likelihood 1e-9
f32 with the same randoms, O(N^2): both back ends (`n2_mode='tensor'`: TF32 mma.sync, V split hi + lo, P rounded to 10
bits; `'fp32_pipe'`: CUDA cores) against the REFERENCE values:
  * gradient (weighted mean statistic): |diff| <= 1e-3 * (|ref| + mean |statistic|)  tensor, 5e-4 fp32 pipe
  * per-particle statistics: 99.9 % of the entries within 5e-3 * (|ref| + mean |statistic|)
    (an f32 CDF-rounding ancestor flip changes single particles; none of the sums over all parents moves by more)
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


@pytest.mark.parametrize("n2_mode,gtol", [("tensor", 1e-3), ("fp32_pipe", 5e-4)])
@pytest.mark.parametrize("name", [n for n in NAMES if "poyiadjis_N2" in n])
def test_f32_n2_back_ends_match_the_reference(name, n2_mode, gtol):
    c, model, kernel, pf, K, N, opts, kw, ref, parts = _replay(name)
    res = _run(c, model, kernel, pf, N, opts, kw, parts, "f32", n2_mode=n2_mode)
    stats = res.tensor("stats")[0].cpu().numpy()[:, :K.p].astype(np.float64)
    scale = np.abs(c["statistics"]) + np.mean(np.abs(c["statistics"]), axis=0)
    frac_ok = np.mean(np.abs(stats - c["statistics"]) <= 5e-3 * scale)
    assert frac_ok >= 0.999, (name, n2_mode, frac_ok)
    avg = po.average_statistic(dict(statistics=c["statistics"], log_weights=c["log_weights"]))
    gscale = np.abs(avg) + np.mean(np.abs(c["statistics"]), axis=0)
    assert np.all(np.abs(res.grad[0] - avg) <= gtol * gscale), (name, n2_mode, res.grad[0], avg)
