"""GPU tests of the O(N^2) smoother back ends (pf.py:84-136): the tensor-core kernel (TF32 mma, FP32
accumulate), the FP32-pipe kernel and the f64 instantiation must agree on identical genealogies.

With the same Philox seed the particle systems of the three runs coincide (the backward kernels only write
statistics), so the comparison is per particle.  Tolerances, tensor vs f32 pipe: 3e-3 * scale per particle
and 1e-3 * scale on the gradient (P is rounded to 10 mantissa bits, the same rounded weight feeds numerator
and denominator; measured worst case 1.1e-3 on a component that cancels to 1e-5), scale = |value| + mean |statistic|."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

THETA = {"svm": [0.95, np.sqrt(2.0), 2.0 + 1e-16, np.sqrt(2.0), 2.0 + 1e-16],
         "lgssm": [0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0],
         "garch": [0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09]}


def _items(model, B, T, seed):
    import sgmcmc_ssm_b200 as sg
    rs = np.random.RandomState(seed)
    it = sg.PFItems()
    for b in range(B):
        y = rs.normal(size=T) * (0.7 if model != "lgssm" else 1.5)
        it.add(y, THETA[model], t1=2, tL=T - 1, weights=1.0 + rs.rand(T - 3), prior_mean=0.0,
               prior_var=1.0 if model == "garch" else 5.0)
    return it


@pytest.mark.parametrize("model", ["svm", "lgssm", "garch"])
@pytest.mark.parametrize("B,N", [(1, 3000), (40, 1300)])          # split-J grid / single-split grid, ragged last tiles
def test_n2_backends_agree(model, B, N):
    import sgmcmc_ssm_b200 as sg
    kern = "prior" if model == "svm" else "optimal"
    it = _items(model, B, 9, 3)
    out = {}
    for tag, dtype, mode in (("f64", "f64", "auto"), ("f32", "f32", "fp32_pipe"), ("tc", "f32", "tensor")):
        r = sg.run_pf(model, kern, "poyiadjis_N2", it, N, dtype=dtype, rng="philox", seed=11, offset=5,
                      resample="multinomial_sorted", n2_mode=mode, want=("stats", "x"))
        out[tag] = (r.grad.copy(), r.tensor("stats").double().cpu().numpy(), r.tensor("x").double().cpu().numpy(), r.loglik.copy())
    # same genealogy in both f32 runs (bit-identical particles); f64 particles differ by round-off cascades, so
    # the f64 comparison is on the estimator only
    np.testing.assert_array_equal(out["f32"][2], out["tc"][2])
    scale = np.abs(out["f32"][1]) + np.mean(np.abs(out["f32"][1]), axis=1, keepdims=True)
    assert np.all(np.abs(out["tc"][1] - out["f32"][1]) <= 3e-3 * scale)
    gscale = np.abs(out["f32"][0]) + np.mean(np.abs(out["f32"][1]), axis=(1,))
    assert np.all(np.abs(out["tc"][0] - out["f32"][0]) <= 1e-3 * gscale)
    np.testing.assert_allclose(out["tc"][3], out["f32"][3], rtol=1e-6)
    # f64 vs f32: different genealogies after the first rounding flip -> Monte-Carlo level agreement of the
    # batch-mean gradient only
    g64, g32 = out["f64"][0].mean(axis=0), out["f32"][0].mean(axis=0)
    assert np.all(np.abs(g64 - g32) <= 0.35 * (np.abs(g64) + np.mean(np.abs(out["f64"][1]))))


def test_n2_tensor_requires_f32():
    import sgmcmc_ssm_b200 as sg
    it = _items("svm", 1, 5, 0)
    with pytest.raises(NotImplementedError):
        sg.run_pf("svm", "prior", "poyiadjis_N2", it, 256, dtype="f64", n2_mode="tensor")


@pytest.mark.parametrize("model", ["svm", "garch"])
def test_n2_matches_on_of_n_estimator_statistically(model):
    """O(N^2) and O(N) Poyiadjis estimate the same gradient; at N = 8192 / 16 items the batch means must agree
    within Monte-Carlo error (O(N) is the noisier one)."""
    import sgmcmc_ssm_b200 as sg
    kern = "prior" if model == "svm" else "optimal"
    it = _items(model, 16, 12, 5)
    g2 = sg.run_pf(model, kern, "poyiadjis_N2", it, 8192, dtype="f32", seed=3, offset=1).grad
    g1 = sg.run_pf(model, kern, "poyiadjis_N", it, 8192, dtype="f32", seed=4, offset=1).grad
    err = np.abs(g2 - g1).mean(axis=0)
    spread = np.abs(g1).mean(axis=0) + 1.0
    assert np.all(err <= 0.5 * spread), (err, spread)


@pytest.mark.parametrize("model,Ntilde", [("svm", 2), ("garch", 2), ("garch", 5), ("lgssm", 3)])
def test_paris_philox_matches_n2_statistically(model, Ntilde):
    """PaRIS with device randoms (guide-table proposals, shared-memory work queue, exact fallback; pf.py:183-341
    in law) estimates what the O(N^2) smoother computes exactly: batch means over 24 items at N = 4096 must agree
    within Monte-Carlo error, every backward index must be valid, and the fallback must be exercised somewhere."""
    import sgmcmc_ssm_b200 as sg
    kern = "prior" if model == "svm" else "optimal"
    it = _items(model, 24, 14, 9)
    rp = sg.run_pf(model, kern, "paris", it, 4096, dtype="f32", seed=21, offset=1, Ntilde=Ntilde, want=("J",))
    J = rp.tensor("J").cpu().numpy()
    assert J.min() >= 0 and J.max() < 4096
    r2 = sg.run_pf(model, kern, "poyiadjis_N2", it, 4096, dtype="f32", seed=22, offset=1)
    gp, g2 = rp.grad.mean(axis=0), r2.grad.mean(axis=0)
    spread = np.abs(r2.grad - g2).mean(axis=0) + 0.05 * np.abs(g2) + 1e-3
    assert np.all(np.abs(gp - g2) <= 1.0 * spread), (gp, g2, spread)
    np.testing.assert_allclose(rp.loglik.mean(), r2.loglik.mean(), rtol=0.02, atol=0.2)
