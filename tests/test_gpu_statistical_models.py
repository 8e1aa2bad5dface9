"""Statistical parity of the PRODUCTION instantiations (device Philox randoms, order-statistics multinomial resampling,
run-time flags folded: `pf_step_kernel<R, Model, SORTED, FM_POY | FM_SHRINK | FM_FILTER, RAGGED>`, the single-launch
kernel for small batches, the Philox PaRIS kernels) against the UNMODIFIED reference, for every model / proposal kernel.

Golden: tests/golden/model_sweep_stats.npz (tests/golden/make_model_sweep_stats.py) -- for each (model, kernel, smoother)
group and each (N, buffer B) cell the mean and standard deviation over `reps` independent runs of the reference's
`buffered_pf_wrapper` + `average_statistic` (the protocol of gradient_error_fig_scripts/*_grad_compare.py: an L = 16
subsequence in the middle of a T = 100 series, B buffer steps on each side).  N = 1000 / 10000 are ragged (not a multiple
of the 256-particle warp tile), N = 4096 is not.

The GPU draws R repetitions of the same cell in ONE batched launch.  Same law iff, per gradient component,
    |mean_gpu - mean_ref| <= 4.5 sqrt(std_ref^2 / reps + std_gpu^2 / R)      (both the bias in N and B and the centre)
    LO <= std_gpu / std_ref <= HI                                            (a reps-sample std: +-9 % at 1 sigma for
                                                                              reps = 64, +-10 % for 48)
and the log-likelihood estimate obeys the same two rules.  A wrong score column, a wrong weight, a biased resampler or
a broken ragged tile in any fast instantiation moves a mean by many sigma (the components differ by O(1) while
std / sqrt(reps) is ~0.05).
"""
import numpy as np
import pytest

from tests import _cases as C

pytestmark = pytest.mark.gpu
LO, HI = 0.6, 1.6


def _groups():
    z = C.load("model_sweep_stats.npz")
    return sorted({k.split("/")[0] for k in z.files if "/" in k})


def _run_cell(group, N, B, R, dtype="f32", variates="native", seed=17, **kw):
    import sgmcmc_ssm_b200 as sg
    z = C.load("model_sweep_stats.npz")
    model, kernel, pf = group.split("_", 2)
    L, t0 = int(z["L"]), int(z["t0"])
    obs = z[group + "/obs"].reshape(-1)
    window = obs[t0 - B:t0 + L + B]
    pk = sg.PackedItems(np.tile(window, R), np.full(R, window.shape[0]), np.full(R, B), np.full(R, L + B), None, None,
                        z[group + "/theta"], float(z[group + "/prior_mean"]), float(z[group + "/prior_var"]))
    # N <= 2048 would run the shared-memory kernel (its own test below): the tile kernels are forced unless asked otherwise
    kw.setdefault("path", "tiles" if (N <= 2048 and not pf == "paris") else "auto")
    res = sg.run_pf(model, kernel, pf, pk, N, dtype=dtype, rng="philox", resample="multinomial_sorted", seed=seed,
                    offset=1000 * B + N % 977, variates=variates, **kw)
    return res


def _check(group, i, j, res, R, what):
    z = C.load("model_sweep_stats.npz")
    reps = int(z[group + "/reps"])
    m_ref, s_ref = z[group + "/mean"][i, j], z[group + "/std"][i, j]
    g = res.grad
    m, s = g.mean(axis=0), g.std(axis=0, ddof=1)
    tol = 4.5 * np.sqrt(s_ref ** 2 / reps + s ** 2 / R)
    assert np.all(np.abs(m - m_ref) <= tol), (what, m, m_ref, tol)
    assert np.all(s / s_ref > LO) and np.all(s / s_ref < HI), (what, s, s_ref)
    lm_ref, ls_ref = z[group + "/loglik_mean"][i, j], z[group + "/loglik_std"][i, j]
    lm, ls = res.loglik.mean(), res.loglik.std(ddof=1)
    assert abs(lm - lm_ref) <= 4.5 * np.sqrt(ls_ref ** 2 / reps + ls ** 2 / R), (what, lm, lm_ref)
    assert LO < ls / ls_ref < HI, (what, ls, ls_ref)


@pytest.mark.parametrize("group", _groups())
def test_production_path_matches_the_reference_distribution(group):
    """f32: every (N, B) cell of the group on the per-step tile kernels (fast modes; N = 1000 / 10000 ragged)."""
    z = C.load("model_sweep_stats.npz")
    paris = group.endswith("paris")
    R = 512 if paris else 2048
    for i, N in enumerate(z[group + "/Ns"]):
        for j, B in enumerate(z[group + "/buffer_sizes"]):
            res = _run_cell(group, int(N), int(B), R)
            _check(group, i, j, res, R, (group, int(N), int(B), "f32"))


@pytest.mark.parametrize("group", [g for g in _groups() if not g.endswith("paris")])
@pytest.mark.parametrize("dtype,R", [("f32", 1024), ("f32", 200), ("f64", 512)])
def test_shared_memory_kernel_matches_the_reference_distribution(group, dtype, R):
    """N = 1000: the one-CTA-per-item shared-memory kernel (small_kernels.cuh; 1024 threads x 1 particle for R <= 296
    items, 512 threads x 2 particles above)."""
    z = C.load("model_sweep_stats.npz")
    i = list(z[group + "/Ns"]).index(1000)
    for j, B in enumerate(z[group + "/buffer_sizes"]):
        res = _run_cell(group, 1000, int(B), R, dtype=dtype, seed=23, path="auto")
        assert res.launches == 1, res.launches
        _check(group, i, j, res, R, (group, 1000, int(B), "small", dtype, R))


@pytest.mark.parametrize("variates", ["native", "f32"])
@pytest.mark.parametrize("group", [g for g in _groups() if not g.endswith("paris")])
def test_f64_production_path_matches_the_reference_distribution(group, variates):
    """The reference's own precision with device randoms (fast-mode f64 instantiations; `variates='f32'` generates the
    random inputs with the f32 transforms and widens them): N = 4096 (full tiles) and N = 1000 (ragged)."""
    z = C.load("model_sweep_stats.npz")
    for N, R in ((4096, 1024), (1000, 2048)):
        i = list(z[group + "/Ns"]).index(N)
        for j, B in enumerate(z[group + "/buffer_sizes"]):
            res = _run_cell(group, N, int(B), R, dtype="f64", variates=variates, seed=29)
            _check(group, i, j, res, R, (group, N, int(B), "f64", variates))


@pytest.mark.parametrize("group", [g for g in _groups() if not g.endswith("paris")])
def test_cluster_kernel_matches_the_reference_distribution(group):
    """The thread-block-cluster kernel (cluster_kernels.cuh: particle system in the distributed shared memory of 2-8 CTAs)
    in every (N, B) cell: N = 1000 -> 4 CTAs x 256 particles, N = 4096 -> 8 x 512, N = 10000 -> 8 x 2048.  A launch holds
    at most 148 / C items, so R = 288 repetitions come from several launches with different call offsets."""
    import sgmcmc_ssm_b200 as sg
    z = C.load("model_sweep_stats.npz")
    model, kernel, pf = group.split("_", 2)
    L, t0 = int(z["L"]), int(z["t0"])
    obs = z[group + "/obs"].reshape(-1)
    for i, N in enumerate(z[group + "/Ns"]):
        per = 36 if N <= 1024 else 18
        for j, B in enumerate(z[group + "/buffer_sizes"]):
            window = obs[t0 - int(B):t0 + L + int(B)]
            pk = sg.PackedItems(np.tile(window, per), np.full(per, window.shape[0]), np.full(per, int(B)), np.full(per, L + int(B)),
                                None, None, z[group + "/theta"], float(z[group + "/prior_mean"]), float(z[group + "/prior_var"]))
            grads, lls = [], []
            for rep in range(288 // per):
                res = sg.run_pf(model, kernel, pf, pk, int(N), dtype="f32", rng="philox", seed=31, offset=100 * rep + int(B) + 1, path="cluster")
                assert res.launches == 1
                grads.append(res.grad); lls.append(res.loglik)

            class _R(object):
                grad, loglik = np.concatenate(grads), np.concatenate(lls)
            _check(group, i, j, _R, 288, (group, int(N), int(B), "cluster"))
