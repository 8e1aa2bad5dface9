#!/usr/bin/env python3
"""Generate tests/golden/svm_replay.npz: RNG-state snapshots + stored reference gradients.

Runs ONLY in the build container (needs /root/reference); the output .npz is committed and is what
the tests read.  Nothing under tests/, bench.py or smoke() imports /root/reference at run time.

What it pins
------------
The reference ships bit-reproducible golden vectors for the SVM `poyiadjis_N` gradient:
`scratch/svm_grad_compare/(0.95, 0.5, 0.5)/trial/dat0_joblib.gz`, written by
`nonlinear_ssm_pf_experiment_scripts/gradient_error_fig_scripts/svm_grad_compare.py:28-150`.
That script is not importable (needs matplotlib/seaborn, and has a syntax hole at :12), so its
logic is re-hosted here in the exact RNG order:

  seed 12345 -> SVMParameters(0.95, 0.5, 0.5) -> generate_svm_data(T=100) -> 10 "truth" runs at
  N=1e6 with buffer L=16 -> for B in [20,18,15,12,10,5,3,2,0]: for rep in range(50): N=100, 1000,
  10000.

For every (B, rep) with rep < KEEP_REPS we snapshot the legacy MT19937 state *before* the three
calls, run the live reference, check it reproduces the stored dat0 rows, and save
(state, observations window, stored gradients).  The oracle test restores the state, draws the
uniforms/normals in the reference's order and must land on the stored values.
The first N=1e6 truth run is saved the same way (state + live reference gradient) for the GPU
large-N parity test.
"""
import os
import sys
import time

import numpy as np

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
import joblib  # noqa: E402
from sgmcmc_ssm.models.svm import SVMParameters, SVMHelper, generate_svm_data  # noqa: E402

KEEP_REPS = 2
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "svm_replay.npz")
DAT0 = "/root/reference/scratch/svm_grad_compare/(0.95, 0.5, 0.5)/trial/dat0_joblib.gz"


def pack_state(st):
    # ('MT19937', key[624] uint32, pos, has_gauss, cached_gaussian)
    return np.concatenate([st[1].astype(np.float64), [float(st[2]), float(st[3]), float(st[4])]])


def convert(g):
    return [float(np.ravel(g["A"])[0]), float(np.ravel(g["LQinv_vec"])[0]), float(np.ravel(g["LRinv_vec"])[0])]


def main():
    T, L, N_reps = 100, 16, 50
    buffer_sizes = [20, 18, 15, 12, 10, 5, 3, 2, 0]
    df = joblib.load(DAT0)
    np.random.seed(12345)
    A = np.eye(1) * 0.95
    Q = np.eye(1) * 0.5
    R = np.eye(1) * 0.5
    LQinv = np.linalg.cholesky(np.linalg.inv(Q))
    LRinv = np.linalg.cholesky(np.linalg.inv(R))
    parameters = SVMParameters(A=A, LQinv=LQinv, LRinv=LRinv)
    data = generate_svm_data(T=T, parameters=parameters)
    t0 = (T + L) // 2
    obs = data["observations"]
    helper = SVMHelper(forward_message=data["initial_message"], **parameters.dim)

    out = dict(
        observations=obs,
        t0=t0, L=L,
        theta=np.array([parameters.A[0, 0], parameters.LQinv[0, 0], parameters.Qinv[0, 0],
                        parameters.LRinv[0, 0], parameters.Rinv[0, 0]]),
        prior_precision=np.asarray(data["initial_message"]["precision"], dtype=float),
        prior_mean_precision=np.asarray(data["initial_message"]["mean_precision"], dtype=float),
    )

    truth_states, truth_grads = [], []
    for rep in range(10):
        st = pack_state(np.random.get_state())
        tic = time.time()
        g = convert(helper.pf_gradient_estimate(
            observations=obs[t0 - L:t0 + 2 * L], parameters=parameters, kernel=None,
            subsequence_start=L, subsequence_end=2 * L, pf="poyiadjis_N", N=1000000))
        print("truth rep", rep, g, "%.1fs" % (time.time() - tic), flush=True)
        if rep < 1:
            truth_states.append(st)
            truth_grads.append(g)
    out["truth_states"] = np.array(truth_states)
    out["truth_grads"] = np.array(truth_grads)  # columns A, LQinv_vec, LRinv_vec

    cell_B, cell_rep, cell_state, cell_stored, cell_live = [], [], [], [], []
    worst = 0.0
    for B in buffer_sizes:
        for rep in range(N_reps):
            st = pack_state(np.random.get_state())
            live, stored = [], []
            for N in (100, 1000, 10000):
                g = convert(helper.pf_gradient_estimate(
                    observations=obs[t0 - B:t0 + L + B], parameters=parameters, kernel=None,
                    subsequence_start=B, subsequence_end=L + B, pf="poyiadjis_N", N=N))
                live.append(g)
                q = df[(df.rep == rep) & (df.buffer_size == B) & (df.sampler == "poyiadjis_%d" % N)]
                stored.append([float(q[q.variable == v].value.iloc[0]) for v in ("A", "LQinv_vec", "LRinv_vec")])
            d = float(np.max(np.abs(np.array(live) - np.array(stored))))
            worst = max(worst, d)
            if rep < KEEP_REPS:
                cell_B.append(B)
                cell_rep.append(rep)
                cell_state.append(st)
                cell_stored.append(stored)
                cell_live.append(live)
        print("B", B, "worst |live-stored| so far", worst, flush=True)
    out.update(cell_B=np.array(cell_B), cell_rep=np.array(cell_rep), cell_state=np.array(cell_state),
               cell_stored=np.array(cell_stored), cell_live=np.array(cell_live),
               replay_worst_abs_diff=np.array(worst))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "worst abs diff live-vs-dat0 over all 1350 cells:", worst)


if __name__ == "__main__":
    main()
