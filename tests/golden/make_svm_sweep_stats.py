#!/usr/bin/env python3
"""Generate tests/golden/svm_sweep_stats.npz from the reference's OWN stored gradient-error sweep
(`scratch/svm_grad_compare/(0.95, 0.5, 0.5)/trial/dat0_joblib.gz`, written by
gradient_error_fig_scripts/svm_grad_compare.py:97-150): for trial 0 and every (buffer size B, N) cell the mean and
standard deviation over its 50 repetitions of the three gradient components [A, LQinv_vec, LRinv_vec], plus the
mean runtime.  The observations / parameters of trial 0 are already in svm_replay.npz.  Build-container only."""
import os
import numpy as np
import joblib

DAT0 = "/root/reference/scratch/svm_grad_compare/(0.95, 0.5, 0.5)/trial/dat0_joblib.gz"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "svm_sweep_stats.npz")
df = joblib.load(DAT0)
Bs = [20, 18, 15, 12, 10, 5, 3, 2, 0]
Ns = [100, 1000, 10000]
vars_ = ["A", "LQinv_vec", "LRinv_vec"]
mean = np.zeros((len(Bs), len(Ns), 3)); std = np.zeros_like(mean); cnt = np.zeros((len(Bs), len(Ns)), dtype=np.int64)
runtime = np.zeros((len(Bs), len(Ns)))
for i, B in enumerate(Bs):
    for j, N in enumerate(Ns):
        sub = df[(df["buffer_size"] == B) & (df["sampler"] == "poyiadjis_{0}".format(N))]
        for k, v in enumerate(vars_):
            vals = sub[sub["variable"] == v]["value"].to_numpy(dtype=float)
            mean[i, j, k], std[i, j, k] = vals.mean(), vals.std(ddof=1)
            cnt[i, j] = vals.shape[0]
        runtime[i, j] = sub[sub["variable"] == "runtime"]["value"].to_numpy(dtype=float).mean()
np.savez_compressed(OUT, buffer_sizes=np.array(Bs), Ns=np.array(Ns), mean=mean, std=std, count=cnt, runtime=runtime)
print("wrote", OUT, cnt.min(), cnt.max()); print(mean[0], std[0])
