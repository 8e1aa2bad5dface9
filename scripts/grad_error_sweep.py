#!/usr/bin/env python3
"""Gradient-error sweep (gradient_error_fig_scripts/{svm,garch}_grad_compare.py protocol) on the GPU:
bias^2 / variance / MSE of the buffered PF gradient vs buffer size B and particle count N, against a "truth" = mean of
10 full-buffer runs at N = 10^6 (svm_grad_compare.py:64-82, garch_grad_compare.py:79-98, 226-235).  The reference
needs ~8 CPU-minutes for the truth runs alone and stops at N = 10^4 (SVM) / 10^6 (GARCH, single runs); here every
cell is one batched launch of `reps` items and N goes to 2^16.

  python scripts/grad_error_sweep.py [--model svm|garch] [--pf poyiadjis_N|paris|poyiadjis_N2] [--reps 256] [--out file.json]
SVM uses trial 0 of the reference's stored sweep (tests/golden/svm_replay.npz) and prints the reference's own trial-0
MSE (from its 50 stored repetitions, tests/golden/svm_sweep_stats.npz) next to the GPU's."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg


def cell(model, kern, pf, theta, window, t1, tL, N, reps, prior, seed, **kw):
    pk = sg.PackedItems(np.tile(window, reps), np.full(reps, window.shape[0]), np.full(reps, t1), np.full(reps, tL), None, None,
                        theta, prior[0], prior[1])
    return sg.run_pf(model, kern, pf, pk, N, dtype="f32", rng="philox", seed=seed, offset=1, **kw).grad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="svm", choices=["svm", "garch"])
    ap.add_argument("--pf", default="poyiadjis_N")
    ap.add_argument("--reps", type=int, default=256)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    L = 16
    if args.model == "svm":
        z = np.load(os.path.join(ROOT, "tests", "golden", "svm_replay.npz"))
        obs, t0, theta = z["observations"][:, 0], int(z["t0"]), z["theta"]
        prior = (0.0, float(np.linalg.inv(z["prior_precision"])[0, 0]))
        kern, Bs, keys = "prior", [20, 18, 15, 12, 10, 5, 3, 2, 0], ["LRinv_vec", "LQinv_vec", "A"]
        truth_B = L                                                      # svm_grad_compare.py:64-82
    else:
        rs = np.random.RandomState(12345)                                # garch_grad_compare.py:36-60 shape
        al, be, ga, tau = 0.1, 0.8, 0.05, 0.3
        mu, phi, lam = al / (1 - be - ga), be + ga, be / (be + ga)
        theta = [al, be, ga, mu, phi, lam, 1 / tau, 1 / tau ** 2, tau ** 2]
        T = 100
        x, s2, obs = 0.0, mu, np.zeros(T)
        for t in range(T):
            s2 = al + be * x * x + ga * s2
            x = np.sqrt(s2) * rs.normal()
            obs[t] = x + tau * rs.normal()
        t0, prior = (T + L) // 2, (0.0, mu)
        kern, Bs, keys = "optimal", [8, 6, 4, 3, 2, 1, 0], ["LRinv_vec", "log_mu", "logit_phi", "logit_lambduh"]
        truth_B = L
    t_start = time.time()
    truth = cell(args.model, kern, "poyiadjis_N", theta, obs[t0 - truth_B:t0 + L + truth_B], truth_B, L + truth_B,
                 1000000, 10, prior, seed=99).mean(axis=0)
    rows = []
    Ns = [1 << k for k in range(7, 17)]
    if args.pf == "poyiadjis_N2":
        Ns = [n for n in Ns if n <= 1 << 13]
    for B in Bs:
        w = obs[t0 - B:t0 + L + B]
        for N in Ns:
            g = cell(args.model, kern, args.pf, theta, w, B, L + B, N, args.reps, prior, seed=1000 + B)
            bias2 = (g.mean(axis=0) - truth) ** 2
            var = g.var(axis=0, ddof=1)
            rows.append(dict(buffer_size=B, N=N, **{k + "_bias_sq": float(bias2[i]) for i, k in enumerate(keys)},
                             **{k + "_var": float(var[i]) for i, k in enumerate(keys)},
                             **{k + "_mse": float(bias2[i] + var[i]) for i, k in enumerate(keys)}))
    out = dict(model=args.model, pf=args.pf, reps=args.reps, L=L, truth=dict(zip(keys, truth.tolist())),
               seconds=time.time() - t_start, rows=rows)
    if args.model == "svm":
        gold = np.load(os.path.join(ROOT, "tests", "golden", "svm_sweep_stats.npz"))
        ref = []
        tr = truth[[2, 1, 0]]                                           # golden order [A, LQinv_vec, LRinv_vec]
        for i, B in enumerate(gold["buffer_sizes"]):
            for j, N in enumerate(gold["Ns"]):
                mse = (gold["mean"][i, j] - tr) ** 2 + gold["std"][i, j] ** 2
                ref.append(dict(buffer_size=int(B), N=int(N), A_mse=float(mse[0]), LQinv_vec_mse=float(mse[1]),
                                LRinv_vec_mse=float(mse[2]), mean_runtime_s=float(gold["runtime"][i, j])))
        out["reference_trial0_stored"] = ref
    txt = json.dumps(out, indent=1)
    if args.out:
        open(args.out, "w").write(txt)
    print("sweep done in %.1f s; truth %s" % (out["seconds"], out["truth"]))
    for r in rows:
        if r["N"] in (128, 1024, 8192, 65536) and r["buffer_size"] in (Bs[0], Bs[len(Bs) // 2], 0):
            print("B=%2d N=%6d " % (r["buffer_size"], r["N"]) + " ".join("%s mse %.4g" % (k, r[k + "_mse"]) for k in keys))


if __name__ == "__main__":
    main()
