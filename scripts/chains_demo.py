#!/usr/bin/env python3
"""BASELINE configs[3] shape: many independent SGLD chains of a Seq SVM sampler on exchange-rate-like data
(49 sequences, 5907 observations; demo/exchange_rate/exchange_rate_full_demo.py:16-42, 96-102: epsilon = 0.001,
subsequence 16, buffer 4, num_sequences 1, N = 1000 / 10000), chains sharded over the torchrun ranks with NO
collective per iteration.  The reference's data file does not travel to the GPU box, so the
series are synthetic SVM draws with the same segment structure.

  python scripts/chains_demo.py [--chains 64] [--iters 200] [--N 1000]
  python -m torch.distributed.run --nproc-per-node 8 scripts/chains_demo.py --chains 64
Prints one JSON line: SGLD iterations/s per chain and in total, ensemble (one launch / iteration) vs chain-by-chain."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import torch
import sgmcmc_ssm_b200 as sg
from sgmcmc_ssm_b200 import parallel
from sgmcmc_ssm_b200.ensemble import ChainEnsemble
from sgmcmc_ssm_b200.models.svm import SeqSVMSampler, SVMParameters, SVMPrior


def data(seed=12345, n_seq=49, total=5907):
    rs = np.random.RandomState(seed)
    cuts = np.sort(rs.choice(np.arange(7, total - 7, 7), n_seq - 1, replace=False))
    lens = np.diff(np.concatenate([[0], cuts, [total]]))
    seqs = []
    for L in lens:
        x, y = rs.normal() * 2.0, np.zeros((int(L), 1))
        for t in range(int(L)):
            x = 0.95 * x + np.sqrt(0.5) * rs.normal()
            y[t, 0] = np.sqrt(0.5) * np.exp(0.5 * x) * rs.normal()
        seqs.append(y)
    return seqs


def chain(seqs, seed):
    np.random.seed(seed)
    s = SeqSVMSampler(n=1, m=1, observations=seqs)        # exchange_rate_full_demo.py:96-99
    s.prior_init()
    s.project_parameters()
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--N", type=int, default=1000)
    args = ap.parse_args()
    rank, world, _ = parallel.init_distributed()
    lo, hi = parallel.shard_bounds(args.chains)
    seqs = data()
    seeds = [12345 + c for c in range(lo, hi)]
    kw = dict(kind="pf", pf="poyiadjis_N", N=args.N, subsequence_length=16, buffer_length=4, minibatch_size=1, num_sequences=1)
    sg.set_seed(7 + rank)
    ens = ChainEnsemble([chain(seqs, s) for s in seeds], seeds=seeds)
    for _ in range(5):
        ens.sample_sgld(epsilon=1e-3, **kw)
    torch.cuda.synchronize(); parallel.barrier()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        ens.sample_sgld(epsilon=1e-3, **kw)
    torch.cuda.synchronize(); parallel.barrier()
    dt_ens = parallel.allreduce_max(time.perf_counter() - t0)
    # chain by chain (one launch per chain and iteration: what a per-process port would do)
    solo = [chain(seqs, s) for s in seeds[:4]]
    n_solo = max(1, args.iters // 4)
    t0 = time.perf_counter()
    for _ in range(n_solo):
        for s in solo:
            s.sample_sgld(epsilon=1e-3, **kw); s.project_parameters()
    torch.cuda.synchronize()
    dt_solo = (time.perf_counter() - t0) / (n_solo * len(solo))
    A = np.array([float(np.ravel(s.parameters.A)[0]) for s in ens.samplers])
    if rank == 0:
        print(json.dumps({"workload": "SeqSVMSampler, 49 sequences / 5907 obs (synthetic, exchange-rate demo shape), SGLD eps=1e-3, "
                                      "S=16, B=4, num_sequences=1, N=%d" % args.N,
                          "chains": args.chains, "n_gpus": world, "iters": args.iters,
                          "ensemble_ms_per_iteration_all_chains": 1e3 * dt_ens / args.iters,
                          "ensemble_chain_iterations_per_sec": args.chains * args.iters / dt_ens,
                          "chain_by_chain_ms_per_iteration_per_chain": 1e3 * dt_solo,
                          "chain_by_chain_chain_iterations_per_sec_one_gpu": 1.0 / dt_solo,
                          "A_mean_rank0": float(A.mean()), "A_finite": bool(np.all(np.isfinite(A)))}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
