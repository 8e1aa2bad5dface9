#!/usr/bin/env python3
"""BASELINE configs[3]: independent SGLD chains of a Seq SVM sampler on the EUR/USD hourly returns
(data/EURUS_hourly.npz = the hourly series of the reference's data/EURUS_processed.npz; split on gaps > 6 h keeping
segments longer than 6 observations -> 49 sequences / 5907 observations, demo/exchange_rate/exchange_rate_full_demo.py:16-42;
`SeqSVMSampler(n=1, m=1)`, `prior_init(); project_parameters()`, SGLD epsilon = 0.001, subsequence 16, buffer 4,
num_sequences 1, pf = poyiadjis_N, N = 1000 / 10000, :96-102), chains sharded over the torchrun ranks with NO collective
per iteration; chain c is seeded 12345 + c on whichever rank it lands.

  python scripts/chains_demo.py [--chains 64] [--iters 2000] [--N 1000]
  python -m torch.distributed.run --nproc-per-node 8 scripts/chains_demo.py --chains 64
Prints one JSON line: chain-iterations/s of the device-resident loop (DeviceChains: the whole iteration on the GPU),
of the host ensemble (one batched launch per iteration, numpy update per chain) and of a chain-by-chain host loop."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import torch
import sgmcmc_ssm_b200 as sg
from sgmcmc_ssm_b200 import parallel
from sgmcmc_ssm_b200.device_loop import DeviceChains
from sgmcmc_ssm_b200.ensemble import ChainEnsemble
from sgmcmc_ssm_b200.models.svm import SeqSVMSampler


def eurus_sequences():
    """exchange_rate_full_demo.py:16-42 (np.timedelta64 instead of pandas.Timedelta: same comparison)."""
    z = np.load(os.path.join(ROOT, "data", "EURUS_hourly.npz"))
    dates = z["hourly_date"]
    observations = z["hourly_log_returns"].reshape(-1, 1) * 1000
    gap_indices = np.where(np.diff(dates) > np.timedelta64(6, "h"))[0].tolist()
    split = []
    for start, end in zip([0] + gap_indices, gap_indices + [observations.size]):
        if end - start > 6:
            split.append(observations[start:end])
    return split


def chain(seqs, seed):
    np.random.seed(seed)
    s = SeqSVMSampler(n=1, m=1, observations=seqs)        # exchange_rate_full_demo.py:96-99
    s.prior_init()
    s.project_parameters()
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--N", type=int, default=1000)
    ap.add_argument("--host-iters", type=int, default=40)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rank, world, _ = parallel.init_distributed()
    lo, hi = parallel.shard_bounds(args.chains)
    seqs = eurus_sequences()
    assert len(seqs) == 49 and sum(len(s) for s in seqs) == 5907
    seeds = [12345 + c for c in range(lo, hi)]
    kw = dict(kind="pf", pf="poyiadjis_N", N=args.N, subsequence_length=16, buffer_length=4, minibatch_size=1, num_sequences=1)
    # ---- device-resident loop: every iteration of every local chain on the GPU, chain streams keyed by the global index
    sg.set_seed(7)
    chains = DeviceChains([chain(seqs, s) for s in seeds], method="SGLD", epsilon=1e-3, chain_id_base=lo,
                          trace_every=max(1, args.iters // 50), max_trace_rows=60, **kw)
    chains.run(max(16, args.iters // 20)).synchronize()
    torch.cuda.synchronize(); parallel.barrier()
    t0 = time.perf_counter()
    chains.run(args.iters).synchronize()
    parallel.barrier()
    dt_dev = parallel.allreduce_max(time.perf_counter() - t0)
    params = chains.pull_parameters()
    A = np.array([float(np.ravel(p.A)[0]) for p in params])
    # ---- host ensemble (round-1 path): one batched launch per iteration, numpy SGLD update per chain
    ens = ChainEnsemble([chain(seqs, s) for s in seeds], seeds=seeds)
    for _ in range(3):
        ens.sample_sgld(epsilon=1e-3, **kw)
    torch.cuda.synchronize(); parallel.barrier()
    t0 = time.perf_counter()
    for _ in range(args.host_iters):
        ens.sample_sgld(epsilon=1e-3, **kw)
    torch.cuda.synchronize(); parallel.barrier()
    dt_ens = parallel.allreduce_max(time.perf_counter() - t0)
    # ---- chain by chain on the host (what a per-process port of the reference script would do)
    solo = chain(seqs, seeds[0])
    n_solo = max(4, args.host_iters // 2)
    t0 = time.perf_counter()
    for _ in range(n_solo):
        solo.sample_sgld(epsilon=1e-3, **kw); solo.project_parameters()
    torch.cuda.synchronize()
    dt_solo = (time.perf_counter() - t0) / n_solo
    ok = parallel.allreduce_sum(np.array([float(np.all(np.isfinite(A)))]))[0] == world
    if rank == 0:
        line = {"workload": "SeqSVMSampler on EUR/USD hourly returns x 1000 (49 sequences / 5907 obs, data/EURUS_hourly.npz), SGLD eps=1e-3, "
                            "S=16, B=4, num_sequences=1, pf=poyiadjis_N, N=%d, prior_init + project_parameters, seeds 12345+c" % args.N,
                "chains": args.chains, "chains_per_gpu": hi - lo, "n_gpus": world, "iters": args.iters,
                "device_loop_chain_iterations_per_sec": args.chains * args.iters / dt_dev,
                "device_loop_ms_per_iteration_all_chains": 1e3 * dt_dev / args.iters,
                "device_loop_persistent_kernel": bool(chains.persistent),
                "host_ensemble_chain_iterations_per_sec": args.chains * args.host_iters / dt_ens,
                "host_ensemble_ms_per_iteration_all_chains": 1e3 * dt_ens / args.host_iters,
                "chain_by_chain_host_iterations_per_sec_one_chain": 1.0 / dt_solo,
                "A_mean_rank0": float(A.mean()), "A_std_rank0": float(A.std()), "all_chains_finite": bool(ok)}
        print(json.dumps(line), flush=True)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(line, f, indent=1)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
