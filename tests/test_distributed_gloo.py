"""World-size-2 gloo test (CPU) of the multi-GPU host logic: contiguous item sharding, the single
all-reduce of per-rank gradient sums, and the sampler's `distributed=True` path (the device call is
replaced by a deterministic CPU stand-in that depends only on the window and the global item index, so
the sharded result must equal the single-process result bit for bit in a fixed reduction order)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_batch(self, windows, parameters, item_id_base=0, **kwargs):
    grads = []
    for k, w in enumerate(windows):
        s = float(np.sum(w["observations"])) + 0.01 * (item_id_base + k)
        wt = 1.0 if w["weights"] is None else float(np.sum(w["weights"]))
        grads.append(dict(LRinv_vec=s, LQinv_vec=2 * s + wt, A=np.sin(s)))
    return grads, None


def _fake_packed(self, packed, parameters, item_id_base=0, **kwargs):
    """Stand-in of Helper.pf_gradient_sum_packed (the vectorised minibatch path): same per-item values as
    _fake_batch, summed over the shard."""
    off = np.concatenate([[0], np.cumsum(packed.T_buf)])
    tot = dict(LRinv_vec=0.0, LQinv_vec=0.0, A=0.0)
    for k in range(len(packed)):
        s = float(np.sum(packed.obs_flat[off[k]:off[k + 1]])) + 0.01 * (item_id_base + k)
        n = int(packed.tL[k] - packed.t1[k])
        wt = 1.0 if packed.wts_off[k] < 0 else float(np.sum(packed.wts_flat[packed.wts_off[k]:packed.wts_off[k] + n]))
        tot["LRinv_vec"] += s
        tot["LQinv_vec"] += 2 * s + wt
        tot["A"] += np.sin(s)
    return tot, None


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
    from sgmcmc_ssm_b200 import parallel
    from sgmcmc_ssm_b200.models.svm import SVMSampler, SVMParameters, SVMHelper
    r, w, _ = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    # (1) sharding is a partition
    covered = []
    for n in (1, 2, 7, 256):
        lo, hi = parallel.shard_bounds(n)
        allb = parallel.allreduce_sum(np.array([hi - lo]))
        assert int(allb[0]) == n
        covered.append((lo, hi))
    # (2) all-reduce of gradient sums
    tot = parallel.allreduce_sum(np.array([1.0 + rank, 10.0 * (rank + 1)]))
    np.testing.assert_array_equal(tot, [3.0, 30.0])
    assert parallel.allreduce_max(float(rank)) == 1.0
    # (3) sampler path
    SVMHelper.pf_gradient_estimate_batch = _fake_batch
    SVMHelper.pf_gradient_sum_packed = _fake_packed
    rs = np.random.RandomState(0)
    y = rs.normal(size=(500, 1))
    p = SVMParameters(A=np.eye(1) * 0.9, LQinv=np.eye(1), LRinv=np.eye(1))
    s = SVMSampler(n=1, m=1, observations=y, parameters=p)
    np.random.seed(5)
    g_dist = s.noisy_gradient(kind="pf", N=8, subsequence_length=20, buffer_length=5, minibatch_size=7, distributed=True)
    np.random.seed(5)
    g_one = s.noisy_gradient(kind="pf", N=8, subsequence_length=20, buffer_length=5, minibatch_size=7)
    for k in g_one:
        np.testing.assert_allclose(g_dist[k], g_one[k], rtol=1e-13, atol=1e-13)
    # the vectorised minibatch path and the per-window loop (explicit buffer_dicts) agree
    np.random.seed(5)
    bds = [s._random_subsequence_and_buffers(buffer_length=5, subsequence_length=20, T=500) for _ in range(7)]
    g_loop = s.noisy_gradient(kind="pf", N=8, subsequence_length=20, buffer_length=5, minibatch_size=7, buffer_dicts=bds,
                              distributed=True)
    for k in g_one:
        np.testing.assert_allclose(g_loop[k], g_one[k], rtol=1e-12, atol=1e-12)
    parallel.barrier()
    out.put((rank, covered, {k: float(np.ravel(v)[0]) for k, v in g_dist.items()}))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo_sharding_and_allreduce():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    (r0, c0, g0), (r1, c1, g1) = res
    for (lo0, hi0), (lo1, hi1) in zip(c0, c1):
        assert lo0 == 0 and hi0 == lo1 and hi1 >= lo1          # contiguous, ordered, disjoint
    assert g0 == g1                                             # every rank ends with the same gradient
