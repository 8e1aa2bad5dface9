"""torchrun smoke of the distributed sampler path with a SMALL minibatch (CUDA-graph replay on every rank, NCCL
all-reduce of the gradient sums): every rank must end with the same gradient, equal to the single-process result."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import torch
import sgmcmc_ssm_b200 as sg
from sgmcmc_ssm_b200 import parallel
from sgmcmc_ssm_b200.models.svm import SVMSampler, SVMParameters
rank, world, _ = parallel.init_distributed()
rs = np.random.RandomState(0)
y = rs.normal(size=(2000, 1))
p = SVMParameters(A=np.eye(1) * 0.95, LQinv=np.eye(1) * np.sqrt(2.0), LRinv=np.eye(1) * np.sqrt(2.0))
s = SVMSampler(n=1, m=1, observations=y, parameters=p)
kw = dict(kind="pf", pf="poyiadjis_N", N=65536, subsequence_length=40, buffer_length=10, minibatch_size=2 * world, dtype="f32")
out = []
for it in range(4):
    sg.set_seed(100 + it); np.random.seed(it)
    g = s.noisy_gradient(distributed=(world > 1), **kw)
    out.append(np.array([float(np.ravel(v)[0]) for v in g.values()]))
out = np.array(out)
ref = []
for it in range(4):                       # the same minibatch on one rank, no sharding
    sg.set_seed(100 + it); np.random.seed(it)
    g = s.noisy_gradient(**kw)
    ref.append(np.array([float(np.ravel(v)[0]) for v in g.values()]))
ref = np.array(ref)
assert np.all(np.isfinite(out))
np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
tot = parallel.allreduce_sum(out.ravel())
np.testing.assert_allclose(tot, world * out.ravel(), rtol=1e-12)
if rank == 0:
    print("dist smoke ok: world", world, "grad", out[0])
if world > 1:
    torch.distributed.destroy_process_group()
