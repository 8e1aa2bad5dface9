"""compute-sanitizer target: one small call of every round-2 kernel family (shared-memory kernel, cluster kernel, cooperative
tile kernel, fast-mode tile kernels f32 / f64, device-resident SG-MCMC loop: per-iteration and persistent)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
from sgmcmc_ssm_b200.device_loop import DeviceChains
from sgmcmc_ssm_b200.models.svm import SeqSVMSampler, SVMParameters
rs = np.random.RandomState(0)
th = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
it = sg.PFItems()
for b in range(3):
    it.add(rs.normal(size=6 + b) * 0.7, th, t1=1, tL=5, weights=1 + rs.rand(4), prior_mean=0.0, prior_var=10.0)
for N, path in ((300, "small"), (1000, "small"), (2000, "small"), (700, "cluster"), (1500, "cluster"), (3000, "cluster"),
                (3000, "tiles"), (5000, "tiles"), (256, "tiles")):
    for pf, kw in (("poyiadjis_N", {}), ("nemeth", dict(lambduh=0.9)), ("filter", {})):
        for dtype in ("f32", "f64"):
            r = sg.run_pf("svm", "prior", pf, it, N, dtype=dtype, path=path, seed=1, offset=1, **kw)
            assert np.all(np.isfinite(r.grad)), (N, path, pf, dtype)
big = sg.PFItems()
for b in range(160):
    big.add(rs.normal(size=5) * 0.7, th, t1=1, tL=4, prior_mean=0.0, prior_var=10.0)
for pf, kw in (("poyiadjis_N", {}), ("nemeth", dict(lambduh=0.9)), ("filter", {})):
    for dtype in ("f32", "f64"):
        r = sg.run_pf("svm", "prior", pf, big, 3000, dtype=dtype, path="tiles", seed=1, offset=1, **kw)     # per-step launches, fast modes (ragged)
        assert np.all(np.isfinite(r.grad))
seqs = [rs.normal(size=(n, 1)) for n in (30, 9, 25)]
ps = SVMParameters(A=np.eye(1) * 0.9, LQinv=np.eye(1) * 1.2, LRinv=np.eye(1) * 1.1)
for N, M in ((300, 1), (1500, 1), (300, 2), (3000, 1)):
    ch = DeviceChains([SeqSVMSampler(n=1, m=1, observations=seqs, parameters=ps.copy()) for _ in range(3)], method="SGLD", epsilon=1e-3,
                      pf="poyiadjis_N", N=N, subsequence_length=8, buffer_length=2, minibatch_size=M, num_sequences=1, trace_every=1, max_trace_rows=4)
    ch.run(3, graph=False)
    ch.pull_parameters()
print("sanitizer target ok")
