"""Host-side profile of the SGLD loop on BASELINE configs[0] (LGSSM T=1000, N=1000, S=40, B=10, minibatch 1):
cProfile of 300 iterations, top entries by own time and by cumulative time."""
import cProfile, io, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
from sgmcmc_ssm_b200.models.lgssm import LGSSMSampler, LGSSMParameters, generate_lgssm_data
np.random.seed(12345)
p = LGSSMParameters(A=np.eye(1) * 0.9, C=np.eye(1), LQinv=np.eye(1) * np.sqrt(10.0), LRinv=np.eye(1))
data = generate_lgssm_data(T=1000, parameters=p)
s = LGSSMSampler(n=1, m=1, observations=data["observations"], parameters=p.copy())
kw = dict(epsilon=0.01, kind="pf", pf="poyiadjis_N", N=1000, subsequence_length=40, buffer_length=10, minibatch_size=1)
for _ in range(20):
    s.sample_sgld(**kw); s.project_parameters()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    s.sample_sgld(**kw); s.project_parameters()
torch.cuda.synchronize()
print("plain: %.1f it/s" % (300 / (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable()
for _ in range(300):
    s.sample_sgld(**kw); s.project_parameters()
pr.disable()
for key in ("tottime", "cumulative"):
    buf = io.StringIO(); pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(28); print(buf.getvalue()[:6000])
