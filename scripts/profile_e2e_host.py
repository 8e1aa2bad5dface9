"""cProfile of the host side of one end-to-end gradient (bench workload: SVM, N = 2^16, 512 windows)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
from sgmcmc_ssm_b200.models.svm import SVMSampler, SVMParameters, generate_svm_data
np.random.seed(1)
params = SVMParameters(A=np.eye(1) * 0.95, LQinv=np.eye(1) * np.sqrt(2.0), LRinv=np.eye(1) * np.sqrt(2.0))
y = generate_svm_data(T=10000, parameters=params)["observations"]
sampler = SVMSampler(n=1, m=1, observations=y, parameters=params)
kw = dict(kind="pf", pf="poyiadjis_N", N=65536, subsequence_length=40, buffer_length=10, minibatch_size=512, dtype="f32", rng="philox",
          resample="multinomial_sorted")
for _ in range(3):
    sampler.noisy_gradient(**kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    sampler.noisy_gradient(**kw)
torch.cuda.synchronize()
print("e2e ms per gradient", (time.perf_counter() - t0) / 10 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    sampler.noisy_gradient(**kw)
pr.disable()
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
