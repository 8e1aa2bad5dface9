"""A/B of the shared-memory kernel's thread shape (SGM_LIB_PATH selects the build): gradient latency at N = 1000 / 1024 and
SGLD iterations/s of the persistent kernel (LGSSM T=1000, N=1000)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
from sgmcmc_ssm_b200.device_loop import DeviceChains
from sgmcmc_ssm_b200.models.lgssm import LGSSMSampler, LGSSMParameters, generate_lgssm_data
rs = np.random.RandomState(0)
th = [0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0]
print("lib", os.environ.get("SGM_LIB_PATH", "default"))
for N in (512, 1000, 1024):
    for B in (1, 8, 64, 148, 296):
        it = sg.PFItems()
        for b in range(B):
            it.add(rs.normal(size=60) * 0.7, th, t1=10, tL=50, weights=np.ones(40) * 25.0, prior_mean=0.0, prior_var=10.0)
        p = sg.engine.PreparedPF("lgssm", "optimal", "poyiadjis_N", it, N, dtype="f32", path="small").upload()
        for k in range(3):
            p.launch(offset=k + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for k in range(20):
            p.launch(offset=10 + k)
        e1.record(); torch.cuda.synchronize()
        print("N=%d B=%d  %.4f ms" % (N, B, e0.elapsed_time(e1) / 20), flush=True)
np.random.seed(12345)
pl = LGSSMParameters(A=np.eye(1) * 0.9, C=np.eye(1), LQinv=np.eye(1) * np.sqrt(10.0), LRinv=np.eye(1))
dl = generate_lgssm_data(T=1000, parameters=pl)
sg.set_seed(1)
ch = DeviceChains([LGSSMSampler(n=1, m=1, observations=dl["observations"], parameters=pl.copy())], method="SGLD", epsilon=0.01,
                  pf="poyiadjis_N", N=1000, subsequence_length=40, buffer_length=10, minibatch_size=1)
ch.run(200).synchronize()
t0 = time.perf_counter(); ch.run(4000).synchronize(); dt = time.perf_counter() - t0
print("sgld it/s persistent N=1000: %.1f" % (4000 / dt))
