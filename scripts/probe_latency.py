"""Latency regime and SG-MCMC loop probe (not the bench): single-item gradient latency, small-N batch throughput of the
shared-memory kernel vs the per-step tile kernels, SGLD iterations/s of the device loop (persistent kernel, launch
sequence per iteration with / without CUDA graph) and of the host loop.   python scripts/probe_latency.py [--json out]"""
import argparse, json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg
from sgmcmc_ssm_b200.device_loop import DeviceChains
from sgmcmc_ssm_b200.models.lgssm import LGSSMSampler, LGSSMParameters, generate_lgssm_data
from sgmcmc_ssm_b200.models.svm import SVMSampler, SeqSVMSampler, SVMParameters, generate_svm_data

ap = argparse.ArgumentParser()
ap.add_argument("--json", default=None)
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
out = {}
rs = np.random.RandomState(0)
TH = {"svm": ([0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0], "prior", 10.0), "lgssm": ([0.9, np.sqrt(10.0), 10.0, 1.0, 1.0, 1.0], "optimal", 10.0)}


def timed(fn, reps):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def prep(model, N, B, path="auto", dtype="f32", T=60):
    th, kern, pv = TH[model]
    it = sg.PFItems()
    for b in range(B):
        it.add(rs.normal(size=T) * 0.7, th, t1=10, tL=T - 10, weights=np.ones(T - 20) * 25.0, prior_mean=0.0, prior_var=pv)
    return sg.engine.PreparedPF(model, kern, "poyiadjis_N", it, N, dtype=dtype, rng="philox", resample="multinomial_sorted", path=path).upload()

# ---- 1. gradient latency / throughput at small N -------------------------------------------------------------------
k = [0]
def go(p):
    def f():
        k[0] += 1
        (p.launch_graph if p.graph_eligible() else p.launch)(offset=k[0])
    return f
for model in ("svm", "lgssm"):
    for N in (256, 1000, 1024, 2048):
        for B, reps in ((1, 50), (8, 30), (64, 20), (296, 10), (4096, 3)):
            for path in ("auto", "small", "tiles"):
                if args.quick and (B == 296 or (path == "tiles" and B == 64)):
                    continue
                if path == "small" and (N <= 512 or B > 64):
                    continue                    # auto already is the shared-memory kernel there
                p = prep(model, N, B, path)
                ms = timed(go(p), reps)
                key = "grad_ms_%s_N%d_B%d_%s" % (model, N, B, path)
                out[key] = ms
                print("%-34s %9.4f ms  %.3e particle-steps/s  launches %d" % (key, ms, N * 60 * B / (ms * 1e-3), p.launches), flush=True)
# larger N single item (tile kernels + CUDA graph)
for N in (4096, 8192, 16384, 65536):
    for path in ("auto", "tiles"):
        p = prep("svm", N, 1, path)
        ms = timed(go(p), 20)
        out["grad_ms_svm_N%d_B1_%s" % (N, path)] = ms
        print("grad_ms_svm_N%d_B1_%s  %9.4f ms  launches %d" % (N, path, ms, p.launches), flush=True)

# ---- 2. SGLD iterations / s ------------------------------------------------------------------------------------------
np.random.seed(12345)
pl = LGSSMParameters(A=np.eye(1) * 0.9, C=np.eye(1), LQinv=np.eye(1) * np.sqrt(10.0), LRinv=np.eye(1))
dl = generate_lgssm_data(T=1000, parameters=pl)
ps = SVMParameters(A=np.eye(1) * 0.95, LQinv=np.eye(1) * np.sqrt(2.0), LRinv=np.eye(1) * np.sqrt(2.0))
ds = generate_svm_data(T=10000, parameters=ps)


def sgld_rate(make, N, iters, **kw):
    sg.set_seed(1)
    ch = DeviceChains([make()], method="SGLD", epsilon=0.01, pf="poyiadjis_N", N=N, subsequence_length=40, buffer_length=10,
                      minibatch_size=1, **kw)
    ch.run(max(8, iters // 10)).synchronize()
    t0 = time.perf_counter()
    ch.run(iters).synchronize()
    dt = time.perf_counter() - t0
    ch.pull_parameters()
    return iters / dt, ch

mk_l = lambda: LGSSMSampler(n=1, m=1, observations=dl["observations"], parameters=pl.copy())
mk_s = lambda: SVMSampler(n=1, m=1, observations=ds["observations"], parameters=ps.copy())
for name, mk, N, iters in (("lgssm_T1000_N1000", mk_l, 1000, 4000), ("svm_T10000_N1024", mk_s, 1024, 4000),
                           ("svm_T10000_N8192", mk_s, 8192, 400), ("svm_T10000_N65536", mk_s, 65536, 100)):
    for mode, kw in (("persistent", {}), ("per_iteration_graph", dict(persistent=False)), ):
        if mode == "persistent" and N > 2048:
            continue
        r, ch = sgld_rate(mk, N, iters, **kw)
        out["sgld_its_%s_%s" % (name, mode)] = r
        print("sgld it/s %-22s %-22s %10.1f  (A=%.4f)" % (name, mode, r, float(ch.samplers[0].parameters.A[0, 0])), flush=True)
# host loop (one Python iteration per step) for comparison
s = mk_l()
kw = dict(epsilon=0.01, kind="pf", pf="poyiadjis_N", N=1000, subsequence_length=40, buffer_length=10, minibatch_size=1)
for _ in range(5):
    s.sample_sgld(**kw); s.project_parameters()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200):
    s.sample_sgld(**kw); s.project_parameters()
torch.cuda.synchronize()
out["sgld_its_lgssm_T1000_N1000_host_loop"] = 200 / (time.perf_counter() - t0)
print("sgld it/s lgssm host loop %.1f" % out["sgld_its_lgssm_T1000_N1000_host_loop"], flush=True)
# through the public API: sampler.fit on the device
s = mk_l()
s.fit("SGLD", 50, epsilon=0.01, subsequence_length=40, buffer_length=10, kind="pf", pf_kwargs=dict(pf="poyiadjis_N", N=1000))
t0 = time.perf_counter()
s.fit("SGLD", 4000, epsilon=0.01, subsequence_length=40, buffer_length=10, kind="pf", pf_kwargs=dict(pf="poyiadjis_N", N=1000))
out["sgld_its_lgssm_T1000_N1000_fit_api"] = 4000 / (time.perf_counter() - t0)
print("sgld it/s lgssm sampler.fit(4000) %.1f" % out["sgld_its_lgssm_T1000_N1000_fit_api"], flush=True)

# ---- 3. many chains (configs[3] shape: Seq SVM, 49 sequences, S = 16, B = 4, one sequence per iteration) -------------
lens = rs.randint(20, 300, size=49)
seqs = [rs.normal(size=(int(n), 1)) for n in lens]
for N in (1000, 10000):
    for C in (8, 64):
        sg.set_seed(2)
        ch = DeviceChains([SeqSVMSampler(n=1, m=1, observations=seqs, parameters=ps.copy()) for _ in range(C)], method="SGLD",
                          epsilon=1e-3, pf="poyiadjis_N", N=N, subsequence_length=16, buffer_length=4, minibatch_size=1, num_sequences=1)
        iters = 400 if N == 1000 else 100
        ch.run(16).synchronize()
        t0 = time.perf_counter()
        ch.run(iters).synchronize()
        dt = time.perf_counter() - t0
        ch.pull_parameters()
        out["chain_its_seqsvm_N%d_C%d" % (N, C)] = C * iters / dt
        print("chains seqsvm N=%d C=%d  %.1f chain-iterations/s (%.3f ms / iteration, persistent=%s)" % (N, C, C * iters / dt, 1e3 * dt / iters, ch.persistent), flush=True)
if args.json:
    json.dump(out, open(args.json, "w"), indent=1)
