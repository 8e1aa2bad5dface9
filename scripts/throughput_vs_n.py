"""Device-timed step throughput of the O(N) gradient path against the particle count (BASELINE config 2's range
N = 2^10..2^16, the reference's own N = 1000 / 10000, and config 5's N = 2^20), SVM prior kernel, poyiadjis_N, f32,
order-statistics resampling, 60-step windows.  The batch keeps ~3.4e7 particles in flight (state > L2) except where
noted.  Writes one JSON document (profiles/throughput_r01_vs_n.json)."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
import sgmcmc_ssm_b200 as sg

rs = np.random.RandomState(0)
TH = [0.95, np.sqrt(2.0), 2.0, np.sqrt(2.0), 2.0]
T, PEAK, BYTES = 60, 6450.6e9, 40
rows = []
for N in (1000, 1024, 2048, 4096, 8192, 10000, 16384, 32768, 65536, 100000, 262144, 1 << 20):
    B = max(8, min(4096, (1 << 25) // N))
    it = sg.PFItems()
    for b in range(B):
        it.add(rs.normal(size=T) * 0.7, TH, t1=10, tL=50, weights=np.ones(40) * 250.0, prior_mean=0.0, prior_var=10.0)
    prep = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", it, N, dtype="f32", rng="philox",
                                resample="multinomial_sorted").upload()
    for k in range(3):
        prep.launch(offset=k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    reps = 5
    for k in range(reps):
        prep.launch(offset=5 + k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ps = B * N * T / (ms * 1e-3)
    path = "single-launch kernel (one CTA per item)" if (N <= 2048 and B <= 1024) else (
        "FAST step kernel" if N % 256 == 0 else "FAST step kernel, ragged last tile")
    rows.append(dict(N=N, items=B, ms_per_gradient_batch=ms, particle_steps_per_s=ps, hbm_roofline_frac=ps * BYTES / PEAK,
                     state_bytes=B * N * 20, kernel=path))
    print("N=%-8d items=%-5d %8.3f ms  %.3e p-s/s  %.1f %%  %s" % (N, B, ms, ps, 100 * ps * BYTES / PEAK, path), flush=True)
json.dump(dict(note=__doc__, peak_hbm_GBps=PEAK / 1e9, bytes_per_particle_step=BYTES, rows=rows),
          open(os.path.join(ROOT, "gpurun_out", "throughput_vs_n.json"), "w"), indent=1)
