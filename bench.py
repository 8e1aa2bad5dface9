#!/usr/bin/env python3
"""bench.py -- buffered particle-filter gradient throughput (BASELINE.json metric: particle-steps/sec).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on the host CPU cores

Workload (BASELINE.json configs[1]): SVM synthetic T=10000 (A, Q, R) = (0.95, 0.5, 0.5), Poyiadjis O(N)
smoother, N = 2^16 particles per subsequence, subsequence 40 + buffers 10 (T_buf = 60 steps), a minibatch
of M = 512 subsequences per GPU (weak scaling), f32 particle arithmetic with f64 CDF offsets, device Philox
randoms, order-statistics multinomial resampling (same law as the reference's multinomial).  The batch runs
as two halves on two CUDA streams (sgm_pf_desc.aux_stream).  `--particles 1048576 --minibatch 128` is one
GPU's share of configs[4].

A "step" = one noisy-gradient evaluation of the minibatch = one pass of the hot path over one batch.
  value : device-timed (CUDA events, max over ranks), inputs resident in HBM; includes the per-step
          NCCL all-reduce of the gradient sums when N > 1.
  e2e   : the same metric through the public API `sampler.noisy_gradient(kind='pf', ...)` with HOST
          observation buffers: window draws, packing, H2D, all kernels, D2H and the all-reduce inside the
          timed region.
  roofline : algorithmic 40 B / particle-step x particles of one time step / time of one time step
          (CUDA events on the launch stream around the step loop / T_buf), against MEASURED_PEAKS.json.
  extra : single-subsequence latency (CUDA-graph replay), SGLD iterations/s (configs[0]), O(N^2) pair-steps/s on
          the tensor cores and on the FP32 pipe, PaRIS particle-steps/s (configs[2] shape).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "particle-steps/sec (buffered PF grad, SVM N=2^16)"
UNIT = "particle-steps/s"
N_PARTICLES = 1 << 16
SUBSEQ, BUFFER, T_SERIES = 40, 10, 10000
ALG_BYTES = 40            # e * (2 (n + p) + 2), SVM f32: 4 * (2 * 4 + 2)


def svm_series(T=T_SERIES, seed=12345):
    """Synthetic SVM series under the reference's generator semantics (svm/parameters.py:75-135),
    vectorised-free re-statement: x_t = A x_{t-1} + N(0, Q), y_t ~ N(0, R exp(x_t))."""
    rs = np.random.RandomState(seed)
    A, Q, R = 0.95, 0.5, 0.5
    x = np.sqrt(Q / (1 - A * A)) * rs.normal()
    y = np.zeros((T, 1))
    for t in range(T):
        x = A * x + np.sqrt(Q) * rs.normal()
        y[t, 0] = np.sqrt(R) * np.exp(0.5 * x) * rs.normal()
    return y


def svm_theta():
    LQinv, LRinv = np.sqrt(1 / 0.5), np.sqrt(1 / 0.5)
    return dict(A=0.95, LQinv=LQinv, Qinv=LQinv * LQinv + 1e-16, LRinv=LRinv, Rinv=LRinv * LRinv + 1e-16)


def uniform_partition_weights(S, T, start):
    """Per-step importance weights of a 'uniform'-partition subsequence [start, start + S)
    (sgmcmc_sampler.py:1994-2008: (T - S + 1) / number of subsequences covering t).  Workload construction only --
    the GPU arm imports nothing from oracle/."""
    t = np.arange(start, start + S)
    cap = min(S, T - S + 1)
    if start + S <= 2 * S:
        num = np.minimum(t + 1, cap)
    elif start >= T - 2 * S - 1:
        num = np.minimum(T - t, cap)
    else:
        num = np.full(S, S)
    return np.ones(S, dtype=float) * (T - S + 1) / num


def draw_windows(n, T=T_SERIES, seed=777):
    """n subsequence starts + 'uniform' partition weights (sgmcmc_sampler.py:1969-2017)."""
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        start = int(rs.randint(0, T - SUBSEQ + 1))
        lo, hi = max(0, start - BUFFER), min(T, start + SUBSEQ + BUFFER)
        out.append(dict(start=start, lo=lo, hi=hi, weights=uniform_partition_weights(SUBSEQ, T, start)))
    return out


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref, the unmodified package; else the oracle port) on all host cores
# ------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def ref_kind():
    """"reference" when the vendored copy of the unmodified reference (oracle/make_ref.py) is present, else "port"."""
    return "reference" if os.path.isdir(os.path.join(REF_DIR, "sgmcmc_ssm")) else "port"


def _cpu_one_gradient(args):
    y, w, seed = args
    if ref_kind() == "reference":
        # the UNMODIFIED reference: sgmcmc_ssm.models.svm.SVMHelper.pf_gradient_estimate (svm/helper.py:67-128)
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import warnings
        warnings.filterwarnings("ignore")
        from sgmcmc_ssm.models.svm import SVMHelper, SVMParameters
        th = svm_theta()
        params = SVMParameters(A=np.eye(1) * th["A"], LQinv=np.eye(1) * th["LQinv"], LRinv=np.eye(1) * th["LRinv"])
        obs = y[w["lo"]:w["hi"]]
        np.random.seed(seed)
        g = SVMHelper(n=1, m=1).pf_gradient_estimate(
            observations=obs, parameters=params, subsequence_start=w["start"] - w["lo"],
            subsequence_end=w["start"] - w["lo"] + SUBSEQ, weights=w["weights"], pf="poyiadjis_N", N=N_PARTICLES)
        return obs.shape[0] * N_PARTICLES, [float(g["LRinv_vec"]), float(g["LQinv_vec"]), float(np.ravel(g["A"])[0])]
    from oracle import pf_oracle as po
    rng = po.LegacyStream(seed, native_choice=True)       # np.random.choice(range(N), p=...) like pf.py:28-29
    obs = y[w["lo"]:w["hi"]]
    g = po.pf_gradient_estimate("svm", obs, svm_theta(), rng, subsequence_start=w["start"] - w["lo"],
                                subsequence_end=w["start"] - w["lo"] + SUBSEQ, weights=w["weights"],
                                pf="poyiadjis_N", N=N_PARTICLES)
    return obs.shape[0] * N_PARTICLES, [g["LRinv_vec"], g["LQinv_vec"], g["A"]]


def cpu_rate(y, windows, cores, steps, warmup):
    """particle-steps/s of the numpy port: each step = `cores` independent subsequence gradients, one per
    process (the reference itself is single-threaded; independent processes are its fair multi-core use)."""
    import multiprocessing as mp
    per_step = cores
    jobs = [(y, windows[i % len(windows)], 1000 + i) for i in range((steps + warmup) * per_step)]
    if cores == 1:
        for j in jobs[:warmup * per_step]:
            _cpu_one_gradient(j)
        t0 = time.perf_counter()
        done = sum(_cpu_one_gradient(j)[0] for j in jobs[warmup * per_step:])
        dt = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            if warmup:
                pool.map(_cpu_one_gradient, jobs[:warmup * per_step], chunksize=1)
            t0 = time.perf_counter()
            done = sum(r[0] for r in pool.map(_cpu_one_gradient, jobs[warmup * per_step:], chunksize=1))
            dt = time.perf_counter() - t0
    return done / dt, dt, done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    cores = os.cpu_count() or 1
    y = svm_series()
    windows = draw_windows(512)
    rate, dt, done = cpu_rate(y, windows, cores, args.steps, min(args.warmup, 1))
    sample = "{0} independent subsequence gradients per step (one per core), N=2^16, T_buf<=60".format(cores)
    line = {"metric": METRIC, "value": rate, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args.gpus, args.minibatch), reference_arm_work_per_step=(
                "{0} subsequence gradients per step (one per host core), not the GPU arm's minibatch: a RATE comparison "
                "on the same per-item workload".format(cores)), reference_arm_items_per_step=cores),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": ref_kind(), "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


def workload_config(world, M):
    return {"workload": "SVM synthetic T=10000 (A,Q,R)=(0.95,0.5,0.5), buffered PF gradient, pf=poyiadjis_N, "
                        "N={2} particles/subsequence, subsequence 40 + buffer 10 (T_buf<=60), "
                        "minibatch {0} subsequences/GPU x {1} GPU".format(M, world, N_PARTICLES),
            "minibatch_per_gpu": M, "n_particles": N_PARTICLES, "subsequence_length": SUBSEQ,
            "buffer_length": BUFFER, "pf": "poyiadjis_N", "resample": "multinomial_sorted", "rng": "philox",
            "parallelism": "items sharded over {0} GPU, one NCCL all-reduce of gradient sums per step".format(world),
            "l2": "particle state per GPU = M*N*(16+4)B*2 buffers >> 126 MB L2 (inputs larger than L2)"}


# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_gpu(args):
    import torch
    import sgmcmc_ssm_b200 as sg
    from sgmcmc_ssm_b200 import parallel
    from sgmcmc_ssm_b200.models.svm import SVMSampler, SVMParameters
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the one JSON line
    rank, world, local = parallel.init_distributed()
    dev = torch.device("cuda", torch.cuda.current_device())
    M = args.minibatch
    y = svm_series()
    windows_all = draw_windows(M * world)
    lo, hi = parallel.shard_bounds(len(windows_all))
    th = svm_theta()
    theta = [th[k] for k in ("A", "LQinv", "Qinv", "LRinv", "Rinv")]
    sg.set_seed(12345)

    # ---- (A) device-resident: inputs packed and uploaded once, K timed launches -----------------------
    items = sg.PFItems()
    for w in windows_all[lo:hi]:
        items.add(y[w["lo"]:w["hi"]], theta, t1=w["start"] - w["lo"], tL=w["start"] - w["lo"] + SUBSEQ,
                  weights=w["weights"], prior_mean=0.0, prior_var=10.0)
    prep = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", items, N_PARTICLES, dtype="f32", rng="philox",
                                resample="multinomial_sorted", item_id_base=lo)
    prep.upload()
    st = prep.st
    so = prep.base_out - st.dev_out.data_ptr()
    grad_view = st.dev_out[so:so + prep.B * 64].view(torch.float64).view(prep.B, 8)
    gsum = torch.zeros(8, dtype=torch.float64, device=dev)

    def one_step(k, events=None):
        prep.launch(offset=k + 1, step_events=events)
        torch.sum(grad_view, dim=0, out=gsum)
        if world > 1:
            torch.distributed.all_reduce(gsum)

    for k in range(args.warmup):
        one_step(k)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:           # create the handles
        a.record(); b.record()
    torch.cuda.synchronize()
    parallel.barrier()
    clocks = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(args.steps):
        one_step(args.warmup + k, ev[k])
    e1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    dev_ms = parallel.allreduce_max(e0.elapsed_time(e1))
    clk = clocks.stop() if rank == 0 else None
    step_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev])) / prep.max_T     # avg pf_step_kernel launch
    step_kernel_ms = parallel.allreduce_max(step_kernel_ms)
    local_ps = prep.particle_steps
    total_ps = parallel.allreduce_sum(np.array([float(local_ps)]))[0]
    value = total_ps * args.steps / (dev_ms * 1e-3)
    res = prep.download().wait()
    assert np.all(np.isfinite(res.grad)) and np.all(res.status == 0), "bench produced non-finite gradients"
    launches_per_step = prep.launches + 1 + (1 if world > 1 else 0)

    # ---- (B) end to end through the public API, host buffers -----------------------------------------
    params = SVMParameters(A=np.eye(1) * 0.95, LQinv=np.eye(1) * th["LQinv"], LRinv=np.eye(1) * th["LRinv"])
    sampler = SVMSampler(n=1, m=1, observations=y, parameters=params)
    np.random.seed(4242)                                   # identical subsequence draws on every rank
    api_kw = dict(kind="pf", pf="poyiadjis_N", N=N_PARTICLES, subsequence_length=SUBSEQ, buffer_length=BUFFER,
                  minibatch_size=M * world, dtype="f32", rng="philox", resample="multinomial_sorted",
                  distributed=(world > 1))
    for _ in range(max(1, min(args.warmup, 2))):
        sampler.noisy_gradient(**api_kw)
    torch.cuda.synchronize()
    parallel.barrier()
    t0 = time.perf_counter()
    e2e_local_ps, e2e_h2d, e2e_d2h = 0, 0, 0
    for _ in range(args.steps):
        g = sampler.noisy_gradient(**api_kw)
        e2e_local_ps += sampler.last_pf_info["particle_steps"]      # N x sum(T_buf) of this rank's packed windows
        e2e_h2d, e2e_d2h = sampler.last_pf_info["h2d_bytes"], sampler.last_pf_info["d2h_bytes"]
    torch.cuda.synchronize()
    parallel.barrier()
    e2e_s = parallel.allreduce_max(time.perf_counter() - t0)
    e2e_value = parallel.allreduce_sum(np.array([float(e2e_local_ps)]))[0] / e2e_s
    assert all(np.all(np.isfinite(v)) for v in g.values())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    multi = {}
    if not args.no_extras:                      # every rank takes part (sharded work + all-reduce inside)
        multi["f64"] = f64_block(sg, torch, parallel, items, lo, args, world, hbm, sampler, api_kw)
        multi["config5_strong"] = config5_strong(sg, torch, parallel, theta, world, hbm)
        multi["config4_chains"] = config4_chains(sg, torch, parallel, world)
    if world > 1:
        torch.distributed.destroy_process_group()
    if rank != 0:
        return 0
    achieved = ALG_BYTES * prep.B * N_PARTICLES / (step_kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        # measured DRAM bytes per particle-step (ncu --set full of the bench command) x particles of one time step
        traffic = json.load(open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")))["dram_bytes_per_particle_step"] * prep.B * N_PARTICLES
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world, M),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(e2e_h2d),
                "d2h_bytes_per_step": int(e2e_d2h), "ms_per_step": 1e3 * e2e_s / args.steps,
                "api": "SVMSampler.noisy_gradient(kind='pf', minibatch_size=M*n_gpus, ...): numpy window draws, packing, one "
                       "H2D copy, kernels, device-side sum of the item gradients, in-place NCCL all-reduce, one D2H copy"},
        "gpu_launches": int(launches_per_step * args.steps),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "pf_step_kernel<float, SvmPrior, SORTED=true> (+ pf_header_kernel; the batch runs as two halves "
                                               "on two streams, avg_launch_ms = time of one time step of the whole batch)", "achieved": achieved, "peak": hbm,
                     "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                     "traffic_source": "ncu --set full capture of this command (profiles/step_kernel_traffic.json: dram bytes per "
                                       "particle-step x particles per time step), not measured in this run",
                     "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                     "alg_bytes_per_particle_step": ALG_BYTES, "particles_per_launch": prep.B * N_PARTICLES,
                     "avg_launch_ms": step_kernel_ms, "step_kernel_share_of_step": step_kernel_ms * prep.max_T * args.steps / dev_ms},
    }
    if world == 1 and not args.no_cpu_baseline:
        windows = draw_windows(8, seed=99)
        # size the sample from one timed gradient (about 1 s on the box's cores) so it fills ~cpu_seconds
        _, dt1, _ = cpu_rate(y, windows, 1, 1, 0)
        n_grad = int(min(64, max(2, args.cpu_seconds / max(dt1, 1e-3))))
        rate, dt, done = cpu_rate(y, windows, 1, n_grad, 0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": 1, "kind": ref_kind(),
                                "sample": "{0} subsequence gradient(s) of the same workload (N=2^16, T_buf<=60), {3}, "
                                          "{1:.1f} s; host has {2} cores".format(
                                              n_grad, dt, os.cpu_count(),
                                              "the unmodified reference (oracle/_ref: SVMHelper.pf_gradient_estimate)"
                                              if ref_kind() == "reference" else "numpy oracle port incl. np.random.choice")}
    if not args.no_extras:
        f64 = multi.pop("f64")
        line["f64"] = f64
        line["roofline_f64"] = f64["native"]["roofline"]
        line["extra"] = multi
        if world == 1:
            line["extra"].update(extras(sg, y, theta, windows_all, torch))
    print(json.dumps(line), flush=True)
    return 0


def _timed_launches(prep, torch, parallel, world, warm, steps):
    """Device-timed launches of a prepared batch (+ the all-reduce of the gradient sums): (ms per launch, ms per time step)."""
    st = prep.st
    so = prep.base_out - st.dev_out.data_ptr()
    grad_view = st.dev_out[so:so + prep.B * 64].view(torch.float64).view(prep.B, 8)
    gsum = torch.zeros(8, dtype=torch.float64, device=grad_view.device)

    def one(k, ev=None):
        prep.launch(offset=k + 1, step_events=ev)
        torch.sum(grad_view, dim=0, out=gsum)
        if world > 1:
            torch.distributed.all_reduce(gsum)
    for k in range(warm):
        one(k)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(); b.record()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    parallel.barrier()
    e0.record()
    for k in range(steps):
        one(warm + k, ev[k])
    e1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    ms = parallel.allreduce_max(e0.elapsed_time(e1)) / steps
    step_ms = parallel.allreduce_max(float(np.mean([a.elapsed_time(b) for a, b in ev])) / prep.max_T)
    assert bool(torch.all(torch.isfinite(gsum)))
    return ms, step_ms


def f64_block(sg, torch, parallel, items, lo, args, world, hbm, sampler, api_kw):
    """The headline workload in the reference's precision: float64 particle arithmetic and storage (80 B / particle-step),
    device-timed and through the public API; `native` = 53-bit device variates, `variates_f32` = the f32 path's random
    variates widened to f64 (every operation on the particle system still f64; include/sgmpf.h SGM_VARIATES_F32)."""
    out = {}
    for name, variates in (("native", "native"), ("variates_f32", "f32")):
        prep = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", items, N_PARTICLES, dtype="f64", rng="philox",
                                    resample="multinomial_sorted", item_id_base=lo, variates=variates).upload()
        ms, step_ms = _timed_launches(prep, torch, parallel, world, 2, 3)
        total_ps = parallel.allreduce_sum(np.array([float(prep.particle_steps)]))[0]
        achieved = 2 * ALG_BYTES * prep.B * N_PARTICLES / (step_ms * 1e-3) / 1e9
        kw = dict(api_kw, dtype="f64", variates=variates)
        sampler.noisy_gradient(**kw)
        torch.cuda.synchronize(); parallel.barrier()
        t0 = time.perf_counter()
        ps = 0
        for _ in range(2):
            sampler.noisy_gradient(**kw)
            ps += sampler.last_pf_info["particle_steps"]
        torch.cuda.synchronize(); parallel.barrier()
        dt = parallel.allreduce_max(time.perf_counter() - t0)
        out[name] = {"dtype": "f64", "variates": variates, "value": total_ps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                     "e2e": {"value": parallel.allreduce_sum(np.array([float(ps)]))[0] / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / 2},
                     "roofline": {"bound": "hbm", "kernel": "pf_step_kernel<double, SvmPrior, SORTED=true, FM_POY> (+ pf_header_kernel)",
                                  "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                                  "alg_bytes_per_particle_step": 2 * ALG_BYTES, "particles_per_launch": prep.B * N_PARTICLES,
                                  "avg_launch_ms": step_ms}}
    return out


def config5_strong(sg, torch, parallel, theta, world, hbm):
    """BASELINE configs[4] as written (strong scaling): SVM T = 10^6, minibatch 1024 subsequences of S=40 / B=10, N = 2^20
    particles each, the 1024 items split over the ranks, one all-reduce of the gradient sums per gradient."""
    from scipy.signal import lfilter
    from sgmcmc_ssm_b200.sgmcmc_sampler import random_subsequences_packed
    N5, M5, T5 = 1 << 20, 1024, 1000000
    rs = np.random.RandomState(2020)
    x = lfilter([1.0], [1.0, -0.95], np.sqrt(0.5) * rs.normal(size=T5))
    y5 = (np.sqrt(0.5) * np.exp(0.5 * x) * rs.normal(size=T5)).reshape(-1, 1)
    lo, hi = parallel.shard_bounds(M5)
    state = np.random.get_state()
    np.random.seed(555)
    arrays = random_subsequences_packed(y5, SUBSEQ, M5, BUFFER, None, lo, hi)
    np.random.set_state(state)
    pk = sg.PackedItems(theta=theta, prior_mean=0.0, prior_var=10.0, **arrays)
    prep = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", pk, N5, dtype="f32", rng="philox",
                                resample="multinomial_sorted", item_id_base=lo).upload()
    ms, step_ms = _timed_launches(prep, torch, parallel, world, 1, 2)
    total_ps = parallel.allreduce_sum(np.array([float(prep.particle_steps)]))[0]
    achieved = ALG_BYTES * prep.B * N5 / (step_ms * 1e-3) / 1e9
    out = {"workload": "SVM T=1e6, 1024 subsequences x N=2^20 particles, S=40 B=10, split over the ranks (strong scaling)",
           "items_per_gpu": hi - lo, "n_gpus": world, "ms_per_gradient": ms, "particle_steps_per_s": total_ps / (ms * 1e-3),
           "frac": achieved / hbm, "workspace_gb_per_gpu": prep.ws_bytes / 1e9}
    prep.st.workspace = None                    # hand the 43 GB (1 GPU) back before the next block
    del prep
    torch.cuda.empty_cache()
    return out


def config4_chains(sg, torch, parallel, world):
    """BASELINE configs[3]: 64 independent SGLD chains of a Seq SVM sampler on the EUR/USD hourly returns (49 sequences /
    5907 observations, demo/exchange_rate/exchange_rate_full_demo.py:16-42, 96-102), chains sharded over the ranks with no
    collective per iteration, every iteration on the device (device_loop.DeviceChains)."""
    from sgmcmc_ssm_b200.device_loop import DeviceChains
    from sgmcmc_ssm_b200.models.svm import SeqSVMSampler
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from chains_demo import eurus_sequences, chain
    seqs = eurus_sequences()
    n_chains = 64
    lo, hi = parallel.shard_bounds(n_chains)
    out = {"workload": "64 SeqSVMSampler SGLD chains on EUR/USD hourly returns (49 seq / 5907 obs), eps=1e-3, S=16, B=4, "
                       "num_sequences=1, pf=poyiadjis_N; chains sharded over the ranks, device-resident loop", "n_gpus": world}
    state = np.random.get_state()
    for N, iters in ((1000, 2000), (10000, 200)):
        chains = DeviceChains([chain(seqs, 12345 + c) for c in range(lo, hi)], method="SGLD", epsilon=1e-3, chain_id_base=lo,
                              kind="pf", pf="poyiadjis_N", N=N, subsequence_length=16, buffer_length=4, minibatch_size=1,
                              num_sequences=1)
        chains.run(max(16, iters // 10)).synchronize()
        parallel.barrier()
        t0 = time.perf_counter()
        chains.run(iters).synchronize()
        parallel.barrier()
        dt = parallel.allreduce_max(time.perf_counter() - t0)
        params = chains.pull_parameters()
        ok = all(np.isfinite(float(np.ravel(p.A)[0])) for p in params)
        out["N%d" % N] = {"chain_iterations_per_sec": n_chains * iters / dt, "ms_per_iteration_all_chains": 1e3 * dt / iters,
                          "chains_per_gpu": hi - lo, "persistent_kernel": bool(chains.persistent), "finite": bool(ok)}
    np.random.set_state(state)
    return out


def extras(sg, y, theta, windows, torch):
    """Secondary numbers: single-subsequence latency (the reference script's minibatch_size=1) and SGLD iters/sec."""
    out = {}
    w = windows[0]
    it = sg.PFItems().add(y[w["lo"]:w["hi"]], theta, t1=w["start"] - w["lo"], tL=w["start"] - w["lo"] + SUBSEQ,
                          weights=w["weights"], prior_mean=0.0, prior_var=10.0)
    prep = sg.engine.PreparedPF("svm", "prior", "poyiadjis_N", it, N_PARTICLES, dtype="f32", rng="philox",
                                resample="multinomial_sorted").upload()
    go = prep.launch_graph if prep.graph_eligible() else prep.launch     # what run_pf / the samplers do for a small batch
    for k in range(3):
        go(offset=k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(20):
        go(offset=10 + k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out["minibatch1_ms_per_gradient"] = ms
    out["minibatch1_particle_steps_per_sec"] = prep.particle_steps / (ms * 1e-3)
    out["minibatch1_cuda_graph"] = bool(prep.graph_eligible())
    # SGLD iterations / s ("full sample_sgld + project_parameters iterations per second").
    # configs[0]: LGSSM T=1000, N=1000, S=40, B=10, minibatch 1 -- through the public API: sampler.fit(iter_type='SGLD', ...)
    # runs the iterations on the device (device_loop.DeviceChains; persistent kernel), timed with host wall clock
    # incl. set-up, the single enqueue and the parameter read-back; `host_loop` = one Python iteration per step (round 1).
    from sgmcmc_ssm_b200.models.lgssm import LGSSMSampler, LGSSMParameters, generate_lgssm_data
    from sgmcmc_ssm_b200.models.svm import SVMSampler, SVMParameters
    np.random.seed(12345)
    p = LGSSMParameters(A=np.eye(1) * 0.9, C=np.eye(1), LQinv=np.eye(1) * np.sqrt(10.0), LRinv=np.eye(1))
    data = generate_lgssm_data(T=1000, parameters=p)
    fit_kw = dict(epsilon=0.01, subsequence_length=40, buffer_length=10, minibatch_size=1, kind="pf")

    def fit_rate(sampler, iters, N):
        sampler.fit("SGLD", max(8, iters // 20), pf_kwargs=dict(pf="poyiadjis_N", N=N), **fit_kw)      # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sampler.fit("SGLD", iters, pf_kwargs=dict(pf="poyiadjis_N", N=N), **fit_kw)
        dt = time.perf_counter() - t0
        assert all(np.all(np.isfinite(v)) for v in sampler.parameters.var_dict.values())
        return iters / dt
    s = LGSSMSampler(n=1, m=1, observations=data["observations"], parameters=p.copy())
    out["sgld_iters_per_sec_lgssm_T1000_N1000"] = fit_rate(s, 4000, 1000)
    s = LGSSMSampler(n=1, m=1, observations=data["observations"], parameters=p.copy())
    kw = dict(epsilon=0.01, kind="pf", pf="poyiadjis_N", N=1000, subsequence_length=40, buffer_length=10, minibatch_size=1)
    for _ in range(5):
        s.sample_sgld(**kw); s.project_parameters()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        s.sample_sgld(**kw); s.project_parameters()
    torch.cuda.synchronize()
    out["sgld_iters_per_sec_lgssm_T1000_N1000_host_loop"] = 100 / (time.perf_counter() - t0)
    # configs[1]: SVM T=10000, minibatch 1, N = 2^10 (persistent kernel), 2^13, 2^16 (tile kernels, CUDA graph per 8 iterations)
    ps = SVMParameters(A=np.eye(1) * 0.95, LQinv=np.eye(1) * np.sqrt(2.0), LRinv=np.eye(1) * np.sqrt(2.0))
    for N, iters in ((1 << 10, 4000), (1 << 13, 400), (1 << 16, 200)):
        s = SVMSampler(n=1, m=1, observations=y, parameters=ps.copy())
        out["sgld_iters_per_sec_svm_T10000_N%d" % N] = fit_rate(s, iters, N)

    # O(N^2) smoother (configs[1]: N up to 2^16) and PaRIS (configs[2]: GARCH N = 2^14), device-timed
    def timed(model, kern, pf, th, N, B, T, **kw):
        rs = np.random.RandomState(1)
        its = sg.PFItems()
        for _ in range(B):
            its.add(rs.normal(size=T) * 0.7, th, t1=2, tL=T - 2, prior_mean=0.0, prior_var=1.0)
        pp = sg.engine.PreparedPF(model, kern, pf, its, N, dtype="f32", rng="philox", resample="multinomial_sorted", **kw).upload()
        pp.launch(offset=1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for k in range(2):
            pp.launch(offset=2 + k)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / 2 * 1e-3
    sec = timed("svm", "prior", "poyiadjis_N2", theta, N_PARTICLES, 1, 8)
    out["n2_tensor_pair_steps_per_sec_svm_N65536"] = N_PARTICLES * float(N_PARTICLES) * 8 / sec
    sec = timed("svm", "prior", "poyiadjis_N2", theta, N_PARTICLES, 1, 8, n2_mode="fp32_pipe")
    out["n2_fp32pipe_pair_steps_per_sec_svm_N65536"] = N_PARTICLES * float(N_PARTICLES) * 8 / sec
    gth = [0.1, 0.8, 0.05, 0.1 / 0.15, 0.85, 0.8 / 0.85, 1 / 0.3, 1 / 0.09, 0.09]
    sec = timed("garch", "optimal", "paris", gth, 1 << 14, 64, 60)
    out["paris_particle_steps_per_sec_garch_N16384_B64"] = (1 << 14) * 64 * 60 / sec
    return out


def main():
    global N_PARTICLES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--minibatch", type=int, default=512, help="subsequences per GPU per step")
    ap.add_argument("--particles", type=int, default=N_PARTICLES,
                    help="particles per subsequence (default 2^16 = the metric's configuration; 2^20 with --minibatch 128 "
                         "is one GPU's share of BASELINE configs[4])")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    N_PARTICLES = int(args.particles)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
