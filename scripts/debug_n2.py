import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import sgmcmc_ssm_b200 as sg
import test_gpu_n2 as T
for model, B, N in (("garch", 40, 1300), ("svm", 40, 1300), ("garch", 1, 3000)):
    kern = "prior" if model == "svm" else "optimal"
    it = T._items(model, B, 9, 3)
    out = {}
    for tag, dtype, mode in (("f64", "f64", "auto"), ("f32", "f32", "fp32_pipe"), ("tc", "f32", "tensor")):
        r = sg.run_pf(model, kern, "poyiadjis_N2", it, N, dtype=dtype, rng="philox", seed=11, offset=5,
                      resample="multinomial_sorted", n2_mode=mode, want=("stats", "x"))
        out[tag] = (r.grad.copy(), r.tensor("stats").double().cpu().numpy(), r.tensor("x").double().cpu().numpy())
    d = np.abs(out["tc"][1] - out["f32"][1])
    scale = np.abs(out["f32"][1]) + np.mean(np.abs(out["f32"][1]), axis=1, keepdims=True)
    ratio = d / scale
    idx = np.unravel_index(np.argmax(ratio), ratio.shape)
    print(model, B, N, "max ratio", ratio.max(), "at", idx, "tc", out["tc"][1][idx[0], idx[1]], "f32", out["f32"][1][idx[0], idx[1]], "x", out["f32"][2][idx[0], idx[1]])
    print("  quantiles of ratio", np.quantile(ratio, [0.5, 0.99, 0.9999]))
    print("  grad diff", np.abs(out["tc"][0] - out["f32"][0]).max(), "grad", out["f32"][0][0])
