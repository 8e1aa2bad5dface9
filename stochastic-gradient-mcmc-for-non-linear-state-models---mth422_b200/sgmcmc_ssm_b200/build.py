"""Build recipe for the CUDA library (nvcc, sm_100a only).  Used by __graft_entry__.build() and by
`python -m sgmcmc_ssm_b200.build`.  The .so is built IN-TREE next to csrc/ so it travels with the repo
snapshot; it is git-ignored."""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(PKG_DIR, "csrc")
# SGM_LIB_PATH: developer override for A/B-testing an experimental build of the same library
LIB_PATH = os.environ.get("SGM_LIB_PATH") or os.path.join(PKG_DIR, "libsgmpf.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]
# translation units: the C-ABI + the two per-dtype instantiations of the launch orchestration (compiled in parallel)
UNITS = ["sgmpf.cu", "sgmpf_f32.cu", "sgmpf_f64.cu", "sgmpf_cl32.cu", "sgmpf_cl64.cu"]
OBJ_DIR = os.path.join(PKG_DIR, "build")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + [
        os.path.join(os.path.dirname(PKG_DIR), "include", "sgmpf.h")]


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if os.environ.get("SGM_LIB_PATH") or (not force and not is_stale()):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = [f for f in os.environ.get("SGM_NVCC_EXTRA", "").split() if f]          # e.g. -DSGM_STEP_CTAS=5 for A/B builds
    os.makedirs(OBJ_DIR, exist_ok=True)
    procs = []
    for unit in UNITS:
        obj = os.path.join(OBJ_DIR, unit.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, unit)]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    objs, log = [], ""
    for cmd, obj, proc in procs:
        out, err = proc.communicate()
        log += out + err
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + out + err)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(log)
    with open(os.path.join(OBJ_DIR, "nvcc_build.log"), "w") as f:
        f.write(log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
