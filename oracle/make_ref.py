#!/usr/bin/env python3
"""Vendor the UNMODIFIED reference package next to the oracle:  /root/reference/sgmcmc_ssm  ->  oracle/_ref/sgmcmc_ssm.

The reference is pure Python (nothing to compile), so "building" it is a copy of its `.py` files where they lie; the
copy is git-ignored (reference sources never enter this repo's history) but NOT gpurun-ignored, so it travels to the GPU
box like a built .so and `bench.py --impl reference` / `cpu_baseline` can time the reference ITSELF
(`cpu_baseline.kind = "reference"`) instead of the oracle port.  Run by `__graft_entry__.build()` whenever
/root/reference is present (build container); on the GPU box the prebuilt copy is used as is."""
import os
import shutil
import sys

SRC = "/root/reference/sgmcmc_ssm"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "sgmcmc_ssm")


def make(force=False):
    if not os.path.isdir(SRC):
        return os.path.isdir(DST)
    if os.path.isdir(DST) and not force:
        return True
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    ok = make(force="--force" in sys.argv)
    print("oracle/_ref:", "present" if ok else "unavailable (no /root/reference)")
