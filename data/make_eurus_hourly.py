#!/usr/bin/env python3
"""Extract the hourly EUR/USD log-returns of the reference's data file (data/EURUS_processed.npz, the input of
demo/exchange_rate/exchange_rate_full_demo.py:16-42 and of BASELINE configs[3]) into a small fixture that travels with
the repo: `hourly_log_returns` (5908,) float64 and `hourly_date` (5908,) datetime64[h] -- 95 KB instead of the 2.8 MB
file with the minute / daily series.  It is data, not code.  Build-container only."""
import os
import numpy as np

SRC = "/root/reference/data/EURUS_processed.npz"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "EURUS_hourly.npz")
z = np.load(SRC)
np.savez_compressed(OUT, hourly_log_returns=z["hourly_log_returns"], hourly_date=z["hourly_date"])
print("wrote", OUT, os.path.getsize(OUT), "bytes")
