#!/usr/bin/env python
"""Latency of one lb2_dp_run call (pack + H2D + kernels + D2H) as a function of batch size and of the
longest task in the batch: the quantities that bound a round of the batch producer (fiber_sched.cu)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lamsa_b200
from lamsa_b200 import workload

ctx = lamsa_b200.Context(0)
for n, qmax in ((1, 60), (64, 60), (800, 60), (800, 300), (2000, 60), (2000, 300), (8000, 60), (8000, 300)):
    tasks, keep = workload.gen_microbench(n, seed=3, qmin=20, qmax=qmax, wmax=12) if "wmax" in workload.gen_microbench.__code__.co_varnames \
        else workload.gen_microbench(n, seed=3, qmax=qmax)
    ctx.run(tasks, keep)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); ctx.run(tasks, keep); ts.append(time.perf_counter() - t0)
    print(json.dumps({"tasks": n, "qmax": qmax, "median_us": round(float(np.median(ts)) * 1e6, 1), "min_us": round(min(ts) * 1e6, 1)}), flush=True)
# one long task among short ones
for long_len in (1000, 5000):
    tasks, keep = workload.gen_microbench(800, seed=5, qmax=60)
    big, keep2 = workload.gen_microbench(1, seed=9, qmin=long_len, qmax=long_len) if "qmin" in workload.gen_microbench.__code__.co_varnames else (None, None)
    if big is None:
        break
    allt = np.concatenate([tasks, big])
    ctx.run(allt, keep + keep2)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); ctx.run(allt, keep + keep2); ts.append(time.perf_counter() - t0)
    print(json.dumps({"tasks": 801, "one_long_task": long_len, "kind": int(big["kind"][0]), "w": int(big["w"][0]), "median_us": round(float(np.median(ts)) * 1e6, 1)}), flush=True)
ctx.close()
