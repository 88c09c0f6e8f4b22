"""Manual GPU triage: latency of small batches through lb2_dp_run."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lamsa_b200
from lamsa_b200 import workload
ctx = lamsa_b200.Context(0)
for n in (1, 16, 128, 1024):
    tasks, keep = workload.gen_microbench(n, seed=9, qmin=20, qmax=150, wmin=5, wmax=20, max_dl=10)
    for rep in range(3):
        ctx.run(tasks, keep)
    t0 = time.perf_counter()
    R = 200
    for rep in range(R):
        b = lamsa_b200.Batch(ctx, tasks, keep); b.upload(); ms = b.compute(); r = b.download(); b.close()
    dt = (time.perf_counter() - t0) / R
    t0 = time.perf_counter()
    for rep in range(R):
        b = lamsa_b200.Batch(ctx, tasks, keep); b.close()
    dc = (time.perf_counter() - t0) / R
    print(f"n={n}: {dt*1e6:.0f} us per batch (create+destroy alone {dc*1e6:.0f} us, kernels {ms*1e3:.0f} us, launches {b.stats() if False else ''})")
