"""Manual GPU triage: wall-clock breakdown of the end-to-end path (host task records in, host results out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lamsa_b200
from lamsa_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
tasks, keep = workload.gen_microbench(n)
ctx = lamsa_b200.Context(0)
for it in range(3):
    t0 = time.perf_counter(); b = lamsa_b200.Batch(ctx, tasks, keep)
    t1 = time.perf_counter(); b.upload(); import torch; 
    t2 = time.perf_counter(); ms = b.compute()
    t3 = time.perf_counter(); res, cig = b.download(copy=False)
    t4 = time.perf_counter(); b.close()
    t5 = time.perf_counter()
    cells = int(res["cells"].sum())
    print(f"iter {it}: create(pack) {t1-t0:.3f}  upload(async) {t2-t1:.3f}  compute {t3-t2:.3f} (kernels {ms/1e3:.3f})  "
          f"download {t4-t3:.3f}  close {t5-t4:.3f}  total {t5-t0:.3f} s -> {cells/(t5-t0)/1e9:.1f} GCUPS e2e; cigar words {len(cig)}", flush=True)

for it in range(3):
    t0 = time.perf_counter(); res, cig = ctx.run(tasks, keep); dt = time.perf_counter() - t0
    print(f"lb2_dp_run (pipelined) iter {it}: {dt:.3f} s -> {int(res['cells'].sum())/dt/1e9:.1f} GCUPS e2e; {ctx.last_run_stats()}", flush=True)

ptasks, pool = workload.pool_tasks(tasks, keep, lamsa_b200.pinned_pool)
for chunk in (100000, 131072, 200000):
    ctx.set_chunk_tasks(chunk)
    for it in range(3):
        t0 = time.perf_counter(); res, cig = ctx.run_pool(ptasks, pool); dt = time.perf_counter() - t0
        print(f"lb2_dp_run_pool chunk {chunk} iter {it}: {dt:.3f} s -> {int(res['cells'].sum())/dt/1e9:.1f} GCUPS e2e; {ctx.last_run_stats()} {ctx.last_run_kernel_ms()}", flush=True)
