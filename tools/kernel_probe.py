#!/usr/bin/env python
"""Kernel-only timing of the C2 microbenchmark batch under the current LB2_* routing knobs
(environment), for quick A/B runs: prints cells, ms per step and GCUPS (actual cells)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lamsa_b200
from lamsa_b200 import workload

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
tasks, keep = workload.gen_microbench(n, seed=20260101)
ctx = lamsa_b200.Context(0)
b = lamsa_b200.Batch(ctx, tasks, keep)
b.upload()
for _ in range(2):
    b.compute()
ms = min(b.compute() for _ in range(4))
st = b.stats()
res, cig = b.download()
cells = int(res["cells"].sum())
b.close()
import time
e2e = []
for _ in range(3):
    t0 = time.perf_counter(); r2, c2 = ctx.run(tasks, keep); e2e.append(time.perf_counter() - t0); del r2, c2
knobs = {k: v for k, v in os.environ.items() if k.startswith("LB2_")}
print(json.dumps({"tasks": n, "knobs": knobs, "ms": round(ms, 3), "fill_ms": round(st["fill_ms"], 3), "trace_ms": round(st["trace_ms"], 3),
                  "gcups": round(cells / ms / 1e6, 1), "e2e_s": round(min(e2e[1:]), 4), "e2e_gcups": round(cells / min(e2e[1:]) / 1e9, 1), "checksum": int(res["score"].astype("int64").sum()), "cigar_words": int(res["n_cigar"].sum())}))
