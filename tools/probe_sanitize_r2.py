"""Manual GPU triage: a small pass through the kernels added in round 2, meant to be run under
`compute-sanitizer --tool memcheck` (one tool per gpurun call):  the register-window extension kernel, the
warp-cooperative traceback, a pooled run (byte offsets into the caller's pool), the split-mapping line kernel and
the local Smith-Waterman kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import lamsa_b200
from lamsa_b200 import workload
import _hash, _sw

ctx = lamsa_b200.Context(0)
tasks, keep = workload.gen_microbench(40, seed=1, qmin=1200, qmax=1500, wmin=5, wmax=15, max_err=0.05)     # lean + long traceback
res, cig = ctx.run(tasks, keep)
t2, k2 = workload.gen_microbench(600, seed=2, qmin=1, qmax=300)
pt, pool = workload.pool_tasks(t2, k2, lamsa_b200.pinned_pool)
res2, cig2 = ctx.run_pool(pt, pool)
lines, hits = _hash.gpu_lines(ctx, _hash.gen_cases(12, 5, max_len=900))
sw = [_sw.gpu_align2(c) for c in _sw.gen_cases(12, 5, qmax=120, tmax=300)]
print("ok", int(res["cells"].sum()), int(res2["cells"].sum()), sum(len(l) for l in lines), len(sw))
ctx.close()
