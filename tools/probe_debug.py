"""Manual GPU triage: runs workloads class by class and prints the first mismatches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import lamsa_b200
from lamsa_b200 import workload
import _oracle

ctx = lamsa_b200.Context(0)
print("int peak:", ctx.int_peak(), flush=True)
for name, gen in [("edge", lambda: workload.gen_edge_cases(7)),
                  ("micro", lambda: workload.gen_microbench(20000, seed=201)),
                  ("small", lambda: workload.gen_microbench(30000, seed=202, qmin=1, qmax=160, wmin=1, wmax=30, max_err=0.3, max_dl=20))]:
    tasks, keep = gen()
    t0 = time.time()
    b = lamsa_b200.Batch(ctx, tasks, keep)
    b.upload(); ms = b.compute(); res, cig = b.download(); st = b.stats(); b.close()
    ores, ocig, secs = _oracle.oracle_run(tasks)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what=name, check_cells=True)
    cells = int(ores["cells"].sum())
    print(f"{name}: n={len(tasks)} cells={cells} kernel_ms={ms:.3f} (fill {st['fill_ms']:.3f} trace {st['trace_ms']:.3f}) "
          f"GCUPS={cells / ms / 1e6:.2f} oracle {secs:.2f}s ({cells / secs / 1e9:.3f} GCUPS, {os.cpu_count()} thr) "
          f"mismatches={len(bad)}", flush=True)
    for x in bad[:12]:
        print("   ", x)
    if bad:
        kinds = {}
        # per class summary of failures
        for f in ("score", "qle", "tle", "n_cigar"):
            d = res[f] != ores[f]
            print(f"    field {f}: {int(d.sum())} tasks differ; by kind: global {int((d & (tasks['kind']==0)).sum())} extend {int((d & (tasks['kind']==1)).sum())}")
