"""Manual GPU triage: small mixed workload for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import lamsa_b200
from lamsa_b200 import workload
import _oracle
ctx = lamsa_b200.Context(0)
for name, gen in [("edge", lambda: workload.gen_edge_cases(7)),
                  ("micro", lambda: workload.gen_microbench(600, seed=301, qmax=500)),
                  ("small", lambda: workload.gen_microbench(1500, seed=302, qmin=1, qmax=120, wmin=1, wmax=30, max_dl=20))]:
    tasks, keep = gen()
    if name == "edge":
        tasks = tasks[(tasks["qlen"] < 1100)]
    res, cig = ctx.run(tasks, keep)
    ores, ocig, _ = _oracle.oracle_run(tasks)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what=name, check_cells=True)
    print(name, len(tasks), "mismatches", len(bad), flush=True)
