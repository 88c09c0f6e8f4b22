import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, lamsa_b200, _oracle
from lamsa_b200 import workload
ctx = lamsa_b200.Context(0)
for name, gen in [("edge", lambda: workload.gen_edge_cases(7)), ("micro", lambda: workload.gen_microbench(4000, seed=301, qmax=500))]:
    tasks, keep = gen()
    res, cig = ctx.run(tasks, keep)
    ores, ocig, _ = _oracle.oracle_run(tasks)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what=name, check_cells=True)
    print(name, len(tasks), "mismatches", len(bad))
    for b in bad[:12]: print("   ", b)
    # which tasks differ in score
    d = np.nonzero((res["score"] != ores["score"]) | (res["cells"] != ores["cells"]) | (res["n_cigar"] != ores["n_cigar"]))[0]
    for i in d[:15]:
        t = tasks[i]; print("    task", i, "kind", t["kind"], "qlen", t["qlen"], "tlen", t["tlen"], "w", t["w"], "h0", t["h0"], "flags", t["flags"], "gpu", res[i]["score"], res[i]["cells"], res[i]["n_cigar"], "orc", ores[i]["score"], ores[i]["cells"], ores[i]["n_cigar"])
