"""Manual GPU triage: per-launch-class timing of one compute step (LB2_CLASS_TIMING=1)."""
import os, sys
os.environ["LB2_CLASS_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import lamsa_b200
from lamsa_b200 import workload
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
tasks, keep = workload.gen_microbench(n)
ctx = lamsa_b200.Context(0)
b = lamsa_b200.Batch(ctx, tasks, keep)
b.upload(); b.compute()
print("---- second step", file=sys.stderr)
ms = b.compute()
res, cig = b.download()
ext = tasks["kind"] == 1
print(f"total {ms:.2f} ms; cells global {int(res['cells'][~ext].sum()):.3e} extend {int(res['cells'][ext].sum()):.3e}; "
      f"static extend {int((tasks['tlen'][ext].astype(np.int64) * np.minimum(tasks['qlen'][ext], 2 * tasks['w'][ext] + 1)).sum()):.3e}", file=sys.stderr)
rows = res["cells"][ext] / np.maximum(tasks["tlen"][ext], 1)
print("extend mean live width per (all) row:", rows.mean(), file=sys.stderr)
