"""GPU: the register-window extension kernel (dp_fill_lean.cuh; band <= 15, normally only for tasks of 192+ rows) forced
onto EVERY extension with a narrow band (LB2_LEAN_ROWS=1, read once per process: hence the child process) and compared
with the oracle word for word: random streams with bands 1..15, all edge cases, long tasks, score-only tasks."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import _oracle, lamsa_b200
from lamsa_b200 import workload
ctx = lamsa_b200.Context(0)
sets = []
sets.append(workload.gen_microbench(30000, seed=301, qmin=1, qmax=400, wmin=1, wmax=15, max_err=0.3, max_dl=20))
sets.append(workload.gen_microbench(4000, seed=302, qmin=800, qmax=3000, wmin=1, wmax=15, max_err=0.1, max_dl=10))
sets.append(workload.gen_microbench(3000, seed=303, qmin=1, qmax=600, wmin=1, wmax=15, cigar=False))
sets.append(workload.gen_edge_cases(seed=5))
total = 0
for tasks, keep in sets:
    tasks = tasks.copy()
    narrow = (tasks["kind"] == 1)
    tasks["w"][narrow] = np.minimum(tasks["w"][narrow], 15)
    res, cig = ctx.run(tasks, keep)
    ores, ocig, _ = _oracle.oracle_run(tasks, 4)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what="lean", check_cells=True)
    assert not bad, "\n".join(bad[:10])
    total += int(narrow.sum())
print("lean ok", total)
'''


def test_gpu_lean_kernel_on_every_narrow_extension():
    env = dict(os.environ, LB2_LEAN_ROWS="1")
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0 and "lean ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
