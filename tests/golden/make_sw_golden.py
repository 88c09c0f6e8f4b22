"""Regenerates tests/golden/sw_golden.npz: the seven kswr_t fields the UNMODIFIED reference ksw_align2 (src/ksw.c:344-371,
through oracle/_ref/libksw_ref.so) returns on tests/_sw.gen_cases(400, 4711).  Run where /root/reference exists
(make -C oracle ref first)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import _sw

n, seed = 400, 4711
res = np.array([_sw.ref_align2(c) for c in _sw.gen_cases(n, seed)], np.int32)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sw_golden.npz"), n=n, seed=seed, res=res)
print("wrote", n, "results")
