"""Regenerates tests/golden/hash_split_golden.npz: CIGAR and return value of the UNMODIFIED reference hash_split_map
(src/split_mapping.c:634-825: k-mer index, line, DP stitching; through oracle/_ref/liblamsa_ref.so) on
tests/_hash.gen_cases(160, 777, max_len=1800) with the presets of tests/_hash.make_ap.  Run in the container that has
/root/reference (make -C oracle sdpref first)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import _hash

n, seed, max_len = 160, 777, 1800
cases = _hash.gen_cases(n, seed, max_len=max_len)
outs = [_hash.ref_split_map(c) for c in cases]
off = np.concatenate(([0], np.cumsum([len(o[0]) for o in outs])))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hash_split_golden.npz"), n=n, seed=seed, max_len=max_len, off=off,
                    cigars=np.concatenate([o[0] for o in outs]).astype(np.int32), res=np.array([o[1] for o in outs], np.int32))
print("wrote", off[-1], "CIGAR words of", n, "cases; split flags", np.bincount(np.array([o[1] for o in outs]), minlength=4).tolist())
