"""Regenerates tests/golden/hash_line_golden.npz: the lines the UNMODIFIED reference (init_hash / hash_hit /
hash_main_line of src/split_mapping.c, through oracle/_ref/liblamsa_ref.so) chains on tests/_hash.gen_cases(240, 4242).
Run in the container that has /root/reference (make -C oracle sdpref first)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import _hash

n, seed = 240, 4242
cases = _hash.gen_cases(n, seed)
lines = [_hash.ref_line(c) for c in cases]
off = np.concatenate(([0], np.cumsum([len(l) for l in lines])))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hash_line_golden.npz"), n=n, seed=seed, off=off,
                    lines=np.concatenate(lines).astype(np.int32), read_bases=sum(len(c["read"]) for c in cases))
print("wrote", off[-1], "line nodes of", n, "cases")
