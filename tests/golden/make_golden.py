"""Generates tests/golden/ksw_golden.npz from the UNMODIFIED reference ksw.c
(oracle/_ref/libksw_ref.so, built by `make -C oracle ref` where /root/reference exists).

The inputs are regenerated from seeds by lamsa_b200.workload (a digest of the
input bytes is stored so generator drift is detected); the file holds only the
reference's outputs: per-task result fields and the concatenated CIGAR words.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from lamsa_b200 import workload  # noqa: E402
import _oracle  # noqa: E402

SETS = {
    "edge": lambda: workload.gen_edge_cases(7),
    "micro": lambda: workload.gen_microbench(3000, seed=11),
    "micro_score": lambda: workload.gen_microbench(1000, seed=12, cigar=False),
}


def main():
    assert _oracle.have_ref(), "build oracle/_ref first: make -C oracle ref"
    out = {}
    for name, gen in SETS.items():
        tasks, keep = gen()
        res, cig, _ = _oracle.ref_run(tasks, 8)
        n = res["n_cigar"].astype(np.int64)
        # store CIGARs in task order
        tot = int(n.sum())
        idx = np.repeat(res["cigar_off"], n) + (np.arange(tot) - np.repeat(np.cumsum(n) - n, n))
        out[f"{name}_cigar"] = cig[idx].astype(np.int32)
        for f in ("score", "qle", "tle", "gtle", "gscore", "max_off", "n_cigar", "m_cigar"):
            out[f"{name}_{f}"] = res[f].astype(np.int32)
        out[f"{name}_inputs_sha1"] = np.array(_oracle.inputs_digest(tasks))
        print(name, len(tasks), "tasks,", tot, "cigar words")
    path = os.path.join(ROOT, "tests", "golden", "ksw_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
