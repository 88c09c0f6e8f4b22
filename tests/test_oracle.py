"""CPU suite: the oracle (oracle/dp_oracle.c) is pinned against
  (1) the committed golden outputs of the unmodified reference (tests/golden/ksw_golden.npz),
  (2) the reference itself (oracle/_ref) where it was built (this container; not the GPU box).
"""
import ctypes as C
import os

import numpy as np
import pytest

import _oracle
from lamsa_b200 import workload
from lamsa_b200._lib import RESULT_DTYPE

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ksw_golden.npz")

SETS = {
    "edge": lambda: workload.gen_edge_cases(7),
    "micro": lambda: workload.gen_microbench(3000, seed=11),
    "micro_score": lambda: workload.gen_microbench(1000, seed=12, cigar=False),
}


def golden_results(name):
    g = np.load(GOLD)
    n = len(g[f"{name}_score"])
    res = np.zeros(n, dtype=RESULT_DTYPE)
    for f in ("score", "qle", "tle", "gtle", "gscore", "max_off", "n_cigar", "m_cigar"):
        res[f] = g[f"{name}_{f}"]
    nc = res["n_cigar"].astype(np.int64)
    res["cigar_off"] = np.cumsum(nc) - nc
    return res, g[f"{name}_cigar"], str(g[f"{name}_inputs_sha1"])


@pytest.mark.parametrize("name", list(SETS))
def test_oracle_matches_golden(name):
    tasks, keep = SETS[name]()
    gres, gcig, sha = golden_results(name)
    assert _oracle.inputs_digest(tasks) == sha, "workload generator drifted from the golden inputs"
    res, cig, _ = _oracle.oracle_run(tasks, 4)
    bad = _oracle.compare(tasks, res, cig, gres, gcig, what=f"oracle-vs-golden[{name}]")
    assert not bad, "\n".join(bad)


@pytest.mark.skipif(not _oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [101, 102])
def test_oracle_matches_reference_random(seed):
    tasks, keep = workload.gen_microbench(4000, seed=seed, qmax=600)
    a = _oracle.oracle_run(tasks, 8)
    b = _oracle.ref_run(tasks, 8)
    bad = _oracle.compare(tasks, a[0], a[1], b[0], b[1], what="oracle-vs-ref")
    assert not bad, "\n".join(bad)


@pytest.mark.skipif(not _oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_small_and_wide():
    # short tasks like LAMSA's median (17-150 bp, w 7-24) and PacBio-style penalties
    tasks, keep = workload.gen_microbench(6000, seed=103, qmin=1, qmax=160, wmin=1, wmax=30, max_err=0.3, max_dl=20)
    a = _oracle.oracle_run(tasks, 8)
    b = _oracle.ref_run(tasks, 8)
    bad = _oracle.compare(tasks, a[0], a[1], b[0], b[1], what="oracle-vs-ref-small")
    assert not bad, "\n".join(bad)


def test_cell_count_is_static_band_for_global():
    tasks, keep = workload.gen_microbench(500, seed=5, qmax=300)
    res, cig, _ = _oracle.oracle_run(tasks, 2)
    g = tasks["kind"] == 0
    q, t = tasks["qlen"][g].astype(np.int64), tasks["tlen"][g].astype(np.int64)
    w = np.maximum(tasks["w"][g], np.abs(q - t) + 3).astype(np.int64)
    exp = np.array([sum(min(qq, i + ww + 1) - max(0, i - ww) for i in range(tt)) for qq, tt, ww in zip(q, t, w)])
    assert (res["cells"][g] == exp).all()
