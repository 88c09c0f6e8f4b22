"""GPU suite, BASELINE configs[1] at FULL size (1 M tasks, ~8e10 cells): too large to replay on the
oracle in seconds, so it is checked through size-independent properties
  * every CIGAR consumes exactly the aligned query / target prefix (global: the whole pair),
  * the global fill evaluates exactly the static-band cell count of src/ksw.c:577-578,
  * a pipelined lb2_dp_run (chunked) and one resident batch agree word for word (digest),
plus a bit-exact oracle comparison on a random 40 k-task sample of the same stream."""
import numpy as np
import pytest

import _oracle
import lamsa_b200
from lamsa_b200 import workload

pytestmark = pytest.mark.gpu
N = 1_000_000


def cigar_sums(res, cig):
    n = res["n_cigar"].astype(np.int64)
    tot = int(n.sum())
    idx = np.repeat(res["cigar_off"], n) + (np.arange(tot) - np.repeat(np.cumsum(n) - n, n))
    w = cig[idx].astype(np.int64)
    op, ln = w & 15, w >> 4
    owner = np.repeat(np.arange(len(res)), n)
    q = np.bincount(owner, weights=ln * ((op == 0) | (op == 1)), minlength=len(res)).astype(np.int64)
    t = np.bincount(owner, weights=ln * ((op == 0) | (op == 2)), minlength=len(res)).astype(np.int64)
    bad_op = np.bincount(owner, weights=(op > 2), minlength=len(res)) > 0
    return q, t, bad_op


def test_full_size_properties(ctx):
    tasks, keep = workload.gen_microbench(N)
    b = lamsa_b200.Batch(ctx, tasks, keep)
    b.upload(); b.compute()
    res, cig = b.download()
    b.close()
    glob = tasks["kind"] == 0
    q, t, bad_op = cigar_sums(res, cig)
    assert not bad_op.any()
    assert (q[glob] == tasks["qlen"][glob]).all() and (t[glob] == tasks["tlen"][glob]).all()
    ext = ~glob
    assert (q[ext] == res["qle"][ext]).all() and (t[ext] == res["tle"][ext]).all()
    assert (res["qle"][ext] <= tasks["qlen"][ext]).all() and (res["tle"][ext] <= tasks["tlen"][ext]).all()
    assert (res["score"][ext] >= tasks["h0"][ext]).all()
    # static band cell count of the global fill
    ql, tl = tasks["qlen"][glob].astype(np.int64), tasks["tlen"][glob].astype(np.int64)
    w = np.maximum(tasks["w"][glob], np.abs(ql - tl) + 3).astype(np.int64)
    i = np.arange(1051, dtype=np.int64)[None, :]
    width = np.minimum(ql[:, None], i + w[:, None] + 1) - np.maximum(0, i - w[:, None])
    width = np.where(i < tl[:, None], np.maximum(width, 0), 0)
    assert (res["cells"][glob] == width.sum(axis=1)).all()
    # pipelined one-shot run == resident batch
    res2, cig2 = ctx.run(tasks, keep)
    assert ctx.last_run_stats()["launches"] > 40
    for f in ("score", "qle", "tle", "n_cigar", "cells"):
        assert (res[f] == res2[f]).all(), f
    assert _oracle.cigar_digest(res, cig) == _oracle.cigar_digest(res2, cig2)
    # bit-exact sample against the oracle
    pick = np.sort(np.random.default_rng(1).choice(N, size=40_000, replace=False))
    sub = tasks[pick]
    ores, ocig, _ = _oracle.oracle_run(sub)
    sres = res[pick].copy()
    bad = _oracle.compare(sub, sres, cig, ores, ocig, what="full-size sample", check_cells=True)
    assert not bad, "\n".join(bad)
