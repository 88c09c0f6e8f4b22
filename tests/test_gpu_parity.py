"""GPU suite (B200): the CUDA path, reached through the C ABI, against
  * the committed golden outputs of the unmodified reference ksw.c,
  * the oracle on seeded random task streams (bit-exact: scores, end points,
    every CIGAR word, evaluated-cell counts),
  * the reference-named drop-in entry points.
Nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest

import _oracle
import lamsa_b200
from lamsa_b200 import workload
from test_oracle import SETS, golden_results

pytestmark = pytest.mark.gpu


def run_gpu(ctx, tasks, keep):
    return ctx.run(tasks, keep)


@pytest.mark.parametrize("name", list(SETS))
def test_gpu_matches_reference_golden(ctx, name):
    tasks, keep = SETS[name]()
    gres, gcig, sha = golden_results(name)
    assert _oracle.inputs_digest(tasks) == sha
    res, cig = run_gpu(ctx, tasks, keep)
    bad = _oracle.compare(tasks, res, cig, gres, gcig, what=f"gpu-vs-golden[{name}]")
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("seed,n,kw", [
    (201, 20000, dict()),
    (202, 30000, dict(qmin=1, qmax=160, wmin=1, wmax=30, max_err=0.3, max_dl=20)),
    (203, 4000, dict(qmin=600, qmax=1000, wmin=150, wmax=200)),
    (204, 6000, dict(cigar=False)),
])
def test_gpu_matches_oracle_random(ctx, seed, n, kw):
    tasks, keep = workload.gen_microbench(n, seed=seed, **kw)
    res, cig = run_gpu(ctx, tasks, keep)
    ores, ocig, _ = _oracle.oracle_run(tasks)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what="gpu-vs-oracle", check_cells=True)
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("pinned", [True, False])
def test_gpu_pooled_run_matches_oracle(ctx, pinned):
    """lb2_dp_run_pool: sequences uploaded as they lie in one host pool (odd alignments, a short chunk size so that
    several chunks with their own pool ranges are pipelined) and re-laid out on the device."""
    tasks, keep = workload.gen_microbench(9000, seed=206, qmin=1, qmax=300)
    edge, ekeep = workload.gen_edge_cases(seed=9)
    tasks = np.concatenate((tasks, edge))
    ptasks, pool = workload.pool_tasks(tasks, alloc=lamsa_b200.pinned_pool if pinned else None)
    c2 = lamsa_b200.Context(0)
    c2.set_chunk_tasks(2000)
    res, cig = c2.run_pool(ptasks, pool)
    st = c2.last_run_stats()
    c2.close()
    assert st["h2d_bytes"] < 1.2 * pool.nbytes + 200 * len(tasks)
    ores, ocig, _ = _oracle.oracle_run(tasks)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what="gpu-pooled-vs-oracle", check_cells=True)
    assert not bad, "\n".join(bad)
    with pytest.raises(RuntimeError, match="outside the pool"):
        ctx.run_pool(tasks[:50], pool)          # the original pointers are not inside the pool


def test_gpu_multi_wave_equals_single_wave(ctx):
    tasks, keep = workload.gen_microbench(3000, seed=205, qmax=500)
    a = run_gpu(ctx, tasks, keep)
    small = lamsa_b200.Context(0)
    small.set_scratch_limit(8 << 20)          # forces many waves through the same scratch
    b = small.run(tasks, keep)
    small.close()
    bad = _oracle.compare(tasks, a[0], a[1], b[0], b[1], what="wave")
    assert not bad, "\n".join(bad)


def test_dropin_entry_points(ctx):
    rng = np.random.default_rng(5)
    mat = lamsa_b200.default_matrix(1, 3)
    AP = lamsa_b200.AlnPara()
    AP.ins_ext_o = AP.del_ext_o = 5
    AP.ins_ext_e = AP.del_ext_e = 2
    AP.ins_gapo = AP.del_gapo = 5
    AP.ins_gape = AP.del_gape = 2
    AP.end_bonus, AP.zdrop, AP.band_w, AP.split_len = 5, 100, 10, 100
    AP.id_rate, AP.aln_mode = 0.04, 0
    orc = C.CDLL(_oracle.ensure_oracle())
    for it in range(40):
        ql = int(rng.integers(0, 300))
        q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        t = q.copy()
        if ql > 10:
            k = int(rng.integers(1, ql // 2))
            t = np.concatenate((q[:k], rng.integers(0, 4, size=int(rng.integers(0, 8)), dtype=np.uint8), q[k + int(rng.integers(0, 5)):]))
        tl = len(t)
        tasks = np.zeros(2, dtype=lamsa_b200.TASK_DTYPE)
        qb, tb = np.concatenate((q, np.zeros(8, np.uint8))), np.concatenate((t, np.zeros(8, np.uint8)))
        for r, kind in zip(tasks, (0, 1)):
            r["kind"], r["flags"], r["qlen"], r["tlen"] = kind, 1, ql, tl
            r["query"], r["target"], r["w"], r["h0"] = qb.ctypes.data, tb.ctypes.data, 10, 19
            r["o_del"], r["e_del"], r["o_ins"], r["e_ins"] = 5, 2, 5, 2
            r["end_bonus"], r["zdrop"], r["m"], r["mat"] = 5, 100, 5, mat.ctypes.data
        ores, ocig, _ = _oracle.oracle_run(tasks, 1)
        s, cg = lamsa_b200.ksw_global2(ql, qb, tl, tb, 5, mat, 5, 2, 5, 2, 10)
        assert s == ores["score"][0]
        assert cg == list(ocig[ores["cigar_off"][0]: ores["cigar_off"][0] + ores["n_cigar"][0]])
        s2, _ = lamsa_b200.ksw_global(ql, qb, tl, tb, 5, mat, 5, 2, 10, want_cigar=False)
        assert s2 == s
        mx, qle, tle, cg, cap = lamsa_b200.ksw_extend_core(ql, qb, tl, tb, 5, mat, 10, 19, AP)
        assert (mx, qle, tle, cap) == (ores["score"][1], ores["qle"][1], ores["tle"][1], ores["m_cigar"][1])
        assert cg == list(ocig[ores["cigar_off"][1]: ores["cigar_off"][1] + ores["n_cigar"][1]])
        rc, qle2, tle2, cg2, _ = lamsa_b200.ksw_extend_c(ql, qb, tl, tb, 5, mat, 10, 19, AP)
        assert (qle2, tle2, cg2) == (qle, tle, cg)
        assert rc == (0 if qle == ql else 1 if tle == tl else 2)


def test_gpu_long_tasks(ctx):
    """Tasks of thousands of rows: one task per warp in the fill (dp_pack.h LB2_LONG_ROWS) and the warp-cooperative
    traceback (dp_trace.cuh trace_long_kernel), with indel runs that push the path out of a staged window, packed
    and int32 kernels (match score 2 keeps a task out of the int16 domain), both kinds; a 20 kbp extension with the
    SV band of SURVEY appendix C (w 2458)."""
    rng = np.random.default_rng(77)
    mats = {1: lamsa_b200.default_matrix(1, 3), 2: lamsa_b200.default_matrix(2, 4)}
    keep, rows = list(mats.values()), []

    def add(kind, q, t, w, h0=0, zdrop=0, pen=(5, 2, 5, 2), end_bonus=5, match=1):
        qb = np.concatenate((q, np.zeros(8, np.uint8))).astype(np.uint8)
        tb = np.concatenate((t, np.zeros(8, np.uint8))).astype(np.uint8)
        keep.extend([qb, tb])
        r = np.zeros(1, dtype=lamsa_b200.TASK_DTYPE)
        r["kind"], r["flags"], r["qlen"], r["tlen"] = kind, 1, len(q), len(t)
        r["query"], r["target"], r["w"], r["h0"] = qb.ctypes.data, tb.ctypes.data, w, h0
        r["o_del"], r["e_del"], r["o_ins"], r["e_ins"] = pen
        r["end_bonus"], r["zdrop"], r["m"], r["mat"] = end_bonus, zdrop, 5, mats[match].ctypes.data
        rows.append(r)

    def indels(q, n, maxlen):
        """copy of q with n indel runs of up to maxlen bases and 3 % substitutions"""
        out, at = [], 0
        for cut in sorted(rng.integers(1, len(q) - 1, size=n)):
            out.append(q[at:cut])
            L = int(rng.integers(1, maxlen + 1))
            if rng.random() < 0.5:
                out.append(rng.integers(0, 4, size=L, dtype=np.uint8)); at = cut
            else:
                at = min(len(q), cut + L)
        out.append(q[at:])
        t = np.concatenate(out).astype(np.uint8)
        sub = rng.random(len(t)) < 0.03
        t[sub] = (t[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
        return t

    for ql in (1600, 3000, 4700, 7000):
        q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        for maxlen, w in ((3, 10), (40, 60), (90, 120)):
            t = indels(q, 12, maxlen)
            for match in (1, 2):
                add(1, q, t, w, h0=100, zdrop=100, match=match)
                add(1, q, t, w, h0=19, zdrop=0, pen=(2, 1, 2, 1), end_bonus=0, match=match)
                add(0, q, t, w, match=match)
                add(0, q, t, w, pen=(1, 1, 1, 1), match=match)
    q = rng.integers(0, 4, size=19982, dtype=np.uint8)
    add(1, q, indels(q, 30, 60), 2458, h0=190, zdrop=100)          # the longest task seen, its band
    add(1, q[:9000], indels(q[:9000], 6, 200), 2458, h0=190, zdrop=0)
    tasks = np.concatenate(rows)
    res, cig = ctx.run(tasks, keep)
    ores, ocig, _ = _oracle.oracle_run(tasks, 4)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what="long", check_cells=True)
    assert not bad, "\n".join(bad)


def test_gpu_huge_windows(ctx):
    """Bands of ~10^4 columns (`-V 10000` gaps): windows beyond shared memory run the
    global-memory-window variant; a 32768-slot packed window still fits shared memory."""
    rng = np.random.default_rng(31)
    mat = lamsa_b200.default_matrix(1, 3)
    keep, rows = [mat], []

    def add(kind, q, t, w, h0=0, zdrop=0, pen=(5, 2, 5, 2), end_bonus=5):
        qb = np.concatenate((q, np.zeros(8, np.uint8))).astype(np.uint8)
        tb = np.concatenate((t, np.zeros(8, np.uint8))).astype(np.uint8)
        keep.extend([qb, tb])
        r = np.zeros(1, dtype=lamsa_b200.TASK_DTYPE)
        r["kind"], r["flags"], r["qlen"], r["tlen"] = kind, 1, len(q), len(t)
        r["query"], r["target"], r["w"], r["h0"] = qb.ctypes.data, tb.ctypes.data, w, h0
        r["o_del"], r["e_del"], r["o_ins"], r["e_ins"] = pen
        r["end_bonus"], r["zdrop"], r["m"], r["mat"] = end_bonus, zdrop, 5, mat.ctypes.data
        rows.append(r)

    q = rng.integers(0, 4, size=20000, dtype=np.uint8)
    t_del = np.concatenate((q[:4000], q[13000:]))            # 9 kbp deletion in the target
    add(0, q, t_del, 10)                                     # global: w -> |dl|+3 = 9003, int32, global window
    add(1, q, t_del, 12000, h0=100)                          # extension, band clamp ~1e4
    add(1, q, q.copy(), 12000, h0=100, zdrop=100)            # long perfect extension (h0+qlen > int16 budget)
    q2 = rng.integers(0, 4, size=9000, dtype=np.uint8)
    t2 = np.concatenate((q2[:3000], rng.integers(0, 4, size=6000, dtype=np.uint8), q2[3000:]))
    add(0, q2, t2, 50, pen=(1, 1, 1, 1))                     # global with a 6 kbp insertion in the target: packed, S=32768
    tasks = np.concatenate(rows)
    res, cig = ctx.run(tasks, keep)
    ores, ocig, _ = _oracle.oracle_run(tasks, 4)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what="huge", check_cells=True)
    assert not bad, "\n".join(bad)


def test_gpu_reference_windows_from_resident_pac(ctx):
    """f1: targets named as (pac coordinate, length) windows of the resident 2-bit reference
    (bntseq .pac layout, reference src/bntseq.c:242) give the same results as unpacked bytes;
    LB2_FLAG_TARGET_REV reads the window back to front (ksw_extend_r, src/ksw.c:829)."""
    rng = np.random.default_rng(41)
    L = 300_000
    ref = rng.integers(0, 4, size=L, dtype=np.uint8)
    pad = np.concatenate((ref, np.zeros((-L) % 4 + 4, np.uint8)))
    q4 = pad[: (len(pad) // 4) * 4].reshape(-1, 4)
    pac = (q4[:, 0] << 6 | q4[:, 1] << 4 | q4[:, 2] << 2 | q4[:, 3]).astype(np.uint8)      # MSB-first
    c2 = lamsa_b200.Context(0)
    c2.set_reference(pac, L)
    base, keep = workload.gen_microbench(4000, seed=42, qmax=400)
    n = len(base)
    tl = base["tlen"].astype(np.int64)
    coor = rng.integers(0, L - 600, size=n).astype(np.int64)
    rev = (np.arange(n) % 3 == 0) & (base["kind"] == 1)
    # explicit-bytes twin of every window (what pac2fa_core + the reversal would hand to ksw)
    toff = np.concatenate(([0], np.cumsum(tl)[:-1]))
    tbytes = np.empty(int(tl.sum()) + 64, dtype=np.uint8)
    for i in range(n):
        wseg = ref[coor[i]: coor[i] + tl[i]]
        tbytes[toff[i]: toff[i] + tl[i]] = wseg[::-1] if rev[i] else wseg
    explicit = base.copy()
    explicit["target"] = tbytes.ctypes.data + toff.astype(np.uint64)
    bypac = base.copy()
    bypac["target"] = 0
    bypac["target_pac"] = coor
    bypac["flags"] = base["flags"] | lamsa_b200.FLAG_TARGET_PAC | np.where(rev, lamsa_b200.FLAG_TARGET_REV, 0)
    ores, ocig, _ = _oracle.oracle_run(explicit)
    res, cig = c2.run(bypac, keep)
    c2.close()
    bad = _oracle.compare(explicit, res, cig, ores, ocig, what="pac-window", check_cells=True)
    assert not bad, "\n".join(bad)
    with pytest.raises(RuntimeError):          # windows need a resident reference
        ctx.run(bypac[:4], keep)


def test_gpu_pipelined_run_equals_oracle():
    """lb2_dp_run cuts large batches into chunks (pack / H2D / kernels overlapped): results and
    CIGAR offsets must be stitched back in task order."""
    c2 = lamsa_b200.Context(0)
    c2.set_chunk_tasks(1500)
    tasks, keep = workload.gen_microbench(20000, seed=206, qmax=300)
    res, cig = c2.run(tasks, keep)
    assert c2.last_run_stats()["launches"] > 30          # really ran as many chunks
    c2.close()
    ores, ocig, _ = _oracle.oracle_run(tasks)
    bad = _oracle.compare(tasks, res, cig, ores, ocig, what="pipelined", check_cells=True)
    assert not bad, "\n".join(bad)


def test_gpu_async_compute_matches_blocking(ctx):
    """lb2_batch_compute_async / _done / _wait (the batch producer's path) = lb2_batch_compute"""
    import time
    tasks, keep = workload.gen_microbench(3000, seed=909, qmax=400)
    a = lamsa_b200.Batch(ctx, tasks, keep); a.upload(); a.compute(); ra, ca = a.download(); a.close()
    b = lamsa_b200.Batch(ctx, tasks, keep); b.upload(); b.compute_async()
    t0 = time.time()
    while not b.done():
        assert time.time() - t0 < 60
        time.sleep(0.0005)
    ms = b.wait()
    rb, cb = b.download(); b.close()
    assert ms > 0
    # CIGARs are reserved in the dense pool in completion order, so offsets differ between runs: compare per task
    bad = _oracle.compare(tasks, ra, ca, rb, cb, what="async-vs-blocking", check_cells=True)
    assert not bad, "\n".join(bad)
