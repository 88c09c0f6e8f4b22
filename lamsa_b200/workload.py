"""Synthetic DP task streams (seeded, vectorised numpy; no reference data needed).

``gen_microbench`` is BASELINE.json configs[1] / SURVEY.md 8d "C2": qlen~U[50,1000],
target = mutated query (error~U[0,0.2], 1/3 sub, 1/3 ins, 1/3 del) with
abs(tlen-qlen)<=50, w~U{10..200}; half ksw_global2 (penalties 5/2/5/2 and
1/1/1/1), half ksw_extend_core (h0 in {8,10,19,50,100}, zdrop 100, end_bonus
{0,5}); matrix +1/-3/N -1.
"""
import numpy as np

from ._lib import FLAG_CIGAR, KIND_EXTEND, KIND_GLOBAL, TASK_DTYPE
from .ksw import default_matrix


def mutate_pool(rng, qseq, qoff, qlen, err):
    """Apply per-task error rates to pooled queries -> (tseq, toff, tlen_raw)."""
    total = int(qlen.sum())
    thr = np.repeat((err * 65535.0 / 3.0).astype(np.float32), qlen)
    r = rng.integers(0, 65536, size=total, dtype=np.uint16).astype(np.float32)
    is_sub = r < thr
    is_del = (~is_sub) & (r < 2 * thr)
    is_ins = (~is_sub) & (~is_del) & (r < 3 * thr)
    cnt = np.ones(total, dtype=np.int64)
    cnt[is_del] = 0
    cnt[is_ins] = 2
    start = np.cumsum(cnt) - cnt
    ttotal = int(cnt.sum())
    src = np.repeat(np.arange(total, dtype=np.int64), cnt)
    tseq = np.empty(ttotal + 256, dtype=np.uint8)
    tseq[:ttotal] = qseq[src]
    tseq[ttotal:] = rng.integers(0, 4, size=256, dtype=np.uint8)
    sub_pos = start[is_sub]
    tseq[sub_pos] = (qseq[:total][is_sub] + rng.integers(1, 4, size=sub_pos.size, dtype=np.uint8)) & 3
    ins_pos = start[is_ins]
    tseq[ins_pos] = rng.integers(0, 4, size=ins_pos.size, dtype=np.uint8)
    csum = np.concatenate(([0], np.cumsum(cnt)))
    toff = csum[qoff]
    tlen = csum[qoff + qlen] - toff
    return tseq, toff.astype(np.int64), tlen.astype(np.int64)


def gen_microbench(n, seed=20260101, qmin=50, qmax=1000, wmin=10, wmax=200, max_err=0.2,
                   max_dl=50, chunk=100_000, cigar=True):
    """Returns (tasks, keep): TASK_DTYPE array and the buffers it points into."""
    rng = np.random.default_rng(seed)
    mat = default_matrix(1, 3)
    parts, keep = [], [mat]
    h0_choices = np.array([8, 10, 19, 50, 100], dtype=np.int32)
    done = 0
    while done < n:
        k = min(chunk, n - done)
        qlen = rng.integers(qmin, qmax + 1, size=k).astype(np.int64)
        qoff = np.concatenate(([0], np.cumsum(qlen)[:-1])).astype(np.int64)
        total = int(qlen.sum())
        qseq = rng.integers(0, 4, size=total + 64, dtype=np.uint8)
        # ~0.5% of query bases become N (code 4): reads may contain N, reference windows may not
        nmask = rng.integers(0, 200, size=total + 64, dtype=np.uint8) == 0
        qseq[nmask] = 4
        err = rng.uniform(0.0, max_err, size=k)
        tseq, toff, tlen = mutate_pool(rng, np.minimum(qseq, 3), qoff, qlen, err)
        tlen = np.clip(tlen, np.maximum(qlen - max_dl, 1), qlen + max_dl)
        tlen = np.minimum(tlen, tseq.size - toff)
        t = np.zeros(k, dtype=TASK_DTYPE)
        kind = (np.arange(done, done + k) & 1).astype(np.int32)       # alternate global / extend
        t["kind"] = np.where(kind == 0, KIND_GLOBAL, KIND_EXTEND)
        t["flags"] = FLAG_CIGAR if cigar else 0
        t["qlen"], t["tlen"] = qlen, tlen
        t["query"] = qseq.ctypes.data + qoff.astype(np.uint64)
        t["target"] = tseq.ctypes.data + toff.astype(np.uint64)
        t["w"] = rng.integers(wmin, wmax + 1, size=k)
        pen = rng.integers(0, 2, size=k)
        t["o_del"] = np.where(pen == 0, 5, 1)
        t["e_del"] = np.where(pen == 0, 2, 1)
        t["o_ins"] = np.where(pen == 0, 5, 1)
        t["e_ins"] = np.where(pen == 0, 2, 1)
        t["h0"] = np.where(kind == 1, h0_choices[rng.integers(0, 5, size=k)], 0)
        t["zdrop"] = np.where(kind == 1, 100, 0)
        t["end_bonus"] = np.where(kind == 1, rng.integers(0, 2, size=k) * 5, 0)
        t["m"] = 5
        t["mat"] = mat.ctypes.data
        parts.append(t)
        keep += [qseq, tseq]
        done += k
    return np.concatenate(parts), keep


def pool_tasks(tasks, keep=(), alloc=None):
    """Copy every sequence of `tasks` into ONE pool, in task order, and re-point the records at it (lb2_pool_pack)
    -> (tasks', pool).  `alloc(nbytes)` returns the pool's uint8 array (lamsa_b200.pinned_pool for page-locked
    memory; default a plain numpy array).  Matrices keep their own buffers."""
    import ctypes as C
    from ._lib import load_library
    lib = load_library()
    t = np.ascontiguousarray(tasks, dtype=TASK_DTYPE).copy()
    pac = (t["flags"] & 2) != 0
    total = int(((t["qlen"].astype(np.int64) + np.where(pac, 0, t["tlen"]).astype(np.int64) + 15) & ~15).sum()) + 64
    pool = alloc(total) if alloc is not None else np.zeros(total, dtype=np.uint8)
    used = C.c_int64()
    if lib.lb2_pool_pack(len(t), t.ctypes.data, pool.ctypes.data, pool.nbytes, C.byref(used)):
        raise RuntimeError("lb2_pool_pack: " + lib.lb2_last_error().decode())
    return t, pool


def gen_edge_cases(seed=7):
    """Small adversarial tasks: empty sequences, band beyond the length difference,
    all-N reads, dying first rows, z-drop, end-bonus ties, narrow and huge bands."""
    rng = np.random.default_rng(seed)
    mat13 = default_matrix(1, 3)
    mat11 = default_matrix(1, 1)
    mat24 = default_matrix(2, 4)
    keep = [mat13, mat11, mat24]
    rows = []

    def add(kind, q, t, w, mat=mat13, pen=(5, 2, 5, 2), h0=0, zdrop=100, end_bonus=5, cigar=True):
        q = np.ascontiguousarray(q, dtype=np.uint8)
        t = np.ascontiguousarray(t, dtype=np.uint8)
        qb = np.concatenate((q, np.zeros(8, np.uint8)))
        tb = np.concatenate((t, np.zeros(8, np.uint8)))
        keep.extend([qb, tb])
        r = np.zeros(1, dtype=TASK_DTYPE)
        r["kind"], r["flags"] = kind, (FLAG_CIGAR if cigar else 0)
        r["qlen"], r["tlen"] = len(q), len(t)
        r["query"], r["target"] = qb.ctypes.data, tb.ctypes.data
        r["w"], r["h0"] = w, h0
        r["o_del"], r["e_del"], r["o_ins"], r["e_ins"] = pen
        r["end_bonus"], r["zdrop"] = end_bonus, zdrop
        r["m"], r["mat"] = 5, mat.ctypes.data
        rows.append(r)

    def rnd(n):
        return rng.integers(0, 4, size=n, dtype=np.uint8)

    def mut(q, e):
        out = []
        for b in q:
            x = rng.random()
            if x < e / 3:
                out.append((int(b) + int(rng.integers(1, 4))) & 3)
            elif x < 2 * e / 3:
                continue
            elif x < e:
                out.append(int(rng.integers(0, 4)))
                out.append(int(b))
            else:
                out.append(int(b))
        return np.array(out, dtype=np.uint8)

    G, E = KIND_GLOBAL, KIND_EXTEND
    # empty and tiny
    for ql, tl in [(0, 0), (0, 5), (5, 0), (1, 1), (1, 7), (7, 1), (2, 2), (0, 40), (40, 0)]:
        for w in (1, 3, 50):
            add(G, rnd(ql), rnd(tl), w)
            add(E, rnd(ql), rnd(tl), w, h0=10)
            add(E, rnd(ql), rnd(tl), w, h0=10, cigar=False)
            add(G, rnd(ql), rnd(tl), w, cigar=False)
    # length difference far beyond the requested band
    for ql, tl in [(30, 90), (90, 30), (200, 260), (500, 380), (31, 33), (32, 32), (33, 31), (63, 64), (64, 63)]:
        q = rnd(ql)
        for w in (1, 5, 10, 64):
            add(G, q, rnd(tl), w)
            add(G, q, rnd(tl), w, pen=(1, 1, 1, 1), mat=mat11)
            add(E, q, rnd(tl), w, h0=50)
    # all-N reads and N-rich reads
    add(G, np.full(60, 4), rnd(60), 10)
    add(E, np.full(60, 4), rnd(60), 10, h0=100)
    qn = rnd(300); qn[::3] = 4
    add(G, qn, mut(np.minimum(qn, 3), 0.05), 20)
    add(E, qn, mut(np.minimum(qn, 3), 0.05), 20, h0=19)
    # tiny h0: row 0 dies; unrelated sequences: z-drop / m==0 exits
    for h0 in (1, 2, 3, 8):
        add(E, rnd(100), rnd(100), 10, h0=h0)
        add(E, rnd(100), rnd(100), 10, h0=h0, pen=(1, 1, 1, 1), mat=mat11, end_bonus=0)
    for _ in range(20):
        q = rnd(int(rng.integers(50, 400)))
        t = np.concatenate((mut(q[: len(q) // 2], 0.05), rnd(len(q))))     # good prefix then junk
        add(E, q, t, int(rng.integers(5, 120)), h0=int(rng.choice([8, 19, 50, 100])))
        add(E, q, t, int(rng.integers(5, 120)), h0=100, pen=(2, 1, 2, 1), mat=mat11, end_bonus=0)
        add(E, q, t, 40, h0=100, zdrop=int(rng.choice([0, 5, 20])), mat=mat24)
    # end-bonus tie territory: perfect matches, query shorter / longer than target
    for ql, tl in [(50, 50), (50, 80), (80, 50), (120, 121)]:
        q = rnd(max(ql, tl))
        for eb in (0, 1, 5, 50):
            add(E, q[:ql], q[:tl], 10, h0=10, end_bonus=eb)
    # similar pairs across all lane widths, including the widest single-warp class
    for ql in (20, 31, 32, 45, 63, 64, 100, 127, 128, 200, 255, 256, 400, 511, 512, 700, 1000, 1023):
        q = rnd(ql)
        for e in (0.0, 0.1, 0.25):
            t = mut(q, e)
            for w in (1, 7, 15, 16, 30, 31, 61, 62, 123, 124, 200, 247, 248, 400, 495):
                if rng.random() < 0.35:
                    add(G, q, t, w)
                    add(E, q, t, w, h0=int(rng.choice([8, 50, 100])))
                    add(E, q, t, w, h0=100, pen=(2, 1, 2, 1), mat=mat11, end_bonus=0)
    # long sequences with a narrow band (chunk hand-over many times)
    for ql in (2000, 5000):
        q = rnd(ql)
        t = mut(q, 0.05)
        for w in (10, 50, 200, 400):
            add(G, q, t, w)
            add(E, q, t, w, h0=100)
            add(E, q, t, w, h0=100, cigar=False)
    return np.concatenate(rows), keep
