"""Test helpers for the sparse-DP chaining (SDP): flat read sets, ctypes bindings of the CPU
checkers (oracle/build/libsdp_oracle.so = this repo's restatement, oracle/_ref/liblamsa_ref.so =
the unmodified reference, only where it was built), the recorder-file parser and synthetic
seed-hit generators.  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "build", "libsdp_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblamsa_ref.so")

HIT_DTYPE = np.dtype([("offset", "<i8"), ("nchr", "<i4"), ("NM", "<i4"), ("len_dif", "<i4"), ("nstrand", "<i4")])
REG_DTYPE = np.dtype([("beg", "<i4"), ("end", "<i4"), ("chr", "<i4"), ("is_rev", "<i4"),
                      ("ref_beg", "<i8"), ("ref_end", "<i8")])
READ_DTYPE = np.dtype([("seed_out", "<i4"), ("seed_all", "<i4"), ("read_len", "<i4"), ("n_reg", "<i4"),
                       ("seed_first", "<i8"), ("hit_first", "<i8"), ("reg_first", "<i8")])
PARA_DTYPE = np.dtype([("seed_len", "<i4"), ("seed_step", "<i4"), ("seed_inv", "<i4"), ("per_aln_m", "<i4"),
                       ("first_loci_thd", "<i4"), ("SV_len_thd", "<i4"), ("ske_max", "<i4"), ("ovlp_rat", "<f4"),
                       ("split_len", "<i4"), ("match_dis", "<i4"), ("mismatch_thd", "<i4"), ("aln_mode", "<i4"),
                       ("bwt_seed_len", "<i4"), ("frag_score_table", "<i4", (10,))])
assert HIT_DTYPE.itemsize == 24 and REG_DTYPE.itemsize == 32 and READ_DTYPE.itemsize == 40 and PARA_DTYPE.itemsize == 92

SCORE_TABLE = (1, 1, 1, 1, -3, -3, -3, -3, -6, -6)      # src/lamsa_aln.c:177-188


def default_para(mode="default"):
    """lamsa_aln_para fields the chaining reads, as `lamsa aln` sets them (src/lamsa_aln.c:1281-1420)."""
    p = np.zeros((), PARA_DTYPE)
    p["per_aln_m"], p["first_loci_thd"], p["SV_len_thd"], p["ske_max"], p["ovlp_rat"] = 200, 2, 10000, 10, 0.7
    p["split_len"], p["bwt_seed_len"] = 100, 19
    p["frag_score_table"] = SCORE_TABLE
    if mode == "default":
        p["seed_len"], p["seed_step"], p["match_dis"], p["mismatch_thd"], p["aln_mode"] = 50, 100, 5, 10, 0
    elif mode == "pacbio":
        p["seed_len"], p["seed_step"], p["match_dis"], p["mismatch_thd"], p["aln_mode"] = 50, 25, 8, 10, 1
    elif mode == "ont2d":
        p["seed_len"], p["seed_step"], p["match_dis"], p["mismatch_thd"], p["aln_mode"] = 50, 25, 3, 10, 3
    else:
        raise ValueError(mode)
    p["seed_inv"] = p["seed_step"] - p["seed_len"]
    return p


class ReadSet:
    """Flat description of a set of reads (include/lamsa_b200.h section 3)."""

    def __init__(self, para, reads, seed_id, map_n, hits, regs=None):
        self.para = np.array(para, dtype=PARA_DTYPE).reshape(())
        self.reads = np.ascontiguousarray(reads, dtype=READ_DTYPE)
        self.seed_id = np.ascontiguousarray(seed_id, dtype=np.int32)
        self.map_n = np.ascontiguousarray(map_n, dtype=np.int32)
        self.hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        self.regs = np.ascontiguousarray(regs if regs is not None else np.zeros(0, REG_DTYPE), dtype=REG_DTYPE)

    def __len__(self):
        return len(self.reads)

    def subset(self, idx):
        """Reads idx (any order) as a new, re-based ReadSet."""
        reads = self.reads[idx].copy()
        sid, mn, hits, regs = [], [], [], []
        s0 = h0 = r0 = 0
        for r in reads:
            so, hf, sf = int(r["seed_out"]), int(r["hit_first"]), int(r["seed_first"])
            m = self.map_n[sf:sf + so]
            nh = int(m.sum())
            sid.append(self.seed_id[sf:sf + so]); mn.append(m); hits.append(self.hits[hf:hf + nh])
            regs.append(self.regs[int(r["reg_first"]):int(r["reg_first"]) + int(r["n_reg"])])
            r["seed_first"], r["hit_first"], r["reg_first"] = s0, h0, r0
            s0 += so; h0 += nh; r0 += int(r["n_reg"])
        cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
        return ReadSet(self.para, reads, cat(sid, np.int32), cat(mn, np.int32), cat(hits, HIT_DTYPE), cat(regs, REG_DTYPE))

    def save_dict(self, prefix=""):
        return {prefix + "para": self.para, prefix + "reads": self.reads, prefix + "seed_id": self.seed_id,
                prefix + "map_n": self.map_n, prefix + "hits": self.hits, prefix + "regs": self.regs}

    @staticmethod
    def from_dict(d, prefix=""):
        return ReadSet(d[prefix + "para"], d[prefix + "reads"], d[prefix + "seed_id"], d[prefix + "map_n"],
                       d[prefix + "hits"], d[prefix + "regs"])


def concat(sets):
    """Concatenate ReadSets that share one parameter block."""
    reads, s0, h0, r0 = [], 0, 0, 0
    for s in sets:
        assert s.para.tobytes() == sets[0].para.tobytes()
        r = s.reads.copy()
        r["seed_first"] += s0; r["hit_first"] += h0; r["reg_first"] += r0
        reads.append(r)
        s0 += len(s.seed_id); h0 += len(s.hits); r0 += len(s.regs)
    return ReadSet(sets[0].para, np.concatenate(reads), np.concatenate([s.seed_id for s in sets]),
                   np.concatenate([s.map_n for s in sets]), np.concatenate([s.hits for s in sets]),
                   np.concatenate([s.regs for s in sets]))


# ------------------------------------------------------------------ CPU checkers --
def ensure_oracle():
    src = os.path.join(ROOT, "oracle", "sdp_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sdp"])
    return ORACLE_SO


def have_ref():
    return os.path.exists(REF_SO)


_libs = {}


def _run(path, fn, rs, stages, with_pairs):
    key = (path, fn)
    if key not in _libs:
        lib = C.CDLL(path)
        f = getattr(lib, fn)
        f.restype = C.c_int
        _libs[key] = (lib, f)
    f = _libs[key][1]
    n = len(rs)
    cap = 64 + 8 * n + 4 * len(rs.hits)
    while True:
        out1, out2 = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        off1, off2 = np.zeros(n + 1, np.int64), np.zeros(n + 1, np.int64)
        pairs = np.zeros(2, np.int64)
        args = [C.c_void_p(rs.para.ctypes.data), C.c_int64(n), C.c_void_p(rs.reads.ctypes.data),
                C.c_void_p(rs.seed_id.ctypes.data), C.c_void_p(rs.map_n.ctypes.data), C.c_void_p(rs.hits.ctypes.data),
                C.c_void_p(rs.regs.ctypes.data), C.c_int(stages),
                C.c_void_p(out1.ctypes.data), C.c_int64(cap), C.c_void_p(off1.ctypes.data),
                C.c_void_p(out2.ctypes.data), C.c_int64(cap), C.c_void_p(off2.ctypes.data)]
        if with_pairs:
            args.append(C.c_void_p(pairs.ctypes.data))
        rc = f(*args)
        if rc == 0:
            break
        cap *= 4
    return (out1[:off1[-1]], off1), (out2[:off2[-1]], off2), pairs


def oracle_run(rs, stages=3):
    """-> ((stream1, off1), (stream2, off2), pairs[2]) from this repo's C restatement."""
    return _run(ensure_oracle(), "orc_sdp_run_batch", rs, stages, True)


def ref_run(rs, stages=3):
    """Same, from the unmodified reference frag_line_BCC / frag_line_remain."""
    return _run(REF_SO, "ref_sdp_run_batch", rs, stages, False)


def diff_streams(a, b, what=""):
    """Compare two (stream, off) results read by read -> list of mismatch strings."""
    (sa, oa), (sb, ob) = a, b
    bad = []
    for r in range(len(oa) - 1):
        x, y = sa[oa[r]:oa[r + 1]], sb[ob[r]:ob[r + 1]]
        if len(x) != len(y) or not np.array_equal(x, y):
            bad.append(f"{what} read {r}: {x[:24].tolist()}... vs {y[:24].tolist()}... (len {len(x)} vs {len(y)})")
            if len(bad) >= 8:
                break
    return bad


# --------------------------------------------------------------- recorder files --
def parse_recording(path):
    """oracle/_ref/lamsa_rec output (layout: oracle/sdp_ref_shim.c) -> (ReadSet, expected1, expected2)
    where expectedK = (stream, off) of the reference for stage K."""
    w = np.fromfile(path, dtype="<i4")
    pw = PARA_DTYPE.itemsize // 4
    reads, sid, mn, hits, regs, e1, e2 = [], [], [], [], [], [], []
    para = None
    p = 0
    s0 = h0 = r0 = 0
    while p < len(w):
        tag, nw = int(w[p]), int(w[p + 1]); q = p + 2; p = q + nw
        if tag == 1:
            pa = w[q:q + pw].view(PARA_DTYPE)[0]; q += pw
            if para is None:
                para = pa
            assert pa.tobytes() == para.tobytes()
            so, sa, rl = (int(v) for v in w[q:q + 3]); q += 3
            sm = w[q:q + 2 * so].reshape(so, 2); q += 2 * so
            nh = int(sm[:, 1].sum())
            hh = w[q:q + 6 * nh].view(HIT_DTYPE); q += 6 * nh
            no = int(w[q]); q += 1
            e1.append(w[q:q + no])
            reads.append((so, sa, rl, 0, s0, h0, r0))
            sid.append(sm[:, 0]); mn.append(sm[:, 1]); hits.append(hh)
            s0 += so; h0 += nh
        else:
            nr = int(w[q]); q += 1
            rg = w[q:q + 8 * nr].view(REG_DTYPE); q += 8 * nr
            no = int(w[q]); q += 1
            e2.append(w[q:q + no])
            so, sa, rl, _, sf, hf, _ = reads[-1]
            reads[-1] = (so, sa, rl, nr, sf, hf, r0)
            regs.append(rg); r0 += nr
    assert len(e1) == len(e2) == len(reads)
    rs = ReadSet(para, np.array(reads, dtype=READ_DTYPE), np.concatenate(sid), np.concatenate(mn),
                 np.concatenate(hits) if hits else np.zeros(0, HIT_DTYPE),
                 np.concatenate(regs) if regs else np.zeros(0, REG_DTYPE))

    def pack(es):
        off = np.zeros(len(es) + 1, np.int64)
        off[1:] = np.cumsum([len(e) for e in es])
        return np.concatenate(es).astype(np.int32), off
    return rs, pack(e1), pack(e2)


# ---------------------------------------------------------- synthetic seed hits --
def gen_reads(n_reads, seed=0, mode="default", read_len=(2000, 12000), repeat_frac=0.15, sv_rate=0.3,
              miss_frac=0.3, max_hits=12, n_chr=3, with_regs=True):
    """Seed-hit sets shaped like GEM output on a genome with repeats and structural variants:
    a true path (with deletions / insertions / inversions / translocations), seeds without hits,
    seeds with extra hits elsewhere, co-linear decoy copies (tandem repeats).  Aligned records for
    stage 2 cover random stretches of the read."""
    rng = np.random.default_rng(seed)
    para = default_para(mode)
    L, S = int(para["seed_len"]), int(para["seed_step"])
    reads, sid, mn, hits, regs = [], [], [], [], []
    s0 = h0 = r0 = 0
    for _ in range(n_reads):
        rl = int(rng.integers(read_len[0], read_len[1] + 1))
        seed_all = max(1, (rl - L) // S + 1)
        chr_, strand = int(rng.integers(0, n_chr)), int(rng.choice([1, -1]))
        pos = int(rng.integers(100000, 50000000))
        rd_sid, rd_mn, rd_hits = [], [], []
        drift = 0
        for k in range(1, seed_all + 1):
            if rng.random() < sv_rate * S / 1500.0:                  # structural event between seeds
                ev = rng.integers(0, 5)
                if ev == 0: drift += int(rng.integers(20, 3000)) * strand        # deletion
                elif ev == 1: drift -= int(rng.integers(20, 600)) * strand       # insertion / duplication
                elif ev == 2: strand = -strand; drift += int(rng.integers(-2000, 2000))  # inversion
                elif ev == 3: chr_ = int(rng.integers(0, n_chr)); pos = int(rng.integers(100000, 50000000)); drift = 0
                else: drift += int(rng.integers(-12, 13))                        # small indel drift
            if rng.random() < miss_frac:
                continue
            true_off = pos + strand * (k - 1) * S + drift + int(rng.integers(-2, 3))
            hs = [(true_off, chr_, int(rng.integers(0, 4)), int(rng.integers(-2, 3)), strand)]
            if rng.random() < repeat_frac:
                for _e in range(int(rng.integers(1, max_hits))):
                    kind = rng.random()
                    if kind < 0.4:        # tandem copy near by
                        hs.append((true_off + int(rng.integers(-4, 5)) * int(rng.integers(40, 400)), chr_,
                                   int(rng.integers(0, 5)), int(rng.integers(-2, 3)), strand))
                    elif kind < 0.7:      # same chromosome, other strand or far away
                        hs.append((int(rng.integers(100000, 50000000)), chr_, int(rng.integers(0, 5)),
                                   int(rng.integers(-2, 3)), int(rng.choice([1, -1]))))
                    else:
                        hs.append((int(rng.integers(100000, 50000000)), int(rng.integers(0, n_chr)),
                                   int(rng.integers(0, 5)), int(rng.integers(-2, 3)), int(rng.choice([1, -1]))))
                order = rng.permutation(len(hs))
                hs = [hs[i] for i in order]
            if rng.random() < 0.03:       # a seed whose true hit is missing but decoys exist
                hs = hs[1:] if len(hs) > 1 else hs
            rd_sid.append(k); rd_mn.append(len(hs)); rd_hits.extend(hs)
        n_reg = 0
        if with_regs and rng.random() < 0.9:
            cur = 1
            while cur < rl and rng.random() < 0.85:
                b = cur + int(rng.integers(0, 1500)); e = min(rl, b + int(rng.integers(100, 4000)))
                if b >= rl:
                    break
                rp = pos + int(rng.integers(-3000, 3000))
                regs.append((b, e, int(rng.integers(0, n_chr)) if rng.random() < 0.2 else chr_, int(strand == -1),
                             rp, rp + (e - b) + int(rng.integers(-50, 50))))
                n_reg += 1
                cur = e + int(rng.integers(-20, 800))
        reads.append((len(rd_sid), seed_all, rl, n_reg, s0, h0, r0))
        sid.extend(rd_sid); mn.extend(rd_mn); hits.extend(rd_hits)
        s0 += len(rd_sid); h0 += len(rd_hits); r0 += n_reg
    return ReadSet(para, np.array(reads, dtype=READ_DTYPE), np.array(sid, np.int32), np.array(mn, np.int32),
                   np.array(hits, dtype=HIT_DTYPE) if hits else np.zeros(0, HIT_DTYPE),
                   np.array(regs, dtype=REG_DTYPE) if regs else np.zeros(0, REG_DTYPE))
