"""Shared by test_wrappers_oracle.py (CPU) and test_gpu_wrappers.py (B200): seeded read/reference
pairs for the host wrappers of the banded DP -- ksw_extend_c / ksw_extend_r (reference
src/ksw.c:809-836), sw_mid_fix (:841-860), ksw_bi_extend (:862-926) -- and per-call ctypes drivers
for the three implementations that export them with the same signatures:
   oracle/_ref/libksw_ref.so   the unmodified reference (takes a genuine lamsa_aln_para, whose
                               layout equals lamsa_b200.AlnPara -- tests/test_abi.py),
   oracle/build/libdp_oracle.so  the C restatement (orc_*, flat parameter structs),
   lamsa_b200/liblamsa_b200.so   the CUDA path behind the reference's own symbol names.
Test infrastructure only."""
import ctypes as C
import os

import numpy as np

import _oracle
from lamsa_b200 import AlnPara, default_matrix

u8p, i8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.POINTER(C.c_int)
cpp = C.POINTER(C.POINTER(C.c_int32))


class OrcExtPar(C.Structure):       # oracle/dp_oracle.h:orc_ext_par
    _fields_ = [(n, C.c_int) for n in ("o_del", "e_del", "o_ins", "e_ins", "end_bonus", "zdrop")]


class OrcBiPar(C.Structure):        # oracle/dp_oracle.h:orc_bi_par
    _fields_ = [("ext", OrcExtPar)] + [(n, C.c_int) for n in ("del_gapo", "del_gape", "ins_gapo", "ins_gape",
                                                               "band_w", "split_len", "aln_mode")] + [("id_rate", C.c_float)]


def make_para(rng):
    """Penalty / threshold sets in the range of the reference's presets (src/lamsa_aln.h:17-77)."""
    AP = AlnPara()
    pen = [(5, 2, 5, 2), (1, 1, 1, 1), (2, 1, 2, 1), (5, 2, 1, 1)][int(rng.integers(0, 4))]
    AP.del_ext_o, AP.del_ext_e, AP.ins_ext_o, AP.ins_ext_e = pen
    pen = [(5, 2, 5, 2), (1, 1, 1, 1), (1, 1, 5, 2)][int(rng.integers(0, 3))]
    AP.del_gapo, AP.del_gape, AP.ins_gapo, AP.ins_gape = pen
    AP.end_bonus = int(rng.choice([0, 5]))
    AP.zdrop = int(rng.choice([100, 100, 30, 10]))
    AP.band_w = int(rng.choice([5, 10, 10, 30]))
    AP.split_len = int(rng.choice([100, 100, 40, 20]))
    AP.aln_mode = int(rng.choice([0, 2, 2, 3]))          # bit 1 = aln_mode_high_id_err (src/lamsa_aln.h:438)
    AP.id_rate = float(rng.choice([0.04, 0.15, 0.3]))
    return AP


def orc_para(AP):
    P = OrcBiPar()
    P.ext = OrcExtPar(AP.del_ext_o, AP.del_ext_e, AP.ins_ext_o, AP.ins_ext_e, AP.end_bonus, AP.zdrop)
    P.del_gapo, P.del_gape, P.ins_gapo, P.ins_gape = AP.del_gapo, AP.del_gape, AP.ins_gapo, AP.ins_gape
    P.band_w, P.split_len, P.aln_mode, P.id_rate = AP.band_w, AP.split_len, AP.aln_mode, AP.id_rate
    return P


def gen_pairs(n, seed, qmax=500):
    """(query, target, lh0, rh0, AP) tuples that reach every exit of ksw_bi_extend: clean pairs (the left
    extension reaches an end), pairs with a junk / inserted / deleted middle (both extensions stop: global
    re-alignment when `near`, else sw_mid_fix with or without the nS mH seam), junk prefixes (the left extension
    dies at once), empty sequences."""
    rng = np.random.default_rng(seed)
    out = []

    def mut(q, e):
        r = rng.random(len(q))
        keep = r >= e / 3
        t = q.copy()
        sub = (r >= e / 3) & (r < 2 * e / 3)
        t[sub] = (t[sub] + rng.integers(1, 4, size=int(sub.sum()))) & 3
        return t[keep]

    for k in range(n):
        AP = make_para(rng)
        ql = int(rng.integers(0, qmax)) if k % 23 else int(rng.integers(0, 3))
        q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        if ql > 4 and rng.random() < 0.1:
            q[rng.integers(0, ql, size=max(1, ql // 50))] = 4              # reads may hold N
        style = int(rng.integers(0, 6))
        e = float(rng.choice([0.0, 0.03, 0.1, 0.2]))
        base = mut(np.minimum(q, 3), e)
        if style == 0 or len(base) < 8:
            t = base
        else:
            a = int(rng.integers(1, len(base) - 1)); b = int(rng.integers(a, len(base)))
            gap = int(rng.choice([3, 15, 40, 120, 300]))
            junk = rng.integers(0, 4, size=gap, dtype=np.uint8)
            if style == 1:   t = np.concatenate((base[:a], junk, base[a:]))          # insertion in the target
            elif style == 2: t = np.concatenate((base[:a], base[min(len(base), a + gap):]))   # deletion
            elif style == 3: t = np.concatenate((base[:a], junk, base[b:]))          # replaced middle
            elif style == 4: t = np.concatenate((junk, base[a:]))                    # junk prefix
            else:            t = np.concatenate((base[:b], junk))                    # junk suffix
        t = np.ascontiguousarray(t, dtype=np.uint8)
        lh0, rh0 = int(rng.choice([10, 19, 50, 100])), int(rng.choice([10, 19, 50, 100]))
        out.append((q, t, lh0, rh0, AP))
    return out


class Impl:
    """Per-call driver of one implementation; results as plain Python tuples."""

    def __init__(self, which):
        self.which = which
        if which == "ref":
            self.lib = C.CDLL(_oracle.REF_SO)
            self.free = C.CDLL(None).free
            pre = "ksw_"
        elif which == "oracle":
            self.lib = C.CDLL(_oracle.ensure_oracle())
            self.free = C.CDLL(None).free
            pre = "orc_"
        else:
            import lamsa_b200
            self.lib = lamsa_b200.load_library()
            self.free = self.lib.lb2_free
            pre = "ksw_"
        self.free.argtypes = [C.c_void_p]; self.free.restype = None
        par_ext = C.POINTER(OrcExtPar) if which == "oracle" else C.POINTER(AlnPara)
        par_bi = C.POINTER(OrcBiPar) if which == "oracle" else C.POINTER(AlnPara)
        self.ext_c = getattr(self.lib, pre + "extend_c"); self.ext_r = getattr(self.lib, pre + "extend_r")
        for f in (self.ext_c, self.ext_r):
            f.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, i8p, C.c_int, C.c_int, par_ext, ip, ip, cpp, ip, ip]
            f.restype = C.c_int
        self.bi = getattr(self.lib, pre + "bi_extend")
        self.bi.argtypes = [C.c_int, u8p, C.c_int, u8p, C.c_int, i8p, C.c_int, C.c_int, par_bi, cpp, ip, ip]
        self.bi.restype = C.c_int
        self.mid = getattr(self.lib, "orc_mid_fix" if which == "oracle" else "sw_mid_fix")
        self.mid.argtypes = [cpp, ip, ip, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), C.c_int,
                             u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int, par_bi, C.c_int, i8p]
        self.mid.restype = None
        self.mat = default_matrix(1, 3)

    def _par(self, AP, bi):
        if self.which != "oracle":
            return C.byref(AP)
        P = orc_para(AP)
        self._keep = P
        return C.byref(P) if bi else C.byref(P.ext)

    def _take(self, c, n):
        out = [int(c[i]) for i in range(n)] if n else []
        if c:
            self.free(C.cast(c, C.c_void_p))
        return out

    @staticmethod
    def _buf(a):
        b = np.concatenate((a, np.zeros(8, np.uint8)))
        return b, b.ctypes.data_as(u8p)

    def extend(self, rev, q, t, w, h0, AP):
        qb, qp = self._buf(q); tb, tp = self._buf(t)
        qle, tle, n, cap = C.c_int(-7), C.c_int(-7), C.c_int(0), C.c_int(0)
        c = C.POINTER(C.c_int32)()
        f = self.ext_r if rev else self.ext_c
        r = f(len(q), qp, len(t), tp, 5, self.mat.ctypes.data_as(i8p), w, h0, self._par(AP, False),
              C.byref(qle), C.byref(tle), C.byref(c), C.byref(n), C.byref(cap))
        return (r, qle.value, tle.value, cap.value, self._take(c, n.value))

    def bi_extend(self, q, t, lh0, rh0, AP):
        qb, qp = self._buf(q); tb, tp = self._buf(t)
        n, cap = C.c_int(0), C.c_int(0)
        c = C.POINTER(C.c_int32)()
        r = self.bi(len(q), qp, len(t), tp, 5, self.mat.ctypes.data_as(i8p), lh0, rh0, self._par(AP, True),
                    C.byref(c), C.byref(n), C.byref(cap))
        # m_cigar is compared only where the reference defines it (it leaves *m_cigar_ = n after the global exits)
        return (r, cap.value, self._take(c, n.value))

    def mid_fix(self, q, t, lc, rc, lqe, rqe, lte, rte, AP):
        """sw_mid_fix on given left / right CIGARs (rc already in forward order, as ksw_bi_extend passes it)."""
        qb, qp = self._buf(q); tb, tp = self._buf(t)
        libc = C.CDLL(None); libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]
        out = C.cast(libc.malloc(40), C.POINTER(C.c_int32))                    # src/ksw.c:911-912
        n, cap = C.c_int(0), C.c_int(10)
        la = (C.c_int32 * max(1, len(lc)))(*lc); ra = (C.c_int32 * max(1, len(rc)))(*rc)
        self.mid(C.byref(out), C.byref(n), C.byref(cap), la, len(lc), ra, len(rc), qp, len(q), lqe, rqe,
                 tp, len(t), lte, rte, self._par(AP, True), 5, self.mat.ctypes.data_as(i8p))
        words = [int(out[i]) for i in range(n.value)]
        C.CDLL(None).free(C.cast(out, C.c_void_p)) if self.which != "gpu" else self.free(C.cast(out, C.c_void_p))
        return (cap.value, words)


def run_all(impl, pairs):
    """Every wrapper on every pair -> list of result tuples (the comparison key)."""
    out = []
    for q, t, lh0, rh0, AP in pairs:
        dl = abs(len(q) - len(t))
        w = max(dl + 3, AP.band_w)                                            # what ksw_bi_extend passes (:873)
        L = impl.extend(False, q, t, w, lh0, AP)
        R = impl.extend(True, q, t, w, rh0, AP)
        B = impl.bi_extend(q, t, lh0, rh0, AP)
        # sw_mid_fix directly, on the two extensions' own outputs (whatever their exit codes were)
        M = impl.mid_fix(q, t, L[4], R[4][::-1], max(L[1], 0), max(R[1], 0), max(L[2], 0), max(R[2], 0), AP)
        out.append((L, R, B, M))
    return out


def classify_exit(L, R, q, t, AP):
    """Which exit of ksw_bi_extend (src/ksw.c:874-924) a pair takes, from the two extension results."""
    dl = abs(len(q) - len(t))
    near = dl < AP.split_len + np.float32(len(t)) * np.float32(AP.id_rate) * (AP.aln_mode & 2)
    if L[0] < 2:
        return "left_end"
    if near and (2 * L[1] > len(q) or 2 * L[2] > len(t)):
        return "left_global"
    if R[0] < 2:
        return "right_end"
    if near and (2 * R[1] > len(q) or 2 * R[2] > len(t)):
        return "right_global"
    Sn, Hn, half = len(q) - L[1] - R[1], len(t) - L[2] - R[2], AP.split_len // 2
    return "mid_clip" if (abs(Sn) >= half or abs(Hn) >= half or abs(Sn - Hn) >= half) else "mid_global"
