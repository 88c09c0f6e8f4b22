"""Test helpers for the local Smith-Waterman entry points (reference src/ksw.c:68-377): seeded pairs and ctypes
drivers of the unmodified reference (oracle/_ref/libksw_ref.so), the oracle (oracle/build/libsw_oracle.so) and
the CUDA drop-ins (liblamsa_b200.so).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "build", "libsw_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libksw_ref.so")
XBYTE, XSTOP, XSUBO, XSTART = 0x10000, 0x20000, 0x40000, 0x80000
_libs = {}


class Kswr(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("score", "te", "qe", "score2", "te2", "tb", "qb")]

    def tup(self):
        return (self.score, self.te, self.qe, self.score2, self.te2, self.tb, self.qb)


def have_ref():
    return os.path.exists(REF_SO)


def _lib(path):
    if path not in _libs:
        if path == ORACLE_SO and not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "sw"])
        _libs[path] = C.CDLL(path)
    return _libs[path]


def _align2(lib, c):
    lib.ksw_align2.restype = Kswr
    lib.ksw_align2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    q = np.ascontiguousarray(c["q"], np.uint8).copy(); t = np.ascontiguousarray(c["t"], np.uint8).copy()
    mat = np.ascontiguousarray(c["mat"], np.int8)
    r = lib.ksw_align2(len(q), q.ctypes.data, len(t), t.ctypes.data, c["m"], mat.ctypes.data, c["o_del"], c["e_del"], c["o_ins"], c["e_ins"], c["xtra"], None)
    assert (q == c["q"]).all() and (t == c["t"]).all()          # the start-point pass reverses in place and restores
    return r.tup()


def ref_align2(c):
    return _align2(_lib(REF_SO), c)


def gpu_align2(c):
    from lamsa_b200 import load_library
    return _align2(load_library(), c)


def oracle_align2(c):
    lib = _lib(ORACLE_SO)
    out = (C.c_int * 7)()
    q = np.ascontiguousarray(c["q"], np.uint8); t = np.ascontiguousarray(c["t"], np.uint8); mat = np.ascontiguousarray(c["mat"], np.int8)
    lib.orc_sw_align2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.orc_sw_align2(len(q), q.ctypes.data, len(t), t.ctypes.data, c["m"], mat.ctypes.data, c["o_del"], c["e_del"], c["o_ins"], c["e_ins"], c["xtra"], out)
    return tuple(out)


def gen_cases(n, seed, qmax=300, tmax=900):
    """query = a piece of the target with errors (or unrelated), sometimes two copies of the piece in the target (a second
    best hit), all flag combinations, both profile widths; byte mode is pushed into overflow by long perfect matches."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        m = 5
        a, b = [(1, 3), (1, 1), (2, 4), (5, 4)][int(rng.integers(0, 4))]
        mat = np.full((5, 5), -b, np.int8); np.fill_diagonal(mat, a); mat[4, :] = -1; mat[:, 4] = -1
        ql = int(rng.integers(1, qmax)); tl = int(rng.integers(1, tmax))
        t = rng.integers(0, 4, size=tl, dtype=np.uint8)
        if k % 5 == 4:
            q = rng.integers(0, 4, size=ql, dtype=np.uint8)
        else:
            s = int(rng.integers(0, max(1, tl - 1))); q = t[s:s + ql].copy()
            e = float(rng.choice([0.0, 0.03, 0.1, 0.25]))
            if len(q) > 3 and e > 0:
                keep = rng.random(len(q)) >= e / 3; q = q[keep]
                sub = rng.random(len(q)) < e / 3; q[sub] = (q[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
                ins = np.flatnonzero(rng.random(len(q)) < e / 3); q = np.insert(q, ins, rng.integers(0, 4, size=len(ins), dtype=np.uint8))
            if k % 3 == 0 and len(q) > 10:          # a second, partial copy further down the target
                t = np.concatenate((t, rng.integers(0, 4, size=int(rng.integers(5, 80)), dtype=np.uint8), q[: int(len(q) * rng.uniform(0.4, 1.0))]))
        if len(q) == 0:
            q = rng.integers(0, 4, size=1, dtype=np.uint8)
        if k % 9 == 5:
            q[rng.integers(0, len(q), size=2)] = 4
        pen = [(5, 2, 5, 2), (1, 1, 1, 1), (6, 1, 6, 1), (3, 2, 5, 1)][int(rng.integers(0, 4))]
        xtra = 0
        if k % 2: xtra |= XBYTE
        if k % 3 != 1: xtra |= XSTART
        if k % 4 >= 2: xtra |= XSUBO | int(rng.integers(5, 60))
        if k % 11 == 7: xtra = (xtra & ~0xffff & ~XSUBO) | XSTOP | int(rng.integers(5, 60))
        out.append(dict(q=q.astype(np.uint8), t=t.astype(np.uint8), m=m, mat=mat.reshape(-1), o_del=pen[0], e_del=pen[1], o_ins=pen[2], e_ins=pen[3], xtra=xtra))
    return out
