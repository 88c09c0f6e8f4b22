"""ctypes bindings of the CPU checkers (test infrastructure):
   oracle/build/libdp_oracle.so  -- this repo's C restatement (oracle/dp_oracle.c)
   oracle/_ref/libksw_ref.so     -- the unmodified reference ksw.c (only where it was built)
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "build", "libdp_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libksw_ref.so")

from lamsa_b200._lib import RESULT_DTYPE, TASK_DTYPE  # noqa: E402


def ensure_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "all"])
    return ORACLE_SO


def have_ref():
    return os.path.exists(REF_SO)


_libs = {}


def _load(path, fn):
    key = (path, fn)
    if key not in _libs:
        lib = C.CDLL(path)
        f = getattr(lib, fn)
        f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                      C.c_int, C.POINTER(C.c_double)]
        f.restype = C.c_int
        _libs[key] = (lib, f)
    return _libs[key]


def _run(path, fn, tasks, nthreads):
    lib, f = _load(path, fn)
    tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
    res = np.zeros(len(tasks), dtype=RESULT_DTYPE)
    pool, pn, secs = C.c_void_p(), C.c_int64(), C.c_double()
    f(len(tasks), tasks.ctypes.data, res.ctypes.data, C.byref(pool), C.byref(pn), nthreads, C.byref(secs))
    cig = np.ctypeslib.as_array(C.cast(pool, C.POINTER(C.c_int32)), shape=(max(pn.value, 1),))[:pn.value].copy()
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    libc.free(pool)
    return res, cig, secs.value


def oracle_run(tasks, nthreads=None):
    """Run tasks through the C restatement -> (results, cigar_pool, seconds)."""
    return _run(ensure_oracle(), "orc_run_batch", tasks, nthreads or os.cpu_count() or 1)


def ref_run(tasks, nthreads=None):
    """Run tasks through the unmodified reference ksw.c -> (results, cigar_pool, seconds)."""
    return _run(REF_SO, "ref_run_batch", tasks, nthreads or os.cpu_count() or 1)


CMP_GLOBAL = ("score", "n_cigar")
CMP_EXT_CIGAR = ("score", "qle", "tle", "n_cigar", "m_cigar")
CMP_EXT_SCORE = ("score", "qle", "tle", "gtle", "gscore", "max_off")


def compare(tasks, ra, ca, rb, cb, what="", check_cells=False):
    """Field-by-field + CIGAR-word comparison; returns list of mismatch strings (empty = identical)."""
    bad = []
    kind, flags = tasks["kind"], tasks["flags"]

    def chk(mask, fields):
        for f in fields:
            d = np.nonzero(mask & (ra[f] != rb[f]))[0]
            for i in d[:5]:
                bad.append(f"{what} task {i} kind={kind[i]} flags={flags[i]} qlen={tasks['qlen'][i]} "
                           f"tlen={tasks['tlen'][i]} w={tasks['w'][i]} h0={tasks['h0'][i]}: {f} {ra[f][i]} != {rb[f][i]}")
    chk(kind == 0, CMP_GLOBAL)
    chk((kind == 1) & (flags & 1 == 1), CMP_EXT_CIGAR)
    chk((kind == 1) & (flags & 1 == 0), CMP_EXT_SCORE)
    if check_cells:
        chk(np.ones(len(tasks), bool), ("cells",))
    if not bad:
        # CIGAR words, task by task (pools may be laid out in different orders)
        n = ra["n_cigar"].astype(np.int64)
        tot = int(n.sum())
        if tot:
            ia = np.repeat(ra["cigar_off"], n) + (np.arange(tot) - np.repeat(np.cumsum(n) - n, n))
            ib = np.repeat(rb["cigar_off"], n) + (np.arange(tot) - np.repeat(np.cumsum(n) - n, n))
            neq = ca[ia] != cb[ib]
            if neq.any():
                owner = np.repeat(np.arange(len(tasks)), n)
                for i in np.unique(owner[neq])[:5]:
                    a = ca[ra["cigar_off"][i]: ra["cigar_off"][i] + n[i]]
                    b = cb[rb["cigar_off"][i]: rb["cigar_off"][i] + n[i]]
                    bad.append(f"{what} task {i} kind={kind[i]} qlen={tasks['qlen'][i]} tlen={tasks['tlen'][i]} "
                               f"w={tasks['w'][i]}: CIGAR {fmt_cigar(a)} != {fmt_cigar(b)}")
    return bad


def fmt_cigar(words):
    return "".join(f"{int(x) >> 4}{'MIDNSHP=XB'[int(x) & 15]}" for x in words)


def cigar_digest(res, cig):
    """Order-independent digest of all CIGARs: sha1 over per-task words in task order."""
    h = hashlib.sha1()
    n = res["n_cigar"].astype(np.int64)
    tot = int(n.sum())
    if tot:
        idx = np.repeat(res["cigar_off"], n) + (np.arange(tot) - np.repeat(np.cumsum(n) - n, n))
        h.update(np.ascontiguousarray(cig[idx], dtype="<i4").tobytes())
    return h.hexdigest()


def inputs_digest(tasks):
    """sha1 of every task's parameters and sequence bytes (detects generator drift)."""
    h = hashlib.sha1()
    for name in ("kind", "flags", "qlen", "tlen", "w", "h0", "o_del", "e_del", "o_ins", "e_ins", "end_bonus", "zdrop"):
        h.update(np.ascontiguousarray(tasks[name]).tobytes())
    for t in tasks:
        h.update(C.string_at(int(t["query"]), int(t["qlen"])))
        h.update(C.string_at(int(t["target"]), int(t["tlen"])))
    return h.hexdigest()
