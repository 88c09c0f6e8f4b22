"""Test helpers for the local split mapping (reference src/split_mapping.c): SV-shaped (window, read) pairs, ctypes
bindings of the CPU checkers (oracle/build/libhash_oracle.so = this repo's restatement of the seed-and-chain half,
oracle/_ref/liblamsa_ref.so = the unmodified reference, only where it was built).  Test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "build", "libhash_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblamsa_ref.so")
_libs = {}


def have_ref():
    return os.path.exists(REF_SO)


def _lib(path):
    if path not in _libs:
        if path == ORACLE_SO and not os.path.exists(path):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "hash"])
        lib = C.CDLL(path)
        sig = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        for name in ("orc_hash_line", "ref_hash_line"):
            if hasattr(lib, name):
                getattr(lib, name).argtypes = sig
        _libs[path] = lib
    return _libs[path]


def _line(fn, c):
    cap = max(16, len(c["read"]) // max(1, c["hash_step"]) + 8)
    a, b, f = (np.zeros(cap, np.int32) for _ in range(3))
    ref = np.ascontiguousarray(c["ref"], np.uint8); read = np.ascontiguousarray(c["read"], np.uint8)
    m = fn(ref.ctypes.data, len(ref), read.ctypes.data, len(read), c["ref_offset"], c["hash_len"], c["hash_step"],
           c["split_len"], c["head"], c["tail"], a.ctypes.data, b.ctypes.data, f.ctypes.data, cap)
    assert m >= 0, "line buffer too small"
    return np.stack((a[:m], b[:m], f[:m]), axis=1)


def oracle_line(c):
    return _line(_lib(ORACLE_SO).orc_hash_line, c)


def ref_line(c):
    return _line(_lib(REF_SO).ref_hash_line, c)


def gen_cases(n, seed, max_len=3000):
    """(reference window, read) pairs the way split mapping meets them: a read piece against the window it came from,
    with a deletion / insertion / duplication / nothing in the middle, 0-15 % errors, low-complexity stretches (k-mers
    with many hits, some beyond the 50-hit cap), reads with N; all head / tail combinations; both parameter presets."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(n):
        preset = k % 3
        hash_len, hash_step = ((10, 10), (8, 4), (8, 4))[preset]
        L = int(rng.integers(hash_len, max_len)) if k % 11 else int(rng.integers(hash_len, 3 * hash_len))
        base = rng.integers(0, 4, size=L + 2400, dtype=np.uint8)
        if k % 4 == 1:          # tandem / low-complexity stretch
            unit = rng.integers(0, 4, size=int(rng.integers(1, 7)), dtype=np.uint8)
            a = int(rng.integers(0, max(1, L // 2)))
            ln = int(rng.integers(20, 900))
            base[a:a + ln] = np.resize(unit, ln)[: len(base[a:a + ln])]
        ref = base[:L].copy()
        kind = k % 5
        cut = int(rng.integers(1, max(2, L - 1)))
        sv = int(rng.integers(1, 2000))
        if kind == 0:
            read = ref.copy()
        elif kind == 1:         # deletion in the read
            read = np.concatenate((ref[:cut], ref[min(L, cut + sv):]))
        elif kind == 2:         # insertion in the read
            read = np.concatenate((ref[:cut], rng.integers(0, 4, size=sv, dtype=np.uint8), ref[cut:]))
        elif kind == 3:         # duplication in the read
            d = min(sv, cut)
            read = np.concatenate((ref[:cut], ref[cut - d:cut], ref[cut:]))
        else:                   # unrelated flank
            read = np.concatenate((ref[:cut], rng.integers(0, 4, size=min(sv, 300), dtype=np.uint8)))
        err = float(rng.choice([0.0, 0.02, 0.05, 0.15]))
        if err > 0 and len(read):
            keep = rng.random(len(read)) >= err / 3
            read = read[keep]
            sub = rng.random(len(read)) < err / 3
            read[sub] = (read[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
            ins = np.flatnonzero(rng.random(len(read)) < err / 3)
            read = np.insert(read, ins, rng.integers(0, 4, size=len(ins), dtype=np.uint8))
        if k % 7 == 3 and len(read) > 5:
            read[rng.integers(0, len(read), size=3)] = 4
        if len(read) < hash_len:
            read = np.concatenate((read, rng.integers(0, 4, size=hash_len - len(read), dtype=np.uint8)))
        head, tail = ((1, 1), (1, 1), (1, 0), (0, 1))[k % 4]
        ref_offset = 0
        if kind == 2 and k % 2 == 0:
            ref_offset = int(rng.integers(1, 50))            # the overlapped-duplication branch of hash_main_dis
        out.append(dict(ref=ref, read=read.astype(np.uint8), ref_offset=ref_offset, hash_len=hash_len, hash_step=hash_step,
                        split_len=100 if preset != 2 else 50, head=head, tail=tail))
    return out


# ---- the whole hash_split_map (index + line + DP stitching): reference-generated golden CIGARs vs the drop-in ----
def make_ap(c):
    """lamsa_aln_para as `lamsa aln` sets it for the case's preset (src/lamsa_aln.c:1281-1420, src/lamsa_aln.h:17-77)."""
    from lamsa_b200 import AlnPara, default_matrix
    AP = AlnPara()
    pacbio = c["hash_len"] == 8
    AP.hash_len, AP.hash_step, AP.hash_key_len, AP.hash_size = c["hash_len"], c["hash_step"], 2, 16
    AP.split_len, AP.split_pen = c["split_len"], 10
    if pacbio:
        AP.match, AP.mis = 1, 1
        AP.ins_gapo, AP.ins_gape, AP.del_gapo, AP.del_gape = 1, 1, 1, 1
        AP.ins_ext_o, AP.ins_ext_e, AP.del_ext_o, AP.del_ext_e = 2, 1, 2, 1
        AP.band_w, AP.end_bonus, AP.id_rate, AP.aln_mode = 200, 0, 0.3, 3
    else:
        AP.match, AP.mis = 1, 3
        AP.ins_gapo, AP.ins_gape, AP.del_gapo, AP.del_gape = 5, 2, 5, 2
        AP.ins_ext_o, AP.ins_ext_e, AP.del_ext_o, AP.del_ext_e = 5, 2, 5, 2
        AP.band_w, AP.end_bonus, AP.id_rate, AP.aln_mode = 10, 5, 0.04, 0
    AP.zdrop = 100
    m = default_matrix(AP.match, AP.mis)
    for k in range(25):
        AP.sc_mat[k] = int(m[k])
    return AP


def ref_split_map(c):
    """reference hash_split_map -> (CIGAR words, return value)"""
    lib = _lib(REF_SO)
    AP = make_ap(c)
    ref = np.ascontiguousarray(c["ref"], np.uint8); read = np.ascontiguousarray(c["read"], np.uint8)
    cap = 4 * (len(ref) + len(read)) + 64
    out = np.zeros(cap, np.int32); res = C.c_int()
    lib.ref_hash_split_map.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    n = lib.ref_hash_split_map(ref.ctypes.data, len(ref), c["ref_offset"], read.ctypes.data, len(read), C.byref(AP), c["head"], c["tail"],
                               out.ctypes.data, cap, C.byref(res))
    assert n >= 0
    return out[:n].copy(), res.value


def gpu_split_map(c):
    """this library's drop-in hash_split_map (GPU line + GPU DP) -> (CIGAR words, return value)"""
    from lamsa_b200 import load_library
    lib = load_library()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p; libc.malloc.argtypes = [C.c_size_t]; libc.free.argtypes = [C.c_void_p]
    AP = make_ap(c)
    ref = np.ascontiguousarray(c["ref"], np.uint8); read = np.ascontiguousarray(c["read"], np.uint8)
    cig = C.c_void_p(libc.malloc(4 * 100)); n = C.c_int(0); m = C.c_int(100)
    lib.hash_split_map.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    res = lib.hash_split_map(C.byref(cig), C.byref(n), C.byref(m), ref.ctypes.data, len(ref), c["ref_offset"], read.ctypes.data, len(read),
                             C.byref(AP), None, None, None, None, c["head"], c["tail"])
    out = np.ctypeslib.as_array(C.cast(cig, C.POINTER(C.c_int32)), shape=(max(n.value, 1),))[: n.value].copy()
    libc.free(cig)
    return out, res


def gpu_lines(ctx, cases):
    """lb2_hash_line_run on all cases as ONE batch -> list of (m_len, 3) arrays"""
    from lamsa_b200._lib import HashTask
    n = len(cases)
    tasks = (HashTask * n)()
    keep = []
    for k, c in enumerate(cases):
        ref = np.ascontiguousarray(c["ref"], np.uint8); read = np.ascontiguousarray(c["read"], np.uint8)
        cap = max(1, (len(read) - c["hash_len"]) // c["hash_step"] + 1) if len(read) >= c["hash_len"] else 1
        line = np.zeros(3 * cap, np.int32)
        keep += [ref, read, line]
        t = tasks[k]
        t.ref, t.ref_len, t.read, t.read_len = ref.ctypes.data, len(ref), read.ctypes.data, len(read)
        t.ref_offset, t.hash_len, t.hash_step, t.split_len = c["ref_offset"], c["hash_len"], c["hash_step"], c["split_len"]
        t.head, t.tail, t.line, t.line_cap = c["head"], c["tail"], line.ctypes.data, cap
    if ctx.lib.lb2_hash_line_run(ctx.handle, n, tasks):
        raise RuntimeError(ctx.lib.lb2_last_error().decode())
    return [keep[3 * k + 2][: 3 * tasks[k].m_len].reshape(-1, 3).copy() for k in range(n)], [tasks[k].n_hits for k in range(n)]
