"""GPU suite for the local Smith-Waterman entry points (include/lamsa_b200.h section 7): the drop-in ksw_align2
(sw_local.cuh through lb2_sw_run) against the oracle, which is pinned against the unmodified reference
(tests/test_sw_oracle.py), and against golden results of the reference itself.  Nothing here reads /root/reference."""
import ctypes as C
import os

import numpy as np
import pytest

import _sw

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sw_golden.npz")


def test_gpu_ksw_align2_matches_reference_golden(ctx):
    g = np.load(GOLDEN)
    cases = _sw.gen_cases(int(g["n"]), int(g["seed"]))
    for k, c in enumerate(cases):
        assert _sw.gpu_align2(c) == tuple(int(v) for v in g["res"][k]), f"case {k} (xtra {c['xtra']:#x})"


def test_gpu_ksw_align2_matches_oracle(ctx):
    cases = _sw.gen_cases(500, 41, qmax=700, tmax=2500)
    for k, c in enumerate(cases):
        got, want = _sw.gpu_align2(c), _sw.oracle_align2(c)
        assert got == want, f"case {k} (xtra {c['xtra']:#x}, qlen {len(c['q'])}, tlen {len(c['t'])}): {got} vs {want}"


def test_gpu_sw_batch_and_profile_reuse(ctx):
    """lb2_sw_run on many pairs at once; ksw_qinit + ksw_u8 / ksw_i16 with one profile against several targets"""
    from lamsa_b200 import load_library
    lib = load_library()

    class Task(C.Structure):
        _fields_ = [("query", C.c_void_p), ("qlen", C.c_int32), ("target", C.c_void_p), ("tlen", C.c_int32), ("m", C.c_int32), ("mat", C.c_void_p),
                    ("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("xtra", C.c_int32), ("size", C.c_int32)]
    cases = [c for c in _sw.gen_cases(300, 43) if not c["xtra"] & _sw.XSTART]
    tasks = (Task * len(cases))(); keep = []
    for k, c in enumerate(cases):
        q = np.ascontiguousarray(c["q"]); t = np.ascontiguousarray(c["t"]); m = np.ascontiguousarray(c["mat"]); keep += [q, t, m]
        tasks[k] = Task(q.ctypes.data, len(q), t.ctypes.data, len(t), c["m"], m.ctypes.data, c["o_del"], c["e_del"], c["o_ins"], c["e_ins"],
                        c["xtra"], 1 if c["xtra"] & _sw.XBYTE else 2)
    out = np.zeros((len(cases), 7), np.int32)
    lib.lb2_sw_run.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    assert lib.lb2_sw_run(ctx.handle, len(cases), tasks, out.ctypes.data) == 0, lib.lb2_last_error()
    for k, c in enumerate(cases):
        assert tuple(int(v) for v in out[k]) == _sw.oracle_align2(c), f"batch case {k}"
    # one profile, several targets, both widths
    lib.ksw_qinit.restype = C.c_void_p
    lib.ksw_qinit.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    for fn in (lib.ksw_u8, lib.ksw_i16):
        fn.restype = _sw.Kswr
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
    c0 = cases[0]
    q = np.ascontiguousarray(c0["q"]); m = np.ascontiguousarray(c0["mat"])
    for size, fn in ((1, lib.ksw_u8), (2, lib.ksw_i16)):
        prof = lib.ksw_qinit(size, len(q), q.ctypes.data, 5, m.ctypes.data)
        for c in cases[:12]:
            t = np.ascontiguousarray(c["t"])
            r = fn(prof, len(t), t.ctypes.data, 5, 2, 5, 2, 0).tup()
            want = _sw.oracle_align2(dict(c0, t=c["t"], o_del=5, e_del=2, o_ins=5, e_ins=2, xtra=_sw.XBYTE if size == 1 else 0))
            assert r == want
        libc.free(prof)


def test_gpu_sw_edge_cases(ctx):
    """one-base sequences, an empty target, a query longer than the target, protein-sized alphabets"""
    rng = np.random.default_rng(6)
    mat5 = np.full((5, 5), -3, np.int8); np.fill_diagonal(mat5, 1)
    base = dict(m=5, mat=mat5.reshape(-1), o_del=5, e_del=2, o_ins=5, e_ins=2)
    cases = []
    for xtra in (0, _sw.XBYTE, _sw.XSTART, _sw.XBYTE | _sw.XSTART | _sw.XSUBO | 3):
        cases += [dict(base, q=np.array([1], np.uint8), t=np.array([1], np.uint8), xtra=xtra),
                  dict(base, q=np.array([1], np.uint8), t=np.array([2], np.uint8), xtra=xtra),
                  dict(base, q=rng.integers(0, 4, size=40, dtype=np.uint8), t=np.zeros(0, np.uint8), xtra=xtra),
                  dict(base, q=rng.integers(0, 4, size=300, dtype=np.uint8), t=rng.integers(0, 4, size=7, dtype=np.uint8), xtra=xtra)]
    m20 = rng.integers(-4, 6, size=(20, 20)).astype(np.int8); m20 = np.minimum(m20, m20.T); np.fill_diagonal(m20, 5)
    for xtra in (0, _sw.XBYTE | _sw.XSTART):
        t = rng.integers(0, 20, size=400, dtype=np.uint8)
        cases.append(dict(m=20, mat=m20.reshape(-1), o_del=10, e_del=1, o_ins=10, e_ins=1, q=t[100:180].copy(), t=t, xtra=xtra))
    for k, c in enumerate(cases):
        assert _sw.gpu_align2(c) == _sw.oracle_align2(c), f"edge case {k}"
