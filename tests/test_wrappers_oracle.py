"""CPU suite: the oracle's restatements of the host wrappers -- orc_extend_c / orc_extend_r /
orc_mid_fix / orc_bi_extend (oracle/dp_oracle.c) -- pinned against the unmodified reference's
ksw_extend_c / ksw_extend_r / sw_mid_fix / ksw_bi_extend (src/ksw.c:809-926, oracle/_ref/libksw_ref.so)
on seeded pairs that reach every exit of ksw_bi_extend, including the float comparison of :881 with
aln_mode & 2.  Only possible where the reference was compiled (this container)."""
import collections

import pytest

import _oracle
import _wrappers

pytestmark = pytest.mark.skipif(not _oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("seed,n", [(901, 6000), (902, 6000)])
def test_oracle_wrappers_match_reference(seed, n):
    pairs = _wrappers.gen_pairs(n, seed)
    ref, orc = _wrappers.Impl("ref"), _wrappers.Impl("oracle")
    a, b = _wrappers.run_all(ref, pairs), _wrappers.run_all(orc, pairs)
    exits = collections.Counter()
    for k, (ra, rb, p) in enumerate(zip(a, b, pairs)):
        assert ra == rb, f"pair {k} (qlen {len(p[0])}, tlen {len(p[1])}): reference {ra} != oracle {rb}"
        exits[_wrappers.classify_exit(ra[0], ra[1], p[0], p[1], p[4])] += 1
    # the generator must really reach all six exits (five of ksw_bi_extend, sw_mid_fix's two branches)
    for e in ("left_end", "left_global", "right_end", "right_global", "mid_clip", "mid_global"):
        assert exits[e] >= 10, exits


def test_zero_cell_calls_are_answered_without_a_gpu_like_the_reference():
    """An empty query or target needs no DP cell: the drop-in symbols answer in closed form (ksw_dropin.cu:
    zero_cell_task), so these calls work on a box without a GPU and must equal the reference's results."""
    import ctypes as C
    import numpy as np
    import lamsa_b200
    rng = np.random.default_rng(77)
    ref, lib = _wrappers.Impl("ref"), _wrappers.Impl("gpu")
    n = 0
    for it in range(400):
        AP = _wrappers.make_para(rng)
        AP.end_bonus = int(rng.choice([0, 5, 5, 20]))            # 20 > oe_del: the (max_ie, qlen-1) end point
        ql, tl = [(0, 0), (0, int(rng.integers(1, 300))), (int(rng.integers(1, 300)), 0)][it % 3]
        q = rng.integers(0, 4, size=ql, dtype=np.uint8); t = rng.integers(0, 4, size=tl, dtype=np.uint8)
        w, h0 = int(rng.integers(1, 100)), int(rng.choice([1, 3, 8, 50, 100]))
        for rev in (False, True):
            assert lib.extend(rev, q, t, w, h0, AP) == ref.extend(rev, q, t, w, h0, AP), (ql, tl, w, h0, rev)
        # ksw_global2 through the library's Python mirror and the reference through ctypes
        s, cg = lamsa_b200.ksw_global2(ql, np.concatenate((q, np.zeros(8, np.uint8))), tl, np.concatenate((t, np.zeros(8, np.uint8))),
                                       5, ref.mat, AP.del_gapo, AP.del_gape, AP.ins_gapo, AP.ins_gape, w)
        nc, cp = C.c_int(0), C.POINTER(C.c_int32)()
        qb, qp = ref._buf(q); tb, tp = ref._buf(t)
        f = ref.lib.ksw_global2
        f.argtypes = [C.c_int, _wrappers.u8p, C.c_int, _wrappers.u8p, C.c_int, _wrappers.i8p] + [C.c_int] * 5 + [_wrappers.ip, _wrappers.cpp]
        rs = f(ql, qp, tl, tp, 5, ref.mat.ctypes.data_as(_wrappers.i8p), AP.del_gapo, AP.del_gape, AP.ins_gapo, AP.ins_gape, w,
               C.byref(nc), C.byref(cp))
        assert (s, cg) == (rs, ref._take(cp, nc.value)), (ql, tl)
        if ql == 0 or tl == 0:
            # every stage of ksw_bi_extend sees an empty side too
            assert lib.bi_extend(q, t, h0, h0, AP) == ref.bi_extend(q, t, h0, h0, AP), (ql, tl)
        n += 1
    assert n == 400
