"""CPU suite: the C-ABI library loads, exports every symbol include/lamsa_b200.h
declares, its struct layouts match the reference headers, and it fails loudly
(no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _oracle
import lamsa_b200
from lamsa_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# SURVEY.md appendix B (printed from the reference headers with gcc 13, x86-64)
REF_OFFSETS = dict(n_thread=0, seed_len=4, seed_step=8, seed_inv=12, per_aln_m=16, first_loci_thd=20,
                   SV_len_thd=24, ske_max=28, ovlp_rat=32, split_len=52, split_pen=56, res_mul_max=60,
                   hash_len=64, hash_key_len=68, hash_step=72, hash_size=76, outp=88, match_dis=96,
                   mismatch_thd=100, frag_score_table=112, ins_gapo=120, ins_gape=124, del_gapo=128,
                   del_gape=132, ins_ext_o=136, ins_ext_e=140, del_ext_o=144, del_ext_e=148, match=152,
                   mis=156, sc_mat=160, band_w=188, end_bonus=192, zdrop=196, ed_rate=200, id_rate=208,
                   read_type=216, aln_mode=220)


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(lamsa_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return lamsa_b200.load_library()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "lamsa_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:ksw|lb2|sw|frag_line|node|build_node)_[A-Za-z0-9_]+|cover_rate|init_hash|hash_split_map)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    abi = open(os.path.join(ROOT, "lamsa_b200", "csrc", "ref_abi.h")).read()
    declared_abi = set(re.findall(r"\b((?:heap|node)_[a-z_]+)\s*\(lb2_ref_node_score", abi))
    assert declared_abi == set(_lib.EXPORTS_REF_ABI), declared_abi ^ set(_lib.EXPORTS_REF_ABI)
    for name in declared | declared_abi:
        assert hasattr(lib, name), name


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liblamsa_ref.so")),
                    reason="oracle/_ref/liblamsa_ref.so not built (needs /root/reference)")
def test_sdp_struct_layouts_match_reference_headers(lib):
    """map_t/map_msg/frag_msg/frag_aln_msg/aln_reg/reg_t/kseq_t/node_score as restated in ref_abi.h"""
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblamsa_ref.so"))
    a, b = (C.c_int * 64)(), (C.c_int * 64)()
    na, nb = ref.ref_sdp_offsets(a), lib.lb2_ref_abi_offsets(b)
    assert na == nb and list(a[:na]) == list(b[:nb])
    na, nb = ref.ref_sdp_sizes(a), lib.lb2_ref_abi_sizes(b)
    assert na == nb and list(a[:na]) == list(b[:nb])


def test_sdp_struct_layouts_match_survey_appendix_b(lib):
    b = (C.c_int * 64)()
    n = lib.lb2_ref_abi_sizes(b)
    # map_t, map_msg, frag_dp_node, frag_msg, frag_aln_msg, lamsa_aln_per_para, aln_reg, reg_t
    assert list(b[:8]) == [1064, 32, 80, 32, 88, 12, 24, 40] and b[10] == 8 and b[11] == 40


class LineNode(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int)]


class NodeScore(C.Structure):
    _fields_ = [("node", C.POINTER(LineNode)), ("score", C.POINTER(C.c_int)), ("NM", C.POINTER(C.c_int)),
                ("min_score_thd", C.c_int), ("max_n", C.c_int), ("node_n", C.c_int)]


def _bind_heap(l):
    l.node_init_score.restype = C.POINTER(NodeScore); l.node_init_score.argtypes = [C.c_int]
    l.node_free_score.argtypes = [C.POINTER(NodeScore)]; l.node_free_score.restype = None
    l.heap_add_node.argtypes = [C.POINTER(NodeScore), LineNode, C.c_int, C.c_int]
    l.node_heap_update_min.argtypes = [C.POINTER(NodeScore), LineNode, C.c_int, C.c_int]
    l.node_pop.argtypes = [C.POINTER(NodeScore), C.POINTER(C.c_int), C.POINTER(C.c_int)]; l.node_pop.restype = LineNode
    l.node_heap_extract_max.argtypes = [C.POINTER(NodeScore), C.POINTER(C.c_int)]; l.node_heap_extract_max.restype = LineNode
    l.node_heap_extract_minpos.argtypes = [C.POINTER(NodeScore)]; l.node_heap_extract_minpos.restype = LineNode
    for f in ("build_node_max_heap", "build_node_min_heap", "build_node_minpos_heap"):
        getattr(l, f).argtypes = [C.POINTER(NodeScore)]; getattr(l, f).restype = None
    l.cover_rate.argtypes = [C.c_int] * 4; l.cover_rate.restype = C.c_float
    return l


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liblamsa_ref.so")),
                    reason="oracle/_ref/liblamsa_ref.so not built (needs /root/reference)")
def test_node_score_helpers_match_reference(lib):
    """the host-side node_score helpers other reference files bind (bwt_aln.c, lamsa_aln.c): same
    array contents and return values as src/lamsa_heap.c / src/lamsa_dp_con.c:29-67 on random traces"""
    ref = _bind_heap(C.CDLL(os.path.join(ROOT, "oracle", "_ref", "liblamsa_ref.so")))
    mine = _bind_heap(lib)
    rng = np.random.default_rng(5)

    def state(ns):
        n = ns.contents.node_n
        return [(ns.contents.node[i].x, ns.contents.node[i].y, ns.contents.score[i], ns.contents.NM[i]) for i in range(n)]

    for trial in range(200):
        cap = int(rng.integers(1, 12))
        a, b = ref.node_init_score(cap), mine.node_init_score(cap)
        ops = rng.integers(0, 100, size=int(rng.integers(1, 40)))
        for op in ops:
            x, y, sc, nm = (int(v) for v in rng.integers(0, 8, size=4))
            if op < 60:
                ra, rb = ref.heap_add_node(a, LineNode(x, y), sc, nm), mine.heap_add_node(b, LineNode(x, y), sc, nm)
            elif op < 70:
                s1, s2, n1, n2 = C.c_int(-9), C.c_int(-9), C.c_int(-9), C.c_int(-9)
                pa, pb = ref.node_pop(a, C.byref(s1), C.byref(n1)), mine.node_pop(b, C.byref(s2), C.byref(n2))
                ra, rb = (pa.x, pa.y, s1.value, n1.value), (pb.x, pb.y, s2.value, n2.value)
            elif op < 80:
                ref.build_node_minpos_heap(a); mine.build_node_minpos_heap(b)
                pa, pb = ref.node_heap_extract_minpos(a), mine.node_heap_extract_minpos(b)
                ra, rb = (pa.x, pa.y), (pb.x, pb.y)
            elif op < 90:
                ref.build_node_max_heap(a); mine.build_node_max_heap(b)
                s1, s2 = C.c_int(-9), C.c_int(-9)
                pa, pb = ref.node_heap_extract_max(a, C.byref(s1)), mine.node_heap_extract_max(b, C.byref(s2))
                ra, rb = (pa.x, pa.y, s1.value), (pb.x, pb.y, s2.value)
            else:
                ref.build_node_min_heap(a); mine.build_node_min_heap(b)
                ra = rb = 0
            assert ra == rb, (trial, op)
            assert state(a) == state(b), (trial, op)
        ref.node_free_score(a); mine.node_free_score(b)
    for _ in range(200):
        v = [int(t) for t in rng.integers(1, 300, size=4)]
        s1, e1, s2, e2 = v[0], v[0] + v[1], v[2], v[2] + v[3]
        assert ref.cover_rate(s1, e1, s2, e2) == mine.cover_rate(s1, e1, s2, e2)


def test_para_layout_matches_survey_offsets():
    assert C.sizeof(lamsa_b200.AlnPara) == 224
    for f, off in REF_OFFSETS.items():
        assert getattr(lamsa_b200.AlnPara, f).offset == off, f


@pytest.mark.skipif(not _oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_para_layout_matches_reference_header():
    ref = C.CDLL(_oracle.REF_SO)
    assert ref.ref_para_sizeof() == C.sizeof(lamsa_b200.AlnPara)
    buf = (C.c_int * 64)()
    n = ref.ref_para_offsets(buf)
    assert n == len(_lib.PARA_FIELDS)
    for name, off in zip(_lib.PARA_FIELDS, buf[:n]):
        assert getattr(lamsa_b200.AlnPara, name).offset == off, name


def test_task_and_result_struct_sizes():
    # must match the C structs in include/lamsa_b200.h (checked by the oracle's batch driver too)
    assert _lib.TASK_DTYPE.itemsize == 88 and _lib.RESULT_DTYPE.itemsize == 48
    assert _lib.TASK_DTYPE.fields["mat"][1] == 72 and _lib.RESULT_DTYPE.fields["cells"][1] == 40


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.lb2_ctx_create(0, C.byref(h)) != 0
    assert b"no CPU path" in lib.lb2_last_error() or b"CUDA" in lib.lb2_last_error()
    with pytest.raises(RuntimeError):
        lamsa_b200.Context(0)


def test_pool_pack_copies_sequences_in_task_order(lib):
    """lb2_pool_pack (host only): the sequences of scattered task records end up in ONE pool, task by task, and the
    records point into it (the layout under which a chunk of lb2_dp_run_pool uploads one tight range)."""
    import numpy as np
    from lamsa_b200 import workload
    tasks, keep = workload.gen_microbench(3000, seed=9, qmin=0, qmax=120)
    ptasks, pool = workload.pool_tasks(tasks, keep)
    base = pool.ctypes.data
    q_off = (ptasks["query"] - np.uint64(base)).astype(np.int64)
    t_off = (ptasks["target"] - np.uint64(base)).astype(np.int64)
    assert (q_off >= 0).all() and (t_off + ptasks["tlen"] <= pool.nbytes).all()
    assert (np.diff(q_off) > 0).all() and (t_off == q_off + ptasks["qlen"]).all()
    import ctypes as C
    for i in (0, 1, 17, 2999):
        for f, n in (("query", "qlen"), ("target", "tlen")):
            a = np.ctypeslib.as_array(C.cast(int(tasks[f][i]), C.POINTER(C.c_uint8)), shape=(max(int(tasks[n][i]), 1),))[: int(tasks[n][i])]
            b = np.ctypeslib.as_array(C.cast(int(ptasks[f][i]), C.POINTER(C.c_uint8)), shape=(max(int(tasks[n][i]), 1),))[: int(tasks[n][i])]
            assert (a == b).all()
    for f in ("kind", "qlen", "tlen", "w", "h0", "mat"):
        assert (ptasks[f] == tasks[f]).all()
