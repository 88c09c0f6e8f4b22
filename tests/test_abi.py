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
    declared = set(re.findall(r"\b((?:ksw|lb2|sw)_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


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
