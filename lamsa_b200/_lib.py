"""ctypes view of include/lamsa_b200.h (structs, constants, library loading)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LB2_LIB_PATH") or os.path.join(HERE, "liblamsa_b200.so")   # env override: kernel experiments

KIND_GLOBAL, KIND_EXTEND = 0, 1
FLAG_CIGAR, FLAG_TARGET_PAC, FLAG_TARGET_REV = 1, 2, 4


class LibraryMissing(RuntimeError):
    pass


# lb2_task / lb2_result as numpy structured dtypes (C layout, x86-64)
TASK_DTYPE = np.dtype([
    ("kind", "<i4"), ("flags", "<i4"), ("qlen", "<i4"), ("tlen", "<i4"),
    ("query", "<u8"), ("target", "<u8"),
    ("w", "<i4"), ("h0", "<i4"), ("o_del", "<i4"), ("e_del", "<i4"), ("o_ins", "<i4"), ("e_ins", "<i4"),
    ("end_bonus", "<i4"), ("zdrop", "<i4"), ("m", "<i4"), ("_pad", "<i4"), ("mat", "<u8"),
    ("target_pac", "<i8"),
], align=True)
assert TASK_DTYPE.itemsize == 88

RESULT_DTYPE = np.dtype([
    ("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"), ("gtle", "<i4"), ("gscore", "<i4"),
    ("max_off", "<i4"), ("n_cigar", "<i4"), ("m_cigar", "<i4"), ("cigar_off", "<i8"), ("cells", "<i8"),
], align=True)
assert RESULT_DTYPE.itemsize == 48


class AlnPara(C.Structure):
    """Mirror of `lamsa_aln_para` (reference src/lamsa_aln.h:384-432), same field order."""
    _fields_ = [
        ("n_thread", C.c_int),
        ("seed_len", C.c_int), ("seed_step", C.c_int), ("seed_inv", C.c_int),
        ("per_aln_m", C.c_int), ("first_loci_thd", C.c_int),
        ("SV_len_thd", C.c_int), ("ske_max", C.c_int), ("ovlp_rat", C.c_float),
        ("bwt_seed_len", C.c_int), ("bwt_max_len", C.c_int), ("bwt_min_len", C.c_int),
        ("fastest", C.c_int),
        ("split_len", C.c_int), ("split_pen", C.c_int), ("res_mul_max", C.c_int),
        ("hash_len", C.c_int), ("hash_key_len", C.c_int), ("hash_step", C.c_int), ("hash_size", C.c_int),
        ("supp_soft", C.c_uint8), ("comm", C.c_uint8),
        ("outp", C.c_void_p),
        ("match_dis", C.c_int), ("mismatch_thd", C.c_int),
        ("del_thd", C.c_int), ("ins_thd", C.c_int),
        ("frag_score_table", C.c_void_p),
        ("ins_gapo", C.c_int), ("ins_gape", C.c_int), ("del_gapo", C.c_int), ("del_gape", C.c_int),
        ("ins_ext_o", C.c_int), ("ins_ext_e", C.c_int), ("del_ext_o", C.c_int), ("del_ext_e", C.c_int),
        ("match", C.c_int), ("mis", C.c_int),
        ("sc_mat", C.c_int8 * 25),
        ("band_w", C.c_int), ("end_bonus", C.c_int), ("zdrop", C.c_int),
        ("ed_rate", C.c_float), ("mis_rate", C.c_float), ("id_rate", C.c_float), ("mat_rate", C.c_float),
        ("read_type", C.c_int),
        ("aln_mode", C.c_uint8),
    ]


class AuxTask(C.Structure):
    """lb2_aux_task"""
    _fields_ = [("cigar", C.c_void_p), ("n_cigar", C.c_int32), ("read_len", C.c_int32), ("read", C.c_void_p), ("ref_pac", C.c_int64)]


class HashTask(C.Structure):
    """lb2_hash_task"""
    _fields_ = [("ref", C.c_void_p), ("ref_len", C.c_int32), ("read", C.c_void_p), ("read_len", C.c_int32), ("ref_offset", C.c_int32),
                ("hash_len", C.c_int32), ("hash_step", C.c_int32), ("split_len", C.c_int32), ("head", C.c_int32), ("tail", C.c_int32),
                ("line", C.c_void_p), ("line_cap", C.c_int32), ("m_len", C.c_int32), ("n_hits", C.c_int32)]


class ClassStat(C.Structure):
    """lb2_class_stat"""
    _fields_ = [("class_id", C.c_int32), ("kind", C.c_int32), ("variant", C.c_int32), ("window_slots", C.c_int32),
                ("tasks", C.c_int64), ("cells", C.c_int64), ("ms", C.c_float), ("kernel", C.c_char * 44)]


# order of oracle/ref_shim.c:ref_para_offsets
PARA_FIELDS = [
    "n_thread", "seed_len", "seed_step", "seed_inv", "per_aln_m", "first_loci_thd", "SV_len_thd", "ske_max",
    "ovlp_rat", "split_len", "split_pen", "res_mul_max", "match_dis", "mismatch_thd", "frag_score_table",
    "ins_gapo", "ins_gape", "del_gapo", "del_gape", "ins_ext_o", "ins_ext_e", "del_ext_o", "del_ext_e",
    "match", "mis", "sc_mat", "band_w", "end_bonus", "zdrop", "ed_rate", "mis_rate", "id_rate", "mat_rate",
    "read_type", "aln_mode",
]

# every symbol include/lamsa_b200.h declares
EXPORTS = [
    "ksw_global2", "ksw_global", "ksw_extend2", "ksw_extend", "ksw_extend_core", "ksw_extend_c",
    "ksw_extend_r", "ksw_bi_extend", "sw_mid_fix", "ksw_qinit", "ksw_u8", "ksw_i16", "ksw_align2", "ksw_align", "lb2_sw_run",
    "lb2_ctx_create", "lb2_ctx_destroy", "lb2_last_error", "lb2_ctx_set_scratch_limit", "lb2_ctx_set_reference", "lb2_ctx_last_run_stats", "lb2_ctx_last_run_kernel_ms", "lb2_ctx_set_chunk_tasks", "lb2_dp_run", "lb2_dp_run_pool", "lb2_pool_pack", "lb2_host_alloc", "lb2_host_free",
    "lb2_batch_create", "lb2_batch_create_pool", "lb2_batch_upload", "lb2_batch_compute", "lb2_batch_compute_async", "lb2_batch_compute_done", "lb2_batch_compute_wait", "lb2_batch_download", "lb2_batch_download_view", "lb2_batch_stats", "lb2_batch_set_class_timing", "lb2_batch_class_stats",
    "lb2_batch_destroy", "lb2_free", "lb2_int_peak",
    "lb2_sdp_create", "lb2_sdp_reset", "lb2_sdp_run_bcc", "lb2_sdp_run_remain", "lb2_sdp_stats", "lb2_sdp_destroy", "lb2_sdp_get_tracked", "lb2_sdp_set_tracked",
    "lb2_aux_run", "lb2_hash_line_run", "init_hash", "hash_split_map", "lb2_init_hash", "lb2_hash_split_map", "lb2_worker_aux_counts", "lb2_producer_set_reference", "lb2_ref_abi_offsets", "lb2_ref_abi_sizes", "lb2_worker_spawn", "lb2_worker_join", "lb2_worker_yield", "lb2_worker_parked_seconds", "lb2_dropin_warmup", "lb2_fiber_selftest",
    "frag_line_BCC", "frag_line_remain", "node_init_score", "node_free_score", "cover_rate",
    "build_node_max_heap", "build_node_min_heap", "build_node_minpos_heap",
]
# by-value line_node helpers declared in lamsa_b200/csrc/ref_abi.h
EXPORTS_REF_ABI = ["heap_add_node", "node_pop", "node_heap_extract_max", "node_heap_extract_minpos", "node_heap_update_min"]

_lib = None


def load_library():
    """Load liblamsa_b200.so; raises LibraryMissing (never falls back to a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU implementation to fall back to)")
    lib = C.CDLL(LIB_PATH)
    P, I, I64 = C.c_void_p, C.c_int, C.c_int64
    lib.lb2_last_error.restype = C.c_char_p
    lib.lb2_ctx_create.argtypes = [I, C.POINTER(P)]
    lib.lb2_ctx_destroy.argtypes = [P]
    lib.lb2_ctx_destroy.restype = None
    lib.lb2_ctx_set_scratch_limit.argtypes = [P, C.c_uint64]
    lib.lb2_ctx_set_reference.argtypes = [P, P, I64]
    lib.lb2_ctx_set_chunk_tasks.argtypes = [P, I64]
    lib.lb2_ctx_last_run_stats.argtypes = [P, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64)]
    lib.lb2_dp_run.argtypes = [P, I64, P, P, C.POINTER(P), C.POINTER(I64)]
    lib.lb2_batch_create.argtypes = [P, I64, P, C.POINTER(P)]
    lib.lb2_batch_create_pool.argtypes = [P, P, I64, I64, P, C.POINTER(P)]
    lib.lb2_dp_run_pool.argtypes = [P, P, I64, I64, P, P, C.POINTER(P), C.POINTER(I64)]
    lib.lb2_pool_pack.argtypes = [I64, P, P, I64, C.POINTER(I64)]
    lib.lb2_ctx_last_run_kernel_ms.argtypes = [P, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.lb2_host_alloc.argtypes = [C.c_size_t, C.POINTER(P)]
    lib.lb2_host_free.argtypes = [P]
    lib.lb2_host_free.restype = None
    lib.lb2_batch_upload.argtypes = [P]
    lib.lb2_batch_compute.argtypes = [P, C.POINTER(C.c_float)]
    lib.lb2_batch_compute_async.argtypes = [P]
    lib.lb2_batch_compute_done.argtypes = [P]
    lib.lb2_batch_compute_wait.argtypes = [P, C.POINTER(C.c_float)]
    lib.lb2_batch_download.argtypes = [P, P, C.POINTER(P), C.POINTER(I64)]
    lib.lb2_batch_download_view.argtypes = [P, P, C.POINTER(P), C.POINTER(I64)]
    lib.lb2_batch_stats.argtypes = [P, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64),
                                    C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.lb2_batch_set_class_timing.argtypes = [P, I]
    lib.lb2_batch_class_stats.argtypes = [P, C.POINTER(ClassStat), I]
    lib.lb2_batch_destroy.argtypes = [P]
    lib.lb2_aux_run.argtypes = [P, I64, C.POINTER(AuxTask), P]
    lib.lb2_hash_line_run.argtypes = [P, I64, C.POINTER(HashTask)]
    lib.lb2_batch_destroy.restype = None
    lib.lb2_free.argtypes = [P]
    lib.lb2_free.restype = None
    lib.lb2_int_peak.argtypes = [P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(I), C.POINTER(I)]
    u8p, i8p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.POINTER(I)
    cpp = C.POINTER(C.POINTER(C.c_int32))
    app = C.POINTER(AlnPara)
    lib.ksw_global2.argtypes = [I, u8p, I, u8p, I, i8p, I, I, I, I, I, ip, cpp]
    lib.ksw_global.argtypes = [I, u8p, I, u8p, I, i8p, I, I, I, ip, cpp]
    lib.ksw_extend2.argtypes = [I, u8p, I, u8p, I, i8p, I, I, I, I, I, I, I, I, ip, ip, ip, ip, ip]
    lib.ksw_extend.argtypes = [I, u8p, I, u8p, I, i8p, I, I, I, I, I, I, ip, ip, ip, ip, ip]
    for name in ("ksw_extend_core", "ksw_extend_c", "ksw_extend_r"):
        getattr(lib, name).argtypes = [I, u8p, I, u8p, I, i8p, I, I, app, ip, ip, cpp, ip, ip]
    lib.ksw_bi_extend.argtypes = [I, u8p, I, u8p, I, i8p, I, I, app, cpp, ip, ip]
    _lib = lib
    return lib
