"""ctypes binding of the sparse-DP chaining interface (include/lamsa_b200.h section 3):
frag_line_BCC / frag_line_remain of the reference (src/lamsa_dp_con.c:1305,1252) for a batch of
reads, on the GPU.  No CPU implementation; raises when the library or a GPU is missing."""
import ctypes as C

import numpy as np

from ._lib import load_library

HIT_DTYPE = np.dtype([("offset", "<i8"), ("nchr", "<i4"), ("NM", "<i4"), ("len_dif", "<i4"), ("nstrand", "<i4")])
REG_DTYPE = np.dtype([("beg", "<i4"), ("end", "<i4"), ("chr", "<i4"), ("is_rev", "<i4"),
                      ("ref_beg", "<i8"), ("ref_end", "<i8")])
READ_DTYPE = np.dtype([("seed_out", "<i4"), ("seed_all", "<i4"), ("read_len", "<i4"), ("n_reg", "<i4"),
                       ("seed_first", "<i8"), ("hit_first", "<i8"), ("reg_first", "<i8")])
PARA_DTYPE = np.dtype([("seed_len", "<i4"), ("seed_step", "<i4"), ("seed_inv", "<i4"), ("per_aln_m", "<i4"),
                       ("first_loci_thd", "<i4"), ("SV_len_thd", "<i4"), ("ske_max", "<i4"), ("ovlp_rat", "<f4"),
                       ("split_len", "<i4"), ("match_dis", "<i4"), ("mismatch_thd", "<i4"), ("aln_mode", "<i4"),
                       ("bwt_seed_len", "<i4"), ("frag_score_table", "<i4", (10,))])

SDP_EXPORTS = ["lb2_sdp_create", "lb2_sdp_run_bcc", "lb2_sdp_run_remain", "lb2_sdp_stats", "lb2_sdp_destroy",
               "frag_line_BCC", "frag_line_remain"]


def _bind(lib):
    if getattr(lib, "_sdp_bound", False):
        return lib
    P, I64 = C.c_void_p, C.c_int64
    lib.lb2_sdp_create.argtypes = [P, P, I64, P, P, P, P, C.POINTER(P)]
    lib.lb2_sdp_run_bcc.argtypes = [P, C.POINTER(P), C.POINTER(P), C.POINTER(C.c_float)]
    lib.lb2_sdp_run_remain.argtypes = [P, P, P, C.POINTER(P), C.POINTER(P), C.POINTER(C.c_float)]
    lib.lb2_sdp_stats.argtypes = [P, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64)]
    lib.lb2_sdp_get_tracked.argtypes = [P, P]
    lib.lb2_sdp_set_tracked.argtypes = [P, P]
    lib.lb2_sdp_reset.argtypes = [P, P, I64, P, P, P, P]
    lib.lb2_sdp_destroy.argtypes = [P]
    lib.lb2_sdp_destroy.restype = None
    lib._sdp_bound = True
    return lib


class SdpBatch:
    """A set of reads resident on the GPU for the two chaining stages.

    para: PARA_DTYPE scalar; reads: READ_DTYPE array; seed_id/map_n: per seed with hits;
    hits: HIT_DTYPE array, seed-major per read."""

    def __init__(self, ctx, para, reads, seed_id, map_n, hits):
        self.lib = _bind(load_library())
        self.ctx = ctx
        self.para = np.array(para, dtype=PARA_DTYPE).reshape(())
        self.reads = np.ascontiguousarray(reads, dtype=READ_DTYPE)
        seed_id = np.ascontiguousarray(seed_id, dtype=np.int32)
        map_n = np.ascontiguousarray(map_n, dtype=np.int32)
        hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
        h = C.c_void_p()
        if self.lib.lb2_sdp_create(ctx.handle, self.para.ctypes.data, len(self.reads), self.reads.ctypes.data,
                                   seed_id.ctypes.data, map_n.ctypes.data, hits.ctypes.data, C.byref(h)):
            raise RuntimeError("lb2_sdp_create: " + self.lib.lb2_last_error().decode())
        self.handle = h
        self.kernel_ms = 0.0
        # hits of the batch, in batch order (the reads may reference the arrays in any order)
        self._map_n_sum = [int(map_n[int(r["seed_first"]):int(r["seed_first"]) + int(r["seed_out"])].sum()) for r in self.reads]

    def _result(self, sp, op):
        n = len(self.reads)
        off = np.ctypeslib.as_array(C.cast(op, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
        total = int(off[-1])
        if total == 0:
            return np.zeros(0, np.int32), off
        return np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_int32)), shape=(total,)).copy(), off

    def run_bcc(self):
        """Stage 1 for every read -> (stream, off): read r owns stream[off[r]:off[r+1]]."""
        sp, op, ms = C.c_void_p(), C.c_void_p(), C.c_float()
        if self.lib.lb2_sdp_run_bcc(self.handle, C.byref(sp), C.byref(op), C.byref(ms)):
            raise RuntimeError("lb2_sdp_run_bcc: " + self.lib.lb2_last_error().decode())
        self.kernel_ms = ms.value
        return self._result(sp, op)

    def run_remain(self, reads, regs):
        """Stage 2: reads carry n_reg / reg_first into regs (the aligned records of stage 1)."""
        reads = np.ascontiguousarray(reads, dtype=READ_DTYPE)
        regs = np.ascontiguousarray(regs, dtype=REG_DTYPE)
        assert len(reads) == len(self.reads)
        sp, op, ms = C.c_void_p(), C.c_void_p(), C.c_float()
        if self.lib.lb2_sdp_run_remain(self.handle, reads.ctypes.data, regs.ctypes.data if len(regs) else None,
                                       C.byref(sp), C.byref(op), C.byref(ms)):
            raise RuntimeError("lb2_sdp_run_remain: " + self.lib.lb2_last_error().decode())
        self.kernel_ms = ms.value
        return self._result(sp, op)

    def get_tracked(self):
        """One flag per hit (batch order): on a stage-1 skeleton.  Call after run_bcc()."""
        n = int(sum(int(m) for m in self._map_n_sum))
        flags = np.zeros(n, np.uint8)
        if self.lib.lb2_sdp_get_tracked(self.handle, flags.ctypes.data if n else None):
            raise RuntimeError("lb2_sdp_get_tracked: " + self.lib.lb2_last_error().decode())
        return flags

    def set_tracked(self, flags):
        """Load the stage-1 flags into a batch that did not run stage 1 itself, then run_remain()."""
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        if self.lib.lb2_sdp_set_tracked(self.handle, flags.ctypes.data if len(flags) else None):
            raise RuntimeError("lb2_sdp_set_tracked: " + self.lib.lb2_last_error().decode())

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.lib.lb2_sdp_stats(self.handle, C.byref(a), C.byref(b), C.byref(c))
        return {"pairs": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    def close(self):
        if self.handle:
            self.lib.lb2_sdp_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
