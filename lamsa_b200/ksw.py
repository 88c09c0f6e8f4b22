"""Host-side mirror of the reference's ksw interface (reference src/ksw.h:64-127).

Two ways in, both going through the C ABI of ``liblamsa_b200.so``:

* ``ksw_global2`` / ``ksw_extend_core`` / ``ksw_bi_extend`` ... -- same names and
  argument meaning as the reference prototypes; each call reaches the drop-in
  symbol of the same name (one blocking GPU task per DP).
* ``Context`` / ``Batch`` -- the batch producer: thousands of tasks per launch.

There is no CPU implementation here; without the library or a GPU these raise.
"""
import ctypes as C
import weakref

import numpy as np

from ._lib import (AlnPara, FLAG_CIGAR, KIND_EXTEND, KIND_GLOBAL, RESULT_DTYPE, TASK_DTYPE,
                   load_library)


def default_matrix(match=1, mis=3):
    """5x5 matrix of the reference's lamsa_fill_mat (src/lamsa_aln.c:1331-1340)."""
    m = np.full((5, 5), -mis, dtype=np.int8)
    for i in range(4):
        m[i, i] = match
    m[4, :] = -1
    m[:, 4] = -1
    return np.ascontiguousarray(m.reshape(-1))


def _err(lib, what):
    return RuntimeError(f"{what}: {lib.lb2_last_error().decode()}")


class Context:
    """One GPU context (lb2_ctx)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = C.c_void_p()
        if self.lib.lb2_ctx_create(int(device), C.byref(h)):
            raise _err(self.lib, "lb2_ctx_create")
        self.handle = h
        self.device = device

    def set_scratch_limit(self, nbytes):
        if self.lib.lb2_ctx_set_scratch_limit(self.handle, int(nbytes)):
            raise _err(self.lib, "lb2_ctx_set_scratch_limit")

    def set_reference(self, pac, l_pac):
        """Upload the 2-bit packed forward reference (bntseq .pac layout) once; see lb2_ctx_set_reference."""
        pac = np.ascontiguousarray(pac, dtype=np.uint8)
        if self.lib.lb2_ctx_set_reference(self.handle, pac.ctypes.data, int(l_pac)):
            raise _err(self.lib, "lb2_ctx_set_reference")

    def int_peak(self):
        a, b = C.c_double(), C.c_double()
        sm, khz = C.c_int(), C.c_int()
        if self.lib.lb2_int_peak(self.handle, C.byref(a), C.byref(b), C.byref(sm), C.byref(khz)):
            raise _err(self.lib, "lb2_int_peak")
        return {"gops_s16x2": a.value, "gops_s32": b.value, "sm_count": sm.value, "clock_khz": khz.value}

    def run(self, tasks, keep=()):
        """One-shot lb2_dp_run (host task records in, host results out; large batches are
        chunked and pipelined inside the library) -> (results, cigar_pool)."""
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        res = np.zeros(len(tasks), dtype=RESULT_DTYPE)
        pool, pn = C.c_void_p(), C.c_int64()
        if self.lib.lb2_dp_run(self.handle, len(tasks), tasks.ctypes.data, res.ctypes.data, C.byref(pool), C.byref(pn)):
            raise _err(self.lib, "lb2_dp_run")
        if not pn.value:
            self.lib.lb2_free(pool)
            return res, np.zeros(0, dtype=np.int32)
        # wrap the malloc'd pool without copying; it is freed when the array goes away
        cig = np.ctypeslib.as_array(C.cast(pool, C.POINTER(C.c_int32)), shape=(pn.value,))
        weakref.finalize(cig, self.lib.lb2_free, pool.value)
        return res, cig

    def run_pool(self, tasks, pool):
        """lb2_dp_run_pool: as run(), with every sequence of `tasks` inside the uint8 array `pool` (see pinned_pool /
        workload.pool_tasks): no host copy of the sequences, the pool range is uploaded as it lies."""
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        res = np.zeros(len(tasks), dtype=RESULT_DTYPE)
        out, pn = C.c_void_p(), C.c_int64()
        if self.lib.lb2_dp_run_pool(self.handle, pool.ctypes.data, pool.nbytes, len(tasks), tasks.ctypes.data,
                                    res.ctypes.data, C.byref(out), C.byref(pn)):
            raise _err(self.lib, "lb2_dp_run_pool")
        if not pn.value:
            self.lib.lb2_free(out)
            return res, np.zeros(0, dtype=np.int32)
        cig = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int32)), shape=(pn.value,))
        weakref.finalize(cig, self.lib.lb2_free, out.value)
        return res, cig

    def aux_counts(self, cigars, reads, ref_pacs):
        """lb2_aux_run (include/lamsa_b200.h section 5): per record (CIGAR words, read codes, pac coordinate) ->
        array of (n_match, n_mismatch, n_ins_open, n_ins_ext, n_del_open, n_del_ext, read_used, ref_used)."""
        from ._lib import AuxTask
        n = len(cigars)
        tasks = (AuxTask * max(n, 1))()
        keep = []
        for i in range(n):
            c = np.ascontiguousarray(cigars[i], dtype=np.int32); r = np.ascontiguousarray(reads[i], dtype=np.uint8)
            keep += [c, r]
            tasks[i].cigar = c.ctypes.data if len(c) else None; tasks[i].n_cigar = len(c)
            tasks[i].read = r.ctypes.data if len(r) else None; tasks[i].read_len = len(r)
            tasks[i].ref_pac = int(ref_pacs[i])
        out = np.zeros((n, 8), dtype=np.int32)
        if self.lib.lb2_aux_run(self.handle, n, tasks, out.ctypes.data):
            raise _err(self.lib, "lb2_aux_run")
        return out

    def set_chunk_tasks(self, n):
        if self.lib.lb2_ctx_set_chunk_tasks(self.handle, int(n)):
            raise _err(self.lib, "lb2_ctx_set_chunk_tasks")

    def last_run_stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.lib.lb2_ctx_last_run_stats(self.handle, C.byref(a), C.byref(b), C.byref(c))
        return {"h2d_bytes": a.value, "d2h_bytes": b.value, "launches": c.value}

    def last_run_kernel_ms(self):
        f, t = C.c_float(), C.c_float()
        self.lib.lb2_ctx_last_run_kernel_ms(self.handle, C.byref(f), C.byref(t))
        return {"fill_ms": f.value, "trace_ms": t.value}

    def close(self):
        if self.handle:
            self.lib.lb2_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_pool(nbytes):
    """uint8 array over page-locked host memory (lb2_host_alloc); freed when the array goes away."""
    lib = load_library()
    p = C.c_void_p()
    if lib.lb2_host_alloc(int(nbytes), C.byref(p)):
        raise _err(lib, "lb2_host_alloc")
    arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(max(int(nbytes), 1),))
    weakref.finalize(arr, lib.lb2_host_free, p.value)
    return arr[:int(nbytes)] if nbytes else arr[:0]


class Batch:
    """Staged batch (lb2_batch): create=pack, upload=H2D, compute=kernels, download=D2H.  With `pool` (the uint8
    array holding every sequence of `tasks`): lb2_batch_create_pool."""

    def __init__(self, ctx, tasks, keep=(), pool=None):
        self.ctx, self.lib = ctx, ctx.lib
        tasks = np.ascontiguousarray(tasks, dtype=TASK_DTYPE)
        self.n = len(tasks)
        self._keep = (tasks, keep, pool)
        h = C.c_void_p()
        if pool is None:
            if self.lib.lb2_batch_create(ctx.handle, self.n, tasks.ctypes.data, C.byref(h)):
                raise _err(self.lib, "lb2_batch_create")
        elif self.lib.lb2_batch_create_pool(ctx.handle, pool.ctypes.data, pool.nbytes, self.n, tasks.ctypes.data, C.byref(h)):
            raise _err(self.lib, "lb2_batch_create_pool")
        self.handle = h

    def upload(self):
        if self.lib.lb2_batch_upload(self.handle):
            raise _err(self.lib, "lb2_batch_upload")

    def compute(self):
        ms = C.c_float()
        if self.lib.lb2_batch_compute(self.handle, C.byref(ms)):
            raise _err(self.lib, "lb2_batch_compute")
        return ms.value

    def compute_async(self):
        """Enqueue the kernels and return (lb2_batch_compute_async); pair with done() / wait()."""
        if self.lib.lb2_batch_compute_async(self.handle):
            raise _err(self.lib, "lb2_batch_compute_async")

    def done(self):
        return bool(self.lib.lb2_batch_compute_done(self.handle))

    def wait(self):
        ms = C.c_float()
        if self.lib.lb2_batch_compute_wait(self.handle, C.byref(ms)):
            raise _err(self.lib, "lb2_batch_compute_wait")
        return ms.value

    def download(self, want_cigar=True, copy=True):
        """-> (results, cigar words).  copy=False returns a view into the batch's pinned
        staging (valid until close() or the next download)."""
        res = np.zeros(self.n, dtype=RESULT_DTYPE)
        pool, pn = C.c_void_p(), C.c_int64()
        fn = self.lib.lb2_batch_download if copy else self.lib.lb2_batch_download_view
        if fn(self.handle, res.ctypes.data, C.byref(pool) if want_cigar else None, C.byref(pn)):
            raise _err(self.lib, "lb2_batch_download")
        cig = np.zeros(0, dtype=np.int32)
        if want_cigar and pn.value:
            cig = np.ctypeslib.as_array(C.cast(pool, C.POINTER(C.c_int32)), shape=(pn.value,))
            if copy:
                cig = cig.copy()
        if want_cigar and copy:
            self.lib.lb2_free(pool)
        return res, cig

    def set_class_timing(self, on=True):
        self.lib.lb2_batch_set_class_timing(self.handle, 1 if on else 0)

    def class_stats(self):
        """Per kernel class of the last compute() with class timing on, after download(): list of dicts."""
        from ._lib import ClassStat
        arr = (ClassStat * 64)()
        n = self.lib.lb2_batch_class_stats(self.handle, arr, 64)
        return [{"kernel": arr[i].kernel.decode(), "kind": "extend" if arr[i].kind else "global", "variant": arr[i].variant,
                 "window_slots": arr[i].window_slots, "tasks": arr[i].tasks, "cells": arr[i].cells, "ms": arr[i].ms}
                for i in range(max(n, 0))]

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        f, t = C.c_float(), C.c_float()
        self.lib.lb2_batch_stats(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(f), C.byref(t))
        return {"h2d_bytes": a.value, "d2h_bytes": b.value, "launches": c.value,
                "fill_ms": f.value, "trace_ms": t.value}

    def close(self):
        if self.handle:
            self.lib.lb2_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_tasks(kind, qseq, qoff, qlen, tseq, toff, tlen, w, mat, *, h0=0, o_del=5, e_del=2, o_ins=5,
               e_ins=2, end_bonus=0, zdrop=0, cigar=True, m=5):
    """Vectorised construction of a TASK_DTYPE array over pooled sequences.

    qseq/tseq: uint8 arrays holding all sequences; task i uses qseq[qoff[i]:qoff[i]+qlen[i]].
    Scalar or per-task arrays are accepted for every parameter.  The caller must
    keep qseq, tseq and mat alive while the tasks are in use.
    """
    n = len(qlen)
    t = np.zeros(n, dtype=TASK_DTYPE)
    t["kind"] = kind
    t["flags"] = np.where(np.broadcast_to(np.asarray(cigar), (n,)), FLAG_CIGAR, 0)
    t["qlen"], t["tlen"] = qlen, tlen
    t["query"] = qseq.ctypes.data + np.asarray(qoff, dtype=np.uint64)
    t["target"] = tseq.ctypes.data + np.asarray(toff, dtype=np.uint64)
    for name, v in (("w", w), ("h0", h0), ("o_del", o_del), ("e_del", e_del), ("o_ins", o_ins),
                    ("e_ins", e_ins), ("end_bonus", end_bonus), ("zdrop", zdrop), ("m", m)):
        t[name] = v
    t["mat"] = mat.ctypes.data
    return t


# ------------------------------------------------------------ drop-in mirrors --
def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.POINTER(C.c_uint8))


def _i8(a):
    a = np.ascontiguousarray(a, dtype=np.int8)
    return a, a.ctypes.data_as(C.POINTER(C.c_int8))


def _take_cigar(lib, ptr, n):
    out = [int(ptr[i]) for i in range(n)] if n else []
    if ptr:
        lib.lb2_free(C.cast(ptr, C.c_void_p))
    return out


def ksw_global2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, w, want_cigar=True):
    """reference src/ksw.c:543 -> (score, cigar words or None)"""
    lib = load_library()
    q, qp = _u8(query)
    t, tp = _u8(target)
    mm, mp = _i8(mat)
    if not want_cigar:
        return lib.ksw_global2(qlen, qp, tlen, tp, m, mp, o_del, e_del, o_ins, e_ins, w, None, None), None
    n, c = C.c_int(), C.POINTER(C.c_int32)()
    s = lib.ksw_global2(qlen, qp, tlen, tp, m, mp, o_del, e_del, o_ins, e_ins, w, C.byref(n), C.byref(c))
    return s, _take_cigar(lib, c, n.value)


def ksw_global(qlen, query, tlen, target, m, mat, gapo, gape, w, want_cigar=True):
    """reference src/ksw.c:655"""
    return ksw_global2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w, want_cigar)


def ksw_extend2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0):
    """reference src/ksw.c:387 -> (score, qle, tle, gtle, gscore, max_off)"""
    lib = load_library()
    q, qp = _u8(query)
    t, tp = _u8(target)
    mm, mp = _i8(mat)
    o = [C.c_int() for _ in range(5)]
    s = lib.ksw_extend2(qlen, qp, tlen, tp, m, mp, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                        *[C.byref(x) for x in o])
    return (s,) + tuple(x.value for x in o)


def ksw_extend(qlen, query, tlen, target, m, mat, gapo, gape, w, end_bonus, zdrop, h0):
    """reference src/ksw.c:492"""
    return ksw_extend2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w, end_bonus, zdrop, h0)


def _ext_call(name, qlen, query, tlen, target, m, mat, w, h0, AP):
    lib = load_library()
    q, qp = _u8(query)
    t, tp = _u8(target)
    mm, mp = _i8(mat)
    qle, tle, n, cap = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    c = C.POINTER(C.c_int32)()
    r = getattr(lib, name)(qlen, qp, tlen, tp, m, mp, w, h0, C.byref(AP), C.byref(qle), C.byref(tle),
                           C.byref(c), C.byref(n), C.byref(cap))
    return r, qle.value, tle.value, _take_cigar(lib, c, n.value), cap.value


def ksw_extend_core(qlen, query, tlen, target, m, mat, w, h0, AP):
    """reference src/ksw.c:667 -> (max, qle, tle, cigar, m_cigar)"""
    return _ext_call("ksw_extend_core", qlen, query, tlen, target, m, mat, w, h0, AP)


def ksw_extend_c(qlen, query, tlen, target, m, mat, w, h0, AP):
    """reference src/ksw.c:809 -> (0|1|2, qle, tle, cigar, m_cigar)"""
    return _ext_call("ksw_extend_c", qlen, query, tlen, target, m, mat, w, h0, AP)


def ksw_extend_r(qlen, query, tlen, target, m, mat, w, h0, AP):
    """reference src/ksw.c:820 -> (0|1|2, qre, tre, cigar, m_cigar)"""
    return _ext_call("ksw_extend_r", qlen, query, tlen, target, m, mat, w, h0, AP)


def ksw_bi_extend(qlen, query, tlen, target, m, mat, lh0, rh0, AP):
    """reference src/ksw.c:862 -> (0|1, cigar, m_cigar)"""
    lib = load_library()
    q, qp = _u8(query)
    t, tp = _u8(target)
    mm, mp = _i8(mat)
    n, cap = C.c_int(), C.c_int()
    c = C.POINTER(C.c_int32)()
    r = lib.ksw_bi_extend(qlen, qp, tlen, tp, m, mp, lh0, rh0, C.byref(AP), C.byref(c), C.byref(n), C.byref(cap))
    return r, _take_cigar(lib, c, n.value), cap.value
