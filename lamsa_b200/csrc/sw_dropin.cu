// sw_dropin.cu -- the reference's local Smith-Waterman entry points (src/ksw.h:62-63, src/ksw.c:68-377: ksw_qinit,
// ksw_u8, ksw_i16, ksw_align2, ksw_align) on top of sw_local.cuh, and the batch call lb2_sw_run underneath them.
// Nothing in LAMSA calls these; they are exported because the reference's header declares them and
// SURVEY 8(b) lists them as part of the boundary.  Every cell is evaluated on the GPU; no CPU path.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/lamsa_b200.h"
#include "ctx_internal.h"
#include "dropin_internal.h"
#include "sw_local.cuh"

using namespace lb2;

namespace {
struct SwBuffers { uint8_t* h = nullptr; uint8_t* d = nullptr; size_t cap = 0; int32_t* d_scratch = nullptr; size_t scratch_cap = 0; std::mutex mu; };
std::mutex g_mu;
std::map<lb2_ctx*, SwBuffers*> g_buffers;
SwBuffers* buffers_of(lb2_ctx* c) {
    std::lock_guard<std::mutex> lk(g_mu);
    SwBuffers*& b = g_buffers[c];
    if (!b) b = new SwBuffers();
    return b;
}
size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }
}  // namespace

#define CUS(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    return lb2::set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } while (0)

extern "C" int lb2_sw_run(lb2_ctx* ctx, int64_t n, const lb2_sw_task* tasks, lb2_sw_result* results) {
    if (!ctx || n < 0 || (n > 0 && (!tasks || !results))) return lb2::set_error("lb2_sw_run: bad argument");
    if (n == 0) return 0;
    SwBuffers* B = buffers_of(ctx);
    std::lock_guard<std::mutex> lk(B->mu);
    CUS(cudaSetDevice(lb2::ctx_device(ctx)));
    cudaStream_t s = lb2::ctx_stream(ctx);
    std::vector<SwTask> dt((size_t)n);
    size_t bytes = 0, ints = 0;
    for (int64_t i = 0; i < n; ++i) {
        const lb2_sw_task& t = tasks[i];
        if (t.qlen < 0 || t.tlen < 0 || t.m < 1 || t.m > 64 || !t.mat || (t.qlen && !t.query) || (t.tlen && !t.target) || (t.size != 1 && t.size != 2))
            return lb2::set_error("lb2_sw_run: task %lld is malformed", (long long)i);
        int maxv = 0;
        for (int a = 0; a < t.m * t.m; ++a) maxv = std::max(maxv, (int)t.mat[a]);
        if (t.e_ins < 0 || t.o_ins < 0 || t.e_del < 0 || t.o_del < 0 ||
            (int64_t)(t.qlen + 16) * (t.e_ins + maxv) + t.o_ins + 8 >= (1 << 24))
            return lb2::set_error("lb2_sw_run: task %lld: qlen %d with these penalties exceeds the 24-bit scan domain", (long long)i, t.qlen);
        SwTask& d = dt[(size_t)i];
        d.qlen = t.qlen; d.tlen = t.tlen; d.m = t.m; d.o_del = t.o_del; d.e_del = t.e_del; d.o_ins = t.o_ins; d.e_ins = t.e_ins;
        d.xtra = t.xtra; d.size = t.size;
        d.q_off = (uint32_t)bytes; bytes += up16((size_t)t.qlen + 1);
        d.t_off = (uint32_t)bytes; bytes += up16((size_t)t.tlen + 1);
        d.mat_off = (uint32_t)bytes; bytes += up16((size_t)t.m * t.m);
        const int lanes = t.size == 1 ? 16 : 8;
        const size_t W = (size_t)((t.qlen + lanes - 1) / lanes) * lanes;
        d.he_off = (uint32_t)ints; ints += 2 * W + 4;
        d.rowmax_off = (uint32_t)ints; ints += (t.xtra & kSwXSubo) ? (size_t)t.tlen + 4 : 4;
        if ((bytes | ints) >> 31) return lb2::set_error("lb2_sw_run: batch too large");
    }
    const size_t o_task = 0, o_seq = up16(sizeof(SwTask) * (size_t)n), o_res = o_seq + up16(bytes), total = o_res + up16(sizeof(SwResult) * (size_t)n);
    if (B->cap < total) {
        const size_t cap = std::max(total + total / 4, B->cap * 2);
        CUS(cudaStreamSynchronize(s));
        cudaFreeHost(B->h); cudaFree(B->d); B->h = nullptr; B->d = nullptr; B->cap = 0;
        CUS(cudaMallocHost(&B->h, cap)); CUS(cudaMalloc(&B->d, cap)); B->cap = cap;
    }
    if (B->scratch_cap < ints) {
        const size_t cap = std::max(ints + ints / 4, B->scratch_cap * 2);
        CUS(cudaStreamSynchronize(s));
        cudaFree(B->d_scratch); B->d_scratch = nullptr; B->scratch_cap = 0;
        CUS(cudaMalloc(&B->d_scratch, cap * 4)); B->scratch_cap = cap;
    }
    memcpy(B->h + o_task, dt.data(), sizeof(SwTask) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const lb2_sw_task& t = tasks[i]; const SwTask& d = dt[(size_t)i];
        if (t.qlen) memcpy(B->h + o_seq + d.q_off, t.query, (size_t)t.qlen);
        if (t.tlen) memcpy(B->h + o_seq + d.t_off, t.target, (size_t)t.tlen);
        memcpy(B->h + o_seq + d.mat_off, t.mat, (size_t)t.m * t.m);
    }
    CUS(cudaMemcpyAsync(B->d, B->h, o_res, cudaMemcpyHostToDevice, s));
    const int wpb = 4;
    sw_local_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, s>>>(reinterpret_cast<const SwTask*>(B->d + o_task), (int)n, B->d + o_seq,
                                                                          B->d_scratch, reinterpret_cast<SwResult*>(B->d + o_res));
    CUS(cudaGetLastError());
    CUS(cudaMemcpyAsync(B->h + o_res, B->d + o_res, sizeof(SwResult) * (size_t)n, cudaMemcpyDeviceToHost, s));
    CUS(cudaStreamSynchronize(s));
    const SwResult* rr = reinterpret_cast<const SwResult*>(B->h + o_res);
    for (int64_t i = 0; i < n; ++i) {
        lb2_sw_result& o = results[i];
        o.score = rr[i].score; o.te = rr[i].te; o.qe = rr[i].qe; o.score2 = rr[i].score2; o.te2 = rr[i].te2; o.tb = rr[i].tb; o.qb = rr[i].qb;
    }
    return 0;
}

// ---- drop-in entry points ------------------------------------------------------------------------------------------
// the query "profile": opaque to callers, who only pass it back and free() it (src/ksw.h:58-60) -- one malloc block
struct _kswq_t {
    int qlen, slen;
    uint8_t shift, mdiff, max, size;           // as ksw_qinit fills them (src/ksw.c:88-96)
    int m;
    // followed by m*m matrix bytes and qlen query codes
};

namespace {
std::mutex g_sw_mu;
kswr_t run_profile(kswq_t* q, int tlen, const uint8_t* target, int o_del, int e_del, int o_ins, int e_ins, int xtra) {
    const int8_t* mat = reinterpret_cast<const int8_t*>(q + 1);
    const uint8_t* query = reinterpret_cast<const uint8_t*>(mat + q->m * q->m);
    lb2_sw_task t; memset(&t, 0, sizeof t);
    t.query = query; t.qlen = q->qlen; t.target = target; t.tlen = tlen; t.m = q->m; t.mat = mat;
    t.o_del = o_del; t.e_del = e_del; t.o_ins = o_ins; t.e_ins = e_ins; t.xtra = xtra; t.size = q->size;
    lb2_sw_result r;
    int rc;
    { std::lock_guard<std::mutex> lk(g_sw_mu); rc = lb2_sw_run(lb2::dropin_ctx(), 1, &t, &r); }
    if (rc) { fprintf(stderr, "[lamsa_b200] local alignment failed: %s\n", lb2_last_error()); exit(1); }
    kswr_t k; k.score = r.score; k.te = r.te; k.qe = r.qe; k.score2 = r.score2; k.te2 = r.te2; k.tb = r.tb; k.qb = r.qb;
    return k;
}
}  // namespace

extern "C" {

kswq_t* ksw_qinit(int size, int qlen, const uint8_t* query, int m, const int8_t* mat) {
    size = size > 1 ? 2 : 1;
    const int p = 8 * (3 - size);
    if (qlen < 0) qlen = 0;
    kswq_t* q = (kswq_t*)malloc(sizeof(kswq_t) + (size_t)m * m + (size_t)qlen + 16);
    q->qlen = qlen; q->slen = (qlen + p - 1) / p; q->size = (uint8_t)size; q->m = m;
    int minv = 127, maxv = 0;
    for (int a = 0; a < m * m; ++a) { minv = mat[a] < minv ? mat[a] : minv; maxv = mat[a] > maxv ? mat[a] : maxv; }
    q->max = (uint8_t)maxv; q->shift = (uint8_t)(256 - (uint8_t)minv); q->mdiff = (uint8_t)(maxv + q->shift);
    int8_t* qm = reinterpret_cast<int8_t*>(q + 1);
    memcpy(qm, mat, (size_t)m * m);
    if (qlen) memcpy(qm + m * m, query, (size_t)qlen);
    return q;
}

kswr_t ksw_u8(kswq_t* q, int tlen, const uint8_t* target, int o_del, int e_del, int o_ins, int e_ins, int xtra) {
    const uint8_t keep = q->size; q->size = 1;
    const kswr_t r = run_profile(q, tlen, target, o_del, e_del, o_ins, e_ins, xtra);
    q->size = keep;
    return r;
}
kswr_t ksw_i16(kswq_t* q, int tlen, const uint8_t* target, int o_del, int e_del, int o_ins, int e_ins, int xtra) {
    const uint8_t keep = q->size; q->size = 2;
    const kswr_t r = run_profile(q, tlen, target, o_del, e_del, o_ins, e_ins, xtra);
    q->size = keep;
    return r;
}

kswr_t ksw_align2(int qlen, uint8_t* query, int tlen, uint8_t* target, int m, const int8_t* mat,
                  int o_del, int e_del, int o_ins, int e_ins, int xtra, kswq_t** qry)
{
    kswq_t* q = (qry && *qry) ? *qry : ksw_qinit((xtra & kSwXByte) ? 1 : 2, qlen, query, m, mat);
    if (qry && *qry == nullptr) *qry = q;
    const int size = q->size;
    kswr_t r = run_profile(q, tlen, target, o_del, e_del, o_ins, e_ins, xtra);
    if (qry == nullptr) free(q);
    if ((xtra & kSwXStart) == 0 || ((xtra & kSwXSubo) && r.score < (xtra & 0xffff))) return r;
    if (r.qe < 0) return r;          // byte overflow (score 255): the reference builds an empty profile here and reads before it
    // start point: the same search over the reversed prefixes, stopped at the score found (src/ksw.c:361-369)
    std::vector<uint8_t> rq((size_t)r.qe + 1), rt((size_t)(tlen > 0 ? tlen : 1));
    for (int k = 0; k <= r.qe; ++k) rq[(size_t)k] = query[r.qe - k];
    if (tlen) memcpy(rt.data(), target, (size_t)tlen);
    for (int k = 0; k <= r.te; ++k) rt[(size_t)k] = target[r.te - k];
    kswq_t* q2 = ksw_qinit(size, r.qe + 1, rq.data(), m, mat);
    const kswr_t rr = run_profile(q2, tlen, rt.data(), o_del, e_del, o_ins, e_ins, kSwXStop | r.score);
    free(q2);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return r;
}

kswr_t ksw_align(int qlen, uint8_t* query, int tlen, uint8_t* target, int m, const int8_t* mat, int gapo, int gape, int xtra, kswq_t** qry) {
    return ksw_align2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, xtra, qry);
}

}  // extern "C"
