// ksw_dropin.cu -- the reference's ksw_* entry points (include/lamsa_b200.h,
// section 1) implemented on top of the batch interface.  Every DP cell is
// evaluated by the CUDA kernels; this file only holds the host control flow
// that the reference keeps in its wrappers (src/ksw.c:809-926) and the CIGAR
// list helpers those wrappers use (src/frag_check.h:139-188).
//
// Calls block; concurrent callers (the reference's n_thread workers) are combined
// into one GPU batch by the submitter below.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lamsa_b200.h"
#include "dropin_internal.h"
#include "cigar_list.h"

namespace {

std::mutex g_mu;
lb2_ctx* g_ctx = nullptr;

thread_local lb2_ctx* tl_ctx = nullptr;        // device threads of the batch producer own a context each

lb2_ctx* default_ctx() {
    if (tl_ctx) return tl_ctx;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) {
        int dev = 0;
        if (const char* e = getenv("LB2_DEVICE")) dev = atoi(e);
        if (lb2_ctx_create(dev, &g_ctx)) {
            fprintf(stderr, "[lamsa_b200] cannot open GPU %d: %s\n", dev, lb2_last_error());
            exit(1);                               // same style as the reference's fatal paths
        }
    }
    return g_ctx;
}

}  // namespace
namespace lb2 { lb2_ctx* dropin_ctx() { return default_ctx(); } }   // shared with sdp_dropin.cu
bool lb2::dropin_has_thread_ctx() { return tl_ctx != nullptr; }
void lb2::dropin_bind_thread_ctx(lb2_ctx* c) { tl_ctx = c; }

// Opens the batch producer's GPUs (contexts, batch slots, device threads: producer.cu) from a helper thread,
// so that CUDA start-up overlaps the caller's own start-up (index loading).  Optional.
extern "C" void lb2_dropin_warmup(void) {
    // CUDA start-up time grows with the number of GPUs it has to initialise: show it only the ones this process
    // will use (LB2_DEVICE .. LB2_DEVICE + LB2_DEVICES - 1), unless the environment already chose
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        const char* nd = getenv("LB2_DEVICES"); const char* b = getenv("LB2_DEVICE");
        const int ndev = nd && atoi(nd) > 0 ? atoi(nd) : 1, base = b ? atoi(b) : 0;
        std::string v;
        for (int k = 0; k < ndev; ++k) v += (k ? "," : "") + std::to_string(base + k);
        setenv("CUDA_VISIBLE_DEVICES", v.c_str(), 0);
        setenv("LB2_DEVICE", "0", 1);
    }
    std::thread([] { lb2::producer_warmup(); }).detach();
}
namespace {

// ---- combining submitter ----------------------------------------------------
// The reference calls ksw_* from n_thread pthreads, one read per thread
// (src/lamsa_aln.c:838-842), each call blocking.  Instead of one launch per
// call, callers park their task in a shared queue; one of them (the leader)
// waits a short gather window, submits everything queued as ONE batch and hands
// the results back.  With `lamsa aln -t 128` a launch carries up to 128 tasks
// without touching the reference's sources.  A lone caller (or -t 1) degrades to
// a batch of one after an empty gather window, so nothing can deadlock.
using Pending = lb2::DpRequest;

std::mutex q_mu;
std::condition_variable q_cv;
std::vector<Pending*> q_wait;
bool q_leader = false;
int q_last_batch = 1;

int gather_us() {
    static const int v = [] { const char* e = getenv("LB2_GATHER_US"); return e && *e ? atoi(e) : 40; }();
    return v;
}

// hand results (and malloc'd CIGARs with the capacity the reference would have grown to) back
void deliver(std::vector<Pending*>& batch, const lb2_result* results, const cigar32_t* pool) {
    for (size_t i = 0; i < batch.size(); ++i) {
        Pending* p = batch[i];
        const lb2_result& r = results[i];
        *p->res = r;
        if (p->cig) {
            if (r.n_cigar > 0) {
                cigar32_t* out = (cigar32_t*)malloc(sizeof(cigar32_t) * (size_t)r.reserved);
                memcpy(out, pool + r.cigar_off, sizeof(cigar32_t) * (size_t)r.n_cigar);
                *p->cig = out;
            } else *p->cig = nullptr;
        }
    }
}

void submit_batch(std::vector<Pending*>& batch) {
    lb2_ctx* c = default_ctx();
    const int64_t n = (int64_t)batch.size();
    std::vector<lb2_task> tasks((size_t)n);
    std::vector<lb2_result> results((size_t)n);
    for (int64_t i = 0; i < n; ++i) tasks[(size_t)i] = batch[(size_t)i]->task;
    cigar32_t* pool = nullptr; int64_t pn = 0;
    int rc;
    static FILE* round_log = [] { const char* e = getenv("LB2_ROUND_LOG"); return e && *e ? fopen(e, "w") : (FILE*)nullptr; }();
    const auto t0 = std::chrono::steady_clock::now();
    if (tl_ctx) rc = lb2_dp_run(c, n, tasks.data(), results.data(), &pool, &pn);      // this thread's own context
    else {
        std::lock_guard<std::mutex> lk(g_mu);      // one stream per context
        rc = lb2_dp_run(c, n, tasks.data(), results.data(), &pool, &pn);
    }
    if (rc) { fprintf(stderr, "[lamsa_b200] DP launch failed: %s\n", lb2_last_error()); exit(1); }
    if (round_log) {        // one line per submitted batch: tasks, longest task, cells, wall and kernel time
        const double us = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6;
        int max_t = 0, max_q = 0, max_w = 0; int64_t cells = 0, launches = 0; float fm = 0, tm = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (tasks[(size_t)i].tlen > max_t) { max_t = tasks[(size_t)i].tlen; max_q = tasks[(size_t)i].qlen; max_w = tasks[(size_t)i].w; }
            cells += results[(size_t)i].cells;
        }
        lb2_ctx_last_run_stats(c, nullptr, nullptr, &launches);
        lb2_ctx_last_run_kernel_ms(c, &fm, &tm);
        static std::mutex log_mu;
        std::lock_guard<std::mutex> lk(log_mu);
        static const auto log_t0 = std::chrono::steady_clock::now();
        const double at = std::chrono::duration<double>(t0 - log_t0).count() * 1e6;
        fprintf(round_log, "%lld %d %d %d %lld %.1f %.3f %.3f %lld %.0f %p\n", (long long)n, max_t, max_q, max_w, (long long)cells, us, fm, tm,
                (long long)launches, at, (void*)c);
    }
    deliver(batch, results.data(), pool);
    lb2_free(pool);
}

// Tasks without a single DP cell -- an empty query or an empty target: about 30 % of the extension calls of
// LAMSA's PacBio mode -- are answered in closed form: what the reference's code leaves behind
// when its cell loop is never entered (global: src/ksw.c:549,569-572,577-579,632-650; extension: :692-694,
// :709-712,:718-725,:758-763,:785-803).  No DP cell is evaluated on the host; a round trip to the GPU for nothing
// would only lengthen the read's chain of dependent calls.  Returns false for the one shape the reference itself
// leaves undefined (ksw_extend2 with an empty query writes past its eh[] array, :399,407), which stays on the GPU path.
cigar32_t* one_op(int len, int op) {
    cigar32_t* c = (cigar32_t*)malloc(4 * sizeof(cigar32_t));          // push_cigar's first allocation, :509-511
    c[0] = (cigar32_t)(len << 4 | op);
    return c;
}
bool zero_cell_task(const lb2_task& t, lb2_result* r, cigar32_t** cig) {
    const bool want = (t.flags & LB2_FLAG_CIGAR) != 0;
    memset(r, 0, sizeof *r);
    if (cig) *cig = nullptr;
    if (t.kind == LB2_KIND_GLOBAL) {
        r->qle = t.qlen; r->tle = t.tlen;
        if (t.qlen == 0 && t.tlen == 0) { r->score = 0; return true; }
        if (t.qlen == 0) {                       // every row only moves eh[0].h = -(o_del + e_del (i+1)); one deletion
            r->score = -(t.o_del + t.e_del * t.tlen);
            if (want) { r->n_cigar = 1; r->reserved = 4; if (cig) *cig = one_op(t.tlen, LB2_CDEL); }
        } else {                                 // no rows: eh[qlen].h of the first row (the band is >= qlen + 3); one insertion
            r->score = -(t.o_ins + t.e_ins * t.qlen);
            if (want) { r->n_cigar = 1; r->reserved = 4; if (cig) *cig = one_op(t.qlen, LB2_CINS); }
        }
        return true;
    }
    if (t.h0 <= 0) return false;
    if (t.tlen == 0) {                           // no rows: max = h0, max_i = max_j = max_ie = -1, gscore = -1
        r->score = t.h0; r->gscore = -1;
        return true;
    }
    if (!want) return false;                     // empty query, score-only form
    // empty query, row 0 has no cells: gscore = h1 = max(h0 - oe_del, 0), max_ie = 0, then m == 0 ends the loop
    r->score = t.h0;
    const int h1 = t.h0 - (t.o_del + t.e_del) > 0 ? t.h0 - (t.o_del + t.e_del) : 0;
    r->gscore = h1;
    if (!(h1 <= 0 || h1 <= t.h0 - t.end_bonus)) {    // end point (max_ie, qlen-1) = (0, -1): one deleted base
        r->tle = 1; r->gtle = 1;
        r->n_cigar = 1; r->reserved = 4;
        if (cig) *cig = one_op(1, LB2_CDEL);
    }
    return true;
}

// run one task; returns malloc'd CIGAR (or NULL) through *cig
void run_one(const lb2_task& t, lb2_result* r, cigar32_t** cig) {
    if ((t.qlen == 0 || t.tlen == 0) && !(t.flags & LB2_FLAG_TARGET_PAC) && zero_cell_task(t, r, cig)) return;
    Pending me{t, r, cig, false};
    if (lb2::fiber_active()) { lb2::fiber_wait_dp(&me); return; }      // worker fiber: park and let the scheduler batch it
    std::unique_lock<std::mutex> lk(q_mu);
    q_wait.push_back(&me);
    for (;;) {
        if (me.done) return;
        if (!q_leader) break;                      // nobody is collecting: I lead
        q_cv.wait(lk);
    }
    q_leader = true;
    // gather window: only worth waiting when the previous batch showed company
    if (q_last_batch > 1 && gather_us() > 0) {
        size_t seen = q_wait.size();
        for (int spin = 0; spin < 8; ++spin) {
            lk.unlock();
            std::this_thread::sleep_for(std::chrono::microseconds(gather_us()));
            lk.lock();
            if (q_wait.size() == seen) break;      // arrivals stopped
            seen = q_wait.size();
        }
    }
    std::vector<Pending*> batch;
    batch.swap(q_wait);
    q_last_batch = (int)batch.size();
    lk.unlock();
    submit_batch(batch);
    lk.lock();
    for (Pending* p : batch) p->done = true;
    q_leader = false;
    lk.unlock();
    q_cv.notify_all();
}

}  // namespace
namespace lb2 { void dropin_submit_dp(std::vector<DpRequest*>& batch) { submit_batch(batch); } }
namespace {

using namespace lb2::cigar_list;

}  // namespace

extern "C" {

int ksw_global2(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                int m, const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int* n_cigar, cigar32_t** cigar)
{
    if (qlen < 0 || tlen < 0) {
        fprintf(stderr, "[ksw_global2] Error: qlen: %d tlen: %d\n", qlen, tlen); exit(-1);
    }
    if (n_cigar) *n_cigar = 0;
    const bool want = n_cigar && cigar;
    lb2_task t; memset(&t, 0, sizeof t);
    t.kind = LB2_KIND_GLOBAL; t.flags = want ? LB2_FLAG_CIGAR : 0;
    t.qlen = qlen; t.tlen = tlen; t.query = query; t.target = target;
    t.w = w; t.h0 = 0; t.o_del = o_del; t.e_del = e_del; t.o_ins = o_ins; t.e_ins = e_ins;
    t.m = m; t.mat = mat;
    lb2_result r;
    cigar32_t* c = nullptr;
    run_one(t, &r, want ? &c : nullptr);
    if (want) { *n_cigar = r.n_cigar; *cigar = c; }
    return r.score;
}

int ksw_global(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
               int m, const int8_t* mat, int gapo, int gape, int w, int* n_cigar, cigar32_t** cigar)
{
    return ksw_global2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w, n_cigar, cigar);
}

int ksw_extend2(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                int m, const int8_t* mat, int o_del, int e_del, int o_ins, int e_ins,
                int w, int end_bonus, int zdrop, int h0,
                int* qle, int* tle, int* gtle, int* gscore, int* max_off)
{
    if (h0 <= 0) { fprintf(stderr, "[ksw_extend2] Assertion `h0 > 0' failed.\n"); abort(); }
    lb2_task t; memset(&t, 0, sizeof t);
    t.kind = LB2_KIND_EXTEND; t.flags = 0;
    t.qlen = qlen; t.tlen = tlen; t.query = query; t.target = target;
    t.w = w; t.h0 = h0; t.o_del = o_del; t.e_del = e_del; t.o_ins = o_ins; t.e_ins = e_ins;
    t.end_bonus = end_bonus; t.zdrop = zdrop; t.m = m; t.mat = mat;
    lb2_result r;
    run_one(t, &r, nullptr);
    if (qle) *qle = r.qle;
    if (tle) *tle = r.tle;
    if (gtle) *gtle = r.gtle;
    if (gscore) *gscore = r.gscore;
    if (max_off) *max_off = r.max_off;
    return r.score;
}

int ksw_extend(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
               int m, const int8_t* mat, int gapo, int gape, int w, int end_bonus, int zdrop, int h0,
               int* qle, int* tle, int* gtle, int* gscore, int* max_off)
{
    return ksw_extend2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w, end_bonus, zdrop, h0,
                       qle, tle, gtle, gscore, max_off);
}

int ksw_extend_core(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                    int m, const int8_t* mat, int w, int h0, lamsa_aln_para* AP,
                    int* _qle, int* _tle, cigar32_t** cigar_, int* n_cigar_, int* m_cigar_)
{
    if (qlen < 0 || tlen < 0) {
        fprintf(stderr, "[ksw_extend_core] Error: qlen: %d tlen: %d\n", qlen, tlen); exit(-1);
    }
    if (h0 <= 0) { fprintf(stderr, "[ksw_extend_core] Assertion `h0 > 0' failed.\n"); abort(); }
    const bool want = n_cigar_ && cigar_;
    lb2_task t; memset(&t, 0, sizeof t);
    t.kind = LB2_KIND_EXTEND; t.flags = want ? LB2_FLAG_CIGAR : 0;
    t.qlen = qlen; t.tlen = tlen; t.query = query; t.target = target;
    t.w = w; t.h0 = h0;
    t.o_del = AP->del_ext_o; t.e_del = AP->del_ext_e; t.o_ins = AP->ins_ext_o; t.e_ins = AP->ins_ext_e;
    t.end_bonus = AP->end_bonus; t.zdrop = AP->zdrop; t.m = m; t.mat = mat;
    lb2_result r;
    cigar32_t* c = nullptr;
    run_one(t, &r, want ? &c : nullptr);
    if (want) {                                  // src/ksw.c:781-803
        if (_qle) *_qle = r.qle;
        if (_tle) *_tle = r.tle;
        *n_cigar_ = r.n_cigar; *cigar_ = c; *m_cigar_ = r.reserved;
    }
    return r.score;
}

int ksw_extend_c(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                 int m, const int8_t* mat, int w, int h0, lamsa_aln_para* AP,
                 int* _qle, int* _tle, cigar32_t** cigar_, int* n_cigar_, int* m_cigar_)
{
    *n_cigar_ = *m_cigar_ = 0;
    ksw_extend_core(qlen, query, tlen, target, m, mat, w, h0, AP, _qle, _tle, cigar_, n_cigar_, m_cigar_);
    return *_qle == qlen ? 0 : *_tle == tlen ? 1 : 2;
}

int ksw_extend_r(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                 int m, const int8_t* mat, int w, int h0, lamsa_aln_para* AP,
                 int* _qre, int* _tre, cigar32_t** cigar_, int* n_cigar_, int* m_cigar_)
{
    *n_cigar_ = *m_cigar_ = 0;
    uint8_t* rq = (uint8_t*)malloc(qlen > 0 ? qlen : 1);
    uint8_t* rt = (uint8_t*)malloc(tlen > 0 ? tlen : 1);
    for (int a = 0; a < qlen; ++a) rq[a] = query[qlen - 1 - a];
    for (int a = 0; a < tlen; ++a) rt[a] = target[tlen - 1 - a];
    ksw_extend_core(qlen, rq, tlen, rt, m, mat, w, h0, AP, _qre, _tre, cigar_, n_cigar_, m_cigar_);
    free(rq); free(rt);
    return *_qre == qlen ? 0 : *_tre == tlen ? 1 : 2;
}

void sw_mid_fix(cigar32_t** cigar, int* cigar_n, int* cigar_m,
                cigar32_t* lcigar, int ln_cigar, cigar32_t* rcigar, int rn_cigar,
                const uint8_t* query, int qlen, int lqe, int rqe,
                const uint8_t* target, int tlen, int lte, int rte,
                lamsa_aln_para* AP, int m, const int8_t* mat)
{
    const int Sn = qlen - lqe - rqe, Hn = tlen - lte - rte, half = AP->split_len / 2;
    if (abs(Sn) >= half || abs(Hn) >= half || abs(Sn - Hn) >= half || tlen < 0 || qlen < 0) {
        list_append(cigar, cigar_n, cigar_m, lcigar, ln_cigar);
        list_add(cigar, cigar_n, cigar_m, (Sn << 4) | LB2_CSOFT_CLIP);
        list_add(cigar, cigar_n, cigar_m, Hn << 4 | LB2_CHARD_CLIP);
        list_append(cigar, cigar_n, cigar_m, rcigar, rn_cigar);
    } else {            // small leftovers: re-align the whole pair globally
        cigar32_t* g = nullptr; int gn = 0;
        ksw_global2(qlen, query, tlen, target, m, mat, AP->del_gapo, AP->del_gape,
                    AP->ins_gapo, AP->ins_gape, AP->band_w, &gn, &g);
        list_append(cigar, cigar_n, cigar_m, g, gn);
        free(g);
    }
}

int ksw_bi_extend(int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                  int m, const int8_t* mat, int lh0, int rh0, lamsa_aln_para* AP,
                  cigar32_t** cigar_, int* n_cigar_, int* m_cigar_)
{
    if (*n_cigar_) *n_cigar_ = 0;
    const int dl = abs(qlen - tlen);
    const int w = dl + 3 > AP->band_w ? dl + 3 : AP->band_w;                  // src/ksw.c:873
    // the reference compares an int with a float expression (:881, :900);
    // aln_mode_high_id_err() evaluates to 0 or 2 (src/lamsa_aln.h:438)
    const bool near = dl < AP->split_len + tlen * AP->id_rate * (AP->aln_mode & 2);

    int lqe = 0, lte = 0, ln = 0, lm = 0; cigar32_t* lc = nullptr;
    int res = ksw_extend_c(qlen, query, tlen, target, m, mat, w, lh0, AP, &lqe, &lte, &lc, &ln, &lm);
    if (res < 2) {
        *cigar_ = lc; *n_cigar_ = ln; *m_cigar_ = lm;
        list_add_nonempty(cigar_, n_cigar_, m_cigar_,
                          res == 0 ? (((tlen - lte) << 4) | LB2_CDEL) : (((qlen - lqe) << 4) | LB2_CINS));
        return 0;
    }
    if (near && ((lqe << 1 > qlen) || (lte << 1 > tlen))) {
        if (lc) free(lc);
        ksw_global2(qlen, query, tlen, target, m, mat, AP->del_gapo, AP->del_gape,
                    AP->ins_gapo, AP->ins_gape, AP->band_w, n_cigar_, cigar_);
        *m_cigar_ = *n_cigar_;
        return 0;
    }
    int rqe = 0, rte = 0, rn = 0, rm = 0; cigar32_t* rc = nullptr;
    res = ksw_extend_r(qlen, query, tlen, target, m, mat, w, rh0, AP, &rqe, &rte, &rc, &rn, &rm);
    if (res < 2) {
        list_add_nonempty(&rc, &rn, &rm,
                          res == 0 ? (((tlen - rte) << 4) | LB2_CDEL) : (((qlen - rqe) << 4) | LB2_CINS));
        list_reverse(rc, rn);
        *cigar_ = rc; *n_cigar_ = rn; *m_cigar_ = rm;
        free(lc);
        return 0;
    }
    if (near && ((rqe << 1 > qlen) || (rte << 1 > tlen))) {
        if (lc) free(lc);
        if (rc) free(rc);
        ksw_global2(qlen, query, tlen, target, m, mat, AP->del_gapo, AP->del_gape,
                    AP->ins_gapo, AP->ins_gape, AP->band_w, n_cigar_, cigar_);
        *m_cigar_ = *n_cigar_;
        return 0;
    }
    list_reverse(rc, rn);
    cigar32_t* out = (cigar32_t*)malloc(10 * sizeof(cigar32_t));             // :911-912
    int on = 0, om = 10;
    const int Sn = qlen - lqe - rqe;
    sw_mid_fix(&out, &on, &om, lc, ln, rc, rn, query, qlen, lqe, rqe, target, tlen, lte, rte, AP, m, mat);
    *cigar_ = out; *n_cigar_ = on; *m_cigar_ = om;
    if (lc) free(lc);
    if (rc) free(rc);
    return Sn >= AP->split_len ? 1 : 0;
}

}  // extern "C"
