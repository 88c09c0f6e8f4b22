// sdp_dropin.cu -- the reference's sparse-DP entry points (include/lamsa_b200.h section 1,
// "sparse-DP chaining entry points") on top of the batch interface: frag_line_BCC and
// frag_line_remain flatten the caller's map_msg / aln_reg into the arrays of lb2_sdp_*, run the
// chaining on the GPU (one read per call here; the batch interface takes thousands) and
// materialise the skeletons as frag_msg.  The node_score helpers that other reference files bind
// (src/bwt_aln.c:103-148, src/lamsa_aln.c:609) live here too: they operate on caller-owned host
// structs and never were part of the device path.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/lamsa_b200.h"
#include "ref_abi.h"
#include "dropin_internal.h"

namespace {

constexpr int kCigarLenM = 1000;               // CIGAR_LEN_M, src/lamsa_aln.h:172
constexpr int kUncovered = 0;                  // UNCOVERED, src/frag_check.h:234

// frag_init_msg, src/frag_check.c:19-37
void fmsg_init(lb2_ref_frag_msg* f) {
    f->frag_max = 1; f->frag_num = 0;
    f->fa_msg = (lb2_ref_frag_aln_msg*)malloc(sizeof(lb2_ref_frag_aln_msg));
    lb2_ref_frag_aln_msg* a = f->fa_msg;
    a->cigar_max = kCigarLenM; a->cigar = (int32_t*)malloc(kCigarLenM * sizeof(int32_t)); a->cigar_len = 0;
    a->seed_i = (int*)malloc(sizeof(int)); a->seed_aln_i = (int*)malloc(sizeof(int));
    a->seed_num = 0; a->seed_max = 1;
}
// frag_set_msg(FRAG_END), src/frag_check.c:58-83: open fragment frag_i with its first seed
void fmsg_open(lb2_ref_frag_msg* f, const lb2_ref_map_msg* m_msg, int seed_i, int aln_i, int frag_i) {
    if (f->frag_num == f->frag_max) {
        f->frag_max <<= 1;
        f->fa_msg = (lb2_ref_frag_aln_msg*)realloc(f->fa_msg, f->frag_max * sizeof(lb2_ref_frag_aln_msg));
        for (int i = f->frag_num; i < f->frag_max; ++i) {
            lb2_ref_frag_aln_msg* a = f->fa_msg + i;
            a->cigar_max = kCigarLenM; a->cigar = (int32_t*)malloc(kCigarLenM * sizeof(int32_t)); a->cigar_len = 0;
            a->seed_i = (int*)malloc(sizeof(int)); a->seed_aln_i = (int*)malloc(sizeof(int));
            a->seed_num = 0; a->seed_max = 1;
        }
    }
    lb2_ref_frag_aln_msg* a = f->fa_msg + frag_i;
    a->chr = m_msg[seed_i].map[aln_i].nchr;
    a->strand = m_msg[seed_i].map[aln_i].nstrand;
    a->seed_i[0] = seed_i; a->seed_aln_i[0] = aln_i;
    a->seed_num = 1; a->flag = kUncovered; a->cigar_len = 0;
}
// frag_set_msg(FRAG_SEED), :86-96
void fmsg_add_seed(lb2_ref_frag_msg* f, int seed_i, int aln_i, int frag_i) {
    lb2_ref_frag_aln_msg* a = f->fa_msg + frag_i;
    if (a->seed_num == a->seed_max) {
        a->seed_max <<= 1;
        a->seed_i = (int*)realloc(a->seed_i, a->seed_max * sizeof(int));
        a->seed_aln_i = (int*)realloc(a->seed_aln_i, a->seed_max * sizeof(int));
    }
    a->seed_i[a->seed_num] = seed_i; a->seed_aln_i[a->seed_num] = aln_i;
    a->seed_num++;
}

// skeleton stream -> frag_msg array, in the call order of frag_dp_path (src/lamsa_dp_con.c:1177-1233)
int stream_to_fmsg(const int32_t* w, int64_t nw, const lb2_ref_map_msg* m_msg, int seed_all, lb2_ref_frag_msg** f_msg) {
    if (nw < 1) return 0;
    const int line_n = w[0];
    if (line_n == 0) return 0;
    *f_msg = (lb2_ref_frag_msg*)malloc(line_n * sizeof(lb2_ref_frag_msg));
    for (int l = 0; l < line_n; ++l) fmsg_init(*f_msg + l);
    int64_t p = 1;
    for (int l = 0; l < line_n; ++l) {
        lb2_ref_frag_msg* f = *f_msg + l;
        const int line_score = w[p++], frag_num = w[p++];
        for (int k = 0; k < frag_num; ++k) {
            const int seed_num = w[p++];
            fmsg_open(f, m_msg, w[p], w[p + 1], k);
            if (k == 0) f->frag_right_bound = seed_all + 1;
            for (int s = 1; s < seed_num; ++s) fmsg_add_seed(f, w[p + 2 * s], w[p + 2 * s + 1], k);
            p += 2 * seed_num;
            f->frag_num = k + 1;                               // FRAG_START, :84-85
        }
        f->frag_left_bound = 0;
        f->line_score = line_score;
    }
    return line_n;
}

void para_from(const lamsa_aln_para* AP, lb2_sdp_para* P) {
    memset(P, 0, sizeof *P);
    P->seed_len = AP->seed_len; P->seed_step = AP->seed_step; P->seed_inv = AP->seed_inv;
    P->per_aln_m = AP->per_aln_m; P->first_loci_thd = AP->first_loci_thd; P->SV_len_thd = AP->SV_len_thd;
    P->ske_max = AP->ske_max; P->ovlp_rat = AP->ovlp_rat; P->split_len = AP->split_len;
    P->match_dis = AP->match_dis; P->mismatch_thd = AP->mismatch_thd; P->aln_mode = AP->aln_mode;
    P->bwt_seed_len = AP->bwt_seed_len;
    for (int i = 0; i < 10; ++i) P->frag_score_table[i] = AP->frag_score_table[i];
}

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "[lamsa_b200] %s: %s\n", what, lb2_last_error());
    exit(1);
}

// Per-worker state, keyed by the worker's f_node scratch pointer (one per reference worker,
// src/lamsa_aln.c:971): the read's flattened hits and the tracked flags stage 1 leaves for stage 2.
std::mutex g_mu;
std::unordered_map<void*, lb2::SdpWorkerState*> g_workers;
lb2::SdpWorkerState* worker_state(void* key, bool create) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_workers.find(key);
    if (it != g_workers.end()) return it->second;
    if (!create) return nullptr;
    return g_workers[key] = new lb2::SdpWorkerState();
}

// aln_sort_reg + aln_merg_reg on the caller's struct (src/lamsa_aln.c:476-519), the in-place side
// effect of get_remain_reg (:558) that later stages of the read pipeline observe
void push_pts(lb2_ref_reg* r, int bn, const lb2_ref_reg_b* b, int en, const lb2_ref_reg_b* e) {
    for (int i = 0; i < bn; ++i) {
        if (r->beg_n == r->beg_m) { r->beg_m <<= 1; r->ref_beg = (lb2_ref_reg_b*)realloc(r->ref_beg, r->beg_m * sizeof(lb2_ref_reg_b)); }
        r->ref_beg[r->beg_n++] = b[i];
    }
    for (int i = 0; i < en; ++i) {
        if (r->end_n == r->end_m) { r->end_m <<= 1; r->ref_end = (lb2_ref_reg_b*)realloc(r->ref_end, r->end_m * sizeof(lb2_ref_reg_b)); }
        r->ref_end[r->end_n++] = e[i];
    }
}
void sort_merge_regions(lb2_ref_aln_reg* a, int thd) {
    // ties (equal `beg`) keep input order: see the note on qsort's tie order in sdp_batch.cu (remaining_regions)
    std::stable_sort(a->reg, a->reg + a->reg_n, [](const lb2_ref_reg& x, const lb2_ref_reg& y) { return x.beg < y.beg; });
    int cur = 0;
    for (int i = 1; i < a->reg_n; ++i) {
        if (a->reg[i].beg - a->reg[cur].end - 1 < thd) {
            if (a->reg[i].end > a->reg[cur].end) a->reg[cur].end = a->reg[i].end;
            push_pts(a->reg + cur, a->reg[i].beg_n, a->reg[i].ref_beg, a->reg[i].end_n, a->reg[i].ref_end);
        } else {
            ++cur;
            if (cur != i) {
                a->reg[cur].beg = a->reg[i].beg; a->reg[cur].end = a->reg[i].end;
                a->reg[cur].beg_n = a->reg[cur].end_n = 0;
                push_pts(a->reg + cur, a->reg[i].beg_n, a->reg[i].ref_beg, a->reg[i].end_n, a->reg[i].ref_end);
            }
        }
    }
    a->reg_n = cur + 1;
}

// serve a request: parked for the fiber scheduler's next batch, or run alone right now
void serve(lb2::SdpRequest* r) {
    if (lb2::fiber_active()) { lb2::fiber_wait_sdp(r); return; }
    std::vector<lb2::SdpRequest*> one{r};
    lb2::dropin_submit_sdp(one);
}

}  // namespace

// All requests of `batch` (same stage, same parameters) as ONE GPU batch.  Each OS thread keeps one
// grow-only batch object, so steady state allocates nothing.
// where the time of the chaining batches goes (LB2_FIBER_STATS): seconds of flattening the requests, of loading the
// batch object (H2D), of the stage (kernels + read-back; the kernels' own event time beside it) and of handing back
namespace lb2 { std::atomic<long long> g_sdp_ns[5] = {}; }
void lb2::dropin_submit_sdp(std::vector<lb2::SdpRequest*>& batch) {
    if (batch.empty()) return;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ns = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
    const auto t_begin = now();
    // one grow-only batch object per context (= per scheduler thread, or the process-wide one, which the
    // reference's pthread workers share: those calls are serialised here, like the DP calls in ksw_dropin.cu)
    static std::mutex shared_ctx_mu;
    std::unique_lock<std::mutex> serial(shared_ctx_mu, std::defer_lock);
    if (!lb2::dropin_has_thread_ctx()) serial.lock();
    static std::mutex batch_mu;
    static std::unordered_map<lb2_ctx*, lb2_sdp_batch*> batch_of;
    lb2_ctx* const my_ctx = lb2::dropin_ctx();
    lb2_sdp_batch* tl_batch = nullptr;
    {
        std::lock_guard<std::mutex> lk(batch_mu);
        auto it = batch_of.find(my_ctx);
        if (it != batch_of.end()) tl_batch = it->second;
    }
    const int stage = batch[0]->stage;
    const int64_t n = (int64_t)batch.size();
    std::vector<lb2_sdp_read> reads((size_t)n);
    std::vector<int32_t> sid, mn; std::vector<lb2_sdp_hit> hits; std::vector<lb2_sdp_reg> regs; std::vector<uint8_t> flags;
    // a batch may consist of reads without a single seed hit (most reads of a 15 %-error set at default seeding):
    // the arrays are empty then, but their pointers must not be NULL
    sid.reserve(16); mn.reserve(16); hits.reserve(16); regs.reserve(16); flags.reserve(16);
    for (int64_t i = 0; i < n; ++i) {
        const lb2::SdpRequest* q = batch[(size_t)i];
        const lb2::SdpWorkerState* ws = q->ws;
        if (q->stage != stage || memcmp(&ws->para, &batch[0]->ws->para, sizeof(lb2_sdp_para))) {
            fprintf(stderr, "[lamsa_b200] chaining requests of one batch must share stage and parameters\n"); exit(1);
        }
        lb2_sdp_read rd = ws->read;
        rd.seed_first = (int64_t)sid.size(); rd.hit_first = (int64_t)hits.size();
        rd.reg_first = (int64_t)regs.size(); rd.n_reg = (int32_t)q->regs.size();
        reads[(size_t)i] = rd;
        sid.insert(sid.end(), ws->seed_id.begin(), ws->seed_id.end());
        mn.insert(mn.end(), ws->map_n.begin(), ws->map_n.end());
        hits.insert(hits.end(), ws->hits.begin(), ws->hits.end());
        regs.insert(regs.end(), q->regs.begin(), q->regs.end());
        if (stage == 2) flags.insert(flags.end(), ws->tracked.begin(), ws->tracked.end());
    }
    const lb2_sdp_para* P = &batch[0]->ws->para;
    const auto t_flat = now();
    if (!tl_batch) {
        if (lb2_sdp_create(my_ctx, P, n, reads.data(), sid.data(), mn.data(), hits.data(), &tl_batch)) die("chaining batch");
        std::lock_guard<std::mutex> lk(batch_mu);
        batch_of[my_ctx] = tl_batch;
    }
    else if (lb2_sdp_reset(tl_batch, P, n, reads.data(), sid.data(), mn.data(), hits.data())) die("chaining batch");
    const int32_t* w; const int64_t* off;
    const auto t_load = now();
    float kms = 0.f;
    if (stage == 1) {
        if (lb2_sdp_run_bcc(tl_batch, &w, &off, &kms)) die("frag_line_BCC");
        flags.resize(hits.size());
        if (lb2_sdp_get_tracked(tl_batch, flags.data())) die("frag_line_BCC");
    } else {
        if (lb2_sdp_set_tracked(tl_batch, flags.data())) die("frag_line_remain");
        if (lb2_sdp_run_remain(tl_batch, reads.data(), regs.data(), &w, &off, &kms)) die("frag_line_remain");
    }
    const auto t_run = now();
    for (int64_t i = 0; i < n; ++i) {
        lb2::SdpRequest* q = batch[(size_t)i];
        q->stream.assign(w + off[i], w + off[i + 1]);
        if (stage == 1) {
            const size_t h0 = (size_t)reads[(size_t)i].hit_first, hn = q->ws->hits.size();
            q->ws->tracked.assign(flags.begin() + h0, flags.begin() + h0 + hn);
        }
    }
    lb2::g_sdp_ns[0] += ns(t_begin, t_flat); lb2::g_sdp_ns[1] += ns(t_flat, t_load); lb2::g_sdp_ns[2] += ns(t_load, t_run);
    lb2::g_sdp_ns[3] += (long long)(kms * 1e6); lb2::g_sdp_ns[4] += ns(t_run, now());
}

extern "C" int frag_line_BCC(lb2_ref_map_msg* m_msg, lb2_ref_frag_msg** f_msg, lb2_ref_per_para* APP, lamsa_aln_para* AP,
                             lb2_ref_kseq* seqs, lb2_ref_line_node*, int*, int*, int*, void*** f_node, lb2_ref_line_node*, int) {
    lb2::SdpWorkerState* ws = worker_state((void*)f_node, true);
    para_from(AP, &ws->para);
    const int S = APP->seed_out;
    ws->seed_id.resize(S); ws->map_n.resize(S); ws->hits.clear(); ws->tracked.clear();
    for (int i = 0; i < S; ++i) {
        ws->seed_id[i] = m_msg[i].seed_id; ws->map_n[i] = m_msg[i].map_n;
        for (int j = 0; j < m_msg[i].map_n; ++j) {
            const lb2_ref_map& m = m_msg[i].map[j];
            ws->hits.push_back(lb2_sdp_hit{m.offset, m.nchr, m.NM, m.len_dif, m.nstrand});
        }
    }
    ws->read = lb2_sdp_read{S, APP->seed_all, (int32_t)seqs->seq.l, 0, 0, 0, 0};
    lb2::SdpRequest req{1, ws, {}, {}};
    serve(&req);
    return stream_to_fmsg(req.stream.data(), (int64_t)req.stream.size(), m_msg, APP->seed_all, f_msg);
}

extern "C" int frag_line_remain(lb2_ref_aln_reg* a_reg, lb2_ref_map_msg* m_msg, lb2_ref_frag_msg** f_msg, lb2_ref_per_para* APP,
                                lamsa_aln_para* AP, lb2_ref_kseq* seqs, lb2_ref_line_node*, int*, int*, int*, void*** f_node,
                                lb2_ref_line_node*, int*, int*, int) {
    lb2::SdpWorkerState* ws = worker_state((void*)f_node, false);
    if (!ws || ws->tracked.size() != ws->hits.size()) {
        fprintf(stderr, "[lamsa_b200] frag_line_remain without a preceding frag_line_BCC on this worker\n"); exit(1);
    }
    lb2::SdpRequest req{2, ws, {}, {}};
    for (int k = 0; k < a_reg->reg_n; ++k) {
        const lb2_ref_reg& g = a_reg->reg[k];
        // get_reg (src/lamsa_aln.c:608-616) hands one begin and one end point per record; records that a
        // caller already merged are passed as one record per point pair
        // a record without a begin or without an end point has no interval on the reference to report (push_reg allows
        // such records; get_reg never produces them): it takes part in the in-place merge below, not in the request
        const int n = (g.beg_n > 0 && g.end_n > 0) ? std::max(g.beg_n, g.end_n) : 0;
        for (int t = 0; t < n; ++t) {
            const lb2_ref_reg_b& pb = g.ref_beg[std::min(t, g.beg_n - 1)];
            const lb2_ref_reg_b& pe = g.ref_end[std::min(t, g.end_n - 1)];
            req.regs.push_back(lb2_sdp_reg{g.beg, g.end, pb.chr, pb.is_rev, pb.ref_pos, pe.ref_pos});
        }
    }
    if (a_reg->reg_n > 0) sort_merge_regions(a_reg, AP->bwt_seed_len);
    serve(&req);
    return stream_to_fmsg(req.stream.data(), (int64_t)req.stream.size(), m_msg, APP->seed_all, f_msg);
}

// ---- node_score helpers (src/lamsa_dp_con.c:29-67, src/lamsa_heap.c) ------------------------
extern "C" lb2_ref_node_score* node_init_score(int n) {
    lb2_ref_node_score* ns = (lb2_ref_node_score*)malloc(sizeof(lb2_ref_node_score));
    ns->max_n = n; ns->node_n = 0;
    ns->node = (lb2_ref_line_node*)malloc(n * sizeof(lb2_ref_line_node));
    ns->score = (int*)malloc(n * sizeof(int));
    ns->NM = (int*)malloc(n * sizeof(int));
    return ns;
}
extern "C" void node_free_score(lb2_ref_node_score* ns) { free(ns->score); free(ns->NM); free(ns->node); free(ns); }
extern "C" float cover_rate(int s1, int e1, int s2, int e2) {
    const int s = s2 > s1 ? s2 : s1, e = e2 < e1 ? e2 : e1;
    const float r1 = (e - s + 1 + 0.0) / (e1 - s1 + 1 + 0.0), r2 = (e - s + 1 + 0.0) / (e2 - s2 + 1 + 0.0);
    return r1 > r2 ? r1 : r2;
}
namespace {
inline void swap_all(lb2_ref_node_score* h, int a, int b, bool with_nm) {
    std::swap(h->node[a], h->node[b]); std::swap(h->score[a], h->score[b]);
    if (with_nm) std::swap(h->NM[a], h->NM[b]);
}
template <class Less> void sift(lb2_ref_node_score* h, int i, Less less, bool with_nm) {
    for (;;) {
        const int l = 2 * i + 1, r = 2 * i + 2;
        int m = i;
        if (l < h->node_n && less(l, i)) m = l;
        if (r < h->node_n && less(r, m)) m = r;
        if (m == i) return;
        swap_all(h, i, m, with_nm); i = m;
    }
}
void sift_min(lb2_ref_node_score* h, int i) {
    sift(h, i, [h](int a, int b) { return h->score[a] < h->score[b] || (h->score[a] == h->score[b] && h->NM[a] > h->NM[b]); }, true);
}
void sift_max(lb2_ref_node_score* h, int i) {       // src/lamsa_heap.c:16-33 leaves NM in place when it swaps
    sift(h, i, [h](int a, int b) { return h->score[a] > h->score[b] || (h->score[a] == h->score[b] && h->NM[a] < h->NM[b]); }, false);
}
void sift_minpos(lb2_ref_node_score* h, int i) {
    sift(h, i, [h](int a, int b) { return h->node[a].x < h->node[b].x; }, true);
}
}  // namespace
extern "C" void build_node_min_heap(lb2_ref_node_score* ns) { for (int i = (ns->node_n - 1) / 2; i >= 0; --i) sift_min(ns, i); }
extern "C" void build_node_max_heap(lb2_ref_node_score* ns) { if (ns->node_n == 0) return; for (int i = (ns->node_n - 1) / 2; i >= 0; --i) sift_max(ns, i); }
extern "C" void build_node_minpos_heap(lb2_ref_node_score* ns) { for (int i = (ns->node_n - 1) / 2; i >= 0; --i) sift_minpos(ns, i); }
extern "C" lb2_ref_line_node node_pop(lb2_ref_node_score* ns, int* score, int* NM) {
    if (ns->node_n < 1) return lb2_ref_line_node{-1, 0};
    --ns->node_n;
    *score = ns->score[ns->node_n]; *NM = ns->NM[ns->node_n];
    return ns->node[ns->node_n];
}
extern "C" lb2_ref_line_node node_heap_extract_max(lb2_ref_node_score* ns, int* score) {
    if (ns->node_n < 1) return lb2_ref_line_node{-1, 0};
    const lb2_ref_line_node top = ns->node[0];
    *score = ns->score[0];
    --ns->node_n;
    ns->node[0] = ns->node[ns->node_n]; ns->score[0] = ns->score[ns->node_n]; ns->NM[0] = ns->NM[ns->node_n];
    sift_max(ns, 0);
    return top;
}
extern "C" lb2_ref_line_node node_heap_extract_minpos(lb2_ref_node_score* ns) {
    if (ns->node_n < 1) return lb2_ref_line_node{-1, 0};
    const lb2_ref_line_node top = ns->node[0];
    --ns->node_n;
    ns->node[0] = ns->node[ns->node_n]; ns->score[0] = ns->score[ns->node_n]; ns->NM[0] = ns->NM[ns->node_n];
    sift_minpos(ns, 0);
    return top;
}
extern "C" int node_heap_update_min(lb2_ref_node_score* ns, lb2_ref_line_node node, int score, int NM) {
    if (ns->score[0] < score || (ns->score[0] == score && ns->NM[0] > NM)) {
        const int ret = ns->node[0].x;
        ns->score[0] = score; ns->NM[0] = NM; ns->node[0] = node;
        sift_min(ns, 0);
        return ret;
    }
    return -2;
}
extern "C" int heap_add_node(lb2_ref_node_score* ns, lb2_ref_line_node node, int score, int NM) {
    if (ns->node_n < ns->max_n) {
        ns->score[ns->node_n] = score; ns->NM[ns->node_n] = NM; ns->node[ns->node_n++] = node;
        if (ns->node_n == ns->max_n) build_node_min_heap(ns);
        return -1;
    }
    return node_heap_update_min(ns, node, score, NM);
}

// ---- layout self-description, compared with the reference headers by tests/test_abi.py ------
// (same order as oracle/sdp_ref_shim.c:ref_sdp_offsets)
extern "C" int lb2_ref_abi_offsets(int* o) {
    int n = 0;
#define OFF(T, f) (int)offsetof(T, f)
    o[n++] = OFF(lb2_ref_map, nstrand); o[n++] = OFF(lb2_ref_map, nchr); o[n++] = OFF(lb2_ref_map, offset);
    o[n++] = OFF(lb2_ref_map, NM); o[n++] = OFF(lb2_ref_map, len_dif);
    o[n++] = OFF(lb2_ref_map_msg, map); o[n++] = OFF(lb2_ref_map_msg, map_n); o[n++] = OFF(lb2_ref_map_msg, seed_id);
    o[n++] = OFF(lb2_ref_frag_msg, frag_max); o[n++] = OFF(lb2_ref_frag_msg, frag_num); o[n++] = OFF(lb2_ref_frag_msg, fa_msg);
    o[n++] = OFF(lb2_ref_frag_msg, line_score); o[n++] = OFF(lb2_ref_frag_msg, frag_left_bound); o[n++] = OFF(lb2_ref_frag_msg, frag_right_bound);
    o[n++] = OFF(lb2_ref_frag_aln_msg, chr); o[n++] = OFF(lb2_ref_frag_aln_msg, strand); o[n++] = OFF(lb2_ref_frag_aln_msg, cigar);
    o[n++] = OFF(lb2_ref_frag_aln_msg, cigar_len); o[n++] = OFF(lb2_ref_frag_aln_msg, cigar_max); o[n++] = OFF(lb2_ref_frag_aln_msg, flag);
    o[n++] = OFF(lb2_ref_frag_aln_msg, seed_max); o[n++] = OFF(lb2_ref_frag_aln_msg, seed_num); o[n++] = OFF(lb2_ref_frag_aln_msg, seed_i);
    o[n++] = OFF(lb2_ref_frag_aln_msg, seed_aln_i);
    o[n++] = OFF(lb2_ref_per_para, seed_all); o[n++] = OFF(lb2_ref_per_para, seed_out);
    o[n++] = OFF(lb2_ref_aln_reg, reg); o[n++] = OFF(lb2_ref_aln_reg, reg_n); o[n++] = OFF(lb2_ref_aln_reg, reg_m); o[n++] = OFF(lb2_ref_aln_reg, read_len);
    o[n++] = OFF(lb2_ref_reg, ref_beg); o[n++] = OFF(lb2_ref_reg, ref_end); o[n++] = OFF(lb2_ref_reg, beg_n); o[n++] = OFF(lb2_ref_reg, end_n);
    o[n++] = OFF(lb2_ref_reg, beg_m); o[n++] = OFF(lb2_ref_reg, end_m); o[n++] = OFF(lb2_ref_reg, beg); o[n++] = OFF(lb2_ref_reg, end);
    o[n++] = OFF(lb2_ref_reg_b, is_rev); o[n++] = OFF(lb2_ref_reg_b, chr); o[n++] = OFF(lb2_ref_reg_b, ref_pos);
    o[n++] = OFF(lb2_ref_kseq, seq) + OFF(lb2_ref_kstring, l);
    o[n++] = OFF(lb2_ref_node_score, node); o[n++] = OFF(lb2_ref_node_score, score); o[n++] = OFF(lb2_ref_node_score, NM);
    o[n++] = OFF(lb2_ref_node_score, min_score_thd); o[n++] = OFF(lb2_ref_node_score, max_n); o[n++] = OFF(lb2_ref_node_score, node_n);
    o[n++] = kCigarLenM; o[n++] = kUncovered;
#undef OFF
    return n;
}
// sizes in the order of oracle/sdp_ref_shim.c:ref_sdp_sizes (frag_dp_node, 80 bytes, is opaque here)
extern "C" int lb2_ref_abi_sizes(int* o) {
    int n = 0;
    o[n++] = (int)sizeof(lb2_ref_map); o[n++] = (int)sizeof(lb2_ref_map_msg); o[n++] = 80;
    o[n++] = (int)sizeof(lb2_ref_frag_msg); o[n++] = (int)sizeof(lb2_ref_frag_aln_msg); o[n++] = (int)sizeof(lb2_ref_per_para);
    o[n++] = (int)sizeof(lb2_ref_aln_reg); o[n++] = (int)sizeof(lb2_ref_reg); o[n++] = (int)sizeof(lb2_ref_reg_b);
    o[n++] = (int)sizeof(lb2_ref_kseq); o[n++] = (int)sizeof(lb2_ref_line_node); o[n++] = (int)sizeof(lb2_ref_node_score);
    return n;
}
