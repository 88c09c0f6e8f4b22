// sdp_kernel.cuh -- sparse-DP anchor chaining on the GPU (sm_100a): one warp per read.
//
// Reference semantics being reproduced (paths relative to the reference tree):
//   frag_line_BCC     src/lamsa_dp_con.c:1305-1445   stage 1, all seed hits of a read
//   frag_line_remain  src/lamsa_dp_con.c:1252-1302   stage 2, inside the unaligned read regions
// and everything below them: edge classification get_fseed_dis :596, node set-up :636-681/:766,
// predecessor scan frag_dp_update :701, tree pruning :786-920, gap refill frag_mini_dp_line :1068,
// region chaining frag_mini_dp_multi_line :923, skeleton clustering :12-494, overlap filter :568,
// path -> fragments frag_dp_path :1152, the array heaps of src/lamsa_heap.c.
//
// Execution model.  A read is owned by one warp.
//  * The O(hits^2) part -- the predecessor scan of every node -- runs ACROSS the 32 lanes: the
//    predecessors of a node are one contiguous range of a per-read "scan order" array (seeds
//    descending, hits ascending = the order of :713-714), lanes classify 32 edges at a time and
//    the winner is picked with warp reductions that reproduce the sequential rule exactly
//    (strictly better score, then smaller NM, then first in scan order; a '-' strand match-like
//    predecessor ends the scan at once, :726-733).  Node initialisation and the co-linear
//    promotion of :1031 are lane-parallel too (independent per node / per seed).
//  * Everything else (tree pruning, gap refill control, clustering, fragment emission) is
//    pointer-chasing with sequential dependences.  It is executed by ALL lanes redundantly and
//    warp-uniformly: every lane computes the same values and issues the same stores (same
//    address, same value -> one transaction), so each lane always reads its own writes and no
//    intra-warp synchronisation is needed except after the lane-partitioned loops.
//  * Node order, every tie rule and every side effect follow the reference; node state lives in
//    per-read scratch in HBM (structure of arrays over the read's hits, L1/L2 resident while the
//    warp works on it).  Son lists are intrusive linked lists (head/tail/next-sibling, bounded by
//    son_n) instead of the reference's realloc'd arrays; insertion order is preserved.
//  * The extend-boundary slots E_LB/E_RB (:450-491,:520-563) are dead stores in the reference
//    (never read anywhere) and are not computed.
#pragma once
#include <cstdint>
#include <climits>
#include <cuda_runtime.h>
#include "../../include/lamsa_b200.h"

namespace lb2 {
namespace sdp {

// edge kinds (src/lamsa_aln.h:103-121), node states (src/split_mapping.h:60-67), skeleton flags (:124-150)
enum { E_MATCH = 0, E_MISMATCH = 2, E_MATCH_THD = 2, E_LONG_MISMATCH = 3, E_INSERT = 4, E_DELETE = 5,
       E_CHR_DIF = 6, E_UNCONNECT = 8, E_INIT = 20 };
enum { ST_MIN = 1, ST_MULTI = 2, ST_WHOLE = 4, ST_TRACKED = 5 };
enum { SK_MERGB = 0, SK_NMERG = 1, SK_MERGH = 2, SK_INTER = 4, SK_DUMP = 8, SK_EXTRA = 5 };
enum { ERR_NONE = 0, ERR_PATH = 1, ERR_STREAM = 2, ERR_EDGE = 3, ERR_STACK = 4 };

// node state arrays (each n_hits ints)
enum { A_SON_FLAG, A_FROM, A_IN_DE, A_SON_N, A_SON_HEAD, A_SON_TAIL, A_SIB, A_MAX_SCORE, A_MAX_NM, A_MAX_NODE,
       A_SCORE, A_TOL_NM, A_MATCH_FLAG, A_DP_FLAG, A_NODE_N, A_COUNT };

constexpr unsigned kFullMask = 0xffffffffu;

struct DRead {
    int32_t seed_out, seed_all, read_len, n_hits;
    int32_t n_region, pad;
    int64_t seed_base;       // seed_id / map_n
    int64_t hoff_base;       // hoff (seed_out + 1 entries)
    int64_t hit_base;        // hits / hseed / rflat
    int64_t scratch;         // first word of this read's scratch
    int64_t region_first;    // stage 2: first region of this read
};
struct DRegion { int32_t beg, end, bn, en; int64_t pt_first; };      // bn begin points, then en end points
struct DPoint { int64_t pos; int32_t chr; int32_t pad; };
struct DOut { int32_t n_words, err; int64_t pairs; };

// per-read scratch, in 4-byte words relative to DRead::scratch
struct Layout {
    int64_t node, line, tline, mini, lsl, tlsl, rank, trank, srank, stack, kheap, clu, trg, tri, out, total;
    int32_t line_cap, out_cap;
};
__host__ __device__ inline Layout make_layout(int H, int S, int K) {
    Layout L;
    int64_t o = 0;
    // every array starts on a 16-byte boundary (int2 / int4 views)
#define LB2_TAKE(field, words) do { L.field = o; o += ((int64_t)(words) + 3) & ~int64_t(3); } while (0)
    L.line_cap = 6 * H + 16 + S;
    L.out_cap = 5 * H + 8;
    LB2_TAKE(node, (int64_t)A_COUNT * H);
    LB2_TAKE(line, 2 * (int64_t)L.line_cap);
    LB2_TAKE(tline, 2 * (int64_t)L.line_cap);
    LB2_TAKE(mini, 2 * (int64_t)(S + 2));
    LB2_TAKE(lsl, 2 * (int64_t)H + 2);
    LB2_TAKE(tlsl, 2 * (int64_t)H + 2);
    LB2_TAKE(rank, H + 1);
    LB2_TAKE(trank, H + 1);
    LB2_TAKE(srank, H + 1);
    LB2_TAKE(stack, 3 * (int64_t)(H + 2));
    LB2_TAKE(kheap, 3 * (int64_t)(K + 1));
    LB2_TAKE(clu, 10 * (int64_t)(H + 1) + 4);   // cluster members (3), cluster starts, kept lists (2), kept starts, counts; sort keys (2)
    LB2_TAKE(trg, 4 * (int64_t)(2 * H + 2));
    LB2_TAKE(tri, 2 * (int64_t)(H + 1));
    LB2_TAKE(out, L.out_cap);
#undef LB2_TAKE
    L.total = o;
    return L;
}

struct Ctx {
    const lb2_sdp_para* P;
    int lane, S, H, seed_all, read_len, K;
    const int32_t *seed_id, *map_n, *hoff, *hseed, *rflat;
    const lb2_sdp_hit* hit;
    int* nd;
    int2 *line, *tline, *mini;
    int *lsl, *tlsl, *rank, *trank, *srank;
    int *stk; int stk_n, stk_thd;
    int *kh; int kh_n;
    int *clu; int4* trg; int *tri_start, *tri_n;
    int *out; int out_n, out_cap;
    long long pairs;
    int err;
};

#define NA(c, a, p) ((c).nd[(size_t)(a) * (c).H + (p)])

__device__ __forceinline__ int flat_of(const Ctx& c, int2 a) { return a.x < 0 ? -1 : c.hoff[a.x] + a.y; }
__device__ __forceinline__ int2 xy_of(const Ctx& c, int p) {
    if (p < 0) return make_int2(-1, 0);
    const int x = c.hseed[p];
    return make_int2(x, p - c.hoff[x]);
}
__device__ __forceinline__ int seed_of(const Ctx& c, int p) { return p < 0 ? -1 : c.hseed[p]; }

// ------------------------------------------------------ edge classification --
// get_fseed_dis, src/lamsa_dp_con.c:596-634.  pf -> cf are flat node ids (-1 = START_NODE).
__device__ __forceinline__ int edge_kind_hits(const lb2_sdp_para& P, int pf, int cf, int sp, int si,
                                              const lb2_sdp_hit& hp, const lb2_sdp_hit& hi) {
    if (sp == si) return pf == cf ? E_MATCH : E_UNCONNECT;
    if (hi.nchr != hp.nchr || hi.nstrand != hp.nstrand) return E_CHR_DIF;
    const int ds = abs(sp - si);
    if (ds * P.seed_step < P.seed_len) return E_UNCONNECT;
    const int64_t expct = hp.offset + hp.nstrand * (si - sp) * P.seed_step;
    const int64_t act = hi.offset;
    const int dis = (int)(hp.nstrand * ((sp < si) ? (act - expct) : (expct - act))
                          - ((hp.nstrand * (sp - si) < 0) ? hp.len_dif : hi.len_dif));
    const int mat_dis = P.match_dis * ((P.aln_mode & 2) ? ds : 1);
    if (dis <= mat_dis && dis >= -mat_dis) {
        if (ds == 1) return E_MATCH;
        if (ds <= 3 * P.mismatch_thd) return E_MISMATCH;
        return E_LONG_MISMATCH;
    }
    if (dis > mat_dis && dis < P.SV_len_thd) return E_DELETE;
    if ((dis < -mat_dis && dis >= 0 - (ds * P.seed_step - P.seed_len)) ||
        (dis < -(P.split_len / 2) && dis >= -P.SV_len_thd)) return E_INSERT;
    return E_UNCONNECT;
}
__device__ __forceinline__ lb2_sdp_hit load_hit(const lb2_sdp_hit* h) {
    const int2* q = reinterpret_cast<const int2*>(h);        // 24-byte records, 8-byte aligned
    const int2 a = q[0], b = q[1], d = q[2];
    lb2_sdp_hit r;
    r.offset = (int64_t)(((uint64_t)(uint32_t)a.y << 32) | (uint32_t)a.x);
    r.nchr = b.x; r.NM = b.y; r.len_dif = d.x; r.nstrand = d.y;
    return r;
}
__device__ __forceinline__ int edge_kind(const Ctx& c, int pf, int cf) {
    if (pf < 0 || cf < 0) return E_MATCH;
    const int sp = c.hseed[pf], si = c.hseed[cf];
    return edge_kind_hits(*c.P, pf, cf, c.seed_id[sp], c.seed_id[si], load_hit(c.hit + pf), load_hit(c.hit + cf)) ;
}
// NB: the reference compares seed INDICES for "same seed" (pre == i) and seed IDS for distances; indices and
// ids are both strictly increasing, so equality of one is equality of the other.

// ------------------------------------------------------------- node set-up --
// fnode_set, :636-655
__device__ __forceinline__ void node_set(Ctx& c, int p, int from, int score, int NM, int match_flag, int dp_flag) {
    NA(c, A_SON_FLAG, p) = E_INIT; NA(c, A_FROM, p) = from; NA(c, A_SCORE, p) = score; NA(c, A_TOL_NM, p) = NM;
    NA(c, A_MATCH_FLAG, p) = match_flag; NA(c, A_DP_FLAG, p) = dp_flag;
    NA(c, A_NODE_N, p) = 1; NA(c, A_IN_DE, p) = 0; NA(c, A_SON_N, p) = 0;
    NA(c, A_MAX_SCORE, p) = score; NA(c, A_MAX_NM, p) = NM; NA(c, A_MAX_NODE, p) = p;
}
// frag_dp_per_init, :766-784
__device__ __forceinline__ void node_init(Ctx& c, int p, int from, int dp_flag) {
    if (from < 0) { node_set(c, p, from, 1, c.hit[p].NM, E_MATCH, dp_flag); return; }
    const int k = edge_kind(c, from, p);
    if (k != E_UNCONNECT && k != E_CHR_DIF)
        node_set(c, p, from, 2 + c.P->frag_score_table[k], c.hit[p].NM + c.hit[from].NM, k, dp_flag);
    else NA(c, A_DP_FLAG, p) = 0 - dp_flag;
}
// fnode_add_son, :683-698 (insertion order kept by the tail pointer)
__device__ __forceinline__ void node_add_son(Ctx& c, int fa, int son) {
    NA(c, A_IN_DE, fa) += 1;
    const int n = NA(c, A_SON_N, fa);
    if (n == 0) NA(c, A_SON_HEAD, fa) = son;
    else NA(c, A_SIB, NA(c, A_SON_TAIL, fa)) = son;
    NA(c, A_SON_TAIL, fa) = son;
    NA(c, A_SON_N, fa) = n + 1;
}

// ----------------------------------------------------- predecessor scan ------
// frag_dp_update, :701-764: lanes classify 32 predecessors per step, in the reference's scan order.
__device__ void node_update(Ctx& c, int me, int start_seed, int dp_flag) {
    const lb2_sdp_para& P = *c.P;
    const int x = c.hseed[me];
    const int rq0 = c.H - c.hoff[x], rq1 = c.H - c.hoff[start_seed];
    const int my_from = NA(c, A_FROM, me), me_NM = NA(c, A_TOL_NM, me);
    int best_from = my_from, best_score = NA(c, A_SCORE, me), best_NM = me_NM, best_flag = NA(c, A_DP_FLAG, me);
    const lb2_sdp_hit hm = load_hit(c.hit + me);
    const int sid_me = c.seed_id[x];
    for (int base = rq0; base < rq1; base += 32) {
        const int rq = base + c.lane;
        int p = -1, k = E_UNCONNECT, sc = INT_MIN, nm = INT_MAX;
        bool pass = false, ok = false, early = false;
        if (rq < rq1) {
            p = c.rflat[rq];
            if (NA(c, A_DP_FLAG, p) == dp_flag) {
                const lb2_sdp_hit hp = load_hit(c.hit + p);
                if (!(hp.nstrand == 1 && NA(c, A_SON_FLAG, p) <= E_MATCH_THD)) {     // :718-720
                    pass = true;
                    k = edge_kind_hits(P, p, me, c.seed_id[c.hseed[p]], sid_me, hp, hm);
                    if (k != E_UNCONNECT && k != E_CHR_DIF) {
                        ok = true;
                        early = hp.nstrand == -1 && k <= E_MATCH_THD;                 // :726
                        sc = NA(c, A_SCORE, p) + 1 + P.frag_score_table[k];
                        nm = NA(c, A_TOL_NM, p) + me_NM;
                    }
                }
            }
        }
        const unsigned em = __ballot_sync(kFullMask, early);
        const unsigned pm = __ballot_sync(kFullMask, pass);
        const unsigned om = __ballot_sync(kFullMask, ok);
        // candidates that come before the first "early" one still compete in order; the early one then
        // overrides whatever was best (goto UPDATE), so only it matters.
        if (em) {
            const int e = __ffs(em) - 1;
            c.pairs += __popc(pm & (0xffffffffu >> (31 - e)));
            best_from = __shfl_sync(kFullMask, p, e); best_score = __shfl_sync(kFullMask, sc, e);
            best_flag = __shfl_sync(kFullMask, k, e); best_NM = __shfl_sync(kFullMask, nm, e);
            break;
        }
        c.pairs += __popc(pm);
        if (om) {
            const int ms = __reduce_max_sync(kFullMask, sc);
            const int mn = __reduce_min_sync(kFullMask, (ok && sc == ms) ? nm : INT_MAX);
            const unsigned wm = __ballot_sync(kFullMask, ok && sc == ms && nm == mn);
            const int wl = __ffs(wm) - 1;
            const int wp = __shfl_sync(kFullMask, p, wl), wk = __shfl_sync(kFullMask, k, wl);
            if (ms > best_score || (ms == best_score && mn < best_NM)) {             // :736-748
                best_from = wp; best_score = ms; best_NM = mn; best_flag = wk;
            }
        }
    }
    if (best_from != my_from) {                                                       // :753-762
        NA(c, A_SON_FLAG, best_from) = best_flag;
        NA(c, A_FROM, me) = best_from; NA(c, A_SCORE, me) = best_score; NA(c, A_TOL_NM, me) = best_NM;
        NA(c, A_MATCH_FLAG, me) = best_flag;
        NA(c, A_NODE_N, me) = NA(c, A_NODE_N, best_from) + 1;
        node_add_son(c, best_from, me);
    }
}

// --------------------------------------------------------------- path stack --
// node_score used as a stack (node_init_score :29, node_add_score :786, node_pop src/lamsa_heap.c:5)
#define STK_NODE(c, i)  ((c).stk[i])
#define STK_SCORE(c, i) ((c).stk[(c).H + 2 + (i)])
#define STK_NM(c, i)    ((c).stk[2 * ((c).H + 2) + (i)])
__device__ void path_push(Ctx& c, int score, int NM, int node) {
    if (score < c.stk_thd) return;
    if (c.stk_n > c.H) { c.err = ERR_STACK; return; }
    STK_SCORE(c, c.stk_n) = score; STK_NM(c, c.stk_n) = NM; STK_NODE(c, c.stk_n) = node; ++c.stk_n;
    NA(c, A_DP_FLAG, node) = ST_TRACKED;
    int guard = 0;
    for (int t = NA(c, A_FROM, node); t >= 0; t = NA(c, A_FROM, t)) {
        NA(c, A_DP_FLAG, t) = ST_TRACKED;
        if (++guard > c.H) { c.err = ERR_PATH; return; }
    }
}
// detach a son from its parent (:842-847, :851-857, :894-899)
__device__ void detach(Ctx& c, int son, int max_node) {
    NA(c, A_FROM, son) = -1;
    const int ms = NA(c, A_MAX_SCORE, son) - (NA(c, A_SCORE, son) - 1);
    const int mn = NA(c, A_MAX_NM, son) - (NA(c, A_TOL_NM, son) - c.hit[son].NM);
    NA(c, A_MAX_SCORE, son) = ms; NA(c, A_MAX_NM, son) = mn;
    NA(c, A_NODE_N, max_node) -= NA(c, A_NODE_N, son) - 1;
    path_push(c, ms, mn, max_node);
}
// get_max_son, :808-829
__device__ int best_son(Ctx& c, int f) {
    int max_score = 0, max_NM = 0, max_dis = 0, flag_thd = E_INIT, best = -1;
    const int x = c.hseed[f], n = NA(c, A_SON_N, f);
    int s = NA(c, A_SON_HEAD, f);
    for (int i = 0; i < n; ++i) {
        const int mf = NA(c, A_MATCH_FLAG, s), sc = NA(c, A_MAX_SCORE, s), nm = NA(c, A_MAX_NM, s), dx = c.hseed[s] - x;
        if (mf <= flag_thd && (sc > max_score || (sc == max_score && (dx < max_dis || nm < max_NM)))) {
            best = s; max_score = sc; max_NM = nm; max_dis = dx;
            if (mf <= E_MATCH_THD) flag_thd = E_MATCH_THD;
        }
        if (i + 1 < n) s = NA(c, A_SIB, s);
    }
    return best;
}
// cut_branch, :831-870
__device__ void cut_branch(Ctx& c, int f) {
    const int keep = best_son(c, f);
    if (keep < 0) { c.err = ERR_PATH; return; }          // the reference would read an uninitialised node here
    const int n = NA(c, A_SON_N, f);
    int s = NA(c, A_SON_HEAD, f);
    for (int i = 0; i < n; ++i) {
        const int nxt = (i + 1 < n) ? NA(c, A_SIB, s) : -1;
        if (s != keep) detach(c, s, NA(c, A_MAX_NODE, s));
        s = nxt;
    }
    if (NA(c, A_SCORE, f) > NA(c, A_MAX_SCORE, keep)) {
        NA(c, A_IN_DE, keep) = -1;
        detach(c, keep, NA(c, A_MAX_NODE, keep));
        NA(c, A_SON_N, f) = 0;
        NA(c, A_MAX_NODE, f) = f; NA(c, A_MAX_SCORE, f) = NA(c, A_SCORE, f); NA(c, A_MAX_NM, f) = NA(c, A_TOL_NM, f);
    } else {
        NA(c, A_SON_N, f) = 1; NA(c, A_SON_HEAD, f) = keep; NA(c, A_SON_TAIL, f) = keep;
        NA(c, A_MAX_NODE, f) = NA(c, A_MAX_NODE, keep); NA(c, A_MAX_SCORE, f) = NA(c, A_MAX_SCORE, keep);
        NA(c, A_MAX_NM, f) = NA(c, A_MAX_NM, keep);
    }
    NA(c, A_IN_DE, f) = 0;
}
// branch_track_new, :873-920
__device__ void track_from_leaf(Ctx& c, int n) {
    int max_score, max_NM, max_node;
    NA(c, A_IN_DE, n) = -1;
    if (NA(c, A_SON_N, n) == 0) {
        max_node = n; NA(c, A_MAX_NODE, n) = n;
        max_score = NA(c, A_SCORE, n); NA(c, A_MAX_SCORE, n) = max_score;
        max_NM = NA(c, A_TOL_NM, n); NA(c, A_MAX_NM, n) = max_NM;
    } else { max_node = NA(c, A_MAX_NODE, n); max_score = NA(c, A_MAX_SCORE, n); max_NM = NA(c, A_MAX_NM, n); }
    int fa = NA(c, A_FROM, n), guard = 0;
    while (fa >= 0) {
        if (NA(c, A_SON_N, fa) != 1) {
            const int d = NA(c, A_IN_DE, fa) - 1;
            NA(c, A_IN_DE, fa) = d;
            if (d == 0) cut_branch(c, fa);
            return;
        }
        if (NA(c, A_SCORE, fa) > max_score) {
            const int s = NA(c, A_SON_HEAD, fa);
            NA(c, A_IN_DE, s) = -1;
            detach(c, s, max_node);
            NA(c, A_SON_N, fa) = 0;
            max_score = NA(c, A_SCORE, fa); max_NM = NA(c, A_TOL_NM, fa); max_node = fa;
        }
        NA(c, A_MAX_SCORE, fa) = max_score; NA(c, A_MAX_NM, fa) = max_NM; NA(c, A_MAX_NODE, fa) = max_node;
        NA(c, A_IN_DE, fa) = -1;
        fa = NA(c, A_FROM, fa);
        if (++guard > c.H || c.err) { if (!c.err) c.err = ERR_PATH; return; }
    }
    path_push(c, max_score, max_NM, max_node);
}

// ---------------------------------------------------------------- gap refill --
// frag_mini_dp_line, :1068-1150.  left/right are (seed, hit) pairs; right.x may be seed_out (no tail).
__device__ int gap_refill(Ctx& c, int2 left, int2 right, int* d_score, int* d_NM, int has_tail) {
    const int* tbl = c.P->frag_score_table;
    const int lf = flat_of(c, left);
    const int head = lf;                                              // _head is always 1 at the two call sites
    const int rf = has_tail ? flat_of(c, right) : -1;
    int old_score, old_NM;
    const int left_NM = lf < 0 ? 0 : c.hit[lf].NM;
    if (!has_tail) { old_score = 1; old_NM = left_NM; }
    else { old_score = 2 + tbl[NA(c, A_MATCH_FLAG, rf)]; old_NM = left_NM + c.hit[rf].NM; }
    const int dp_flag = ST_MULTI;
    const int p0 = c.hoff[left.x + 1], p1 = c.hoff[right.x];          // nodes of the seeds strictly between
    for (int p = p0 + c.lane; p < p1; p += 32) {                      // lane-parallel: independent per node
        const int f = NA(c, A_DP_FLAG, p);
        if (f == dp_flag || f == 0 - dp_flag) node_init(c, p, head, dp_flag);
    }
    __syncwarp();
    for (int p = c.hoff[min(left.x + 2, right.x)]; p < p1; ++p)
        if (NA(c, A_DP_FLAG, p) == dp_flag) node_update(c, p, left.x + 1, dp_flag);
    int max_score, max_NM = 0, max_n = 0, max_node = head;
    if (!has_tail) {
        max_score = old_score;
        for (int i = right.x - 1; i > left.x; --i)
            for (int p = c.hoff[i]; p < c.hoff[i + 1]; ++p) {
                if (NA(c, A_DP_FLAG, p) != dp_flag) continue;
                const int sc = NA(c, A_SCORE, p), nm = NA(c, A_TOL_NM, p);
                if (sc > max_score || (sc == max_score && nm < max_NM)) { max_score = sc; max_NM = nm; max_node = p; max_n = NA(c, A_NODE_N, p); }
            }
    } else {
        NA(c, A_FROM, rf) = head; NA(c, A_SCORE, rf) = old_score; NA(c, A_TOL_NM, rf) = old_NM; NA(c, A_NODE_N, rf) = 1;
        node_update(c, rf, left.x + 1, dp_flag);
        max_score = NA(c, A_SCORE, rf); max_NM = NA(c, A_TOL_NM, rf); max_node = NA(c, A_FROM, rf); max_n = NA(c, A_NODE_N, rf) - 1;
    }
    int k = max_n - 1;
    for (int t = max_node; seed_of(c, t) != left.x; t = NA(c, A_FROM, t)) {
        if (k < 0 || t < 0) { c.err = ERR_PATH; return 0; }
        c.mini[k--] = xy_of(c, t);
    }
    if (k >= 0) { c.err = ERR_PATH; return 0; }
    *d_score += max_score - old_score;
    *d_NM += max_NM - old_NM;
    return max_n;
}

// frag_min_extend, :1031-1066: lanes take the other seeds, each stops at its first co-linear hit
__device__ void promote_colinear(Ctx& c, int me, int aln_min, int dp_flag) {
    const int x = c.hseed[me];
    for (int i = c.lane; i < c.S; i += 32) {
        if (i == x || c.map_n[i] <= aln_min) continue;
        for (int p = c.hoff[i]; p < c.hoff[i + 1]; ++p) {
            const int k = i < x ? edge_kind(c, p, me) : edge_kind(c, me, p);
            if (k == E_MATCH || k == E_MISMATCH || k == E_LONG_MISMATCH) { NA(c, A_DP_FLAG, p) = dp_flag; break; }
        }
    }
}

// ------------------------------------------------------ skeleton clustering --
#define SK_NODE(L, sl, i) ((L) + (sl)[(i) << 1])
#define SK_LEN(sl, i)     ((sl)[((i) << 1) + 1])
#define T_LB(l, n) ((l)[(n)].x)
#define T_RB(l, n) ((l)[(n)].y)
#define T_MF(l, n) ((l)[(n) + 2].x)
#define T_MH(l, n) ((l)[(n) + 2].y)
#define T_LS(l, n) ((l)[(n) + 3].x)
#define T_BS(l, n) ((l)[(n) + 3].y)
#define T_NM(l, n) ((l)[(n) + 4].x)
#define MFI(i) T_MF(SK_NODE(L, sl, i), SK_LEN(sl, i))
#define MHI(i) T_MH(SK_NODE(L, sl, i), SK_LEN(sl, i))
#define LSI(i) T_LS(SK_NODE(L, sl, i), SK_LEN(sl, i))
#define NMI(i) T_NM(SK_NODE(L, sl, i), SK_LEN(sl, i))

// line_sort_endpos, :10-27: descending last-seed index, stable (glibc qsort = merge sort)
__device__ void sort_by_end(Ctx& c, int2* L, int* sl, int* rank, int* srank, int len) {
    int* pos = c.clu; int* id = c.clu + len;
    for (int i = 0; i < len; ++i) { id[i] = i; pos[i] = SK_NODE(L, sl, i)[SK_LEN(sl, i) - 1].x; }
    for (int i = 1; i < len; ++i) {
        const int p = pos[i], v = id[i];
        int k = i - 1;
        while (k >= 0 && pos[k] < p) { pos[k + 1] = pos[k]; id[k + 1] = id[k]; --k; }
        pos[k + 1] = p; id[k + 1] = v;
    }
    for (int i = 0; i < len; ++i) { const int v = id[i]; rank[i] = v; srank[v] = i; }
}
// line_merge, :69-112
__device__ void merge_pair(int a, int b, int2* L, int* sl, float ovlp_r) {
    int2 *na = SK_NODE(L, sl, a), *nb = SK_NODE(L, sl, b);
    const int la = SK_LEN(sl, a), lb = SK_LEN(sl, b);
    int s1, e1, hi, lhi;
    const int s2 = na[0].x, e2 = na[la - 1].x;
    int2* nhi;
    if (T_MF(nb, lb) & SK_NMERG) { hi = b; nhi = nb; lhi = lb; s1 = nb[0].x; e1 = nb[lb - 1].x; }
    else { hi = T_MH(nb, lb); nhi = SK_NODE(L, sl, hi); lhi = SK_LEN(sl, hi); s1 = T_LB(nhi, lhi); e1 = T_RB(nhi, lhi); }
    const int st = s2 > s1 ? s2 : s1, en = e2 < e1 ? e2 : e1;
    const float r1 = (float)((en - st + 1 + 0.0) / (e1 - s1 + 1 + 0.0)), r2 = (float)((en - st + 1 + 0.0) / (e2 - s2 + 1 + 0.0));
    if (r1 < ovlp_r && r2 < ovlp_r) { T_MF(na, la) = SK_NMERG; return; }
    if (T_LS(na, la) <= T_LS(nb, lb) / 2 || T_LS(na, la) <= T_BS(nb, lb) / 2) {
        T_LB(nhi, lhi) = s1; T_RB(nhi, lhi) = e1;
        T_MF(nhi, lhi) = SK_MERGH;
        T_MH(na, la) = hi;
        T_MF(na, la) = SK_MERGB | SK_DUMP;
        return;
    }
    T_LB(nhi, lhi) = s1 + s2 - st; T_RB(nhi, lhi) = e1 + e2 - en;
    T_MF(nhi, lhi) = SK_MERGH;
    T_MF(na, la) = SK_MERGB; T_MH(na, la) = hi;
    if (T_BS(nb, lb) > T_BS(na, la)) T_BS(na, la) = T_BS(nb, lb);
}

// bounded top-k heap of skeleton ids (heap_add_node :44-59, src/lamsa_heap.c:152-201, :102-150)
#define KH_ID(c, i)    ((c).kh[i])
#define KH_SCORE(c, i) ((c).kh[(c).K + 1 + (i)])
#define KH_NM(c, i)    ((c).kh[2 * ((c).K + 1) + (i)])
__device__ __forceinline__ void kh_swap(Ctx& c, int a, int b) {
    int t = KH_ID(c, a); KH_ID(c, a) = KH_ID(c, b); KH_ID(c, b) = t;
    t = KH_SCORE(c, a); KH_SCORE(c, a) = KH_SCORE(c, b); KH_SCORE(c, b) = t;
    t = KH_NM(c, a); KH_NM(c, a) = KH_NM(c, b); KH_NM(c, b) = t;
}
__device__ void kh_sift_min(Ctx& c, int i) {
    for (;;) {
        const int l = 2 * i + 1, r = 2 * i + 2;
        int m = i;
        if (l < c.kh_n && (KH_SCORE(c, l) < KH_SCORE(c, i) || (KH_SCORE(c, l) == KH_SCORE(c, i) && KH_NM(c, l) > KH_NM(c, i)))) m = l;
        if (r < c.kh_n && (KH_SCORE(c, r) < KH_SCORE(c, m) || (KH_SCORE(c, r) == KH_SCORE(c, m) && KH_NM(c, r) > KH_NM(c, m)))) m = r;
        if (m == i) return;
        kh_swap(c, i, m); i = m;
    }
}
__device__ void kh_sift_minpos(Ctx& c, int i) {
    for (;;) {
        const int l = 2 * i + 1, r = 2 * i + 2;
        int m = i;
        if (l < c.kh_n && KH_ID(c, l) < KH_ID(c, i)) m = l;
        if (r < c.kh_n && KH_ID(c, r) < KH_ID(c, m)) m = r;
        if (m == i) return;
        kh_swap(c, i, m); i = m;
    }
}
__device__ int kh_offer(Ctx& c, int id, int score, int NM) {
    if (c.kh_n < c.K) {
        KH_SCORE(c, c.kh_n) = score; KH_NM(c, c.kh_n) = NM; KH_ID(c, c.kh_n) = id; ++c.kh_n;
        if (c.kh_n == c.K) for (int i = (c.kh_n - 1) / 2; i >= 0; --i) kh_sift_min(c, i);
        return -1;
    }
    if (KH_SCORE(c, 0) < score || (KH_SCORE(c, 0) == score && KH_NM(c, 0) > NM)) {
        const int out = KH_ID(c, 0);
        KH_SCORE(c, 0) = score; KH_NM(c, 0) = NM; KH_ID(c, 0) = id;
        kh_sift_min(c, 0);
        return out;
    }
    return -2;
}
__device__ int kh_take_minpos(Ctx& c) {
    if (c.kh_n < 1) return -1;
    const int top = KH_ID(c, 0);
    --c.kh_n;
    KH_ID(c, 0) = KH_ID(c, c.kh_n); KH_SCORE(c, 0) = KH_SCORE(c, c.kh_n); KH_NM(c, 0) = KH_NM(c, c.kh_n);
    kh_sift_minpos(c, 0);
    return top;
}

// One cluster (line_filter :160-235 / line_filter1 :345-400).  Members cx/cy/cz[0..cn) in rank order;
// kept[0] = best, kept[1..] = head then bodies; returns the kept count.
__device__ int pick_in_cluster(Ctx& c, int2* L, int* sl, const int* cx, const int* cy, const int* cz, int cn,
                               int* kept, int* tri_n) {
    int b_score = 0, s_score = 0, kn = 1;
    for (int j = 0; j < cn; ++j) {
        if (cy[j] > b_score) { s_score = b_score; b_score = cy[j]; }
        else if (cy[j] > s_score) s_score = cy[j];
    }
    if (s_score >= b_score / 2) {
        c.kh_n = 0;
        for (int j = 0; j < cn; ++j) {
            if (cy[j] >= b_score / 2) {
                const int ret = kh_offer(c, cx[j], cy[j], cz[j]);
                if (ret == -2) { MFI(cx[j]) |= SK_DUMP; if (tri_n) tri_n[cx[j]] = 0; }
                else if (ret != -1) { MFI(ret) |= SK_DUMP; if (tri_n) tri_n[ret] = 0; }
            } else { MFI(cx[j]) |= SK_DUMP; if (tri_n) tri_n[cx[j]] = 0; }
        }
        for (int i = (c.kh_n - 1) / 2; i >= 0; --i) kh_sift_minpos(c, i);
        const int head = kh_take_minpos(c);
        int2* hn = SK_NODE(L, sl, head); const int hl = SK_LEN(sl, head);
        T_MF(hn, hl) = SK_MERGH;
        int min_l = hn[0].x, max_r = hn[hl - 1].x, body;
        if (T_LS(hn, hl) == b_score) kept[0] = head;
        kept[kn++] = head;
        while ((body = kh_take_minpos(c)) != -1) {
            int2* bn = SK_NODE(L, sl, body); const int bl = SK_LEN(sl, body);
            T_MF(bn, bl) = SK_MERGB; T_MH(bn, bl) = head;
            min_l = min(min_l, bn[0].x); max_r = max(max_r, bn[bl - 1].x);
            if (T_LS(bn, bl) == b_score) kept[0] = body;
            kept[kn++] = body;
        }
        T_LB(hn, hl) = min(hn[0].x, min_l);
        T_RB(hn, hl) = max(hn[hl - 1].x, max_r);
    } else {
        for (int j = 0; j < cn; ++j) {
            if (cy[j] == b_score) { MFI(cx[j]) = SK_NMERG; kept[0] = cx[j]; kept[kn++] = cx[j]; }
            else { MFI(cx[j]) |= SK_DUMP; if (tri_n) tri_n[cx[j]] = 0; }
        }
    }
    return kn;
}

// line_filter :122-319 (stage1) / line_filter1 :321-404 (stage 2).  The reference's len x len work
// arrays are flattened: clusters are contiguous in rank order, so members and kept lists are
// segments of two flat arrays.
__device__ void filter_clusters(Ctx& c, int2* L, int* sl, int* rank, int* srank, int len, bool stage1) {
    int* cx = c.clu; int* cy = cx + (len + 1); int* cz = cy + (len + 1);
    int* cstart = cz + (len + 1);                 // len + 2
    int* kept = cstart + (len + 2);               // 2 * len + 2
    int* kstart = kept + (2 * len + 2);           // len + 1
    int* kcnt = kstart + (len + 1);               // len + 1
    int m_i = -1, n_mem = 0;
    for (int _i = 0; _i < len; ++_i) {
        const int i = rank[_i], mf = MFI(i);
        if (mf & SK_DUMP) continue;
        if (mf & SK_NMERG) {
            if (!stage1) continue;
            ++m_i; cstart[m_i] = n_mem; cx[n_mem] = i; cy[n_mem] = -2; cz[n_mem] = 0; ++n_mem;
        } else if (mf & SK_MERGH) {
            ++m_i; cstart[m_i] = n_mem; cx[n_mem] = i; cy[n_mem] = LSI(i); cz[n_mem] = NMI(i); ++n_mem;
        } else {
            if (m_i < 0) { c.err = ERR_PATH; return; }
            cx[n_mem] = i; cy[n_mem] = LSI(i); cz[n_mem] = NMI(i); ++n_mem;
        }
    }
    cstart[m_i + 1] = n_mem;
    int n_kept = 0;
    for (int i = 0; i <= m_i; ++i) {
        const int a = cstart[i], cn = cstart[i + 1] - a;
        kstart[i] = n_kept;
        int* kp = kept + n_kept;
        if (cy[a] == -2) { kp[0] = cx[a]; kcnt[i] = 1; n_kept += 1; continue; }
        const int kn = pick_in_cluster(c, L, sl, cx + a, cy + a, cz + a, cn, kp, stage1 ? c.tri_n : nullptr);
        kcnt[i] = kn; n_kept += kn;
        if (!stage1) continue;
        // inter-skeleton candidates inside the triggers of the kept skeletons, :236-274
        for (int ii = 1; ii < kn; ++ii) {
            const int j = kp[ii], _j = srank[j];
            for (int k = 0; k < c.tri_n[j]; ++k) {
                int head = -1;
                const int4 tg = c.trg[c.tri_start[j] + k];       // n1 = (x,y), n2 = (z,w)
                for (int _l = _j + 1; _l < len; ++_l) {
                    const int l = rank[_l];
                    int2* nl = SK_NODE(L, sl, l); const int ll = SK_LEN(sl, l);
                    if ((T_MF(nl, ll) & 0x3) != 0) break;
                    if (!(nl[0].x > tg.x && nl[ll - 1].x < tg.z)) continue;
                    const int f2 = c.hoff[tg.z] + tg.w, f1 = c.hoff[tg.x] + tg.y;
                    const int mfk = NA(c, A_MATCH_FLAG, f2);
                    if (mfk != E_MISMATCH && mfk != E_LONG_MISMATCH) continue;
                    const lb2_sdp_hit hs = load_hit(c.hit + flat_of(c, nl[0])), he = load_hit(c.hit + flat_of(c, nl[ll - 1]));
                    const lb2_sdp_hit h1 = load_hit(c.hit + f1), h2 = load_hit(c.hit + f2);
                    const int st = hs.nstrand;
                    if (st == h1.nstrand || hs.nchr != h1.nchr || st * hs.offset < st * h2.offset || st * he.offset > st * h1.offset) continue;
                    T_MF(nl, ll) = SK_INTER;
                    if (head == -1) { T_MF(nl, ll) |= SK_NMERG; head = l; }
                    else { T_MF(nl, ll) |= SK_MERGB; T_MH(nl, ll) = head; MFI(head) = SK_INTER | SK_MERGH; }
                }
            }
        }
    }
    // a cluster spanning fewer than 3 seeds at either end, next to a real one, is dropped, :279-316
    if (stage1 && m_i > 0) {
        for (int side = 0; side < 2; ++side) {
            const int a = side == 0 ? 0 : m_i, b = side == 0 ? 1 : m_i - 1;
            const int ia = kept[kstart[a]], ib = kept[kstart[b]];
            int2* na = SK_NODE(L, sl, ia); const int la = SK_LEN(sl, ia);
            int2* nb = SK_NODE(L, sl, ib); const int lb = SK_LEN(sl, ib);
            if (na[la - 1].x - na[0].x < 2 && nb[lb - 1].x - nb[0].x >= 2) {
                for (int i = 1; i < kcnt[a]; ++i) {
                    const int v = kept[kstart[a] + i];
                    MFI(v) = SK_DUMP;
                    for (int _j = 0; _j < len; ++_j) {
                        const int j = rank[_j], mf = MFI(j);
                        if (!(mf & SK_NMERG) && !(mf & SK_MERGH) && !(mf & SK_DUMP) && MHI(j) == v) MFI(j) = SK_DUMP;
                    }
                }
            }
        }
    }
}

// line_set_bound :425-441 / line_set_bound1 :496-511 (+ line_remove :406-423); returns the kept count
__device__ int cluster_skeletons(Ctx& c, int2* L, int* sl, int* rank, int* srank, int len, bool stage1) {
    if (len <= 0) return len;
    sort_by_end(c, L, sl, rank, srank, len);
    MFI(rank[0]) = SK_NMERG;
    for (int i = 1; i < len; ++i) merge_pair(rank[i], rank[i - 1], L, sl, c.P->ovlp_rat);
    filter_clusters(c, L, sl, rank, srank, len, stage1);
    int cur = 0;
    for (int _l = 0; _l < len; ++_l) {
        const int l = rank[_l];
        if (!(MFI(l) & SK_DUMP)) rank[cur++] = l;
    }
    return cur;
}

// ------------------------------------------------------ fragments -> stream --
__device__ __forceinline__ void put(Ctx& c, int v) {
    if (c.out_n < c.out_cap) c.out[c.out_n] = v; else c.err = ERR_STREAM;
    ++c.out_n;
}
// line_filter_overlap :568-594 + frag_dp_path :1152-1250
__device__ void emit_skeletons(Ctx& c, int2* L, int* sl, int* rank, int line_n) {
    put(c, line_n);
    if (line_n == 0) return;
    if (c.P->aln_mode & 1) {
        const int seed_len = c.P->seed_len;
        for (int _i = 0; _i < line_n; ++_i) {
            const int i = rank[_i];
            int2* ni = SK_NODE(L, sl, i); const int li = SK_LEN(sl, i);
            int last = 0;
            for (int j = 1; j < li; ++j) {
                const bool is_tail = (j == li - 1);
                const int cf = flat_of(c, ni[j]), pf = flat_of(c, ni[last]);
                const lb2_sdp_hit hc = load_hit(c.hit + cf), hp = load_hit(c.hit + pf);
                const bool ovl = (int64_t)(seed_len + ((hc.nstrand == 1) ? hp.len_dif : hc.len_dif)) > hc.nstrand * (hc.offset - hp.offset)
                                 && NA(c, A_MATCH_FLAG, cf) != E_INSERT;
                if (is_tail) { if (ovl) ni[last].x = -1; }
                else if (ovl) ni[j].x = -1;
                else last = j;
            }
        }
    }
    for (int _l = 0; _l < line_n; ++_l) {
        const int l = rank[_l];
        int2* ln = SK_NODE(L, sl, l); const int ll = SK_LEN(sl, l);
        put(c, T_LS(ln, ll));
        const int frag_num_at = c.out_n; put(c, 0);
        int frag_num = 0, seed_num_at = c.out_n, seed_num = 1;
        put(c, 1);
        int2 pre = ln[ll - 1], cur;
        put(c, pre.x); put(c, pre.y);
        for (int i = ll - 1; i > 0; --i) {
            cur = pre;
            if (ln[i - 1].x < 0) continue;
            pre = ln[i - 1];
            const int mf = cur.x < 0 ? E_MATCH : NA(c, A_MATCH_FLAG, flat_of(c, cur));
            if (mf == E_INSERT || mf == E_DELETE || mf == E_MISMATCH || mf == E_LONG_MISMATCH) {
                if (seed_num_at < c.out_cap) c.out[seed_num_at] = seed_num;
                ++frag_num;
                seed_num_at = c.out_n; put(c, 1); seed_num = 1;
                put(c, pre.x); put(c, pre.y);
            } else if (mf == E_MATCH) {
                put(c, pre.x); put(c, pre.y); ++seed_num;
            } else { c.err = ERR_EDGE; return; }
        }
        if (seed_num_at < c.out_cap) c.out[seed_num_at] = seed_num;
        if (frag_num_at < c.out_cap) c.out[frag_num_at] = frag_num + 1;
    }
}

// --------------------------------------------------------------------- stage 1 --
// frag_line_BCC, :1305-1445
__device__ void stage_bcc(Ctx& c) {
    const lb2_sdp_para& P = *c.P;
    int min_n = P.first_loci_thd, min_num = 0;
    for (int i = c.lane; i < c.S; i += 32) min_num += (c.map_n[i] <= min_n);
    min_num = __reduce_add_sync(kFullMask, min_num);
    const bool all_min = (min_num == 0 || min_num * 3 < c.S);
    for (int p = c.lane; p < c.H; p += 32)
        node_set(c, p, -1, 1, c.hit[p].NM, E_MATCH, (all_min || c.map_n[c.hseed[p]] <= min_n) ? ST_MIN : ST_MULTI);
    if (all_min) min_n = P.per_aln_m;
    __syncwarp();
    if (min_n != P.per_aln_m) {
        for (int i = 0; i < c.S; ++i)
            if (c.map_n[i] <= min_n)
                for (int p = c.hoff[i]; p < c.hoff[i + 1]; ++p) promote_colinear(c, p, min_n, ST_MIN);
        __syncwarp();
    }
    for (int p = c.S > 1 ? c.hoff[1] : c.H; p < c.H; ++p)
        if (NA(c, A_DP_FLAG, p) == ST_MIN) node_update(c, p, 0, ST_MIN);

    c.stk_n = 0; c.stk_thd = 2;
    for (int i = c.S - 1; i >= 0; --i)
        for (int p = c.hoff[i]; p < c.hoff[i + 1]; ++p)
            if (NA(c, A_DP_FLAG, p) == ST_MIN && NA(c, A_IN_DE, p) == 0) { track_from_leaf(c, p); if (c.err) return; }

    int l_i = 0, next_start = 0, n_trg = 0;
    int2* L = c.line; int* sl = c.lsl;
    for (;;) {
        if (c.stk_n < 1) break;                                   // node_pop: a stack, src/lamsa_heap.c:5-13
        --c.stk_n;
        int line_score = STK_SCORE(c, c.stk_n), line_NM = STK_NM(c, c.stk_n);
        const int2 max_node = xy_of(c, STK_NODE(c, c.stk_n));
        int node_i = 0, mini_len;
        c.tri_start[l_i] = n_trg; c.tri_n[l_i] = 0; sl[l_i << 1] = next_start;
        int2* ln = L + next_start;
        int2 last_n, right, left;
#define ADD_TRIGGER(a, b) do { c.trg[n_trg] = make_int4((a).x, (a).y, (b).x, (b).y); ++n_trg; ++c.tri_n[l_i]; } while (0)
        if (max_node.x < c.S - 1) {                               // refill to the right of the path end
            mini_len = gap_refill(c, max_node, make_int2(c.S, 0), &line_score, &line_NM, 0);
            if (c.err) return;
            for (int k = mini_len - 1; k >= 0; --k) { ln[node_i++] = c.mini[k]; NA(c, A_DP_FLAG, flat_of(c, c.mini[k])) = ST_TRACKED; }
            ln[node_i] = max_node;
            last_n = ln[0];
            for (int k = mini_len - 1; k >= 0; --k) {
                if (last_n.x - ln[node_i - k].x > 2) ADD_TRIGGER(ln[node_i - k], last_n);
                last_n = ln[node_i - k];
            }
        }
        right = max_node;
        while (right.x != -1) {
            ln[node_i++] = right;
            left = xy_of(c, NA(c, A_FROM, flat_of(c, right)));
            if (left.x < right.x - 1) {                           // refill every gap of the path
                mini_len = gap_refill(c, left, right, &line_score, &line_NM, 1);
                if (c.err) return;
                for (int k = mini_len - 1; k >= 0; --k) { ln[node_i++] = c.mini[k]; NA(c, A_DP_FLAG, flat_of(c, c.mini[k])) = ST_TRACKED; }
                ln[node_i] = left;
                last_n = right;
                for (int k = mini_len; k >= 0; --k) {
                    if (last_n.x - ln[node_i - k].x > 2) {
                        if (ln[node_i - k].x == -1) continue;
                        ADD_TRIGGER(ln[node_i - k], last_n);
                    }
                    last_n = ln[node_i - k];
                }
            }
            right = left;
            if (node_i > c.H + 1) { c.err = ERR_PATH; return; }
        }
#undef ADD_TRIGGER
        for (int k = 0; k < node_i / 2; ++k) { const int2 t = ln[k]; ln[k] = ln[node_i - k - 1]; ln[node_i - k - 1] = t; }
        sl[(l_i << 1) + 1] = node_i;
        T_LS(ln, node_i) = line_score; T_BS(ln, node_i) = line_score; T_NM(ln, node_i) = line_NM;
        T_MF(ln, node_i) = 0; T_MH(ln, node_i) = 0; T_LB(ln, node_i) = 0; T_RB(ln, node_i) = 0;
        ++l_i; next_start += node_i + SK_EXTRA;
    }
    const int kept = cluster_skeletons(c, L, sl, c.rank, c.srank, l_i, true);
    if (c.err) return;
    emit_skeletons(c, L, sl, c.rank, kept);
}

// --------------------------------------------------------------------- stage 2 --
// frag_mini_dp_multi_line, :923-1017
__device__ int region_dp(Ctx& c, int left_b, int right_b, const DRegion& rg, const DPoint* pts, int2* L, int* sl) {
    if (left_b + 1 >= right_b) return 0;
    const lb2_sdp_para& P = *c.P;
    const int start = left_b + 1, end = right_b - 1, dp_flag = ST_WHOLE;
    const int p0 = c.hoff[start], p1 = c.hoff[end + 1];
    for (int p = p0 + c.lane; p < p1; p += 32)
        if (NA(c, A_DP_FLAG, p) != ST_TRACKED) node_set(c, p, -1, 1, c.hit[p].NM, E_MATCH, dp_flag);
    __syncwarp();
    for (int p = c.hoff[min(start + 1, end + 1)]; p < p1; ++p)
        if (NA(c, A_DP_FLAG, p) == dp_flag) node_update(c, p, start, dp_flag);
    c.stk_n = 0; c.stk_thd = 0;
    for (int i = end; i >= start; --i)
        for (int p = c.hoff[i]; p < c.hoff[i + 1]; ++p)
            if (NA(c, A_DP_FLAG, p) == dp_flag && NA(c, A_IN_DE, p) == 0) { track_from_leaf(c, p); if (c.err) return 0; }
    int l_i = 0, next_start = 0;
    while (c.stk_n >= 1) {
        --c.stk_n;
        int score = STK_SCORE(c, c.stk_n);
        const int NM = STK_NM(c, c.stk_n), rf = STK_NODE(c, c.stk_n);
        int node_i = NA(c, A_NODE_N, rf) - 1;
        if (node_i < 0 || next_start + node_i + 1 + SK_EXTRA > 6 * c.H + 16 + c.S) { c.err = ERR_PATH; return 0; }
        sl[l_i << 1] = next_start; sl[(l_i << 1) + 1] = node_i + 1;
        next_start += node_i + 1 + SK_EXTRA;
        const lb2_sdp_hit hr = load_hit(c.hit + rf);
        const int rx = c.hseed[rf];
        bool near = false;                                        // :979-996
        for (int i = 0; i < rg.bn + rg.en && !near; ++i) {
            const DPoint pt = pts[rg.pt_first + i];
            if (hr.nchr == pt.chr && llabs((hr.offset - pt.pos) - (long long)((rx - left_b) * P.seed_step)) < P.SV_len_thd) near = true;
        }
        if (near) { if (score > 1) score += score / 2; else score++; }
        int2* node = SK_NODE(L, sl, l_i); const int len = SK_LEN(sl, l_i);
        T_LS(node, len) = score; T_BS(node, len) = score; T_NM(node, len) = NM;
        T_MF(node, len) = 0; T_MH(node, len) = 0; T_LB(node, len) = 0; T_RB(node, len) = 0;
        for (int t = rf; t >= 0; t = NA(c, A_FROM, t)) {
            if (node_i < 0) { c.err = ERR_PATH; return 0; }
            node[node_i--] = xy_of(c, t);
        }
        if (node_i >= 0) { c.err = ERR_PATH; return 0; }
        ++l_i;
    }
    return l_i;
}

// frag_line_remain, :1252-1302 (the regions come from the host: get_remain_reg, src/lamsa_aln.c:548-572)
__device__ void stage_remain(Ctx& c, const DRegion* regions, int n_region, const DPoint* pts) {
    const lb2_sdp_para& P = *c.P;
    int l_n = 0, next_start = 0;
    for (int i = 0; i < n_region; ++i) {
        const DRegion rg = regions[i];
        const int left_id = (rg.beg + P.seed_inv - 1) / P.seed_step + 1;
        int right_id = (rg.end - 1) / P.seed_step + 1;
        if (right_id > c.seed_all) right_id -= 1;
        int left = -2, right = -2;
        for (int j = 0; j < c.S; ++j) if (c.seed_id[j] >= left_id) { left = j - 1; break; }
        if (left == -2) continue;
        for (int j = c.S - 1; j >= 0; --j) if (c.seed_id[j] <= right_id) { right = j + 1; break; }
        if (right == -2) continue;
        const int l0 = region_dp(c, left, right, rg, pts, c.tline, c.tlsl);
        if (c.err) return;
        const int l = cluster_skeletons(c, c.tline, c.tlsl, c.trank, c.srank, l0, false);
        if (c.err) return;
        for (int _j = 0; _j < l; ++_j) {
            const int j = c.trank[_j], len = c.tlsl[(j << 1) + 1];
            if (next_start + len + SK_EXTRA > 6 * c.H + 16 + c.S) { c.err = ERR_PATH; return; }
            c.lsl[(l_n + _j) << 1] = next_start; c.lsl[((l_n + _j) << 1) + 1] = len;
            for (int k = 0; k < len + SK_EXTRA; ++k) c.line[next_start + k] = c.tline[c.tlsl[j << 1] + k];
            next_start += len + SK_EXTRA;
            c.rank[l_n + _j] = l_n + _j;
        }
        l_n += l;
    }
    emit_skeletons(c, c.line, c.lsl, c.rank, l_n);
}

// ----------------------------------------------------------------------- kernel --
// Persistent warps pull reads (largest first) from an atomic counter.
#ifndef LB2_SDP_MIN_BLOCKS
#define LB2_SDP_MIN_BLOCKS 8      // 64 registers, 32 warps per SM: 25.4 vs 20.4 Gpairs/s at 4 (124 registers) on the 20k-read pacbio set
#endif
__global__ void __launch_bounds__(128, LB2_SDP_MIN_BLOCKS)
sdp_kernel(const int stage, const __grid_constant__ lb2_sdp_para P, const int n_reads, const DRead* __restrict__ reads,
           const int32_t* __restrict__ order, const int32_t* __restrict__ seed_id, const int32_t* __restrict__ map_n,
           const int32_t* __restrict__ hoff, const lb2_sdp_hit* __restrict__ hits, const int32_t* __restrict__ hseed,
           const int32_t* __restrict__ rflat, const DRegion* __restrict__ regions, const DPoint* __restrict__ pts,
           int* scratch, DOut* outs, unsigned int* counter) {
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(counter, 1u);
        slot = __shfl_sync(kFullMask, slot, 0);
        if (slot >= (unsigned)n_reads) return;
        const int r = order[slot];
        const DRead rd = reads[r];
        Ctx c;
        c.P = &P; c.lane = lane; c.S = rd.seed_out; c.H = rd.n_hits; c.seed_all = rd.seed_all; c.read_len = rd.read_len;
        c.K = P.ske_max;
        c.seed_id = seed_id + rd.seed_base; c.map_n = map_n + rd.seed_base; c.hoff = hoff + rd.hoff_base;
        c.hit = hits + rd.hit_base; c.hseed = hseed + rd.hit_base; c.rflat = rflat + rd.hit_base;
        const Layout lay = make_layout(rd.n_hits, rd.seed_out, P.ske_max);
        int* base = scratch + rd.scratch;
        c.nd = base + lay.node;
        c.line = reinterpret_cast<int2*>(base + lay.line); c.tline = reinterpret_cast<int2*>(base + lay.tline);
        c.mini = reinterpret_cast<int2*>(base + lay.mini);
        c.lsl = base + lay.lsl; c.tlsl = base + lay.tlsl; c.rank = base + lay.rank; c.trank = base + lay.trank; c.srank = base + lay.srank;
        c.stk = base + lay.stack; c.stk_n = 0; c.stk_thd = 0;
        c.kh = base + lay.kheap; c.kh_n = 0;
        c.clu = base + lay.clu; c.trg = reinterpret_cast<int4*>(base + lay.trg);
        c.tri_start = base + lay.tri; c.tri_n = c.tri_start + (rd.n_hits + 1);
        c.out = base + lay.out; c.out_n = 0; c.out_cap = lay.out_cap;
        c.pairs = 0; c.err = ERR_NONE;
        if (stage == 1) stage_bcc(c);
        else stage_remain(c, regions + rd.region_first, rd.n_region, pts);
        __syncwarp();
        if (lane == 0) { DOut o; o.n_words = c.err ? 0 : c.out_n; o.err = c.err; o.pairs = c.pairs; outs[r] = o; }
    }
}

// ---- ordered compaction of the per-read streams into one dense pool -------------
// exclusive scan of the stream lengths (one block; n_reads is at most a few million)
__global__ void __launch_bounds__(1024) sdp_scan_kernel(const DOut* __restrict__ outs, int n, long long* off) {
    __shared__ long long part[1024];
    const int t = threadIdx.x;
    const int per = (n + 1023) / 1024;
    const int b = min(n, t * per), e = min(n, b + per);
    long long s = 0;
    for (int i = b; i < e; ++i) s += outs[i].n_words;
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const long long v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = part[t] - s;
    for (int i = b; i < e; ++i) { off[i] = run; run += outs[i].n_words; }
    if (t == 1023) off[n] = part[1023];
}
__global__ void sdp_gather_kernel(const DRead* __restrict__ reads, const DOut* __restrict__ outs, int n, int K,
                                  const int* __restrict__ scratch, const long long* __restrict__ off, int* dense) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const DRead rd = reads[w];
    const Layout lay = make_layout(rd.n_hits, rd.seed_out, K);
    const int* src = scratch + rd.scratch + lay.out;
    int* dst = dense + off[w];
    const int nw = outs[w].n_words;
    for (int i = lane; i < nw; i += 32) dst[i] = src[i];
}

// The only node state stage 2 needs from stage 1 is "this hit is on a stage-1 skeleton"
// (dp_flag == TRACKED_FLAG, tested at src/lamsa_dp_con.c:948).  get: flags[h] = tracked; set: dp_flag of
// EVERY hit := tracked ? TRACKED : 0, which lets stage 2 run on a batch object that never ran stage 1.
__global__ void sdp_tracked_kernel(const DRead* __restrict__ reads, int n, int K, int* scratch, uint8_t* flags, int set) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const DRead rd = reads[w];
    const Layout lay = make_layout(rd.n_hits, rd.seed_out, K);
    int* dp = scratch + rd.scratch + lay.node + (int64_t)A_DP_FLAG * rd.n_hits;
    uint8_t* f = flags + rd.hit_base;
    for (int p = lane; p < rd.n_hits; p += 32) {
        if (set) dp[p] = f[p] ? ST_TRACKED : 0;
        else f[p] = dp[p] == ST_TRACKED;
    }
}

}  // namespace sdp
}  // namespace lb2
