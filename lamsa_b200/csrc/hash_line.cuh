// hash_line.cuh -- the seed-and-chain half of the reference's local split mapping on the GPU, one warp per request:
//   k-mer index of the reference window      init_hash              src/split_mapping.c:181-208
//   look-up of the read's k-mers             hash_split_map         :654-675  (at most 50 hits per k-mer, :669)
//   chaining of the hits into one line       hash_main_line         :492-602
//     relation of two hits                   hash_main_dis          :218-261
//     node initialisation                    hash_dp_init / hash_mini_dp_init   :264-310 / :399-441
//     single-hit diagonals                   hash_min_extend        :312-338
//     best predecessor                       hash_dp_update         :341-397
//     refill between two line nodes          mini_hash_main_line    :444-488
// What comes back per request is the line: read position, diagonal and relation of every node -- exactly what the
// stitching of :688-821 (host control flow around DP calls, hash_dropin.cu) reads.
//
// Formulation (the same as oracle/hash_oracle.c, which is pinned against the reference): nodes are numbered head = 0,
// the hits seed-major 1..N, tail = N+1; `from` is a node number.  The index is a hash table of the READ's distinct
// k-mers (open addressing); the window's positions are counted into it, then written, then each list (<= 50) sorted:
// ascending positions, the order init_hash_core produces.  The quadratic part -- the predecessor scan of every node --
// runs across the lanes over a scan-order array (seeds descending, hits ascending: the reference's loop order); the
// winner is the strictly best score, first in scan order (__reduce_max_sync over (score, 31 - lane) keys, chunks in
// order).  Node updates have side effects on their predecessors (:374-375), so nodes stay sequential.  Scalar control
// is executed by all lanes; every store is done by lane 0 and followed by __syncwarp().
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lb2 {

struct HashReq {
    uint32_t ref_off, read_off;        // byte offsets into the batch's sequence pool
    int32_t ref_len, read_len, ref_offset;
    int32_t hash_len, hash_step, split_len, head_on, tail_on;
    int32_t S, R, logT;                // read k-mers, window k-mers, log2 of the table size
    uint32_t fixed_off, out_off;       // int offsets: per-request fixed scratch / output (3 ints per line node, capacity S)
};
struct HashRes { int32_t m_len, status, n_nodes, pad; };      // status 0 ok, 1 node pool exhausted (retry with a larger one)

namespace hl {
enum { F_MATCH = 0, F_MISMATCH = 2, F_MATCH_THD = 2, F_LONG_MISMATCH = 3, F_INSERT = 4, F_DELETE = 5, F_UNCONNECT = 8, F_UNMATCH = 9 };
enum { FLAG_MIN = 1, FLAG_MULTI = 2, FLAG_UNLIMITED = 3, SV_PEN = 2, MAX_HITS = 50 };
constexpr unsigned kAll = 0xffffffffu;

struct Par { int L, st, split_len, ref_len, read_len, ref_offset; };

__device__ __forceinline__ int relation(const Par& P, int a_i, int a_off, int b_i, int b_off) {
    const int d = a_i > b_i ? a_off - b_off : b_off - a_off;
    const int gap = abs(b_i - a_i);
    if (d == 0) return gap < P.L + 2 * P.st ? F_MATCH : gap < P.L + 6 * P.st ? F_MISMATCH : F_LONG_MISMATCH;
    if (d > 0) return F_DELETE;
    if (d >= -(gap - P.L)) return F_INSERT;
    if (d > -(P.split_len / 2)) return F_UNCONNECT;
    const bool fwd = b_i > a_i;
    const int lo_i = fwd ? a_i : b_i, lo_off = fwd ? a_off : b_off, hi_i = fwd ? b_i : a_i, hi_off = fwd ? b_off : a_off;
    bool ok;
    if (P.ref_offset > 0) ok = P.read_len - P.ref_len + hi_off >= -(lo_i + P.L - 1) && P.read_len - lo_off >= hi_i;
    else ok = hi_off >= -(lo_i - 1) && P.ref_len - lo_off >= hi_i;
    return ok ? F_INSERT : F_UNCONNECT;
}
__device__ __forceinline__ int edge_pen(int rel) { return rel <= F_MATCH_THD ? 0 : SV_PEN; }

// node state, structure of arrays over node numbers
struct Nodes { int *x, *off, *from, *score, *cnt, *mflag, *dflag; int step; };
__device__ __forceinline__ int read_i_of(const Nodes& nd, int id, int head_ri, int tail_ri, int tail) {
    return id == 0 ? head_ri : id == tail ? tail_ri : (nd.x[id] - 1) * nd.step;
}

__device__ __forceinline__ uint32_t kmer_code(const uint8_t* p, int L) {
    uint32_t c = 0;
    for (int i = 0; i < L; ++i) { uint32_t b = p[i]; b = b >= 4u ? 2u : b; c = c << 2 | b; }      // N hashes as G (src/bntseq.c:78)
    return c;
}
__device__ __forceinline__ uint32_t table_home(uint32_t code, int logT) { return (code * 2654435761u) >> (32 - logT); }
}  // namespace hl

// one warp per request
__global__ void __launch_bounds__(128)
hash_line_kernel(const HashReq* __restrict__ reqs, int n, const uint8_t* __restrict__ pool, int32_t* __restrict__ fixed,
                 int32_t* __restrict__ node_pool, unsigned long long* __restrict__ pool_cursor, unsigned long long pool_cap,
                 int32_t* __restrict__ out, HashRes* __restrict__ results)
{
    using namespace hl;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const HashReq q = reqs[r];
    const uint8_t* __restrict__ ref = pool + q.ref_off;
    const uint8_t* __restrict__ read = pool + q.read_off;
    const int S = q.S, R = q.R, T = 1 << q.logT, TM = T - 1;
    const Par P{q.hash_len, q.hash_step, q.split_len, q.ref_len, q.read_len, q.ref_offset};
    // fixed scratch: table key / count / base [T each], slot of every seed [S+2], start / len [S+2 each], hits [R],
    // bitmap of single-hit diagonals
    int32_t* fx = fixed + q.fixed_off;
    uint32_t* tkey = reinterpret_cast<uint32_t*>(fx); fx += T;
    int32_t* tcnt = fx; fx += T;
    int32_t* tbase = fx; fx += T;
    int32_t* tcur = fx; fx += T;
    int32_t* sslot = fx; fx += S + 2;
    int32_t* start = fx; fx += S + 2;
    int32_t* len = fx; fx += S + 2;
    int32_t* hits = fx; fx += R + 1;
    uint32_t* bitmap = reinterpret_cast<uint32_t*>(fx);
    const int bm_words = (q.ref_len + q.read_len + 2 + 31) / 32 + 1;

    // ---- A. table of the read's distinct k-mers
    for (int k = lane; k < T; k += 32) { tkey[k] = 0u; tcnt[k] = 0; tcur[k] = 0; }
    __syncwarp();
    for (int s = 1 + lane; s <= S; s += 32) {
        const uint32_t key = kmer_code(read + (size_t)(s - 1) * q.hash_step, q.hash_len) + 1u;
        uint32_t h = table_home(key, q.logT);
        for (;;) {
            const uint32_t old = atomicCAS(&tkey[h], 0u, key);
            if (old == 0u || old == key) break;
            h = (h + 1) & TM;
        }
        sslot[s] = (int)h;
    }
    __syncwarp();
    // ---- B. count the window's positions per k-mer
    for (int p = lane; p < R; p += 32) {
        const uint32_t key = kmer_code(ref + p, q.hash_len) + 1u;
        uint32_t h = table_home(key, q.logT);
        for (;;) {
            const uint32_t k = tkey[h];
            if (k == key) { atomicAdd(&tcnt[h], 1); break; }
            if (k == 0u) break;
            h = (h + 1) & TM;
        }
    }
    __syncwarp();
    // ---- C. where each list (1..50 positions) starts
    {
        int run = 0;
        for (int k0 = 0; k0 < T; k0 += 32) {
            const int k = k0 + lane;
            const int c = tcnt[k];
            const int mine = (c > 0 && c <= MAX_HITS) ? c : 0;
            int inc = mine;
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kAll, inc, d); if (lane >= d) inc += v; }
            tbase[k] = run + inc - mine;
            run += __shfl_sync(kAll, inc, 31);
        }
    }
    __syncwarp();
    // ---- D. write the positions, E. sort every list
    for (int p = lane; p < R; p += 32) {
        const uint32_t key = kmer_code(ref + p, q.hash_len) + 1u;
        uint32_t h = table_home(key, q.logT);
        for (;;) {
            const uint32_t k = tkey[h];
            if (k == key) {
                const int c = tcnt[h];
                if (c <= MAX_HITS) hits[tbase[h] + atomicAdd(&tcur[h], 1)] = p;
                break;
            }
            if (k == 0u) break;
            h = (h + 1) & TM;
        }
    }
    __syncwarp();
    for (int k = lane; k < T; k += 32) {
        const int c = tcnt[k];
        if (c > 1 && c <= MAX_HITS) {
            int32_t* a = hits + tbase[k];
            for (int i = 1; i < c; ++i) { const int v = a[i]; int j = i - 1; while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; } a[j + 1] = v; }
        }
    }
    __syncwarp();
    // ---- F. hits per seed, node numbers
    int N = 0, have_single = 0;
    {
        int run = 1;
        for (int s0 = 1; s0 <= S; s0 += 32) {
            const int s = s0 + lane;
            int c = 0;
            if (s <= S) { c = tcnt[sslot[s]]; if (c > MAX_HITS) c = 0; }
            int inc = c;
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(kAll, inc, d); if (lane >= d) inc += v; }
            if (s <= S) { len[s] = c; start[s] = run + inc - c; }
            run += __shfl_sync(kAll, inc, 31);
            have_single |= __any_sync(kAll, c == 1);
        }
        N = run - 1;
        if (lane == 0) { start[0] = 0; len[0] = 1; start[S + 1] = N + 1; len[S + 1] = 1; }
    }
    const int tail = N + 1;
    // ---- G. node arrays from the batch's pool
    unsigned long long base = 0;
    const unsigned long long need = 10ull * (unsigned long long)(N + 2);
    if (lane == 0) base = atomicAdd(pool_cursor, need);
    base = __shfl_sync(kAll, base, 0);
    if (base + need > pool_cap) {
        if (lane == 0) results[r] = HashRes{0, 1, N, 0};
        return;
    }
    Nodes nd;
    {
        int32_t* a = node_pool + base;
        nd.x = a; nd.off = a + (N + 2); nd.from = a + 2 * (N + 2); nd.score = a + 3 * (N + 2); nd.cnt = a + 4 * (N + 2);
        nd.mflag = a + 5 * (N + 2); nd.dflag = a + 6 * (N + 2); nd.step = q.hash_step;
    }
    int32_t* rorder = node_pool + base + 7 * (N + 2);
    int32_t* line = node_pool + base + 8 * (N + 2);
    int32_t* mini = node_pool + base + 9 * (N + 2);
    const int head_ri = q.head_on ? -q.hash_len : -1, tail_ri = q.tail_on ? q.read_len : -1;
    __syncwarp();

    // (re)initialise node `id` against node `hd` (hash_dp_init / hash_mini_dp_init); each lane its own node
    auto init_node = [&](int id, int hd, int hd_ri, int hd_off, int hd_dflag, int flag) {
        if (hd_dflag == FLAG_UNLIMITED) { nd.from[id] = hd; nd.score[id] = 1; nd.cnt[id] = 1; nd.mflag[id] = F_MATCH; nd.dflag[id] = flag; return; }
        const int rel = relation(P, hd_ri, hd_off, (nd.x[id] - 1) * nd.step, nd.off[id]);
        if (rel == F_UNCONNECT) { nd.from[id] = -1; nd.score[id] = 0; nd.cnt[id] = 0; nd.mflag[id] = rel; nd.dflag[id] = -flag; }
        else { nd.from[id] = hd; nd.score[id] = 2 - edge_pen(rel); nd.cnt[id] = 1; nd.mflag[id] = rel; nd.dflag[id] = flag; }
    };
    // best predecessor of node `id` among the nodes of seeds [first_x, x-1] carrying `flag` (hash_dp_update); warp-wide
    auto update_node = [&](int id, int first_x, int flag) {
        const int vx = nd.x[id], v_ri = read_i_of(nd, id, head_ri, tail_ri, tail), v_off = nd.off[id];
        const int v_from = nd.from[id], v_dflag = nd.dflag[id];
        int best = nd.score[id], from = v_from, best_rel = 0;
        const int q0 = N + 1 - start[vx], q1 = N + 1 - start[first_x];
        const bool unlimited = v_dflag == FLAG_UNLIMITED;
        for (int c0 = q0; c0 < q1; c0 += 32) {
            const int qq = c0 + lane;
            unsigned key = 0u; int rel = 0, cand = -1;
            if (qq < q1) {
                cand = rorder[qq];
                if (nd.dflag[cand] == flag) {
                    if (unlimited) key = ((unsigned)(nd.score[cand] + 4) << 5) | (unsigned)(31 - lane);
                    else {
                        rel = relation(P, (nd.x[cand] - 1) * nd.step, nd.off[cand], v_ri, v_off);
                        if (rel != F_UNCONNECT) key = ((unsigned)(nd.score[cand] + 1 - edge_pen(rel) + 4) << 5) | (unsigned)(31 - lane);
                    }
                }
            }
            const unsigned top = __reduce_max_sync(kAll, key);
            if (top) {
                const int sc = (int)(top >> 5) - 4;
                if (sc > best) {
                    const int wl = 31 - (int)(top & 31u);
                    best = sc; from = __shfl_sync(kAll, cand, wl); best_rel = __shfl_sync(kAll, rel, wl);
                }
            }
        }
        if (!unlimited) {
            if (from != v_from) {
                const int add = nd.cnt[from];
                __syncwarp();
                if (lane == 0) {
                    nd.score[id] = best; nd.from[id] = from; nd.mflag[id] = best_rel;
                    if (best_rel == F_MATCH) nd.dflag[from] = -flag;
                    nd.cnt[id] = nd.cnt[id] + add;
                }
            }
        } else if (best > 0) {
            const int c = nd.cnt[from];
            __syncwarp();
            if (lane == 0) { nd.from[id] = from; nd.off[id] = -1; nd.score[id] = best; nd.cnt[id] = c; nd.mflag[id] = F_MATCH; nd.dflag[id] = FLAG_UNLIMITED; }
        }
        __syncwarp();
    };

    // ---- H. head, tail, hits
    if (lane == 0) {
        nd.x[0] = 0; nd.off[0] = q.head_on ? 0 : -1; nd.from[0] = -1; nd.score[0] = 0; nd.cnt[0] = 0; nd.mflag[0] = F_MATCH;
        nd.dflag[0] = q.head_on ? FLAG_MIN : FLAG_UNLIMITED;
        nd.x[tail] = S + 1; nd.off[tail] = q.tail_on ? q.ref_len - q.read_len : -1; nd.from[tail] = 0; nd.score[tail] = 0; nd.cnt[tail] = 0;
        nd.mflag[tail] = q.tail_on ? (q.head_on ? F_UNMATCH : F_MATCH) : F_MATCH;
        nd.dflag[tail] = q.tail_on ? FLAG_MIN : FLAG_UNLIMITED;
    }
    __syncwarp();
    {
        const int h_off = q.head_on ? 0 : -1, h_dflag = q.head_on ? FLAG_MIN : FLAG_UNLIMITED;
        for (int s = 1 + lane; s <= S; s += 32) {
            const int c = len[s], b = tbase[sslot[s]], st0 = start[s];
            for (int i = 0; i < c; ++i) {
                const int id = st0 + i;
                nd.x[id] = s; nd.off[id] = hits[b + i] - (s - 1) * q.hash_step;
                init_node(id, 0, head_ri, h_off, h_dflag, c == 1 ? FLAG_MIN : FLAG_MULTI);
                rorder[N - (st0 - 1) - c + i] = id;
            }
        }
    }
    __syncwarp();

    int n_line = 0;
    if (have_single) {
        // ---- I. multi-hit nodes on the diagonal of a single-hit node (head and tail included) join the single-hit pass
        for (int k = lane; k < bm_words; k += 32) bitmap[k] = 0u;
        __syncwarp();
        const int bias = q.read_len + 1;
        for (int s = lane; s <= S + 1; s += 32)
            if (len[s] == 1) { const int o = nd.off[start[s]] + bias; atomicOr(&bitmap[o >> 5], 1u << (o & 31)); }
        __syncwarp();
        for (int s = 1 + lane; s <= S; s += 32) {
            const int c = len[s];
            if (c <= 1) continue;
            for (int i = 0; i < c; ++i) {
                const int id = start[s] + i;
                if (nd.dflag[id] < 0) continue;
                const int o = nd.off[id] + bias;
                if ((bitmap[o >> 5] >> (o & 31)) & 1u) nd.dflag[id] = FLAG_MIN;
            }
        }
        __syncwarp();
        for (int id = start[2 <= S ? 2 : S + 1]; id <= N; ++id)
            if (nd.dflag[id] == FLAG_MIN) update_node(id, 1, FLAG_MIN);
        update_node(tail, 1, FLAG_MIN);
        // ---- walk back from the tail; a gap behind a non-match edge is refilled with multi-hit nodes
        int right = tail, left = nd.from[tail];
        for (;;) {
            if (nd.mflag[right] != F_MATCH && nd.x[left] < nd.x[right] - 1) {
                const int lx = nd.x[left], rx = nd.x[right];
                const int l_ri = read_i_of(nd, left, head_ri, tail_ri, tail), l_off = nd.off[left], l_dflag = nd.dflag[left];
                __syncwarp();
                for (int id = start[lx + 1] + lane; id < start[rx]; id += 32) init_node(id, left, l_ri, l_off, l_dflag, FLAG_MULTI);
                if (lane == 0) { nd.from[right] = left; nd.score[right] = 0; nd.cnt[right] = 0; nd.dflag[right] = FLAG_MULTI; }
                __syncwarp();
                for (int id = start[lx + 2 <= rx ? lx + 2 : rx]; id < start[rx]; ++id)
                    if (nd.dflag[id] == FLAG_MULTI) update_node(id, lx + 1, FLAG_MULTI);
                update_node(right, lx + 1, FLAG_MULTI);
                const int m = nd.cnt[right];
                int k = m - 1;
                for (int id = nd.from[right]; nd.x[id] != lx; id = nd.from[id]) { if (k >= 0 && lane == 0) mini[k] = id; --k; }
                __syncwarp();
                for (int i = m - 1 - lane; i >= 0; i -= 32) line[n_line + (m - 1 - i)] = mini[i];
                n_line += m;
                __syncwarp();
            }
            if (nd.x[left] == 0) break;
            if (lane == 0) line[n_line] = left;
            ++n_line;
            right = left; left = nd.from[right];
        }
        __syncwarp();
        // the line was collected back to front
        for (int i = lane; i < n_line; i += 32) {
            const int id = line[n_line - 1 - i];
            out[q.out_off + 3 * i] = (nd.x[id] - 1) * nd.step; out[q.out_off + 3 * i + 1] = nd.off[id]; out[q.out_off + 3 * i + 2] = nd.mflag[id];
        }
    } else {
        for (int id = start[2 <= S ? 2 : S + 1]; id <= N; ++id)
            if (nd.dflag[id] == FLAG_MULTI) update_node(id, 1, FLAG_MULTI);
        update_node(tail, 1, FLAG_MULTI);
        n_line = nd.cnt[tail];
        int k = n_line - 1;
        for (int id = nd.from[tail]; nd.x[id] != 0; id = nd.from[id]) {
            if (k >= 0 && lane == 0) { out[q.out_off + 3 * k] = (nd.x[id] - 1) * nd.step; out[q.out_off + 3 * k + 1] = nd.off[id]; out[q.out_off + 3 * k + 2] = nd.mflag[id]; }
            --k;
        }
    }
    if (lane == 0) results[r] = HashRes{n_line, 0, N, 0};
}

}  // namespace lb2
