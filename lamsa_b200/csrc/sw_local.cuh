// sw_local.cuh -- the reference's local Smith-Waterman (ksw_u8 src/ksw.c:116-235, ksw_i16 :237-335) on the GPU, one
// warp per (query, target) pair.  The striped SSE2 code is reproduced as the recurrence it evaluates (derivation and
// the pinning against the reference: oracle/sw_oracle.c):
//   T(i,j)   = max(0, H(i-1,j-1) + s(i,j), E(i,j))             over slen*lanes columns: the padding columns behind the
//                                                               query score 0 and take part in the row maximum
//   F        = max-plus prefix scan of T over the row (what main loop + lazy-F loop produce together): Ffull;
//              the same scan restarted at every multiple of slen (what the main loop alone sees): Fseg
//   H(i,j)   = max(T, Ffull),   E(i+1,j) = max(0, E - e_del, max(T, Fseg) - o_del - e_del)
// Both scans run across the lanes (inclusive prefix maximum by shuffles; the segmented one carries the segment number
// above the value, so that a later segment's entries always win and a foreign segment is recognised).
// Row maxima, the end point (first column of the best row's maximum), the merged list of rows reaching the 2nd-best
// threshold (:186-194), the early exits (KSW_XSTOP, byte overflow -> 255) follow the reference statement by statement.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lb2 {

struct SwTask {
    uint32_t q_off, t_off, mat_off;     // byte offsets into the batch's pool
    int32_t qlen, tlen, m;
    int32_t o_del, e_del, o_ins, e_ins, xtra, size;
    uint32_t he_off, rowmax_off;        // int offsets into the scratch: H and E rows (2 * W), row maxima (tlen)
};
struct SwResult { int32_t score, te, qe, score2, te2, tb, qb, rows; };

constexpr int kSwXByte = 0x10000, kSwXStop = 0x20000, kSwXSubo = 0x40000, kSwXStart = 0x80000;

__global__ void __launch_bounds__(128)
sw_local_kernel(const SwTask* __restrict__ tasks, int n, const uint8_t* __restrict__ pool, int32_t* __restrict__ scratch,
                SwResult* __restrict__ results)
{
    constexpr unsigned kAll = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const SwTask t = tasks[r];
    const uint8_t* __restrict__ q = pool + t.q_off;
    const uint8_t* __restrict__ tg = pool + t.t_off;
    const int8_t* __restrict__ mat = reinterpret_cast<const int8_t*>(pool + t.mat_off);
    const int lanes = t.size == 1 ? 16 : 8;
    const int slen = (t.qlen + lanes - 1) / lanes, W = slen * lanes;
    int32_t* __restrict__ H = scratch + t.he_off;
    int32_t* __restrict__ E = H + W;
    int32_t* __restrict__ rowmax = scratch + t.rowmax_off;
    int minv = 127, maxv = 0;
    for (int a = 0; a < t.m * t.m; ++a) { const int v = mat[a]; minv = v < minv ? v : minv; maxv = v > maxv ? v : maxv; }
    const int shift = (int)(uint8_t)(256 - (int)(uint8_t)minv);
    const int minsc = (t.xtra & kSwXSubo) ? (t.xtra & 0xffff) : 0x10000, endsc = (t.xtra & kSwXStop) ? (t.xtra & 0xffff) : 0x10000;
    const bool subo = (t.xtra & kSwXSubo) != 0;
    SwResult res{0, -1, -1, -1, -1, -1, -1, 0};
    if (t.qlen <= 0) { if (lane == 0) results[r] = res; return; }
    for (int j = lane; j < 2 * W; j += 32) H[j] = 0;              // H and E rows
    __syncwarp();
    const int OFF = t.o_ins + 1;
    const int NEG = -(1 << 29);
    int gmax = 0, te = -1, qe = 0, rows = 0;
    for (int i = 0; i < t.tlen; ++i) {
        const int8_t* __restrict__ srow = mat + (int)tg[i] * t.m;
        int carry_diag = 0, carry_full = NEG, rmax = 0;
        unsigned carry_key = 0u;
        for (int c0 = 0; c0 < W; c0 += 32) {
            const int j = c0 + lane;
            const bool act = j < W;
            const int hold = act ? H[j] : 0, e = act ? E[j] : 0;
            int diag = __shfl_up_sync(kAll, hold, 1);
            if (lane == 0) diag = carry_diag;
            carry_diag = __shfl_sync(kAll, hold, 31);
            const int s = j < t.qlen ? (int)srow[q[j]] : 0;
            int T = diag + s; T = T > 0 ? T : 0; T = T > e ? T : e;
            const int w = act ? T - t.o_ins + j * t.e_ins : NEG;
            const unsigned seg = (unsigned)(j / slen);
            const unsigned key = act ? ((seg << 24) | (unsigned)(w + OFF)) : 0u;
            int incl = w; unsigned ikey = key;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(kAll, incl, d); const unsigned kv = __shfl_up_sync(kAll, ikey, d);
                if (lane >= d) { incl = incl > v ? incl : v; ikey = ikey > kv ? ikey : kv; }
            }
            int excl = __shfl_up_sync(kAll, incl, 1); unsigned ekey = __shfl_up_sync(kAll, ikey, 1);
            if (lane == 0) { excl = NEG; ekey = 0u; }
            excl = excl > carry_full ? excl : carry_full; ekey = ekey > carry_key ? ekey : carry_key;
            const int tot = __shfl_sync(kAll, incl, 31); const unsigned ktot = __shfl_sync(kAll, ikey, 31);
            carry_full = carry_full > tot ? carry_full : tot; carry_key = carry_key > ktot ? carry_key : ktot;
            int ffull = excl - j * t.e_ins; ffull = ffull > 0 ? ffull : 0;
            int fseg = 0;
            if (ekey != 0u && (ekey >> 24) == seg) { fseg = (int)(ekey & 0xffffffu) - OFF - j * t.e_ins; fseg = fseg > 0 ? fseg : 0; }
            const int hs = T > fseg ? T : fseg, h = T > ffull ? T : ffull;
            int en = e - t.e_del; const int open = hs - t.o_del - t.e_del;
            en = en > open ? en : open; en = en > 0 ? en : 0;
            if (act) { H[j] = h; E[j] = en; rmax = rmax > h ? rmax : h; }
        }
        rmax = __reduce_max_sync(kAll, rmax);
        if (t.size == 1 && rmax >= 255 - shift) rmax = 255 - shift;
        if (subo && lane == 0) rowmax[i] = rmax;
        rows = i + 1;
        __syncwarp();
        if (rmax > gmax) {
            gmax = rmax; te = i;
            qe = -1;
            for (int c0 = 0; c0 < W && qe < 0; c0 += 32) {           // first column of the row's maximum (:211-214)
                const int j = c0 + lane;
                const unsigned bal = __ballot_sync(kAll, j < W && H[j] == rmax);
                if (bal) qe = c0 + __ffs(bal) - 1;
            }
            if ((t.size == 1 && gmax + shift >= 255) || gmax >= endsc) break;
        }
    }
    const bool over = t.size == 1 && gmax + shift >= 255;
    res.score = over ? 255 : gmax; res.te = te; res.rows = rows;
    if (!over) {
        res.qe = te >= 0 ? qe : 0;
        if (subo && lane == 0) {                                      // the merged list of :186-194, then :217-225
            __threadfence_block();
            const int x = (res.score + maxv - 1) / (maxv > 0 ? maxv : 1), low = te - x, high = te + x;
            int lv = -1, lr = -2; bool have = false;
            for (int i = 0; i < rows; ++i) {
                const int v = rowmax[i];
                if (v < minsc) continue;
                if (!have || lr + 1 != i) {
                    if (have && (lr < low || lr > high) && lv > res.score2) { res.score2 = lv; res.te2 = lr; }
                    lv = v; lr = i; have = true;
                } else if (lv < v) { lv = v; lr = i; }
            }
            if (have && (lr < low || lr > high) && lv > res.score2) { res.score2 = lv; res.te2 = lr; }
        }
    }
    if (lane == 0) results[r] = res;
}

}  // namespace lb2
