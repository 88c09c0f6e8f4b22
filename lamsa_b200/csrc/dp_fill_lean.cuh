// dp_fill_lean.cuh -- ksw_extend_core (src/ksw.c:667-807) for LONG, NARROW extensions: band w <= 15, one task per warp,
// the whole eh[] window in registers.
//
// Why: real reads' long DP tasks are extensions of thousands of rows inside the default band of 10 (SURVEY appendix C:
// p99 4 650 rows, max 19 982), and a batch of the producer lasts as long as its longest task -- what counts for them is
// the time of ONE row on ONE warp, i.e. the length of the dependent instruction chain per row, not lanes kept busy.
// The general kernels keep the window in shared memory and walk tiles (dp_fill.cuh: about 1 100 cycles per row alone
// on an SM, the lane-group kernels about 2 100).  Here the static window of a row, columns [i-w, i+w+1], has at most
// 2w+2 <= 32 columns, so column j lives in lane j & 31 for as long as it is inside the window: {H(i-1,j-1), E(i,j)} of
// the reference's eh_t are two registers per lane, the diagonal arrives by one shuffle from lane (j-1) & 31, F is one
// rotated 5-step prefix maximum, the row maximum with its last column one REDUX over (h << 6 | rank) keys, and the
// band trim (:775-778) one ballot rotated to the band's first column.  int32 throughout: no value-range condition
// (the 20 kbp extensions do not fit the packed int16 kernels anyway).
//
// Stale-slot semantics hold as in the window kernels: a lane's registers keep whatever the
// column held when the band moved off it, and a column is initialised to its row -1 value (:692-694) exactly when the
// static window first reaches it (it was never inside any band before).  Direction nibbles and row bands are written
// in the layout of the G = 1 int32 kernel (one byte per cell from the row's `beg`), so the traceback kernels read them
// unchanged.
#pragma once
#include "dp_device.cuh"
#include "dp_fill.cuh"

namespace lb2 {

__device__ void fill_task_lean(const DTask& T, const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac,
                               uint8_t* __restrict__ zbase, DResult* __restrict__ res,
                               const uint2* __restrict__ smat, const int lane)
{
    const int qlen = T.qlen, tlen = T.tlen, w = T.w, h0 = T.h0;
    const int o_del = T.o_del, e_del = T.e_del, o_ins = T.o_ins, e_ins = T.e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins, zdrop = T.zdrop, end_bonus = T.end_bonus;
    const uint8_t* __restrict__ qseq = query_ptr(T, pool);
    const TargetSrc tsrc = make_target(T, pool, pac);
    const bool want = (T.want_dir & kWantDir) != 0;
    int2* __restrict__ rowmeta = reinterpret_cast<int2*>(zbase + T.z_off);
    uint8_t* __restrict__ zdir = zbase + T.z_off + ext_meta_bytes(tlen);
    const size_t zrow_bytes = (size_t)T.row_chunks * 32;
    const uint2* __restrict__ mrows = smat + (int)T.mat_id * 8;
    const int NEG = -(1 << 29);

    // my slot: columns 0 .. min(w+1, qlen) start with their row -1 values
    int slot_hi = (w + 1 < qlen) ? w + 1 : qlen;
    int h = 0, e = 0;
    uint32_t qsel = 0;
    if (lane <= slot_hi) {
        h = init_h<kKindExtend>(lane, qlen, w, h0, o_ins, e_ins);
        if (lane < qlen) qsel = sel_for_code(qseq[lane]);
    }
    const int lm1 = (lane - 1) & 31, lm2 = (lane - 2) & 31, lm3 = (lane - 3) & 31, lm4 = (lane - 4) & 31, lm8 = (lane - 8) & 31,
              lm12 = (lane - 12) & 31, lm16 = (lane - 16) & 31;
    int beg = 0, end = qlen;
    int mx = h0, mx_i = -1, mx_j = -1, mx_ie = -1, gscore = -1, max_off = 0;
    long long cells = 0;
    uint32_t tcur = tsrc.at(lane), tnext = tsrc.at(32 + lane);
    // s(i, my column) is looked up one row ahead, off the row's dependent chain (the lane that admits a new column in
    // row i+1 gets a stale value: that column is not inside the band before row i+2)
    int s;
    { const uint2 mr = mrows[__shfl_sync(kFull, (int)tcur, 0) & 7]; s = prmt_s8(mr.x, mr.y, qsel); }
    uint8_t* zrow = zdir;
    int i = 0;
    for (; i < tlen; ++i, zrow += zrow_bytes) {
        const int lo = i - w;
        const int send = i + w + 1 < qlen ? i + w + 1 : qlen;
        if (send > slot_hi) {                        // the column entering the static window: one lane re-initialises
            slot_hi = send;
            if (lane == (send & 31)) {
                h = init_h<kKindExtend>(send, qlen, w, h0, o_ins, e_ins); e = 0;
                qsel = send < qlen ? sel_for_code(qseq[send]) : 0u;
            }
        }
        if (beg < lo) beg = lo;
        if (end > send) end = send;
        int h1init = 0;
        if (beg == 0) { h1init = h0 - (o_del + e_del * (i + 1)); if (h1init < 0) h1init = 0; }
        const int j = lo + ((lane - lo) & 31);       // the column of the static window that lives in this lane
        const int r = j - beg;                       // its rank in the band
        const bool act = j >= beg && j < end;
        const int M = h ? h + s : 0;                 // :737
        {   // next row's scores
            if (((i + 1) & 31) == 0) { tcur = tnext; tnext = tsrc.at(i + 1 + 32 + lane); }
            const uint2 mr = mrows[__shfl_sync(kFull, (int)tcur, (i + 1) & 31) & 7];
            s = prmt_s8(mr.x, mr.y, qsel);
        }
        int tI = M - oe_ins; tI = tI > 0 ? tI : 0;
        // F(i,j) = max_{k<j} (tI_k - (j-1-k) e_ins), F(i,beg) = 0: exclusive prefix maximum over the ranks of
        // v_k = tI_k + (rank_k + 1) e_ins, across the lanes in band order (rotated by beg)
        // exclusive form from the start (x = left neighbour's v), then three radix-4 rounds (the three shuffles of a round
        // are independent): 1+3 dependent shuffle latencies instead of 5+1
        const int v = act ? tI + (r + 1) * e_ins : NEG;
        int x = __shfl_sync(kFull, v, lm1);
        if (r < 1) x = NEG;
        {
            const int a = __shfl_sync(kFull, x, lm1), b = __shfl_sync(kFull, x, lm2), c = __shfl_sync(kFull, x, lm3);
            int m3 = r >= 2 ? a : NEG; m3 = (r >= 3 && b > m3) ? b : m3; m3 = (r >= 4 && c > m3) ? c : m3;
            x = x > m3 ? x : m3;                      // covers ranks r-4 .. r-1
        }
        {
            const int a = __shfl_sync(kFull, x, lm4), b = __shfl_sync(kFull, x, lm8), c = __shfl_sync(kFull, x, lm12);
            int m3 = r >= 5 ? a : NEG; m3 = (r >= 9 && b > m3) ? b : m3; m3 = (r >= 13 && c > m3) ? c : m3;
            x = x > m3 ? x : m3;                      // covers r-16 .. r-1
        }
        {
            const int a = __shfl_sync(kFull, x, lm16);
            if (r >= 17 && a > x) x = a;              // covers r-32 .. r-1
        }
        int f = x - r * e_ins;
        if (r <= 0) f = 0;
        // H, directions (ties: E over M, F over both, :738-741)
        uint32_t d = M > e ? 0u : 1u;
        int hn = M > e ? M : e;
        d = hn > f ? d : 2u;
        hn = hn > f ? hn : f;
        int t = M - oe_del; t = t > 0 ? t : 0;
        int en = e - e_del;
        d |= en > t ? 4u : 0u;
        en = en > t ? en : t;
        d |= (f - e_ins) > tI ? 8u : 0u;
        if (want) {
            zrow[r & 31] = (uint8_t)d;              // all 32 lanes: one whole sector per row (ranks are distinct mod 32; cells outside the band are never read)
            if (lane == 0) rowmeta[i] = make_int2(beg, end);
        }
        cells += end > beg ? end - beg : 0;
        // row maximum and its LAST column (:743-744)
        const unsigned gkey = __reduce_max_sync(kFull, act ? ((unsigned)hn << 6) | (unsigned)r : 0u);
        const int gm = (int)(gkey >> 6), gmj = beg + (int)(gkey & 63u);
        // the slots after this row: eh[j].h = H(i,j-1) for beg < j <= end, eh[beg].h = first-column value, eh[end].e = 0
        const int hleft = __shfl_sync(kFull, hn, lm1);
        if (j >= beg && j <= end) { h = j == beg ? h1init : hleft; e = j == end ? 0 : en; }
        const int jfin = beg > end ? beg : end;
        if (jfin == qlen) {                               // :759-762
            int h1 = __shfl_sync(kFull, hn, (qlen - 1) & 31);
            if (!(end > beg)) h1 = h1init;
            mx_ie = gscore > h1 ? mx_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (gm == 0) { ++i; break; }                      // :763
        if (gm > mx) {
            mx = gm; mx_i = i; mx_j = gmj;
            int off = gmj - i; off = off < 0 ? -off : off;
            max_off = max_off > off ? max_off : off;
        } else if (zdrop > 0) {                           // :767-773
            const int di = i - mx_i, dj = gmj - mx_j;
            bool drop;
            if (di > dj) drop = mx - gm - (di - dj) * e_del > zdrop;
            else         drop = mx - gm - (dj - di) * e_ins > zdrop;
            if (drop) { ++i; break; }
        }
        // band trim (:775-778): first non-zero slot in [beg,end), last one in [beg',end]; one ballot in band order
        const unsigned bal = __ballot_sync(kFull, j >= beg && j <= end && (h | e) != 0);
        const unsigned rot = __funnelshift_r(bal, bal, beg & 31);          // bit k = column beg + k
        const int span = end - beg;                                         // 1 .. 31 here (gm > 0)
        const unsigned lowb = rot & ((1u << span) - 1u);
        const int nb = lowb ? beg + __ffs(lowb) - 1 : end;
        const unsigned hib = rot & (span >= 31 ? 0xffffffffu : ((2u << span) - 1u)) & ~((1u << (nb - beg)) - 1u);
        const int nh = hib ? beg + 31 - __clz(hib) : nb - 1;
        beg = nb;
        end = nh + 2 < qlen ? nh + 2 : qlen;
    }

    int ti, tk;
    if (gscore <= 0 || gscore <= mx - end_bonus) { ti = mx_i; tk = mx_j; }   // :785-789
    else { ti = mx_ie; tk = qlen - 1; }
    if (lane == 0) {
        DResult rr;
        rr.score = mx; rr.max_i = mx_i; rr.max_j = mx_j; rr.max_ie = mx_ie;
        rr.gscore = gscore; rr.max_off = max_off; rr.ti = ti; rr.tk = tk;
        rr.n_cigar = 0; rr.rows = i; rr.cigar_off = 0; rr.cells = cells;
        *res = rr;
    }
}

// persistent warps pulling tasks from the launch counter, like the other fill kernels
__global__ void __launch_bounds__(128)
fill_lean_kernel(const DTask* __restrict__ tasks, const int32_t* __restrict__ order, int n,
                 const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac, uint8_t* __restrict__ zbase,
                 DResult* __restrict__ results, const uint2* __restrict__ gmat,
                 unsigned int* __restrict__ counter, int /*S*/, uint8_t* __restrict__ /*gwin*/)
{
    __shared__ uint2 smat[kMaxMats * 8];
    for (int k = threadIdx.x; k < kMaxMats * 8; k += blockDim.x) smat[k] = gmat[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= (unsigned)n) break;
        const int idx = order[t];
        fill_task_lean(tasks[idx], pool, pac, zbase, results + idx, smat, lane);
    }
}

}  // namespace lb2
