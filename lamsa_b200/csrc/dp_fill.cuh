// dp_fill.cuh -- row-parallel banded affine-gap fill, one warp per task (int32 lanes).
//
// Why a row can be computed column-parallel: in both reference recurrences the
// gap states are fed by the DIAGONAL term M only (src/ksw.c:603,608 / :745,751),
// so inside row i
//     M(i,j)   = H(i-1,j-1) + s(i,j)            (extension: 0 when H(i-1,j-1)==0)
//     E(i+1,j) = max(E(i,j) - e_del, M(i,j) - oe_del)          per-column state
//     F(i,j+1) = max(F(i,j) - e_ins, M(i,j) - oe_ins)          max-plus prefix scan
//     H(i,j)   = max(M, E, F)
// F(i,j) is the exclusive prefix maximum of u_k = t_k + (k+1)*e_ins (k < j)
// minus j*e_ins.  Integer max/add are exact and associative, so every value
// (including the "-inf" arithmetic of the global fill) is bit-identical to
// the sequential loop.
//
// Data layout.  The reference's eh[] array (src/ksw.c:383-385) lives in SHARED
// memory, one circular window per warp: hb[j & (S-1)], eb[j & (S-1)] hold slot
// j, S >= band window + 64.  Slots are initialised to the "row -1" values
// (src/ksw.c:569-572 / :692-694) one column ahead of the band's right edge and
// keep whatever they held when a row does not visit them, exactly like the
// reference array (the adaptive band of the extension reads such slots later:
// SURVEY.md A.2-8).  qb[] holds, per query column, the PRMT selector that
// extracts (sign-extended) the score of that query code from the 8-byte
// scoring-matrix row of the current target base.
//
// Work mapping.  A row's live columns [beg,end) are cut into tiles of 32*G
// columns; in a tile lane L owns G consecutive columns (G in {1,2,4}), so the
// number of warp iterations per row follows the ACTUAL band (adaptive in the
// extension), not the static one.  Per tile: vector LDS of h/e/selectors, G
// cells of straight-line integer code, one Kogge-Stone prefix-max over the
// lane totals, vector STS of the new h/e, one store of G direction nibbles.
#pragma once
#include "dp_device.cuh"
#include <climits>

namespace lb2 {

__device__ __forceinline__ int prmt_s8(uint32_t lo, uint32_t hi, uint32_t sel) {
    int r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
    return r;
}

// selector that extracts byte `code` (0..7) of {lo,hi} and sign-extends it
__device__ __forceinline__ uint32_t sel_for_code(uint32_t code) {
    code &= 7u;
    return code + (code | 8u) * 0x1110u;
}

// "row -1" value of slot j: src/ksw.c:569-572 (global) / :692-694 (extension)
template <int KIND>
__device__ __forceinline__ int init_h(int j, int qlen, int w, int h0, int o_ins, int e_ins) {
    if (KIND == kKindGlobal) {
        if (j == 0) return 0;
        return (j <= qlen && j <= w) ? -(o_ins + e_ins * j) : kNegInf;
    } else {
        if (j == 0) return h0;
        if (j > qlen) return 0;
        int v = h0 - o_ins - e_ins * j;     // h0 - oe_ins - (j-1)*e_ins
        return v > 0 ? v : 0;
    }
}

template <int G> struct Vec;
template <> struct Vec<1> {
    static __device__ __forceinline__ void ld(const int* p, int (&v)[1]) { v[0] = p[0]; }
    static __device__ __forceinline__ void st(int* p, const int (&v)[1]) { p[0] = v[0]; }
    static __device__ __forceinline__ void ldq(const uint16_t* p, uint32_t (&v)[1]) { v[0] = p[0]; }
};
template <> struct Vec<2> {
    static __device__ __forceinline__ void ld(const int* p, int (&v)[2]) {
        int2 t = *reinterpret_cast<const int2*>(p); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void st(int* p, const int (&v)[2]) {
        *reinterpret_cast<int2*>(p) = make_int2(v[0], v[1]); }
    static __device__ __forceinline__ void ldq(const uint16_t* p, uint32_t (&v)[2]) {
        uint32_t t = *reinterpret_cast<const uint32_t*>(p); v[0] = t & 0xffffu; v[1] = t >> 16; }
};
template <> struct Vec<4> {
    static __device__ __forceinline__ void ld(const int* p, int (&v)[4]) {
        int4 t = *reinterpret_cast<const int4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void st(int* p, const int (&v)[4]) {
        *reinterpret_cast<int4*>(p) = make_int4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void ldq(const uint16_t* p, uint32_t (&v)[4]) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = t.x & 0xffffu; v[1] = t.x >> 16; v[2] = t.y & 0xffffu; v[3] = t.y >> 16; }
};


template <int G, int KIND>
__device__ void fill_task(const DTask& T, const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac,
                          uint8_t* __restrict__ zbase, DResult* __restrict__ res,
                          const uint2* __restrict__ smat /* [kMaxMats][8] */,
                          int* __restrict__ hb, int* __restrict__ eb, uint16_t* __restrict__ qb,
                          const int S, const int lane)
{
    constexpr int GS = (G == 1 ? 0 : G == 2 ? 1 : 2);
    constexpr int EINIT = (KIND == kKindGlobal) ? kNegInf : 0;
    constexpr int FINIT = (KIND == kKindGlobal) ? kNegInf : 0;
    const int SM = S - 1;
    const int qlen = T.qlen, tlen = T.tlen, w = T.w, h0 = T.h0;
    const int o_del = T.o_del, e_del = T.e_del, o_ins = T.o_ins, e_ins = T.e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const uint8_t* __restrict__ qseq = query_ptr(T, pool);
    const TargetSrc tsrc = make_target(T, pool, pac);
    const bool want = (T.want_dir & kWantDir) != 0;
    const int RT = T.row_chunks;                        // tiles per stored row
    int2* __restrict__ rowmeta = reinterpret_cast<int2*>(zbase + T.z_off);
    uint8_t* __restrict__ zdir = zbase + T.z_off + (KIND == kKindExtend ? ext_meta_bytes(tlen) : 0);
    const size_t zrow_bytes = (size_t)RT * 32 * dir_lane_bytes(G);
    const uint2* __restrict__ mrows = smat + (int)T.mat_id * 8;
    const int qpad = (qlen + 1 + 31) & ~31;

    // ---- window initialisation: slots 0..send_0, selectors for the first columns
    int slot_hi = (w + 1 < qlen) ? w + 1 : qlen;        // slots [0, slot_hi] are initialised
    for (int j = lane; j <= slot_hi; j += 32) {
        hb[j & SM] = init_h<KIND>(j, qlen, w, h0, o_ins, e_ins);
        eb[j & SM] = EINIT;
    }
    int q_hi = 0;                                       // selectors of columns [.., q_hi) are in qb
    while (q_hi < slot_hi + 1 && q_hi < qpad) {
        qb[(q_hi + lane) & SM] = (uint16_t)sel_for_code(qseq[q_hi + lane]);
        q_hi += 32;
    }
    uint32_t qpre = (q_hi < qpad) ? qseq[q_hi + lane] : 0u;    // next 32 codes, prefetched

    int beg = 0, end = qlen;
    int mx = h0, mx_i = -1, mx_j = -1, mx_ie = -1, gscore = -1, max_off = 0;
    long long cells = 0;
    uint32_t tcur = tsrc.at(lane);
    uint32_t tnext = tsrc.at(32 + lane);
    __syncwarp();
    int i = 0;
    for (; i < tlen; ++i) {
        if ((i & 31) == 0 && i) {
            tcur = tnext;
            tnext = tsrc.at(i + 32 + lane);
        }
        const int tb = __shfl_sync(kFull, (int)tcur, i & 31) & 7;
        const int sbeg = i > w ? i - w : 0;
        const int send = i + w + 1 < qlen ? i + w + 1 : qlen;
        if (KIND == kKindExtend) {
            if (beg < i - w) beg = i - w;
            if (end > send) end = send;
        } else {
            beg = sbeg;
            end = send;
        }
        // admit the column that enters the static window, keep selectors ahead of it
        bool touched = false;
        if (send > slot_hi) {
            slot_hi = send;
            if (lane == 0) {
                hb[send & SM] = init_h<KIND>(send, qlen, w, h0, o_ins, e_ins);
                eb[send & SM] = EINIT;
            }
            touched = true;
        }
        if (q_hi < send + 1 && q_hi < qpad) {
            qb[(q_hi + lane) & SM] = (uint16_t)sel_for_code(qpre);
            q_hi += 32;
            qpre = (q_hi < qpad) ? qseq[q_hi + lane] : 0u;
            touched = true;
        }
        if (touched) __syncwarp();

        const uint2 mrow = mrows[tb];
        int h1init;
        if (KIND == kKindExtend) {
            h1init = 0;
            if (beg == 0) { h1init = h0 - (o_del + e_del * (i + 1)); if (h1init < 0) h1init = 0; }
        } else {
            h1init = beg == 0 ? -(o_del + e_del * (i + 1)) : kNegInf;
        }
        const int base = beg & ~(G - 1);
        const int ntile = end >= base ? ((end - base) >> (5 + GS)) + 1 : 0;
        int carryF = FINIT + beg * e_ins;       // max over finished tiles of u_k, seeded with F(i,beg)
        int carryH = h1init;                    // H(i, j0-1) for lane 0 of the next tile
        int m = 0, mj = -1;                     // row maximum, last argmax (extension)
        int nzlo = INT_MAX, nzhi = -1;          // first / last non-zero slot after this row (extension)
        int hq = 0;                             // H(i, qlen-1) when this lane computes it
        uint8_t* zrow = zdir + (size_t)i * zrow_bytes;

        for (int tile = 0; tile < ntile; ++tile) {
            const int j0 = base + ((tile << 5) + lane) * G;
            const int s0 = j0 & SM;
            const bool mine = (j0 <= end) && (j0 + G > beg);     // any slot of mine in [beg, end]
            int hv[G], ev[G]; uint32_t qs[G];
            int Mv[G], pre[G];
            int lu = INT_MIN;
            if (mine) {
                Vec<G>::ld(hb + s0, hv); Vec<G>::ld(eb + s0, ev); Vec<G>::ldq(qb + s0, qs);
                int ue = (j0 + 1) * e_ins;
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    const int j = j0 + c;
                    const int raw = hv[c];
                    const int s = prmt_s8(mrow.x, mrow.y, qs[c]);
                    int M;
                    if (KIND == kKindExtend) M = raw ? raw + s : 0; else M = raw + s;
                    Mv[c] = M;
                    int t = M - oe_ins;
                    if (KIND == kKindExtend) t = t > 0 ? t : 0;
                    const int u = t + ue;
                    ue += e_ins;
                    pre[c] = lu;
                    if (j >= beg && j < end) lu = lu > u ? lu : u;
                }
            }
            // inclusive prefix maximum of the lane totals
            int incl = lu;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl = incl > o ? incl : o;
            }
            int pin = __shfl_up_sync(kFull, incl, 1);
            if (lane == 0) pin = INT_MIN;
            pin = pin > carryF ? pin : carryF;
            {
                const int tot = __shfl_sync(kFull, incl, 31);
                carryF = carryF > tot ? carryF : tot;
            }
            uint32_t dirw = 0;
            int hnew[G];
            if (mine) {
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    const int j = j0 + c;
                    const bool act = (j >= beg && j < end);
                    const int M = Mv[c];
                    int e = ev[c];
                    const int px = pin > pre[c] ? pin : pre[c];
                    const int f = px - j * e_ins;
                    int h; uint32_t d;
                    if (KIND == kKindExtend) {          // ties: E over M, F over both (src/ksw.c:738-741)
                        d = M > e ? 0u : 1u; h = M > e ? M : e;
                        d = h > f ? d : 2u;  h = h > f ? h : f;
                    } else {                            // ties: M over E over F (src/ksw.c:598-601)
                        d = M >= e ? 0u : 1u; h = M >= e ? M : e;
                        d = h >= f ? d : 2u;  h = h >= f ? h : f;
                    }
                    hnew[c] = h;
                    if (KIND == kKindExtend) {
                        if (act) { mj = m > h ? mj : j; m = m > h ? m : h; }   // last argmax
                        if (j == qlen - 1) hq = h;
                    }
                    int t = M - oe_del;
                    if (KIND == kKindExtend) t = t > 0 ? t : 0;
                    e -= e_del;
                    d |= e > t ? 4u : 0u;
                    e = e > t ? e : t;
                    if (act) ev[c] = e;
                    if (j == end) ev[c] = EINIT;                       // eh[end].e (src/ksw.c:632 / :758)
                    t = M - oe_ins;
                    if (KIND == kKindExtend) t = t > 0 ? t : 0;
                    d |= (f - e_ins) > t ? 8u : 0u;
                    dirw |= d << (4 * c);
                }
            }
            // shifted H row: slot j <- H(i, j-1) for beg <= j <= end
            const int hlast = mine ? hnew[G - 1] : 0;
            int hleft = __shfl_up_sync(kFull, hlast, 1);
            if (lane == 0) hleft = carryH;
            carryH = __shfl_sync(kFull, hlast, 31);
            if (mine) {
#pragma unroll
                for (int c = G - 1; c >= 0; --c) {
                    const int j = j0 + c;
                    int nh = c == 0 ? hleft : hnew[c - 1];
                    if (j == beg) nh = h1init;
                    if (j >= beg && j <= end) hv[c] = nh;
                    if (KIND == kKindExtend) {
                        if (j >= beg && j <= end && (hv[c] | ev[c]) != 0) {
                            if (j < end) nzlo = nzlo < j ? nzlo : j;
                            nzhi = nzhi > j ? nzhi : j;
                        }
                    }
                }
                Vec<G>::st(hb + s0, hv); Vec<G>::st(eb + s0, ev);
                if (want && j0 < end) {
                    uint8_t* p = zrow + (size_t)((tile << 5) + lane) * dir_lane_bytes(G);
                    if (G == 4) *reinterpret_cast<uint16_t*>(p) = (uint16_t)dirw;
                    else *p = (uint8_t)dirw;
                }
            }
        }
        __syncwarp();
        if (want && KIND == kKindExtend && lane == 0) rowmeta[i] = make_int2(beg, end);
        cells += end > beg ? end - beg : 0;
        if (KIND == kKindExtend) {
            const int gm = __reduce_max_sync(kFull, m);
            const int gmj = __reduce_max_sync(kFull, m == gm ? mj : -1);
            const int jfin = beg > end ? beg : end;          // value of j when the column loop ends
            if (jfin == qlen) {                              // src/ksw.c:759-762
                int h1 = h1init;
                if (end > beg) h1 = __shfl_sync(kFull, hq, ((qlen - 1 - base) >> GS) & 31);
                mx_ie = gscore > h1 ? mx_ie : i;
                gscore = gscore > h1 ? gscore : h1;
            }
            if (gm == 0) { ++i; break; }                     // :763
            if (gm > mx) {
                mx = gm; mx_i = i; mx_j = gmj;
                int off = gmj - i; off = off < 0 ? -off : off;
                max_off = max_off > off ? max_off : off;
            } else if (T.zdrop > 0) {                        // :767-773
                const int di = i - mx_i, dj = gmj - mx_j;
                bool drop;
                if (di > dj) drop = mx - gm - (di - dj) * e_del > T.zdrop;
                else         drop = mx - gm - (dj - di) * e_ins > T.zdrop;
                if (drop) { ++i; break; }
            }
            // band trim (:775-778) over the slots as they stand after this row
            int nb = __reduce_min_sync(kFull, nzlo);
            int nh = __reduce_max_sync(kFull, nzhi);
            nb = nb < end ? nb : end;
            if (nh < nb) nh = nb - 1;
            beg = nb;
            end = nh + 2 < qlen ? nh + 2 : qlen;
        }
    }

    // ---- results
    int score = 0, ti = -1, tk = -1;
    if (KIND == kKindGlobal) {
        score = hb[qlen & SM];                               // eh[qlen].h (src/ksw.c:634)
        ti = tlen - 1;
        tk = (ti + w + 1 < qlen ? ti + w + 1 : qlen) - 1;     // :638
    } else {
        score = mx;
        if (gscore <= 0 || gscore <= mx - T.end_bonus) { ti = mx_i; tk = mx_j; }   // :785-789
        else { ti = mx_ie; tk = qlen - 1; }
    }
    if (lane == 0) {
        DResult r;
        r.score = score; r.max_i = mx_i; r.max_j = mx_j; r.max_ie = mx_ie;
        r.gscore = gscore; r.max_off = max_off; r.ti = ti; r.tk = tk;
        r.n_cigar = 0; r.rows = i; r.cigar_off = 0; r.cells = cells;
        *res = r;
    }
    __syncwarp();
}

// Persistent warps: each warp pulls the next task index of its class from a
// global counter (tasks are pre-sorted by descending cost on the host).
// The eh[] windows live in dynamic shared memory (blockDim.x/32 windows of
// warp_smem_bytes(S)); GW = true is the variant for windows that do not fit
// there (bands of tens of thousands of columns, `-V 10000` gaps): the same
// code over a per-warp slice of an L2-resident global scratch.
template <int G, int KIND, bool GW>
__global__ void __launch_bounds__(256)
fill_kernel(const DTask* __restrict__ tasks, const int32_t* __restrict__ order, int n,
            const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac, uint8_t* __restrict__ zbase,
            DResult* __restrict__ results, const uint2* __restrict__ gmat,
            unsigned int* __restrict__ counter, int S, uint8_t* __restrict__ gwin)
{
    __shared__ uint2 smat[kMaxMats * 8];
    extern __shared__ __align__(16) uint8_t dyn[];
    for (int k = threadIdx.x; k < kMaxMats * 8; k += blockDim.x) smat[k] = gmat[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* mine;
    if (GW) mine = gwin + ((size_t)blockIdx.x * (blockDim.x >> 5) + wid) * warp_smem_bytes(S);
    else mine = dyn + (size_t)wid * warp_smem_bytes(S);
    int* hb = reinterpret_cast<int*>(mine);
    int* eb = hb + S;
    uint16_t* qb = reinterpret_cast<uint16_t*>(eb + S);
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= (unsigned)n) break;
        const int idx = order[t];
        fill_task<G, KIND>(tasks[idx], pool, pac, zbase, results + idx, smat, hb, eb, qb, S, lane);
    }
}

}  // namespace lb2
