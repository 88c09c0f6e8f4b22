// dp_fill.cuh -- row-parallel banded affine-gap fill, one warp per task (int32 lanes).
//
// Why a row can be computed column-parallel: in both reference recurrences the
// gap states are fed by the DIAGONAL term M only (src/ksw.c:603,608 / :745,751),
// so inside row i
//     M(i,j)   = H(i-1,j-1) + s(i,j)            (extension: 0 when H(i-1,j-1)==0)
//     E(i+1,j) = max(E(i,j) - e_del, M(i,j) - oe_del)          per-column state
//     F(i,j+1) = max(F(i,j) - e_ins, M(i,j) - oe_ins)          max-plus prefix scan
//     H(i,j)   = max(M, E, F)
// F is an exclusive prefix maximum of u_j = t_j + (j+1)*e_ins, which the warp
// evaluates with one Kogge-Stone pass over the per-lane maxima.  Integer
// max/add are exact and associative, so every value (including the "-inf"
// arithmetic of the global fill) is bit-identical to the sequential loop.
//
// Ownership: lane L holds the reference's eh[] slots of column chunk q
// (columns q*C .. q*C+C-1) with q % 32 == L, in registers.  The band window
// [i-w, i+w+1] slides right one column per row; when a lane's chunk has fallen
// completely left of the window the lane adopts chunk q+32 and initialises the
// slots to the "row -1" values (src/ksw.c:569-572 / :692-694).  31*C >= 2w+1
// (or 32*C >= qlen+1) guarantees the adopted chunk is not needed before the
// old one is dead.  Slots keep whatever they held when a row does not visit
// them, exactly like the reference array (the adaptive band of the extension
// can read such slots later: SURVEY.md A.2-8).
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

template <int C> struct QChunk { uint32_t w[(C + 3) / 4]; };

template <int C>
__device__ __forceinline__ QChunk<C> load_qchunk(const uint8_t* __restrict__ q, int chunk, int qpad) {
    QChunk<C> r;
#pragma unroll
    for (int k = 0; k < (C + 3) / 4; ++k) r.w[k] = 0;
    const long off = (long)chunk * C;
    if (off < qpad) {
        if constexpr (C == 1) r.w[0] = q[off];
        else if constexpr (C == 2) r.w[0] = *reinterpret_cast<const uint16_t*>(q + off);
        else if constexpr (C == 4) r.w[0] = *reinterpret_cast<const uint32_t*>(q + off);
        else if constexpr (C == 8) {
            uint2 v = *reinterpret_cast<const uint2*>(q + off); r.w[0] = v.x; r.w[1] = v.y;
        } else if constexpr (C == 16) {
            uint4 v = *reinterpret_cast<const uint4*>(q + off);
            r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
        } else {
            static_assert(C == 32, "unsupported chunk");
            uint4 v = *reinterpret_cast<const uint4*>(q + off);
            uint4 u = *reinterpret_cast<const uint4*>(q + off + 16);
            r.w[0] = v.x; r.w[1] = v.y; r.w[2] = v.z; r.w[3] = v.w;
            r.w[4] = u.x; r.w[5] = u.y; r.w[6] = u.z; r.w[7] = u.w;
        }
    }
    return r;
}

template <int C>
__device__ __forceinline__ void store_dir(uint8_t* p, const uint32_t (&d)[(C + 7) / 8]) {
    if constexpr (C == 1) *p = (uint8_t)d[0];          // one nibble per byte (C==1 only)
    else if constexpr (C == 2) *p = (uint8_t)d[0];
    else if constexpr (C == 4) *reinterpret_cast<uint16_t*>(p) = (uint16_t)d[0];
    else if constexpr (C == 8) *reinterpret_cast<uint32_t*>(p) = d[0];
    else if constexpr (C == 16) *reinterpret_cast<uint2*>(p) = make_uint2(d[0], d[1]);
    else *reinterpret_cast<uint4*>(p) = make_uint4(d[0], d[1], d[2], d[3]);
}

// bytes of direction storage per chunk
__host__ __device__ constexpr int dir_chunk_bytes(int C) { return C >= 2 ? C / 2 : 1; }

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

template <int C, int KIND>
__device__ void fill_task(const DTask& T, const uint8_t* __restrict__ pool,
                          uint8_t* __restrict__ zbase, DResult* __restrict__ res,
                          const uint2* __restrict__ smat /* [kMaxMats][8] */, const int lane)
{
    constexpr int CS = (C == 1 ? 0 : C == 2 ? 1 : C == 4 ? 2 : C == 8 ? 3 : C == 16 ? 4 : 5);
    constexpr int EINIT = (KIND == kKindGlobal) ? kNegInf : 0;
    constexpr int NDW = (C + 7) / 8;
    const int qlen = T.qlen, tlen = T.tlen, w = T.w, h0 = T.h0;
    const int o_del = T.o_del, e_del = T.e_del, o_ins = T.o_ins, e_ins = T.e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const uint8_t* __restrict__ qseq = pool + (size_t)T.q_off32 * 32;
    const uint8_t* __restrict__ tseq = pool + (size_t)T.t_off32 * 32;
    const int qpad = (qlen + 1 + 31) & ~31;
    const bool want = T.want_dir != 0;
    const int RW = T.row_chunks;
    int2* __restrict__ rowmeta = reinterpret_cast<int2*>(zbase + T.z_off);
    uint8_t* __restrict__ zdir = zbase + T.z_off + (KIND == kKindExtend ? ext_meta_bytes(tlen) : 0);
    const uint2* __restrict__ mrows = smat + (int)T.mat_id * 8;

    int hs[C], es[C];
    uint32_t qsel[C];
    int chunk = lane;
    QChunk<C> qnext;

    auto adopt = [&](const QChunk<C>& qc) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j = chunk * C + c;
            hs[c] = init_h<KIND>(j, qlen, w, h0, o_ins, e_ins);
            es[c] = EINIT;
            qsel[c] = sel_for_code((qc.w[c / 4] >> (8 * (c & 3))) & 0xffu);
        }
    };
    {
        QChunk<C> q0 = load_qchunk<C>(qseq, chunk, qpad);
        qnext = load_qchunk<C>(qseq, chunk + 32, qpad);
        adopt(q0);
    }

    int beg = 0, end = qlen;
    int mx = h0, mx_i = -1, mx_j = -1, mx_ie = -1, gscore = -1, max_off = 0;
    long long cells = 0;
    const int tpad = (tlen + 31) & ~31;
    uint32_t tcur = (lane < tpad) ? tseq[lane] : 0u;
    uint32_t tnext = (32 + lane < tpad) ? tseq[32 + lane] : 0u;
    int i = 0;
    for (; i < tlen; ++i) {
        if ((i & 31) == 0 && i) {
            tcur = tnext;
            tnext = (i + 32 + lane < tpad) ? tseq[i + 32 + lane] : 0u;
        }
        const int tb = __shfl_sync(kFull, (int)tcur, i & 31) & 7;
        const int sbeg = i > w ? i - w : 0;
        if (KIND == kKindExtend) {
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
        } else {
            beg = sbeg;
            end = i + w + 1 < qlen ? i + w + 1 : qlen;
        }
        if ((chunk + 1) * C <= sbeg) {          // my chunk is dead: adopt the next one
            chunk += 32;
            QChunk<C> qc = qnext;
            qnext = load_qchunk<C>(qseq, chunk + 32, qpad);
            adopt(qc);
        }
        const int j0 = chunk * C;
        const uint2 mrow = mrows[tb];
        int h1init;
        if (KIND == kKindExtend) {
            h1init = 0;
            if (beg == 0) { h1init = h0 - (o_del + e_del * (i + 1)); if (h1init < 0) h1init = 0; }
        } else {
            h1init = beg == 0 ? -(o_del + e_del * (i + 1)) : kNegInf;
        }
        // lane-local active range [lo, hi) in chunk coordinates
        int lo = beg - j0; lo = lo < 0 ? 0 : (lo > C ? C : lo);
        int hi = end - j0; hi = hi < 0 ? 0 : (hi > C ? C : hi);

        // ---- pass 1: M and the lane maximum of u_j = t_j + (j+1)*e_ins
        int Mv[C];
        int U = INT_MIN;
        int ue = (j0 + 1) * e_ins;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int raw = hs[c];
            const int s = prmt_s8(mrow.x, mrow.y, qsel[c]);
            int M;
            if (KIND == kKindExtend) M = raw ? raw + s : 0; else M = raw + s;
            Mv[c] = M;
            int t = M - oe_ins;
            if (KIND == kKindExtend) t = t > 0 ? t : 0;
            const int u = t + ue;
            ue += e_ins;
            if (c >= lo && c < hi) U = U > u ? U : u;
        }
        // ---- exclusive prefix maximum across lanes, in column order
        const int lane_beg = (beg >> CS) & 31;
        const int rank = (lane - lane_beg) & 31;
        const int nact = end > beg ? ((end - 1) >> CS) - (beg >> CS) + 1 : 0;   // lanes with active cells
        int v = U;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            if (d < nact) {
                const int o = __shfl_sync(kFull, v, (lane - d) & 31);
                if (rank >= d) v = v > o ? v : o;
            }
        }
        int P = __shfl_sync(kFull, v, (lane - 1) & 31);
        if (rank == 0) P = INT_MIN;
        const int ja = j0 > beg ? j0 : beg;
        {
            const int base = (KIND == kKindGlobal ? kNegInf : 0) + beg * e_ins;
            P = P > base ? P : base;
        }
        int f = P - ja * e_ins;

        // ---- pass 2: H, E', F', direction nibbles, row maximum
        int m = 0, mj = -1;
        uint32_t dirw[NDW];
#pragma unroll
        for (int k = 0; k < NDW; ++k) dirw[k] = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const bool act = (c >= lo && c < hi);
            const int M = Mv[c];
            int e = es[c];
            int h; uint32_t d;
            if (KIND == kKindExtend) {          // ties: E over M, F over both (src/ksw.c:738-741)
                d = M > e ? 0u : 1u; h = M > e ? M : e;
                d = h > f ? d : 2u;  h = h > f ? h : f;
            } else {                            // ties: M over E over F (src/ksw.c:598-601)
                d = M >= e ? 0u : 1u; h = M >= e ? M : e;
                d = h >= f ? d : 2u;  h = h >= f ? h : f;
            }
            Mv[c] = h;                          // Mv now holds H(i, j)
            if (KIND == kKindExtend) {
                if (act) { mj = m > h ? mj : j0 + c; m = m > h ? m : h; }   // last argmax
            }
            int t = M - oe_del;
            if (KIND == kKindExtend) t = t > 0 ? t : 0;
            e -= e_del;
            d |= e > t ? 4u : 0u;
            e = e > t ? e : t;
            if (act) es[c] = e;
            t = M - oe_ins;
            if (KIND == kKindExtend) t = t > 0 ? t : 0;
            int f2 = f - e_ins;
            d |= f2 > t ? 8u : 0u;
            f2 = f2 > t ? f2 : t;
            if (act) f = f2;
            dirw[c / 8] |= d << (4 * (c & 7));
        }
        // ---- commit the shifted H row: slot j <- H(i, j-1) for beg <= j <= end
        const int hleft = __shfl_sync(kFull, Mv[C - 1], (lane - 1) & 31);
        {
            const int clo = beg - j0, chi = end - j0;   // inclusive range [clo, chi]
#pragma unroll
            for (int c = C - 1; c >= 0; --c) {
                int hv = c == 0 ? hleft : Mv[c - 1];
                if (c == clo) hv = h1init;
                if (c >= clo && c <= chi) hs[c] = hv;
                if (c == chi && chi >= clo) es[c] = EINIT;
            }
        }
        if (want) {
            if (KIND == kKindExtend && lane == 0) rowmeta[i] = make_int2(beg, end);
            if (hi > lo) {
                const long rel = (long)i * RW + (chunk - (sbeg >> CS));
                store_dir<C>(zdir + rel * dir_chunk_bytes(C), dirw);
            }
        }
        cells += end > beg ? end - beg : 0;
        if (KIND == kKindExtend) {
            const int gm = __reduce_max_sync(kFull, m);
            const int gmj = __reduce_max_sync(kFull, m == gm ? mj : -1);
            const int jfin = beg > end ? beg : end;          // value of j when the column loop ends
            if (jfin == qlen) {                              // src/ksw.c:759-762
                int h1 = h1init;
                if (end > beg) {                             // H(i, qlen-1) lives in some lane's Mv
                    const int cq = (qlen - 1) & (C - 1);
                    int hv = 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) if (c == cq) hv = Mv[c];
                    h1 = __shfl_sync(kFull, hv, ((qlen - 1) >> CS) & 31);
                }
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
            uint32_t nzmask = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) nzmask |= ((hs[c] | es[c]) != 0 ? 1u : 0u) << c;
            const int clo = beg - j0, chi = end - j0;        // slots [beg, end]
            uint32_t inmask = 0;
            if (chi >= 0 && clo < C && chi >= clo) {
                const int a = clo < 0 ? 0 : clo, b = chi > C - 1 ? C - 1 : chi;
                inmask = (b - a + 1 >= 32) ? 0xffffffffu : (((1u << (b - a + 1)) - 1u) << a);
            }
            nzmask &= inmask;
            uint32_t lomask = nzmask;
            if (chi >= 0 && chi < C) lomask &= ~(1u << chi);     // first scan excludes slot `end`
            const int mylo = lomask ? j0 + __ffs(lomask) - 1 : INT_MAX;
            const int myhi = nzmask ? j0 + 31 - __clz(nzmask) : -1;
            int nb = __reduce_min_sync(kFull, mylo);
            int nh = __reduce_max_sync(kFull, myhi);
            nb = nb < end ? nb : end;
            if (nh < nb) nh = nb - 1;
            beg = nb;
            end = nh + 2 < qlen ? nh + 2 : qlen;
        }
    }

    // ---- results
    int score = 0, ti = -1, tk = -1;
    if (KIND == kKindGlobal) {
        // eh[qlen].h after the last row (src/ksw.c:634)
        const int cq = qlen & (C - 1);
        int hv = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) if (c == cq) hv = hs[c];
        const bool owner = chunk == (qlen >> CS);
        const unsigned who = __ballot_sync(kFull, owner);
        score = who ? __shfl_sync(kFull, hv, __ffs(who) - 1) : init_h<KIND>(qlen, qlen, w, h0, o_ins, e_ins);
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
}

// Persistent warps: each warp pulls the next task index of its class from a
// global counter (tasks are pre-sorted by descending cost on the host).
template <int C, int KIND>
__global__ void __launch_bounds__(128)
fill_kernel(const DTask* __restrict__ tasks, const int32_t* __restrict__ order, int n,
            const uint8_t* __restrict__ pool, uint8_t* __restrict__ zbase,
            DResult* __restrict__ results, const uint2* __restrict__ gmat,
            unsigned int* __restrict__ counter)
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
        fill_task<C, KIND>(tasks[idx], pool, zbase, results + idx, smat, lane);
    }
}

}  // namespace lb2
