// dp_fill16.cuh -- the packed-int16 fill: same row-parallel scheme as dp_fill.cuh
// (shared-memory eh[] window, band-following tiles, warp prefix-max for F), but
// two adjacent query columns share one 32-bit register and every max/add is a
// Blackwell DPX instruction on s16x2 lanes:
//     VIMNMX.S16x2 (with the two "which operand won" predicates -> direction
//     bits for free), VIADDMNMX.S16x2(.RELU), VIADD.16x2, PRMT for half shifts.
// Lane L of a tile owns G = 2*NP consecutive columns = NP packed registers.
//
// Exactness.  The host routes a task here only when every value the reference
// computes in int32 provably fits the int16 domain used below:
//   extension: 0 <= H,E,F <= h0 + qlen*max(mat) (<= 16383, because 2H is formed),
//              max(mat) <= 1, plus the row-relative scan offset (ncol+2G+2)*e_ins;
//   global:    every finite value >= L (all-mismatch / all-gap bound) > kNeg16+512,
//              the reference's -2^30 "minus infinity" becomes kNeg16: it only ever
//              LOSES a max against a finite value and is subtracted from at most
//              once before being replaced, so no decision changes.
// Extension detail: M = H ? H+s : 0 (src/ksw.c:737) is evaluated as
// min(H+s, 2H); for H==0 this yields min(s,0) <= 0 instead of 0, and every use
// of M (M>e with e>=0, max(M,e), max(M-oe,0)) is identical for all M <= 0.
//
// Direction nibble of a cell here holds the four raw predicates
//   bit0  (E >= M) [extension]  /  (M >= E) [global]
//   bit1  (F >= h) [extension]  /  (h >= F) [global]
//   bit2  (M-oe_del clamp >= E-e_del)   -> E' opened   (extension flag = !bit2)
//   bit3  (M-oe_ins clamp >= F-e_ins)   -> F' opened   (extension flag = !bit3)
// (DTask::dir_fmt = 1); dp_trace.cuh decodes both formats.
#pragma once
#include "dp_device.cuh"
#include "dp_fill.cuh"

#ifndef LB2_FILL16_MIN_BLOCKS
#define LB2_FILL16_MIN_BLOCKS 2
#endif

namespace lb2 {

constexpr int kNeg16 = -32000;

__device__ __forceinline__ uint32_t pk2(int lo, int hi) {
    return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16);
}
__device__ __forceinline__ uint32_t dup2(int v) { return pk2(v, v); }
__device__ __forceinline__ int lo16(uint32_t x) { return (int)(short)(x & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t x) { return (int)x >> 16; }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
// (f(a,b,c)) bitwise select: (a & m) | (b & ~m)
__device__ __forceinline__ uint32_t blend(uint32_t a, uint32_t b, uint32_t m) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(r) : "r"(m), "r"(a), "r"(b));   // m ? a : b
    return r;
}

// d += bit when the predicate holds.  The integer ALU pipe is what bounds these kernels while the
// FMA pipe idles, so the add is written as a guarded multiply-add by a value ptxas cannot prove to
// be 1: it issues as `@P IMAD` on the FMA pipe (a plain guarded add would go to the ALU pipe, and
// the C expression becomes a SEL + IADD3 tree there).
__device__ __forceinline__ void add_flag(uint32_t& d, bool p, uint32_t bit, uint32_t one) {
    asm("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q mad.lo.u32 %0, %2, %3, %0; }" : "+r"(d) : "r"((uint32_t)p), "r"(one), "r"(bit));
}

// selector pair for two adjacent query codes: {row[c0], sign, row[c1], sign}
__device__ __forceinline__ uint32_t sel_for_pair(uint32_t c0, uint32_t c1) {
    c0 &= 7u; c1 &= 7u;
    return c0 | ((c0 | 8u) << 4) | (c1 << 8) | ((c1 | 8u) << 12);
}

// "row -1" value of slot j in the packed domain.  The GLOBAL fill works in the "hat" domain
// X^(i,j) = X(i,j) + j*e_ins (slot j holds H(i-1,j-1) + (j-1)*e_ins): there F^ is a plain prefix
// maximum of u^_k = M^_k - o_ins, so the row-relative scan offsets disappear; every comparison
// of the reference is between quantities of the same column and is unchanged by the offset.
template <int KIND>
__device__ __forceinline__ int init_h16(int j, int qlen, int w, int h0, int o_ins, int e_ins) {
    if (KIND == kKindGlobal) {
        if (j == 0) return -e_ins;                                   // H(-1,-1) = 0, column -1
        return (j <= qlen && j <= w) ? -(o_ins + e_ins) : kNeg16;    // -(o+e*j) + (j-1)*e
    }
    return init_h<KIND>(j, qlen, w, h0, o_ins, e_ins);
}

// per-task copy of the scoring-matrix rows: global fill adds e_ins to every byte (s^ = s + e_ins)
__device__ __forceinline__ uint32_t add_bytes(uint32_t x, uint32_t e4) {
    return ((x & 0x7f7f7f7fu) + (e4 & 0x7f7f7f7fu)) ^ ((x ^ e4) & 0x80808080u);
}
template <int KIND>
__device__ __forceinline__ void stage_matrix(uint2* __restrict__ mrw, const uint2* __restrict__ rows, int e_ins, int gl) {
    if (gl < 8) {
        uint2 v = rows[gl];
        if (KIND == kKindGlobal) {
            const uint32_t e4 = (uint32_t)(e_ins & 0xff) * 0x01010101u;
            v.x = add_bytes(v.x, e4); v.y = add_bytes(v.y, e4);
        }
        mrw[gl] = v;
    }
}

// per-block table of half masks: entry [lo*(G+1)+hi][p] = mask of columns c in [lo,hi) of pair p
template <int NP> __host__ __device__ constexpr int mask_table_words() { return (2 * NP + 1) * (2 * NP + 1) * NP; }

template <int NP> struct PVec;
template <> struct PVec<2> {
    static __device__ __forceinline__ void ld(const int16_t* p, uint32_t (&v)[2]) {
        uint2 t = *reinterpret_cast<const uint2*>(p); v[0] = t.x; v[1] = t.y; }
    static __device__ __forceinline__ void st(int16_t* p, const uint32_t (&v)[2]) {
        *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1]); }
    static __device__ __forceinline__ void ldq(const uint16_t* p, uint32_t (&v)[2]) {
        uint32_t t = *reinterpret_cast<const uint32_t*>(p); v[0] = t & 0xffffu; v[1] = t >> 16; }
    static __device__ __forceinline__ void ldm(const uint32_t* p, uint32_t (&v)[2]) {
        uint2 t = *reinterpret_cast<const uint2*>(p); v[0] = t.x; v[1] = t.y; }
};
template <> struct PVec<4> {
    static __device__ __forceinline__ void ld(const int16_t* p, uint32_t (&v)[4]) {
        uint4 t = *reinterpret_cast<const uint4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    static __device__ __forceinline__ void st(int16_t* p, const uint32_t (&v)[4]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]); }
    static __device__ __forceinline__ void ldq(const uint16_t* p, uint32_t (&v)[4]) {
        uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = t.x & 0xffffu; v[1] = t.x >> 16; v[2] = t.y & 0xffffu; v[3] = t.y >> 16; }
    static __device__ __forceinline__ void ldm(const uint32_t* p, uint32_t (&v)[4]) {
        uint4 t = *reinterpret_cast<const uint4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};

template <int NP, int KIND>
__device__ void fill_task16(const DTask& T, const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac,
                            uint8_t* __restrict__ zbase, DResult* __restrict__ res,
                            const uint2* __restrict__ smat, const uint32_t* __restrict__ mtab,
                            int16_t* __restrict__ hb, int16_t* __restrict__ eb, uint16_t* __restrict__ qb,
                            uint2* __restrict__ mrw, const int S, const int lane)
{
    constexpr int G = 2 * NP;
    constexpr int GS = (NP == 2 ? 2 : 3);
    constexpr bool EXT = (KIND == kKindExtend);
    constexpr int EINIT = EXT ? 0 : kNeg16;
    constexpr int FINIT = EXT ? 0 : kNeg16;
    constexpr int POISON = EXT ? 0 : kNeg16;           // H of dead slots left of the band (see below)
    const int SM = S - 1, SMQ = SM >> 1;
    // the value 1, opaque to ptxas (S is a power of two): lets add_flag issue on the FMA pipe
    const uint32_t one = (uint32_t)S >> (31 - __clz(S));
    const int qlen = T.qlen, tlen = T.tlen, w = T.w, h0 = T.h0;
    const int o_del = T.o_del, e_del = T.e_del, o_ins = T.o_ins, e_ins = T.e_ins;
    const uint8_t* __restrict__ qseq = query_ptr(T, pool);
    const TargetSrc tsrc = make_target(T, pool, pac);
    const bool want = (T.want_dir & kWantDir) != 0;
    int2* __restrict__ rowmeta = reinterpret_cast<int2*>(zbase + T.z_off);
    uint8_t* __restrict__ zdir = zbase + T.z_off + (EXT ? ext_meta_bytes(tlen) : 0) + (size_t)lane * (G / 2);
    const size_t zrow_bytes = (size_t)T.row_chunks * 32 * (G / 2);
    const uint2* __restrict__ mrows = smat + (int)T.mat_id * 8;
    const int qpad = (qlen + 1 + 31) & ~31;

    // packed constants
    const uint32_t NEGP = dup2(kNeg16);
    const uint32_t N_OE_INS = dup2(-(o_ins + e_ins)), N_OE_DEL = dup2(-(o_del + e_del));
    const uint32_t N_E_INS = dup2(-e_ins), N_E_DEL = dup2(-e_del), N_O_INS = dup2(-o_ins);
    const uint32_t TILE_STEP = dup2(32 * G * e_ins), N_TILE_STEP = dup2(-32 * G * e_ins);
    // row-relative scan offsets of my pairs in tile 0: RO1 = ((r+1)e,(r+2)e), NRO = (-r e, -(r+1)e)
    uint32_t RO1_0[NP], NRO_0[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int r = lane * G + 2 * p;
        RO1_0[p] = pk2((r + 1) * e_ins, (r + 2) * e_ins);
        NRO_0[p] = pk2(-r * e_ins, -(r + 1) * e_ins);
    }

    stage_matrix<KIND>(mrw, mrows, e_ins, lane);
    // ---- window initialisation: slots [0, send_0]; selectors for columns [0, w+66)
    int slot_hi = (w + 1 < qlen) ? w + 1 : qlen;        // slots [0, slot_hi] are initialised
    for (int j = lane; j <= slot_hi; j += 32) {
        hb[j & SM] = (int16_t)init_h16<KIND>(j, qlen, w, h0, o_ins, e_ins);
        eb[j & SM] = (int16_t)EINIT;
    }
    int q_hi = 0;                                       // selectors of columns [.., q_hi) are in qb
    {
        const int want_q = w + 66 < qpad ? w + 66 : qpad;
        while (q_hi < want_q) {
            if (lane < 16) {
                const uint32_t cc = ld_pair(qseq + q_hi + 2 * lane);
                qb[((q_hi >> 1) + lane) & SMQ] = (uint16_t)sel_for_pair(cc & 0xffu, cc >> 8);
            }
            q_hi += 32;
        }
    }
    uint32_t qpre = (q_hi < qpad && lane < 16) ? ld_pair(qseq + q_hi + 2 * lane) : 0u;
    uint32_t tcur = 0u;
    uint32_t tnext = tsrc.at(lane);

    int beg = 0, end = qlen;
    int mx = h0, mx_i = -1, mx_j = -1, mx_ie = -1, gscore = -1, max_off = 0;
    long long cells = 0;
    int i = 0;
    for (; i < tlen; ++i) {
        if ((i & 31) == 0) {                            // every 32 rows: next target codes, next 32 selectors
            tcur = tnext;
            tnext = tsrc.at(i + 32 + lane);
            if (i && q_hi < qpad) {
                if (lane < 16) qb[((q_hi >> 1) + lane) & SMQ] = (uint16_t)sel_for_pair(qpre & 0xffu, qpre >> 8);
                q_hi += 32;
                qpre = (q_hi < qpad && lane < 16) ? ld_pair(qseq + q_hi + 2 * lane) : 0u;
            }
        }
        const int tb = __shfl_sync(kFull, (int)tcur, i & 31) & 7;
        const int send = i + w + 1 < qlen ? i + w + 1 : qlen;
        if (EXT) {
            if (beg < i - w) beg = i - w;
            if (end > send) end = send;
        } else {
            beg = i > w ? i - w : 0;
            end = send;
        }
        const int base = beg & ~(G - 1);
        // admit the column entering the static window; poison the dead slots [base, beg): the row
        // computes them unmasked, and H = POISON keeps their u below the F(i,beg) seed
        if (lane == 0 && send > slot_hi) {
            hb[send & SM] = (int16_t)init_h16<KIND>(send, qlen, w, h0, o_ins, e_ins);
            eb[send & SM] = (int16_t)EINIT;
        }
        slot_hi = send > slot_hi ? send : slot_hi;
        if (lane < G - 1 && base + lane < beg) hb[(base + lane) & SM] = (int16_t)POISON;
        __syncwarp();

        const uint2 mrow = mrw[tb];
        int h1init;
        if (EXT) {
            h1init = 0;
            if (beg == 0) { h1init = h0 - (o_del + e_del * (i + 1)); if (h1init < 0) h1init = 0; }
        } else {
            h1init = beg == 0 ? -(o_del + e_del * (i + 1)) - e_ins : kNeg16;    // hat: column -1
        }
        const int rb = beg - base, re = end - base;            // live columns, row-relative [rb, re)
        const int ntile = end >= base ? ((end - base) >> (5 + GS)) + 1 : 0;
        uint32_t carryF = dup2(EXT ? FINIT + rb * e_ins : kNeg16);            // seeded with F(i,beg) in the u-domain
        uint32_t carryH = 0;                                   // left neighbour's last pair, previous tile
        uint32_t mrowmax[NP];                                  // extension: running row maxima per pair
        int mt_lo[NP], mt_hi[NP];                              // tile of the last (tie-)update
#pragma unroll
        for (int p = 0; p < NP; ++p) { mrowmax[p] = 0; mt_lo[p] = -1; mt_hi[p] = -1; }
        uint8_t* zp = zdir + (size_t)i * zrow_bytes;
        int r0 = lane * G;

        for (int tile = 0; tile < ntile; ++tile, r0 += 32 * G, zp += 32 * (G / 2)) {
            const int s0 = (base + r0) & SM;
            uint32_t H[NP], E[NP], qs[NP];
            PVec<NP>::ld(hb + s0, H); PVec<NP>::ld(eb + s0, E); PVec<NP>::ldq(qb + (s0 >> 1), qs);
            uint32_t am[NP], cm[NP];
            if (EXT) {
                // lane-local live ranges in columns 0..G: cells [lo,hi), slots [lo,hc)
                const int lo = __viaddmin_s32_relu(rb, -r0, G);          // max(min(rb-r0,G),0)
                const int hi = __viaddmin_s32_relu(re, -r0, G);
                const int hc = __viaddmin_s32_relu(re + 1, -r0, G);
                PVec<NP>::ldm(mtab + (lo * (G + 1) + hi) * NP, am);
                PVec<NP>::ldm(mtab + (lo * (G + 1) + hc) * NP, cm);
            }

            uint32_t M[NP], tI[NP], pre[NP];
            uint32_t run = NEGP;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t s = prmt(mrow.x, mrow.y, qs[p]);
                if (EXT) {
                    M[p] = __viaddmin_s16x2(H[p], s, __vadd2(H[p], H[p]));
                    tI[p] = __viaddmax_s16x2_relu(M[p], N_OE_INS, 0u);
                } else {
                    M[p] = __vadd2(H[p], s);                     // hat domain: s already holds s + e_ins
                    tI[p] = __vadd2(M[p], N_O_INS);              // u^ = M^ - o_ins
                }
                const uint32_t u = EXT ? __vadd2(tI[p], RO1_0[p]) : tI[p];
                // exclusive prefix inside the lane: (run, max(run, u.lo))
                pre[p] = __vmaxs2(run, prmt(u, NEGP, 0x1054));
                run = __vimax3_s16x2(run, u, prmt(u, 0u, 0x1032));
            }
            // inclusive prefix maximum of the lane totals (both halves carry the same value)
            uint32_t incl = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) incl = __vmaxs2(incl, __shfl_up_sync(kFull, incl, d));
            uint32_t pin = __shfl_up_sync(kFull, incl, 1);
            if (lane == 0) pin = NEGP;
            pin = __vmaxs2(pin, carryF);
            carryF = __vmaxs2(carryF, __shfl_sync(kFull, incl, 31));
            if (EXT) carryF = __vadd2(carryF, N_TILE_STEP);          // scan offsets are tile-relative: move the carry into the next tile's domain

            uint32_t dirw = 0;
            uint32_t Hn[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t F = EXT ? __vadd2(__vmaxs2(pin, pre[p]), NRO_0[p]) : __vmaxs2(pin, pre[p]);
                bool a_hi, a_lo, b_hi, b_lo, c_hi, c_lo, d_hi, d_lo;
                uint32_t h;
                if (EXT) {                           // ties: E over M, F over both (src/ksw.c:738-741)
                    h = __vibmax_s16x2(E[p], M[p], &a_hi, &a_lo);
                    h = __vibmax_s16x2(F, h, &b_hi, &b_lo);
                } else {                             // ties: M over E over F (src/ksw.c:598-601)
                    h = __vibmax_s16x2(M[p], E[p], &a_hi, &a_lo);
                    h = __vibmax_s16x2(h, F, &b_hi, &b_lo);
                }
                Hn[p] = h;
                uint32_t tD;
                if (EXT) tD = __viaddmax_s16x2_relu(M[p], N_OE_DEL, 0u);
                else tD = __vadd2(M[p], N_OE_DEL);
                const uint32_t En = __vibmax_s16x2(tD, __vadd2(E[p], N_E_DEL), &c_hi, &c_lo);
                // F' opened?  tI >= F - e_ins; in the hat domain that is u^ >= F^
                (void)__vibmax_s16x2(tI[p], EXT ? __vadd2(F, N_E_INS) : F, &d_hi, &d_lo);
                E[p] = EXT ? blend(En, E[p], am[p]) : En;
                add_flag(dirw, a_lo, 1u << (8 * p), one);  add_flag(dirw, b_lo, 2u << (8 * p), one);
                add_flag(dirw, c_lo, 4u << (8 * p), one);  add_flag(dirw, d_lo, 8u << (8 * p), one);
                add_flag(dirw, a_hi, 16u << (8 * p), one); add_flag(dirw, b_hi, 32u << (8 * p), one);
                add_flag(dirw, c_hi, 64u << (8 * p), one); add_flag(dirw, d_hi, 128u << (8 * p), one);
                if (EXT) {
                    bool m_hi, m_lo;
                    const uint32_t hm = h | ~am[p];              // inactive columns read -1: never >= max
                    mrowmax[p] = __vibmax_s16x2(hm, mrowmax[p], &m_hi, &m_lo);
                    if (m_lo) mt_lo[p] = tile;
                    if (m_hi) mt_hi[p] = tile;
                }
            }
            // shifted H row: slot j <- H(i, j-1); my first slot takes the left neighbour's last column
            uint32_t left = __shfl_up_sync(kFull, Hn[NP - 1], 1);
            if (lane == 0) left = carryH;
            carryH = __shfl_sync(kFull, Hn[NP - 1], 31);
#pragma unroll
            for (int p = NP - 1; p >= 0; --p) {
                const uint32_t sh = prmt(p == 0 ? left : Hn[p - 1], Hn[p], 0x5432);    // (prev.hi, Hn[p].lo)
                H[p] = EXT ? blend(sh, H[p], cm[p]) : sh;
            }
            // lanes right of `end` hold nothing of this row (and would alias live slots of the window)
            if (r0 <= re) { PVec<NP>::st(hb + s0, H); PVec<NP>::st(eb + s0, E); }
            if (want) {
                if (NP == 2) *reinterpret_cast<uint16_t*>(zp) = (uint16_t)dirw;
                else *reinterpret_cast<uint32_t*>(zp) = dirw;
            }
        }
        __syncwarp();
        // single-slot fix-ups of the row: eh[beg].h = first-column value, eh[end].e = init
        if (lane == 0 && end >= beg) {
            hb[beg & SM] = (int16_t)h1init;
            eb[end & SM] = (int16_t)EINIT;
        }
        if (EXT && want && lane == 0) rowmeta[i] = make_int2(beg, end);
        cells += end > beg ? end - beg : 0;
        if (EXT) {
            __syncwarp();
            // row maximum and its LAST column (src/ksw.c:743-744): one key per accumulator,
            // value in the high half, column+1 in the low half (values <= 16000, columns < 65535
            // on this path), so a single max-reduction yields both; untouched accumulators are key 0
            // = (m 0, mj -1), the reference's initial state
            uint32_t key = 0;
            const int cbase = base + lane * G + 1;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t klo = (mrowmax[p] << 16) | (uint32_t)(mt_lo[p] < 0 ? 0 : cbase + (mt_lo[p] << (5 + GS)) + 2 * p);
                const uint32_t khi = (mrowmax[p] & 0xffff0000u) | (uint32_t)(mt_hi[p] < 0 ? 0 : cbase + (mt_hi[p] << (5 + GS)) + 2 * p + 1);
                key = key > klo ? key : klo;
                key = key > khi ? key : khi;
            }
            const uint32_t gkey = __reduce_max_sync(kFull, key);
            const int gm = (int)(gkey >> 16), gmj = (int)(gkey & 0xffffu) - 1;
            const int jfin = beg > end ? beg : end;
            if (jfin == qlen) {                              // src/ksw.c:759-762
                const int h1 = end > beg ? (int)hb[qlen & SM] : h1init;     // eh[end].h == H(i, qlen-1)
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
            // band trim (:775-778): first non-zero slot in [beg,end), last one in [beg',end]
            int nb = end, nh;
            if (end - beg < 32) {                            // whole band under one ballot
                const int j = beg + lane;
                const bool nz = j <= end && (hb[j & SM] != 0 || eb[j & SM] != 0);
                const unsigned bal = __ballot_sync(kFull, nz);
                const unsigned lowb = bal & ~(1u << (end - beg));          // first scan excludes slot `end`
                if (lowb) nb = beg + __ffs(lowb) - 1;
                const unsigned hib = bal & (0xffffffffu << (nb - beg));     // second scan starts at beg'
                nh = hib ? beg + 31 - __clz(hib) : nb - 1;
            } else {
                for (int st = beg; st < end; st += 32) {
                    const int j = st + lane;
                    const bool nz = j < end && (hb[j & SM] != 0 || eb[j & SM] != 0);
                    const unsigned bal = __ballot_sync(kFull, nz);
                    if (bal) { nb = st + __ffs(bal) - 1; break; }
                }
                nh = nb - 1;
                for (int st = end; st >= nb; st -= 32) {
                    const int j = st - lane;
                    const bool nz = j >= nb && (hb[j & SM] != 0 || eb[j & SM] != 0);
                    const unsigned bal = __ballot_sync(kFull, nz);
                    if (bal) { nh = st - (__ffs(bal) - 1); break; }
                }
            }
            beg = nb;
            end = nh + 2 < qlen ? nh + 2 : qlen;
        }
    }
    __syncwarp();

    // ---- results
    int score = 0, ti = -1, tk = -1;
    if (!EXT) {
        score = (int)hb[qlen & SM] - (qlen - 1) * e_ins;     // eh[qlen].h (src/ksw.c:634), out of the hat domain
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

template <int NP, int KIND>
__global__ void __launch_bounds__(256, (NP == 4 && KIND == kKindGlobal) ? 3 : LB2_FILL16_MIN_BLOCKS)   // measured per kernel
fill16_kernel(const DTask* __restrict__ tasks, const int32_t* __restrict__ order, int n,
              const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac, uint8_t* __restrict__ zbase,
              DResult* __restrict__ results, const uint2* __restrict__ gmat,
              unsigned int* __restrict__ counter, int S, uint8_t* __restrict__ /*gwin: unused*/)
{
    constexpr int G = 2 * NP;
    __shared__ uint2 smat[kMaxMats * 8];
    __shared__ __align__(16) uint32_t mtab[mask_table_words<NP>()];
    extern __shared__ __align__(16) uint8_t dyn[];
    for (int k = threadIdx.x; k < kMaxMats * 8; k += blockDim.x) smat[k] = gmat[k];
    for (int k = threadIdx.x; k < mask_table_words<NP>(); k += blockDim.x) {
        const int p = k % NP, lohi = k / NP, lo = lohi / (G + 1), hi = lohi % (G + 1);
        const int c0 = 2 * p, c1 = 2 * p + 1;
        mtab[k] = ((c0 >= lo && c0 < hi) ? 0x0000ffffu : 0u) | ((c1 >= lo && c1 < hi) ? 0xffff0000u : 0u);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint8_t* mine = dyn + (size_t)wid * warp_smem_bytes16(S);
    int16_t* hb = reinterpret_cast<int16_t*>(mine);
    int16_t* eb = hb + S;
    uint16_t* qb = reinterpret_cast<uint16_t*>(eb + S);
    uint2* mrw = reinterpret_cast<uint2*>(qb + (S >> 1));
    for (;;) {
        unsigned int t = 0;
        if (lane == 0) t = atomicAdd(counter, 1u);
        t = __shfl_sync(kFull, t, 0);
        if (t >= (unsigned)n) break;
        const int idx = order[t];
        fill_task16<NP, KIND>(tasks[idx], pool, pac, zbase, results + idx, smat, mtab, hb, eb, qb, mrw, S, lane);
    }
}

}  // namespace lb2
