// dp_fill16d.cuh -- the packed fill of dp_fill16.cuh in SUB-WARP LANE GROUPS with dynamic refill: a warp runs 32/L tasks
// at once (L = 8 or 16 lanes each, tiles of L*2NP columns, collectives group-scoped), so that the per-row scalar work of
// the extension is issued once per bundle and narrow / adaptive bands do not idle most of a 32-lane tile.  Every group
// walks its own task with its own row index and, when that task is over (target end reached, m == 0 or z-drop exit),
// commits it and takes the next task from the launch's atomic counter while the other groups keep going: extensions end
// early at unpredictable rows (z-drop: the cells the reference evaluates are 64 % of the static band area on the C2
// workload), so with a common row index a finished group would idle for most of its bundle (the static-bundle build of
// round 1 was 7 % slower and is gone).
#pragma once
#include "dp_fill16.cuh"

namespace lb2 {

template <int NP, int KIND, int L>
__device__ void fill_bundle16_dyn(const DTask* __restrict__ tasks, const int32_t* __restrict__ order,
                              unsigned int* __restrict__ counter, const int n,
                              const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac,
                              uint8_t* __restrict__ zbase, DResult* __restrict__ results,
                              const uint2* __restrict__ smat, const uint32_t* __restrict__ mtab,
                              uint8_t* __restrict__ warp_smem, const int S, const int lane)
{
    constexpr int G = 2 * NP;
    constexpr int GS = (NP == 2 ? 2 : 3);
    constexpr int LS = (L == 32 ? 5 : L == 16 ? 4 : 3);
    constexpr int QR = (2 * L < 32 ? 2 * L : 32);      // selectors refilled per period (and the period, in rows)
    constexpr bool EXT = (KIND == kKindExtend);
    constexpr int EINIT = EXT ? 0 : kNeg16;
    constexpr int FINIT = EXT ? 0 : kNeg16;
    constexpr int POISON = EXT ? 0 : kNeg16;
    const int g = lane >> LS, gl = lane & (L - 1);
    const unsigned gmask = (L == 32) ? kFull : (((1u << L) - 1u) << (g * L));
    const int SM = S - 1, SMQ = SM >> 1;
    // the value 1, opaque to ptxas (S is a power of two): lets add_flag issue on the FMA pipe
    const uint32_t one = (uint32_t)S >> (31 - __clz(S));

    int16_t* __restrict__ hb = reinterpret_cast<int16_t*>(warp_smem + (size_t)g * warp_smem_bytes16(S));
    int16_t* __restrict__ eb = hb + S;
    uint16_t* __restrict__ qb = reinterpret_cast<uint16_t*>(eb + S);
    uint2* __restrict__ mrw = reinterpret_cast<uint2*>(qb + (S >> 1));
    const uint32_t NEGP = dup2(kNeg16);

    // ---- per-task state of my group (re-loaded whenever the group takes its next task)
    bool have = false, exhausted = false, live = false;
    int idx = 0, i = 0;
    int qlen = 0, tlen = 0, w = 0, h0 = 0, o_del = 0, e_del = 0, o_ins = 0, e_ins = 0, zdrop = 0, end_bonus = 0;
    const uint8_t* __restrict__ qseq = pool;
    TargetSrc tsrc; tsrc.bytes = pool; tsrc.pac = nullptr; tsrc.coor = 0; tsrc.tlen = 0; tsrc.tpad = 0; tsrc.rev = 0;
    bool want = false;
    int2* __restrict__ rowmeta = reinterpret_cast<int2*>(zbase);
    uint8_t* __restrict__ zdir = zbase;
    size_t zrow_bytes = 0;
    int qpad = 0;
    uint32_t N_OE_INS = 0, N_OE_DEL = 0, N_E_INS = 0, N_E_DEL = 0, N_O_INS = 0, TILE_STEP = 0, N_TILE_STEP = 0;
    uint32_t RO1_0[NP], NRO_0[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { RO1_0[p] = 0; NRO_0[p] = 0; }
    int slot_hi = 0, q_hi = 0;
    uint32_t qpre = 0u, tcur = 0u, tnext = 0u;
    int beg = 0, end = 0;
    int mx = 0, mx_i = -1, mx_j = -1, mx_ie = -1, gscore = -1, max_off = 0;
    long long cells = 0;
    int rows = 0;

    for (;;) {
        // ---- a group whose task is over (or that has none yet) commits it and takes the next one from the
        // counter; the other groups of the warp wait here (group-scoped collectives only inside)
        if (!exhausted && !(live && i < tlen)) {
            if (have) {
                int score, ti, tk;
                if (!EXT) {
                    score = (int)hb[qlen & SM] - (qlen - 1) * e_ins;     // out of the hat domain
                    ti = tlen - 1;
                    tk = (ti + w + 1 < qlen ? ti + w + 1 : qlen) - 1;
                } else {
                    score = mx;
                    if (gscore <= 0 || gscore <= mx - end_bonus) { ti = mx_i; tk = mx_j; }
                    else { ti = mx_ie; tk = qlen - 1; }
                }
                if (gl == 0) {
                    DResult r;
                    r.score = score; r.max_i = mx_i; r.max_j = mx_j; r.max_ie = mx_ie;
                    r.gscore = gscore; r.max_off = max_off; r.ti = ti; r.tk = tk;
                    r.n_cigar = 0; r.rows = rows; r.cigar_off = 0; r.cells = cells;
                    results[idx] = r;
                }
                __syncwarp(gmask);                                       // the group is done reading its window
            }
            unsigned t = 0;
            if (gl == 0) t = atomicAdd(counter, 1u);
            t = __shfl_sync(gmask, t, g * L);
            const bool got = t < (unsigned)n;
            if (!got) exhausted = true;
            if (got || !have) {                                          // a group that never had a task shadows task 0
                idx = order[got ? t : 0];
                const DTask T = tasks[idx];
                qlen = T.qlen; tlen = T.tlen; w = T.w; h0 = T.h0;
                o_del = T.o_del; e_del = T.e_del; o_ins = T.o_ins; e_ins = T.e_ins; zdrop = T.zdrop; end_bonus = T.end_bonus;
                qseq = query_ptr(T, pool);
                tsrc = make_target(T, pool, pac);
                want = got && (T.want_dir & kWantDir) != 0;
                rowmeta = reinterpret_cast<int2*>(zbase + T.z_off);
                zdir = zbase + T.z_off + (EXT ? ext_meta_bytes(tlen) : 0) + (size_t)gl * (G / 2);
                zrow_bytes = (size_t)T.row_chunks * 32 * (G / 2);
                const uint2* __restrict__ mrows = smat + (int)T.mat_id * 8;
                qpad = (qlen + 1 + 31) & ~31;
                N_OE_INS = dup2(-(o_ins + e_ins)); N_OE_DEL = dup2(-(o_del + e_del));
                N_E_INS = dup2(-e_ins); N_E_DEL = dup2(-e_del); N_O_INS = dup2(-o_ins);
                TILE_STEP = dup2(L * G * e_ins); N_TILE_STEP = dup2(-L * G * e_ins);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int r = gl * G + 2 * p;
                    RO1_0[p] = pk2((r + 1) * e_ins, (r + 2) * e_ins);
                    NRO_0[p] = pk2(-r * e_ins, -(r + 1) * e_ins);
                }
                stage_matrix<KIND>(mrw, mrows, e_ins, gl);
                // window initialisation: slots [0, send_0]; selectors for columns [0, w+66)
                slot_hi = (w + 1 < qlen) ? w + 1 : qlen;
                for (int j = gl; j <= slot_hi; j += L) {
                    hb[j & SM] = (int16_t)init_h16<KIND>(j, qlen, w, h0, o_ins, e_ins);
                    eb[j & SM] = (int16_t)EINIT;
                }
                q_hi = 0;
                {
                    const int want_q = w + 66 < qpad ? w + 66 : qpad;
                    while (q_hi < want_q) {
                        if (gl < QR / 2) {
                            const uint32_t cc = ld_pair(qseq + q_hi + 2 * gl);
                            qb[((q_hi >> 1) + gl) & SMQ] = (uint16_t)sel_for_pair(cc & 0xffu, cc >> 8);
                        }
                        q_hi += QR;
                    }
                }
                qpre = (q_hi < qpad && gl < QR / 2) ? ld_pair(qseq + q_hi + 2 * gl) : 0u;
                tcur = 0u;
                tnext = tsrc.at(gl);
                beg = 0; end = qlen;
                mx = h0; mx_i = -1; mx_j = -1; mx_ie = -1; gscore = -1; max_off = 0;
                cells = 0; rows = 0; i = 0;
            }
            have = got;
            live = got;
        }
        __syncwarp();
        const bool lr = live && i < tlen;                 // my group computes row i of its task
        if (!__any_sync(kFull, lr)) {
            if (__all_sync(kFull, exhausted)) break;      // every group is out of tasks
            continue;                                     // some group still has to commit / fetch
        }
        if ((i & (L - 1)) == 0) {                         // next L target codes
            tcur = tnext;
            tnext = tsrc.at(i + L + gl);
        }
        if ((i & (QR - 1)) == 0 && i && q_hi < qpad) {    // next QR selectors (pad region is readable)
            if (gl < QR / 2) qb[((q_hi >> 1) + gl) & SMQ] = (uint16_t)sel_for_pair(qpre & 0xffu, qpre >> 8);
            q_hi += QR;
            qpre = (q_hi < qpad && gl < QR / 2) ? ld_pair(qseq + q_hi + 2 * gl) : 0u;
        }
        const int tb = __shfl_sync(kFull, (int)tcur, i & (L - 1), L) & 7;
        const int send = i + w + 1 < qlen ? i + w + 1 : qlen;
        if (lr) {
            if (EXT) {
                if (beg < i - w) beg = i - w;
                if (end > send) end = send;
            } else {
                beg = i > w ? i - w : 0;
                end = send;
            }
        }
        const int base = beg & ~(G - 1);
        if (lr) {
            if (gl == 0 && send > slot_hi) {
                hb[send & SM] = (int16_t)init_h16<KIND>(send, qlen, w, h0, o_ins, e_ins);
                eb[send & SM] = (int16_t)EINIT;
            }
            slot_hi = send > slot_hi ? send : slot_hi;
            if (gl < G - 1 && base + gl < beg) hb[(base + gl) & SM] = (int16_t)POISON;
        }
        __syncwarp();

        const uint2 mrow = mrw[tb];
        int h1init;
        if (EXT) {
            h1init = 0;
            if (beg == 0) { h1init = h0 - (o_del + e_del * (i + 1)); if (h1init < 0) h1init = 0; }
        } else {
            h1init = beg == 0 ? -(o_del + e_del * (i + 1)) - e_ins : kNeg16;    // hat: column -1
        }
        const int rb = beg - base, re = end - base;
        const int ntile = (lr && end >= base) ? ((end - base) >> (LS + GS)) + 1 : 0;
        const int ntile_max = __reduce_max_sync(kFull, ntile);
        uint32_t carryF = dup2(EXT ? FINIT + rb * e_ins : kNeg16);
        uint32_t carryH = 0;
        uint32_t mrowmax[NP];
        int mt_lo[NP], mt_hi[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) { mrowmax[p] = 0; mt_lo[p] = -1; mt_hi[p] = -1; }
        uint8_t* zp = zdir + (size_t)i * zrow_bytes;
        int r0 = gl * G;

        for (int tile = 0; tile < ntile_max; ++tile, r0 += L * G, zp += L * (G / 2)) {
            const bool ta = tile < ntile;                  // my group has this tile
            const int s0 = (base + r0) & SM;
            uint32_t H[NP], E[NP], qs[NP];
            PVec<NP>::ld(hb + s0, H); PVec<NP>::ld(eb + s0, E); PVec<NP>::ldq(qb + (s0 >> 1), qs);
            uint32_t am[NP], cm[NP];
            if (EXT) {
                const int lo = __viaddmin_s32_relu(rb, -r0, G);
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
                pre[p] = __vmaxs2(run, prmt(u, NEGP, 0x1054));
                run = __vimax3_s16x2(run, u, prmt(u, 0u, 0x1032));
            }
            uint32_t incl = run;
#pragma unroll
            for (int d = 1; d < L; d <<= 1) incl = __vmaxs2(incl, __shfl_up_sync(kFull, incl, d, L));
            uint32_t pin = __shfl_up_sync(kFull, incl, 1, L);
            if (gl == 0) pin = NEGP;
            pin = __vmaxs2(pin, carryF);
            // the scan offsets are TILE-relative (column index inside the tile): the carry moves into the next tile's domain,
            // one subtraction per tile instead of two offset updates per register
            carryF = __vmaxs2(carryF, __shfl_sync(kFull, incl, L - 1, L));
            if (EXT) carryF = __vadd2(carryF, N_TILE_STEP);

            uint32_t dirw = 0;
            uint32_t Hn[NP];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t F = EXT ? __vadd2(__vmaxs2(pin, pre[p]), NRO_0[p]) : __vmaxs2(pin, pre[p]);
                bool a_hi, a_lo, b_hi, b_lo, c_hi, c_lo, d_hi, d_lo;
                uint32_t h;
                if (EXT) {
                    h = __vibmax_s16x2(E[p], M[p], &a_hi, &a_lo);
                    h = __vibmax_s16x2(F, h, &b_hi, &b_lo);
                } else {
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
                    const uint32_t hm = h | ~am[p];
                    // no `ta` guard: beyond the group's last tile the mask am is empty, hm reads -1 and never wins
                    mrowmax[p] = __vibmax_s16x2(hm, mrowmax[p], &m_hi, &m_lo);
                    if (m_lo) mt_lo[p] = tile;
                    if (m_hi) mt_hi[p] = tile;
                }
            }
            uint32_t left = __shfl_up_sync(kFull, Hn[NP - 1], 1, L);
            if (gl == 0) left = carryH;
            carryH = __shfl_sync(kFull, Hn[NP - 1], L - 1, L);
#pragma unroll
            for (int p = NP - 1; p >= 0; --p) {
                const uint32_t sh = prmt(p == 0 ? left : Hn[p - 1], Hn[p], 0x5432);
                H[p] = EXT ? blend(sh, H[p], cm[p]) : sh;
            }
            if (ta && r0 <= re) { PVec<NP>::st(hb + s0, H); PVec<NP>::st(eb + s0, E); }
            // every lane of a tile pass stores, also those right of the band: a group's stores then fill whole 32-byte
            // sectors and nothing has to be read back to merge a partly written one when it leaves the L2 (ncu: 1.2 GB of
            // DRAM reads per 200 k-task launch of the dominant kernel came from exactly that); the bytes lie inside the
            // row's allocation and the traceback never looks outside [beg, end)
            if (ta && want) {
                if (NP == 2) *reinterpret_cast<uint16_t*>(zp) = (uint16_t)dirw;
                else *reinterpret_cast<uint32_t*>(zp) = dirw;
            }
        }
        __syncwarp();
        if (lr) {
            if (gl == 0 && end >= beg) {
                hb[beg & SM] = (int16_t)h1init;
                eb[end & SM] = (int16_t)EINIT;
            }
            if (EXT && want && gl == 0) rowmeta[i] = make_int2(beg, end);
            cells += end > beg ? end - beg : 0;
            rows = i + 1;
        }
        if (EXT) {
            __syncwarp();
            uint32_t key = 0;
            const int cbase = base + gl * G + 1;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const uint32_t klo = (mrowmax[p] << 16) | (uint32_t)(mt_lo[p] < 0 ? 0 : cbase + (mt_lo[p] << (LS + GS)) + 2 * p);
                const uint32_t khi = (mrowmax[p] & 0xffff0000u) | (uint32_t)(mt_hi[p] < 0 ? 0 : cbase + (mt_hi[p] << (LS + GS)) + 2 * p + 1);
                key = key > klo ? key : klo;
                key = key > khi ? key : khi;
            }
            const uint32_t gkey = __reduce_max_sync(gmask, key);
            const int gm = (int)(gkey >> 16), gmj = (int)(gkey & 0xffffu) - 1;
            // band trim over the slots as they stand after this row (group-wide ballots)
            int nb = end, nh;
            {
                constexpr unsigned LM = (L == 32) ? 0xffffffffu : ((1u << L) - 1u);
                bool found = false;
                for (int st = beg;; st += L) {                             // first scan, ascending over [beg,end)
                    const bool pending = lr && !found && st < end;
                    if (!__any_sync(kFull, pending)) break;
                    const int j = st + gl;
                    const bool nz = pending && j < end && (hb[j & SM] != 0 || eb[j & SM] != 0);
                    const unsigned bal = (__ballot_sync(kFull, nz) >> (g * L)) & LM;
                    if (pending && bal) { nb = st + __ffs(bal) - 1; found = true; }
                }
                nh = nb - 1;
                found = false;
                for (int st = end;; st -= L) {                             // second scan, descending over [beg',end]
                    const bool pending = lr && !found && st >= nb;
                    if (!__any_sync(kFull, pending)) break;
                    const int j = st - gl;
                    const bool nz = pending && j >= nb && (hb[j & SM] != 0 || eb[j & SM] != 0);
                    const unsigned bal = (__ballot_sync(kFull, nz) >> (g * L)) & LM;
                    if (pending && bal) { nh = st - (__ffs(bal) - 1); found = true; }
                }
            }
            if (lr) {
                const int jfin = beg > end ? beg : end;
                if (jfin == qlen) {                              // src/ksw.c:759-762
                    const int h1 = end > beg ? (int)hb[qlen & SM] : h1init;
                    mx_ie = gscore > h1 ? mx_ie : i;
                    gscore = gscore > h1 ? gscore : h1;
                }
                if (gm == 0) live = false;                       // :763
                else {
                    if (gm > mx) {
                        mx = gm; mx_i = i; mx_j = gmj;
                        int off = gmj - i; off = off < 0 ? -off : off;
                        max_off = max_off > off ? max_off : off;
                    } else if (zdrop > 0) {                    // :767-773
                        const int di = i - mx_i, dj = gmj - mx_j;
                        bool drop;
                        if (di > dj) drop = mx - gm - (di - dj) * e_del > zdrop;
                        else         drop = mx - gm - (dj - di) * e_ins > zdrop;
                        if (drop) live = false;
                    }
                    if (live) {                                  // :775-778
                        beg = nb;
                        end = nh + 2 < qlen ? nh + 2 : qlen;
                    }
                }
            }
        }
        ++i;
    }
}

template <int NP, int KIND, int L>
__global__ void __launch_bounds__(256, 2)
fill16d_kernel(const DTask* __restrict__ tasks, const int32_t* __restrict__ order, int n,
               const uint8_t* __restrict__ pool, const uint8_t* __restrict__ pac, uint8_t* __restrict__ zbase,
               DResult* __restrict__ results, const uint2* __restrict__ gmat,
               unsigned int* __restrict__ counter, int S, uint8_t* __restrict__ /*gwin: unused*/)
{
    constexpr int G = 2 * NP;
    constexpr int NSUB = 32 / L;
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
    uint8_t* mine = dyn + (size_t)wid * NSUB * warp_smem_bytes16(S);
    fill_bundle16_dyn<NP, KIND, L>(tasks, order, counter, n, pool, pac, zbase, results, smat, mtab, mine, S, lane);
}

}  // namespace lb2
