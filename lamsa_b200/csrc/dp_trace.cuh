// dp_trace.cuh -- traceback over the direction nibbles written by the fill
// kernels, one thread per task (latency-bound pointer chase; thousands of
// tasks in flight hide it).  Reproduces src/ksw.c:636-649 (global) and
// :792-802 (extension), including the reference's treatment of cells the
// extension never computed: there the byte is the 255 fill of :707, which
// decodes as "insertion" in every state.
#pragma once
#include "dp_device.cuh"

namespace lb2 {

// tasks whose path is at least this long (qlen + tlen) are walked by a whole warp (trace_long_kernel): a matter of
// LATENCY (a batch lasts as long as its longest task).  For throughput one thread per task is better -- a million
// 500-row walks keep every SM busy at a thirty-second of the issue slots a warp per walk would take; measured on the
// C2 batch with the threshold at 1536: traceback 14.8 -> 25.6 ms -- so the threshold sits above that workload's tasks
// (the same 1200 rows at which the fill goes to one task per warp, dp_pack.h).
constexpr int kLongTrace = 2400;

struct CigarWriter {
    int32_t* top;      // next free word is top[-1]; words are written downwards
    int32_t* floor;    // lowest usable address
    int n;
    int op, len;
    bool overflow;
    bool writer = true;    // warp-uniform walks: every lane keeps the state, one lane stores
    __device__ void flush() {
        if (len > 0) {
            if (top > floor) { --top; if (writer) *top = (int32_t)((uint32_t)len << 4 | (uint32_t)op); ++n; }
            else overflow = true;
        }
    }
    __device__ void push(int o, int l) {      // run-length merge: src/ksw.c:506-516
        if (len > 0 && o == op) len += l;
        else { flush(); op = o; len = l; }
    }
};

__global__ void __launch_bounds__(128)
trace_kernel(const DTask* __restrict__ tasks, const int32_t* __restrict__ order, int n,
             const uint8_t* __restrict__ zbase, DResult* __restrict__ results,
             int32_t* __restrict__ ctmp, int32_t* __restrict__ cdense,
             unsigned long long* __restrict__ cursor, unsigned long long dense_cap,
             int* __restrict__ err, int skip_long)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int idx = order[t];
    const DTask T = tasks[idx];
    if (!(T.want_dir & kWantDir)) return;
    if (skip_long && T.qlen + T.tlen >= kLongTrace) return;       // trace_long_kernel's
    DResult* R = results + idx;
    int i = R->ti, k = R->tk;
    const int w = T.w, G = 1 << T.cshift, gs = T.cshift;
    const bool ext = T.kind == kKindExtend;
    const int2* __restrict__ rowmeta = reinterpret_cast<const int2*>(zbase + T.z_off);
    const uint8_t* __restrict__ zdir = zbase + T.z_off + (ext ? ext_meta_bytes(T.tlen) : 0);
    const int lb = G >= 2 ? G / 2 : 1;                       // bytes per lane per tile
    const size_t zrow_bytes = (size_t)T.row_chunks * 32 * lb;

    CigarWriter W;
    W.top = ctmp + T.ctmp_end; W.floor = W.top - T.ctmp_cap;
    W.n = 0; W.op = -1; W.len = 0; W.overflow = false;

    int which = 0;
    int meta_row = -1; int rb = 0, re = 0;
    while (i >= 0 && k >= 0) {
        const int sbeg = i > w ? i - w : 0;
        bool computed;
        if (ext) {
            if (meta_row != i) { const int2 mm = rowmeta[i]; rb = mm.x; re = mm.y; meta_row = i; }
            computed = k >= rb && k < re;
        } else {
            const int send = i + w + 1 < T.qlen ? i + w + 1 : T.qlen;
            computed = k >= sbeg && k < send;
        }
        if (computed) {
            // fill layout: row i stores, from column base = beg_i & ~(G-1), G nibbles per lane
            const int base = (ext ? rb : sbeg) & ~(G - 1);
            const int rel = k - base;
            const uint8_t* p = zdir + (size_t)i * zrow_bytes + (size_t)(rel >> gs) * lb;
            uint32_t v = (G == 8) ? *reinterpret_cast<const uint32_t*>(p)
                       : (G == 4) ? *reinterpret_cast<const uint16_t*>(p) : *p;
            uint32_t nib = (v >> (4 * (rel & (G - 1)))) & 15u;
            if (T.dir_fmt == 1) {
                // raw predicates of dp_fill16.cuh -> canonical {source of H, E extended, F extended}
                const uint32_t b0 = nib & 1u, b1 = (nib >> 1) & 1u;
                const uint32_t pe = ext ? b0 : b0 ^ 1u, pf = ext ? b1 : b1 ^ 1u;
                nib = (pf ? 2u : pe) | ((~nib) & 12u);
            }
            // byte the reference would hold: f<<4 | e<<2 | h  (src/ksw.c:556)
            if (which == 0) which = nib & 3;
            else if (which == 1) which = (nib >> 2) & 1;
            else if (which == 2) which = (nib & 8) ? 2 : 0;
            else which = 0;                    // bits 6-7 of a computed byte are 0
        } else which = 3;                      // 255 fill
        if (which == 0) { W.push(0, 1); --i; --k; }
        else if (which == 1) { W.push(2, 1); --i; }
        else { W.push(1, 1); --k; }
    }
    if (i >= 0) W.push(2, i + 1);
    if (k >= 0) W.push(1, k + 1);
    W.flush();
    if (W.overflow) { atomicExch(err, 1); R->n_cigar = 0; return; }
    // words [top, top+n) are already in forward (reversed-traversal) order
    const unsigned long long off = atomicAdd(cursor, (unsigned long long)W.n);
    if (off + W.n > dense_cap) { atomicExch(err, 2); R->n_cigar = 0; return; }
    for (int a = 0; a < W.n; ++a) cdense[off + a] = W.top[a];
    R->n_cigar = W.n;
    R->cigar_off = (long long)off;
}

// The same walk for LONG tasks, one warp per task.  A single thread pays one dependent L2/HBM access per row (the row's
// band, then the cell's nibble: about 0.6 us per row, 3 ms for a 5 000-row extension, and a batch lasts as long as its
// longest task).  Here the warp stages sixteen rows at a time -- lane pair (2l, 2l+1) loads row i-l's band and the two
// 16-byte pieces of its direction nibbles that end at the path's current column -- into shared memory and then walks
// them there, every lane in step (lane 0 stores the CIGAR words).  A path that leaves the staged window (a long
// insertion run) simply stages again from where it stands.
constexpr int kTraceRows = 16;
constexpr int kTraceWarps = 4;

__global__ void __launch_bounds__(kTraceWarps * 32)
trace_long_kernel(const DTask* __restrict__ tasks, const int32_t* __restrict__ list, int nlist,
                  const uint8_t* __restrict__ zbase, DResult* __restrict__ results,
                  int32_t* __restrict__ ctmp, int32_t* __restrict__ cdense,
                  unsigned long long* __restrict__ cursor, unsigned long long dense_cap,
                  int* __restrict__ err)
{
    __shared__ __align__(16) uint4 sdir[kTraceWarps][kTraceRows][2];
    __shared__ int4 smeta[kTraceWarps][kTraceRows];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t = blockIdx.x * kTraceWarps + wid;
    if (t >= nlist) return;
    const int idx = list[t];
    const DTask T = tasks[idx];
    DResult* R = results + idx;
    int i = R->ti, k = R->tk;
    const int w = T.w, G = 1 << T.cshift;
    const bool ext = T.kind == kKindExtend;
    const int2* __restrict__ rowmeta = reinterpret_cast<const int2*>(zbase + T.z_off);
    const uint8_t* __restrict__ zdir = zbase + T.z_off + (ext ? ext_meta_bytes(T.tlen) : 0);
    const int lb = G >= 2 ? G / 2 : 1;
    const size_t zrow_bytes = (size_t)T.row_chunks * 32 * lb;
    const int cs = G >= 2 ? 5 : 4;                 // log2(cells per 16-byte piece): a nibble per cell, a byte when G == 1

    CigarWriter W;
    W.top = ctmp + T.ctmp_end; W.floor = W.top - T.ctmp_cap;
    W.n = 0; W.op = -1; W.len = 0; W.overflow = false; W.writer = lane == 0;

    int which = 0;
    while (i >= 0 && k >= 0) {
        {   // ---- stage rows i .. i-15
            const int l = lane >> 1, h = lane & 1;
            const int r = i - l;
            int rb = 0, re = 0, base = 0, c = 0;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (r >= 0) {
                if (ext) { const int2 mm = rowmeta[r]; rb = mm.x; re = mm.y; }
                else { rb = r > w ? r - w : 0; re = r + w + 1 < T.qlen ? r + w + 1 : T.qlen; }
                base = rb & ~(G - 1);
                const int relk = k - base;             // the path's column in this row is never right of k
                c = relk >= 0 ? relk >> cs : 0;
                const int piece = c - 1 + h;
                if (piece >= 0 && (size_t)piece * 16 < zrow_bytes)
                    v = *reinterpret_cast<const uint4*>(zdir + (size_t)r * zrow_bytes + (size_t)piece * 16);
            }
            sdir[wid][l][h] = v;
            if (h == 0) smeta[wid][l] = make_int4(rb, re, base, c);
        }
        __syncwarp();
        const int i0 = i;
        if (which == 0) {
            // runs of diagonal steps -- most of an alignment -- are taken at once: lane l looks at cell (i-l, k-l) of the
            // staged rows, the run is the number of leading lanes whose cell was computed and came from the diagonal
            const int l = lane, rr = i - l, kk = k - l;
            bool is_m = false;
            if (l < kTraceRows && rr >= 0 && kk >= 0) {
                const int4 mt = smeta[wid][l];
                const int rel = kk - mt.z, piece = (rel >> cs) - (mt.w - 1);
                if (kk >= mt.x && kk < mt.y && (piece == 0 || piece == 1)) {
                    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(&sdir[wid][l][0]);
                    uint32_t nib;
                    if (G == 1) nib = bytes[piece * 16 + (rel & 15)] & 15u;
                    else nib = (bytes[piece * 16 + ((rel & 31) >> 1)] >> (4 * (rel & 1))) & 15u;
                    if (T.dir_fmt == 1) {           // raw predicates: H came from the diagonal iff neither E nor F won
                        const uint32_t b0 = nib & 1u, b1 = (nib >> 1) & 1u;
                        is_m = ext ? (b0 == 0u && b1 == 0u) : (b0 == 1u && b1 == 1u);
                    } else is_m = (nib & 3u) == 0u;
                }
            }
            const unsigned bal = __ballot_sync(kFull, is_m);
            const int run = __ffs(~bal) - 1;        // leading ones (at most kTraceRows: lanes beyond never vote)
            if (run > 0) { W.push(0, run); i -= run; k -= run; }
        }
        while (i >= 0 && k >= 0 && i > i0 - kTraceRows) {
            const int4 mt = smeta[wid][i0 - i];
            if (k >= mt.x && k < mt.y) {
                const int rel = k - mt.z;
                const int piece = (rel >> cs) - (mt.w - 1);
                if (piece < 0 || piece > 1) break;     // left of the staged window: stage again from (i, k)
                const uint8_t* bytes = reinterpret_cast<const uint8_t*>(&sdir[wid][i0 - i][0]);
                uint32_t nib;
                if (G == 1) nib = bytes[piece * 16 + (rel & 15)] & 15u;
                else nib = (bytes[piece * 16 + ((rel & 31) >> 1)] >> (4 * (rel & 1))) & 15u;
                if (T.dir_fmt == 1) {
                    const uint32_t b0 = nib & 1u, b1 = (nib >> 1) & 1u;
                    const uint32_t pe = ext ? b0 : b0 ^ 1u, pf = ext ? b1 : b1 ^ 1u;
                    nib = (pf ? 2u : pe) | ((~nib) & 12u);
                }
                if (which == 0) which = nib & 3;
                else if (which == 1) which = (nib >> 2) & 1;
                else if (which == 2) which = (nib & 8) ? 2 : 0;
                else which = 0;
            } else which = 3;
            if (which == 0) { W.push(0, 1); --i; --k; }
            else if (which == 1) { W.push(2, 1); --i; }
            else { W.push(1, 1); --k; }
        }
        __syncwarp();
    }
    if (i >= 0) W.push(2, i + 1);
    if (k >= 0) W.push(1, k + 1);
    W.flush();
    if (W.overflow) { if (lane == 0) { atomicExch(err, 1); R->n_cigar = 0; } return; }
    unsigned long long off = 0;
    if (lane == 0) off = atomicAdd(cursor, (unsigned long long)W.n);
    off = __shfl_sync(kFull, off, 0);
    if (off + W.n > dense_cap) { if (lane == 0) { atomicExch(err, 2); R->n_cigar = 0; } return; }
    __syncwarp();
    for (int a = lane; a < W.n; a += 32) cdense[off + a] = W.top[a];
    if (lane == 0) { R->n_cigar = W.n; R->cigar_off = (long long)off; }
}

}  // namespace lb2
