// dp_trace.cuh -- traceback over the direction nibbles written by the fill
// kernels, one thread per task (latency-bound pointer chase; thousands of
// tasks in flight hide it).  Reproduces src/ksw.c:636-649 (global) and
// :792-802 (extension), including the reference's treatment of cells the
// extension never computed: there the byte is the 255 fill of :707, which
// decodes as "insertion" in every state.
#pragma once
#include "dp_device.cuh"

namespace lb2 {

struct CigarWriter {
    int32_t* top;      // next free word is top[-1]; words are written downwards
    int32_t* floor;    // lowest usable address
    int n;
    int op, len;
    bool overflow;
    __device__ void flush() {
        if (len > 0) {
            if (top > floor) { *--top = (int32_t)((uint32_t)len << 4 | (uint32_t)op); ++n; }
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
             int* __restrict__ err)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int idx = order[t];
    const DTask T = tasks[idx];
    if (!(T.want_dir & kWantDir)) return;
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

}  // namespace lb2
