// aux_scan.cuh -- the reference-touching half of lamsa_res_aux (src/frag_check.c:793-853): walk an alignment
// record's CIGAR against the read and the RESIDENT 2-bit reference and count matches / mismatches / gap opens and
// extensions, from which the host derives NM and AS exactly as the reference does (:835-838).  One warp per
// record: the CIGAR operations are walked by every lane (uniform), the bases of a match run are compared 32 at a
// time across the lanes.  HBM-bound byte work: per record its CIGAR words, read_len bytes of read and ref_len/4
// bytes of packed reference in, 32 bytes out.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/lamsa_b200.h"

namespace lb2 {

struct AuxRec {
    uint32_t cigar_off, n_cigar;     // words inside the batch's CIGAR pool
    uint32_t read_off, read_len;     // bytes inside the batch's read pool
    uint64_t ref_pac;                // forward pac coordinate of the record's first reference base
};

__global__ void __launch_bounds__(128)
aux_scan_kernel(const AuxRec* __restrict__ recs, int n, const int32_t* __restrict__ cigars, const uint8_t* __restrict__ reads,
                const uint8_t* __restrict__ pac, long long l_pac, lb2_aux_result* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= n) return;
    const AuxRec R = recs[r];
    const int32_t* __restrict__ cg = cigars + R.cigar_off;
    const uint8_t* __restrict__ rd = reads + R.read_off;
    int read_i = 0, ref_i = 0, n_io = 0, n_ie = 0, n_do = 0, n_de = 0, m_len = 0, bad = 0;
    int mm = 0;                                              // this lane's share of the mismatches
    for (uint32_t c = 0; c < R.n_cigar; ++c) {
        const int op = cg[c] & 0xf, len = (int)((uint32_t)cg[c] >> 4);
        if (op == LB2_CMATCH) {
            for (int t = lane; t < len; t += 32) {
                const long long k = (long long)R.ref_pac + ref_i + t;
                // base k of the forward reference: pac[k>>2] >> ((~k&3)<<1) & 3 (src/bntseq.c:242); past the end: never equal
                const unsigned rf = k < l_pac ? (pac[k >> 2] >> ((~k & 3) << 1)) & 3u : 255u;
                const unsigned q = read_i + t < (int)R.read_len ? rd[read_i + t] : 254u;
                mm += q != rf;
            }
            read_i += len; ref_i += len; m_len += len;
        } else if (op == LB2_CINS) { read_i += len; n_ie += len; ++n_io; }
        else if (op == LB2_CDEL) { ref_i += len; n_de += len; ++n_do; }
        else if (op == LB2_CSOFT_CLIP) { read_i += len; }
        else { bad = op + 1; break; }                        // the reference prints and exits (:827-829)
    }
    mm = __reduce_add_sync(0xffffffffu, mm);
    if (lane == 0) {
        lb2_aux_result o;
        o.n_match = m_len - mm; o.n_mismatch = mm;
        o.n_ins_open = n_io; o.n_ins_ext = n_ie; o.n_del_open = n_do; o.n_del_ext = n_de;
        o.read_used = read_i; o.ref_used = bad ? -bad : ref_i;
        out[r] = o;
    }
}

}  // namespace lb2
