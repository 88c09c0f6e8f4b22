// dp_device.cuh -- device-side data layout shared by the fill and traceback
// kernels of the banded affine-gap DP (sm_100a).
//
// Reference semantics being reproduced (paths relative to the reference tree):
//   ksw_global2      src/ksw.c:543-653
//   ksw_extend_core  src/ksw.c:667-807   (ksw_extend2 :387-490 is the same fill
//                                          without direction bits)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lb2 {

constexpr int kNegInf = -0x40000000;     // src/ksw.c:504
constexpr int kMaxMats = 16;             // distinct scoring matrices per batch
constexpr int kKindGlobal = 0;
constexpr int kKindExtend = 1;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kWantDir = 1, kTargetPac = 2, kTargetRev = 4, kRawOff = 8;    // bits of DTask::want_dir

// One DP task as the kernels see it.  Sequences live in one pooled byte array
// (`pool`), every sequence starting on a 32-byte boundary and padded so that a
// whole column chunk can always be fetched with one aligned vector load.
struct __align__(16) DTask {
    uint32_t q_off32, t_off32;   // offsets into pool, in 32-byte units -- in BYTES with kRawOff (pooled batches: the caller's bytes as they lie, any
                                 // alignment, no padding: what follows a sequence is readable but arbitrary) (t_off32: pac coordinate when kTargetPac)
    int32_t qlen, tlen;
    int32_t w;                   // FINAL band: after src/ksw.c:549 resp. :696-704
    int32_t h0;
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t end_bonus, zdrop;
    uint8_t kind, want_dir /* bit0 directions, kTargetPac, kTargetRev, kRawOff */, mat_id, cshift;   // cshift = log2(G), G = columns per lane per tile
    int32_t row_chunks;          // tiles (32*G columns) of direction nibbles stored per row
    uint64_t z_off;              // byte offset of this task's direction scratch
    uint64_t ctmp_end;           // word offset one past this task's CIGAR scratch
    int32_t ctmp_cap;            // words available below ctmp_end
    int32_t dir_fmt;             // 0: canonical nibble (dp_fill.cuh), 1: raw predicates (dp_fill16.cuh)
};
static_assert(sizeof(DTask) == 80, "DTask layout");

struct __align__(16) DResult {
    int32_t score;               // global: eh[qlen].h ; extend: max
    int32_t max_i, max_j, max_ie, gscore, max_off;   // extend bookkeeping
    int32_t ti, tk;              // traceback start cell
    int32_t n_cigar, rows;       // rows = rows entered before the loop ended
    int64_t cigar_off;
    int64_t cells;
};
static_assert(sizeof(DResult) == 64, "DResult layout");

// Target code of row `idx`: from the pooled bytes, or from the resident 2-bit reference
// (bntseq layout: base k = pac[k>>2] >> ((~k&3)<<1) & 3, reference src/bntseq.c:242).
struct TargetSrc {
    const uint8_t* bytes;      // pooled codes (padded to a multiple of 32)
    const uint8_t* pac;        // resident reference, or nullptr
    uint32_t coor;             // pac coordinate of the window's first base
    int tlen, tpad, rev;
    __device__ __forceinline__ uint32_t at(int idx) const {
        if (idx >= tpad) return 0u;
        if (!pac) return bytes[idx];
        if (idx >= tlen) return 0u;
        const uint64_t k = (uint64_t)coor + (uint32_t)(rev ? tlen - 1 - idx : idx);
        return (pac[k >> 2] >> ((~k & 3u) << 1)) & 3u;
    }
};
__device__ __forceinline__ TargetSrc make_target(const struct DTask& T, const uint8_t* pool, const uint8_t* pac);

// direction nibble: bits0-1 source of H (0 diagonal, 1 E, 2 F), bit2 E was an
// extension, bit3 F was an extension.  The reference byte (src/ksw.c:556) is
// nib&3 | (nib&4) | (nib&8)<<2.
__host__ __device__ inline uint64_t ext_meta_bytes(int tlen) {
    return ((uint64_t)tlen * 8 + 31) & ~uint64_t(31);      // 32: the direction rows behind it start on a sector boundary
}

// bytes of direction storage per lane per tile (G nibbles; one byte when G==1)
__host__ __device__ constexpr int dir_lane_bytes(int G) { return G >= 2 ? G / 2 : 1; }
// tiles a row can span: columns [beg & ~(G-1), end] with end-beg <= ncol
__host__ __device__ inline int row_tiles_for(long ncol, int G) { return (int)((ncol + G) / (32 * G)) + 1; }
// shared-memory bytes one warp needs for a window of S slots
__host__ __device__ constexpr size_t warp_smem_bytes(int S) { return (size_t)S * 10; }
// ... of the packed-int16 kernels (h16, e16, one selector per pair, + 8 staged matrix rows)
__host__ __device__ constexpr size_t warp_smem_bytes16(int S) { return (size_t)S * 5 + 64; }

// start of the task's query codes; two adjacent codes (alignment-free: pooled batches keep the caller's layout)
__device__ __forceinline__ const uint8_t* query_ptr(const DTask& T, const uint8_t* pool) {
    return pool + (size_t)T.q_off32 * ((T.want_dir & kRawOff) ? 1 : 32);
}
__device__ __forceinline__ uint32_t ld_pair(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

__device__ __forceinline__ TargetSrc make_target(const DTask& T, const uint8_t* pool, const uint8_t* pac) {
    TargetSrc t;
    const bool p = (T.want_dir & kTargetPac) != 0;
    t.bytes = pool + (p ? 0 : (size_t)T.t_off32 * ((T.want_dir & kRawOff) ? 1 : 32));
    t.pac = p ? pac : nullptr;
    t.coor = T.t_off32;
    t.tlen = T.tlen; t.tpad = (T.tlen + 31) & ~31; t.rev = (T.want_dir & kTargetRev) ? 1 : 0;
    return t;
}

}  // namespace lb2
