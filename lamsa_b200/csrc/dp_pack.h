// dp_pack.h -- host-side classification and packing of ONE DP task, shared by the two producers of
// batches: lb2_batch_create (dp_batch.cu: an array of lb2_task records) and the batch producer
// (producer.cu: worker threads classify and copy their own requests while they park them, so the
// per-GPU submitter only concatenates).  No DP arithmetic here.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/lamsa_b200.h"
#include "dp_device.cuh"

namespace lb2 {

// A launch class = (kind, variant, S = window slots).
// variants: 0..2 int32 lanes with G = 1,2,4 columns per lane; 3,4 packed int16 with NP = 2,4 pairs
// per lane; 5 = int32 G=4 with the window in global memory (does not fit shared memory);
// 6,7 = packed NP=2 in sub-warp groups of L = 16 / 8 lanes per task (dp_fill16d.cuh); 8, 9 = packed NP=4, L = 8 / 16;
// 10 = extensions of many rows inside a band of at most 15: one task per warp, the window in registers (dp_fill_lean.cuh).
constexpr int kMinLogS = 6, kMaxLogS = 18;          // 64 .. 262144 slots per warp
constexpr int kNumLogS = kMaxLogS - kMinLogS + 1;
constexpr int kNumVar = 11;
constexpr int kVarGmem = 5;
constexpr int kVarBlock = 10;
constexpr int kNumClass = 2 * kNumVar * kNumLogS;
constexpr size_t kMaxDynSmem = 200 * 1024;
constexpr int kCostBins = 48;
inline int class_id(int kind, int var, int logS) { return (kind * kNumVar + var) * kNumLogS + (logS - kMinLogS); }
inline int class_kind(int c) { return c / (kNumVar * kNumLogS); }
inline int class_var(int c) { return (c / kNumLogS) % kNumVar; }
inline int class_logS(int c) { return c % kNumLogS + kMinLogS; }
inline int var_gshift(int var) { return var == kVarBlock ? 0 : var >= 8 ? 3 : var == kVarGmem || var >= 6 ? 2 : var < 3 ? var : var - 1; }     // log2(columns per lane)
inline bool var_packed(int var) { return var == 3 || var == 4 || (var >= 6 && var <= 9); }
inline int var_tasks_per_warp(int var) { return var == 6 || var == 9 ? 2 : (var >= 7 && var <= 8) ? 4 : 1; }
inline size_t var_warp_smem(int var, int S) {
    return var_packed(var) ? warp_smem_bytes16(S) * var_tasks_per_warp(var) : warp_smem_bytes(S);
}

inline int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}

// largest / smallest entry of a scoring matrix (floored / capped at 0 like the reference's scans); tasks of a batch
// share a handful of matrices, so the last one looked at is remembered per thread
struct MatRange { int maxs, mins; };
inline MatRange mat_range(int m, const int8_t* mat) {
    thread_local const int8_t* last = nullptr; thread_local int last_m = 0; thread_local int8_t copy[64]; thread_local MatRange r{0, 0};
    if (mat == last && m == last_m && !memcmp(copy, mat, (size_t)m * m)) return r;
    r.maxs = 0; r.mins = 0;
    for (int a = 0; a < m * m; ++a) { r.maxs = r.maxs > mat[a] ? r.maxs : mat[a]; r.mins = r.mins < mat[a] ? r.mins : mat[a]; }
    last = mat; last_m = m; memcpy(copy, mat, (size_t)m * m);
    return r;
}

// src/ksw.c:696-704 -- double division, truncation toward zero
inline int extend_band(int w, int qlen, int m, const int8_t* mat, int end_bonus,
                       int o_del, int e_del, int o_ins, int e_ins) {
    const int best = mat_range(m, mat).maxs;
    int lim = (int)((double)(qlen * best + end_bonus - o_ins) / e_ins + 1.);
    lim = lim > 1 ? lim : 1;
    w = w < lim ? w : lim;
    lim = (int)((double)(qlen * best + end_bonus - o_del) / e_del + 1.);
    lim = lim > 1 ? lim : 1;
    w = w < lim ? w : lim;
    return w;
}

// Can every value of this task live in the packed-int16 domain of dp_fill16.cuh?
inline bool fits_int16(const lb2_task& t, int w) {
    const MatRange mr = mat_range(t.m, t.mat);
    const int maxs = mr.maxs, mins = mr.mins;
    const long ncol = std::min<long>(t.qlen, 2L * w + 1);
    const long maxo = std::max(t.o_del, t.o_ins), maxe = std::max(t.e_del, t.e_ins);
    if (t.o_del < 0 || t.o_ins < 0 || maxe > 255 || maxo + maxe > 500) return false;
    const long scan = (ncol + 300) * (long)t.e_ins;
    if (t.kind == LB2_KIND_EXTEND) {
        if (maxs > 1) return false;                       // M = min(H+s, 2H) needs s <= H for H >= 1
        const long maxh = (long)t.h0 + (long)t.qlen * maxs;
        return maxh <= 16000 && maxh + scan <= 32000;
    }
    // global fill, hat domain (values carry + column*e_ins): s + e_ins must stay an int8
    if (maxs + t.e_ins > 127) return false;
    const long lower = (long)(-mins) * std::min(t.qlen, t.tlen) + 2 * maxo + maxe * ((long)t.qlen + t.tlen + 2) + (maxo + maxe) + 64;
    return lower <= 30000 && (long)t.qlen * (maxs + t.e_ins) + 64 <= 32000;
}

// kernel variant from the widest band a row can have
inline int pick_variant(const lb2_task& t, int w, long ncol, int logS) {
    const int S_ = 1 << logS;
    // long extensions inside a narrow band: the latency of one row on one warp is what counts (dp_fill_lean.cuh)
    static const int lean_rows = env_int("LB2_LEAN_ROWS", 192);
    if (lean_rows > 0 && t.kind == LB2_KIND_EXTEND && w <= 15 && t.tlen >= lean_rows) return kVarBlock;
    static const int force_gmem = env_int("LB2_FORCE_GMEM", 0);              // test hook
    if (force_gmem || warp_smem_bytes16(S_) > kMaxDynSmem) return kVarGmem;   // window beyond shared memory
    static const int use16 = env_int("LB2_P16", 1), p16_min = env_int("LB2_P16_MIN", 37),
                     np4_min = env_int("LB2_NP4_MIN", 200), np4_min_ext = env_int("LB2_NP4_MIN_EXT", 1000000);
    // narrow bands (the short interval fills of real reads) run 4 tasks per warp; so does every extension
    // whose static band is below 410 columns: its LIVE band (src/ksw.c:775-778) is a few dozen columns wide,
    // and the 32-column tiles of an 8-lane group follow it with less idle lanes than 128-column warp tiles
    // (1 M-task C2: 437 vs 423 GCUPS, tools/kernel_probe.py)
    static const int sub_l = env_int("LB2_SUBWARP", 8), sub_max_ext = env_int("LB2_SUBWARP_MAX_EXT", 410),
                     sub_max_glb = env_int("LB2_SUBWARP_MAX_GLB", 200);
    // long tasks: what matters is the time of ONE row on ONE warp (a batch lasts as long as its longest task), not lanes kept
    // busy -- one task per warp (no group bookkeeping in the row loop) instead of the lane groups
    static const int long_rows = env_int("LB2_LONG_ROWS", 1200), long_var = env_int("LB2_LONG_VAR", 3);
    if (use16 && fits_int16(t, w)) {
        const bool wide = ncol >= (t.kind == LB2_KIND_EXTEND ? np4_min_ext : np4_min);
        if (t.tlen >= long_rows && ncol < 2000 && long_var >= 0) return ncol >= 1000 ? 4 : long_var;
        const int sub_max = t.kind == LB2_KIND_EXTEND ? sub_max_ext : sub_max_glb;
        // wide bands in 8-lane groups with 8 columns per lane (64-column tiles): LB2_SUB_NP4_MIN_EXT / _GLB
        static const int sub4_ext = env_int("LB2_SUB_NP4_MIN_EXT", 160), sub4_glb = env_int("LB2_SUB_NP4_MIN_GLB", 100);
        static const int sub16_ext = env_int("LB2_SUB16_NP4_MIN_EXT", 375);     // 16-lane groups, 128-column tiles: windows of 1024 slots halve the resident 8-lane groups
        static const int sub16_glb = env_int("LB2_SUB16_NP4_MIN_GLB", 1000000), sub16_glb_max = env_int("LB2_SUB16_NP4_MAX_GLB", 1000000);
        if (sub_l == 8 && t.kind == LB2_KIND_EXTEND && ncol < sub_max && ncol >= sub16_ext &&
            warp_smem_bytes16(S_) * 2 * 2 <= kMaxDynSmem) return 9;
        if (sub_l == 8 && t.kind == LB2_KIND_GLOBAL && ncol >= sub16_glb && ncol < sub16_glb_max &&
            warp_smem_bytes16(S_) * 2 * 2 <= kMaxDynSmem) return 9;
        if (sub_l == 8 && ncol < sub_max && ncol >= (t.kind == LB2_KIND_EXTEND ? sub4_ext : sub4_glb) &&
            warp_smem_bytes16(S_) * 4 * 2 <= kMaxDynSmem) return 8;
        if (sub_l && ncol < sub_max && warp_smem_bytes16(S_) * (32 / sub_l) * 2 <= kMaxDynSmem) return sub_l == 16 ? 6 : 7;
        if (ncol >= p16_min) return wide ? 4 : 3;
    }
    if (warp_smem_bytes(S_) > kMaxDynSmem) return kVarGmem;
    return ncol <= 36 ? 0 : ncol <= 72 ? 1 : 2;
}
// window slots: the whole eh[] array when it is small, else band window + look-ahead
inline int pick_logS(int qlen, int w) {
    const long qpad = ((long)qlen + 1 + 31) & ~31L;
    const long need = std::min<long>(qpad, 2L * w + 140);
    int l = kMinLogS;
    while ((1L << l) < need && l <= kMaxLogS) ++l;
    return l <= kMaxLogS ? l : -1;
}

// One DP task, validated, classified and sized; sequence offsets are relative to whatever pool its
// producer copied the sequences into (32-byte units), direction / CIGAR scratch offsets are unset.
struct PackedTask {
    DTask d;
    uint64_t zsz;        // direction scratch bytes (0: no CIGAR wanted)
    int32_t ctmpw;       // CIGAR scratch words
    int16_t cls;         // launch class
    uint8_t flags;       // LB2_FLAG_*
    uint8_t bin;         // descending log-spaced cost bin inside the class (0 = most cells)
};

inline uint64_t pool_bytes_query(int qlen) { return ((uint64_t)qlen + 1 + 31) & ~uint64_t(31); }
inline uint64_t pool_bytes_target(int tlen) { return ((uint64_t)tlen + 31) & ~uint64_t(31); }

// Validates and classifies `t` into `o` (everything but q_off32 / t_off32 for byte targets, mat_id, z_off,
// ctmp_end).  Returns 0, or 1 with a message in err.  `l_pac` < 0: no resident reference.
inline int classify_task(const lb2_task& t, int64_t l_pac, PackedTask& o, char* err, size_t errn) {
    if (t.qlen < 0 || t.tlen < 0) { snprintf(err, errn, "qlen %d tlen %d", t.qlen, t.tlen); return 1; }
    if (t.kind != LB2_KIND_GLOBAL && t.kind != LB2_KIND_EXTEND) { snprintf(err, errn, "kind %d", t.kind); return 1; }
    if (t.m < 1 || t.m > 8 || !t.mat) { snprintf(err, errn, "alphabet size %d unsupported (1..8)", t.m); return 1; }
    const bool tpac = (t.flags & LB2_FLAG_TARGET_PAC) != 0;
    if ((t.qlen && !t.query) || (t.tlen && !tpac && !t.target)) { snprintf(err, errn, "NULL sequence"); return 1; }
    if (tpac && (l_pac < 0 || t.target_pac < 0 || t.target_pac + t.tlen > l_pac)) {
        snprintf(err, errn, "reference window [%lld,+%d) outside the resident reference (%lld bases)",
                 (long long)t.target_pac, t.tlen, (long long)(l_pac < 0 ? 0 : l_pac));
        return 1;
    }
    if (t.e_del <= 0 || t.e_ins <= 0) { snprintf(err, errn, "gap extension penalties must be > 0"); return 1; }
    int w = t.w;
    if (t.kind == LB2_KIND_GLOBAL) {
        const int dl = std::abs(t.qlen - t.tlen);
        w = dl + 3 < w ? w : dl + 3;                                  // src/ksw.c:549
    } else {
        if (t.h0 <= 0) { snprintf(err, errn, "h0 must be > 0 (src/ksw.c:682)"); return 1; }
        w = extend_band(w, t.qlen, t.m, t.mat, t.end_bonus, t.o_del, t.e_del, t.o_ins, t.e_ins);
    }
    if (w < 0) { snprintf(err, errn, "negative band"); return 1; }
    const long ncol = std::min<long>(t.qlen, 2L * w + 1);
    const int ls = pick_logS(t.qlen, w);
    if (ls < 0) { snprintf(err, errn, "qlen %d with band %d needs a window beyond %d slots (not supported yet)", t.qlen, w, 1 << kMaxLogS); return 1; }
    const int var = pick_variant(t, w, ncol, ls);
    const int cs = var_gshift(var);
    DTask& d = o.d;
    memset(&d, 0, sizeof d);
    d.t_off32 = tpac ? (uint32_t)t.target_pac : 0u;
    d.qlen = t.qlen; d.tlen = t.tlen; d.w = w; d.h0 = t.h0;
    d.o_del = t.o_del; d.e_del = t.e_del; d.o_ins = t.o_ins; d.e_ins = t.e_ins;
    d.end_bonus = t.end_bonus; d.zdrop = t.zdrop;
    d.kind = (uint8_t)t.kind;
    d.want_dir = (uint8_t)(((t.flags & LB2_FLAG_CIGAR) ? kWantDir : 0) | (tpac ? kTargetPac : 0) |
                           ((tpac && (t.flags & LB2_FLAG_TARGET_REV)) ? kTargetRev : 0));
    d.cshift = (uint8_t)cs;
    d.row_chunks = row_tiles_for(ncol, 1 << cs);
    d.dir_fmt = var_packed(var) ? 1 : 0;
    o.flags = (uint8_t)t.flags;
    o.cls = (int16_t)class_id(t.kind, var, ls);
    if (t.flags & LB2_FLAG_CIGAR) {
        const int G = 1 << cs;
        uint64_t z = (uint64_t)t.tlen * d.row_chunks * 32 * dir_lane_bytes(G);
        if (t.kind == LB2_KIND_EXTEND) z += ext_meta_bytes(t.tlen);
        o.zsz = (z + 31) & ~uint64_t(31);                 // tasks start on a 32-byte sector boundary
        o.ctmpw = t.qlen + t.tlen + 2;
    } else { o.zsz = 0; o.ctmpw = 0; }
    const int64_t cost = (int64_t)t.tlen * ncol + 1;                 // ~3 bins per octave, heaviest first
    const int l2 = 63 - __builtin_clzll((unsigned long long)cost);
    const int frac = l2 >= 2 ? (int)((cost >> (l2 - 2)) & 3) : 0;
    const int bin = std::min(kCostBins - 1, std::max(0, (l2 * 4 + frac) / 3 - 4));
    o.bin = (uint8_t)(kCostBins - 1 - bin);
    return 0;
}

// Sequence bytes into a 32-byte aligned, 32-byte padded pool slot.  `stream`: with non-temporal stores -- the pinned
// staging of a large batch is written once and read by the DMA engine only, so the destination lines need not be
// read into the cache first (a write-allocate copy moves the destination twice over the memory bus; with eight
// processes packing a gigabyte each per step the host memory system is what bounds the end-to-end rate).
inline void put_sequence(uint8_t* dst, const uint8_t* src, size_t len, size_t padded, bool rev, bool stream) {
    if (rev) { std::reverse_copy(src, src + len, dst); memset(dst + len, 0, padded - len); return; }
#if defined(__SSE2__)
    if (stream && len >= 64) {
        size_t k = 0;
        for (; k + 16 <= len; k += 16)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k), _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + k)));
        uint8_t tail[48];
        memset(tail, 0, sizeof tail);
        memcpy(tail, src + k, len - k);
        for (size_t q = 0; k + q < padded; q += 16)
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst + k + q), _mm_loadu_si128(reinterpret_cast<const __m128i*>(tail + q)));
        return;
    }
#endif
    if (len) memcpy(dst, src, len);
    memset(dst + len, 0, padded - len);
}

// copy a task's sequences into a pool at byte offset `off` (32-byte aligned); returns the bytes used.
// `rev`: store both sequences back to front (what ksw_extend_r does before its fill, src/ksw.c:826-830).
inline uint64_t copy_sequences(const lb2_task& t, PackedTask& o, uint8_t* pool, uint64_t off, bool rev = false, bool stream = false) {
    const bool tpac = (t.flags & LB2_FLAG_TARGET_PAC) != 0;
    const uint64_t qp = pool_bytes_query(t.qlen);
    o.d.q_off32 = (uint32_t)(off >> 5);
    put_sequence(pool + off, t.query, (size_t)t.qlen, (size_t)qp, rev, stream);
    uint64_t used = qp;
    if (!tpac) {
        const uint64_t tp = pool_bytes_target(t.tlen);
        o.d.t_off32 = (uint32_t)((off + qp) >> 5);
        put_sequence(pool + off + qp, t.target, (size_t)t.tlen, (size_t)tp, rev, stream);
        used += tp;
    }
    return used;
}

// Tasks classified and copied by one producer thread, ready to be concatenated into a batch.
struct TaskBlob {
    std::vector<PackedTask> tasks;
    std::vector<uint8_t> pool;                          // sequences, 32-byte granular; offsets in tasks are relative to it
    std::vector<std::array<int8_t, 64>> mats;           // distinct scoring matrices (8x8, zero padded); DTask::mat_id indexes this
    void clear() { tasks.clear(); pool.clear(); mats.clear(); }
    int matrix_id(int m, const int8_t* mat) {
        std::array<int8_t, 64> m8; m8.fill(0);
        for (int a = 0; a < m; ++a) for (int c = 0; c < m; ++c) m8[(size_t)a * 8 + c] = mat[a * m + c];
        for (size_t k = 0; k < mats.size(); ++k) if (mats[k] == m8) return (int)k;
        mats.push_back(m8);
        return (int)mats.size() - 1;
    }
    // classify + copy; returns 0 or 1 (message in err)
    int add(const lb2_task& t, int64_t l_pac, char* err, size_t errn, bool rev = false) {
        PackedTask p;
        if (classify_task(t, l_pac, p, err, errn)) return 1;
        p.d.mat_id = (uint8_t)matrix_id(t.m, t.mat);
        const uint64_t off = pool.size();
        const uint64_t need = pool_bytes_query(t.qlen) + ((t.flags & LB2_FLAG_TARGET_PAC) ? 0 : pool_bytes_target(t.tlen));
        pool.resize(off + need);
        copy_sequences(t, p, pool.data(), off, rev);
        tasks.push_back(p);
        return 0;
    }
};

}  // namespace lb2
