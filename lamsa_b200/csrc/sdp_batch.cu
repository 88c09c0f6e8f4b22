// sdp_batch.cu -- host side of the sparse-DP chaining interface (include/lamsa_b200.h, section 3):
// lays the reads out for the kernel (per-read hit prefix, seed-of-hit, scan order, scratch),
// derives the unaligned read regions for stage 2 (the reference does that on the host too:
// get_remain_reg, src/lamsa_aln.c:548-572), launches sdp_kernel and gathers the skeleton streams.
// No chaining decision is taken here.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../include/lamsa_b200.h"
#include "ctx_internal.h"
#include "sdp_kernel.cuh"
#include <chrono>

using namespace lb2;
using namespace lb2::sdp;

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    return set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } while (0)

namespace {
template <class T> struct DevBuf {
    T* p = nullptr; size_t cap = 0;
    // geometric growth with a floor: cudaMalloc / cudaFree synchronise the device, and a producer that
    // reloads one batch object again and again must stop allocating after warm-up
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        const size_t want = std::max(std::max(n + n / 4, cap * 2), (size_t)4096 / sizeof(T) + 1);
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p && !view) cudaFree(p); p = nullptr; cap = 0; }
    bool view = false;                         // p points into another buffer (the upload arena)
    void point_at(void* q) { view = true; p = static_cast<T*>(q); cap = 0; }
};
// page-locked host staging, grow-only.  Every copy of a chaining batch goes through one: copies from or to pageable
// memory are staged by the driver in pieces, each piece waited for -- with the producer's sleeping waits
// (cudaDeviceScheduleBlockingSync) that made the H2D of a batch of a few hundred kilobytes cost milliseconds, ten times
// its kernels (LB2_FIBER_STATS: 4.6 s of loading and 6.3 s of stages around 0.3 s of kernels on 20 000 reads).
struct PinBuf {
    uint8_t* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        const size_t want = std::max(std::max(n + n / 4, cap * 2), (size_t)1 << 16);
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
size_t up64(size_t x) { return (x + 63) & ~(size_t)63; }
// Wait for everything queued on `st`: poll an event for a few hundred microseconds (the chaining kernels of a batch
// last that long; a sleeping wait costs a wake-up on top), then sleep.
cudaError_t wait_stream(cudaStream_t st, cudaEvent_t ev) {
    cudaError_t e = cudaEventRecord(ev, st);
    if (e != cudaSuccess) return e;
    const auto t0 = std::chrono::steady_clock::now();
    for (int spin = 0;; ++spin) {
        e = cudaEventQuery(ev);
        if (e != cudaErrorNotReady) return e;
        if ((spin & 63) == 63 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(400)) break;
    }
    return cudaEventSynchronize(ev);
}
}  // namespace

struct lb2_sdp_batch {
    lb2_ctx* ctx = nullptr;
    lb2_sdp_para para{};
    int64_t n = 0, n_hits = 0, n_seeds = 0;
    DRead* reads = nullptr;                   // in `up` (page-locked)
    PinBuf up, up2, down, flags_pin;          // uploads of a reset / of a remain stage, read-backs of a stage, tracked flags
    cudaEvent_t ev_wait = nullptr;
    DevBuf<uint8_t> d_up;                     // device copy of the page-locked arena `up`: ONE copy per load; the next eight point into it
    DevBuf<DRead> d_reads; DevBuf<int32_t> d_order, d_seed_id, d_map_n, d_hoff, d_hseed, d_rflat;
    DevBuf<lb2_sdp_hit> d_hits;
    DevBuf<int> d_scratch, d_dense; DevBuf<DOut> d_outs; DevBuf<unsigned int> d_counter; DevBuf<long long> d_off;
    DevBuf<DRegion> d_regions; DevBuf<DPoint> d_pts; DevBuf<uint8_t> d_flags;
    const int32_t* stream = nullptr; std::vector<int64_t> off; const DOut* outs = nullptr;     // stream / outs: in `down`
    int64_t pairs = 0, h2d = 0, d2h = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

extern "C" void lb2_sdp_destroy(lb2_sdp_batch* b) {
    if (!b) return;
    cudaSetDevice(ctx_device(b->ctx));
    b->d_up.release(); b->d_reads.release(); b->d_order.release(); b->d_seed_id.release(); b->d_map_n.release(); b->d_hoff.release();
    b->d_hseed.release(); b->d_rflat.release(); b->d_hits.release(); b->d_scratch.release(); b->d_dense.release();
    b->d_outs.release(); b->d_counter.release(); b->d_off.release(); b->d_regions.release(); b->d_pts.release(); b->d_flags.release();
    b->up.release(); b->up2.release(); b->down.release(); b->flags_pin.release();
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->ev_wait) cudaEventDestroy(b->ev_wait);
    delete b;
}

extern "C" int lb2_sdp_reset(lb2_sdp_batch* b, const lb2_sdp_para* para, int64_t n_reads, const lb2_sdp_read* reads,
                             const int32_t* seed_id, const int32_t* map_n, const lb2_sdp_hit* hits);

extern "C" int lb2_sdp_create(lb2_ctx* ctx, const lb2_sdp_para* para, int64_t n_reads, const lb2_sdp_read* reads,
                              const int32_t* seed_id, const int32_t* map_n, const lb2_sdp_hit* hits,
                              lb2_sdp_batch** out) {
    if (!ctx || !out) return set_error("lb2_sdp_create: bad arguments");
    CU(cudaSetDevice(ctx_device(ctx)));
    lb2_sdp_batch* b = new lb2_sdp_batch();
    b->ctx = ctx;
    cudaError_t e = cudaEventCreate(&b->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&b->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_wait, cudaEventDisableTiming);
    if (e != cudaSuccess) { lb2_sdp_destroy(b); return set_error("lb2_sdp_create: %s", cudaGetErrorString(e)); }
    if (lb2_sdp_reset(b, para, n_reads, reads, seed_id, map_n, hits)) { lb2_sdp_destroy(b); return 1; }
    *out = b;
    return 0;
}

// (re)load a batch object with a new set of reads; device buffers are grow-only, so a producer
// that keeps one object per worker (the drop-in entry points do) stops allocating after warm-up
extern "C" int lb2_sdp_reset(lb2_sdp_batch* b, const lb2_sdp_para* para, int64_t n_reads, const lb2_sdp_read* reads,
                             const int32_t* seed_id, const int32_t* map_n, const lb2_sdp_hit* hits) {
    if (!b || !para || n_reads < 0 || (n_reads > 0 && (!reads || !seed_id || !map_n)))
        return set_error("lb2_sdp_reset: bad arguments");
    if (para->seed_step <= 0 || para->ske_max < 1 || para->per_aln_m < 1)
        return set_error("lb2_sdp_reset: seed_step, ske_max and per_aln_m must be positive");
    if (n_reads > INT32_MAX) return set_error("lb2_sdp_reset: too many reads");
    lb2_ctx* ctx = b->ctx;
    CU(cudaSetDevice(ctx_device(ctx)));
    b->para = *para; b->n = n_reads;
    b->h2d = b->d2h = b->pairs = 0;
    // pass 1: validate, count
    int64_t n_seeds = 0, n_hits = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        const lb2_sdp_read& rd = reads[r];
        if (rd.seed_out < 0 || rd.seed_first < 0 || rd.hit_first < 0) return set_error("lb2_sdp_reset: read %lld: negative field", (long long)r);
        int64_t H = 0;
        for (int i = 0; i < rd.seed_out; ++i) {
            const int m = map_n[rd.seed_first + i];
            if (m < 0 || m > para->per_aln_m) return set_error("lb2_sdp_reset: read %lld seed %d: map_n %d outside [0, per_aln_m]", (long long)r, i, m);
            if (i > 0 && seed_id[rd.seed_first + i] <= seed_id[rd.seed_first + i - 1]) return set_error("lb2_sdp_reset: read %lld: seed ids must increase", (long long)r);
            H += m;
        }
        if (H > (1 << 26)) return set_error("lb2_sdp_reset: read %lld has %lld hits", (long long)r, (long long)H);
        n_seeds += rd.seed_out; n_hits += H;
    }
    b->n_seeds = n_seeds; b->n_hits = n_hits;
    // the page-locked arena of this load: read descriptors, launch order, seeds / hits of the reads in batch order and
    // the derived index arrays.  It stays untouched until the next reset, so nothing below waits for the copies.
    const size_t o_reads = 0, o_order = o_reads + up64(sizeof(DRead) * (size_t)n_reads), o_sid = o_order + up64(4 * (size_t)n_reads),
                 o_mn = o_sid + up64(4 * (size_t)n_seeds), o_hoff = o_mn + up64(4 * (size_t)n_seeds),
                 o_hseed = o_hoff + up64(4 * (size_t)(n_seeds + n_reads)), o_rflat = o_hseed + up64(4 * (size_t)n_hits),
                 o_hits = o_rflat + up64(4 * (size_t)n_hits), o_end = o_hits + up64(sizeof(lb2_sdp_hit) * (size_t)n_hits);
    {
        cudaError_t e0 = b->up.reserve(o_end + 64);
        if (e0 != cudaSuccess) return set_error("lb2_sdp_reset: pinned staging: %s", cudaGetErrorString(e0));
    }
    b->reads = reinterpret_cast<DRead*>(b->up.p + o_reads);
    int32_t* h_order = reinterpret_cast<int32_t*>(b->up.p + o_order);
    int32_t* h_sid = reinterpret_cast<int32_t*>(b->up.p + o_sid); int32_t* h_mn = reinterpret_cast<int32_t*>(b->up.p + o_mn);
    int32_t* h_hoff = reinterpret_cast<int32_t*>(b->up.p + o_hoff); int32_t* h_hseed = reinterpret_cast<int32_t*>(b->up.p + o_hseed);
    int32_t* h_rflat = reinterpret_cast<int32_t*>(b->up.p + o_rflat);
    lb2_sdp_hit* h_hits = reinterpret_cast<lb2_sdp_hit*>(b->up.p + o_hits);
    int64_t at_seed = 0, at_hit = 0, scratch = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        const lb2_sdp_read& rd = reads[r];
        DRead& d = b->reads[r];
        d.seed_out = rd.seed_out; d.seed_all = rd.seed_all; d.read_len = rd.read_len;
        d.n_region = 0; d.pad = 0; d.region_first = 0;
        d.seed_base = at_seed; d.hoff_base = at_seed + r; d.hit_base = at_hit; d.scratch = scratch;
        int32_t* hoff = h_hoff + d.hoff_base;
        int acc = 0;
        for (int i = 0; i < rd.seed_out; ++i) {
            h_sid[d.seed_base + i] = seed_id[rd.seed_first + i];
            const int m = map_n[rd.seed_first + i];
            h_mn[d.seed_base + i] = m;
            hoff[i] = acc;
            for (int j = 0; j < m; ++j) h_hseed[d.hit_base + acc + j] = i;
            acc += m;
        }
        hoff[rd.seed_out] = acc;
        d.n_hits = acc;
        if (acc) memcpy(h_hits + d.hit_base, hits + rd.hit_first, (size_t)acc * sizeof(lb2_sdp_hit));
        // scan order of the predecessor loops (src/lamsa_dp_con.c:713-714): seeds descending, hits ascending
        int q = 0;
        for (int i = rd.seed_out - 1; i >= 0; --i)
            for (int p = hoff[i]; p < hoff[i + 1]; ++p) h_rflat[d.hit_base + q++] = p;
        at_seed += rd.seed_out; at_hit += acc;
        scratch += make_layout(acc, rd.seed_out, para->ske_max).total;
    }
    std::iota(h_order, h_order + n_reads, 0);
    std::stable_sort(h_order, h_order + n_reads, [&](int32_t x, int32_t y) { return b->reads[x].n_hits > b->reads[y].n_hits; });

    cudaStream_t st = ctx_stream(ctx);
    cudaError_t e = b->d_up.reserve(o_end + 64);
    if (e != cudaSuccess) return set_error("lb2_sdp_reset: cudaMalloc: %s", cudaGetErrorString(e));
    b->d_reads.point_at(b->d_up.p + o_reads); b->d_order.point_at(b->d_up.p + o_order); b->d_seed_id.point_at(b->d_up.p + o_sid);
    b->d_map_n.point_at(b->d_up.p + o_mn); b->d_hoff.point_at(b->d_up.p + o_hoff); b->d_hseed.point_at(b->d_up.p + o_hseed);
    b->d_rflat.point_at(b->d_up.p + o_rflat); b->d_hits.point_at(b->d_up.p + o_hits);
    if (o_end) {
        e = cudaMemcpyAsync(b->d_up.p, b->up.p, o_end, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return set_error("lb2_sdp_reset: H2D: %s", cudaGetErrorString(e));
        b->h2d += (int64_t)o_end;
    }
    e = b->d_scratch.reserve((size_t)scratch);
    if (e == cudaSuccess) e = b->d_outs.reserve(n_reads);
    if (e == cudaSuccess) e = b->d_counter.reserve(1);
    if (e == cudaSuccess) e = b->d_off.reserve(n_reads + 1);
    if (e != cudaSuccess) return set_error("lb2_sdp_reset: %s", cudaGetErrorString(e));
    return 0;
}

static const char* err_name(int e) {
    switch (e) {
        case ERR_PATH: return "inconsistent chain (the reference aborts with a BUG message here)";
        case ERR_STREAM: return "skeleton stream overflow";
        case ERR_EDGE: return "unknown edge kind on a skeleton";
        case ERR_STACK: return "path list overflow";
        default: return "unknown";
    }
}

// launch one stage and bring the ordered streams back
static int run_stage(lb2_sdp_batch* b, int stage, const int32_t** stream, const int64_t** off, float* kernel_ms) {
    cudaStream_t st = ctx_stream(b->ctx);
    CU(cudaSetDevice(ctx_device(b->ctx)));
    const int n = (int)b->n;
    b->off.assign(n + 1, 0);
    b->stream = nullptr; b->outs = nullptr;
    b->pairs = 0;
    if (kernel_ms) *kernel_ms = 0.f;
    if (n > 0) {
        CU(cudaMemsetAsync(b->d_counter.p, 0, sizeof(unsigned int), st));
        const int warps_per_block = 4;
        const int grid = std::max(1, std::min((n + warps_per_block - 1) / warps_per_block, ctx_sm_count(b->ctx) * 8));
        CU(cudaEventRecord(b->ev0, st));
        sdp_kernel<<<grid, warps_per_block * 32, 0, st>>>(stage, b->para, n, b->d_reads.p, b->d_order.p, b->d_seed_id.p, b->d_map_n.p,
                                                          b->d_hoff.p, b->d_hits.p, b->d_hseed.p, b->d_rflat.p, b->d_regions.p,
                                                          b->d_pts.p, b->d_scratch.p, b->d_outs.p, b->d_counter.p);
        CU(cudaGetLastError());
        sdp_scan_kernel<<<1, 1024, 0, st>>>(b->d_outs.p, n, b->d_off.p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(b->ev1, st));
        // read-backs land in page-locked memory: [outs][offsets][stream words]; the stream's size is known after the first wait
        const size_t o_outs = 0, o_off = up64(sizeof(DOut) * (size_t)n), o_dense = o_off + up64(8 * ((size_t)n + 1));
        CU(b->down.reserve(o_dense + 64));
        CU(cudaMemcpyAsync(b->down.p + o_outs, b->d_outs.p, (size_t)n * sizeof(DOut), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(b->down.p + o_off, b->d_off.p, (size_t)(n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, st));
        CU(wait_stream(st, b->ev_wait));
        b->d2h += (int64_t)n * sizeof(DOut) + (int64_t)(n + 1) * 8;
        const DOut* outs = reinterpret_cast<const DOut*>(b->down.p + o_outs);
        const long long* h_off = reinterpret_cast<const long long*>(b->down.p + o_off);
        for (int r = 0; r < n; ++r) {
            if (outs[r].err) return set_error("lb2_sdp: read %d: %s", r, err_name(outs[r].err));
            b->pairs += outs[r].pairs;
        }
        for (int r = 0; r <= n; ++r) b->off[r] = h_off[r];
        const long long total = h_off[n];
        CU(b->d_dense.reserve((size_t)total));
        if (total > 0) {
            if (b->down.cap < o_dense + (size_t)total * 4 + 64) {       // grow the landing zone (outs / offsets were consumed above)
                CU(b->down.reserve(o_dense + (size_t)total * 4 + 64));
            }
            const int threads = 256, blocks = (int)(((long long)n * 32 + threads - 1) / threads);
            sdp_gather_kernel<<<blocks, threads, 0, st>>>(b->d_reads.p, b->d_outs.p, n, b->para.ske_max, b->d_scratch.p, b->d_off.p, b->d_dense.p);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(b->down.p + o_dense, b->d_dense.p, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, st));
            CU(wait_stream(st, b->ev_wait));
            b->d2h += total * 4;
        }
        b->stream = reinterpret_cast<const int32_t*>(b->down.p + o_dense);
        if (kernel_ms) CU(cudaEventElapsedTime(kernel_ms, b->ev0, b->ev1));
    }
    static const int32_t none = 0;
    if (stream) *stream = b->stream ? b->stream : &none;
    if (off) *off = b->off.data();
    return 0;
}

extern "C" int lb2_sdp_run_bcc(lb2_sdp_batch* b, const int32_t** stream, const int64_t** off, float* kernel_ms) {
    if (!b) return set_error("lb2_sdp_run_bcc: batch is NULL");
    return run_stage(b, 1, stream, off, kernel_ms);
}

// The read intervals stage 1 left unaligned, each with the reference ends of its flanks:
// get_remain_reg (src/lamsa_aln.c:548-572) with aln_sort_reg (:477, stable like glibc's merge sort)
// and aln_merg_reg (:498-519, neighbours closer than bwt_seed_len are merged).
namespace {
struct Aligned { int beg, end; std::vector<DPoint> rb, re; };
void remaining_regions(const lb2_sdp_para& P, int read_len, const lb2_sdp_reg* regs, int n_reg,
                       std::vector<DRegion>& out, std::vector<DPoint>& pts) {
    const int lo = P.seed_len, hi = read_len;
    auto emit = [&](int beg, int end, const std::vector<DPoint>* b, const std::vector<DPoint>* e) {
        DRegion r; r.beg = beg; r.end = end; r.bn = b ? (int)b->size() : 0; r.en = e ? (int)e->size() : 0;
        r.pt_first = (int64_t)pts.size();
        if (b) pts.insert(pts.end(), b->begin(), b->end());
        if (e) pts.insert(pts.end(), e->begin(), e->end());
        out.push_back(r);
    };
    if (n_reg == 0) {
        if (lo < read_len && read_len <= hi) emit(1, read_len, nullptr, nullptr);
        return;
    }
    std::vector<int> ord(n_reg);
    std::iota(ord.begin(), ord.end(), 0);
    // TIE ORDER: the reference sorts with qsort(reg_comp) (src/lamsa_aln.c:476), whose order of records with equal `beg` is
    // whatever the C library's qsort does: glibc up to 2.36 merge-sorts (stable, the order reproduced here and by the
    // oracle); glibc >= 2.37 uses an introsort that is not stable.  Records with equal `beg` are therefore the one place
    // where a reference binary built against another libc may legitimately differ; ties are broken by input order here.
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return regs[a].beg < regs[b].beg; });
    std::vector<Aligned> A;
    for (int k = 0; k < n_reg; ++k) {
        const lb2_sdp_reg& g = regs[ord[k]];
        const DPoint pb{g.ref_beg, g.chr, 0}, pe{g.ref_end, g.chr, 0};
        if (!A.empty() && g.beg - A.back().end - 1 < P.bwt_seed_len) {
            Aligned& c = A.back();
            if (g.end > c.end) c.end = g.end;
            c.rb.push_back(pb); c.re.push_back(pe);
        } else {
            A.push_back(Aligned{g.beg, g.end, {pb}, {pe}});
        }
    }
    const int n = (int)A.size();
    if (A[0].beg > lo && A[0].beg - 1 <= hi) emit(1, A[0].beg - 1, nullptr, &A[0].rb);
    for (int i = 1; i < n; ++i)
        if (A[i].beg - A[i - 1].end > lo && A[i].beg - 1 - A[i - 1].end <= hi)
            emit(A[i - 1].end + 1, A[i].beg - 1, &A[i - 1].re, &A[i].rb);
    if (read_len - A[n - 1].end > lo && read_len - A[n - 1].end <= hi) emit(A[n - 1].end + 1, read_len, &A[n - 1].re, nullptr);
}
}  // namespace

extern "C" int lb2_sdp_run_remain(lb2_sdp_batch* b, const lb2_sdp_read* reads, const lb2_sdp_reg* regs,
                                  const int32_t** stream, const int64_t** off, float* kernel_ms) {
    if (!b || (b->n > 0 && !reads)) return set_error("lb2_sdp_run_remain: bad arguments");
    CU(cudaSetDevice(ctx_device(b->ctx)));
    std::vector<DRegion> h_regions; std::vector<DPoint> h_pts;
    for (int64_t r = 0; r < b->n; ++r) {
        if (reads[r].n_reg < 0 || (reads[r].n_reg > 0 && !regs)) return set_error("lb2_sdp_run_remain: read %lld: bad region list", (long long)r);
        DRead& d = b->reads[r];
        d.region_first = (int64_t)h_regions.size();
        remaining_regions(b->para, d.read_len, regs ? regs + reads[r].reg_first : nullptr, reads[r].n_reg, h_regions, h_pts);
        d.n_region = (int32_t)(h_regions.size() - d.region_first);
    }
    cudaStream_t st = ctx_stream(b->ctx);
    CU(b->d_regions.reserve(h_regions.size()));
    CU(b->d_pts.reserve(h_pts.size()));
    // through the page-locked staging of this stage (untouched until the next remain stage: no wait here)
    const size_t o_pts = up64(h_regions.size() * sizeof(DRegion));
    CU(b->up2.reserve(o_pts + up64(h_pts.size() * sizeof(DPoint)) + 64));
    if (!h_regions.empty()) {
        memcpy(b->up2.p, h_regions.data(), h_regions.size() * sizeof(DRegion));
        CU(cudaMemcpyAsync(b->d_regions.p, b->up2.p, h_regions.size() * sizeof(DRegion), cudaMemcpyHostToDevice, st));
    }
    if (!h_pts.empty()) {
        memcpy(b->up2.p + o_pts, h_pts.data(), h_pts.size() * sizeof(DPoint));
        CU(cudaMemcpyAsync(b->d_pts.p, b->up2.p + o_pts, h_pts.size() * sizeof(DPoint), cudaMemcpyHostToDevice, st));
    }
    if (b->n > 0) CU(cudaMemcpyAsync(b->d_reads.p, b->reads, (size_t)b->n * sizeof(DRead), cudaMemcpyHostToDevice, st));
    b->h2d += (int64_t)(h_regions.size() * sizeof(DRegion) + h_pts.size() * sizeof(DPoint) + (size_t)b->n * sizeof(DRead));
    return run_stage(b, 2, stream, off, kernel_ms);
}

extern "C" int lb2_sdp_stats(const lb2_sdp_batch* b, int64_t* pairs, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!b) return set_error("lb2_sdp_stats: batch is NULL");
    if (pairs) *pairs = b->pairs;
    if (h2d_bytes) *h2d_bytes = b->h2d;
    if (d2h_bytes) *d2h_bytes = b->d2h;
    return 0;
}

static int tracked_io(lb2_sdp_batch* b, uint8_t* flags, int set) {
    if (!b || (b->n_hits > 0 && !flags)) return set_error("lb2_sdp_%s_tracked: bad arguments", set ? "set" : "get");
    if (b->n == 0 || b->n_hits == 0) return 0;
    CU(cudaSetDevice(ctx_device(b->ctx)));
    cudaStream_t st = ctx_stream(b->ctx);
    CU(b->d_flags.reserve((size_t)b->n_hits));
    CU(b->flags_pin.reserve((size_t)b->n_hits + 64));
    if (set) {      // no wait: the staging is not written again before the stage that follows has been waited for
        memcpy(b->flags_pin.p, flags, (size_t)b->n_hits);
        CU(cudaMemcpyAsync(b->d_flags.p, b->flags_pin.p, (size_t)b->n_hits, cudaMemcpyHostToDevice, st)); b->h2d += b->n_hits;
    }
    const int threads = 256, blocks = (int)((b->n * 32 + threads - 1) / threads);
    sdp_tracked_kernel<<<blocks, threads, 0, st>>>(b->d_reads.p, (int)b->n, b->para.ske_max, b->d_scratch.p, b->d_flags.p, set);
    CU(cudaGetLastError());
    if (!set) {
        CU(cudaMemcpyAsync(b->flags_pin.p, b->d_flags.p, (size_t)b->n_hits, cudaMemcpyDeviceToHost, st)); b->d2h += b->n_hits;
        CU(wait_stream(st, b->ev_wait));
        memcpy(flags, b->flags_pin.p, (size_t)b->n_hits);
    }
    return 0;
}
extern "C" int lb2_sdp_get_tracked(lb2_sdp_batch* b, uint8_t* flags) { return tracked_io(b, flags, 0); }
extern "C" int lb2_sdp_set_tracked(lb2_sdp_batch* b, const uint8_t* flags) { return tracked_io(b, const_cast<uint8_t*>(flags), 1); }
