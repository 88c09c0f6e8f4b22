// dp_batch.cu -- host side of the batch interface declared in include/lamsa_b200.h:
// packs DP tasks, sizes the direction scratch, launches the fill / traceback
// kernels class by class, and unpacks results.  No DP arithmetic happens here.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lamsa_b200.h"
#include "dp_fill.cuh"
#include "dp_fill16.cuh"
#include "dp_fill16d.cuh"
#include "dp_fill_lean.cuh"
#include "dp_trace.cuh"
#include "int_peak.cuh"
#include "aux_scan.cuh"
#include "ctx_internal.h"
#include "dp_pack.h"

using namespace lb2;

// ------------------------------------------------------------------ errors --
static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return 1;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    return fail("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } while (0)

extern "C" const char* lb2_last_error(void) { return g_err.c_str(); }
extern "C" void lb2_free(void* p) { free(p); }

// ----------------------------------------------------------------- context --
// launch classes, variants and the per-task classification live in dp_pack.h
static inline int class_warps(int var, int logS) {   // warps per block
    if (var == kVarGmem || var == kVarBlock) return 4;
    int wpb = 8;
    while (wpb > 1 && (size_t)wpb * var_warp_smem(var, 1 << logS) > 160 * 1024) wpb >>= 1;
    return wpb;
}

// Grow-only staging owned by a batch while it lives and parked in the context
// in between, so that a producer that submits batch after batch (or the drop-in
// entry points, one task at a time) never re-allocates pinned or device memory.
struct Buffers {
    uint8_t* h_pool = nullptr;  size_t h_pool_cap = 0;
    DTask* h_tasks = nullptr;   DResult* h_results = nullptr;  int32_t* h_order = nullptr;  size_t h_n_cap = 0;
    uint2* h_mats = nullptr;
    cigar32_t* h_cigar = nullptr; size_t h_cigar_cap = 0;      // pinned landing zone of the dense CIGAR pool
    uint8_t* d_pool = nullptr;  size_t d_pool_cap = 0;
    DTask* d_tasks = nullptr;   DResult* d_results = nullptr;  int32_t* d_order = nullptr;  size_t d_n_cap = 0;
    uint2* d_mats = nullptr;
    int32_t* d_cdense = nullptr; size_t dense_cap = 0;
    unsigned long long* d_cursor = nullptr;
    unsigned int* d_counters = nullptr;  size_t counters_cap = 0;
    int* d_err = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t up_ev = nullptr;                        // H2D of this batch finished
    cudaEvent_t done_ev = nullptr;                      // all kernels finished; blocking-sync flavour (no spinning host thread)
    bool valid = false;
    void release() {
        cudaFreeHost(h_pool); cudaFreeHost(h_tasks); cudaFreeHost(h_results); cudaFreeHost(h_order); cudaFreeHost(h_mats);
        cudaFreeHost(h_cigar);
        cudaFree(d_pool); cudaFree(d_tasks); cudaFree(d_results); cudaFree(d_order); cudaFree(d_mats);
        cudaFree(d_cdense); cudaFree(d_cursor); cudaFree(d_counters); cudaFree(d_err);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        if (up_ev) cudaEventDestroy(up_ev);
        if (done_ev) cudaEventDestroy(done_ev);
        *this = Buffers();
    }
};

struct lb2_ctx {
    int device = 0;
    std::mutex mu;
    static constexpr int kParked = 3;
    Buffers parked[kParked];                            // buffers of the last destroyed batches (lb2_dp_run keeps two chunks enqueued while it drains a third)
    cudaStream_t copy = nullptr;                        // H2D of batch k+1 overlaps the kernels of batch k
    int64_t run_h2d = 0, run_d2h = 0, run_launches = 0; // counters of the last lb2_dp_run
    float run_fill_ms = 0, run_trace_ms = 0;
    int64_t chunk_tasks = 131072;                       // lb2_dp_run pipelines chunks of about this many tasks
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint64_t scratch_limit = 0;
    uint8_t* d_z = nullptr;    size_t z_cap = 0;        // direction nibbles (+ row bands)
    int32_t* d_ctmp = nullptr; size_t ctmp_cap = 0;     // per-task reversed CIGAR scratch (words)
    // second scratch set + compute stream: lb2_dp_run alternates its chunks between the two, so the fills of chunk
    // k+1 start on the SMs the tail of chunk k leaves idle and run beside chunk k's traceback
    uint8_t* d_z1 = nullptr;    size_t z1_cap = 0;
    int32_t* d_ctmp1 = nullptr; size_t ctmp1_cap = 0;
    cudaStream_t stream1 = nullptr;
    uint8_t* d_gwin = nullptr; size_t gwin_cap = 0;     // eh[] windows too large for shared memory
    uint8_t* d_pac = nullptr;  int64_t l_pac = 0;       // resident 2-bit forward reference (lb2_ctx_set_reference)
    // grow-only staging of lb2_aux_run (pinned host + device): records, CIGAR words, read bytes, results
    uint8_t* aux_h = nullptr; uint8_t* aux_d = nullptr; size_t aux_cap = 0;
    int occ[kNumClass] = {0};                           // resident blocks per SM, filled lazily
    // side streams: the launch classes of a wave run concurrently, so the drain of one
    // class (few long tasks left) is filled by the blocks of the next
    static constexpr int kAux = 4;
    cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t fork_ev = nullptr, join_ev[kAux] = {nullptr, nullptr, nullptr, nullptr};
};

typedef void (*fill_fn)(const DTask*, const int32_t*, int, const uint8_t*, const uint8_t*, uint8_t*, DResult*,
                        const uint2*, unsigned int*, int, uint8_t*);
static fill_fn fill_table(int kind, int var) {
    if (var == kVarBlock) return fill_lean_kernel;
    if (kind == kKindGlobal) {
        switch (var) {
            case 0: return fill_kernel<1, kKindGlobal, false>;   case 1: return fill_kernel<2, kKindGlobal, false>;
            case 2: return fill_kernel<4, kKindGlobal, false>;   case 3: return fill16_kernel<2, kKindGlobal>;
            case 4: return fill16_kernel<4, kKindGlobal>;        case 5: return fill_kernel<4, kKindGlobal, true>;
            case 6: return fill16d_kernel<2, kKindGlobal, 16>;   case 7: return fill16d_kernel<2, kKindGlobal, 8>;
            case 8: return fill16d_kernel<4, kKindGlobal, 8>;    default: return fill16d_kernel<4, kKindGlobal, 16>;
        }
    }
    switch (var) {
        case 0: return fill_kernel<1, kKindExtend, false>;   case 1: return fill_kernel<2, kKindExtend, false>;
        case 2: return fill_kernel<4, kKindExtend, false>;   case 3: return fill16_kernel<2, kKindExtend>;
        case 4: return fill16_kernel<4, kKindExtend>;        case 5: return fill_kernel<4, kKindExtend, true>;
        case 6: return fill16d_kernel<2, kKindExtend, 16>;   case 7: return fill16d_kernel<2, kKindExtend, 8>;
        case 8: return fill16d_kernel<4, kKindExtend, 8>;    default: return fill16d_kernel<4, kKindExtend, 16>;
    }
}

extern "C" int lb2_ctx_create(int device, lb2_ctx** out) {
    if (!out) return fail("lb2_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("lb2_ctx_create: no CUDA device (%s); this library has no CPU path",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return fail("lb2_ctx_create: device %d of %d", device, ndev);
    CU(cudaSetDevice(device));
    // per-device facts and the one-time kernel attributes are looked up once per process: a batch producer
    // opens several contexts; the kernels' shared-memory attribute is set when a kernel is first launched on a
    // device (compute_enqueue), so that opening a context does not load every kernel of the library
    static std::mutex dev_mu;
    static int dev_major[64], dev_minor[64], dev_sms[64];
    static bool dev_known[64] = {false};
    {
        std::lock_guard<std::mutex> lk(dev_mu);
        if (device >= 64) return fail("lb2_ctx_create: device %d", device);
        if (!dev_known[device]) {
            CU(cudaDeviceGetAttribute(&dev_major[device], cudaDevAttrComputeCapabilityMajor, device));
            CU(cudaDeviceGetAttribute(&dev_minor[device], cudaDevAttrComputeCapabilityMinor, device));
            CU(cudaDeviceGetAttribute(&dev_sms[device], cudaDevAttrMultiProcessorCount, device));
            dev_known[device] = true;
        }
    }
    if (dev_major[device] < 10) return fail("lb2_ctx_create: device %d is sm_%d%d, need sm_100", device, dev_major[device], dev_minor[device]);
    lb2_ctx* c = new lb2_ctx();
    c->device = device;
    c->sm_count = dev_sms[device];
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
    for (int k = 0; k < lb2_ctx::kAux; ++k) {
        CU(cudaStreamCreateWithFlags(&c->aux[k], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->join_ev[k], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    size_t fr = 0, tot = 0;
    CU(cudaMemGetInfo(&fr, &tot));
    c->scratch_limit = (uint64_t)(fr * 0.40);
    *out = c;
    return 0;
}

extern "C" void lb2_ctx_destroy(lb2_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->d_z) cudaFree(c->d_z);
    if (c->d_ctmp) cudaFree(c->d_ctmp);
    if (c->d_z1) cudaFree(c->d_z1);
    if (c->d_ctmp1) cudaFree(c->d_ctmp1);
    if (c->stream1) cudaStreamDestroy(c->stream1);
    if (c->d_gwin) cudaFree(c->d_gwin);
    if (c->d_pac) cudaFree(c->d_pac);
    if (c->aux_h) cudaFreeHost(c->aux_h);
    if (c->aux_d) cudaFree(c->aux_d);
    for (auto& p : c->parked) if (p.valid) p.release();
    if (c->copy) cudaStreamDestroy(c->copy);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int k = 0; k < lb2_ctx::kAux; ++k) {
        if (c->aux[k]) cudaStreamDestroy(c->aux[k]);
        if (c->join_ev[k]) cudaEventDestroy(c->join_ev[k]);
    }
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    delete c;
}

extern "C" int lb2_ctx_set_reference(lb2_ctx* c, const uint8_t* pac, int64_t l_pac) {
    if (!c || !pac || l_pac <= 0) return fail("lb2_ctx_set_reference: bad argument");
    if (l_pac >= ((int64_t)1 << 32)) return fail("lb2_ctx_set_reference: %lld bases exceed the 32-bit window coordinate", (long long)l_pac);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_pac) CU(cudaFree(c->d_pac));
    c->d_pac = nullptr; c->l_pac = 0;
    const size_t bytes = (size_t)(l_pac / 4 + 1);
    CU(cudaMalloc(&c->d_pac, bytes + 16));
    CU(cudaMemcpy(c->d_pac, pac, bytes, cudaMemcpyHostToDevice));
    c->l_pac = l_pac;
    return 0;
}

extern "C" int lb2_ctx_set_chunk_tasks(lb2_ctx* c, int64_t tasks) {
    if (!c || tasks < 1) return fail("lb2_ctx_set_chunk_tasks: bad argument");
    c->chunk_tasks = tasks;
    return 0;
}

extern "C" int lb2_ctx_set_scratch_limit(lb2_ctx* c, uint64_t bytes) {
    if (!c) return fail("ctx is NULL");
    c->scratch_limit = bytes < (1u << 20) ? (1u << 20) : bytes;
    return 0;
}

// ------------------------------------------------------------------- batch --
struct Wave {
    int cls_off[kNumClass + 1];      // ranges inside `order`, by class
    int first, count;                // range inside `order`
    int long_count;                  // tasks walked by trace_long_kernel: order[n + first .. + long_count)
    uint64_t z_bytes, ctmp_words;
};

struct lb2_batch {
    lb2_ctx* ctx = nullptr;
    int64_t n = 0;
    Buffers B;
    size_t pool_bytes = 0;
    uint64_t dense_cap = 0;
    int n_counters = 0;
    // aliases into B (set by lb2_batch_create)
    uint8_t* h_pool = nullptr; DTask* h_tasks = nullptr; int32_t* h_order = nullptr; uint2* h_mats = nullptr;
    DResult* h_results = nullptr;
    uint8_t* d_pool = nullptr; DTask* d_tasks = nullptr; int32_t* d_order = nullptr; uint2* d_mats = nullptr;
    DResult* d_results = nullptr; int32_t* d_cdense = nullptr; unsigned long long* d_cursor = nullptr;
    unsigned int* d_counters = nullptr; int* d_err = nullptr;
    std::vector<Wave> waves;
    std::vector<uint8_t> flags;      // per task: LB2_FLAG_*
    std::vector<int16_t> cls;        // per task: launch class
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // aliases of B.ev
    int64_t h2d_bytes = 0, d2h_bytes = 0, launches = 0;
    float fill_ms = 0, trace_ms = 0;
    bool uploaded = false, enqueued = false, computed = false;
    int slot = 0;                                                // which scratch set / compute stream of the context (0, 1)
    const uint8_t* raw_src = nullptr; size_t raw_bytes = 0;      // pooled batch: the range of the caller's pool its tasks use
    std::vector<cudaEvent_t> wave_ev;      // only when scratch forces several waves
    bool class_timing = false;             // lb2_batch_set_class_timing: classes one after the other, each with its own events
    struct ClassRun { int cls; int tasks; cudaEvent_t t0, t1; float ms; };
    std::vector<ClassRun> class_runs;
};

extern "C" void lb2_batch_destroy(lb2_batch* b) {
    if (!b) return;
    if (b->ctx) {
        cudaSetDevice(b->ctx->device);
        std::lock_guard<std::mutex> lk(b->ctx->mu);
        for (auto& p : b->ctx->parked)
            if (b->B.valid && !p.valid) { p = b->B; b->B = Buffers(); }
    }
    if (b->B.valid) b->B.release();
    for (auto& e : b->wave_ev) cudaEventDestroy(e);
    for (auto& r : b->class_runs) { cudaEventDestroy(r.t0); cudaEventDestroy(r.t1); }
    delete b;
}

static unsigned hw_threads() {
    unsigned t = std::thread::hardware_concurrency();
    const char* env = getenv("LB2_HOST_THREADS");
    if (env && atoi(env) > 0) t = (unsigned)atoi(env);
    if (t == 0) t = 1;
    return t > 32 ? 32 : t;
}

template <class F>
static void parallel_for(int64_t n, F f) {
    unsigned nt = hw_threads();
    if (n < 4096 || nt == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const int64_t step = (n + nt - 1) / nt;
    for (unsigned k = 0; k < nt; ++k) {
        const int64_t a = k * step, e = std::min<int64_t>(n, a + step);
        if (a >= e) break;
        th.emplace_back([=] { f(a, e); });
    }
    for (auto& t : th) t.join();
}

// Grow-only buffers grow geometrically with a floor: every cudaMalloc / cudaFree / cudaMallocHost is a
// device-wide synchronisation, and a producer that submits batch after batch of slightly different size
// (the fiber scheduler, many contexts at once) must reach a steady state without them.
static size_t grown(size_t need, size_t old_cap, size_t floor_) {
    return std::max(std::max(need + need / 4, old_cap * 2), floor_);
}

// ---- the three steps every batch creator shares ------------------------------------------------
// (1) pinned host staging for n tasks and pool_bytes of sequences (grow-only, taken from the context's parked set)
static int alloc_host(lb2_batch* b, size_t pool_bytes, bool host_pool = true) {
    lb2_ctx* ctx = b->ctx;
    b->pool_bytes = pool_bytes + 64;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        // prefer the parked set whose pinned pool is large enough
        int pick = -1;
        for (int k = 0; k < lb2_ctx::kParked; ++k)
            if (ctx->parked[k].valid && (pick < 0 || ctx->parked[k].h_pool_cap >= b->pool_bytes)) pick = k;
        if (pick >= 0) { b->B = ctx->parked[pick]; ctx->parked[pick] = Buffers(); }
    }
    Buffers& B = b->B;
    B.valid = true;
    const size_t n1c = (size_t)std::max<int64_t>(b->n, 1);
    if (host_pool && B.h_pool_cap < b->pool_bytes) {
        const size_t cap = grown(b->pool_bytes, B.h_pool_cap, (size_t)1 << 20);
        cudaFreeHost(B.h_pool); B.h_pool = nullptr; B.h_pool_cap = 0;
        CU(cudaMallocHost(&B.h_pool, cap)); B.h_pool_cap = cap;
    }
    if (B.h_n_cap < n1c) {
        const size_t cap = grown(n1c, B.h_n_cap, 4096);
        cudaFreeHost(B.h_tasks); cudaFreeHost(B.h_results); cudaFreeHost(B.h_order);
        B.h_tasks = nullptr; B.h_results = nullptr; B.h_order = nullptr; B.h_n_cap = 0;
        CU(cudaMallocHost(&B.h_tasks, sizeof(DTask) * cap));
        CU(cudaMallocHost(&B.h_results, sizeof(DResult) * cap));
        CU(cudaMallocHost(&B.h_order, sizeof(int32_t) * cap * 2));       // second half: the tasks of trace_long_kernel, per wave
        B.h_n_cap = cap;
    }
    if (!B.h_mats) CU(cudaMallocHost(&B.h_mats, sizeof(uint2) * kMaxMats * 8));
    b->h_pool = B.h_pool; b->h_tasks = B.h_tasks; b->h_results = B.h_results; b->h_order = B.h_order; b->h_mats = B.h_mats;
    memset(b->h_mats, 0, sizeof(uint2) * kMaxMats * 8);
    return 0;
}

// (2) waves (consecutive tasks whose scratch fits the limit), launch order inside a wave (counting sort by
// class, then descending log-spaced cost bin: the persistent warps only need an approximately longest-first
// order), scratch offsets.  h_tasks[i] must hold task i's descriptor; cls / bin / zsz / ctmpw describe it.
static int layout_waves(lb2_batch* b, const int16_t* cls, const uint8_t* bin, const uint64_t* zsz, const int32_t* ctmpw) {
    const int64_t n = b->n;
    const uint64_t zlimit = b->ctx->scratch_limit;
    uint64_t dense = 0;
    {
        int64_t i = 0;
        while (i < n) {
            Wave wv; wv.first = (int)i; wv.z_bytes = 0; wv.ctmp_words = 0;
            int64_t j = i;
            while (j < n) {
                const uint64_t nz = wv.z_bytes + zsz[j], nc = wv.ctmp_words + (uint64_t)ctmpw[j];
                if (j > i && nz + nc * 4 > zlimit) break;        // a single task larger than the limit still runs alone
                wv.z_bytes = nz; wv.ctmp_words = nc; ++j;
            }
            wv.count = (int)(j - i);
            dense += wv.ctmp_words;
            b->waves.push_back(wv);
            i = j;
        }
    }
    b->cls.assign(cls, cls + n);
    for (auto& wv : b->waves) {
        // classes in use are few: histogram only those (a producer submits thousands of small batches)
        thread_local std::vector<int32_t> keys_tl;
        if (keys_tl.size() < (size_t)wv.count + 1) keys_tl.resize((size_t)wv.count + 1);
        int32_t* const keys = keys_tl.data();
        parallel_for(wv.count, [&](int64_t lo_i, int64_t hi_i) {
            for (int64_t k = lo_i; k < hi_i; ++k) keys[(size_t)k] = (int32_t)cls[wv.first + k] * kCostBins + bin[wv.first + k];
        });
        bool used[kNumClass] = {false};
        for (int k = 0; k < wv.count; ++k) used[keys[(size_t)k] / kCostBins] = true;
        int slot_of[kNumClass], nused = 0;
        for (int c = 0; c < kNumClass; ++c) slot_of[c] = used[c] ? nused++ : -1;
        std::vector<int32_t> hist((size_t)nused * kCostBins + 1, 0);
        for (int k = 0; k < wv.count; ++k) {
            const int c = keys[(size_t)k] / kCostBins, bb = keys[(size_t)k] % kCostBins;
            keys[(size_t)k] = slot_of[c] * kCostBins + bb;
            ++hist[(size_t)keys[(size_t)k] + 1];
        }
        for (size_t k = 1; k < hist.size(); ++k) hist[k] += hist[k - 1];
        for (int c = 0, u = 0; c <= kNumClass; ++c) {
            wv.cls_off[c] = hist[(size_t)u * kCostBins];
            if (c < kNumClass && used[c]) ++u;
        }
        int32_t* ord = b->h_order + wv.first;
        for (int k = 0; k < wv.count; ++k) ord[hist[(size_t)keys[(size_t)k]]++] = wv.first + k;
        wv.long_count = 0;
        for (int k = 0; k < wv.count; ++k) {
            const DTask& d = b->h_tasks[wv.first + k];
            if ((d.want_dir & kWantDir) && d.qlen + d.tlen >= kLongTrace) b->h_order[n + wv.first + wv.long_count++] = wv.first + k;
        }
        uint64_t z = 0, cw = 0;
        for (int k = 0; k < wv.count; ++k) {          // scratch offsets follow the original order
            const int64_t a = wv.first + k;
            DTask& d = b->h_tasks[a];
            d.z_off = z; z += zsz[a];
            cw += (uint64_t)ctmpw[a]; d.ctmp_end = cw; d.ctmp_cap = ctmpw[a];
        }
    }
    b->dense_cap = dense + 16;
    return 0;
}

// (3) device buffers (grow-only, reused across batches) and the context's scratch
static int alloc_device(lb2_batch* b) {
    lb2_ctx* ctx = b->ctx;
    Buffers& B = b->B;
    const size_t n1c = (size_t)std::max<int64_t>(b->n, 1);
    if (B.d_pool_cap < b->pool_bytes) {
        const size_t cap = grown(b->pool_bytes, B.d_pool_cap, (size_t)1 << 20);
        cudaFree(B.d_pool); B.d_pool = nullptr; B.d_pool_cap = 0;
        CU(cudaMalloc(&B.d_pool, cap)); B.d_pool_cap = cap;
    }
    if (B.d_n_cap < n1c) {
        const size_t cap = grown(n1c, B.d_n_cap, 4096);
        cudaFree(B.d_tasks); cudaFree(B.d_results); cudaFree(B.d_order);
        B.d_tasks = nullptr; B.d_results = nullptr; B.d_order = nullptr; B.d_n_cap = 0;
        CU(cudaMalloc(&B.d_tasks, sizeof(DTask) * cap));
        CU(cudaMalloc(&B.d_results, sizeof(DResult) * cap));
        CU(cudaMalloc(&B.d_order, sizeof(int32_t) * cap * 2));
        B.d_n_cap = cap;
    }
    if (!B.d_mats) CU(cudaMalloc(&B.d_mats, sizeof(uint2) * kMaxMats * 8));
    if (B.dense_cap < b->dense_cap) {
        const size_t cap = grown(b->dense_cap, B.dense_cap, (size_t)1 << 18);
        cudaFree(B.d_cdense); B.d_cdense = nullptr; B.dense_cap = 0;
        CU(cudaMalloc(&B.d_cdense, sizeof(int32_t) * cap)); B.dense_cap = cap;
    }
    if (!B.d_cursor) CU(cudaMalloc(&B.d_cursor, sizeof(unsigned long long)));
    b->n_counters = (int)b->waves.size() * kNumClass + 1;
    if (B.counters_cap < (size_t)b->n_counters) {
        cudaFree(B.d_counters); B.d_counters = nullptr; B.counters_cap = 0;
        CU(cudaMalloc(&B.d_counters, sizeof(unsigned int) * b->n_counters * 2)); B.counters_cap = (size_t)b->n_counters * 2;
    }
    if (!B.d_err) CU(cudaMalloc(&B.d_err, sizeof(int)));
    for (auto& e : B.ev) if (!e) CU(cudaEventCreate(&e));
    if (!B.up_ev) CU(cudaEventCreateWithFlags(&B.up_ev, cudaEventDisableTiming));
    if (!B.done_ev) CU(cudaEventCreateWithFlags(&B.done_ev, cudaEventDisableTiming | cudaEventBlockingSync));
    b->d_pool = B.d_pool; b->d_tasks = B.d_tasks; b->d_results = B.d_results; b->d_order = B.d_order; b->d_mats = B.d_mats;
    b->d_cdense = B.d_cdense; b->d_cursor = B.d_cursor; b->d_counters = B.d_counters; b->d_err = B.d_err;
    for (int k = 0; k < 4; ++k) b->ev[k] = B.ev[k];
    // grow the context's scratch.  Kernels of an earlier batch of this context may still be using it
    // (lb2_dp_run pipelines chunks): wait for them explicitly rather than rely on cudaFree's implicit barrier.
    uint64_t zmax = 16, cmax = 16;
    for (auto& wv : b->waves) { zmax = std::max(zmax, wv.z_bytes); cmax = std::max(cmax, wv.ctmp_words); }
    if (b->slot && !ctx->stream1) CU(cudaStreamCreateWithFlags(&ctx->stream1, cudaStreamNonBlocking));
    uint8_t*& dz = b->slot ? ctx->d_z1 : ctx->d_z;          size_t& zcap = b->slot ? ctx->z1_cap : ctx->z_cap;
    int32_t*& dct = b->slot ? ctx->d_ctmp1 : ctx->d_ctmp;   size_t& ccap = b->slot ? ctx->ctmp1_cap : ctx->ctmp_cap;
    if (zcap < zmax || ccap < cmax) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        CU(cudaStreamSynchronize(b->slot ? ctx->stream1 : ctx->stream));
        for (int q = 0; q < lb2_ctx::kAux; ++q) CU(cudaStreamSynchronize(ctx->aux[q]));
        if (zcap < zmax) {
            const size_t cap = grown(zmax, zcap, (size_t)16 << 20);
            if (dz) CU(cudaFree(dz));
            dz = nullptr; zcap = 0;
            CU(cudaMalloc(&dz, cap + 64)); zcap = cap;
        }
        if (ccap < cmax) {
            const size_t cap = grown(cmax, ccap, (size_t)1 << 20);
            if (dct) CU(cudaFree(dct));
            dct = nullptr; ccap = 0;
            CU(cudaMalloc(&dct, (cap + 16) * 4)); ccap = cap;
        }
    }
    return 0;
}

// `pool` != nullptr: every sequence of `tasks` lies inside [pool, pool + pool_bytes) (lb2_batch_create_pool); the range
// of the pool the tasks use is then uploaded as it lies and the kernels read it in place (DTask offsets in bytes,
// kRawOff) instead of a 32-byte aligned, padded copy made here.
static int batch_create_impl(lb2_ctx* ctx, int64_t n, const lb2_task* tasks, const uint8_t* upool, int64_t upool_bytes, lb2_batch** out, int slot = 0) {
    if (!ctx || !out || (n > 0 && !tasks)) return fail("lb2_batch_create: NULL argument");
    if (n < 0 || n > (int64_t)1 << 30) return fail("lb2_batch_create: n=%lld out of range", (long long)n);
    CU(cudaSetDevice(ctx->device));
    lb2_batch* b = new lb2_batch();
    b->ctx = ctx; b->n = n; b->slot = slot;
    struct Guard { lb2_batch* b; bool ok = false; ~Guard() { if (!ok) lb2_batch_destroy(b); } } guard{b};

    // ---- pass 1 (parallel): validate, final band, kernel variant, sizes
    // host scratch of this function is kept per calling thread (a run of equally sized chunks re-uses it: fresh vectors
    // of tens of megabytes per chunk cost more in page faults than the classification itself)
    thread_local std::vector<PackedTask> pk_tl;
    thread_local std::vector<uint64_t> qoff_tl, zsz_tl;
    thread_local std::vector<int16_t> cls_tl; thread_local std::vector<uint8_t> bin_tl; thread_local std::vector<int32_t> ctmpw_tl;
    if (pk_tl.size() < (size_t)n + 1) {
        pk_tl.resize((size_t)n + 1); qoff_tl.resize((size_t)n + 1); zsz_tl.resize((size_t)n + 1); cls_tl.resize((size_t)n + 1);
        bin_tl.resize((size_t)n + 1); ctmpw_tl.resize((size_t)n + 1);
    }
    // plain pointers: the worker threads of parallel_for must see THIS thread's vectors
    PackedTask* const pk = pk_tl.data(); uint64_t* const qoff = qoff_tl.data(); uint64_t* const zsz = zsz_tl.data();
    int16_t* const cls = cls_tl.data(); uint8_t* const bin = bin_tl.data(); int32_t* const ctmpw = ctmpw_tl.data();
    uint64_t raw_min = UINT64_MAX, raw_max = 0;
    std::vector<std::vector<int8_t>> mats;           // distinct matrices, each 64 entries (8x8, zero padded)
    std::mutex mats_mu;
    std::mutex err_mu; int64_t err_i = -1; std::string err_msg;
    auto matrix_id = [&](const lb2_task& t) -> int {
        int8_t m8[64]; memset(m8, 0, sizeof m8);
        for (int a = 0; a < t.m; ++a) for (int c = 0; c < t.m; ++c) m8[a * 8 + c] = t.mat[a * t.m + c];
        std::lock_guard<std::mutex> lk(mats_mu);
        for (size_t k = 0; k < mats.size(); ++k) if (!memcmp(mats[k].data(), m8, 64)) return (int)k;
        if ((int)mats.size() == kMaxMats) return -1;
        mats.emplace_back(m8, m8 + 64);
        return (int)mats.size() - 1;
    };
    const int64_t l_pac = ctx->d_pac ? ctx->l_pac : -1;
    parallel_for(n, [&](int64_t lo_i, int64_t hi_i) {
        const int8_t* last_mat = nullptr; int last_m = 0, last_id = -1;     // per-thread cache: tasks share matrices
        char msg[200];
        uint64_t my_min = UINT64_MAX, my_max = 0;
        for (int64_t i = lo_i; i < hi_i; ++i) {
            const lb2_task& t = tasks[i];
            int bad = classify_task(t, l_pac, pk[(size_t)i], msg, sizeof msg);
            if (!bad && upool) {        // the range of the caller's pool this batch reads
                const bool tp = (t.flags & LB2_FLAG_TARGET_PAC) != 0;
                for (int q = 0; q < (tp ? 1 : 2) && !bad; ++q) {
                    const int len = q ? t.tlen : t.qlen;
                    if (!len) continue;
                    const uint8_t* ptr = q ? t.target : t.query;
                    if (ptr < upool || ptr + len > upool + upool_bytes) { snprintf(msg, sizeof msg, "%s lies outside the pool", q ? "target" : "query"); bad = 1; break; }
                    my_min = std::min<uint64_t>(my_min, (uint64_t)(ptr - upool)); my_max = std::max<uint64_t>(my_max, (uint64_t)(ptr - upool) + len);
                }
            }
            if (!bad && (t.mat != last_mat || t.m != last_m)) {
                last_id = matrix_id(t); last_mat = t.mat; last_m = t.m;
                if (last_id < 0) { snprintf(msg, sizeof msg, "more than %d distinct scoring matrices in one batch", kMaxMats); bad = 1; }
            }
            if (bad) {
                std::lock_guard<std::mutex> lk(err_mu);
                if (err_i < 0 || i < err_i) { err_i = i; err_msg = "task " + std::to_string(i) + ": " + msg; }
                return;
            }
            pk[(size_t)i].d.mat_id = (uint8_t)last_id;
        }
        std::lock_guard<std::mutex> lk(err_mu);
        raw_min = std::min(raw_min, my_min); raw_max = std::max(raw_max, my_max);
    });
    if (err_i >= 0) return fail("%s", err_msg.c_str());
    uint64_t pool = 0;
    for (int64_t i = 0; i < n; ++i) {
        qoff[(size_t)i] = pool;
        pool += pool_bytes_query(tasks[i].qlen);
        if (!(tasks[i].flags & LB2_FLAG_TARGET_PAC)) pool += pool_bytes_target(tasks[i].tlen);
    }
    if (pool >> 37) return fail("sequence pool of %llu bytes is too large for one batch", (unsigned long long)pool);
    if (alloc_host(b, pool, upool == nullptr)) return 1;
    for (size_t k = 0; k < mats.size(); ++k)
        for (int r = 0; r < 8; ++r) memcpy(&b->h_mats[k * 8 + r], mats[k].data() + r * 8, 8);
    uint64_t raw_lo = 0;
    if (upool) {
        // tasks in pool order give one tight range per chunk
        uint64_t lo = raw_min, hi = raw_max;
        if (lo > hi) { lo = 0; hi = 0; }
        raw_lo = lo & ~uint64_t(15);
        b->raw_src = upool + raw_lo; b->raw_bytes = (size_t)(hi - raw_lo);
        if (b->raw_bytes >> 32) return fail("a pooled batch may span at most 4 GiB of its pool (this one spans %llu bytes): lower lb2_ctx_set_chunk_tasks or order the pool like the tasks", (unsigned long long)b->raw_bytes);
        b->pool_bytes = b->raw_bytes + 64;           // what alloc_device sizes the device pool by
    }

    // ---- pass 2: descriptors and sequences into the pinned staging (parallel)
    b->flags.resize((size_t)n);
    uint8_t* hp = b->h_pool;
    DTask* ht = b->h_tasks;
    const bool stream = pool > ((uint64_t)8 << 20);          // large staging: written once, read by the DMA engine only
    parallel_for(n, [&, hp, ht](int64_t a, int64_t e) {
        for (int64_t i = a; i < e; ++i) {
            PackedTask& p = pk[(size_t)i];
            if (upool) {
                const lb2_task& t = tasks[i];
                p.d.want_dir |= kRawOff;
                p.d.q_off32 = t.qlen ? (uint32_t)((uint64_t)(t.query - upool) - raw_lo) : 0u;
                if (!(t.flags & LB2_FLAG_TARGET_PAC)) p.d.t_off32 = t.tlen ? (uint32_t)((uint64_t)(t.target - upool) - raw_lo) : 0u;
            } else
                copy_sequences(tasks[i], p, hp, qoff[(size_t)i], false, stream);
            ht[i] = p.d;
            b->flags[(size_t)i] = p.flags; cls[(size_t)i] = p.cls; bin[(size_t)i] = p.bin;
            zsz[(size_t)i] = p.zsz; ctmpw[(size_t)i] = p.ctmpw;
        }
#if defined(__SSE2__)
        if (stream) _mm_sfence();
#endif
    });
    if (layout_waves(b, cls, bin, zsz, ctmpw)) return 1;
    if (alloc_device(b)) return 1;
    guard.ok = true;
    *out = b;
    return 0;
}

extern "C" int lb2_batch_create(lb2_ctx* ctx, int64_t n, const lb2_task* tasks, lb2_batch** out) {
    return batch_create_impl(ctx, n, tasks, nullptr, 0, out);
}
extern "C" int lb2_batch_create_pool(lb2_ctx* ctx, const uint8_t* pool, int64_t pool_bytes, int64_t n, const lb2_task* tasks, lb2_batch** out) {
    if (!pool || pool_bytes < 0) return fail("lb2_batch_create_pool: NULL pool");
    return batch_create_impl(ctx, n, tasks, pool, pool_bytes, out);
}
// page-locked host memory for pools handed to lb2_batch_create_pool / lb2_dp_run_pool (DMA reads it in place)
extern "C" int lb2_host_alloc(size_t bytes, void** out) {
    if (!out) return fail("lb2_host_alloc: out is NULL");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return 0;
}
extern "C" void lb2_host_free(void* p) { if (p) cudaFreeHost(p); }

// The batch producer's creator: tasks arrive classified and copied by the threads that parked them
// (dp_pack.h: TaskBlob); this only concatenates the blobs into the pinned staging and lays out the launch.
int lb2::batch_create_staged(lb2_ctx* ctx, const TaskBlob* const* blobs, int nblobs, lb2_batch** out) {
    if (!ctx || !out) return fail("batch_create_staged: NULL argument");
    CU(cudaSetDevice(ctx->device));
    int64_t n = 0; uint64_t pool = 0;
    for (int k = 0; k < nblobs; ++k) { n += (int64_t)blobs[k]->tasks.size(); pool += blobs[k]->pool.size(); }
    if (pool >> 37) return fail("sequence pool of %llu bytes is too large for one batch", (unsigned long long)pool);
    lb2_batch* b = new lb2_batch();
    b->ctx = ctx; b->n = n;
    struct Guard { lb2_batch* b; bool ok = false; ~Guard() { if (!ok) lb2_batch_destroy(b); } } guard{b};
    if (alloc_host(b, pool)) return 1;
    // batch-wide matrix table
    std::vector<std::array<int8_t, 64>> mats;
    std::vector<std::vector<uint8_t>> remap((size_t)nblobs);
    for (int k = 0; k < nblobs; ++k) {
        for (const auto& m8 : blobs[k]->mats) {
            size_t id = 0;
            while (id < mats.size() && mats[id] != m8) ++id;
            if (id == mats.size()) {
                if ((int)mats.size() == kMaxMats) return fail("more than %d distinct scoring matrices in one batch", kMaxMats);
                mats.push_back(m8);
            }
            remap[(size_t)k].push_back((uint8_t)id);
        }
    }
    for (size_t k = 0; k < mats.size(); ++k)
        for (int r = 0; r < 8; ++r) memcpy(&b->h_mats[k * 8 + r], mats[k].data() + r * 8, 8);
    b->flags.resize((size_t)n);
    std::vector<int16_t> cls((size_t)n); std::vector<uint8_t> bin((size_t)n);
    std::vector<uint64_t> zsz((size_t)n); std::vector<int32_t> ctmpw((size_t)n);
    int64_t at = 0; uint64_t pool_at = 0;
    for (int k = 0; k < nblobs; ++k) {
        const TaskBlob& bl = *blobs[k];
        if (!bl.pool.empty()) memcpy(b->h_pool + pool_at, bl.pool.data(), bl.pool.size());
        const uint32_t base32 = (uint32_t)(pool_at >> 5);
        for (const PackedTask& p : bl.tasks) {
            DTask d = p.d;
            d.q_off32 += base32;
            if (!(d.want_dir & kTargetPac)) d.t_off32 += base32;
            d.mat_id = remap[(size_t)k][d.mat_id];
            b->h_tasks[at] = d;
            b->flags[(size_t)at] = p.flags; cls[(size_t)at] = p.cls; bin[(size_t)at] = p.bin;
            zsz[(size_t)at] = p.zsz; ctmpw[(size_t)at] = p.ctmpw;
            ++at;
        }
        pool_at += bl.pool.size();
    }
    if (layout_waves(b, cls.data(), bin.data(), zsz.data(), ctmpw.data())) return 1;
    if (alloc_device(b)) return 1;
    guard.ok = true;
    *out = b;
    return 0;
}

extern "C" int lb2_batch_upload(lb2_batch* b) {
    if (!b) return fail("batch is NULL");
    lb2_ctx* c = b->ctx;
    CU(cudaSetDevice(c->device));
    const int64_t n1 = std::max<int64_t>(b->n, 1);
    cudaStream_t s = c->copy;
    if (!b->raw_src) CU(cudaMemcpyAsync(b->d_pool, b->h_pool, b->pool_bytes, cudaMemcpyHostToDevice, s));
    else if (b->raw_bytes) CU(cudaMemcpyAsync(b->d_pool, b->raw_src, b->raw_bytes, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b->d_tasks, b->h_tasks, sizeof(DTask) * n1, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(b->d_order, b->h_order, sizeof(int32_t) * n1, cudaMemcpyHostToDevice, s));
    {
        int64_t any_long = 0;
        for (const auto& wv : b->waves) any_long += wv.long_count;
        if (any_long) CU(cudaMemcpyAsync(b->d_order + b->n, b->h_order + b->n, sizeof(int32_t) * n1, cudaMemcpyHostToDevice, s));
    }
    CU(cudaMemcpyAsync(b->d_mats, b->h_mats, sizeof(uint2) * kMaxMats * 8, cudaMemcpyHostToDevice, s));
    CU(cudaEventRecord(b->B.up_ev, s));
    b->h2d_bytes = (int64_t)(b->raw_src ? b->raw_bytes : b->pool_bytes) + (int64_t)(sizeof(DTask) + 4) * n1 + (int64_t)sizeof(uint2) * kMaxMats * 8;
    b->uploaded = true;
    return 0;
}

// Enqueue every kernel of the batch on the context's compute stream(s); no host sync.
static int compute_enqueue(lb2_batch* b) {
    if (!b) return fail("batch is NULL");
    if (!b->uploaded) return fail("lb2_batch_compute before lb2_batch_upload");
    lb2_ctx* c = b->ctx;
    CU(cudaSetDevice(c->device));
    cudaStream_t s = b->slot ? c->stream1 : c->stream;
    uint8_t* const d_z = b->slot ? c->d_z1 : c->d_z;
    int32_t* const d_ctmp = b->slot ? c->d_ctmp1 : c->d_ctmp;
    CU(cudaStreamWaitEvent(s, b->B.up_ev, 0));
    CU(cudaMemsetAsync(b->d_cursor, 0, sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(b->d_counters, 0, sizeof(unsigned int) * b->n_counters, s));
    CU(cudaMemsetAsync(b->d_err, 0, sizeof(int), s));
    b->launches = 0; b->fill_ms = 0; b->trace_ms = 0;
    for (auto& r : b->class_runs) { cudaEventDestroy(r.t0); cudaEventDestroy(r.t1); }
    b->class_runs.clear();
    CU(cudaEventRecord(b->ev[0], s));
    const bool one_wave = b->waves.size() == 1;
    for (auto& e : b->wave_ev) cudaEventDestroy(e);
    b->wave_ev.clear();
    static const int class_timing = env_int("LB2_CLASS_TIMING", 0);
    for (size_t wi = 0; wi < b->waves.size(); ++wi) {
        const Wave& wv = b->waves[wi];
        if (!one_wave) {       // per-wave (start, fills done, trace done) events, read back in compute_finish
            for (int q = 0; q < 3; ++q) { cudaEvent_t e; CU(cudaEventCreate(&e)); b->wave_ev.push_back(e); }
            CU(cudaEventRecord(b->wave_ev[wi * 3], s));
        }
        static const int multi = env_int("LB2_MULTI_STREAM", 1);
        const bool fan = multi && !class_timing && !b->class_timing;
        if (fan) {
            CU(cudaEventRecord(c->fork_ev, s));
            for (int k = 0; k < lb2_ctx::kAux; ++k) CU(cudaStreamWaitEvent(c->aux[k], c->fork_ev, 0));
        }
        int nlaunch = 0;
        // heaviest classes first (class ids grow with window size / lanes per tile)
        for (int k = kNumClass - 1; k >= 0; --k) {
            const int cnt = wv.cls_off[k + 1] - wv.cls_off[k];
            if (!cnt) continue;
            const int kind = class_kind(k), var = class_var(k), ls = class_logS(k);
            const int wpb = class_warps(var, ls);
            const size_t smem = (var == kVarGmem || var == kVarBlock) ? 0 : (size_t)wpb * var_warp_smem(var, 1 << ls);
            if (!c->occ[k]) {
                // first launch of this kernel through this context: opt in to the large dynamic shared memory, then
                // ask how many blocks fit (both load the kernel's module on demand)
                CU(cudaFuncSetAttribute(fill_table(kind, var), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
                int nb = 0;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fill_table(kind, var), wpb * 32, smem));
                c->occ[k] = nb > 0 ? nb : 1;
            }
            const int tpb = wpb * var_tasks_per_warp(var);
            int grid = (cnt + tpb - 1) / tpb;
            const int cap = c->sm_count * c->occ[k];
            if (grid > cap) grid = cap;
            if (var == kVarGmem) {                    // per-warp windows in global scratch (one class at a time)
                if (grid > c->sm_count) grid = c->sm_count;
                const size_t need = (size_t)grid * wpb * warp_smem_bytes(1 << ls);
                if (c->gwin_cap < need) {
                    CU(cudaStreamSynchronize(s));
                    for (int q = 0; q < lb2_ctx::kAux; ++q) CU(cudaStreamSynchronize(c->aux[q]));
                    if (c->d_gwin) CU(cudaFree(c->d_gwin));
                    c->d_gwin = nullptr; c->gwin_cap = 0;
                    CU(cudaMalloc(&c->d_gwin, need)); c->gwin_cap = need;
                }
            }
            cudaEvent_t t0 = nullptr, t1 = nullptr;
            if (class_timing || b->class_timing) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventRecord(t0, s); }
            // global-window classes share one scratch: keep them all on aux[0] (in order)
            cudaStream_t ls_ = fan ? c->aux[var == kVarGmem ? 0 : nlaunch++ % lb2_ctx::kAux] : s;
            fill_table(kind, var)<<<grid, wpb * 32, smem, ls_>>>(b->d_tasks, b->d_order + wv.first + wv.cls_off[k], cnt,
                                                             b->d_pool, c->d_pac, d_z, b->d_results, b->d_mats,
                                                             b->d_counters + wi * kNumClass + k, 1 << ls, c->d_gwin);
            if (b->class_timing && !class_timing) { cudaEventRecord(t1, s); b->class_runs.push_back({k, cnt, t0, t1, 0.f}); }
            if (class_timing) {
                cudaEventRecord(t1, s); cudaEventSynchronize(t1);
                float ms = 0; cudaEventElapsedTime(&ms, t0, t1);
                double est = 0;
                for (int q = 0; q < cnt; ++q) {
                    const DTask& d = b->h_tasks[b->h_order[wv.first + wv.cls_off[k] + q]];
                    est += (double)d.tlen * std::min<long>(d.qlen, 2L * d.w + 1);
                }
                fprintf(stderr, "[lb2] wave %zu kind %d var %d S %5d: %7d tasks grid %4d x %d warps occ %d  %8.3f ms  static cells %.3e  (%.1f Gcell/s static)\n",
                        wi, kind, var, 1 << ls, cnt, grid, wpb, c->occ[k], ms, est, est / ms / 1e6);
                cudaEventDestroy(t0); cudaEventDestroy(t1);
            }
            CU(cudaGetLastError());
            ++b->launches;
        }
        if (fan) {
            for (int k = 0; k < lb2_ctx::kAux; ++k) {
                CU(cudaEventRecord(c->join_ev[k], c->aux[k]));
                CU(cudaStreamWaitEvent(s, c->join_ev[k], 0));
            }
        }
        CU(cudaEventRecord(one_wave ? b->ev[2] : b->wave_ev[wi * 3 + 1], s));
        if (wv.ctmp_words) {
            static const int long_trace = env_int("LB2_LONG_TRACE", 1);
            const int nlong = long_trace ? wv.long_count : 0;
            if (nlong) {          // the long walks first: they are what the wave waits for
                trace_long_kernel<<<(nlong + kTraceWarps - 1) / kTraceWarps, kTraceWarps * 32, 0, s>>>(
                    b->d_tasks, b->d_order + b->n + wv.first, nlong, d_z, b->d_results, d_ctmp, b->d_cdense,
                    b->d_cursor, b->dense_cap, b->d_err);
                CU(cudaGetLastError());
                ++b->launches;
            }
            if (nlong < wv.count) {
                trace_kernel<<<(wv.count + 127) / 128, 128, 0, s>>>(b->d_tasks, b->d_order + wv.first, wv.count,
                                                                    d_z, b->d_results, d_ctmp, b->d_cdense,
                                                                    b->d_cursor, b->dense_cap, b->d_err, nlong ? 1 : 0);
                CU(cudaGetLastError());
                ++b->launches;
            }
        }
        if (!one_wave) CU(cudaEventRecord(b->wave_ev[wi * 3 + 2], s));
    }
    CU(cudaEventRecord(b->ev[3], s));
    CU(cudaEventRecord(b->B.done_ev, s));
    b->enqueued = true;
    return 0;
}

// Wait for the batch's kernels, read the event timings and the device error flag.
static int compute_finish(lb2_batch* b, float* kernel_ms) {
    if (!b || !b->enqueued) return fail("lb2_batch_compute: nothing enqueued");
    lb2_ctx* c = b->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(b->ev[3]));
    float total = 0, fill_acc = 0, trace_acc = 0;
    CU(cudaEventElapsedTime(&total, b->ev[0], b->ev[3]));
    if (b->waves.size() == 1) {
        CU(cudaEventElapsedTime(&fill_acc, b->ev[0], b->ev[2]));
        CU(cudaEventElapsedTime(&trace_acc, b->ev[2], b->ev[3]));
    } else {
        for (size_t wi = 0; wi * 3 + 2 < b->wave_ev.size(); ++wi) {
            float a = 0, t = 0;
            CU(cudaEventElapsedTime(&a, b->wave_ev[wi * 3], b->wave_ev[wi * 3 + 1]));
            CU(cudaEventElapsedTime(&t, b->wave_ev[wi * 3 + 1], b->wave_ev[wi * 3 + 2]));
            fill_acc += a; trace_acc += t;
        }
    }
    b->fill_ms = fill_acc; b->trace_ms = trace_acc;
    for (auto& r : b->class_runs) CU(cudaEventElapsedTime(&r.ms, r.t0, r.t1));
    if (kernel_ms) *kernel_ms = total;
    // read-backs go through the COPY stream: the compute stream may already hold the kernels of the next
    // chunk (lb2_dp_run pipelines chunks), and a copy queued behind them would stall the pipeline
    int err = 0;
    CU(cudaMemcpyAsync(&err, b->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->copy));
    CU(cudaStreamSynchronize(c->copy));
    if (err) return fail("traceback kernel reported CIGAR scratch overflow (code %d)", err);
    b->computed = true;
    b->enqueued = false;
    return 0;
}

extern "C" int lb2_batch_compute(lb2_batch* b, float* kernel_ms) {
    if (compute_enqueue(b)) return 1;
    return compute_finish(b, kernel_ms);
}

// split form of lb2_batch_compute for producers that overlap a batch with other work
extern "C" int lb2_batch_compute_async(lb2_batch* b) { return compute_enqueue(b); }
extern "C" int lb2_batch_compute_done(lb2_batch* b) {
    if (!b || !b->enqueued) return 1;
    return cudaEventQuery(b->ev[3]) == cudaSuccess ? 1 : 0;
}
extern "C" int lb2_batch_compute_wait(lb2_batch* b, float* kernel_ms) { return compute_finish(b, kernel_ms); }

static int cigar_capacity(int n) {      // capacity after the doubling pushes of src/ksw.c:506-516
    if (n == 0) return 0;
    int m = 4; while (m < n) m <<= 1; return m;
}

static int download_impl(lb2_batch* b, lb2_result* results, bool want_cigar, unsigned long long* used_out) {
    if (!b || (b->n && !results)) return fail("lb2_batch_download: NULL argument");
    if (!b->computed) return fail("lb2_batch_download before lb2_batch_compute");
    lb2_ctx* c = b->ctx;
    CU(cudaSetDevice(c->device));
    cudaStream_t s = c->copy;                // see compute_finish: never queue read-backs behind the next chunk's kernels
    unsigned long long used = 0;
    CU(cudaMemcpyAsync(&used, b->d_cursor, sizeof used, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(b->h_results, b->d_results, sizeof(DResult) * std::max<int64_t>(b->n, 1), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (want_cigar && used) {
        Buffers& B = b->B;
        if (B.h_cigar_cap < used) {
            const size_t cap = grown(used, B.h_cigar_cap, (size_t)1 << 18);
            cudaFreeHost(B.h_cigar); B.h_cigar = nullptr; B.h_cigar_cap = 0;
            CU(cudaMallocHost(&B.h_cigar, cap * sizeof(cigar32_t))); B.h_cigar_cap = cap;
        }
        CU(cudaMemcpyAsync(B.h_cigar, b->d_cdense, sizeof(cigar32_t) * used, cudaMemcpyDeviceToHost, s));
    }
    b->d2h_bytes = (int64_t)sizeof(DResult) * b->n + (want_cigar ? (int64_t)used * 4 : 0) + 8;
    const DResult* hr = b->h_results;
    const DTask* ht = b->h_tasks;
    const uint8_t* fl = b->flags.data();
    parallel_for(b->n, [=](int64_t a, int64_t e) {          // overlaps the CIGAR copy
        for (int64_t i = a; i < e; ++i) {
            const DResult& r = hr[i];
            lb2_result& o = results[i];
            o.score = r.score;
            if (ht[i].kind == kKindExtend) {
                if (fl[i] & LB2_FLAG_CIGAR) { o.qle = r.tk + 1; o.tle = r.ti + 1; }
                else { o.qle = r.max_j + 1; o.tle = r.max_i + 1; }
                o.gtle = r.max_ie + 1; o.gscore = r.gscore; o.max_off = r.max_off;
            } else {
                o.qle = ht[i].qlen; o.tle = ht[i].tlen; o.gtle = 0; o.gscore = 0; o.max_off = 0;
            }
            o.n_cigar = r.n_cigar; o.reserved = cigar_capacity(r.n_cigar);
            o.cigar_off = r.cigar_off; o.cells = r.cells;
        }
    });
    CU(cudaStreamSynchronize(s));
    *used_out = used;
    return 0;
}

extern "C" int lb2_batch_download(lb2_batch* b, lb2_result* results, cigar32_t** cigar_pool, int64_t* cigar_pool_n) {
    unsigned long long used = 0;
    if (download_impl(b, results, cigar_pool != nullptr, &used)) return 1;
    if (cigar_pool) {
        cigar32_t* pool = (cigar32_t*)malloc(sizeof(cigar32_t) * (used ? used : 1));
        if (!pool) return fail("out of host memory for %llu CIGAR words", used);
        const cigar32_t* src = b->B.h_cigar;
        parallel_for((int64_t)used, [=](int64_t a, int64_t e) { memcpy(pool + a, src + a, sizeof(cigar32_t) * (size_t)(e - a)); });
        *cigar_pool = pool;
    }
    if (cigar_pool_n) *cigar_pool_n = (int64_t)used;
    return 0;
}

// Same as lb2_batch_download, but the CIGAR pool is handed out in place (pinned
// staging owned by the batch; valid until the batch is destroyed or downloaded again).
extern "C" int lb2_batch_download_view(lb2_batch* b, lb2_result* results, const cigar32_t** cigar_pool, int64_t* cigar_pool_n) {
    unsigned long long used = 0;
    if (download_impl(b, results, cigar_pool != nullptr, &used)) return 1;
    if (cigar_pool) *cigar_pool = b->B.h_cigar;
    if (cigar_pool_n) *cigar_pool_n = (int64_t)used;
    return 0;
}

extern "C" int lb2_batch_stats(const lb2_batch* b, int64_t* h2d, int64_t* d2h, int64_t* launches,
                               float* fill_ms, float* trace_ms) {
    if (!b) return fail("batch is NULL");
    if (h2d) *h2d = b->h2d_bytes;
    if (d2h) *d2h = b->d2h_bytes;
    if (launches) *launches = b->launches;
    if (fill_ms) *fill_ms = b->fill_ms;
    if (trace_ms) *trace_ms = b->trace_ms;
    return 0;
}

static const char* kernel_name(int kind, int var) {
    static const char* g[] = {"fill_kernel<1,global>", "fill_kernel<2,global>", "fill_kernel<4,global>", "fill16_kernel<2,global>", "fill16_kernel<4,global>",
                              "fill_kernel<4,global,gmem window>", "fill16d_kernel<2,global,16>", "fill16d_kernel<2,global,8>", "fill16d_kernel<4,global,8>",
                              "fill16d_kernel<4,global,16>", "fill_lean_kernel"};
    static const char* e[] = {"fill_kernel<1,extend>", "fill_kernel<2,extend>", "fill_kernel<4,extend>", "fill16_kernel<2,extend>", "fill16_kernel<4,extend>",
                              "fill_kernel<4,extend,gmem window>", "fill16d_kernel<2,extend,16>", "fill16d_kernel<2,extend,8>", "fill16d_kernel<4,extend,8>",
                              "fill16d_kernel<4,extend,16>", "fill_lean_kernel<extend>"};
    return (kind == kKindGlobal ? g : e)[var];
}

extern "C" int lb2_batch_set_class_timing(lb2_batch* b, int on) {
    if (!b) return fail("batch is NULL");
    b->class_timing = on != 0;
    return 0;
}

// Per launch class of the last lb2_batch_compute with class timing on (and after a download, which brings the cell
// counts back): the kernel, its tasks, the DP cells it evaluated and its own CUDA-event time, launched ALONE.
extern "C" int lb2_batch_class_stats(const lb2_batch* b, lb2_class_stat* out, int cap) {
    if (!b || (cap > 0 && !out)) return -1;
    int n = 0;
    // classes may appear once per wave: merge by class id
    for (const auto& r : b->class_runs) {
        int at = -1;
        for (int q = 0; q < n; ++q) if (out[q].class_id == r.cls) at = q;
        if (at < 0) {
            if (n == cap) break;
            at = n++;
            memset(&out[at], 0, sizeof out[at]);
            out[at].class_id = r.cls; out[at].kind = class_kind(r.cls); out[at].variant = class_var(r.cls);
            out[at].window_slots = 1 << class_logS(r.cls);
            snprintf(out[at].kernel, sizeof out[at].kernel, "%s", kernel_name(class_kind(r.cls), class_var(r.cls)));
        }
        out[at].tasks += r.tasks; out[at].ms += r.ms;
    }
    if (b->computed || true)
        for (int64_t i = 0; i < b->n; ++i)
            for (int q = 0; q < n; ++q)
                if (out[q].class_id == b->cls[(size_t)i]) { out[q].cells += b->h_results[i].cells; break; }
    return n;
}

// One-shot run.  Large batches are cut into chunks and pipelined: while the GPU
// runs chunk k, the host packs chunk k+1 and its H2D copy goes out on the copy
// stream; results and CIGAR words are appended in task order.
static int dp_run_impl(lb2_ctx* ctx, const uint8_t* upool, int64_t upool_bytes, int64_t n, const lb2_task* tasks, lb2_result* results,
                       cigar32_t** cigar_pool, int64_t* cigar_pool_n) {
    if (!ctx) return fail("ctx is NULL");
    static const int64_t chunk_env = env_int("LB2_CHUNK_TASKS", 0);
    const int64_t chunk_min = chunk_env > 0 ? chunk_env : ctx->chunk_tasks;
    int64_t K = chunk_min > 0 ? n / chunk_min : 1;
    K = std::max<int64_t>(1, std::min<int64_t>(K, 16));
    // chunk boundaries: the first chunk is a quarter of the others, so the GPU starts early (the pipeline's fill time is
    // the host work + H2D of chunk 0), and the last one half, so little is left to read back once the kernels are done
    std::vector<double> wgt((size_t)K, 1.0);
    if (K >= 2) wgt[0] = 0.25;
    if (K >= 3) wgt[(size_t)K - 1] = 0.5;
    double wsum = 0; for (double x : wgt) wsum += x;
    std::vector<int64_t> cut((size_t)K + 1, 0);
    { double acc = 0; for (int64_t k = 0; k < K; ++k) { acc += wgt[(size_t)k]; cut[(size_t)k + 1] = (int64_t)((double)n * acc / wsum); } }
    cut[(size_t)K] = n;
    ctx->run_h2d = ctx->run_d2h = ctx->run_launches = 0;
    ctx->run_fill_ms = ctx->run_trace_ms = 0;
    cigar32_t* pool = nullptr; int64_t pool_n = 0, pool_cap = 0;
    struct InFlight { lb2_batch* b; int64_t lo; };
    std::vector<InFlight> fly;          // enqueued, not yet drained (oldest first)
    int rc = 0;
    auto drain = [&](lb2_batch* b, int64_t lo) -> int {       // finish + download chunk starting at task `lo`
        if (compute_finish(b, nullptr)) return 1;
        const cigar32_t* view = nullptr; int64_t used = 0;
        if (lb2_batch_download_view(b, results + lo, cigar_pool ? &view : nullptr, &used)) return 1;
        ctx->run_h2d += b->h2d_bytes; ctx->run_d2h += b->d2h_bytes; ctx->run_launches += b->launches;
        ctx->run_fill_ms += b->fill_ms; ctx->run_trace_ms += b->trace_ms;
        if (cigar_pool) {
            if (pool_n + used > pool_cap) {
                pool_cap = std::max<int64_t>((pool_n + used) * (K > 1 ? 2 : 1), 1024);
                cigar32_t* np = (cigar32_t*)realloc(pool, sizeof(cigar32_t) * (size_t)pool_cap);
                if (!np) return fail("out of host memory for the CIGAR pool");
                pool = np;
            }
            cigar32_t* dst = pool + pool_n;
            parallel_for(used, [=](int64_t a, int64_t e) { memcpy(dst + a, view + a, sizeof(cigar32_t) * (size_t)(e - a)); });
            if (pool_n) {
                lb2_result* r = results + lo; const int64_t cnt = b->n, base = pool_n;
                parallel_for(cnt, [=](int64_t a, int64_t e) { for (int64_t i = a; i < e; ++i) r[i].cigar_off += base; });
            }
        }
        pool_n += used;
        return 0;
    };
    static const int trace = env_int("LB2_RUN_TRACE", 0), overlap = env_int("LB2_RUN_OVERLAP", 1);
    // chunks kept enqueued while an older one is read back: with two, the GPU always has the next chunk's kernels queued
    // when a chunk ends (chunk k and k+2 share a scratch set and a compute stream, so they are ordered by the stream)
    const size_t ahead = overlap ? 2 : 1;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    for (int64_t k = 0; k < K && !rc; ++k) {
        const int64_t lo = cut[(size_t)k], hi = cut[(size_t)k + 1];
        lb2_batch* b = nullptr;
        const double t0 = now();
        if (batch_create_impl(ctx, hi - lo, tasks + lo, upool, upool_bytes, &b, overlap ? (int)(k & 1) : 0)) { rc = 1; break; }      // host packing
        const double t1 = now();
        if (lb2_batch_upload(b) || compute_enqueue(b)) { lb2_batch_destroy(b); rc = 1; break; }
        const double t2 = now();
        fly.push_back({b, lo});
        if (fly.size() > ahead) { rc = drain(fly.front().b, fly.front().lo); lb2_batch_destroy(fly.front().b); fly.erase(fly.begin()); }
        if (trace) fprintf(stderr, "[lb2 run] chunk %lld: %lld tasks; at %.1f ms create %.1f, upload+enqueue %.1f, drain of an earlier one %.1f ms\n",
                           (long long)k, (long long)(hi - lo), (t0 - t_begin) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (now() - t2) * 1e3);
    }
    const double t_last = now();
    for (auto& f : fly) { if (!rc) rc = drain(f.b, f.lo); lb2_batch_destroy(f.b); }
    fly.clear();
    if (trace) fprintf(stderr, "[lb2 run] last drains %.1f ms; total %.1f ms; kernels fill %.1f + trace %.1f ms (sums of per-chunk event times: chunks overlap)\n", (now() - t_last) * 1e3, (now() - t_begin) * 1e3, ctx->run_fill_ms, ctx->run_trace_ms);
    if (rc) { free(pool); return 1; }
    if (cigar_pool) *cigar_pool = pool ? pool : (cigar32_t*)malloc(sizeof(cigar32_t));
    if (cigar_pool_n) *cigar_pool_n = pool_n;
    return 0;
}

// Copies the sequences of `tasks` into `pool` in task order (query, then target; 16-byte granular) and re-points
// the records at the copies: the layout under which every chunk of lb2_dp_run_pool uploads one tight range.
extern "C" int lb2_pool_pack(int64_t n, lb2_task* tasks, uint8_t* pool, int64_t pool_bytes, int64_t* used) {
    if (n < 0 || (n > 0 && (!tasks || !pool))) return fail("lb2_pool_pack: NULL argument");
    std::vector<int64_t> off((size_t)n + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (tasks[i].qlen < 0 || tasks[i].tlen < 0) return fail("lb2_pool_pack: task %lld has a negative length", (long long)i);
        const int64_t t = (tasks[i].flags & LB2_FLAG_TARGET_PAC) ? 0 : tasks[i].tlen;
        off[(size_t)i + 1] = off[(size_t)i] + (((int64_t)tasks[i].qlen + t + 15) & ~(int64_t)15);
    }
    if (used) *used = off[(size_t)n];
    if (off[(size_t)n] > pool_bytes) return fail("lb2_pool_pack: pool of %lld bytes, %lld needed", (long long)pool_bytes, (long long)off[(size_t)n]);
    parallel_for(n, [&](int64_t a, int64_t e) {
        for (int64_t i = a; i < e; ++i) {
            lb2_task& t = tasks[i];
            uint8_t* q = pool + off[(size_t)i];
            if (t.qlen) memcpy(q, t.query, (size_t)t.qlen);
            t.query = q;
            if (!(t.flags & LB2_FLAG_TARGET_PAC)) {
                if (t.tlen) memcpy(q + t.qlen, t.target, (size_t)t.tlen);
                t.target = q + t.qlen;
            }
        }
    });
    return 0;
}

extern "C" int lb2_dp_run(lb2_ctx* ctx, int64_t n, const lb2_task* tasks, lb2_result* results,
                          cigar32_t** cigar_pool, int64_t* cigar_pool_n) {
    return dp_run_impl(ctx, nullptr, 0, n, tasks, results, cigar_pool, cigar_pool_n);
}
extern "C" int lb2_dp_run_pool(lb2_ctx* ctx, const uint8_t* pool, int64_t pool_bytes, int64_t n, const lb2_task* tasks,
                               lb2_result* results, cigar32_t** cigar_pool, int64_t* cigar_pool_n) {
    if (!pool || pool_bytes < 0) return fail("lb2_dp_run_pool: NULL pool");
    return dp_run_impl(ctx, pool, pool_bytes, n, tasks, results, cigar_pool, cigar_pool_n);
}

extern "C" int lb2_ctx_last_run_stats(const lb2_ctx* ctx, int64_t* h2d, int64_t* d2h, int64_t* launches) {
    if (!ctx) return fail("ctx is NULL");
    if (h2d) *h2d = ctx->run_h2d;
    if (d2h) *d2h = ctx->run_d2h;
    if (launches) *launches = ctx->run_launches;
    return 0;
}

extern "C" int lb2_ctx_last_run_kernel_ms(const lb2_ctx* ctx, float* fill_ms, float* trace_ms) {
    if (!ctx) return fail("ctx is NULL");
    if (fill_ms) *fill_ms = ctx->run_fill_ms;
    if (trace_ms) *trace_ms = ctx->run_trace_ms;
    return 0;
}

// ---- alignment record statistics (include/lamsa_b200.h section 5, aux_scan.cuh) ----------------------------------
extern "C" int lb2_aux_run(lb2_ctx* c, int64_t n, const lb2_aux_task* tasks, lb2_aux_result* results) {
    if (!c || n < 0 || (n > 0 && (!tasks || !results))) return fail("lb2_aux_run: bad argument");
    if (!c->d_pac) return fail("lb2_aux_run: no resident reference (lb2_ctx_set_reference)");
    if (n == 0) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    // one staging block: [AuxRec n][cigar words][read bytes][results n], each part 16-byte aligned
    size_t words = 0, bytes = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (tasks[i].n_cigar < 0 || tasks[i].read_len < 0 || (tasks[i].n_cigar && !tasks[i].cigar) || (tasks[i].read_len && !tasks[i].read))
            return fail("lb2_aux_run: task %lld is malformed", (long long)i);
        words += (size_t)tasks[i].n_cigar; bytes += ((size_t)tasks[i].read_len + 15) & ~(size_t)15;
    }
    auto up = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_rec = 0, o_cig = up(sizeof(AuxRec) * (size_t)n), o_rd = o_cig + up(words * 4), o_res = o_rd + up(bytes),
                 total = o_res + sizeof(lb2_aux_result) * (size_t)n;
    if (c->aux_cap < total) {
        const size_t cap = grown(total, c->aux_cap, (size_t)1 << 20);
        CU(cudaStreamSynchronize(c->stream));
        cudaFreeHost(c->aux_h); cudaFree(c->aux_d); c->aux_h = nullptr; c->aux_d = nullptr; c->aux_cap = 0;
        CU(cudaMallocHost(&c->aux_h, cap)); CU(cudaMalloc(&c->aux_d, cap)); c->aux_cap = cap;
    }
    AuxRec* rec = reinterpret_cast<AuxRec*>(c->aux_h + o_rec);
    int32_t* cg = reinterpret_cast<int32_t*>(c->aux_h + o_cig);
    uint8_t* rd = c->aux_h + o_rd;
    size_t w = 0, b = 0;
    for (int64_t i = 0; i < n; ++i) {
        const lb2_aux_task& t = tasks[i];
        if (t.ref_pac < 0) return fail("lb2_aux_run: task %lld: negative reference coordinate", (long long)i);
        rec[i] = AuxRec{(uint32_t)w, (uint32_t)t.n_cigar, (uint32_t)b, (uint32_t)t.read_len, (uint64_t)t.ref_pac};
        if (t.n_cigar) memcpy(cg + w, t.cigar, (size_t)t.n_cigar * 4);
        if (t.read_len) memcpy(rd + b, t.read, (size_t)t.read_len);
        w += (size_t)t.n_cigar; b += ((size_t)t.read_len + 15) & ~(size_t)15;
    }
    cudaStream_t s = c->stream;
    CU(cudaMemcpyAsync(c->aux_d, c->aux_h, o_res, cudaMemcpyHostToDevice, s));
    const int wpb = 4;
    aux_scan_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, s>>>(
        reinterpret_cast<const AuxRec*>(c->aux_d + o_rec), (int)n, reinterpret_cast<const int32_t*>(c->aux_d + o_cig),
        c->aux_d + o_rd, c->d_pac, (long long)c->l_pac, reinterpret_cast<lb2_aux_result*>(c->aux_d + o_res));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(c->aux_h + o_res, c->aux_d + o_res, sizeof(lb2_aux_result) * (size_t)n, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(results, c->aux_h + o_res, sizeof(lb2_aux_result) * (size_t)n);
    return 0;
}

extern "C" int lb2_int_peak(lb2_ctx* ctx, double* gops_s16x2, double* gops_s32, int* sm_count, int* clock_khz) {
    if (!ctx) return fail("ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, ctx->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    if (clock_khz) *clock_khz = khz;
    double a = 0, c = 0;
    if (int_peak_measure(ctx->stream, prop.multiProcessorCount, &a, &c)) return fail("int_peak: %s", cudaGetErrorString(cudaGetLastError()));
    if (gops_s16x2) *gops_s16x2 = a;
    if (gops_s32) *gops_s32 = c;
    return 0;
}

// ---- internal accessors for the other translation units (ctx_internal.h) ----
namespace lb2 {
int ctx_reserve(lb2_ctx* ctx, int64_t n_tasks, size_t pool_bytes, size_t z_bytes, size_t cigar_words) {
    if (!ctx) return fail("ctx_reserve: ctx is NULL");
    CU(cudaSetDevice(ctx->device));
    lb2_batch* b = new lb2_batch();
    b->ctx = ctx; b->n = n_tasks;
    struct Guard { lb2_batch* b; ~Guard() { lb2_batch_destroy(b); } } guard{b};
    // take no parked set: this call is there to CREATE one
    Buffers keep[lb2_ctx::kParked];
    { std::lock_guard<std::mutex> lk(ctx->mu); for (int k = 0; k < lb2_ctx::kParked; ++k) { keep[k] = ctx->parked[k]; ctx->parked[k] = Buffers(); } }
    int rc = alloc_host(b, pool_bytes);
    if (!rc) {
        Wave wv; memset(&wv, 0, sizeof wv);
        wv.z_bytes = z_bytes; wv.ctmp_words = cigar_words;
        b->waves.push_back(wv);
        b->dense_cap = cigar_words + 16;
        rc = alloc_device(b);
        b->waves.clear();
    }
    if (!rc && b->B.h_cigar_cap < cigar_words) {
        cudaFreeHost(b->B.h_cigar); b->B.h_cigar = nullptr; b->B.h_cigar_cap = 0;
        if (cudaMallocHost(&b->B.h_cigar, cigar_words * sizeof(cigar32_t)) != cudaSuccess) rc = fail("ctx_reserve: pinned CIGAR buffer");
        else b->B.h_cigar_cap = cigar_words;
    }
    { std::lock_guard<std::mutex> lk(ctx->mu); for (int k = 0; k < lb2_ctx::kParked; ++k) if (keep[k].valid) ctx->parked[k] = keep[k]; }
    return rc;          // the guard parks the new set in a free place (or releases it when both are taken)
}
int batch_wait_blocking(lb2_batch* b) {
    if (!b || !b->enqueued) return fail("batch_wait_blocking: nothing enqueued");
    CU(cudaSetDevice(b->ctx->device));
    CU(cudaEventSynchronize(b->B.done_ev));
    return 0;
}
cudaStream_t ctx_stream(lb2_ctx* c) { return c->stream; }
int ctx_device(lb2_ctx* c) { return c->device; }
int ctx_sm_count(lb2_ctx* c) { return c->sm_count; }
int set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return 1;
}
}
