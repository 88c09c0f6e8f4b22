// producer.cu -- the batch producer: the per-read host control flow of LAMSA's aligner (the worker
// loop of src/lamsa_aln.c:825-891, restated by lamsa_b200/host/aln_core.c) runs as thousands of
// user-level fibers on the host cores, and every banded-DP / chaining call a fiber makes is parked,
// classified and copied by the thread that runs it, and served by ONE submitter per GPU:
//
//   worker threads (one per host core)      device (one per GPU)
//   ----------------------------------      -------------------------------------------------------
//   run fibers until they park        --->  pending groups (tasks already packed: dp_pack.h TaskBlob)
//   hand the parked requests over            submitter thread: concatenates what is pending into one
//   as a packed group                        batch per free slot (lb2::batch_create_staged), H2D, launch
//   <--- inbox: fibers whose requests        one completer thread per slot: waits for the batch's event,
//        were served                         reads results + CIGARs back, writes them into the callers'
//                                            out-pointers, returns the fibers to their home threads
//                                            chaining thread: parked frag_line_BCC / _remain requests of
//                                            all threads as one batch per stage
//
// A launch therefore carries the requests of ALL threads that parked since the last launch (thousands
// when tens of thousands of reads are in flight), batches overlap (slot k+1 is packed and launched while
// slot k runs; long tasks travel in a slot of their own so that short ones are not held back by them),
// and nothing but these few device threads ever touches CUDA.  Dependent calls of one read
// (merge_cigar loops, the three stages of ksw_bi_extend, the read-level stages) simply take several
// rounds while the other reads keep the batches full; results are bit-identical because every request
// is served by the same kernels as a batch of one.
//
// lb2_worker_spawn / lb2_worker_join have the signatures of pthread_create / pthread_join.
#include <pthread.h>
#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ctx_internal.h"
#include "dp_pack.h"
#include "dropin_internal.h"
namespace lb2 { extern std::atomic<long long> g_sdp_ns[5]; }      // sdp_dropin.cu: where the chaining batches' time goes

namespace {

using Clock = std::chrono::steady_clock;
double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

#if defined(__x86_64__) && !defined(LB2_FIBER_UCONTEXT)
#define LB2_FAST_SWITCH 1
extern "C" void lb2_fiber_swap(void** save_sp, void* load_sp);      // fiber_switch.cpp
extern "C" void lb2_fiber_entry_thunk(void);
#else
#define LB2_FAST_SWITCH 0
#endif

int env_i(const char* name, int dflt) { const char* e = getenv(name); return e && *e ? atoi(e) : dflt; }
bool verbose() { static const bool v = getenv("LB2_FIBER_STATS") != nullptr; return v; }

// LB2_TRACE=<file>: one line per device-thread event (seconds since the first event, thread, what, slot, tasks)
struct Trace {
    FILE* f = nullptr; std::mutex mu; Clock::time_point t0;
    Trace() { const char* e = getenv("LB2_TRACE"); if (e && *e) { f = fopen(e, "w"); t0 = Clock::now(); } }
    void ev(const char* who, const char* what, int slot, long n) {
        if (!f) return;
        const double t = secs(t0, Clock::now());
        std::lock_guard<std::mutex> lk(mu);
        fprintf(f, "%.6f %s %s %d %ld\n", t, who, what, slot, n);
    }
    ~Trace() { if (f) fclose(f); }
};
Trace& trace() { static Trace* t = new Trace(); return *t; }

struct Worker;
struct Fiber {
#if LB2_FAST_SWITCH
    void* sp = nullptr;                  // saved stack pointer while the fiber is not running
#else
    ucontext_t ctx;
#endif
    void* map = nullptr; size_t map_bytes = 0;      // the mapping: one guard page + the stack
    void* (*fn)(void*) = nullptr; void* arg = nullptr;
    Worker* home = nullptr;
    bool done = false;
    double parked_s = 0;                 // total time this worker spent parked on requests
};

// requests one worker thread parked since its last hand-over, packed for the device
struct DpGroup {
    Clock::time_point t_first;            // when the first request was parked
    lb2::TaskBlob blob;
    std::vector<lb2::DpRequest*> reqs;
    std::vector<Fiber*> owners;
};
struct AuxRequest { int64_t n; const lb2_aux_task* tasks; lb2_aux_result* results; lb2_hash_task* hash; };   // hash != nullptr: one lb2_hash_line_run request
struct AuxGroup {
    Clock::time_point t_first;
    std::vector<AuxRequest*> reqs;
    std::vector<Fiber*> owners;
};
struct SdpGroup {
    Clock::time_point t_first;
    std::vector<lb2::SdpRequest*> reqs;
    std::vector<Fiber*> owners;
};

struct Device;
struct Worker {                          // one OS thread
    int index = 0;
    Device* dev = nullptr;
#if LB2_FAST_SWITCH
    void* main_sp = nullptr;
#else
    ucontext_t main;
#endif
    Fiber* cur = nullptr;
    std::deque<Fiber*> resumed, fresh;   // run queue: fibers with a served request first, then fibers not started yet
    std::mutex in_mu; std::condition_variable in_cv; std::vector<Fiber*> inbox;
    std::atomic<int> inbox_n{0};
    DpGroup* fast = nullptr; DpGroup* slow = nullptr; SdpGroup* sdp = nullptr; AuxGroup* aux = nullptr;
    size_t live = 0;
    int64_t switches = 0, handovers = 0, dp_n = 0, sdp_n = 0;
    double dev_wait_s = 0, idle_s = 0, fiber_s = 0, pack_s = 0, dp_lat_s = 0, sdp_lat_s = 0, dp_lat_max = 0;
};
thread_local Worker* tl_worker = nullptr;

// ---- the device side ---------------------------------------------------------------------------
// batch slots: the first kFastSlots serve ordinary tasks, the others tasks of more than LB2_FAST_ROWS rows (a batch lasts
// as long as its longest task; long tasks are a tenth of the requests of SV reads but, with a single slot, were two
// thirds of the time a read spent waiting)
constexpr int kFastSlots = 3, kSlowSlots = 3, kSlots = kFastSlots + kSlowSlots;
// chaining batches in flight at once: a batch lasts about as long as its slowest read (one warp walks a read), so a
// request that arrives while one batch runs should not have to wait for it
constexpr int kSdpThreads = 3;
struct Slot {
    lb2_ctx* ctx = nullptr;
    lb2_batch* batch = nullptr;
    std::vector<DpGroup*> groups;
    bool busy = true, launched = false;         // busy until the slot's thread has sized its buffers
    std::thread completer;
};
struct Device {
    int device = 0;
    bool loopback = false;                       // self test: no GPU, requests are handed straight back
    lb2_ctx* sdp_ctx[kSdpThreads] = {nullptr};
    lb2_ctx* aux_ctx = nullptr;                  // record statistics (lb2_aux_run): holds the resident reference
    std::mutex mu;
    std::condition_variable cv_submit, cv_slot, cv_sdp, cv_aux;
    std::vector<DpGroup*> pend_fast, pend_slow;
    int64_t pend_fast_tasks = 0, pend_slow_tasks = 0;
    Clock::time_point pend_fast_since;           // arrival of the oldest pending group
    std::vector<SdpGroup*> pend_sdp;
    std::vector<AuxGroup*> pend_aux;
    Slot slot[kSlots];
    bool stop = false;
    std::thread submitter, sdp_thread[kSdpThreads], aux_thread;
    // statistics (under mu)
    std::vector<int> batch_tasks;
    int64_t dp_tasks = 0, slow_batches = 0, slow_tasks = 0, sdp_reqs = 0, sdp_batches = 0, hash_n = 0, hash_batches = 0;
    double pack_s = 0, deliver_s = 0, sdp_s = 0, kernel_ms = 0;
};

void route_home(std::vector<Fiber*>& owners) {
    // group by home thread: one lock and one wake-up per thread
    std::sort(owners.begin(), owners.end(), [](Fiber* a, Fiber* b) { return a->home < b->home; });
    for (size_t i = 0; i < owners.size();) {
        Worker* w = owners[i]->home;
        size_t j = i;
        while (j < owners.size() && owners[j]->home == w) ++j;
        {
            std::lock_guard<std::mutex> lk(w->in_mu);
            w->inbox.insert(w->inbox.end(), owners.begin() + (long)i, owners.begin() + (long)j);
            w->inbox_n.store((int)w->inbox.size(), std::memory_order_release);
        }
        w->in_cv.notify_one();
        i = j;
    }
}

[[noreturn]] void die(const char* what) {
    fprintf(stderr, "[lamsa_b200] %s: %s\n", what, lb2_last_error());
    exit(1);
}

// results of a finished batch -> the callers' out-pointers (CIGARs malloc'd with the capacity the reference's
// doubling pushes would have reached, src/ksw.c:506-516)
void deliver_slot(Device* d, Slot& s) {
    const auto t0 = Clock::now();
    int64_t n = 0;
    for (DpGroup* g : s.groups) n += (int64_t)g->reqs.size();
    std::vector<lb2_result> results((size_t)n);
    const cigar32_t* pool = nullptr; int64_t pn = 0;
    float ms = 0;
    const int slot_k = (int)(&s - d->slot);
    if (lb2::batch_wait_blocking(s.batch)) die("DP batch failed");
    trace().ev("completer", "kernels_done", slot_k, (long)n);
    if (lb2_batch_compute_wait(s.batch, &ms) || lb2_batch_download_view(s.batch, results.data(), &pool, &pn)) die("DP batch failed");
    trace().ev("completer", "downloaded", slot_k, (long)n);
    int64_t at = 0;
    std::vector<Fiber*> owners;
    owners.reserve((size_t)n);
    for (DpGroup* g : s.groups) {
        for (size_t i = 0; i < g->reqs.size(); ++i, ++at) {
            lb2::DpRequest* p = g->reqs[i];
            const lb2_result& r = results[(size_t)at];
            *p->res = r;
            if (p->cig) {
                if (r.n_cigar > 0) {
                    cigar32_t* out = (cigar32_t*)malloc(sizeof(cigar32_t) * (size_t)r.reserved);
                    memcpy(out, pool + r.cigar_off, sizeof(cigar32_t) * (size_t)r.n_cigar);
                    *p->cig = out;
                } else *p->cig = nullptr;
            }
        }
        owners.insert(owners.end(), g->owners.begin(), g->owners.end());
        delete g;
    }
    s.groups.clear();
    lb2_batch_destroy(s.batch);
    s.batch = nullptr;
    route_home(owners);
    trace().ev("completer", "delivered", slot_k, (long)n);
    std::lock_guard<std::mutex> lk(d->mu);
    d->deliver_s += secs(t0, Clock::now());
    d->kernel_ms += ms;
}

void completer_main(Device* d, int k) {
    Slot& s = d->slot[k];
    cudaSetDevice(d->device);
    {
        // steady-state sizes up front (two sets: one batch being packed while the previous one is read back): growing
        // pinned or device buffers later would stall every slot for milliseconds.  Each slot's thread does its own.
        const bool slow = k >= kFastSlots;
        for (int r = 0; r < 2; ++r)
            if (lb2::ctx_reserve(s.ctx, slow ? 4096 : 16384, (size_t)(slow ? 32 : 16) << 20,
                                 (size_t)(slow ? 1024 : 256) << 20, (size_t)(slow ? 8 : 4) << 20)) die("cannot size the batch buffers");
        { std::lock_guard<std::mutex> lk(d->mu); s.busy = false; }
        d->cv_submit.notify_one();
    }
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv_slot.wait(lk, [&] { return d->stop || s.launched; });
            if (!s.launched) return;
        }
        deliver_slot(d, s);
        {
            std::lock_guard<std::mutex> lk(d->mu);
            s.launched = false; s.busy = false;
        }
        d->cv_submit.notify_one();
    }
}

// Gather policy.  A free slot takes everything that is pending at once when the GPU has nothing else to do
// (latency first); while other batches are in flight it waits until LB2_MIN_BATCH tasks have gathered or the
// oldest pending group is LB2_GATHER_US old (throughput: fewer, larger launches).
void submitter_main(Device* d) {
    cudaSetDevice(d->device);
    static const int min_batch = env_i("LB2_MIN_BATCH", 2048), gather_us = env_i("LB2_GATHER_US", 300);
    for (;;) {
        std::vector<DpGroup*> take;
        int k = -1;
        bool slow = false;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            for (;;) {
                if (d->stop) return;
                int free_fast = -1, busy_fast = 0;
                for (int q = 0; q < kFastSlots; ++q) { if (!d->slot[q].busy) { if (free_fast < 0) free_fast = q; } else ++busy_fast; }
                int free_slow = -1;
                for (int q = kFastSlots; q < kSlots; ++q) if (!d->slot[q].busy) { free_slow = q; break; }
                if (!d->pend_slow.empty() && free_slow >= 0) { k = free_slow; slow = true; break; }
                if (!d->pend_fast.empty() && free_fast >= 0) {
                    const auto due = d->pend_fast_since + std::chrono::microseconds(gather_us);
                    if (busy_fast == 0 || d->pend_fast_tasks >= min_batch || Clock::now() >= due) { k = free_fast; break; }
                    // something is running: give the other threads a moment to add to this launch
                    d->cv_submit.wait_until(lk, due);
                    continue;
                }
                d->cv_submit.wait(lk);
            }
            if (slow) { take.swap(d->pend_slow); d->slow_tasks += d->pend_slow_tasks; d->pend_slow_tasks = 0; ++d->slow_batches; }
            else { take.swap(d->pend_fast); d->batch_tasks.push_back((int)d->pend_fast_tasks); d->dp_tasks += d->pend_fast_tasks; d->pend_fast_tasks = 0; }
            d->slot[k].busy = true;
        }
        const auto t0 = Clock::now();
        Slot& s = d->slot[k];
        { long nt = 0; for (DpGroup* g : take) nt += (long)g->reqs.size(); trace().ev("submitter", "pack_begin", k, nt); }
        std::vector<const lb2::TaskBlob*> blobs;
        for (DpGroup* g : take) blobs.push_back(&g->blob);
        if (lb2::batch_create_staged(s.ctx, blobs.data(), (int)blobs.size(), &s.batch) || lb2_batch_upload(s.batch) ||
            lb2_batch_compute_async(s.batch)) die("DP launch failed");
        s.groups.swap(take);
        trace().ev("submitter", "launched", k, 0);
        {
            std::lock_guard<std::mutex> lk(d->mu);
            s.launched = true;
            d->pack_s += secs(t0, Clock::now());
        }
        d->cv_slot.notify_all();
    }
}

void sdp_main(Device* d, int k) {
    cudaSetDevice(d->device);
    lb2::dropin_bind_thread_ctx(d->sdp_ctx[k]);
    for (;;) {
        std::vector<SdpGroup*> take;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv_sdp.wait(lk, [&] { return d->stop || !d->pend_sdp.empty(); });
            if (d->pend_sdp.empty()) return;
            // a batch costs a dozen driver calls whatever its size: give the other worker threads a moment to add theirs
            static const int min_reads = env_i("LB2_SDP_MIN_BATCH", 64), gather_us = env_i("LB2_SDP_GATHER_US", 200);
            if (gather_us > 0) {
                const auto due = Clock::now() + std::chrono::microseconds(gather_us);
                for (;;) {
                    size_t have = 0;
                    for (SdpGroup* g : d->pend_sdp) have += g->reqs.size();
                    if (d->stop || have >= (size_t)min_reads || d->pend_sdp.empty()) break;
                    if (d->cv_sdp.wait_until(lk, due) == std::cv_status::timeout) break;
                }
                if (d->pend_sdp.empty()) continue;          // another chaining thread took them
            }
            take.swap(d->pend_sdp);
        }
        const auto t0 = Clock::now();
        int64_t nreq = 0; int nb = 0;
        { long nt = 0; for (SdpGroup* g : take) nt += (long)g->reqs.size(); trace().ev("chaining", "begin", 0, nt); }
        // One batch per stage (at most LB2_SDP_MAX_BATCH reads).  The chaining kernel walks a read with one warp, so a
        // launch lasts about as long as its slowest read whatever the number of reads: large batches are the fast
        // ones (measured on 10 kbp PacBio-mode reads: a request waits 30-50 ms uncapped, 250 ms with 512-read batches).
        static const size_t max_batch = (size_t)env_i("LB2_SDP_MAX_BATCH", 1 << 20);
        // (both stages are sorted out before anything is served: a served request lives on its owner's stack and is
        // gone as soon as the owner runs again)
        std::vector<lb2::SdpRequest*> grp[2]; std::vector<Fiber*> own[2];
        for (SdpGroup* g : take)
            for (size_t i = 0; i < g->reqs.size(); ++i) {
                const int s2 = g->reqs[i]->stage == 2 ? 0 : 1;           // stage 2 first: those reads are closer to their end
                grp[s2].push_back(g->reqs[i]); own[s2].push_back(g->owners[i]);
            }
        for (int q = 0; q < 2; ++q)
            for (size_t lo = 0; lo < grp[q].size(); lo += max_batch) {
                const size_t hi = std::min(grp[q].size(), lo + max_batch);
                std::vector<lb2::SdpRequest*> part(grp[q].begin() + (long)lo, grp[q].begin() + (long)hi);
                std::vector<Fiber*> owners(own[q].begin() + (long)lo, own[q].begin() + (long)hi);
                lb2::dropin_submit_sdp(part);
                route_home(owners);
                ++nb; nreq += (int64_t)part.size();
            }
        for (SdpGroup* g : take) delete g;
        trace().ev("chaining", "end", 0, (long)nreq);
        std::lock_guard<std::mutex> lk(d->mu);
        d->sdp_s += secs(t0, Clock::now()); d->sdp_reqs += nreq; d->sdp_batches += nb;
    }
}

// the reference for lb2_aux_run, given once by the host program (lb2_producer_set_reference)
const uint8_t* g_ref_pac = nullptr; int64_t g_ref_l_pac = 0;

void aux_main(Device* d) {
    cudaSetDevice(d->device);
    bool have_ref = false;
    for (;;) {
        std::vector<AuxGroup*> take;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv_aux.wait(lk, [&] { return d->stop || !d->pend_aux.empty(); });
            if (d->pend_aux.empty()) return;
            take.swap(d->pend_aux);
        }
        std::vector<lb2_aux_task> tasks; std::vector<lb2_aux_result> results; std::vector<Fiber*> owners;
        std::vector<lb2_hash_task> hashes; std::vector<lb2_hash_task*> hash_home;
        for (AuxGroup* g : take) for (AuxRequest* q : g->reqs) {
            if (q->hash) { hashes.push_back(*q->hash); hash_home.push_back(q->hash); }
            else tasks.insert(tasks.end(), q->tasks, q->tasks + q->n);
        }
        if (!tasks.empty()) {
            if (!have_ref) {
                if (!g_ref_pac || lb2_ctx_set_reference(d->aux_ctx, g_ref_pac, g_ref_l_pac)) die("record statistics need the reference (lb2_producer_set_reference)");
                have_ref = true;
            }
            results.resize(tasks.size());
            if (lb2_aux_run(d->aux_ctx, (int64_t)tasks.size(), tasks.data(), results.data())) die("record statistics failed");
        }
        if (!hashes.empty()) {      // the seed-and-chain requests of hash_split_map (hash_dropin.cu): one launch
            if (lb2_hash_line_run(d->aux_ctx, (int64_t)hashes.size(), hashes.data())) { fprintf(stderr, "[lamsa_b200] %s\n", lb2_last_error()); die("split-mapping lines failed"); }
            for (size_t k = 0; k < hashes.size(); ++k) { hash_home[k]->m_len = hashes[k].m_len; hash_home[k]->n_hits = hashes[k].n_hits; }
            d->hash_n += (int64_t)hashes.size(); ++d->hash_batches;
        }
        size_t at = 0;
        for (AuxGroup* g : take) {
            for (AuxRequest* q : g->reqs) if (!q->hash) { std::copy(results.begin() + (long)at, results.begin() + (long)(at + (size_t)q->n), q->results); at += (size_t)q->n; }
            owners.insert(owners.end(), g->owners.begin(), g->owners.end());
            delete g;
        }
        route_home(owners);
    }
}

// self test: the "device" hands every request straight back from another thread
void loopback_main(Device* d) {
    for (;;) {
        std::vector<DpGroup*> take;
        {
            std::unique_lock<std::mutex> lk(d->mu);
            d->cv_submit.wait(lk, [&] { return d->stop || !d->pend_fast.empty(); });
            if (d->pend_fast.empty()) return;
            take.swap(d->pend_fast);
        }
        std::vector<Fiber*> owners;
        for (DpGroup* g : take) { owners.insert(owners.end(), g->owners.begin(), g->owners.end()); delete g; }
        route_home(owners);
    }
}

// devices are opened once per process (from lb2_dropin_warmup's helper thread, or by the first join) and
// never torn down: the process is over when the aligner returns
std::mutex g_dev_mu;
std::vector<Device*> g_devices;
Device* g_loopback = nullptr;

// Opens the GPUs on first use.  A process whose workers never park a request (a build that links CPU
// implementations of the entry points, used to test the read pipeline on a box without a GPU) may run without
// one: the failure is kept and raised by the first request that needs a device.
std::string g_dev_error;
std::vector<Device*>& devices() {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_devices.empty() || !g_dev_error.empty()) return g_devices;
    const auto t0 = Clock::now();
    const int ndev = std::max(1, env_i("LB2_DEVICES", 1)), base = env_i("LB2_DEVICE", 0);
    const uint64_t scratch = (uint64_t)env_i("LB2_SLOT_SCRATCH_MB", 8192) << 20;
    // one opener thread per GPU: creating a device's primary context takes of the order of a second
    std::vector<Device*> opened((size_t)ndev, nullptr);
    std::vector<std::string> errs((size_t)ndev);
    {
        std::vector<std::thread> openers;
        for (int k = 0; k < ndev; ++k)
            openers.emplace_back([&, k] {
                Device* d = new Device();
                d->device = base + k;
                bool ok = true;
                // this process runs ten device threads per GPU next to the workers: waits on the GPU should sleep, not spin
                if (!env_i("LB2_SPIN_SYNC", 0)) { cudaSetDevice(d->device); cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync); }
                for (int q = 0; q < kSlots && ok; ++q) {
                    ok = lb2_ctx_create(d->device, &d->slot[q].ctx) == 0;
                    if (ok) lb2_ctx_set_scratch_limit(d->slot[q].ctx, scratch);
                }
                for (int q = 0; q < kSdpThreads && ok; ++q) ok = lb2_ctx_create(d->device, &d->sdp_ctx[q]) == 0;
                ok = ok && lb2_ctx_create(d->device, &d->aux_ctx) == 0;
                if (!ok) errs[(size_t)k] = lb2_last_error();        // lb2_last_error is per thread
                else opened[(size_t)k] = d;
            });
        for (auto& t : openers) t.join();
    }
    for (int k = 0; k < ndev; ++k)
        if (!opened[(size_t)k]) { g_dev_error = errs[(size_t)k]; return g_devices; }      // g_devices stays empty
    const double t_ctx = secs(t0, Clock::now());
    for (Device* d : opened) {
        d->submitter = std::thread(submitter_main, d);
        for (int q = 0; q < kSdpThreads; ++q) d->sdp_thread[q] = std::thread(sdp_main, d, q);
        d->aux_thread = std::thread(aux_main, d);
        for (int q = 0; q < kSlots; ++q) d->slot[q].completer = std::thread(completer_main, d, q);
    }
    g_devices = opened;
    if (verbose()) fprintf(stderr, "[lamsa_b200] %d GPU(s) opened in %.3f s (%d batch slots each; buffers are sized by the slots' own threads)\n", ndev, t_ctx, kSlots);
    return g_devices;
}
[[noreturn]] void no_device() {
    fprintf(stderr, "[lamsa_b200] cannot open the GPU: %s \n", g_dev_error.c_str());
    exit(1);
}

// ---- fibers ------------------------------------------------------------------------------------
size_t stack_bytes() {
    static const size_t v = (size_t)env_i("LB2_FIBER_STACK_KB", 1024) * 1024;
    return v;
}

#if LB2_FAST_SWITCH
inline void switch_to_fiber(Worker* w, Fiber* f) { lb2_fiber_swap(&w->main_sp, f->sp); }
inline void switch_to_sched(Worker* w, Fiber* f) { lb2_fiber_swap(&f->sp, w->main_sp); }
#else
void trampoline(unsigned lo, unsigned hi) {
    Fiber* f = (Fiber*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    f->fn(f->arg);
    f->done = true;
    swapcontext(&f->ctx, &tl_worker->main);       // never resumed
}
inline void switch_to_fiber(Worker* w, Fiber* f) { swapcontext(&w->main, &f->ctx); }
inline void switch_to_sched(Worker* w, Fiber* f) { swapcontext(&f->ctx, &w->main); }
#endif

void yield_to_scheduler() {
    Worker* w = tl_worker;
    ++w->switches;
    switch_to_sched(w, w->cur);
}

[[noreturn]] void no_device();
void bind_device(Worker* w) {
    if (w->dev) return;
    const auto t0 = Clock::now();
    std::vector<Device*>& devs = devices();
    if (devs.empty()) no_device();
    w->dev = devs[(size_t)w->index % devs.size()];
    w->dev_wait_s += secs(t0, Clock::now());
}

void hand_over_dp(Worker* w) {
    bind_device(w);
    Device* d = w->dev;
    const bool f = w->fast && !w->fast->reqs.empty(), s = w->slow && !w->slow->reqs.empty();
    if (!f && !s) return;
    {
        std::lock_guard<std::mutex> lk(d->mu);
        if (f) { if (d->pend_fast.empty()) d->pend_fast_since = Clock::now(); d->pend_fast_tasks += (int64_t)w->fast->reqs.size(); d->pend_fast.push_back(w->fast); }
        if (s) { d->pend_slow_tasks += (int64_t)w->slow->reqs.size(); d->pend_slow.push_back(w->slow); }
    }
    if (f) w->fast = nullptr;
    if (s) w->slow = nullptr;
    ++w->handovers;
    d->cv_submit.notify_one();
}
void hand_over_sdp(Worker* w) {
    if (!w->sdp || w->sdp->reqs.empty()) return;
    bind_device(w);
    Device* d = w->dev;
    {
        std::lock_guard<std::mutex> lk(d->mu);
        d->pend_sdp.push_back(w->sdp);
    }
    w->sdp = nullptr;
    d->cv_sdp.notify_one();
}

void hand_over_aux(Worker* w) {
    if (!w->aux || w->aux->reqs.empty()) return;
    bind_device(w);
    Device* d = w->dev;
    {
        std::lock_guard<std::mutex> lk(d->mu);
        d->pend_aux.push_back(w->aux);
    }
    w->aux = nullptr;
    d->cv_aux.notify_one();
}

void worker_main(Worker* w) {
    tl_worker = w;
    static const size_t flush_dp = (size_t)env_i("LB2_FLUSH_TASKS", 256), flush_sdp = (size_t)env_i("LB2_FLUSH_READS", 32);
    static const auto flush_wait = std::chrono::microseconds(env_i("LB2_FLUSH_US", 250));
    while (w->live > 0) {
        if (w->inbox_n.load(std::memory_order_acquire) > 0) {
            std::lock_guard<std::mutex> lk(w->in_mu);
            w->resumed.insert(w->resumed.end(), w->inbox.begin(), w->inbox.end());
            w->inbox.clear();
            w->inbox_n.store(0, std::memory_order_release);
        }
        Fiber* f = nullptr;
        if (!w->resumed.empty()) { f = w->resumed.front(); w->resumed.pop_front(); }
        else if (!w->fresh.empty()) { f = w->fresh.front(); w->fresh.pop_front(); }
        if (f) {
            w->cur = f;
            const auto tf0 = Clock::now();
            switch_to_fiber(w, f);
            const auto tf1 = Clock::now();
            w->fiber_s += secs(tf0, tf1);
            w->cur = nullptr;
            if (f->done) { --w->live; munmap(f->map, f->map_bytes); f->map = nullptr; }
            // hand over early when enough has gathered or the oldest parked request has waited long enough: the
            // device should not wait for this thread's whole run queue
            if ((w->fast && (w->fast->reqs.size() >= flush_dp || tf1 - w->fast->t_first > flush_wait)) ||
                (w->slow && (w->slow->reqs.size() >= flush_dp / 8 + 1 || tf1 - w->slow->t_first > flush_wait))) hand_over_dp(w);
            if (w->sdp && (w->sdp->reqs.size() >= flush_sdp || tf1 - w->sdp->t_first > flush_wait)) hand_over_sdp(w);
            if (w->aux && (w->aux->reqs.size() >= flush_sdp || tf1 - w->aux->t_first > flush_wait)) hand_over_aux(w);
            continue;
        }
        if (w->live == 0) break;
        hand_over_dp(w);
        hand_over_sdp(w);
        hand_over_aux(w);
        const auto t0 = Clock::now();
        std::unique_lock<std::mutex> lk(w->in_mu);
        w->in_cv.wait(lk, [&] { return !w->inbox.empty(); });
        w->idle_s += secs(t0, Clock::now());
    }
    trace().ev("worker", "exit", w->index, (long)w->switches);
    tl_worker = nullptr;
}

std::mutex g_spawn_mu;
std::vector<Fiber*> g_spawned;       // workers created since the last join
int g_selftest_threads = 0;

int host_threads() {
    if (g_selftest_threads > 0) return g_selftest_threads;
    // the device threads (submitter, completers, chaining) need cores of their own
    int v = env_i("LB2_HOST_THREADS", std::max(1, (int)std::thread::hardware_concurrency() - 3));
    if (v < 1) v = 1;
    return v > 256 ? 256 : v;
}

void run_all(std::vector<Fiber*>& fibers) {
    const auto t0 = Clock::now();
    // The GPUs are opened by lb2_dropin_warmup's helper thread (or by the first request); the workers start at once
    // and only the first hand-over of a thread waits for its device, so reading and parsing the first reads overlap
    // CUDA start-up.
    if (g_selftest_threads > 0 && !g_loopback) {
        g_loopback = new Device(); g_loopback->loopback = true; g_loopback->submitter = std::thread(loopback_main, g_loopback);
    }
    const int K = std::max(1, std::min(host_threads(), (int)fibers.size()));
    std::vector<Worker> workers((size_t)K);
    for (int k = 0; k < K; ++k) { workers[(size_t)k].index = k; workers[(size_t)k].dev = g_selftest_threads > 0 ? g_loopback : nullptr; }
    for (size_t i = 0; i < fibers.size(); ++i) {
        Worker& w = workers[i % (size_t)K];
        fibers[i]->home = &w;
        w.fresh.push_back(fibers[i]);
        ++w.live;
    }
    std::vector<std::thread> th;
    for (int k = 1; k < K; ++k) th.emplace_back(worker_main, &workers[(size_t)k]);
    worker_main(&workers[0]);
    for (auto& t : th) t.join();
    std::vector<Device*> devs;
    if (g_selftest_threads == 0) { std::lock_guard<std::mutex> lk(g_dev_mu); devs = g_devices; }
    if (verbose() && g_selftest_threads == 0) {
        int64_t sw = 0, ho = 0, dn = 0, sn = 0; double dw = 0, idle = 0, fib = 0, pack = 0, dl = 0, sl = 0, dmax = 0;
        for (Worker& w : workers) { dw = std::max(dw, w.dev_wait_s); sw += w.switches; ho += w.handovers; idle += w.idle_s; fib += w.fiber_s; pack += w.pack_s;
                                    dn += w.dp_n; sn += w.sdp_n; dl += w.dp_lat_s; sl += w.sdp_lat_s; dmax = std::max(dmax, w.dp_lat_max); }
        const double wall = secs(t0, Clock::now());
        fprintf(stderr, "[lamsa_b200] %zu workers on %d threads: %.3f s, %lld switches, %lld hand-overs; of the threads' time %.0f %% inside workers "
                        "(%.0f %% of it packing DP requests), %.0f %% idle, up to %.3f s waiting for the GPU to open; a DP request waits %.2f ms on average (max %.1f), a chaining request %.2f ms\n",
                fibers.size(), K, wall, (long long)sw, (long long)ho, 100.0 * fib / (wall * K), fib > 0 ? 100.0 * pack / fib : 0.0, 100.0 * idle / (wall * K), dw,
                dn ? 1e3 * dl / dn : 0.0, 1e3 * dmax, sn ? 1e3 * sl / sn : 0.0);
        for (Device* d : devs) {
            std::lock_guard<std::mutex> lk(d->mu);
            std::vector<int> bt = d->batch_tasks;
            std::sort(bt.begin(), bt.end());
            const int med = bt.empty() ? 0 : bt[bt.size() / 2], mx = bt.empty() ? 0 : bt.back();
            fprintf(stderr, "[lamsa_b200] GPU %d: %lld DP tasks in %zu launches (median %d per launch, max %d), %lld long DP tasks in %lld side batches, "
                            "%lld chaining requests in %lld batches; submitter packing %.3f s, completers %.3f s, kernels %.3f s, chaining thread %.3f s\n",
                    d->device, (long long)d->dp_tasks, bt.size(), med, mx, (long long)d->slow_tasks, (long long)d->slow_batches,
                    (long long)d->sdp_reqs, (long long)d->sdp_batches, d->pack_s, d->deliver_s, d->kernel_ms * 1e-3, d->sdp_s);
            fprintf(stderr, "[lamsa_b200] chaining batches, seconds summed over the threads: flattening the requests %.3f, loading the batch (H2D) %.3f, stage incl. read-back %.3f (kernels alone %.3f), handing back %.3f\n",
                    lb2::g_sdp_ns[0] * 1e-9, lb2::g_sdp_ns[1] * 1e-9, lb2::g_sdp_ns[2] * 1e-9, lb2::g_sdp_ns[3] * 1e-9, lb2::g_sdp_ns[4] * 1e-9);
            if (d->hash_n) fprintf(stderr, "[lamsa_b200] GPU %d: %lld split-mapping lines (k-mer index + chaining of a window) in %lld launches\n", d->device, (long long)d->hash_n, (long long)d->hash_batches);
        }
    }
    for (Fiber* f : fibers) delete f;
    fibers.clear();
}

}  // namespace

namespace lb2 {
bool fiber_active() { return tl_worker && tl_worker->cur; }

// tasks of more than LB2_FAST_ROWS target rows travel in the slot for long tasks: a batch lasts as long as its
// longest task, and the owners of short tasks should resume early
void fiber_wait_dp(DpRequest* r) {
    Worker* w = tl_worker;
    static const int fast_rows = env_i("LB2_FAST_ROWS", 256);
    const bool slow = fast_rows > 0 && r->task.tlen > fast_rows;
    DpGroup*& g = slow ? w->slow : w->fast;
    const auto t0 = Clock::now();
    if (!g) { g = new DpGroup(); g->t_first = t0; }
    if (g_selftest_threads == 0) {
        char msg[200];
        if (g->blob.add(r->task, -1, msg, sizeof msg)) { fprintf(stderr, "[lamsa_b200] DP task rejected: %s\n", msg); exit(1); }
    }
    g->reqs.push_back(r); g->owners.push_back(w->cur);
    const auto t1 = Clock::now();
    w->pack_s += secs(t0, t1);
    yield_to_scheduler();
    w = tl_worker;
    const double lat = secs(t1, Clock::now());
    w->cur->parked_s += lat;
    w->dp_lat_s += lat; ++w->dp_n; if (lat > w->dp_lat_max) w->dp_lat_max = lat;
}
void fiber_wait_sdp(SdpRequest* r) {
    Worker* w = tl_worker;
    if (!w->sdp) { w->sdp = new SdpGroup(); w->sdp->t_first = Clock::now(); }
    w->sdp->reqs.push_back(r); w->sdp->owners.push_back(w->cur);
    const auto t1 = Clock::now();
    yield_to_scheduler();
    w = tl_worker;
    { const double lat = secs(t1, Clock::now()); w->cur->parked_s += lat; w->sdp_lat_s += lat; ++w->sdp_n; }
}
void producer_warmup() { devices(); }
}  // namespace lb2

#if LB2_FAST_SWITCH
extern "C" void lb2_fiber_main(void* p) {                 // entered once per fiber from lb2_fiber_entry_thunk
    Fiber* f = (Fiber*)p;
    f->fn(f->arg);
    f->done = true;
    switch_to_sched(tl_worker, f);                         // never resumed
    abort();
}
#endif

extern "C" void lb2_producer_set_reference(const uint8_t* pac, int64_t l_pac) { g_ref_pac = pac; g_ref_l_pac = l_pac; }

// lb2_aux_run for a worker fiber: parks; all parked requests of all workers are one launch
extern "C" int lb2_worker_aux_counts(int64_t n, const lb2_aux_task* tasks, lb2_aux_result* results) {
    Worker* w = tl_worker;
    if (!w || !w->cur) return lb2::set_error("lb2_worker_aux_counts: not called from a worker of the batch producer");
    if (n <= 0) return 0;
    AuxRequest r{n, tasks, results, nullptr};
    if (!w->aux) { w->aux = new AuxGroup(); w->aux->t_first = Clock::now(); }
    w->aux->reqs.push_back(&r); w->aux->owners.push_back(w->cur);
    const auto t1 = Clock::now();
    yield_to_scheduler();
    tl_worker->cur->parked_s += secs(t1, Clock::now());
    return 0;
}

// lb2_hash_line_run for a worker fiber (hash_dropin.cu): parks; all parked requests of all workers are one launch
int lb2::worker_hash_line(lb2_hash_task* t) {
    Worker* w = tl_worker;
    if (!w || !w->cur) return lb2::set_error("worker_hash_line: not called from a worker of the batch producer");
    AuxRequest r{0, nullptr, nullptr, t};
    if (!w->aux) { w->aux = new AuxGroup(); w->aux->t_first = Clock::now(); }
    w->aux->reqs.push_back(&r); w->aux->owners.push_back(w->cur);
    const auto t1 = Clock::now();
    yield_to_scheduler();
    tl_worker->cur->parked_s += secs(t1, Clock::now());
    return 0;
}

// Seconds the calling worker has spent parked on DP / chaining requests so far (statistics of a host pipeline).
extern "C" double lb2_worker_parked_seconds(void) {
    Worker* w = tl_worker;
    return w && w->cur ? w->cur->parked_s : 0.0;
}

// Let the other workers of this thread run (a worker that has to wait for something a sibling produces).
extern "C" void lb2_worker_yield(void) {
    Worker* w = tl_worker;
    if (!w || !w->cur) { std::this_thread::yield(); return; }
    w->fresh.push_back(w->cur);
    yield_to_scheduler();
}

// ---- self test of the scheduler and the context switch (no GPU needed; tests/test_fibers.py) ----------
namespace {
struct SelfTestArg { int id, yields; double result; long long sum; };
void* selftest_worker(void* p) {
    SelfTestArg* a = (SelfTestArg*)p;
    double acc = (double)a->id;
    long long sum = 0;
    volatile int local[64];
    for (int k = 0; k < 64; ++k) local[k] = a->id * 64 + k;
    for (int y = 0; y < a->yields; ++y) {
        acc = acc * 1.0000001 + (double)y * 0.5;          // floating-point state across switches
        for (int k = 0; k < 64; ++k) sum += local[k] ^ y;  // stack contents across switches
        if (y & 1) lb2_worker_yield();                     // plain yield inside the thread
        else { lb2::DpRequest r{}; lb2::fiber_wait_dp(&r); }   // park, travel through the (loopback) device, come back
    }
    a->result = acc; a->sum = sum;
    return nullptr;
}
}  // namespace
// Runs n workers that yield `yields` times each (alternating between a plain yield and a request that goes
// through another thread and back) on `threads` scheduler threads; returns the number of workers whose results
// differ from the same computation done without any switch (0 = pass).
extern "C" int lb2_fiber_selftest(int n, int yields, int threads) {
    g_selftest_threads = threads > 0 ? threads : 1;
    std::vector<SelfTestArg> args((size_t)n);
    std::vector<pthread_t> ids((size_t)n);
    for (int i = 0; i < n; ++i) { args[(size_t)i] = SelfTestArg{i, yields, 0.0, 0}; lb2_worker_spawn(&ids[(size_t)i], nullptr, selftest_worker, &args[(size_t)i]); }
    for (int i = 0; i < n; ++i) lb2_worker_join(ids[(size_t)i], nullptr);
    g_selftest_threads = 0;
    int bad = 0;
    for (int i = 0; i < n; ++i) {
        double acc = (double)i; long long sum = 0;
        for (int y = 0; y < yields; ++y) { acc = acc * 1.0000001 + (double)y * 0.5; for (int k = 0; k < 64; ++k) sum += (i * 64 + k) ^ y; }
        if (acc != args[(size_t)i].result || sum != args[(size_t)i].sum) ++bad;
    }
    return bad;
}

// pthread_create-shaped: registers a worker; it starts when the first of the workers is joined
extern "C" int lb2_worker_spawn(pthread_t* id, const pthread_attr_t*, void* (*fn)(void*), void* arg) {
    Fiber* f = new Fiber();
    f->fn = fn; f->arg = arg;
    // one inaccessible page below the stack: an overflow faults instead of running into the next mapping
    const size_t page = (size_t)sysconf(_SC_PAGESIZE), sb = stack_bytes();
    f->map_bytes = sb + page;
    f->map = mmap(nullptr, f->map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
    if (f->map == MAP_FAILED) { fprintf(stderr, "[lamsa_b200] cannot map a %zu-byte worker stack\n", f->map_bytes); exit(1); }
    mprotect(f->map, page, PROT_NONE);
    char* stack = (char*)f->map + page;
#if LB2_FAST_SWITCH
    {   // first switch "returns" into lb2_fiber_entry_thunk with the Fiber* in r12 (frame laid out as lb2_fiber_swap pops it)
        uintptr_t top = ((uintptr_t)stack + sb) & ~(uintptr_t)15;
        uint64_t* a = (uint64_t*)(top - 8);                 // return address slot: rsp == top (16-aligned) inside the thunk
        a[0] = (uint64_t)(uintptr_t)&lb2_fiber_entry_thunk;
        a[-1] = 0;                                          // rbp
        a[-2] = 0;                                          // rbx
        a[-3] = (uint64_t)(uintptr_t)f;                     // r12
        a[-4] = 0; a[-5] = 0; a[-6] = 0;                    // r13, r14, r15
        a[-7] = (uint64_t)0x1F80u | ((uint64_t)0x037Fu << 32);   // MXCSR and x87 control word defaults
        f->sp = (void*)(a - 7);
    }
#else
    getcontext(&f->ctx);
    f->ctx.uc_stack.ss_sp = stack; f->ctx.uc_stack.ss_size = sb; f->ctx.uc_link = nullptr;
    const uintptr_t p = (uintptr_t)f;
    makecontext(&f->ctx, (void (*)())trampoline, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
#endif
    std::lock_guard<std::mutex> lk(g_spawn_mu);
    g_spawned.push_back(f);
    if (id) *id = (pthread_t)g_spawned.size();
    return 0;
}

// pthread_join-shaped: the first join after a series of spawns runs ALL spawned workers to completion
extern "C" int lb2_worker_join(pthread_t, void** ret) {
    std::vector<Fiber*> batch;
    {
        std::lock_guard<std::mutex> lk(g_spawn_mu);
        batch.swap(g_spawned);
    }
    if (!batch.empty()) run_all(batch);
    if (ret) *ret = nullptr;
    return 0;
}
