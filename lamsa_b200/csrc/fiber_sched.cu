// fiber_sched.cu -- the batch producer: the reference's worker pool (src/lamsa_aln.c:1151-1162,
// `-t N` pthreads, each aligning one read at a time in lamsa_main_aln :825-891) run as N user-level
// fibers on a few OS threads.  A worker that reaches a DP or chaining call parks its request and
// yields; when every worker of an OS thread is parked, the scheduler submits ALL parked requests as
// one GPU batch (thousands of independent DP tasks per launch instead of one), hands the results
// back and resumes the workers.  The data-dependent host control flow of the reference
// (merge_cigar loops, ksw_bi_extend's three stages, the read-level stages) is untouched: dependent
// calls of one read simply take several rounds while the other reads keep the batches full.
//
// Entry points lb2_worker_spawn / lb2_worker_join have the signatures of pthread_create /
// pthread_join, so a build of the reference's lamsa_aln.c with those two names redirected
// (oracle/fiber_wrapper.c, two #defines; no source change) uses fibers instead of threads.
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
#include <mutex>
#include <thread>
#include <vector>

#include "dropin_internal.h"

namespace {

#if defined(__x86_64__) && !defined(LB2_FIBER_UCONTEXT)
#define LB2_FAST_SWITCH 1
extern "C" void lb2_fiber_swap(void** save_sp, void* load_sp);      // fiber_switch.cpp
extern "C" void lb2_fiber_entry_thunk(void);
#else
#define LB2_FAST_SWITCH 0
#endif

struct Fiber {
#if LB2_FAST_SWITCH
    void* sp = nullptr;                  // saved stack pointer while the fiber is not running
#else
    ucontext_t ctx;
#endif
    void* stack = nullptr; size_t stack_bytes = 0;
    void* (*fn)(void*) = nullptr; void* arg = nullptr;
    bool done = false;
};

struct InFlight { lb2::DpAsync* a = nullptr; std::vector<Fiber*> owners; int slot = 0; };

struct Sched {                       // one per OS thread
#if LB2_FAST_SWITCH
    void* main_sp = nullptr;
#else
    ucontext_t main;
#endif
    Fiber* cur = nullptr;
    std::vector<Fiber*> fibers, runnable;
    std::vector<lb2::DpRequest*> dp_wait;   std::vector<Fiber*> dp_owner;
    std::vector<lb2::SdpRequest*> sdp_wait; std::vector<Fiber*> sdp_owner;
    // parked DP requests by class (short / long tasks), batches in flight, workers not yet started
    std::vector<lb2::DpRequest*> fast_wait, slow_wait; std::vector<Fiber*> fast_owner, slow_owner;
    std::vector<InFlight> inflight;
    bool slot_busy[lb2::kAsyncSlots] = {false};
    std::vector<Fiber*> staged;
    int index = 0;
    int64_t switches = 0, flushes = 0, dp_tasks = 0, sdp_reqs = 0, max_dp = 0, slow_batches = 0, slow_tasks = 0, sdp_batches = 0;
    double gpu_s = 0, ctx_s = 0, sdp_s = 0, submit_s = 0, wait_s = 0;
};
thread_local Sched* tl_sched = nullptr;

std::mutex g_spawn_mu;
std::vector<Fiber*> g_spawned;       // workers created since the last join
bool g_verbose() { static const bool v = getenv("LB2_FIBER_STATS") != nullptr; return v; }

size_t stack_bytes() {
    static const size_t v = [] { const char* e = getenv("LB2_FIBER_STACK_KB"); return (size_t)(e && *e ? atol(e) : 1024) * 1024; }();
    return v;
}
// scheduler threads per GPU unless LB2_HOST_THREADS says otherwise: two with CUDA's default 8 hardware work
// queues, four with 16+, eight with 32 (CUDA_DEVICE_MAX_CONNECTIONS; see lb2_dropin_warmup in ksw_dropin.cu).  Measured on a
// 16-thread B200 host, steady-state chunk of 4 096 reads x 5 kbp:
//    8 queues: 1 thread 0.58 s, 2 threads 0.38 s, 3 threads 0.44 s, 4 threads 0.60 s
//   16 queues: 4 threads 0.27 s;   32 queues: 4 threads 0.25 s, 8 threads 0.22 s, 16 threads 0.45 s
// (with few queues the small launches of many contexts serialise; with few threads the host control flow of
// 1 024 workers per thread is the bound).
// wall-clock marks for LB2_FIBER_STATS: process start (library load), first spawn, end of the last join, exit
const std::chrono::steady_clock::time_point g_t_load = std::chrono::steady_clock::now();
double g_first_spawn_s = -1, g_last_join_s = -1, g_in_chunks_s = 0;
int g_real_chunks = 0;                 // chunks run with GPU contexts (not the self test)
double since_load() { return std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t_load).count(); }
// Runs when the program exits (registered at the first chunk, i.e. after CUDA registered its own handlers, so it
// runs before them).  A `lamsa aln` process is done at this point: tearing down tens of CUDA contexts and their
// pinned buffers costs about 0.7 s that the operating system does for free, so after flushing every stdio stream
// the process leaves with _exit and the status it was given (LB2_FAST_EXIT=0 keeps the ordinary teardown).
void report_at_exit(int status, void*) {
    if (g_verbose() && g_first_spawn_s >= 0) {
        const double t = since_load();
        fprintf(stderr, "[lamsa_b200] wall clock since library load: first worker spawned at %.3f s, %.3f s inside worker chunks, "
                        "%.3f s between/after chunks until the last join (%.3f s), exit at %.3f s\n",
                g_first_spawn_s, g_in_chunks_s, g_last_join_s - g_first_spawn_s - g_in_chunks_s, g_last_join_s, t);
    }
    const char* e = getenv("LB2_FAST_EXIT");
    if (g_real_chunks > 0 && !(e && *e == '0')) { fflush(nullptr); _exit(status); }
}
bool g_selftest = false;              // lb2_fiber_selftest: no GPU contexts, workers yield through selftest_yield
int g_selftest_threads = 0;
int host_threads() {
    if (g_selftest_threads > 0) return g_selftest_threads;
    const char* e = getenv("LB2_HOST_THREADS");
    const char* d = getenv("LB2_DEVICES");
    const char* q = getenv("CUDA_DEVICE_MAX_CONNECTIONS");
    const int ndev = d && *d && atoi(d) > 0 ? atoi(d) : 1;
    const int per_gpu = q && atoi(q) >= 32 ? 8 : q && atoi(q) >= 16 ? 4 : 2;
    int v = e && *e ? atoi(e) : std::min(per_gpu * ndev, (int)std::thread::hardware_concurrency());
    if (v < 1) v = 1;
    return v > 64 ? 64 : v;
}

#if LB2_FAST_SWITCH
inline void switch_to_fiber(Sched* s, Fiber* f) { lb2_fiber_swap(&s->main_sp, f->sp); }
inline void switch_to_sched(Sched* s, Fiber* f) { lb2_fiber_swap(&f->sp, s->main_sp); }
#else
void trampoline(unsigned lo, unsigned hi) {
    Fiber* f = (Fiber*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    f->fn(f->arg);
    f->done = true;
    swapcontext(&f->ctx, &tl_sched->main);       // never resumed
}
inline void switch_to_fiber(Sched* s, Fiber* f) { swapcontext(&s->main, &f->ctx); }
inline void switch_to_sched(Sched* s, Fiber* f) { swapcontext(&f->ctx, &s->main); }
#endif

void yield_to_scheduler() {
    Sched* s = tl_sched;
    Fiber* f = s->cur;
    ++s->switches;
    switch_to_sched(s, f);
}

// ---- rounds ------------------------------------------------------------------------------
// Each OS thread owns its contexts (own streams and scratch; thread k of a process that sees D GPUs
// uses GPU k mod D) and is never synchronised with the other threads.  Whenever none of its workers
// can run, it submits what they have parked as ASYNCHRONOUS batches (up to kAsyncSlots in flight)
// and, if still nothing can run, waits for the oldest batch and resumes its owners.
//  * The workers are started in two halves, so that while one half's batch is on the GPU the other
//    half does its host work: submission, kernels and host control flow overlap.
//  * A batch lasts as long as its longest DP task (one warp walks the rows of a task: about 1.5 us
//    per row with the traceback), while the median task has 50-150 rows and the longest thousands
//    (SURVEY.md appendix C).  Parked tasks are therefore split: tasks of more than LB2_FAST_ROWS
//    target rows travel in a batch of their own, so the owners of short tasks resume early.
int fast_rows() {
    static const int v = [] { const char* e = getenv("LB2_FAST_ROWS"); return e && *e ? atoi(e) : 256; }();
    return v;
}

bool submit_group(Sched* s, std::vector<lb2::DpRequest*>& reqs, std::vector<Fiber*>& owners, bool slow) {
    if (reqs.empty()) return true;
    int slot = -1;
    for (int k = 0; k < lb2::kAsyncSlots; ++k) if (!s->slot_busy[k]) { slot = k; break; }
    if (slot < 0) return false;
    InFlight f;
    f.a = lb2::dropin_dp_async_submit(reqs, slot);
    f.owners.swap(owners);
    f.slot = slot;
    s->slot_busy[slot] = true;
    s->inflight.push_back(std::move(f));
    if (slow) { ++s->slow_batches; s->slow_tasks += (int64_t)reqs.size(); }
    else { ++s->flushes; s->dp_tasks += (int64_t)reqs.size(); s->max_dp = std::max<int64_t>(s->max_dp, (int64_t)reqs.size()); }
    reqs.clear();
    return true;
}
void finish_at(Sched* s, size_t i) {                 // waits if the batch is still running
    InFlight f = std::move(s->inflight[i]);
    s->inflight.erase(s->inflight.begin() + (long)i);
    lb2::dropin_dp_async_finish(f.a);
    s->slot_busy[f.slot] = false;
    for (Fiber* w : f.owners) s->runnable.push_back(w);
}

// nothing can run: submit, reap, wait
void flush(Sched* s) {
    const auto t0 = std::chrono::steady_clock::now();
    // chaining requests (two per read) are synchronous batches: worth a launch when enough of them have
    // gathered, or when there is no DP work to overlap with; until then their owners stay parked
    static const size_t sdp_min = [] { const char* e = getenv("LB2_SDP_MIN_BATCH"); return (size_t)(e && *e ? atoi(e) : 64); }();
    const bool other_work = !s->dp_wait.empty() || !s->fast_wait.empty() || !s->slow_wait.empty() || !s->inflight.empty() || !s->staged.empty();
    if (!s->sdp_wait.empty() && (s->sdp_wait.size() >= sdp_min || !other_work)) {
        const auto ts0 = std::chrono::steady_clock::now();
        for (int stage = 1; stage <= 2; ++stage) {
            std::vector<lb2::SdpRequest*> grp;
            for (lb2::SdpRequest* q : s->sdp_wait) if (q->stage == stage) grp.push_back(q);
            if (!grp.empty()) { lb2::dropin_submit_sdp(grp); ++s->sdp_batches; }
        }
        for (Fiber* f : s->sdp_owner) s->runnable.push_back(f);
        s->sdp_reqs += (int64_t)s->sdp_wait.size();
        s->sdp_wait.clear(); s->sdp_owner.clear();
        s->sdp_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - ts0).count();
    }
    const int lim = fast_rows();
    for (size_t i = 0; i < s->dp_wait.size(); ++i) {
        const bool slow = lim > 0 && s->dp_wait[i]->task.tlen > lim;
        (slow ? s->slow_wait : s->fast_wait).push_back(s->dp_wait[i]);
        (slow ? s->slow_owner : s->fast_owner).push_back(s->dp_owner[i]);
    }
    s->dp_wait.clear(); s->dp_owner.clear();
    const auto tb0 = std::chrono::steady_clock::now();
    submit_group(s, s->fast_wait, s->fast_owner, false);
    submit_group(s, s->slow_wait, s->slow_owner, true);
    const auto tb1 = std::chrono::steady_clock::now();
    s->submit_s += std::chrono::duration<double>(tb1 - tb0).count();
    for (size_t i = 0; i < s->inflight.size();)      // reap what is ready
        if (lb2::dropin_dp_async_done(s->inflight[i].a)) finish_at(s, i); else ++i;
    if (s->runnable.empty() && !s->staged.empty()) { s->runnable.swap(s->staged); }   // second half of the workers starts now
    if (s->runnable.empty() && !s->inflight.empty()) finish_at(s, 0);                  // wait for the oldest batch
    s->wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - tb1).count();
    s->gpu_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

void run_scheduler(Sched* s) {
    tl_sched = s;
    const auto tc0 = std::chrono::steady_clock::now();
    if (!g_selftest) lb2::dropin_use_thread_ctx(s->index);
    s->ctx_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - tc0).count();
    size_t live = s->fibers.size();
    // s->fibers is in reverse worker order (the run queue pops from the back)
    const size_t half = s->fibers.size() / 2;
    for (size_t i = 0; i < s->fibers.size(); ++i) (i < half ? s->staged : s->runnable).push_back(s->fibers[i]);
    while (live > 0) {
        while (!s->runnable.empty()) {
            Fiber* f = s->runnable.back(); s->runnable.pop_back();
            s->cur = f;
            switch_to_fiber(s, f);
            s->cur = nullptr;
            if (f->done) { --live; munmap(f->stack, f->stack_bytes); f->stack = nullptr; }
        }
        if (live == 0) break;
        if (s->dp_wait.empty() && s->sdp_wait.empty() && s->slow_wait.empty() && s->fast_wait.empty() && s->inflight.empty() && s->staged.empty()) {
            fprintf(stderr, "[lamsa_b200] fiber scheduler: %zu workers alive but nothing parked\n", live); exit(1);
        }
        flush(s);
    }
    tl_sched = nullptr;
}

void run_all(std::vector<Fiber*>& fibers) {
    const auto t0 = std::chrono::steady_clock::now();
    static const bool hooked = [] { on_exit(report_at_exit, nullptr); return true; }();
    (void)hooked;
    const int K = std::max(1, std::min(host_threads(), (int)fibers.size()));
    std::vector<Sched> scheds((size_t)K);
    for (size_t i = 0; i < fibers.size(); ++i) scheds[i % K].fibers.push_back(fibers[i]);
    // spawn order = worker order; the scheduler pops from the back, so reverse to start worker 0 first
    for (Sched& s : scheds) std::reverse(s.fibers.begin(), s.fibers.end());
    for (int k = 0; k < K; ++k) scheds[(size_t)k].index = k;
    std::vector<std::thread> th;
    for (int k = 1; k < K; ++k) th.emplace_back(run_scheduler, &scheds[(size_t)k]);
    run_scheduler(&scheds[0]);
    for (auto& t : th) t.join();
    if (g_verbose()) {
        int64_t sw = 0, fl = 0, dp = 0, sd = 0, mx = 0, sb = 0, stt = 0, sdb = 0; double g = 0, cx = 0, ss = 0, su = 0, wa = 0;
        for (Sched& s : scheds) { cx = std::max(cx, s.ctx_s); ss = std::max(ss, s.sdp_s); su = std::max(su, s.submit_s); wa = std::max(wa, s.wait_s);
                                  sw += s.switches; fl += s.flushes; dp += s.dp_tasks; sd += s.sdp_reqs; mx = std::max(mx, s.max_dp); g = std::max(g, s.gpu_s);
                                  sb += s.slow_batches; stt += s.slow_tasks; sdb += s.sdp_batches; }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[lamsa_b200] %zu workers on %d threads: %.3f s, %lld rounds (%.0f per thread), %lld DP tasks (%.0f per launch, max %lld), "
                        "%lld long DP tasks in %lld side batches, %lld chaining requests in %lld batches, %lld switches, %.3f s inside GPU submissions (slowest thread; max per thread: contexts %.3f, chaining %.3f, DP submit %.3f, DP wait %.3f)\n",
                fibers.size(), K, wall, (long long)fl, (double)fl / K, (long long)dp, fl ? (double)dp / fl : 0.0, (long long)mx,
                (long long)stt, (long long)sb, (long long)sd, (long long)sdb, (long long)sw, g, cx, ss, su, wa);
    }
    for (Fiber* f : fibers) delete f;
    fibers.clear();
    g_in_chunks_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    g_last_join_s = since_load();
    if (!g_selftest) ++g_real_chunks;
}

}  // namespace

namespace lb2 {
bool fiber_active() { return tl_sched && tl_sched->cur; }
void fiber_wait_dp(DpRequest* r) {
    Sched* s = tl_sched;
    s->dp_wait.push_back(r); s->dp_owner.push_back(s->cur);
    yield_to_scheduler();
}
void fiber_wait_sdp(SdpRequest* r) {
    Sched* s = tl_sched;
    s->sdp_wait.push_back(r); s->sdp_owner.push_back(s->cur);
    yield_to_scheduler();
}
}  // namespace lb2

#if LB2_FAST_SWITCH
extern "C" void lb2_fiber_main(void* p) {                 // entered once per fiber from lb2_fiber_entry_thunk
    Fiber* f = (Fiber*)p;
    f->fn(f->arg);
    f->done = true;
    switch_to_sched(tl_sched, f);                          // never resumed
    abort();
}
#endif

// ---- self test of the scheduler and the context switch (no GPU needed; tests/test_fibers.py) ----------
namespace {
struct SelfTestArg { int id, yields; double result; long long sum; };
void selftest_yield() {               // give the other workers of this thread a turn
    Sched* s = tl_sched;
    s->runnable.insert(s->runnable.begin(), s->cur);
    yield_to_scheduler();
}
void* selftest_worker(void* p) {
    SelfTestArg* a = (SelfTestArg*)p;
    double acc = (double)a->id;
    long long sum = 0;
    volatile int local[64];
    for (int k = 0; k < 64; ++k) local[k] = a->id * 64 + k;
    for (int y = 0; y < a->yields; ++y) {
        acc = acc * 1.0000001 + (double)y * 0.5;          // floating-point state across switches
        for (int k = 0; k < 64; ++k) sum += local[k] ^ y;  // stack contents across switches
        selftest_yield();
    }
    a->result = acc; a->sum = sum;
    return nullptr;
}
}  // namespace
// Runs n workers that yield `yields` times each on `threads` scheduler threads; returns the number of workers
// whose results differ from the same computation done without any switch (0 = pass).
extern "C" int lb2_fiber_selftest(int n, int yields, int threads) {
    g_selftest = true; g_selftest_threads = threads > 0 ? threads : 1;
    std::vector<SelfTestArg> args((size_t)n);
    std::vector<pthread_t> ids((size_t)n);
    for (int i = 0; i < n; ++i) { args[(size_t)i] = SelfTestArg{i, yields, 0.0, 0}; lb2_worker_spawn(&ids[(size_t)i], nullptr, selftest_worker, &args[(size_t)i]); }
    for (int i = 0; i < n; ++i) lb2_worker_join(ids[(size_t)i], nullptr);
    g_selftest = false; g_selftest_threads = 0;
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
    if (g_first_spawn_s < 0) g_first_spawn_s = since_load();
    Fiber* f = new Fiber();
    f->fn = fn; f->arg = arg;
    f->stack_bytes = stack_bytes();
    f->stack = mmap(nullptr, f->stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
    if (f->stack == MAP_FAILED) { fprintf(stderr, "[lamsa_b200] cannot map a %zu-byte worker stack\n", f->stack_bytes); exit(1); }
#if LB2_FAST_SWITCH
    {   // first switch "returns" into lb2_fiber_entry_thunk with the Fiber* in r12 (frame laid out as lb2_fiber_swap pops it)
        uintptr_t top = ((uintptr_t)f->stack + f->stack_bytes) & ~(uintptr_t)15;
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
    f->ctx.uc_stack.ss_sp = f->stack; f->ctx.uc_stack.ss_size = f->stack_bytes; f->ctx.uc_link = nullptr;
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
