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

struct Fiber {
    ucontext_t ctx;
    void* stack = nullptr; size_t stack_bytes = 0;
    void* (*fn)(void*) = nullptr; void* arg = nullptr;
    bool done = false;
};

struct Sched {                       // one per OS thread
    ucontext_t main;
    Fiber* cur = nullptr;
    std::vector<Fiber*> fibers, runnable;
    std::vector<lb2::DpRequest*> dp_wait;   std::vector<Fiber*> dp_owner;
    std::vector<lb2::SdpRequest*> sdp_wait; std::vector<Fiber*> sdp_owner;
    int index = 0;
    int64_t switches = 0, flushes = 0, dp_tasks = 0, sdp_reqs = 0, max_dp = 0;
    double gpu_s = 0;
};
thread_local Sched* tl_sched = nullptr;

std::mutex g_spawn_mu;
std::vector<Fiber*> g_spawned;       // workers created since the last join
bool g_verbose() { static const bool v = getenv("LB2_FIBER_STATS") != nullptr; return v; }

size_t stack_bytes() {
    static const size_t v = [] { const char* e = getenv("LB2_FIBER_STACK_KB"); return (size_t)(e && *e ? atol(e) : 1024) * 1024; }();
    return v;
}
int host_threads() {
    const char* e = getenv("LB2_HOST_THREADS");
    int v = e && *e ? atoi(e) : (int)std::thread::hardware_concurrency();
    if (v < 1) v = 1;
    return v > 64 ? 64 : v;
}

void trampoline(unsigned lo, unsigned hi) {
    Fiber* f = (Fiber*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    f->fn(f->arg);
    f->done = true;
    swapcontext(&f->ctx, &tl_sched->main);       // never resumed
}

void yield_to_scheduler() {
    Sched* s = tl_sched;
    Fiber* f = s->cur;
    ++s->switches;
    swapcontext(&f->ctx, &s->main);
}

// ---- rounds ------------------------------------------------------------------------------
// Each OS thread owns a context (its own streams and scratch; thread k of a process that sees D
// GPUs uses GPU k mod D) and advances in rounds of its own: it runs its workers until all of them
// are parked, submits everything parked as one chaining batch per stage and one DP batch, and
// resumes them.  Threads do not wait for each other, so a long DP task only delays the workers
// that share its batch, and the batches of different threads overlap on the GPU(s).
void flush(Sched* s) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int stage = 1; stage <= 2; ++stage) {
        std::vector<lb2::SdpRequest*> grp;
        for (lb2::SdpRequest* q : s->sdp_wait) if (q->stage == stage) grp.push_back(q);
        if (!grp.empty()) lb2::dropin_submit_sdp(grp);
    }
    if (!s->dp_wait.empty()) lb2::dropin_submit_dp(s->dp_wait);
    s->gpu_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    ++s->flushes; s->dp_tasks += (int64_t)s->dp_wait.size(); s->sdp_reqs += (int64_t)s->sdp_wait.size();
    s->max_dp = std::max<int64_t>(s->max_dp, (int64_t)s->dp_wait.size());
    for (Fiber* f : s->sdp_owner) s->runnable.push_back(f);
    for (Fiber* f : s->dp_owner) s->runnable.push_back(f);
    s->dp_wait.clear(); s->dp_owner.clear(); s->sdp_wait.clear(); s->sdp_owner.clear();
}

void run_scheduler(Sched* s) {
    tl_sched = s;
    lb2::dropin_use_thread_ctx(s->index);
    size_t live = s->fibers.size();
    for (Fiber* f : s->fibers) s->runnable.push_back(f);
    while (live > 0) {
        while (!s->runnable.empty()) {
            Fiber* f = s->runnable.back(); s->runnable.pop_back();
            s->cur = f;
            swapcontext(&s->main, &f->ctx);
            s->cur = nullptr;
            if (f->done) { --live; munmap(f->stack, f->stack_bytes); f->stack = nullptr; }
        }
        if (live == 0) break;
        if (s->dp_wait.empty() && s->sdp_wait.empty()) {
            fprintf(stderr, "[lamsa_b200] fiber scheduler: %zu workers alive but nothing parked\n", live); exit(1);
        }
        flush(s);
    }
    tl_sched = nullptr;
}

void run_all(std::vector<Fiber*>& fibers) {
    const auto t0 = std::chrono::steady_clock::now();
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
        int64_t sw = 0, fl = 0, dp = 0, sd = 0, mx = 0; double g = 0;
        for (Sched& s : scheds) { sw += s.switches; fl += s.flushes; dp += s.dp_tasks; sd += s.sdp_reqs; mx = std::max(mx, s.max_dp); g = std::max(g, s.gpu_s); }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[lamsa_b200] %zu workers on %d threads: %.3f s, %lld rounds (%.0f per thread), %lld DP tasks (%.0f per launch, max %lld), "
                        "%lld chaining requests, %lld switches, %.3f s inside GPU submissions (slowest thread)\n",
                fibers.size(), K, wall, (long long)fl, (double)fl / K, (long long)dp, fl ? (double)dp / fl : 0.0, (long long)mx,
                (long long)sd, (long long)sw, g);
    }
    for (Fiber* f : fibers) delete f;
    fibers.clear();
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

// pthread_create-shaped: registers a worker; it starts when the first of the workers is joined
extern "C" int lb2_worker_spawn(pthread_t* id, const pthread_attr_t*, void* (*fn)(void*), void* arg) {
    Fiber* f = new Fiber();
    f->fn = fn; f->arg = arg;
    f->stack_bytes = stack_bytes();
    f->stack = mmap(nullptr, f->stack_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE | MAP_STACK, -1, 0);
    if (f->stack == MAP_FAILED) { fprintf(stderr, "[lamsa_b200] cannot map a %zu-byte worker stack\n", f->stack_bytes); exit(1); }
    getcontext(&f->ctx);
    f->ctx.uc_stack.ss_sp = f->stack; f->ctx.uc_stack.ss_size = f->stack_bytes; f->ctx.uc_link = nullptr;
    const uintptr_t p = (uintptr_t)f;
    makecontext(&f->ctx, (void (*)())trampoline, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
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
