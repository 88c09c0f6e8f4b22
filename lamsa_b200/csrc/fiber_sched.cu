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
    // DP tasks too long for a round: parked until the next asynchronous batch, and the batch in flight
    std::vector<lb2::DpRequest*> slow_wait; std::vector<Fiber*> slow_owner;
    lb2::DpAsync* slow_inflight = nullptr;  std::vector<Fiber*> slow_inflight_owner;
    int index = 0;
    int64_t switches = 0, flushes = 0, dp_tasks = 0, sdp_reqs = 0, max_dp = 0, slow_batches = 0, slow_tasks = 0;
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
// Each OS thread owns two contexts (own streams and scratch; thread k of a process that sees D
// GPUs uses GPU k mod D) and advances in rounds of its own: it runs its workers until all of them
// are parked, submits what is parked and resumes the owners.  Threads do not wait for each other.
//
// A round lasts as long as its longest DP task (one warp walks the rows of a task: about 1.5 us
// per row with the traceback), while the median task has 50-150 rows and the longest thousands
// (SURVEY.md appendix C).  So the parked DP tasks are split: tasks of at most LB2_FAST_ROWS target
// rows go out at once as the round's batch; longer ones go to the thread's second context as an
// asynchronous batch that runs beside the following rounds, and their owners resume when it is done.
int fast_rows() {
    static const int v = [] { const char* e = getenv("LB2_FAST_ROWS"); return e && *e ? atoi(e) : 512; }();
    return v;
}

void start_slow_batch(Sched* s) {
    if (s->slow_inflight || s->slow_wait.empty()) return;
    s->slow_inflight = lb2::dropin_dp_async_submit(s->slow_wait);
    s->slow_inflight_owner.swap(s->slow_owner);
    ++s->slow_batches; s->slow_tasks += (int64_t)s->slow_wait.size();
    s->slow_wait.clear();
}
void finish_slow_batch(Sched* s) {          // waits if the batch is still running
    lb2::dropin_dp_async_finish(s->slow_inflight);
    s->slow_inflight = nullptr;
    for (Fiber* f : s->slow_inflight_owner) s->runnable.push_back(f);
    s->slow_inflight_owner.clear();
}

void flush(Sched* s) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int stage = 1; stage <= 2; ++stage) {
        std::vector<lb2::SdpRequest*> grp;
        for (lb2::SdpRequest* q : s->sdp_wait) if (q->stage == stage) grp.push_back(q);
        if (!grp.empty()) lb2::dropin_submit_sdp(grp);
    }
    for (Fiber* f : s->sdp_owner) s->runnable.push_back(f);
    s->sdp_reqs += (int64_t)s->sdp_wait.size();
    s->sdp_wait.clear(); s->sdp_owner.clear();

    std::vector<lb2::DpRequest*> fast; std::vector<Fiber*> fast_owner;
    const int lim = fast_rows();
    for (size_t i = 0; i < s->dp_wait.size(); ++i) {
        if (lim > 0 && s->dp_wait[i]->task.tlen > lim) { s->slow_wait.push_back(s->dp_wait[i]); s->slow_owner.push_back(s->dp_owner[i]); }
        else { fast.push_back(s->dp_wait[i]); fast_owner.push_back(s->dp_owner[i]); }
    }
    s->dp_wait.clear(); s->dp_owner.clear();
    start_slow_batch(s);                       // runs beside the fast batch below and the next rounds
    if (!fast.empty()) {
        lb2::dropin_submit_dp(fast);
        ++s->flushes; s->dp_tasks += (int64_t)fast.size();
        s->max_dp = std::max<int64_t>(s->max_dp, (int64_t)fast.size());
        for (Fiber* f : fast_owner) s->runnable.push_back(f);
    }
    if (s->slow_inflight && (s->runnable.empty() || lb2::dropin_dp_async_done(s->slow_inflight))) {
        finish_slow_batch(s);                  // nothing else to do, or it is ready anyway
        start_slow_batch(s);
    }
    s->gpu_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
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
        if (s->dp_wait.empty() && s->sdp_wait.empty() && s->slow_wait.empty() && !s->slow_inflight) {
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
        int64_t sw = 0, fl = 0, dp = 0, sd = 0, mx = 0, sb = 0, stt = 0; double g = 0;
        for (Sched& s : scheds) { sw += s.switches; fl += s.flushes; dp += s.dp_tasks; sd += s.sdp_reqs; mx = std::max(mx, s.max_dp); g = std::max(g, s.gpu_s);
                                  sb += s.slow_batches; stt += s.slow_tasks; }
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[lamsa_b200] %zu workers on %d threads: %.3f s, %lld rounds (%.0f per thread), %lld DP tasks (%.0f per launch, max %lld), "
                        "%lld long DP tasks in %lld side batches, %lld chaining requests, %lld switches, %.3f s inside GPU submissions (slowest thread)\n",
                fibers.size(), K, wall, (long long)fl, (double)fl / K, (long long)dp, fl ? (double)dp / fl : 0.0, (long long)mx,
                (long long)stt, (long long)sb, (long long)sd, (long long)sw, g);
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
