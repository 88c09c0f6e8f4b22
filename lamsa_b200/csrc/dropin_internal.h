// dropin_internal.h -- shared by the drop-in translation units (ksw_dropin.cu, sdp_dropin.cu,
// fiber_sched.cu): the request records a blocked caller parks, and the hooks between the entry
// points and the fiber scheduler.  Not part of the C ABI.
#pragma once
#include <cstdint>
#include <vector>
#include "../../include/lamsa_b200.h"

namespace lb2 {

lb2_ctx* dropin_ctx();                       // the context of the drop-in symbols: the calling thread's own
                                             // (dropin_use_thread_ctx) or else the process-wide one
// make `c` the context of the calling OS thread (the batch producer's chaining thread)
void dropin_bind_thread_ctx(lb2_ctx* c);
bool dropin_has_thread_ctx();                // false: the caller shares the process-wide context with other threads
// open the batch producer's GPUs now (producer.cu)
void producer_warmup();

// one blocked banded-DP call (ksw_dropin.cu)
struct DpRequest {
    lb2_task task;
    lb2_result* res;
    cigar32_t** cig;                         // NULL: caller wants no CIGAR
    bool done;
};
// run requests as ONE GPU batch and hand the results back (ksw_dropin.cu)
void dropin_submit_dp(std::vector<DpRequest*>& batch);

// one blocked chaining call (sdp_dropin.cu).  The read's flattened seed hits and the tracked
// flags stage 1 leaves for stage 2 live in the worker's state.
struct SdpWorkerState {
    lb2_sdp_para para;
    lb2_sdp_read read;                       // offsets are 0: one read
    std::vector<int32_t> seed_id, map_n;
    std::vector<lb2_sdp_hit> hits;
    std::vector<uint8_t> tracked;            // per hit, written by stage 1
};
struct SdpRequest {
    int stage;                               // 1 = frag_line_BCC, 2 = frag_line_remain
    SdpWorkerState* ws;
    std::vector<lb2_sdp_reg> regs;           // stage 2
    std::vector<int32_t> stream;             // result
};
// run requests of one stage as ONE GPU batch (sdp_dropin.cu)
void dropin_submit_sdp(std::vector<SdpRequest*>& batch);

// fiber scheduler (fiber_sched.cu): true when the caller runs inside a worker fiber; the wait
// functions park the request, switch to the scheduler and return once the request was served.
bool fiber_active();
void fiber_wait_dp(DpRequest* r);
void fiber_wait_sdp(SdpRequest* r);
// lb2_hash_line_run for a worker fiber (producer.cu): parks; all parked requests of all workers are one launch
int worker_hash_line(lb2_hash_task* t);

}  // namespace lb2
