// ctx_internal.h -- what the other translation units of liblamsa_b200.so may see of a context
// (the struct itself is private to dp_batch.cu).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include "../../include/lamsa_b200.h"
struct lb2_ctx;
namespace lb2 {
struct TaskBlob;                          // dp_pack.h
cudaStream_t ctx_stream(lb2_ctx* c);
int ctx_device(lb2_ctx* c);
int ctx_sm_count(lb2_ctx* c);
int set_error(const char* fmt, ...);     // stores the message for lb2_last_error(), returns 1
// lb2_batch_create for tasks that were classified and copied by their producers (dp_batch.cu)
int batch_create_staged(lb2_ctx* ctx, const TaskBlob* const* blobs, int nblobs, lb2_batch** out);
// pre-size one set of a context's grow-only batch buffers (pinned host + device) and its scratch
int ctx_reserve(lb2_ctx* ctx, int64_t n_tasks, size_t pool_bytes, size_t z_bytes, size_t cigar_words);
// block (no spinning) until the kernels of an enqueued batch have finished
int batch_wait_blocking(lb2_batch* b);
}
