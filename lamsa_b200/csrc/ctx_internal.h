// ctx_internal.h -- what the other translation units of liblamsa_b200.so may see of a context
// (the struct itself is private to dp_batch.cu).  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
struct lb2_ctx;
namespace lb2 {
cudaStream_t ctx_stream(lb2_ctx* c);
int ctx_device(lb2_ctx* c);
int ctx_sm_count(lb2_ctx* c);
int set_error(const char* fmt, ...);     // stores the message for lb2_last_error(), returns 1
}
