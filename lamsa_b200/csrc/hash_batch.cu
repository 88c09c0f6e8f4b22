// hash_batch.cu -- host side of lb2_hash_line_run (include/lamsa_b200.h section 6): packs the (reference window,
// read) pairs of a batch, sizes the per-request scratch, launches hash_line_kernel (hash_line.cuh) and hands the
// lines back.  The node arrays of all requests come out of one pool (a request needs 10 ints per hit and the hits
// are only known on the device): when the pool runs out the batch is run again with the size the kernel asked for.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/lamsa_b200.h"
#include "ctx_internal.h"
#include "hash_line.cuh"

using namespace lb2;

namespace {
struct HashBuffers {       // grow-only, one set per context
    uint8_t* h = nullptr; uint8_t* d = nullptr; size_t cap = 0;          // staging: requests, sequences | results, lines
    int32_t* d_fixed = nullptr; size_t fixed_cap = 0;
    int32_t* d_nodes = nullptr; size_t nodes_cap = 0;
    unsigned long long* d_cursor = nullptr;
    std::mutex mu;
};
std::mutex g_mu;
std::map<lb2_ctx*, HashBuffers*> g_buffers;
HashBuffers* buffers_of(lb2_ctx* c) {
    std::lock_guard<std::mutex> lk(g_mu);
    HashBuffers*& b = g_buffers[c];
    if (!b) b = new HashBuffers();
    return b;
}
size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }
}  // namespace

#define CUH(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) \
    return lb2::set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); } while (0)

extern "C" int lb2_hash_line_run(lb2_ctx* ctx, int64_t n, lb2_hash_task* tasks) {
    if (!ctx || n < 0 || (n > 0 && !tasks)) return lb2::set_error("lb2_hash_line_run: bad argument");
    if (n == 0) return 0;
    HashBuffers* B = buffers_of(ctx);
    std::lock_guard<std::mutex> lk(B->mu);
    CUH(cudaSetDevice(lb2::ctx_device(ctx)));
    cudaStream_t s = lb2::ctx_stream(ctx);
    // ---- layout
    std::vector<HashReq> reqs((size_t)n);
    size_t seq_bytes = 0, fixed_ints = 0, out_ints = 0, node_guess = 0;
    for (int64_t i = 0; i < n; ++i) {
        lb2_hash_task& t = tasks[i];
        if (t.ref_len < 0 || t.read_len < 0 || (t.ref_len && !t.ref) || (t.read_len && !t.read))
            return lb2::set_error("lb2_hash_line_run: task %lld is malformed", (long long)i);
        if (t.hash_len < 1 || t.hash_len > 15 || t.hash_step < 1)
            return lb2::set_error("lb2_hash_line_run: task %lld: hash_len %d / hash_step %d unsupported (1..15, >= 1)", (long long)i, t.hash_len, t.hash_step);
        HashReq& q = reqs[(size_t)i];
        q.ref_len = t.ref_len; q.read_len = t.read_len; q.ref_offset = t.ref_offset;
        q.hash_len = t.hash_len; q.hash_step = t.hash_step; q.split_len = t.split_len; q.head_on = t.head ? 1 : 0; q.tail_on = t.tail ? 1 : 0;
        q.S = t.read_len >= t.hash_len ? (t.read_len - t.hash_len) / t.hash_step + 1 : 0;
        q.R = t.ref_len >= t.hash_len ? t.ref_len - t.hash_len + 1 : 0;
        int lt = 6; while ((1 << lt) < 2 * q.S + 2) ++lt;
        q.logT = lt;
        q.ref_off = (uint32_t)seq_bytes; seq_bytes += up16((size_t)t.ref_len + 16);
        q.read_off = (uint32_t)seq_bytes; seq_bytes += up16((size_t)t.read_len + 16);
        if (seq_bytes >> 31) return lb2::set_error("lb2_hash_line_run: more than 2 GB of sequences in one batch");
        q.fixed_off = (uint32_t)fixed_ints;
        fixed_ints += (size_t)4 * ((size_t)1 << lt) + 3 * ((size_t)q.S + 2) + (size_t)q.R + 1 + ((size_t)(t.ref_len + t.read_len + 2 + 31) / 32 + 1) + 8;
        q.out_off = (uint32_t)out_ints; out_ints += 3 * (size_t)std::max(q.S, 1);
        node_guess += 10 * ((size_t)2 * q.S + 64);
        if (t.line_cap < q.S) return lb2::set_error("lb2_hash_line_run: task %lld: line_cap %d, %d nodes possible", (long long)i, t.line_cap, q.S);
        if ((fixed_ints | out_ints) >> 31) return lb2::set_error("lb2_hash_line_run: batch too large");
    }
    const size_t o_req = 0, o_seq = up16(sizeof(HashReq) * (size_t)n), o_res = o_seq + up16(seq_bytes),
                 o_out = o_res + up16(sizeof(HashRes) * (size_t)n), total = o_out + up16(out_ints * 4);
    auto grown = [](size_t need, size_t old) { return std::max(need + need / 4, old * 2); };
    if (B->cap < total) {
        const size_t cap = grown(total, B->cap);
        CUH(cudaStreamSynchronize(s));
        cudaFreeHost(B->h); cudaFree(B->d); B->h = nullptr; B->d = nullptr; B->cap = 0;
        CUH(cudaMallocHost(&B->h, cap)); CUH(cudaMalloc(&B->d, cap)); B->cap = cap;
    }
    if (B->fixed_cap < fixed_ints) {
        const size_t cap = grown(fixed_ints, B->fixed_cap);
        CUH(cudaStreamSynchronize(s));
        cudaFree(B->d_fixed); B->d_fixed = nullptr; B->fixed_cap = 0;
        CUH(cudaMalloc(&B->d_fixed, cap * 4)); B->fixed_cap = cap;
    }
    if (!B->d_cursor) CUH(cudaMalloc(&B->d_cursor, sizeof(unsigned long long)));
    memcpy(B->h + o_req, reqs.data(), sizeof(HashReq) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const lb2_hash_task& t = tasks[i];
        uint8_t* a = B->h + o_seq + reqs[(size_t)i].ref_off;
        if (t.ref_len) memcpy(a, t.ref, (size_t)t.ref_len);
        memset(a + t.ref_len, 0, 16);
        uint8_t* b = B->h + o_seq + reqs[(size_t)i].read_off;
        if (t.read_len) memcpy(b, t.read, (size_t)t.read_len);
        memset(b + t.read_len, 0, 16);
    }
    CUH(cudaMemcpyAsync(B->d, B->h, o_res, cudaMemcpyHostToDevice, s));
    size_t want_nodes = std::max(node_guess, (size_t)1 << 16);
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (B->nodes_cap < want_nodes) {
            const size_t cap = grown(want_nodes, B->nodes_cap);
            CUH(cudaStreamSynchronize(s));
            cudaFree(B->d_nodes); B->d_nodes = nullptr; B->nodes_cap = 0;
            CUH(cudaMalloc(&B->d_nodes, cap * 4)); B->nodes_cap = cap;
        }
        CUH(cudaMemsetAsync(B->d_cursor, 0, sizeof(unsigned long long), s));
        const int wpb = 4;
        hash_line_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, s>>>(
            reinterpret_cast<const HashReq*>(B->d + o_req), (int)n, B->d + o_seq, B->d_fixed, B->d_nodes, B->d_cursor,
            (unsigned long long)B->nodes_cap, reinterpret_cast<int32_t*>(B->d + o_out), reinterpret_cast<HashRes*>(B->d + o_res));
        CUH(cudaGetLastError());
        unsigned long long used = 0;
        CUH(cudaMemcpyAsync(&used, B->d_cursor, sizeof used, cudaMemcpyDeviceToHost, s));
        CUH(cudaMemcpyAsync(B->h + o_res, B->d + o_res, total - o_res, cudaMemcpyDeviceToHost, s));
        CUH(cudaStreamSynchronize(s));
        if (used <= B->nodes_cap) break;
        want_nodes = (size_t)used;                       // the pool ran out: every request reported what it needs
        if (attempt == 2) return lb2::set_error("lb2_hash_line_run: node pool could not be sized");
    }
    const HashRes* res = reinterpret_cast<const HashRes*>(B->h + o_res);
    const int32_t* out = reinterpret_cast<const int32_t*>(B->h + o_out);
    for (int64_t i = 0; i < n; ++i) {
        if (res[i].status) return lb2::set_error("lb2_hash_line_run: request %lld failed (status %d)", (long long)i, res[i].status);
        tasks[i].m_len = res[i].m_len; tasks[i].n_hits = res[i].n_nodes;
        if (res[i].m_len > 0) memcpy(tasks[i].line, out + reqs[(size_t)i].out_off, sizeof(int32_t) * 3 * (size_t)res[i].m_len);
    }
    return 0;
}
