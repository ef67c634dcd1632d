// Host-side plumbing shared by the translation units of librectpack_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "../../include/rectpack_b200.h"

struct bp_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int sm_clock_khz = 0;
    int64_t launches = 0;
    // grow-only device scratch for the batched calls
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // grow-only arena lent to one one-shot bp_flat_search at a time
    void* arena = nullptr;
    size_t arena_bytes = 0;
    bool arena_busy = false;
    // side streams / events of the flat search (one per shape class in a batch), created on
    // first use and kept: creating and destroying them cost ~1 ms per one-shot search
    std::vector<cudaStream_t> side_streams;
    std::vector<cudaEvent_t> side_events;
    cudaEvent_t fork_event = nullptr;
    // resident blocks per SM of the search kernels, by class (0 = not asked yet)
    int occ_rollout[256] = {0};
    int occ_lanes[256] = {0};
    int occ_resident[32] = {0};
};

namespace bp {
// k-th side stream / event of the context (created on first use); nullptr on failure
// BP_STREAM_PRIORITY=1: side stream k gets priority rank k (the search gives stream 0 to the chain
// with the most work, whose launch sequence is the critical path).  Measured on B200
// (profiles/r2/shard_sweep_prio.txt): 2 % better on a 125-problem shard, 1.2 % worse on the full
// set -- off by default.
inline cudaStream_t side_stream(bp_context* ctx, size_t k)
{
    while (ctx->side_streams.size() <= k) {
        cudaStream_t st = nullptr;
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);      // numerically lower = higher priority
        static const bool flat = !(getenv("BP_STREAM_PRIORITY") && atoi(getenv("BP_STREAM_PRIORITY")) != 0);
        const int want = flat ? least : greatest + (int)ctx->side_streams.size();
        const int prio = want > least ? least : want;
        if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio) != cudaSuccess) return nullptr;
        ctx->side_streams.push_back(st);
    }
    return ctx->side_streams[k];
}
inline cudaEvent_t side_event(bp_context* ctx, size_t k)
{
    while (ctx->side_events.size() <= k) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        ctx->side_events.push_back(e);
    }
    return ctx->side_events[k];
}
}  // namespace bp

namespace bp {

void set_error(const char* fmt, ...);

#define BP_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            bp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                          __FILE__, __LINE__);                                          \
            return BP_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define BP_REQUIRE(cond, msg)                                                           \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            bp::set_error("invalid argument: %s", msg);                                 \
            return BP_ERR_INVALID;                                                      \
        }                                                                               \
    } while (0)

// carve 256-byte aligned slices out of the context scratch buffer
struct Scratch {
    bp_context* ctx;
    size_t used = 0;
    explicit Scratch(bp_context* c) : ctx(c) {}
    static size_t align(size_t n) { return (n + 255) & ~size_t(255); }
    template <typename T>
    T* take(size_t count)
    {
        T* p = reinterpret_cast<T*>(static_cast<char*>(ctx->scratch) + used);
        used += align(count * sizeof(T));
        return p;
    }
};
int ensure_scratch(bp_context* ctx, size_t bytes);

// ---- shape classes -------------------------------------------------------------
// Kernels are templated on RHO (rows per lane), W (words per row), EJ (entries per
// lane), each in {1, 2, 4}; a shape is rounded up to the next class (extra rows,
// columns and entry slots are padding).
inline int round_class(int need) { return need <= 1 ? 1 : (need <= 2 ? 2 : 4); }
inline int class_index_of(int v) { return v == 1 ? 0 : (v == 2 ? 1 : 2); }
inline int shape_class(int rows, int cols, int n_entries)
{
    const int rho = round_class((rows + 31) / 32);
    const int w = round_class((cols + 31) / 32);
    const int ej = round_class((n_entries + 31) / 32);
    return (class_index_of(rho) * 3 + class_index_of(w)) * 3 + class_index_of(ej);
}
constexpr int N_CLASSES = 27;

template <int V>
using IC = std::integral_constant<int, V>;

// f(IC<RHO>, IC<W>, IC<EJ>) for class index cls
template <typename F>
inline void dispatch_class(int cls, F&& f)
{
    const int ej = cls % 3, w = (cls / 3) % 3, rho = cls / 9;
    auto with_ej = [&](auto R, auto Wc) {
        switch (ej) {
            case 0: f(R, Wc, IC<1>{}); break;
            case 1: f(R, Wc, IC<2>{}); break;
            default: f(R, Wc, IC<4>{}); break;
        }
    };
    auto with_w = [&](auto R) {
        switch (w) {
            case 0: with_ej(R, IC<1>{}); break;
            case 1: with_ej(R, IC<2>{}); break;
            default: with_ej(R, IC<4>{}); break;
        }
    };
    switch (rho) {
        case 0: with_w(IC<1>{}); break;
        case 1: with_w(IC<2>{}); break;
        default: with_w(IC<4>{}); break;
    }
}

inline bool shape_ok(int rows, int cols, int n_entries)
{
    return rows >= 1 && rows <= BP_MAX_ROWS && cols >= 1 && cols <= BP_MAX_COLS &&
           n_entries >= 0 && n_entries <= BP_MAX_ENTRIES;
}

}  // namespace bp
