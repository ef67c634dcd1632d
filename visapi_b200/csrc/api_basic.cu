// Context lifetime, the a1 scan, batched transitions (a2-a8, a13) and raw rollouts
// (a10) of the C ABI declared in include/rectpack_b200.h.
#include <algorithm>
#include <cstring>

#include "board.cuh"
#include "common.cuh"

namespace bp {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int ensure_scratch(bp_context* ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return BP_OK;
    if (ctx->scratch) {
        BP_CUDA(cudaStreamSynchronize(ctx->stream));
        BP_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    size_t want = std::max(bytes, size_t(1) << 20);
    want = (want + (size_t(1) << 20) - 1) & ~((size_t(1) << 20) - 1);
    BP_CUDA(cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return BP_OK;
}

constexpr int WARPS_PER_BLOCK = 4;
constexpr int BLOCK = WARPS_PER_BLOCK * 32;

// ---------------------------------------------------------------- a1 find_first
__global__ void find_first_kernel(int32_t needle, const int32_t* __restrict__ hay, long long n,
                                  unsigned long long* out)
{
    unsigned long long best = ~0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool hit = (i < n) && (hay[i] == needle);
        const uint32_t b = __ballot_sync(FULLMASK, hit);
        if (b) {   // first hit of this warp in this sweep; later sweeps only have larger indices
            best = (unsigned long long)(base + (threadIdx.x & ~31) + (__ffs(b) - 1));
            break;
        }
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(out, best);
}

// -------------------------------------------------------- a2-a4, a8: lfb / masks
template <int RHO, int W, int EJ>
__global__ void __launch_bounds__(BLOCK)
valid_masks_kernel(int n, int rows, int cols, int wpr, int n_entries,
                   const uint32_t* __restrict__ occ, const uint8_t* __restrict__ tiles,
                   uint8_t* __restrict__ out_mask, int32_t* __restrict__ out_lfb)
{
    const int i = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = lane_id();
    WarpState<RHO, W, EJ> s;
    load_board(s, occ + (size_t)i * rows * wpr, rows, cols, wpr);
    if (tiles) load_entries(s, tiles + (size_t)i * n_entries * 2, n_entries);
    else {
#pragma unroll
        for (int j = 0; j < EJ; ++j) s.ent[j] = 0u;
    }
    int r = -1, c = -1, run = 0;
    const bool found = find_lfb(s, r, c, run);
    if (!found) { r = -1; c = -1; run = 0; }
    if (out_lfb && lane == 0) {
        out_lfb[3 * i + 0] = r;
        out_lfb[3 * i + 1] = c;
        out_lfb[3 * i + 2] = run;
    }
    if (out_mask) {
        uint32_t legal[EJ];
#pragma unroll
        for (int j = 0; j < EJ; ++j) legal[j] = 0u;
        if (found) legal_masks(s, rows, r, run, legal);
#pragma unroll
        for (int j = 0; j < EJ; ++j) {
            const int e = lane + 32 * j;
            if (e < n_entries) out_mask[(size_t)i * n_entries + e] = (legal[j] >> lane) & 1u;
        }
    }
}

// ----------------------------------------------- a5-a7, a13: place + remove pair
// shift the live entries of s to the front of the table (order kept)
template <int RHO, int W, int EJ>
__device__ __forceinline__ void compact_entries(WarpState<RHO, W, EJ>& s, uint16_t* buf /* [32*EJ] smem */)
{
    const int lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < EJ; ++j) buf[lane + 32 * j] = 0;
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const uint32_t m = __ballot_sync(FULLMASK, s.ent[j] != 0u);
        if (s.ent[j] != 0u) buf[base + __popc(m & lt)] = (uint16_t)s.ent[j];
        base += __popc(m);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < EJ; ++j) s.ent[j] = buf[lane + 32 * j];
    __syncwarp();
}

template <int RHO, int W, int EJ>
__global__ void __launch_bounds__(BLOCK)
transitions_kernel(int n, int rows, int cols, int wpr, int n_entries,
                   const uint32_t* __restrict__ occ, const uint8_t* __restrict__ tiles,
                   const int32_t* __restrict__ actions, int compact, uint32_t* __restrict__ out_occ,
                   uint8_t* __restrict__ out_tiles, int32_t* __restrict__ out_verdict,
                   int32_t* __restrict__ out_lfb)
{
    __shared__ uint16_t cbuf[WARPS_PER_BLOCK][32 * EJ];
    const int warp = threadIdx.x >> 5;
    const int i = blockIdx.x * WARPS_PER_BLOCK + warp;
    if (i >= n) return;
    const int lane = lane_id();
    WarpState<RHO, W, EJ> s;
    load_board(s, occ + (size_t)i * rows * wpr, rows, cols, wpr);
    load_entries(s, tiles + (size_t)i * n_entries * 2, n_entries);
    const int n_alive = count_alive(s);
    int r = -1, c = -1, run = 0;
    int verdict;
    const bool found = find_lfb(s, r, c, run);
    if (!found) {
        r = -1; c = -1; run = 0;
        verdict = n_alive == 0 ? BP_ALL_TILES_USED : BP_NO_NEXT_POSITION_TILES_UNUSED;
    } else {
        const int a = actions[i];
        uint32_t v = 0u;
        if (a >= 0 && a < n_entries) {
            uint32_t mine = s.ent[0];
#pragma unroll
            for (int j = 1; j < EJ; ++j) mine = ((a >> 5) == j) ? s.ent[j] : mine;
            v = __shfl_sync(FULLMASK, mine, a & 31);
        }
        const int h = (int)(v & 0xFFu), w = (int)(v >> 8);
        if (v == 0u || w > run || r + h > rows) {
            verdict = BP_TILE_CANNOT_BE_PLACED;
        } else {
            WarpState<RHO, W, EJ> t = s;
            place(t, r, c, h, w);
            if (!kill_pair(t, v)) {
                verdict = BP_PAIR_MISSING;
            } else {
                verdict = BP_PLACED;
                s = t;
                if (compact) compact_entries(s, cbuf[warp]);
            }
        }
    }
    if (lane == 0) {
        out_verdict[i] = verdict;
        if (out_lfb) {
            out_lfb[3 * i + 0] = r;
            out_lfb[3 * i + 1] = c;
            out_lfb[3 * i + 2] = run;
        }
    }
    if (out_occ) store_board(s, out_occ + (size_t)i * rows * wpr, rows, cols, wpr);
    if (out_tiles) store_entries(s, out_tiles + (size_t)i * n_entries * 2, n_entries);
}

// ------------------------------------------------------------ a13/a14 game ended
template <int RHO, int W, int EJ>
__global__ void __launch_bounds__(BLOCK)
game_ended_kernel(int n, int rows, int cols, int wpr, int n_entries,
                  const uint32_t* __restrict__ occ, const uint8_t* __restrict__ tiles,
                  double* __restrict__ out_value)
{
    const int i = blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (i >= n) return;
    WarpState<RHO, W, EJ> s;
    load_board(s, occ + (size_t)i * rows * wpr, rows, cols, wpr);
    load_entries(s, tiles + (size_t)i * n_entries * 2, n_entries);
    const int n_alive = count_alive(s);
    int r, c, run, k = 0;
    if (n_alive != 0 && find_lfb(s, r, c, run)) {
        uint32_t legal[EJ];
        k = legal_masks(s, rows, r, run, legal);
    }
    double value = 0.0;                              // game still running
    if (n_alive == 0 || k == 0) {                    // BinPackLogic.py:91
        int filled = 0;                              // padding is stored as occupied
#pragma unroll
        for (int kk = 0; kk < RHO; ++kk)
#pragma unroll
            for (int q = 0; q < W; ++q) filled += __popc(s.occ[kk][q]);
        filled = __reduce_add_sync(FULLMASK, filled);
        const int total = rows * cols;
        filled -= 32 * RHO * 32 * W - total;         // remove the padding cells
        value = (filled == total) ? 1.0 : (double)filled / (double)total;
    }
    if (lane_id() == 0) out_value[i] = value;
}

// --------------------------------------------------------------- a10 raw rollouts
constexpr int SIMS_PER_TASK = 32;

template <int RHO, int W, int EJ, bool RECORD>
__global__ void __launch_bounds__(BLOCK)
rollouts_kernel(int n_roots, int rows, int cols, int wpr, int n_entries,
                const uint32_t* __restrict__ occ, const uint8_t* __restrict__ tiles, uint32_t seed,
                const uint32_t* __restrict__ streams, const uint32_t* __restrict__ tags,
                uint32_t sim_begin, int n_sims, uint8_t* __restrict__ out_steps,
                uint8_t* __restrict__ out_solved, uint8_t* __restrict__ out_moves)
{
    const int lane = lane_id();
    const int chunks = (n_sims + SIMS_PER_TASK - 1) / SIMS_PER_TASK;
    const long long n_tasks = (long long)n_roots * chunks;
    const long long warp0 = (long long)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * WARPS_PER_BLOCK;
    for (long long task = warp0; task < n_tasks; task += n_warps) {
        const int root = (int)(task / chunks);
        const int chunk = (int)(task % chunks);
        WarpState<RHO, W, EJ> s;
        load_board(s, occ + (size_t)root * rows * wpr, rows, cols, wpr);
        load_entries(s, tiles + (size_t)root * n_entries * 2, n_entries);
        const int n_alive = count_alive(s);
        const uint32_t stream = streams ? streams[root] : (uint32_t)root;
        const uint32_t tag = tags ? tags[root] : 0u;
        const int first = chunk * SIMS_PER_TASK;
        const int count = min(SIMS_PER_TASK, n_sims - first);
        int my_steps = 0, my_solved = 0;
        for (int i = 0; i < count; ++i) {
            bool solved;
            uint8_t* mv = nullptr;
            if (RECORD)
                mv = out_moves + ((size_t)root * n_sims + first + i) * (size_t)(n_entries / 2) * 2;
            const int steps = rollout<RHO, W, EJ, RECORD>(s, rows, n_alive, seed, stream, tag,
                                                          sim_begin + first + i, solved, mv);
            if (lane == i) { my_steps = steps; my_solved = solved ? 1 : 0; }
        }
        if (lane < count) {
            out_steps[(size_t)root * n_sims + first + lane] = (uint8_t)my_steps;
            out_solved[(size_t)root * n_sims + first + lane] = (uint8_t)my_solved;
        }
    }
}

// ------------------------------------------------------- issue-rate micro-benchmark
// Dependency-free integer streams: 8 chains per thread, each step one instruction per chain
// whose second operand is the neighbouring chain (so steps cannot be folded together):
// ALU mode = 8 LOP3 per step (ALU pipe only), MIXED mode = 4 LOP3 + 4 IMAD (ALU + FMA pipes).
// Measures what "issue peak" means for integer code on this chip.
template <bool MIXED>
__global__ void __launch_bounds__(256) issue_microbench_kernel(uint32_t* out, int iters, uint32_t seed)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8u + i;
    const uint32_t k1 = (seed | 1u) + (threadIdx.x << 8);      // per-thread: a register operand
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t n[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MIXED && (i & 1))                                           // IMAD
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(n[i]) : "r"(a[i]), "r"(k1), "r"(a[(i + 1) & 7]));
                else                                                            // LOP3: (a & k1) ^ b
                    asm volatile("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(n[i]) : "r"(a[(i + 1) & 7]), "r"(a[i]), "r"(k1));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = n[i];
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i];
    if (x == 0xDEADBEEFu) out[0] = x;                             // keep the chains alive
}

}  // namespace bp

// =================================================================== C ABI
using namespace bp;

extern "C" {

const char* bp_last_error(void) { return g_error; }

int bp_create(int device_id, bp_context** out)
{
    BP_REQUIRE(out != nullptr, "out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); rectpack_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return BP_ERR_NO_DEVICE;
    }
    BP_REQUIRE(device_id >= 0 && device_id < count, "device_id out of range");
    BP_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    BP_CUDA(cudaGetDeviceProperties(&prop, device_id));
    bp_context* ctx = new bp_context();
    ctx->device = device_id;
    ctx->sm_count = prop.multiProcessorCount;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device_id);
    ctx->sm_clock_khz = khz;
    *out = ctx;
    return BP_OK;
}

int bp_destroy(bp_context* ctx)
{
    if (!ctx) return BP_OK;
    cudaSetDevice(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->arena) cudaFree(ctx->arena);
    for (cudaStream_t st : ctx->side_streams) cudaStreamDestroy(st);
    for (cudaEvent_t e : ctx->side_events) cudaEventDestroy(e);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    delete ctx;
    return BP_OK;
}

int bp_set_stream(bp_context* ctx, void* cuda_stream)
{
    BP_REQUIRE(ctx != nullptr, "ctx is null");
    cudaStream_t next = reinterpret_cast<cudaStream_t>(cuda_stream);
    if (next == ctx->stream) return BP_OK;
    // Work queued on the old stream may still use the context's scratch, arena or tree buffers:
    // the new stream waits for it (an event, no host synchronisation).  The stream belongs to the
    // caller and must be a stream of the context's device.
    BP_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->fork_event) BP_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    BP_CUDA(cudaEventRecord(ctx->fork_event, ctx->stream));
    BP_CUDA(cudaStreamWaitEvent(next, ctx->fork_event, 0));
    ctx->stream = next;
    return BP_OK;
}

int bp_synchronize(bp_context* ctx)
{
    BP_REQUIRE(ctx != nullptr, "ctx is null");
    BP_CUDA(cudaSetDevice(ctx->device));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));
    return BP_OK;
}

int bp_device_info(bp_context* ctx, int* sm_count, int* sm_clock_khz, int64_t* kernel_launches)
{
    BP_REQUIRE(ctx != nullptr, "ctx is null");
    if (sm_count) *sm_count = ctx->sm_count;
    if (sm_clock_khz) *sm_clock_khz = ctx->sm_clock_khz;
    if (kernel_launches) *kernel_launches = ctx->launches;
    return BP_OK;
}

int bp_issue_microbench(bp_context* ctx, double* alu_only_ginst, double* alu_fma_ginst)
{
    BP_REQUIRE(ctx && alu_only_ginst && alu_fma_ginst, "null argument");
    BP_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_scratch(ctx, 1024);
    if (rc) return rc;
    uint32_t* d_out = static_cast<uint32_t*>(ctx->scratch);
    const int iters = 2000, grid = ctx->sm_count * 8, block = 256;
    cudaEvent_t e0, e1;
    BP_CUDA(cudaEventCreate(&e0));
    BP_CUDA(cudaEventCreate(&e1));
    double res[2] = {0.0, 0.0};
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {                       // the last repetition is timed
            BP_CUDA(cudaEventRecord(e0, ctx->stream));
            if (mode == 0) issue_microbench_kernel<false><<<grid, block, 0, ctx->stream>>>(d_out, iters, 12345u);
            else issue_microbench_kernel<true><<<grid, block, 0, ctx->stream>>>(d_out, iters, 12345u);
            BP_CUDA(cudaEventRecord(e1, ctx->stream));
            BP_CUDA(cudaEventSynchronize(e1));
            ctx->launches++;
        }
        float ms = 0.f;
        BP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        // 8 chains x 8 repeats = 64 instructions per iteration per warp
        const double warp_inst = (double)grid * (block / 32) * (double)iters * 64.0;
        res[mode] = warp_inst / (ms * 1e-3) / 1e9;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    BP_CUDA(cudaGetLastError());
    *alu_only_ginst = res[0];
    *alu_fma_ginst = res[1];
    return BP_OK;
}

int bp_find_first(bp_context* ctx, int32_t needle, const int32_t* haystack, int64_t n,
                  int64_t* out_index)
{
    BP_REQUIRE(ctx && out_index && n >= 0 && (haystack || n == 0), "find_first arguments");
    if (n == 0) { *out_index = -1; return BP_OK; }
    BP_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_scratch(ctx, Scratch::align(n * sizeof(int32_t)) + 256);
    if (rc) return rc;
    Scratch sc(ctx);
    int32_t* d_hay = sc.take<int32_t>(n);
    unsigned long long* d_out = sc.take<unsigned long long>(1);
    BP_CUDA(cudaMemcpyAsync(d_hay, haystack, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(cudaMemsetAsync(d_out, 0xFF, sizeof(unsigned long long), ctx->stream));
    const int block = 256;
    const int grid = (int)std::min<int64_t>((n + block - 1) / block, (int64_t)ctx->sm_count * 8);
    find_first_kernel<<<grid, block, 0, ctx->stream>>>(needle, d_hay, n, d_out);
    ctx->launches++;
    BP_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    BP_CUDA(cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));
    *out_index = (h == ~0ull) ? -1 : (int64_t)h;
    return BP_OK;
}

static int check_batch(bp_context* ctx, int n, int rows, int cols, int n_entries)
{
    BP_REQUIRE(ctx != nullptr, "ctx is null");
    BP_REQUIRE(n >= 0, "n < 0");
    BP_REQUIRE(shape_ok(rows, cols, n_entries), "rows/cols must be 1..128, entries 0..128");
    return BP_OK;
}

int bp_valid_masks(bp_context* ctx, int n, int rows, int cols, int n_entries, const uint32_t* occ,
                   const uint8_t* tiles, uint8_t* out_mask, int32_t* out_lfb)
{
    int rc = check_batch(ctx, n, rows, cols, n_entries);
    if (rc) return rc;
    BP_REQUIRE(occ && (tiles || !out_mask), "occ/tiles null");
    if (n == 0) return BP_OK;
    BP_CUDA(cudaSetDevice(ctx->device));
    const int wpr = (cols + 31) / 32;
    const size_t occ_n = (size_t)n * rows * wpr, tiles_n = (size_t)n * n_entries * 2;
    rc = ensure_scratch(ctx, Scratch::align(occ_n * 4) + Scratch::align(tiles_n) +
                                 Scratch::align((size_t)n * n_entries) + Scratch::align((size_t)n * 12) + 1024);
    if (rc) return rc;
    Scratch sc(ctx);
    uint32_t* d_occ = sc.take<uint32_t>(occ_n);
    uint8_t* d_tiles = sc.take<uint8_t>(tiles_n + 2);
    uint8_t* d_mask = sc.take<uint8_t>((size_t)n * n_entries + 1);
    int32_t* d_lfb = sc.take<int32_t>((size_t)n * 3);
    BP_CUDA(cudaMemcpyAsync(d_occ, occ, occ_n * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (tiles && tiles_n)
        BP_CUDA(cudaMemcpyAsync(d_tiles, tiles, tiles_n, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    dispatch_class(shape_class(rows, cols, n_entries), [&](auto R, auto Wc, auto E) {
        valid_masks_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
            <<<grid, BLOCK, 0, ctx->stream>>>(n, rows, cols, wpr, n_entries, d_occ,
                                              tiles ? d_tiles : nullptr,
                                              out_mask ? d_mask : nullptr, d_lfb);
    });
    ctx->launches++;
    BP_CUDA(cudaGetLastError());
    if (out_mask && n_entries)
        BP_CUDA(cudaMemcpyAsync(out_mask, d_mask, (size_t)n * n_entries, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_lfb)
        BP_CUDA(cudaMemcpyAsync(out_lfb, d_lfb, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));
    return BP_OK;
}

int bp_lfb(bp_context* ctx, int n, int rows, int cols, const uint32_t* occ, int32_t* out_lfb)
{
    BP_REQUIRE(out_lfb != nullptr, "out_lfb is null");
    return bp_valid_masks(ctx, n, rows, cols, 0, occ, nullptr, nullptr, out_lfb);
}

int bp_transitions(bp_context* ctx, int n, int rows, int cols, int n_entries, const uint32_t* occ,
                   const uint8_t* tiles, const int32_t* actions, int compact, uint32_t* out_occ,
                   uint8_t* out_tiles, int32_t* out_verdict, int32_t* out_lfb)
{
    int rc = check_batch(ctx, n, rows, cols, n_entries);
    if (rc) return rc;
    BP_REQUIRE(occ && tiles && actions && out_verdict, "occ/tiles/actions/out_verdict null");
    BP_REQUIRE(n_entries >= 1, "n_entries < 1");
    if (n == 0) return BP_OK;
    BP_CUDA(cudaSetDevice(ctx->device));
    const int wpr = (cols + 31) / 32;
    const size_t occ_n = (size_t)n * rows * wpr, tiles_n = (size_t)n * n_entries * 2;
    rc = ensure_scratch(ctx, 2 * Scratch::align(occ_n * 4) + 2 * Scratch::align(tiles_n) +
                                 3 * Scratch::align((size_t)n * 12) + 2048);
    if (rc) return rc;
    Scratch sc(ctx);
    uint32_t* d_occ = sc.take<uint32_t>(occ_n);
    uint32_t* d_occ2 = sc.take<uint32_t>(occ_n);
    uint8_t* d_tiles = sc.take<uint8_t>(tiles_n);
    uint8_t* d_tiles2 = sc.take<uint8_t>(tiles_n);
    int32_t* d_act = sc.take<int32_t>(n);
    int32_t* d_ver = sc.take<int32_t>(n);
    int32_t* d_lfb = sc.take<int32_t>((size_t)n * 3);
    BP_CUDA(cudaMemcpyAsync(d_occ, occ, occ_n * 4, cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(cudaMemcpyAsync(d_tiles, tiles, tiles_n, cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(cudaMemcpyAsync(d_act, actions, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    dispatch_class(shape_class(rows, cols, n_entries), [&](auto R, auto Wc, auto E) {
        transitions_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
            <<<grid, BLOCK, 0, ctx->stream>>>(n, rows, cols, wpr, n_entries, d_occ, d_tiles, d_act,
                                              compact, out_occ ? d_occ2 : nullptr,
                                              out_tiles ? d_tiles2 : nullptr, d_ver, d_lfb);
    });
    ctx->launches++;
    BP_CUDA(cudaGetLastError());
    if (out_occ) BP_CUDA(cudaMemcpyAsync(out_occ, d_occ2, occ_n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_tiles) BP_CUDA(cudaMemcpyAsync(out_tiles, d_tiles2, tiles_n, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(cudaMemcpyAsync(out_verdict, d_ver, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_lfb) BP_CUDA(cudaMemcpyAsync(out_lfb, d_lfb, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));
    return BP_OK;
}

int bp_game_ended(bp_context* ctx, int n, int rows, int cols, int n_entries, const uint32_t* occ,
                  const uint8_t* tiles, double* out_value)
{
    int rc = check_batch(ctx, n, rows, cols, n_entries);
    if (rc) return rc;
    BP_REQUIRE(occ && tiles && out_value, "occ/tiles/out_value null");
    BP_REQUIRE(n_entries >= 1, "n_entries < 1");
    if (n == 0) return BP_OK;
    BP_CUDA(cudaSetDevice(ctx->device));
    const int wpr = (cols + 31) / 32;
    const size_t occ_n = (size_t)n * rows * wpr, tiles_n = (size_t)n * n_entries * 2;
    rc = ensure_scratch(ctx, Scratch::align(occ_n * 4) + Scratch::align(tiles_n) +
                                 Scratch::align((size_t)n * 8) + 1024);
    if (rc) return rc;
    Scratch sc(ctx);
    uint32_t* d_occ = sc.take<uint32_t>(occ_n);
    uint8_t* d_tiles = sc.take<uint8_t>(tiles_n);
    double* d_val = sc.take<double>(n);
    BP_CUDA(cudaMemcpyAsync(d_occ, occ, occ_n * 4, cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(cudaMemcpyAsync(d_tiles, tiles, tiles_n, cudaMemcpyHostToDevice, ctx->stream));
    const int grid = (n + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK;
    dispatch_class(shape_class(rows, cols, n_entries), [&](auto R, auto Wc, auto E) {
        game_ended_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
            <<<grid, BLOCK, 0, ctx->stream>>>(n, rows, cols, wpr, n_entries, d_occ, d_tiles, d_val);
    });
    ctx->launches++;
    BP_CUDA(cudaGetLastError());
    BP_CUDA(cudaMemcpyAsync(out_value, d_val, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));
    return BP_OK;
}

int bp_rollouts(bp_context* ctx, int n_roots, int rows, int cols, int n_entries,
                const uint32_t* occ, const uint8_t* tiles, uint32_t seed, const uint32_t* streams,
                const uint32_t* tags, uint32_t sim_begin, int n_sims, uint8_t* out_steps,
                uint8_t* out_solved, uint8_t* out_moves)
{
    int rc = check_batch(ctx, n_roots, rows, cols, n_entries);
    if (rc) return rc;
    BP_REQUIRE(occ && tiles && out_steps && out_solved, "occ/tiles/out_steps/out_solved null");
    BP_REQUIRE(n_sims >= 0 && n_entries >= 1, "n_sims < 0 or n_entries < 1");
    if (n_roots == 0 || n_sims == 0) return BP_OK;
    BP_CUDA(cudaSetDevice(ctx->device));
    const int wpr = (cols + 31) / 32;
    const size_t occ_n = (size_t)n_roots * rows * wpr, tiles_n = (size_t)n_roots * n_entries * 2;
    const size_t res_n = (size_t)n_roots * n_sims;
    const size_t mv_n = out_moves ? res_n * (size_t)(n_entries / 2) * 2 : 0;
    rc = ensure_scratch(ctx, Scratch::align(occ_n * 4) + Scratch::align(tiles_n) +
                                 2 * Scratch::align((size_t)n_roots * 4) + 2 * Scratch::align(res_n) +
                                 Scratch::align(mv_n) + 2048);
    if (rc) return rc;
    Scratch sc(ctx);
    uint32_t* d_occ = sc.take<uint32_t>(occ_n);
    uint8_t* d_tiles = sc.take<uint8_t>(tiles_n);
    uint32_t* d_streams = sc.take<uint32_t>(n_roots);
    uint32_t* d_tags = sc.take<uint32_t>(n_roots);
    uint8_t* d_steps = sc.take<uint8_t>(res_n);
    uint8_t* d_solved = sc.take<uint8_t>(res_n);
    uint8_t* d_moves = sc.take<uint8_t>(mv_n + 1);
    BP_CUDA(cudaMemcpyAsync(d_occ, occ, occ_n * 4, cudaMemcpyHostToDevice, ctx->stream));
    BP_CUDA(cudaMemcpyAsync(d_tiles, tiles, tiles_n, cudaMemcpyHostToDevice, ctx->stream));
    if (streams) BP_CUDA(cudaMemcpyAsync(d_streams, streams, (size_t)n_roots * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (tags) BP_CUDA(cudaMemcpyAsync(d_tags, tags, (size_t)n_roots * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (out_moves) BP_CUDA(cudaMemsetAsync(d_moves, 0, mv_n, ctx->stream));
    const long long chunks = (n_sims + SIMS_PER_TASK - 1) / SIMS_PER_TASK;
    const long long n_tasks = (long long)n_roots * chunks;
    const int grid = (int)std::min<long long>((n_tasks + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK,
                                              (long long)ctx->sm_count * 16);
    dispatch_class(shape_class(rows, cols, n_entries), [&](auto R, auto Wc, auto E) {
        constexpr int r = decltype(R)::value, w = decltype(Wc)::value, e = decltype(E)::value;
        if (out_moves)
            rollouts_kernel<r, w, e, true><<<grid, BLOCK, 0, ctx->stream>>>(
                n_roots, rows, cols, wpr, n_entries, d_occ, d_tiles, seed,
                streams ? d_streams : nullptr, tags ? d_tags : nullptr, sim_begin, n_sims, d_steps,
                d_solved, d_moves);
        else
            rollouts_kernel<r, w, e, false><<<grid, BLOCK, 0, ctx->stream>>>(
                n_roots, rows, cols, wpr, n_entries, d_occ, d_tiles, seed,
                streams ? d_streams : nullptr, tags ? d_tags : nullptr, sim_begin, n_sims, d_steps,
                d_solved, nullptr);
    });
    ctx->launches++;
    BP_CUDA(cudaGetLastError());
    BP_CUDA(cudaMemcpyAsync(out_steps, d_steps, res_n, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(cudaMemcpyAsync(out_solved, d_solved, res_n, cudaMemcpyDeviceToHost, ctx->stream));
    if (out_moves) BP_CUDA(cudaMemcpyAsync(out_moves, d_moves, mv_n, cudaMemcpyDeviceToHost, ctx->stream));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));
    return BP_OK;
}

}  // extern "C"
