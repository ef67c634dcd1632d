// Flat Monte-Carlo search (CustomMCTS.predict, binpack/MCTS.py:44-149) resident on the
// device, host side: many problems in flight, no host round trip until the results are read.
// Two forms of one search (bp_search_run picks, results are identical):
//   launch-per-depth   every depth = one rollout launch + one advance launch per shape class,
//                      each class's sequence on its own stream, the whole sequence one CUDA graph
//                      from the second run on (search_kernels.cuh; on the skyline state:
//                      search_resident.cuh);
//   resident           ONE launch plays every depth of every problem: unit ring, server SMs that
//                      advance problems, in-kernel peer-memory exchange for partitioned searches
//                      (search_resident.cuh, search_queue.cuh, resident_k*.cu).
//
// HBM layout (P problems, ES = entry stride, RS/WS = max rows / words per row of the batch):
//   desc[P]            rows, cols, n_entries, class, Philox stream           (read-only)
//   occ[P][RS][WS]     current board of each problem (row-major words)
//   tiles[P][ES][2]    current live tile list, compacted, (0,0) padded
//   lfb[P]             (row, col, free run, live entries) of the current board
//   legal[P]           4 x 32-bit legal-entry masks of the current board
//   rec[P][ES][NC]     one 8-byte record per (child, chunk of 32 rollouts)
//   stats[P][ES]       per-child statistics (root-parallel exchange buffer)
//   child_list[]       (problem, entry) of every child of the current depth, per class
//   pstate[P]          skyline state: one 128-byte record per problem (search_types.cuh)
//   slots / urec / ulist / done_slots   unit ring, unit records, legal-entry lists, advance queue
// A rollout never touches HBM except to read its root (once per 32 rollouts) and to
// write 8 bytes per 32 rollouts; it is bound by instruction issue, not bandwidth.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "search_kernels.cuh"
#include "resident_variant.cuh"
#include "search_types.cuh"

// ================================================================= host side
using namespace bp;

// lane-kernel instantiation of a search class: NB height planes, W words per plane, EW words
// of entry mask
template <typename F>
static void dispatch_lanes(int scls, F&& f)
{
    const int cls = scls % N_CLASSES, nb = 5 + (scls / N_CLASSES) % 4;
    const bool three_words = scls / N_CLASSES >= 4;
    const int ej = cls % 3, w = (cls / 3) % 3;
    auto with_ew = [&](auto NBc, auto Wc) {
        if (ej <= 1) f(NBc, Wc, IC<2>{}); else f(NBc, Wc, IC<4>{});
    };
    auto with_w = [&](auto NBc) {
        switch (w) {
            case 0: with_ew(NBc, IC<1>{}); break;
            case 1: with_ew(NBc, IC<2>{}); break;
            default:
                if (three_words) with_ew(NBc, IC<3>{}); else with_ew(NBc, IC<4>{});
                break;
        }
    };
    switch (nb) {
        case 5: with_w(IC<5>{}); break;
        case 6: with_w(IC<6>{}); break;
        case 7: with_w(IC<7>{}); break;
        default: with_w(IC<8>{}); break;
    }
}

// resident-kernel variant: 0 = <2,2>, 1 = <2,4>, 2 = <4,2>, 3 = <4,4> (widest plane words, entry-mask words)
template <typename F>
static void dispatch_resident(int variant, F&& f)
{
    switch (variant) {
        case 0: f(IC<2>{}, IC<2>{}); break;
        case 1: f(IC<2>{}, IC<4>{}); break;
        case 2: f(IC<4>{}, IC<2>{}); break;
        default: f(IC<4>{}, IC<4>{}); break;
    }
}

struct bp_search {
    bp_context* ctx = nullptr;
    SearchParams p{};
    int max_depth = 0;
    int cur_depth = 0;         // next depth to play
    bool ready = false;
    std::vector<ProblemDesc> h_desc;
    // class bookkeeping
    std::vector<int> classes;                   // classes present
    // indexed by CHAIN (class + FS_CLASSES * sub-chain); `classes` lists the chains present
    std::vector<int> cls_count = std::vector<int>(FS_CHAINS, 0);
    std::vector<int> cls_prob_off = std::vector<int>(FS_CHAINS, 0);    // offset into d_cls_problems
    std::vector<int> cls_list_off = std::vector<int>(FS_CHAINS, 0);    // offset into child_list
    std::vector<int> grid_rollout = std::vector<int>(FS_CHAINS, 0);
    std::vector<int> grid_lanes = std::vector<int>(FS_CHAINS, 0);
    // launch-per-depth form on the skyline state: per chain a region of the unit ring used as a
    // plain list, per depth and chain the number of units in it
    std::vector<size_t> chain_unit_off = std::vector<size_t>(FS_CHAINS, 0);
    unsigned long long* d_sky_count = nullptr;   // [MAX_DEPTHS + 1][FS_CHAINS]
    bool sky_stepped = false;                   // the last stepped run used the skyline kernels
    int* d_cls_problems = nullptr;
    ProblemDesc* d_desc = nullptr;
    uint32_t* d_init_occ = nullptr;            // initial boards (NULL: empty boards)
    uint8_t* d_init_tiles = nullptr;           // initial tile tables
    char* arena = nullptr;
    size_t arena_used = 0, arena_bytes = 0;
    // regions of the arena (see search_create): uploaded inputs, zeroed / -1 state, result block
    char *up_begin = nullptr, *up_end = nullptr, *zero_begin = nullptr, *zero_end = nullptr;
    char *res_end = nullptr, *ff_begin = nullptr, *ff_end = nullptr;
    std::vector<char> h_upload, h_results;      // host images of the upload / result block
    bool in_profiled_run = false;               // bp_search_run is stepping the depths itself
#ifdef BP_TIMELINE
    unsigned long long* d_timeline = nullptr;   // debug build: per-launch warp time stamps
#endif
    bool arena_from_ctx = false;
    size_t counters_bytes = 0;
    // optional per-depth CUDA-event timing of the rollout and advance launches
    bool profiling = false;
    std::vector<cudaEvent_t> ev;               // 3 per depth: before rollouts, between, after advance
    int ev_depths = 0;
    // one stream per search class: classes hold disjoint problems, so their depth loops run
    // concurrently and one class's advance (latency-bound) hides under another's rollouts
    std::vector<cudaStream_t> cls_stream;
    std::vector<cudaEvent_t> cls_done;
    std::vector<int> cls_depth;
    cudaEvent_t fork_ev = nullptr;
    // the forked launch-per-depth run captured once as a CUDA graph (from the second run on): one
    // graph launch replaces ~2 x depths x chains kernel launches and their stream fork / join
    cudaGraphExec_t graph_exec = nullptr;
    int graph_launches = 0;                     // kernel nodes in the graph
    int graph_key = 0;                          // which form the graph holds (0: launch-per-depth, 1 + K: hybrid)
    int n_runs = 0;
    // resident search (one launch runs every depth of every problem): queue, ring, variant
    int mode = BP_SEARCH_AUTO;
    bool ran_resident = false;
    bool seeded = false;                        // depth 0 is on the queue
    bool queue_dirty = true;                    // ring / tickets must be cleared before the next run
    ResidentParams x{};
    int variant = 0;
    int grid_resident = 0;
    size_t ring_cap = 0;
    size_t list_base = 0;                       // first slot of the launch-per-depth unit lists
    int hybrid_depths = 0;                      // depths the last hybrid run played launch-per-depth
    UnitQueue h_queue{};                        // header as read back by bp_search_results
    std::vector<uint8_t> h_home;                // home queue by SM arrival rank (uploaded)
    uint8_t* d_home = nullptr;
    // root-parallel exchange over peer memory (resident search with n_ranks > 1)
    void* d_exchange = nullptr;                 // own buffer: gathered[4][R][P][ES] + arrived[P][8] + peer table
    size_t exchange_bytes = 0, exchange_flags_off = 0, exchange_table_off = 0;
    bool exchange_connected = false;
    std::vector<void*> peer_opened;             // cudaIpcOpenMemHandle mappings to close
    ~bp_search()
    {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        // streams and fork/join events belong to the context (side_stream / side_event)
    }
};

// All device arrays of a search live in one arena: a single cudaMalloc (or, for the one-shot
// bp_flat_search, a grow-only arena cached in the context, so repeated calls allocate nothing).
// dev_alloc is called twice over the same sequence: a dry run that only sizes the arena, then
// the run that hands out the slices.
template <typename T>
static int dev_alloc(bp_search* s, T** out, size_t count)
{
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + 255) & ~size_t(255);
    *out = s->arena ? reinterpret_cast<T*>(s->arena + s->arena_used) : nullptr;
    s->arena_used += bytes;
    return BP_OK;
}

extern "C" {

int bp_search_destroy(bp_search* s)
{
    if (!s) return BP_OK;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    for (void* mapped : s->peer_opened) cudaIpcCloseMemHandle(mapped);
    if (s->d_exchange) cudaFree(s->d_exchange);
    if (s->arena_from_ctx) s->ctx->arena_busy = false;
    else if (s->arena) cudaFree(s->arena);
    delete s;
    return BP_OK;
}

}  // extern "C"

static int search_create(bp_context* ctx, int n_problems, const int32_t* rows, const int32_t* cols,
                         const int32_t* n_entries, const uint8_t* tiles, int entry_stride,
                         const uint32_t* boards, const uint32_t* streams, int n_sim, int strategy,
                         uint32_t seed, bool use_ctx_arena, bp_search** out)
{
    BP_REQUIRE(ctx && out && rows && cols && n_entries && tiles, "null argument");
    BP_REQUIRE(n_problems >= 1 && n_problems < (1 << 24), "n_problems must be 1..2^24-1");
    BP_REQUIRE(entry_stride >= 2 && entry_stride <= BP_MAX_ENTRIES && entry_stride % 2 == 0,
               "entry_stride must be even, 2..128");
    BP_REQUIRE(n_sim >= 1 && n_sim <= (1 << 24), "n_sim must be 1..2^24");
    BP_REQUIRE(strategy == BP_MAX_DEPTH || strategy == BP_AVG_DEPTH, "strategy");
    int rs = 0, max_cols = 0, max_n = 0;
    for (int i = 0; i < n_problems; ++i) {
        BP_REQUIRE(shape_ok(rows[i], cols[i], n_entries[i]) && n_entries[i] <= entry_stride &&
                       n_entries[i] % 2 == 0,
                   "problem shape out of range (rows, cols 1..128; entries even, <= entry_stride)");
        rs = std::max(rs, rows[i]);
        max_cols = std::max(max_cols, cols[i]);
        max_n = std::max(max_n, n_entries[i] / 2);
    }
    BP_CUDA(cudaSetDevice(ctx->device));
    bp_search* s = new bp_search();
    s->ctx = ctx;
    SearchParams& p = s->p;
    p.n_problems = n_problems;
    p.es = entry_stride;
    p.rs = rs;
    p.ws = (max_cols + 31) / 32;
    p.n_sim = n_sim;
    p.strategy = strategy;
    p.seed = seed;
    p.sim_begin = 0;
    p.n_local = n_sim;
    p.rank = 0;
    p.n_ranks = 1;
    p.n_chunks = (n_sim + CHUNK - 1) / CHUNK;
    s->max_depth = max_n;

    s->h_desc.resize(n_problems);
    std::vector<std::vector<int>> by_class(FS_CHAINS);
    std::vector<int> class_seen(FS_CLASSES, 0);
    for (int i = 0; i < n_problems; ++i) {
        ProblemDesc& d = s->h_desc[i];
        d.rows = rows[i]; d.cols = cols[i]; d.n_entries = n_entries[i];
        const int nb = rows[i] <= 31 ? 5 : (rows[i] <= 63 ? 6 : (rows[i] <= 127 ? 7 : 8));
        const int three_words = (cols[i] > 64 && cols[i] <= 96) ? 4 : 0;
        // rows per lane of the warp-per-board kernels follow the plane count (31 / 63 / 128 rows),
        // not ceil(rows / 32): boards of exactly 32 or 64 rows would otherwise form classes of
        // their own, each with its own chain of launches
        const int rows_of_class = nb == 5 ? 32 : (nb == 6 ? 64 : 128);
        d.cls = (nb - 5 + three_words) * N_CLASSES + shape_class(rows_of_class, cols[i], n_entries[i]);
        d.stream = streams ? streams[i] : (uint32_t)i;
        d.pad0 = d.pad1 = d.pad2 = 0;
        class_seen[d.cls]++;
    }
    // sub-chains per class: a batch too small to fill the GPU with one launch sequence per class
    // gets several (problems dealt round-robin), so that about 8 sequences are in flight
    int n_present = 0;
    for (int c = 0; c < FS_CLASSES; ++c) n_present += class_seen[c] ? 1 : 0;
    int subs = 1;
    // (measured on B200, profiles/r2/shard_sweep_chains.txt: 1 / 2 / 4 / 8 chains per class take
    // 2.09 / 2.19 / 2.27 / 3.6 ms on a 125-problem shard and 11.28 / 11.63 / 12.17 / 13.36 ms on the
    // full set -- every extra launch sequence costs more in launches than its overlap returns, so
    // one chain per class is the default and BP_SUBCHAINS only an experiment knob)
    (void)n_present;
    if (const char* env = getenv("BP_SUBCHAINS")) subs = std::max(1, std::min(FS_MAX_SUBCHAINS, atoi(env)));
    {
        std::vector<int> dealt(FS_CLASSES, 0);
        for (int i = 0; i < n_problems; ++i) {
            const int c = s->h_desc[i].cls;
            const int k = std::min(subs, std::max(1, class_seen[c] / 4));     // at least ~4 problems per chain
            by_class[c + FS_CLASSES * (dealt[c]++ % k)].push_back(i);
        }
    }

    std::vector<int> cls_problems;
    int list_off = 0;
    for (int c = 0; c < FS_CHAINS; ++c) {
        s->cls_count[c] = (int)by_class[c].size();
        s->cls_prob_off[c] = (int)cls_problems.size();
        s->cls_list_off[c] = list_off;
        if (s->cls_count[c]) {
            s->classes.push_back(c);
            cls_problems.insert(cls_problems.end(), by_class[c].begin(), by_class[c].end());
            list_off += s->cls_count[c] * entry_stride;
        }
    }

    // lane-per-rollout kernel: needs the skyline invariant (empty initial boards) and equal
    // tiles adjacent in every list (true for the reference's sorted lists);
    // BP_ROLLOUT=warp forces the warp-per-board kernel
    bool lanes_ok = boards == nullptr;
    if (const char* env = getenv("BP_ROLLOUT")) lanes_ok = lanes_ok && strcmp(env, "warp") != 0;
    {   // equal tiles adjacent in the list?  closed_in[v] = generation of the problem in which the
        // run of value v has ended; meeting v again in the same problem breaks the rule
        static thread_local std::vector<uint32_t> closed_in(1u << 16, 0u);
        static thread_local uint32_t generation = 0;
        for (int i = 0; i < n_problems && lanes_ok; ++i) {
            if (++generation == 0u) { std::fill(closed_in.begin(), closed_in.end(), 0u); generation = 1; }
            const uint8_t* t = tiles + (size_t)i * entry_stride * 2;
            int prev = -1;
            for (int e = 0; e < n_entries[i]; ++e) {
                const int v = t[2 * e] | (t[2 * e + 1] << 8);
                if (v == 0 || v == prev) continue;
                if (closed_in[v] == generation) { lanes_ok = false; break; }
                if (prev >= 0) closed_in[prev] = generation;
                prev = v;
            }
        }
    }
    p.use_lanes = lanes_ok ? 1 : 0;

    // resident search: kernel variant by the widest plane / entry mask of the batch; ring sized
    // for every unit that can be outstanding at once (each live entry of each problem split into
    // at most gpc_max units), at most 2^20 slots of 8 bytes
    if (p.use_lanes && n_problems < (1 << RS_PROBLEM_BITS)) {
        size_t sum_entries = 0;
        for (int i = 0; i < n_problems; ++i) sum_entries += (size_t)n_entries[i];
        const bool wide_entries = 2 * max_n > 64;
        s->variant = (p.ws > 2 ? 2 : 0) + (wide_entries ? 1 : 0);
        const size_t gpc_cap = (size_t)std::min(p.n_chunks, 256);
        size_t cap = 1024;
        while (cap < sum_entries * gpc_cap && cap < (size_t(1) << 20)) cap <<= 1;
        while (cap < sum_entries) cap <<= 1;
        s->x.gpc_max = (int)std::max<size_t>(1, std::min(gpc_cap, cap / std::max<size_t>(sum_entries, 1)));
        // One unit queue per playout class (NB planes x W words x EW entry words), heaviest first;
        // classes beyond RS_MAX_QUEUES - 1 share the last queue.  Each queue has its own ring, sized
        // for every unit its problems can have outstanding.
        std::vector<long long> weight(96, 0);
        std::vector<size_t> entries_of(96, 0);
        auto play_class = [&](int i) {
            const int nbi = rows[i] <= 31 ? 0 : (rows[i] <= 63 ? 1 : 2);
            return (nbi * 4 + ((cols[i] + 31) / 32 - 1)) * 2 + (n_entries[i] <= 64 ? 0 : 1);
        };
        for (int i = 0; i < n_problems; ++i) {
            const int w = (cols[i] + 31) / 32;
            weight[play_class(i)] += (long long)rows[i] * cols[i] * (w == 1 ? 10 : (w == 2 ? 13 : (w == 3 ? 18 : 22)));
            entries_of[play_class(i)] += (size_t)n_entries[i];
        }
        std::vector<int> order;
        for (int c = 0; c < 96; ++c) if (entries_of[c]) order.push_back(c);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return weight[a] > weight[b]; });
        std::vector<int> queue_of(96, 0);
        s->x.n_queues = (int)std::min<size_t>(order.size(), RS_MAX_QUEUES);
        std::vector<long long> qweight(RS_MAX_QUEUES, 0);
        std::vector<size_t> qentries(RS_MAX_QUEUES, 0);
        for (size_t r = 0; r < order.size(); ++r) {
            const int k = (int)std::min<size_t>(r, RS_MAX_QUEUES - 1);
            queue_of[order[r]] = k;
            qweight[k] += weight[order[r]];
            qentries[k] += entries_of[order[r]];
        }
        for (int i = 0; i < n_problems; ++i) s->h_desc[i].pad0 = queue_of[play_class(i)];
        size_t ring_total = 0;
        for (int k = 0; k < RS_MAX_QUEUES; ++k) {
            size_t ck = 64;
            while (ck < qentries[k] * (size_t)s->x.gpc_max) ck <<= 1;
            s->x.ring_off[k] = (unsigned)ring_total;
            s->x.ring_mask[k] = (unsigned)(ck - 1);
            if (k < s->x.n_queues) ring_total += ck;
        }
        // (the launch-per-depth form on the skyline state keeps its chains' unit lists behind the
        // rings: sum_entries * gpc_max slots.  Separate memory, because the hybrid form plays the first
        // depths from the lists and hands over to the resident kernel chain by chain.)
        s->ring_cap = ring_total + sum_entries * (size_t)s->x.gpc_max;
        s->list_base = ring_total;
        // home queue of the SM with arrival rank r: the playing SMs (ranks n_server_sms + 1 ..) are
        // dealt to the queues in proportion to their weight (area x a cost factor per plane width)
        s->h_home.assign(256, 0);
        s->x.seed_in_flight = n_problems;
        size_t done_cap = 64;
        while (done_cap < (size_t)n_problems) done_cap <<= 1;
        s->x.done_mask = (unsigned)(done_cap - 1);
        // SMs that advance problems instead of playing (32 warps each)
        s->x.n_server_sms = n_problems >= 64 ? 2 : 1;
        if (const char* env = getenv("BP_RESIDENT_SERVERS")) s->x.n_server_sms = std::max(1, atoi(env));
        {
            long long total_w = 0;
            for (int k = 0; k < s->x.n_queues; ++k) total_w += std::max<long long>(qweight[k], 1);
            const int playing = std::max(1, ctx->sm_count - s->x.n_server_sms);
            // arrival rank r -> the j-th playing SM -> the queue whose cumulative share covers the
            // middle of that SM's slice
            for (int r = 0; r < 256; ++r) {
                const int j = ((r - s->x.n_server_sms - 1) % playing + playing) % playing;
                const long long pos = ((2ll * j + 1) * total_w) / (2ll * playing);
                long long acc = 0;
                int k = 0;
                while (k + 1 < s->x.n_queues && pos >= acc + std::max<long long>(qweight[k], 1)) {
                    acc += std::max<long long>(qweight[k], 1);
                    ++k;
                }
                s->h_home[r] = (uint8_t)k;
            }
            static const bool no_affinity = getenv("BP_RESIDENT_AFFINITY") && atoi(getenv("BP_RESIDENT_AFFINITY")) == 0;
            if (no_affinity) std::fill(s->h_home.begin(), s->h_home.end(), 0);
        }
        s->x.poll_ns = getenv("BP_RESIDENT_POLL_NS") ? std::max(32, atoi(getenv("BP_RESIDENT_POLL_NS"))) : 256;
        s->x.poll_tier = getenv("BP_RESIDENT_POLL_TIER") ? std::max(0, atoi(getenv("BP_RESIDENT_POLL_TIER"))) : 0;
        double upw = 4.0;
        if (const char* env = getenv("BP_RESIDENT_UPW")) upw = atof(env);
        s->x.units_per_warp_x4 = std::max(1, (int)(4.0 * upw + 0.5));
        size_t off = s->list_base;
        for (int c : s->classes) {
            s->chain_unit_off[c] = off;
            for (int i : by_class[c]) off += (size_t)n_entries[i] * s->x.gpc_max;
        }
    }

    int rc = BP_OK;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
    const size_t P = n_problems;
    // Arena order: [static inputs, uploaded once] [result block: what bp_search_results reads,
    // zeroed by reset] [more state zeroed by reset] [state set to -1 by reset] [the rest].
    // One reset = two memsets and the copies of the initial states; one result read = the
    // result block in a single copy.
    auto mark = [&]() { return s->arena ? s->arena + s->arena_used : nullptr; };
    auto carve = [&]() -> int {
        s->arena_used = 0;
        s->up_begin = mark();
        TRY(dev_alloc(s, &s->d_desc, P));
        p.desc = s->d_desc;
        TRY(dev_alloc(s, &s->d_cls_problems, cls_problems.size()));
        TRY(dev_alloc(s, &s->d_init_tiles, P * p.es * 2));
        if (s->ring_cap) TRY(dev_alloc(s, &s->d_home, 256));
        s->up_end = mark();
        if (boards) TRY(dev_alloc(s, &s->d_init_occ, P * p.rs * p.ws));
        // queue header: its first 256 bytes (head, tail) survive a reset, the rest is zeroed with
        // the state below and comes back to the host with the result block
        TRY(dev_alloc(s, &s->x.q, 1));
        s->zero_begin = s->arena ? reinterpret_cast<char*>(s->x.q) + RS_QUEUE_RESET_OFFSET : nullptr;
        TRY(dev_alloc(s, &p.status, P));
        TRY(dev_alloc(s, &p.depth, P));
        TRY(dev_alloc(s, &p.n_order, P));
        TRY(dev_alloc(s, &p.n_tiles_placed, P));
        TRY(dev_alloc(s, &p.rollouts, P));
        TRY(dev_alloc(s, &p.rollout_placements, P));
        TRY(dev_alloc(s, &p.rollouts_executed, P));
        TRY(dev_alloc(s, &p.placements_executed, P));
        TRY(dev_alloc(s, &p.order, P * (p.es / 2 + 2) * 2));
        s->res_end = mark();
        TRY(dev_alloc(s, &p.prev_tile, P));
        s->counters_bytes = (size_t)(MAX_DEPTHS + 1) * FS_CHAINS * sizeof(unsigned);
        TRY(dev_alloc(s, &p.child_count, (size_t)(MAX_DEPTHS + 1) * FS_CHAINS));
        TRY(dev_alloc(s, &p.task_counter, (size_t)(MAX_DEPTHS + 1) * FS_CHAINS));
        if (p.use_lanes) TRY(dev_alloc(s, &s->d_sky_count, (size_t)(MAX_DEPTHS + 1) * FS_CHAINS));
        if (p.use_lanes) TRY(dev_alloc(s, &p.pending, P));
        if (!boards) TRY(dev_alloc(s, &p.occ, P * p.rs * p.ws));       // empty boards: zeroed too
        s->zero_end = mark();
        if (boards) TRY(dev_alloc(s, &p.occ, P * p.rs * p.ws));
        s->ff_begin = mark();
        TRY(dev_alloc(s, &p.path, P * (p.es / 2)));
        TRY(dev_alloc(s, &p.child_scores, P * (p.es / 2) * p.es));
        s->ff_end = mark();
        TRY(dev_alloc(s, &p.tiles, P * p.es * 2));
        TRY(dev_alloc(s, &p.lfb, P));
        TRY(dev_alloc(s, &p.legal, P));
        TRY(dev_alloc(s, &p.stats, P * p.es));
        TRY(dev_alloc(s, &p.child_list, (size_t)std::max(list_off, 1)));
        if (p.use_lanes) {
            TRY(dev_alloc(s, &p.planes, P * LANES_PLANE_STRIDE));
            TRY(dev_alloc(s, &p.pairs, P * p.es * LANES_PAIR_WORDS));
            TRY(dev_alloc(s, &p.tile, P * p.es));
            TRY(dev_alloc(s, &p.wmask, P * LANES_TABLE_ROWS * LANES_MASK_WORDS));
            TRY(dev_alloc(s, &p.hmask, P * LANES_TABLE_ROWS * LANES_MASK_WORDS));
            TRY(dev_alloc(s, &p.alive, P));
            s->x.urec_stride = std::min(p.es * s->x.gpc_max, (int)RS_MAX_UNITS);
            TRY(dev_alloc(s, &s->x.urec, P * (size_t)s->x.urec_stride));
            TRY(dev_alloc(s, &s->x.cacc, P * (size_t)p.es));
            TRY(dev_alloc(s, &s->x.ulist, P * p.es));
            TRY(dev_alloc(s, &p.pstate, P));
            TRY(dev_alloc(s, &s->x.slots, s->ring_cap));
            TRY(dev_alloc(s, &s->x.done_slots, (size_t)s->x.done_mask + 1));
        }
        TRY(dev_alloc(s, &p.rec, P * p.es * (size_t)p.n_chunks));
        return BP_OK;
    };
    rc = carve();                                   // dry run: size
    if (rc) { bp_search_destroy(s); return rc; }
    const size_t need = s->arena_used;
    if (use_ctx_arena && !ctx->arena_busy) {
        if (ctx->arena_bytes < need) {
            if (ctx->arena) {
                cudaStreamSynchronize(ctx->stream);
                cudaFree(ctx->arena);
                ctx->arena = nullptr;
                ctx->arena_bytes = 0;
            }
            const size_t want = need + need / 4;
            if (cudaMalloc(&ctx->arena, want) != cudaSuccess) {
                set_error("cudaMalloc of %zu bytes failed", want);
                bp_search_destroy(s);
                return BP_ERR_CUDA;
            }
            ctx->arena_bytes = want;
        }
        s->arena = static_cast<char*>(ctx->arena);
        s->arena_from_ctx = true;
        ctx->arena_busy = true;
    } else {
        void* ptr = nullptr;
        if (cudaMalloc(&ptr, need) != cudaSuccess) {
            set_error("cudaMalloc of %zu bytes failed", need);
            bp_search_destroy(s);
            return BP_ERR_CUDA;
        }
        s->arena = static_cast<char*>(ptr);
    }
    s->arena_bytes = need;
    rc = carve();                                   // hand out the slices
    if (rc) { bp_search_destroy(s); return rc; }
#undef TRY
    // static inputs: one host image, one copy
    s->h_upload.assign((size_t)(s->up_end - s->up_begin), 0);
    auto put = [&](const void* dev, const void* src, size_t bytes) {
        memcpy(s->h_upload.data() + (reinterpret_cast<const char*>(dev) - s->up_begin), src, bytes);
    };
    put(s->d_desc, s->h_desc.data(), P * sizeof(ProblemDesc));
    put(s->d_cls_problems, cls_problems.data(), cls_problems.size() * sizeof(int));
    put(s->d_init_tiles, tiles, P * p.es * 2);
    if (s->ring_cap) {
        put(s->d_home, s->h_home.data(), 256);
        s->x.home_tab = s->d_home;
    }
    BP_CUDA(cudaMemcpyAsync(s->up_begin, s->h_upload.data(), s->h_upload.size(), cudaMemcpyHostToDevice,
                            ctx->stream));
    if (boards)
        BP_CUDA(cudaMemcpyAsync(s->d_init_occ, boards, P * p.rs * p.ws * 4, cudaMemcpyHostToDevice,
                                ctx->stream));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));

    // persistent grid per class: every SM filled to the occupancy limit of that kernel
    for (int chain : s->classes) {
        const int c = chain % FS_CLASSES;
        if (!ctx->occ_rollout[c]) {
            int per_sm = 1;
            dispatch_class(c % N_CLASSES, [&](auto R, auto Wc, auto E) {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                    &per_sm, fs_rollout_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>,
                    128, 0);
            });
            ctx->occ_rollout[c] = std::max(per_sm, 1);
        }
        s->grid_rollout[chain] = ctx->sm_count * ctx->occ_rollout[c];
        if (p.use_lanes) {
            if (!ctx->occ_lanes[c]) {
                int per_sm_l = 1;
                dispatch_lanes(c, [&](auto NBc, auto Wc, auto E) {
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                        &per_sm_l,
                        fs_rollout_lanes_kernel<decltype(NBc)::value, decltype(Wc)::value, decltype(E)::value>,
                        128, 0);
                });
                ctx->occ_lanes[c] = std::max(per_sm_l, 1);
            }
            s->grid_lanes[chain] = ctx->sm_count * ctx->occ_lanes[c];
        }
    }
    if (s->ring_cap) {
        if (!ctx->occ_resident[s->variant]) {
            int per_sm = 1;
            dispatch_resident(s->variant, [&](auto MW, auto ME) {
                per_sm = ResidentVariant<decltype(MW)::value, decltype(ME)::value>::blocks_per_sm();
            });
            ctx->occ_resident[s->variant] = std::max(per_sm, 1);
        }
        s->grid_resident = ctx->sm_count * ctx->occ_resident[s->variant];
        if (const char* env = getenv("BP_RESIDENT_BLOCKS_PER_SM"))
            s->grid_resident = ctx->sm_count * std::max(1, std::min(atoi(env), ctx->occ_resident[s->variant]));
    }
    // side streams and fork / join events come from the context's pool
    if (!ctx->fork_event) BP_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    s->fork_ev = ctx->fork_event;
    // heaviest chain first: stream k has priority rank k (common.cuh), so the chain whose launch
    // sequence is the critical path gets the SM slots first and the lighter chains fill the gaps
    {
        auto weight = [&](int c) {
            long long wgt = 0;
            for (int i : by_class[c]) wgt += (long long)s->h_desc[i].rows * s->h_desc[i].cols;
            return wgt;
        };
        std::stable_sort(s->classes.begin(), s->classes.end(), [&](int a, int b) { return weight(a) > weight(b); });
    }
    for (size_t k = 0; k < s->classes.size(); ++k) {
        cudaStream_t st = side_stream(ctx, k);
        cudaEvent_t e = side_event(ctx, k);
        if (!st || !e) { set_error("cannot create a side stream / event"); bp_search_destroy(s); return BP_ERR_CUDA; }
        s->cls_stream.push_back(st);
        s->cls_done.push_back(e);
        const int c = s->classes[k];
        int md = 0;
        for (int i : by_class[c]) md = std::max(md, s->h_desc[i].n_entries / 2);
        s->cls_depth.push_back(md);
    }
    *out = s;
    return BP_OK;
}

extern "C" {

int bp_search_create(bp_context* ctx, int n_problems, const int32_t* rows, const int32_t* cols,
                     const int32_t* n_entries, const uint8_t* tiles, int entry_stride,
                     const uint32_t* boards, const uint32_t* streams, int n_sim, int strategy,
                     uint32_t seed, bp_search** out)
{
    return search_create(ctx, n_problems, rows, cols, n_entries, tiles, entry_stride, boards,
                         streams, n_sim, strategy, seed, false, out);
}

int bp_search_set_partition(bp_search* s, int rank, int n_ranks, int sim_begin, int n_local)
{
    BP_REQUIRE(s != nullptr, "search is null");
    BP_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "rank / n_ranks");
    BP_REQUIRE(!s->d_exchange, "the partition is fixed once the exchange buffer exists");
    BP_REQUIRE(sim_begin >= 0 && n_local >= 0 && sim_begin + n_local <= s->p.n_sim, "sim range");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    BP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    s->p.rank = rank;
    s->p.n_ranks = n_ranks;
    s->p.sim_begin = sim_begin;
    s->p.n_local = n_local;
    s->p.n_chunks = std::max(1, (n_local + CHUNK - 1) / CHUNK);
    if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }   // parameters are baked in
    s->ready = false;
    return BP_OK;    // the record buffer was sized for all n_sim rollouts
}

}  // extern "C"

extern "C" {

// Exchange buffer of a partitioned (root-parallel) resident search: the per-child statistics of
// every rank for the depth in flight, four generations (run parity x depth parity), and one
// arrival stamp per (problem, rank).  Peers store into it over NVLink from inside their kernels.
int bp_search_exchange_create(bp_search* s, void* ipc_handle_64, void** local_ptr)
{
    BP_REQUIRE(s != nullptr, "search is null");
    BP_REQUIRE(s->p.n_ranks > 1 && s->p.n_ranks <= RS_MAX_RANKS, "needs a partition over 2..8 ranks");
    BP_REQUIRE(s->ring_cap != 0, "the resident kernel is not available for this search");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    if (!s->d_exchange) {
        const size_t stats = (size_t)4 * s->p.n_ranks * s->p.n_problems * s->p.es * sizeof(ChildStats);
        s->exchange_flags_off = (stats + 255) & ~size_t(255);
        s->exchange_table_off =
            (s->exchange_flags_off + (size_t)s->p.n_problems * RS_MAX_RANKS * sizeof(int) + 255) & ~size_t(255);
        s->exchange_bytes = s->exchange_table_off + sizeof(ResidentParams::PeerTable);
        BP_CUDA(cudaMalloc(&s->d_exchange, s->exchange_bytes));
        BP_CUDA(cudaMemset(s->d_exchange, 0, s->exchange_bytes));
    }
    if (ipc_handle_64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        BP_CUDA(cudaIpcGetMemHandle(&h, s->d_exchange));
        memcpy(ipc_handle_64, &h, sizeof(h));
    }
    if (local_ptr) *local_ptr = s->d_exchange;
    return BP_OK;
}

// ipc_handles: uint8[n_ranks][64] gathered from all ranks (other processes, one per GPU), or
// local_ptrs: the exchange buffers of n_ranks searches of THIS process (several ranks emulated on
// one device, or several devices of one process with peer access enabled).
int bp_search_exchange_connect(bp_search* s, const void* ipc_handles, void* const* local_ptrs)
{
    BP_REQUIRE(s != nullptr && s->d_exchange != nullptr, "call bp_search_exchange_create first");
    BP_REQUIRE((ipc_handles != nullptr) != (local_ptrs != nullptr), "give IPC handles or local pointers");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    ResidentParams::PeerTable table{};
    for (int r = 0; r < s->p.n_ranks; ++r) {
        char* base = nullptr;
        if (r == s->p.rank) {
            base = static_cast<char*>(s->d_exchange);
        } else if (local_ptrs) {
            base = static_cast<char*>(local_ptrs[r]);
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, static_cast<const char*>(ipc_handles) + (size_t)r * 64, sizeof(h));
            void* mapped = nullptr;
            BP_CUDA(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
            s->peer_opened.push_back(mapped);
            base = static_cast<char*>(mapped);
        }
        BP_REQUIRE(base != nullptr, "null peer buffer");
        table.gathered[r] = reinterpret_cast<ChildStats*>(base);
        table.arrived[r] = reinterpret_cast<int*>(base + s->exchange_flags_off);
    }
    char* own = static_cast<char*>(s->d_exchange);
    BP_CUDA(cudaMemcpy(own + s->exchange_table_off, &table, sizeof(table), cudaMemcpyHostToDevice));
    s->x.peers = reinterpret_cast<const ResidentParams::PeerTable*>(own + s->exchange_table_off);
    s->exchange_connected = true;
    return BP_OK;
}

// blocks per SM of the resident kernel (0 = as many as fit); several searches that must be
// resident on one device at the same time (ranks emulated on one GPU) share the SMs this way
int bp_search_set_resident_blocks(bp_search* s, int blocks_per_sm)
{
    BP_REQUIRE(s != nullptr && s->ring_cap != 0, "the resident kernel is not available for this search");
    const int fit = s->ctx->occ_resident[s->variant];
    s->grid_resident = s->ctx->sm_count * (blocks_per_sm <= 0 ? fit : std::min(blocks_per_sm, fit));
    return BP_OK;
}

}  // extern "C"

// Kernel launch, optionally as a programmatic dependent launch: the kernel may start while the
// kernel before it in the stream is still running and waits for it in grid_dependency_wait().
// A chain's launches are rollouts(d) -> advance(d) -> rollouts(d + 1) ...; on the device timeline
// (tools/timeline.py) the 4-8 us between the end of one and the first warp of the next were a
// quarter of a late depth's ~45 us.
template <typename... KArgs, typename... Args>
static void launch_kernel(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, bool pdl, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <bool FIRST>
static void launch_advance_class(bp_search* s, int c, int next_slot, const ChildStats* gathered,
                                 cudaStream_t st, bool pdl = false)
{
    const int n = s->cls_count[c];
    const int grid = (n + ADV_WARPS - 1) / ADV_WARPS;
    dispatch_class((c % FS_CLASSES) % N_CLASSES, [&](auto R, auto Wc, auto E) {
        launch_kernel(fs_advance_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value, FIRST>,
                      grid, ADV_WARPS * 32, st, pdl, s->p, c, s->cls_list_off[c],
                      (const int*)(s->d_cls_problems + s->cls_prob_off[c]), n, next_slot, gathered);
    });
    s->ctx->launches++;
}

template <bool FIRST>
static void launch_advance(bp_search* s, int next_slot, const ChildStats* gathered)
{
    // (the first advances of a reset overlap: every class after the first is a dependent launch)
    static const bool pdl = !(getenv("BP_PDL") && atoi(getenv("BP_PDL")) == 0);
    bool first = true;
    for (int c : s->classes) {
        launch_advance_class<FIRST>(s, c, next_slot, gathered, s->ctx->stream, FIRST && pdl && !first);
        first = false;
    }
}

// `overlapped`: other classes' launches run beside this one on their own streams.  Measured on
// problems_g (4 classes, 1000 problems, N = 1000; profiles/r1): with >= 8 units per resident warp
// a whole search takes 13.12 ms, with 4 / 2 / 1 it takes 12.61 / 12.27 / 12.00 ms -- larger units
// amortise the per-unit table fill and queue atomic, and the longer tail of a launch is filled
// by the other classes' kernels.  A class running alone has nobody to fill its tail: there the
// sum of the rollout kernels is 13.52 / 13.53 / 14.04 / 14.88 ms, so it keeps small units.
// Half a unit per warp (whole children almost always) gives another 2 % (4 % at N = 100); smaller
// values change nothing, the unit is capped at one child.
static void launch_rollouts_class(bp_search* s, int c, int depth, cudaStream_t st, bool overlapped,
                                  bool pdl = false)
{
    int upw = overlapped ? 1 : 16;                            // in half units: 0.5 / 8 units per warp
    if (const char* env = getenv("BP_UNITS_PER_WARP")) upw = std::max(1, (int)(2.0 * atof(env) + 0.5));
    if (s->p.use_lanes) {
        // no more blocks than the units this depth can have: every problem of the chain has at most
        // (entries - 2 depth) children, and a launch is cut into about n_warps * upw / 2 units when
        // the children are fewer than that (see the kernel) -- a full grid of blocks that find the
        // counter exhausted costs the launch several microseconds on every SM
        const int full = s->grid_lanes[c];
        const long long children = (long long)s->cls_count[c] * std::max(0, s->p.es - 2 * depth);
        const long long units = std::min<long long>(children * s->p.n_chunks,
                                                    children + ((long long)full * 4 * upw) / 2);
        static const bool shrink = !(getenv("BP_SHRINK_GRID") && atoi(getenv("BP_SHRINK_GRID")) == 0);
        const int grid = shrink ? (int)std::max<long long>(s->ctx->sm_count, std::min<long long>(full, (units + 3) / 4))
                                : full;
        dispatch_lanes(c % FS_CLASSES, [&](auto NBc, auto Wc, auto E) {
            launch_kernel(fs_rollout_lanes_kernel<decltype(NBc)::value, decltype(Wc)::value, decltype(E)::value>,
                          grid, 128, st, pdl, s->p, c, s->cls_list_off[c], depth, upw, full * 4);
        });
    }
    else
        dispatch_class((c % FS_CLASSES) % N_CLASSES, [&](auto R, auto Wc, auto E) {
            fs_rollout_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
                <<<s->grid_rollout[c], 128, 0, st>>>(s->p, c, s->cls_list_off[c], depth);
        });
    s->ctx->launches++;
}

// ---- launch-per-depth form on the skyline state (search_resident.cuh) ------------------------
// The resident kernel's playout and advance with launches as barriers.  The advance is the compact
// skyline advance (no table is rebuilt) and, since the units fold their statistics into one
// accumulator per child, it reads one record per child: with dependent launches this pair of
// kernels is faster than the warp-per-board advance with per-depth tables at every size measured
// (profiles/r2/sky_vs_legacy_final.txt: 125 problems 1.79 / 1.92 ms, 500 problems 5.61 / 5.86 ms, the full
// set 10.89 / 11.17 ms at N = 1000).  BP_STEPPED=sky|legacy forces one of them.
static bool sky_stepped_ok(const bp_search* s)
{
    if (!(s->ring_cap != 0 && s->p.use_lanes && s->p.n_ranks == 1)) return false;
    if (const char* env = getenv("BP_STEPPED")) {
        if (strcmp(env, "legacy") == 0) return false;
        if (strcmp(env, "sky") == 0) return true;
    }
    return true;
}
static int sky_upw_x4()
{
    static const int v = getenv("BP_SKY_UPW") ? std::max(1, (int)(4.0 * atof(getenv("BP_SKY_UPW")) + 0.5)) : 4;
    return v;
}
static UnitSink sky_sink(bp_search* s, int c, int depth_slot)
{
    UnitSink sink;
    sink.tail = s->d_sky_count + (size_t)depth_slot * FS_CHAINS + c;
    sink.slots = s->x.slots + s->chain_unit_off[c];
    sink.mask = 0xFFFFFFFFu;
    sink.in_flight = s->cls_count[c];
    sink.n_warps = (unsigned)s->grid_lanes[c] * 4u;
    // (the child counters of the warp-per-board form double as the chains' active-problem counts)
    sink.active_now = depth_slot > 0 ? s->p.child_count + (size_t)(depth_slot - 1) * FS_CHAINS + c : nullptr;
    sink.active_next = s->p.child_count + (size_t)depth_slot * FS_CHAINS + c;
    return sink;
}
static void launch_seed_sky(bp_search* s, int c, cudaStream_t st)
{
    const int n = s->cls_count[c];
    ResidentParams x = s->x;
    x.units_per_warp_x4 = sky_upw_x4();                      // half a unit per warp: whole children when there is plenty
    fs_resident_seed_kernel<0><<<(n + ADV_WARPS - 1) / ADV_WARPS, ADV_WARPS * 32, 0, st>>>(
        s->p, x, sky_sink(s, c, 0), s->d_cls_problems + s->cls_prob_off[c], n);
    s->ctx->launches++;
}
static void launch_rollouts_sky(bp_search* s, int c, int depth, cudaStream_t st, bool pdl = false)
{
    const int full = s->grid_lanes[c];
    const long long children = (long long)s->cls_count[c] * std::max(0, s->p.es - 2 * depth);
    const long long units = std::min<long long>(children * s->p.n_chunks, children + (long long)full * 2);
    const int grid = (int)std::max<long long>(s->ctx->sm_count, std::min<long long>(full, (units + 3) / 4));
    const unsigned long long* list = s->x.slots + s->chain_unit_off[c];
    const unsigned long long* count = s->d_sky_count + (size_t)depth * FS_CHAINS + c;
    unsigned* counter = s->p.task_counter + (size_t)depth * FS_CHAINS + c;
    dispatch_lanes(c % FS_CLASSES, [&](auto NBc, auto Wc, auto E) {
        constexpr int NB = decltype(NBc)::value == 7 ? 8 : decltype(NBc)::value;   // the skyline playouts: 5 / 6 / 8 planes
        launch_kernel(fs_rollout_sky_kernel<NB, decltype(Wc)::value, decltype(E)::value>, grid, 128, st, pdl,
                      s->p, s->x, list, count, counter);
    });
    s->ctx->launches++;
}
static void launch_advance_sky(bp_search* s, int c, int depth, cudaStream_t st, bool handover = false,
                               bool pdl = false)
{
    const int n = s->cls_count[c];
    const int grid = (n + ADV_WARPS - 1) / ADV_WARPS;
    const int* problems = s->d_cls_problems + s->cls_prob_off[c];
    if (handover) {
        // the chain's last launch-per-depth advance: units of the next depth go to the resident
        // kernel's queues, sized for its warps
        UnitSink sink = sky_sink(s, c, depth + 1);
        sink.n_warps = (unsigned)s->x.n_roll_warps;
        launch_kernel(fs_advance_sky_kernel<1>, grid, ADV_WARPS * 32, st, pdl, s->p, s->x, sink, problems, n,
                      s->p.n_problems);
    } else {
        ResidentParams x = s->x;
        x.units_per_warp_x4 = sky_upw_x4();
        launch_kernel(fs_advance_sky_kernel<0>, grid, ADV_WARPS * 32, st, pdl, s->p, x, sky_sink(s, c, depth + 1),
                      problems, n, s->p.n_problems);
    }
    s->ctx->launches++;
}

// Hybrid form: the depths with plenty of work are played launch-per-depth (per-class kernels, no
// SM given to server warps), the rest by the resident kernel.  The device timeline of the
// launch-per-depth form (tools/timeline.py, profiles/r2/timeline_*.txt) shows why: once a depth is a
// few tens of microseconds of work, each one costs ~45 us -- rollout ramp and tail, two launch
// gaps of 4-8 us, an advance launch of ~15 us -- and every chain pays it for every remaining depth
// (the last 13 depths of a 125-problem shard are 600 us of its 2.1 ms with the GPU a third full).
// In the resident kernel a problem's next depth starts as soon as its own units are done.
// Work of depth d ~ 2 P N (D - d)^2 placements (P problems, N rollouts per child, D tiles); the
// hand-over is at the first depth whose work is below BP_HYBRID_WORK (default 5e7).
static int hybrid_handover_depth(const bp_search* s)
{
    if (const char* env = getenv("BP_HYBRID_DEPTH")) return std::max(0, std::min(atoi(env), s->max_depth));
    double work = 5e7;
    if (const char* env = getenv("BP_HYBRID_WORK")) work = atof(env);
    const double per = 2.0 * (double)s->p.n_problems * (double)s->p.n_sim;
    const int rest = (int)std::floor(std::sqrt(work / per));
    return std::max(0, std::min(s->max_depth - rest, s->max_depth));
}

extern "C" {

int bp_search_reset(bp_search* s)
{
    BP_REQUIRE(s != nullptr, "search is null");
    bp_context* ctx = s->ctx;
    SearchParams& p = s->p;
    BP_CUDA(cudaSetDevice(ctx->device));
    const size_t P = p.n_problems;
    cudaStream_t st = ctx->stream;
    // the initial states are resident in HBM (uploaded by bp_search_create)
    BP_CUDA(cudaMemsetAsync(s->zero_begin, 0, (size_t)(s->zero_end - s->zero_begin), st));
    BP_CUDA(cudaMemsetAsync(s->ff_begin, 0xFF, (size_t)(s->ff_end - s->ff_begin), st));
    if (s->d_init_occ)
        BP_CUDA(cudaMemcpyAsync(p.occ, s->d_init_occ, P * p.rs * p.ws * 4, cudaMemcpyDeviceToDevice, st));
    BP_CUDA(cudaMemcpyAsync(p.tiles, s->d_init_tiles, P * p.es * 2, cudaMemcpyDeviceToDevice, st));
    if (s->ring_cap && s->queue_dirty) {
        // first use (or after an aborted run): empty ring, tickets from zero
        BP_CUDA(cudaMemsetAsync(s->x.q, 0, RS_QUEUE_RESET_OFFSET, st));
        BP_CUDA(cudaMemsetAsync(s->x.slots, 0, s->ring_cap * sizeof(unsigned long long), st));
        BP_CUDA(cudaMemsetAsync(s->x.done_slots, 0, ((size_t)s->x.done_mask + 1) * sizeof(unsigned long long), st));
        s->queue_dirty = false;
    }
    s->seeded = false;
    s->x.run_epoch += 1;
    launch_advance<true>(s, 0, nullptr);
    BP_CUDA(cudaGetLastError());
    s->ran_resident = false;
    s->sky_stepped = false;
    s->cur_depth = 0;
    s->ev_depths = 0;
    s->ready = true;
    return BP_OK;
}

int bp_search_set_profiling(bp_search* s, int on)
{
    BP_REQUIRE(s != nullptr, "search is null");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    s->profiling = on != 0;
    if (s->profiling && s->ev.empty()) {
        s->ev.resize((size_t)3 * (s->max_depth + 1));
        for (cudaEvent_t& e : s->ev) BP_CUDA(cudaEventCreate(&e));
    }
    return BP_OK;
}

int bp_search_set_mode(bp_search* s, int mode)
{
    BP_REQUIRE(s != nullptr, "search is null");
    BP_REQUIRE(mode == BP_SEARCH_AUTO || mode == BP_SEARCH_STEPPED || mode == BP_SEARCH_RESIDENT ||
               mode == BP_SEARCH_HYBRID, "mode");
    s->mode = mode;
    return BP_OK;
}

int bp_search_profile(bp_search* s, double* rollout_ms, double* advance_ms, int* n_depths)
{
    BP_REQUIRE(s && rollout_ms && advance_ms && n_depths, "null argument");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    BP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    double r = 0.0, a = 0.0;
    for (int d = 0; d < s->ev_depths; ++d) {
        float ms = 0.f;
        BP_CUDA(cudaEventElapsedTime(&ms, s->ev[3 * d], s->ev[3 * d + 1]));
        r += ms;
        BP_CUDA(cudaEventElapsedTime(&ms, s->ev[3 * d + 1], s->ev[3 * d + 2]));
        a += ms;
    }
    *rollout_ms = r;
    *advance_ms = s->ran_resident ? 0.0 : a;
    *n_depths = s->ev_depths;
    return BP_OK;
}

int bp_search_step_rollouts(bp_search* s)
{
    BP_REQUIRE(s != nullptr, "search is null");
    if (!s->ready || s->cur_depth >= s->max_depth) {
        set_error("bp_search_step_rollouts: call bp_search_reset first / search finished");
        return BP_ERR_STATE;
    }
    BP_CUDA(cudaSetDevice(s->ctx->device));
    if (s->profiling) BP_CUDA(cudaEventRecord(s->ev[3 * s->cur_depth], s->ctx->stream));
    // (a profiled bp_search_run serialises the classes but keeps the unit size of the real run,
    // so the per-kernel times and instruction counts describe the kernel that is timed)
    for (int c : s->classes)
        launch_rollouts_class(s, c, s->cur_depth, s->ctx->stream, s->in_profiled_run && s->classes.size() > 1);
    BP_CUDA(cudaGetLastError());
    if (s->profiling) BP_CUDA(cudaEventRecord(s->ev[3 * s->cur_depth + 1], s->ctx->stream));
    return BP_OK;
}

int bp_search_step_reduce(bp_search* s)
{
    BP_REQUIRE(s != nullptr, "search is null");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    for (int c : s->classes) {
        const int n = s->cls_count[c];
        const int grid = (n + ADV_WARPS - 1) / ADV_WARPS;
        const int ej = ((c % FS_CLASSES) % N_CLASSES) % 3 == 0 ? 1 : (((c % FS_CLASSES) % N_CLASSES) % 3 == 1 ? 2 : 4);
        const int* list = s->d_cls_problems + s->cls_prob_off[c];
        if (ej == 1) fs_reduce_kernel<1><<<grid, ADV_WARPS * 32, 0, s->ctx->stream>>>(s->p, list, n);
        else if (ej == 2) fs_reduce_kernel<2><<<grid, ADV_WARPS * 32, 0, s->ctx->stream>>>(s->p, list, n);
        else fs_reduce_kernel<4><<<grid, ADV_WARPS * 32, 0, s->ctx->stream>>>(s->p, list, n);
        s->ctx->launches++;
    }
    BP_CUDA(cudaGetLastError());
    return BP_OK;
}

int bp_search_step_commit(bp_search* s, const void* gathered_stats_dev)
{
    BP_REQUIRE(s != nullptr, "search is null");
    if (!s->ready || s->cur_depth >= s->max_depth) {
        set_error("bp_search_step_commit: no depth in flight");
        return BP_ERR_STATE;
    }
    BP_REQUIRE(s->p.n_ranks == 1 || gathered_stats_dev != nullptr,
               "a partitioned search needs the gathered statistics of all ranks");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    launch_advance<false>(s, s->cur_depth + 1, static_cast<const ChildStats*>(gathered_stats_dev));
    BP_CUDA(cudaGetLastError());
    if (s->profiling) {
        BP_CUDA(cudaEventRecord(s->ev[3 * s->cur_depth + 2], s->ctx->stream));
        s->ev_depths = s->cur_depth + 1;
    }
    s->cur_depth += 1;
    return BP_OK;
}

int bp_search_stats_buffer(bp_search* s, void** stats_dev, int64_t* n_bytes)
{
    BP_REQUIRE(s && stats_dev && n_bytes, "null argument");
    *stats_dev = s->p.stats;
    *n_bytes = (int64_t)s->p.n_problems * s->p.es * (int64_t)sizeof(ChildStats);
    return BP_OK;
}

int bp_search_info(bp_search* s, int* lane_kernel, int* n_classes)
{
    BP_REQUIRE(s != nullptr, "search is null");
    if (lane_kernel) *lane_kernel = s->p.use_lanes;
    if (n_classes) *n_classes = (int)s->classes.size();
    return BP_OK;
}

int bp_search_last_run(bp_search* s, int* resident)
{
    BP_REQUIRE(s && resident, "null argument");
    *resident = s->ran_resident ? (s->hybrid_depths > 0 ? 2 : 1) : (s->sky_stepped ? 3 : 0);
    return BP_OK;
}

int bp_search_resident_stats(bp_search* s, double* out13)
{
    double* out8 = out13;
    BP_REQUIRE(s && out8, "null argument");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    BP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    for (int i = 0; i < 13; ++i) out8[i] = 0.0;
    if (!s->ran_resident) return BP_OK;
    UnitQueue q;
    BP_CUDA(cudaMemcpy(&q, s->x.q, sizeof(UnitQueue), cudaMemcpyDeviceToHost));
    out8[0] = (double)q.cyc_total;
    out8[1] = (double)q.cyc_wait;
    out8[2] = (double)q.cyc_play;
    out8[3] = (double)q.cyc_advance;
    out8[4] = (double)q.n_units_played;
    out8[5] = (double)q.n_advances;
    out8[6] = (double)q.n_polls;
    out8[7] = (double)q.cyc_exchange;
    out8[8] = (double)q.n_units_published;
    out8[9] = (double)s->grid_resident * 4.0;
    out8[10] = (double)q.cyc_adv_merge;
    out8[11] = (double)q.cyc_adv_commit;
    out8[12] = (double)q.cyc_adv_publish;
    return BP_OK;
}

int bp_search_max_depth(bp_search* s, int* max_depth)
{
    BP_REQUIRE(s && max_depth, "null argument");
    *max_depth = s->max_depth;
    return BP_OK;
}

// AUTO: the form measured faster on B200 (profiles/r2/auto_threshold.txt, c4_modes.txt, DESIGN.md
// section 4).  With programmatic dependent launches, the grid bound, the CUDA graph and the skyline
// kernels the launch-per-depth form wins every batch of 20-tile problems from 1 to 1000 at N = 100
// and 1000 (0.18 / 0.28 ms for one instance, 1.79 / 2.1 ms for 125, 10.9 / 12.4 ms for 1000): its
// per-class kernels run ~6 % fewer instructions and give no SM to server warps.  The resident kernel
// wins where a search is a long chain of moves on one or a few instances with very many rollouts
// per child (50 tiles: 3.43 / 4.74 ms at N = 5000, but 1.70 / 1.64 ms at N = 1000 and 1.48 / 0.96 ms at
// N = 100), and it is the only form that runs a partitioned (root-parallel) search without a host
// step per move.
static int default_mode(const bp_search* s)
{
    if (!s->ring_cap) return BP_SEARCH_STEPPED;
    if (s->p.n_ranks > 1) return BP_SEARCH_RESIDENT;
    return (s->p.n_problems <= 4 && s->max_depth > 32 && s->p.n_chunks >= 64) ? BP_SEARCH_RESIDENT
                                                                               : BP_SEARCH_STEPPED;
}

int bp_search_run(bp_search* s)
{
    BP_REQUIRE(s != nullptr, "search is null");
    int mode = s->mode;
    if (const char* env = getenv("BP_SEARCH")) {
        if (strcmp(env, "resident") == 0) mode = BP_SEARCH_RESIDENT;
        else if (strcmp(env, "stepped") == 0) mode = BP_SEARCH_STEPPED;
        else if (strcmp(env, "hybrid") == 0) mode = BP_SEARCH_HYBRID;
    }
    if (mode == BP_SEARCH_AUTO) mode = default_mode(s);
    int handover = -1;                           // hybrid: depths played launch-per-depth
    if (mode == BP_SEARCH_HYBRID) {
        if (!s->ring_cap || s->p.n_ranks != 1) {
            set_error("the hybrid search needs the lane kernel (empty initial boards, equal tiles adjacent "
                      "in every list) and an unpartitioned search");
            return BP_ERR_STATE;
        }
        handover = hybrid_handover_depth(s);
        if (handover <= 0) mode = BP_SEARCH_RESIDENT;
        else if (handover >= s->max_depth - 1) mode = BP_SEARCH_STEPPED;
    }
    s->hybrid_depths = mode == BP_SEARCH_HYBRID ? handover : 0;
    if (s->p.n_ranks != 1 && mode != BP_SEARCH_RESIDENT) {
        set_error("bp_search_run of a partitioned search needs the resident kernel (or use the "
                  "step API with an all-gather per move)");
        return BP_ERR_STATE;
    }
    if (mode == BP_SEARCH_RESIDENT && !s->ring_cap) {
        set_error("the resident search kernel needs the lane kernel (empty initial boards, equal "
                  "tiles adjacent in every list; BP_ROLLOUT != warp) and fewer than 2^20 problems");
        return BP_ERR_STATE;
    }
    if (mode == BP_SEARCH_RESIDENT) {
        // resident search: seed the queue with depth 0, then one launch plays every depth of
        // every problem (all shape classes in the same launch)
        if (!s->ready || s->cur_depth != 0) {
            set_error("bp_search_run: call bp_search_reset first");
            return BP_ERR_STATE;
        }
        if (s->p.n_ranks > 1 && !s->exchange_connected) {
            set_error("a partitioned resident search needs bp_search_exchange_connect first");
            return BP_ERR_STATE;
        }
        BP_CUDA(cudaSetDevice(s->ctx->device));
        cudaStream_t st = s->ctx->stream;
        const int blocks_per_sm = s->grid_resident / s->ctx->sm_count;
        if (s->x.n_server_sms >= s->ctx->sm_count) {
            set_error("resident search: %d server SMs leave nothing to play on", s->x.n_server_sms);
            return BP_ERR_STATE;
        }
        s->x.n_roll_warps = (s->ctx->sm_count - s->x.n_server_sms) * blocks_per_sm * 4;
        const unsigned n_warps = (unsigned)s->x.n_roll_warps;
        s->hybrid_depths = 0;
        if (s->profiling) BP_CUDA(cudaEventRecord(s->ev[0], st));
        if (!s->seeded) {
            const int grid = (s->p.n_problems + ADV_WARPS - 1) / ADV_WARPS;
            UnitSink sink;
            sink.tail = nullptr;                              // (per problem: the queue of its class)
            sink.slots = nullptr;
            sink.mask = 0;
            sink.in_flight = s->x.seed_in_flight;
            sink.n_warps = n_warps;
            sink.active_now = nullptr;
            sink.active_next = nullptr;
            fs_resident_seed_kernel<0><<<grid, ADV_WARPS * 32, 0, st>>>(s->p, s->x, sink, nullptr, s->p.n_problems);
            s->ctx->launches++;
            s->seeded = true;
        }
        dispatch_resident(s->variant, [&](auto MW, auto ME) {
            ResidentVariant<decltype(MW)::value, decltype(ME)::value>::launch(s->grid_resident, st, s->p, s->x);
        });
        s->ctx->launches++;
        BP_CUDA(cudaGetLastError());
        if (s->profiling) {
            BP_CUDA(cudaEventRecord(s->ev[1], st));
            BP_CUDA(cudaEventRecord(s->ev[2], st));
            s->ev_depths = 1;
        }
        s->ran_resident = true;
        s->cur_depth = s->max_depth;
        return BP_OK;
    }
    const bool hybrid = mode == BP_SEARCH_HYBRID;
    static const bool fork_single = !(getenv("BP_FORK_SINGLE") && atoi(getenv("BP_FORK_SINGLE")) == 0);
    if (!hybrid && (s->profiling || (s->classes.size() == 1 && !fork_single) || getenv("BP_SERIAL_CLASSES"))) {
        s->in_profiled_run = s->profiling;
        s->sky_stepped = false;
        int rc = BP_OK;
        while (!rc && s->cur_depth < s->max_depth) {
            rc = bp_search_step_rollouts(s);
            if (!rc) rc = bp_search_step_commit(s, nullptr);
        }
        s->in_profiled_run = false;
        return rc;
    }
    if (!s->ready || s->cur_depth != 0) {
        set_error("bp_search_run: call bp_search_reset first");
        return BP_ERR_STATE;
    }
    if (hybrid) {
        if (s->x.n_server_sms >= s->ctx->sm_count) {
            set_error("resident search: %d server SMs leave nothing to play on", s->x.n_server_sms);
            return BP_ERR_STATE;
        }
        s->x.n_roll_warps = (s->ctx->sm_count - s->x.n_server_sms) * (s->grid_resident / s->ctx->sm_count) * 4;
    }
    // fork: every chain runs its whole depth loop on its own stream, then join.  Launches are
    // issued depth-major (all chains' depth 0, then depth 1, ...) so that every chain starts
    // at once instead of waiting for the host to issue the chains before it.
    BP_CUDA(cudaSetDevice(s->ctx->device));
    static const bool pdl = !(getenv("BP_PDL") && atoi(getenv("BP_PDL")) == 0);
    auto issue = [&]() -> int {
#ifdef BP_TIMELINE
        if (!s->d_timeline) {
            BP_CUDA(cudaMalloc(&s->d_timeline, (size_t)(MAX_DEPTHS + 1) * FS_CHAINS * 2 * 4 * 8));
            BP_CUDA(cudaMemcpyToSymbol(g_timeline, &s->d_timeline, sizeof(s->d_timeline)));
        }
        BP_CUDA(cudaMemsetAsync(s->d_timeline, 0, (size_t)(MAX_DEPTHS + 1) * FS_CHAINS * 2 * 4 * 8, s->ctx->stream));
#endif
        BP_CUDA(cudaEventRecord(s->fork_ev, s->ctx->stream));
        for (size_t k = 0; k < s->classes.size(); ++k)
            BP_CUDA(cudaStreamWaitEvent(s->cls_stream[k], s->fork_ev, 0));
        const bool sky = hybrid || sky_stepped_ok(s);
        s->sky_stepped = sky;
        if (sky)
            for (size_t k = 0; k < s->classes.size(); ++k) launch_seed_sky(s, s->classes[k], s->cls_stream[k]);
        const int n_stepped = hybrid ? handover : s->max_depth;
        for (int d = 0; d < n_stepped; ++d) {
            for (size_t k = 0; k < s->classes.size(); ++k) {
                if (d >= s->cls_depth[k]) continue;
                if (sky) {
                    launch_rollouts_sky(s, s->classes[k], d, s->cls_stream[k], pdl);
                    launch_advance_sky(s, s->classes[k], d, s->cls_stream[k], hybrid && d == n_stepped - 1, pdl);
                } else {
                    launch_rollouts_class(s, s->classes[k], d, s->cls_stream[k], true, pdl && d > 0);
                    launch_advance_class<false>(s, s->classes[k], d + 1, nullptr, s->cls_stream[k], pdl);
                }
            }
        }
        BP_CUDA(cudaGetLastError());
        for (size_t k = 0; k < s->classes.size(); ++k) {
            BP_CUDA(cudaEventRecord(s->cls_done[k], s->cls_stream[k]));
            BP_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->cls_done[k], 0));
        }
        if (hybrid) {
            // every chain has handed its live problems over: one launch plays what is left
            dispatch_resident(s->variant, [&](auto MW, auto ME) {
                ResidentVariant<decltype(MW)::value, decltype(ME)::value>::launch(s->grid_resident, s->ctx->stream,
                                                                                   s->p, s->x);
            });
            s->ctx->launches++;
            BP_CUDA(cudaGetLastError());
        }
        return BP_OK;
    };
    // From the second run of a search on, the whole fork / launch / join sequence is one CUDA
    // graph: its launch parameters never change between resets (a one-shot search never pays for
    // the capture).  BP_GRAPH=0 keeps plain launches; the legacy NULL stream cannot be captured.
    static const bool graph_ok = !(getenv("BP_GRAPH") && atoi(getenv("BP_GRAPH")) == 0);
    s->n_runs++;
    const int graph_key = hybrid ? 1 + handover : 0;
    if (s->graph_exec && s->graph_key != graph_key) {          // the search changed form: capture again
        cudaGraphExecDestroy(s->graph_exec);
        s->graph_exec = nullptr;
    }
    s->graph_key = graph_key;
    if (graph_ok && s->ctx->stream != nullptr && (s->n_runs >= 2 || s->graph_exec)) {
        if (!s->graph_exec) {
            const int64_t before = s->ctx->launches;
            cudaGraph_t graph = nullptr;
            BP_CUDA(cudaStreamBeginCapture(s->ctx->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = issue();
            const cudaError_t end = cudaStreamEndCapture(s->ctx->stream, &graph);
            s->graph_launches = (int)(s->ctx->launches - before);
            s->ctx->launches = before;
            if (rc || end != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                set_error("CUDA graph capture of the search failed: %s", cudaGetErrorString(end));
                return BP_ERR_CUDA;
            }
            const cudaError_t inst = cudaGraphInstantiate(&s->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            if (inst != cudaSuccess) {
                s->graph_exec = nullptr;
                set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(inst));
                return BP_ERR_CUDA;
            }
        }
        BP_CUDA(cudaGraphLaunch(s->graph_exec, s->ctx->stream));
        s->ctx->launches += s->graph_launches;
    } else {
        const int rc = issue();
        if (rc) return rc;
    }
    s->ran_resident = hybrid;
    s->cur_depth = s->max_depth;
    return BP_OK;
}

int bp_search_results(bp_search* s, bp_search_result* results, uint8_t* order, int32_t* path,
                      int32_t* child_scores)
{
    BP_REQUIRE(s && results, "null argument");
    bp_context* ctx = s->ctx;
    const SearchParams& p = s->p;
    BP_CUDA(cudaSetDevice(ctx->device));
    const size_t P = p.n_problems;
    // the result block (status .. order) is contiguous in the arena: one copy, then views
    cudaStream_t st = ctx->stream;
    const size_t res_bytes = (size_t)(s->res_end - s->zero_begin);
    s->h_results.resize(res_bytes);
    BP_CUDA(cudaMemcpyAsync(s->h_results.data(), s->zero_begin, res_bytes, cudaMemcpyDeviceToHost, st));
    auto view = [&](const void* dev) {
        return s->h_results.data() + (reinterpret_cast<const char*>(dev) - s->zero_begin);
    };
    const int* status = reinterpret_cast<const int*>(view(p.status));
    const int* depth = reinterpret_cast<const int*>(view(p.depth));
    const int* n_order = reinterpret_cast<const int*>(view(p.n_order));
    const unsigned long long* ntp = reinterpret_cast<const unsigned long long*>(view(p.n_tiles_placed));
    const unsigned long long* roll = reinterpret_cast<const unsigned long long*>(view(p.rollouts));
    const unsigned long long* plc = reinterpret_cast<const unsigned long long*>(view(p.rollout_placements));
    const unsigned long long* exe = reinterpret_cast<const unsigned long long*>(view(p.rollouts_executed));
    const unsigned long long* pexe = reinterpret_cast<const unsigned long long*>(view(p.placements_executed));
    if (path)
        BP_CUDA(cudaMemcpyAsync(path, p.path, P * (p.es / 2) * 4, cudaMemcpyDeviceToHost, st));
    if (child_scores)
        BP_CUDA(cudaMemcpyAsync(child_scores, p.child_scores, P * (p.es / 2) * p.es * 4,
                                cudaMemcpyDeviceToHost, st));
    BP_CUDA(cudaStreamSynchronize(st));
#ifdef BP_TIMELINE
    if (s->d_timeline && getenv("BP_TIMELINE_OUT")) {
        const size_t n = (size_t)(MAX_DEPTHS + 1) * FS_CHAINS * 2;
        std::vector<unsigned long long> h(n * 4);
        BP_CUDA(cudaMemcpy(h.data(), s->d_timeline, n * 32, cudaMemcpyDeviceToHost));
        unsigned long long base = ~0ull;
        for (size_t i = 0; i < n; ++i)
            if (h[4 * i + 3]) base = std::min(base, ~h[4 * i]);
        if (FILE* f = fopen(getenv("BP_TIMELINE_OUT"), "w")) {
            for (size_t k = 0; k < s->classes.size(); ++k)
                fprintf(f, "{\"chain\": %d, \"problems\": %d, \"depths\": %d}\n", s->classes[k],
                        s->cls_count[s->classes[k]], s->cls_depth[k]);
            for (size_t i = 0; i < n; ++i)
                if (h[4 * i + 3])
                    fprintf(f, "{\"depth\": %d, \"chain\": %d, \"kind\": \"%s\", \"start_us\": %.2f, "
                               "\"end_us\": %.2f, \"warp_us\": %.1f, \"warps\": %llu}\n",
                            (int)(i / 2 / FS_CHAINS), (int)(i / 2 % FS_CHAINS), (i & 1) ? "advance" : "rollout",
                            (~h[4 * i] - base) * 1e-3, (h[4 * i + 1] - base) * 1e-3, h[4 * i + 2] * 1e-3,
                            h[4 * i + 3]);
            fclose(f);
        }
    }
#endif
    if (s->ran_resident) {
        // the resettable part of the queue header came back with the result block
        memcpy(reinterpret_cast<char*>(&s->h_queue) + RS_QUEUE_RESET_OFFSET, s->h_results.data(),
               sizeof(UnitQueue) - RS_QUEUE_RESET_OFFSET);
        const UnitQueue& q = s->h_queue;
        if (q.abort || q.in_flight != 0) {
            int stuck = -1, n_stuck = 0;
            for (size_t i = 0; i < P; ++i)
                if (status[i] == ST_ACTIVE) {
                    if (stuck < 0) stuck = (int)i;
                    n_stuck++;
                }
            s->queue_dirty = true;
            set_error("resident search kernel gave up (abort=%d, %d problems in flight): %s; %d active, "
                      "first: problem %d depth %d",
                      q.abort, q.in_flight,
                      q.abort == 2 ? "unit limit exceeded" : "a warp waited too long for work or for a peer",
                      n_stuck, stuck, stuck >= 0 ? depth[stuck] : -1);
            return BP_ERR_STATE;
        }
    }
    if (order) memcpy(order, view(p.order), P * (p.es / 2 + 2) * 2);
    for (size_t i = 0; i < P; ++i) {
        bp_search_result& r = results[i];
        r.depth = depth[i];
        r.solution_found = status[i] == ST_SOLVED ? 1 : 0;
        r.score = r.solution_found ? 0 : s->h_desc[i].n_entries / 2 - depth[i];
        r.n_order = n_order[i];
        r.n_tiles_placed = (int64_t)ntp[i];
        r.rollouts = (int64_t)roll[i];
        r.rollout_placements = (int64_t)plc[i];
        r.rollouts_executed = (int64_t)exe[i];
        r.placements_executed = (int64_t)pexe[i];
    }
    return BP_OK;
}

int bp_flat_search(bp_context* ctx, int n_problems, const int32_t* rows, const int32_t* cols,
                   const int32_t* n_entries, const uint8_t* tiles, int entry_stride,
                   const uint32_t* boards, const uint32_t* streams, int n_sim, int strategy,
                   uint32_t seed, bp_search_result* results, uint8_t* order, int32_t* path,
                   int32_t* child_scores)
{
    // BP_TIMING=1: host wall time of each phase on stderr (synchronising after each)
    static const bool timing = getenv("BP_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::micro>(b - a).count();
    };
    const auto t0 = now();
    bp_search* s = nullptr;
    int rc = search_create(ctx, n_problems, rows, cols, n_entries, tiles, entry_stride, boards,
                           streams, n_sim, strategy, seed, true, &s);
    if (rc) return rc;
    const auto t1 = now();
    rc = bp_search_reset(s);
    if (timing) cudaStreamSynchronize(ctx->stream);
    const auto t2 = now();
    if (!rc) rc = bp_search_run(s);
    const auto t3 = now();
    if (timing) cudaStreamSynchronize(ctx->stream);
    const auto t4 = now();
    if (!rc) rc = bp_search_results(s, results, order, path, child_scores);
    const auto t5 = now();
    bp_search_destroy(s);
    const auto t6 = now();
    if (timing)
        fprintf(stderr, "[bp_flat_search] create %.0f us, reset %.0f, run (issue) %.0f, run (device) %.0f, "
                        "results %.0f, destroy %.0f\n",
                us(t0, t1), us(t1, t2), us(t2, t3), us(t3, t4), us(t4, t5), us(t5, t6));
    return rc;
}

}  // extern "C"
