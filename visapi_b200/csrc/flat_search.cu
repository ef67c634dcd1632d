// Flat Monte-Carlo search (CustomMCTS.predict, binpack/MCTS.py:44-149) resident on the
// device: many problems in flight, every depth = one rollout launch + one advance launch
// per shape class, no host round trip until the results are read.
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
// A rollout never touches HBM except to read its root (once per 32 rollouts) and to
// write 8 bytes per 32 rollouts; it is bound by instruction issue, not bandwidth.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "board.cuh"
#include "common.cuh"
#include "lanes.cuh"

namespace bp {

constexpr int CHUNK = 32;              // rollouts per task = one result per lane
constexpr int ADV_WARPS = 4;
constexpr uint32_t NONE_U32 = 0xFFFFFFFFu;
constexpr int MAX_DEPTHS = BP_MAX_ENTRIES / 2 + 1;
// search class = (height planes the rows need: 5..8) x (warp shape class RHO, W, EJ)
// (+ 4 on the plane index: boards of 65..96 columns, whose lane kernel runs with three words per
// plane instead of four; the warp-per-board kernels of that class keep W = 4)
constexpr int FS_CLASSES = 8 * N_CLASSES;

enum Status : int { ST_ACTIVE = 0, ST_SOLVED = 1, ST_DEAD = 2 };

struct ProblemDesc {
    int rows, cols, n_entries, cls;
    uint32_t stream;
    int pad0, pad1, pad2;
};

struct __align__(8) ChunkRec {
    uint8_t max_steps;       // max placements over the chunk's rollouts
    uint8_t solved_lane;     // first solved rollout of the chunk, 0xFF = none
    uint16_t sum_steps;      // placements of all rollouts of the chunk
    uint16_t prefix_steps;   // placements of rollouts 0..solved_lane (== sum when none)
    uint16_t n_sims;
};

struct __align__(8) ChildStats {
    int32_t max_steps;
    uint32_t first_solved;   // global sim index of the first solved rollout, NONE_U32 = none
    long long sum_steps;     // placements over all of this rank's rollouts
    long long prefix_steps;  // placements up to and including first_solved (this rank)
};

struct SearchParams {
    int n_problems, es, rs, ws;
    int n_sim, strategy, n_chunks;          // n_chunks: chunks of THIS rank
    int sim_begin, n_local, rank, n_ranks;
    uint32_t seed;
    const ProblemDesc* desc;
    uint32_t* occ;
    uint8_t* tiles;
    int4* lfb;
    uint4* legal;
    int* status;
    int* depth;
    unsigned long long* n_tiles_placed;
    unsigned long long* rollouts;
    unsigned long long* rollout_placements;
    unsigned long long* rollouts_executed;
    unsigned long long* placements_executed;
    uint16_t* prev_tile;
    int* n_order;
    uint8_t* order;          // [P][ES/2+2][2]
    int* path;               // [P][ES/2]
    int* child_scores;       // [P][ES/2][ES]
    ChunkRec* rec;           // [P][ES][NC]
    ChildStats* stats;       // [P][ES]
    uint32_t* child_list;    // per class segments
    unsigned* child_count;   // [MAX_DEPTHS + 1][FS_CLASSES]
    unsigned* task_counter;  // [MAX_DEPTHS + 1][FS_CLASSES]
    // lane-per-rollout tables of the current state of each problem (lanes.cuh)
    int use_lanes;
    uint32_t* planes;        // [P][8][4] bit-sliced column heights, inverted
    uint32_t* pairs;         // [P][ES][4] bits of the two entries that leave when e is played
    uint4* tile;             // [P][ES] (h, w, width mask (1 << w) - 1 as lo / hi word)
    uint32_t* wmask;         // [P][129][4] entries with w <= x
    uint32_t* hmask;         // [P][129][4] entries with h <= y
    // resident (single-launch) search: units of the current depth of each problem still running
    int* pending;            // [P]
    int res_flags;           // debugging knobs (BP_RESIDENT_FLAGS): 1 = single-chunk units, 2 = never leave
};

// ---- work queue of the resident search kernel (one per search class) -------------------
// A unit = `chunk_count` consecutive 32-rollout chunks of one child.  Producers (the warp that
// advances a problem) reserve a ticket range with one atomicAdd on `tail`, write the payloads,
// fence, then write each slot's `seq` = ticket + 1.  Consumers take a ticket from `head` and
// wait for their slot's seq.  The ring holds every unit that can be outstanding at once.
struct UnitQueue {
    // every hot word on its own 128-byte line: the ticket counters take one atomic per unit,
    // the flags are polled by waiting warps
    unsigned long long head;     // tickets handed to consumers
    unsigned pad_h[30];
    unsigned long long tail;     // tickets reserved by producers
    unsigned pad_t[30];
    int in_flight;               // problems of this class still searching
    int abort;                   // != 0: a warp gave up waiting (watchdog) or saw a torn slot
    int est_units;               // decaying average of units per published problem-depth
    unsigned pad_f[29];
    unsigned init_next;          // prologue: next problem whose depth-0 units are to be published
    unsigned init_done;          // prologue: problems published
    unsigned cap_mask;           // ring capacity - 1
    unsigned pad_i[29];
    // where the warps' time went (SM clock cycles summed over warps, added when a warp leaves)
    unsigned long long cyc_total, cyc_wait, cyc_play, cyc_advance, n_units_played, n_advances;
    unsigned pad_c[20];
};
static_assert(sizeof(UnitQueue) == 640, "queue header layout");
struct __align__(16) UnitSlot {
    uint32_t seq;                // (uint32_t)(ticket + 1) once the payload is valid
    uint32_t child;              // problem | entry << 24
    uint32_t chunk_begin, chunk_count;
};
constexpr int MAX_UNITS_PER_CHILD = 256;

// loads of state another warp of the SAME launch may have written: through L2 (ld.global.cg),
// never from this SM's L1
template <bool CG, typename T>
__device__ __forceinline__ T ld_state(const T* ptr)
{
    if (CG) return __ldcg(ptr);
    return *ptr;
}
template <typename T>
__device__ __forceinline__ T ld_volatile(const T* ptr) { return *reinterpret_cast<const volatile T*>(ptr); }
template <typename T>
__device__ __forceinline__ void st_volatile(T* ptr, T v) { *reinterpret_cast<volatile T*>(ptr) = v; }

template <bool CG, int RHO, int W, int EJ>
__device__ __forceinline__ void load_board_t(WarpState<RHO, W, EJ>& s, const uint32_t* occ, int rows,
                                             int cols, int wpr)
{
    const int lane = lane_id();
#pragma unroll
    for (int k = 0; k < RHO; ++k) {
        const int row = lane + 32 * k;
#pragma unroll
        for (int q = 0; q < W; ++q) {
            uint32_t v = FULLMASK;
            if (row < rows && q < wpr) {
                v = ld_state<CG>(occ + row * wpr + q);
                const int live = cols - 32 * q;
                if (live < 32) v |= (live <= 0) ? FULLMASK : (FULLMASK << live);
            }
            s.occ[k][q] = v;
        }
    }
}

template <bool CG, int RHO, int W, int EJ>
__device__ __forceinline__ void load_entries_t(WarpState<RHO, W, EJ>& s, const uint8_t* tiles, int n_entries)
{
    const int lane = lane_id();
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(tiles);
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const int e = lane + 32 * j;
        s.ent[j] = (e < n_entries) ? (uint32_t)ld_state<CG>(t16 + e) : 0u;
    }
}

__device__ __forceinline__ long long warp_sum_ll(long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}

// ------------------------------------------------------------------ rollout kernel
// Persistent warps pull (child, chunk) tasks; a task builds the child board from its
// parent (one placement) and plays CHUNK rollouts from it.
template <int RHO, int W, int EJ>
__global__ void __launch_bounds__(128)
fs_rollout_kernel(SearchParams p, int cls, int cls_list_offset, int depth_slot)
{
    const int lane = lane_id();
    const unsigned n_children = p.child_count[depth_slot * FS_CLASSES + cls];
    const unsigned long long n_tasks = (unsigned long long)n_children * p.n_chunks;
    unsigned* counter = &p.task_counter[depth_slot * FS_CLASSES + cls];
    for (;;) {
        unsigned long long task = 0;
        if (lane == 0) task = atomicAdd(counter, 1u);
        task = __shfl_sync(FULLMASK, task, 0);
        if (task >= n_tasks) break;
        const unsigned child = (unsigned)(task / p.n_chunks);
        const int chunk = (int)(task % p.n_chunks);
        const uint32_t packed = p.child_list[cls_list_offset + child];
        const int prob = (int)(packed & 0xFFFFFFu);
        const int entry = (int)(packed >> 24);
        const ProblemDesc d = p.desc[prob];
        const int4 lfb = p.lfb[prob];

        WarpState<RHO, W, EJ> s;
        load_board(s, p.occ + (size_t)prob * p.rs * p.ws, d.rows, d.cols, p.ws);
        load_entries(s, p.tiles + (size_t)prob * p.es * 2, d.n_entries);
        uint32_t mine = s.ent[0];
#pragma unroll
        for (int j = 1; j < EJ; ++j) mine = ((entry >> 5) == j) ? s.ent[j] : mine;
        const uint32_t v = __shfl_sync(FULLMASK, mine, entry & 31);
        place(s, lfb.x, lfb.y, (int)(v & 0xFFu), (int)(v >> 8));
        kill_pair(s, v);
        const int n_alive = lfb.w - 2;
        const uint32_t tag = ((uint32_t)p.depth[prob] << 16) | (uint32_t)entry;

        const int first = chunk * CHUNK;
        const int count = min(CHUNK, p.n_local - first);
        int my_steps = 0;
        bool my_solved = false;
        for (int i = 0; i < count; ++i) {
            bool solved;
            const int steps = rollout<RHO, W, EJ, false>(
                s, d.rows, n_alive, p.seed, d.stream, tag, (uint32_t)(p.sim_begin + first + i),
                solved, nullptr);
            if (lane == i) { my_steps = steps; my_solved = solved; }
        }
        const uint32_t solved_mask = __ballot_sync(FULLMASK, my_solved);
        const int first_lane = solved_mask ? (__ffs(solved_mask) - 1) : 0xFF;
        const int mx = __reduce_max_sync(FULLMASK, my_steps);
        const int sum = __reduce_add_sync(FULLMASK, my_steps);
        const int pre = __reduce_add_sync(FULLMASK, (lane <= first_lane) ? my_steps : 0);
        if (lane == 0) {
            ChunkRec r;
            r.max_steps = (uint8_t)mx;
            r.solved_lane = (uint8_t)first_lane;
            r.sum_steps = (uint16_t)sum;
            r.prefix_steps = (uint16_t)pre;
            r.n_sims = (uint16_t)count;
            p.rec[((size_t)prob * p.es + entry) * p.n_chunks + chunk] = r;
        }
    }
}

// ------------------------------------------------------- lane-per-rollout kernel
// Same tasks and the same 8-byte record as fs_rollout_kernel, but lane i plays rollout i of
// the chunk on its own bit-sliced board (lanes.cuh): ~25x fewer warp instructions per
// placement.  Used when every problem starts from an empty board (skyline invariant) and equal
// tiles are adjacent in its tile list.
// Shared memory of one warp of the lane kernels: the tables of the problem state it is playing.
template <int NB, int W, int EW>
struct LaneTables {
    static constexpr int WROWS = 32 * W + 1, HROWS = (1 << NB) + 1, NE = 32 * EW;
    uint32_t wmask[WROWS * EW];
    uint32_t hmask[HROWS * EW];
    LaneEntry<W, EW> entries[NE];
};

// One work unit: chunks [chunk_begin, chunk_end) of child `entry` of problem `prob`.  The
// problem's tables (shared memory) are refilled when `cached_key` (problem, depth) changes; the
// child board is built once and every chunk writes its 8-byte record.
// CG: the state may have been written by another warp of this launch (resident search).
template <bool CG, int NB, int W, int EW>
__device__ __forceinline__ void lanes_play_unit(const SearchParams& p, LaneTables<NB, W, EW>& tb,
                                                const uint8_t* sel8, int prob, int entry,
                                                int chunk_begin, int chunk_end, int& cached_key)
{
    constexpr int WROWS = LaneTables<NB, W, EW>::WROWS, HROWS = LaneTables<NB, W, EW>::HROWS,
                  NE = LaneTables<NB, W, EW>::NE;
    const int lane = lane_id();
    const ProblemDesc d = p.desc[prob];
    const int4 lfb = ld_state<CG>(p.lfb + prob);      // (hmin, col, run, live entries) of the parent
    const int depth = ld_state<CG>(p.depth + prob);
    const int key = prob | (depth << 24);

    if (key != cached_key) {                          // tables of this problem's current state
        __syncwarp();
        const uint32_t* gw = p.wmask + (size_t)prob * LANES_TABLE_ROWS * LANES_MASK_WORDS;
        const uint32_t* gh = p.hmask + (size_t)prob * LANES_TABLE_ROWS * LANES_MASK_WORDS;
        for (int i = lane; i < WROWS * EW; i += 32) {
            const int x = i / EW, q = i % EW;
            tb.wmask[i] = (x <= d.cols) ? ld_state<CG>(gw + x * LANES_MASK_WORDS + q) : 0u;
        }
        // row inv of the shared table = entries with h <= rows - hmin, hmin = 2^NB - 1 - inv
        for (int i = lane; i < HROWS * EW; i += 32) {
            const int y = i / EW - ((1 << NB) - 1 - d.rows), q = i % EW;
            tb.hmask[i] = (y >= 0 && y <= d.rows) ? ld_state<CG>(gh + y * LANES_MASK_WORDS + q) : 0u;
        }
        const uint4* gg = reinterpret_cast<const uint4*>(p.pairs) + (size_t)prob * p.es;
        const uint4* gt = p.tile + (size_t)prob * p.es;
        for (int i = lane; i < NE; i += 32) {
            const bool in = i < d.n_entries;
            const uint4 pr = in ? ld_state<CG>(gg + i) : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t pair[4] = {pr.x, pr.y, pr.z, pr.w};
            tb.entries[i].set(in ? ld_state<CG>(gt + i) : make_uint4(0u, 0u, 0u, 0u), pair);
        }
        __syncwarp();
        cached_key = key;
    }

    // the child: parent + entry placed at the parent's first free cell, pair removed
    LaneBoard<NB, W, EW> s;
    const uint32_t* gp = p.planes + (size_t)prob * LANES_PLANE_STRIDE;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int q = 0; q < W; ++q) s.pl[b][q] = ld_state<CG>(gp + b * 4 + q);
#pragma unroll
    for (int q = 0; q < EW; ++q) s.alive[q] = word_range(0, lfb.w, q);
    const uint4 et = ld_state<CG>(p.tile + (size_t)prob * p.es + entry);
    uint32_t foot[W];
#pragma unroll
    for (int q = 0; q < W; ++q) foot[q] = word_range(lfb.y, lfb.y + (int)et.y, q);
    lanes_place(s, foot, lfb.x, lfb.x + (int)et.x);
#pragma unroll
    for (int q = 0; q < EW; ++q) s.alive[q] &= ~tb.entries[entry].pair(q);
    const int n_alive = lfb.w - 2;
    const uint32_t tag = ((uint32_t)depth << 16) | (uint32_t)entry;

    for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
        const int first = chunk * CHUNK;
        const int count = min(CHUNK, p.n_local - first);
        int my_solved;
        const int my_steps = lanes_rollout<NB, W, EW>(
            s, d.rows, n_alive, tb.wmask, tb.hmask, tb.entries, sel8, p.seed, d.stream, tag,
            (uint32_t)(p.sim_begin + first + lane), lane < count ? 1 : 0, my_solved);

        const uint32_t solved_mask = __ballot_sync(FULLMASK, my_solved != 0);
        const int first_lane = solved_mask ? (__ffs(solved_mask) - 1) : 0xFF;
        const int mx = __reduce_max_sync(FULLMASK, my_steps);
        const int sum = __reduce_add_sync(FULLMASK, my_steps);
        const int pre = __reduce_add_sync(FULLMASK, (lane <= first_lane) ? my_steps : 0);
        if (lane == 0) {
            ChunkRec r;
            r.max_steps = (uint8_t)mx;
            r.solved_lane = (uint8_t)first_lane;
            r.sum_steps = (uint16_t)sum;
            r.prefix_steps = (uint16_t)pre;
            r.n_sims = (uint16_t)count;
            p.rec[((size_t)prob * p.es + entry) * p.n_chunks + chunk] = r;
        }
    }
}

template <int NB, int W, int EW>
__global__ void __launch_bounds__(128)
fs_rollout_lanes_kernel(SearchParams p, int cls, int cls_list_offset, int depth_slot, int units_per_warp)
{
    __shared__ LaneTables<NB, W, EW> s_tb[4];
    __shared__ uint8_t s_sel8[256 * 8];
    fill_select_table(s_sel8);
    __syncthreads();
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const unsigned n_children = p.child_count[depth_slot * FS_CLASSES + cls];
    // Work unit = `group` consecutive 32-rollout chunks of ONE child: the problem's tables (shared
    // memory) and the child board (registers) are set up once per unit.  Units stay small
    // when there is little work (load balance) and grow up to a whole child when there is
    // plenty (>= units_per_warp units per resident warp; see bp_search_run for the choice).
    const unsigned long long n_tasks = (unsigned long long)n_children * p.n_chunks;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * 4ull;
    const int group = (int)min((unsigned long long)p.n_chunks,
                               max(1ull, n_tasks / (n_warps * (unsigned long long)units_per_warp)));
    const int groups_per_child = (p.n_chunks + group - 1) / group;
    const unsigned long long n_units = (unsigned long long)n_children * groups_per_child;
    unsigned* counter = &p.task_counter[depth_slot * FS_CLASSES + cls];
    int cached_key = -1;
    for (;;) {
        unsigned long long unit = 0;
        if (lane == 0) unit = atomicAdd(counter, 1u);
        unit = __shfl_sync(FULLMASK, unit, 0);
        if (unit >= n_units) break;
        const unsigned child = (unsigned)(unit / groups_per_child);
        const int chunk_begin = (int)(unit % groups_per_child) * group;
        const int chunk_end = min(chunk_begin + group, p.n_chunks);
        const uint32_t packed = p.child_list[cls_list_offset + child];
        lanes_play_unit<false>(p, s_tb[warp], s_sel8, (int)(packed & 0xFFFFFFu), (int)(packed >> 24),
                               chunk_begin, chunk_end, cached_key);
    }
}

// Tables the lane-per-rollout kernel reads, for the state `s` of one problem that was reached
// by placing a (ph x pw) tile at row pr, column pc of the previous state (ph == 0: `s` is the
// initial, empty board).  The lane kernel only runs on boards that start empty, so a board is
// its skyline and the planes are updated, not rebuilt: the columns under the tile all had height
// pr (the tile sits on the row-major first free cell) and now have height pr + ph.
template <bool CG, int RHO, int W, int EJ>
__device__ __forceinline__ void write_lane_tables(const SearchParams& p, int prob, const ProblemDesc& d,
                                                  const WarpState<RHO, W, EJ>& s, int n_alive, int pr,
                                                  int pc, int ph, int pw)
{
    const int lane = lane_id();
    // ---- column heights, bit-sliced and inverted: lane = plane * 4 + word -----------------
    {
        uint32_t* gp = p.planes + (size_t)prob * LANES_PLANE_STRIDE;
        const int b = lane >> 2, q = lane & 3;
        uint32_t word;
        if (ph == 0) {
            word = word_range(0, d.cols, q);           // height 0 in real columns, padding never free
        } else {
            const uint32_t diff = (uint32_t)((255 - pr) ^ (255 - pr - ph));
            word = ld_state<CG>(gp + lane);
            if ((diff >> b) & 1u) word ^= word_range(pc, pc + pw, q);
        }
        gp[lane] = word;
    }

    // ---- per-entry tile record -------------------------------------------------------------
    uint4* gt = p.tile + (size_t)prob * p.es;
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const int e = lane + 32 * j;
        if (e < d.n_entries) {
            const uint32_t wdt = s.ent[j] >> 8;
            const unsigned long long wm = wdt >= 64u ? ~0ull : ((1ull << wdt) - 1ull);
            gt[e] = make_uint4(s.ent[j] & 0xFFu, wdt, (uint32_t)wm, (uint32_t)(wm >> 32));
        }
    }

    // ---- pair masks: entry e leaves with its partner, the entry of the rotated tile with the
    // same rank inside its group of equal entries (a square pairs with its neighbour: rank ^ 1).
    // Equal entries are adjacent (host-checked), so a group is a run of the list: B = the
    // entries that start a run; lane = entry; one warp-uniform pass over the runs tells every
    // entry where the run of its rotation starts and how long it is.
    uint32_t B[EJ];
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        uint32_t pv = __shfl_up_sync(FULLMASK, s.ent[j], 1);
        const uint32_t last_prev = __shfl_sync(FULLMASK, s.ent[j > 0 ? j - 1 : 0], 31);
        if (lane == 0) pv = j > 0 ? last_prev : 0u;
        B[j] = __ballot_sync(FULLMASK, s.ent[j] != 0u && s.ent[j] != pv);
    }
    int start[EJ], rfirst[EJ], rcount[EJ];
    uint32_t rot[EJ];
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        rot[j] = ((s.ent[j] & 0xFFu) << 8) | (s.ent[j] >> 8);
        rfirst[j] = -1;
        rcount[j] = 0;
        int st = -1;                                     // start of this entry's own run
#pragma unroll
        for (int jj = EJ - 1; jj >= 0; --jj) {
            if (jj <= j && st < 0) {
                const uint32_t m = (jj == j) ? (B[jj] & (FULLMASK >> (31 - lane))) : B[jj];
                if (m) st = 32 * jj + 31 - __clz(m);
            }
        }
        start[j] = st;
    }
    {
        int run_start = -1;                              // the run being closed (warp-uniform)
        uint32_t run_value = 0u;
        auto close_run = [&](int end) {
            if (run_start >= 0) {
#pragma unroll
                for (int j = 0; j < EJ; ++j)
                    if (s.ent[j] != 0u && rot[j] == run_value) {
                        rfirst[j] = run_start;
                        rcount[j] = end - run_start;
                    }
            }
        };
#pragma unroll
        for (int jg = 0; jg < EJ; ++jg) {
            uint32_t m = B[jg];
            while (m) {                                  // warp-uniform
                const int bit = __ffs(m) - 1;
                m &= m - 1u;
                const uint32_t v = __shfl_sync(FULLMASK, s.ent[jg], bit);
                close_run(32 * jg + bit);
                run_start = 32 * jg + bit;
                run_value = v;
            }
        }
        close_run(n_alive);
    }
    uint4* gg = reinterpret_cast<uint4*>(p.pairs) + (size_t)prob * p.es;
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const int e = lane + 32 * j;
        if (e < n_alive) {
            const int rank = e - start[j];
            const int want = (s.ent[j] == rot[j]) ? (rank ^ 1) : rank;
            const int partner = (rfirst[j] >= 0 && want < rcount[j]) ? rfirst[j] + want : -1;
            uint32_t out[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                out[q] = ((e >> 5) == q) ? (1u << (e & 31)) : 0u;
                if (partner >= 0 && (partner >> 5) == q) out[q] |= 1u << (partner & 31);
            }
            gg[e] = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }

    // ---- wmask[x] = entries with w <= x, hmask[y] = entries with h <= y ------------------
    uint32_t* gw = p.wmask + (size_t)prob * LANES_TABLE_ROWS * LANES_MASK_WORDS;
    uint32_t* gh = p.hmask + (size_t)prob * LANES_TABLE_ROWS * LANES_MASK_WORDS;
    for (int x = 0; x <= d.cols; ++x) {
        uint32_t m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < EJ; ++j)
            m[j] = __ballot_sync(FULLMASK, s.ent[j] != 0u && (int)(s.ent[j] >> 8) <= x);
        if (lane < 4) gw[x * LANES_MASK_WORDS + lane] = lane == 0 ? m[0] : (lane == 1 ? m[1] : (lane == 2 ? m[2] : m[3]));
    }
    for (int y = 0; y <= d.rows; ++y) {
        uint32_t m[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < EJ; ++j)
            m[j] = __ballot_sync(FULLMASK, s.ent[j] != 0u && (int)(s.ent[j] & 0xFFu) <= y);
        if (lane < 4) gh[y * LANES_MASK_WORDS + lane] = lane == 0 ? m[0] : (lane == 1 ? m[1] : (lane == 2 ? m[2] : m[3]));
    }
}

// --------------------------------------------------------- per-child statistics
// this rank's statistics of child `entry` of problem `prob`, from its chunk records
template <bool CG>
__device__ __forceinline__ ChildStats stats_from_records(const SearchParams& p, int prob, int entry)
{
    const ChunkRec* rec = p.rec + ((size_t)prob * p.es + entry) * p.n_chunks;
    ChildStats st;
    st.max_steps = 0;
    st.first_solved = NONE_U32;
    st.sum_steps = 0;
    st.prefix_steps = 0;
#pragma unroll 8
    for (int ch = 0; ch < p.n_chunks; ++ch) {          // independent 8-byte loads, 8 in flight
        const uint2 raw = ld_state<CG>(reinterpret_cast<const uint2*>(rec + ch));
        ChunkRec r;
        memcpy(&r, &raw, sizeof(r));
        st.max_steps = max(st.max_steps, (int)r.max_steps);
        st.sum_steps += r.sum_steps;
        if (st.first_solved == NONE_U32) {
            if (r.solved_lane != 0xFF) {
                st.first_solved = (uint32_t)(p.sim_begin + ch * CHUNK + r.solved_lane);
                st.prefix_steps += r.prefix_steps;
            } else {
                st.prefix_steps += r.sum_steps;
            }
        }
    }
    return st;
}

// merge the per-rank statistics (ranks hold ascending contiguous sim ranges)
__device__ __forceinline__ ChildStats stats_from_gathered(const SearchParams& p,
                                                          const ChildStats* gathered, int prob,
                                                          int entry)
{
    ChildStats st;
    st.max_steps = 0;
    st.first_solved = NONE_U32;
    st.sum_steps = 0;
    st.prefix_steps = 0;
    for (int g = 0; g < p.n_ranks; ++g) {
        const ChildStats r = gathered[((size_t)g * p.n_problems + prob) * p.es + entry];
        st.max_steps = max(st.max_steps, r.max_steps);
        st.sum_steps += r.sum_steps;
        if (st.first_solved == NONE_U32) {
            st.prefix_steps += r.prefix_steps;     // == r.sum_steps when r has no solved sim
            st.first_solved = r.first_solved;
        }
    }
    return st;
}

template <int EJ>
__global__ void __launch_bounds__(ADV_WARPS * 32)
fs_reduce_kernel(SearchParams p, const int* __restrict__ cls_problems, int n_cls_problems)
{
    const int idx = blockIdx.x * ADV_WARPS + (threadIdx.x >> 5);
    if (idx >= n_cls_problems) return;
    const int prob = cls_problems[idx];
    if (p.status[prob] != ST_ACTIVE) return;
    const int lane = lane_id();
    const uint4 lg = p.legal[prob];
    const uint32_t legal[4] = {lg.x, lg.y, lg.z, lg.w};
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        if ((legal[j] >> lane) & 1u) {
            const int e = lane + 32 * j;
            p.stats[(size_t)prob * p.es + e] = stats_from_records<false>(p, prob, e);
        }
    }
}

// compacted live tile list (order kept) through a shared-memory buffer of the warp
template <int RHO, int W, int EJ>
__device__ __forceinline__ void compact_entries_adv(WarpState<RHO, W, EJ>& s, uint16_t* buf)
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

// ---- resident search: publish the units of a problem's new depth ------------------------
// Called by a whole warp after the problem's state, legal mask and lane tables are written.
// `ebuf` (shared, >= 32 * EJ entries) receives the legal entries in table order.
template <int EJ>
__device__ __forceinline__ void publish_units(const SearchParams& p, UnitQueue* q, UnitSlot* slots,
                                              int prob, const uint32_t (&legal)[EJ], int k,
                                              uint16_t* ebuf, unsigned n_warps)
{
    const int lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    __syncwarp();
    int base = 0;
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        if ((legal[j] >> lane) & 1u) ebuf[base + __popc(legal[j] & lt)] = (uint16_t)(lane + 32 * j);
        base += __popc(legal[j]);
    }
    __syncwarp();
    // Unit size: whole children while the problems in flight offer >= 2 units per resident warp,
    // single chunks when few problems are left (their critical path is what remains).
    // (read by one lane: the count changes under our feet and every lane must agree on the split)
    const int in_flight = max(__shfl_sync(FULLMASK, ld_volatile(&q->in_flight), 0), 1);
    const long long want = (long long)in_flight * k * p.n_chunks / (2ll * n_warps);
    const int group_min = (p.n_chunks + MAX_UNITS_PER_CHILD - 1) / MAX_UNITS_PER_CHILD;
    int group = (int)min((long long)p.n_chunks, max((long long)group_min, want));
    if (p.res_flags & 1) group = group_min;
    const int gpc = (p.n_chunks + group - 1) / group;
    group = (p.n_chunks + gpc - 1) / gpc;                      // equal units
    const int n_units = k * gpc;
    unsigned long long t0 = 0;
    if (lane == 0) {
        st_volatile(p.pending + prob, n_units);
        t0 = atomicAdd(&q->tail, (unsigned long long)n_units);
        const int est = ld_volatile(&q->est_units);
        st_volatile(&q->est_units, est == 0 ? n_units : (est * 7 + n_units + 7) / 8);   // racy by design
    }
    t0 = __shfl_sync(FULLMASK, t0, 0);
    const unsigned mask = q->cap_mask;
    for (int i = lane; i < n_units; i += 32) {
        const int ci = i / gpc, g = i - ci * gpc;
        UnitSlot* sl = slots + ((t0 + i) & mask);
        st_volatile(&sl->child, (uint32_t)prob | ((uint32_t)ebuf[ci] << 24));
        st_volatile(&sl->chunk_begin, (uint32_t)(g * group));
        st_volatile(&sl->chunk_count, (uint32_t)min(group, p.n_chunks - g * group));
    }
    // state, tables, pending and payloads before any seq: every lane fences its own writes,
    // the warp barrier orders them before the seq stores of all lanes
    __threadfence();
    __syncwarp();
    for (int i = lane; i < n_units; i += 32)
        st_volatile(&slots[(t0 + i) & mask].seq, (uint32_t)(t0 + i + 1));
    __syncwarp();
}

// ------------------------------------------------------------------ advance
// One warp, one problem.  FIRST: set up depth 0.  Otherwise: score the children of the
// current depth (lane = child), stop at the first child with a solved rollout, else commit
// the first maximum; then expand the new state (LFB, legal entries, lane tables) and hand the
// children of the next depth to the rollout kernels: appended to the class's child list
// (one launch per depth) or, RESIDENT, published as units on the queue.
// Returns the problem's status.
template <int RHO, int W, int EJ, bool FIRST, bool RESIDENT>
__device__ __forceinline__ int advance_problem(const SearchParams& p, int prob, int cls,
                                               int cls_list_offset, int next_slot,
                                               const ChildStats* __restrict__ gathered,
                                               uint16_t* cbuf, UnitQueue* q, UnitSlot* slots,
                                               unsigned n_warps)
{
    constexpr bool CG = RESIDENT;
    const int lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    const ProblemDesc d = p.desc[prob];

    WarpState<RHO, W, EJ> s;
    int n_alive;
    int placed_r = 0, placed_c = 0, placed_h = 0, placed_w = 0;      // the committed placement
    if (FIRST) {
        load_board_t<CG>(s, p.occ + (size_t)prob * p.rs * p.ws, d.rows, d.cols, p.ws);
        load_entries_t<CG>(s, p.tiles + (size_t)prob * p.es * 2, d.n_entries);
        compact_entries_adv(s, cbuf);
        n_alive = count_alive(s);
    } else {
        if (ld_state<CG>(p.status + prob) != ST_ACTIVE) return ST_DEAD;
        const int depth = ld_state<CG>(p.depth + prob);
        const int4 lfb = ld_state<CG>(p.lfb + prob);
        n_alive = lfb.w;
        const uint4 lg = ld_state<CG>(p.legal + prob);
        const uint32_t legal[4] = {lg.x, lg.y, lg.z, lg.w};
        load_board_t<CG>(s, p.occ + (size_t)prob * p.rs * p.ws, d.rows, d.cols, p.ws);
        load_entries_t<CG>(s, p.tiles + (size_t)prob * p.es * 2, d.n_entries);

        // ---- per-child statistics, lane = child ---------------------------------
        ChildStats st[EJ];
        uint32_t solved_m[EJ];
        int first_child = -1;
#pragma unroll
        for (int j = 0; j < EJ; ++j) {
            const int e = lane + 32 * j;
            const bool is_child = (legal[j] >> lane) & 1u;
            st[j].max_steps = 0; st[j].first_solved = NONE_U32;
            st[j].sum_steps = 0; st[j].prefix_steps = 0;
            if (is_child)
                st[j] = gathered ? stats_from_gathered(p, gathered, prob, e)
                                 : stats_from_records<CG>(p, prob, e);
            solved_m[j] = __ballot_sync(FULLMASK, is_child && st[j].first_solved != NONE_U32);
            if (first_child < 0 && solved_m[j]) first_child = 32 * j + __ffs(solved_m[j]) - 1;
        }

        // ---- what the reference would have counted (MCTS.py:60,76,134-140) -------
        long long roll = 0, plc = 0, plc_exec = 0;
        int best_score = -1, best_entry = -1;
        int* scores = p.child_scores + ((size_t)prob * (p.es / 2) + depth) * p.es;
#pragma unroll
        for (int j = 0; j < EJ; ++j) {
            const int e = lane + 32 * j;
            const bool is_child = (legal[j] >> lane) & 1u;
            int score = -1;
            if (is_child) {
                plc_exec += st[j].sum_steps;
                if (first_child < 0 || e < first_child) {
                    roll += p.n_sim;
                    plc += st[j].sum_steps;
                    score = p.strategy == BP_MAX_DEPTH ? st[j].max_steps : (int)st[j].sum_steps;
                } else if (e == first_child) {
                    roll += (long long)st[j].first_solved + 1;
                    plc += st[j].prefix_steps;
                    score = -2;
                }
            }
            if (e < p.es) scores[e] = score;
            // first maximum in entry order (get_max_index, MCTS.py:20-29)
            const int key_score = (is_child && score >= 0) ? score : -1;
            const int mx = __reduce_max_sync(FULLMASK, key_score);
            if (mx > best_score) {                               // strict: earlier j wins ties
                const uint32_t who = __ballot_sync(FULLMASK, key_score == mx);
                best_score = mx;
                best_entry = 32 * j + __ffs(who) - 1;
            }
        }
        roll = warp_sum_ll(roll);
        plc = warp_sum_ll(plc);
        plc_exec = warp_sum_ll(plc_exec);
        if (lane == 0) {        // only this problem's advancing warp touches these; L2 atomics
            atomicAdd(p.placements_executed + prob, (unsigned long long)plc_exec);
            atomicAdd(p.rollouts + prob, (unsigned long long)roll);
            atomicAdd(p.rollout_placements + prob, (unsigned long long)plc);
            atomicAdd(p.n_tiles_placed + prob, (unsigned long long)plc);
        }

        const int take = first_child >= 0 ? first_child : best_entry;
        uint32_t mine = s.ent[0];
#pragma unroll
        for (int j = 1; j < EJ; ++j) mine = ((take >> 5) == j) ? s.ent[j] : mine;
        const uint32_t v = __shfl_sync(FULLMASK, mine, take & 31);
        const uint16_t prev = ld_state<CG>(p.prev_tile + prob);
        int n_order = ld_state<CG>(p.n_order + prob);
        uint8_t* order = p.order + (size_t)prob * (p.es / 2 + 2) * 2;

        place(s, lfb.x, lfb.y, (int)(v & 0xFFu), (int)(v >> 8));
        kill_pair(s, v);
        placed_r = lfb.x; placed_c = lfb.y; placed_h = (int)(v & 0xFFu); placed_w = (int)(v >> 8);

        if (first_child >= 0) {
            // ---- solved inside a rollout (MCTS.py:77-93) -------------------------
            if (lane == 0) {
                // entries after the solving child were never attempted (:58-60)
                atomicAdd(p.n_tiles_placed + prob,
                          0ull - (unsigned long long)(n_alive - first_child - 1));
                if (prev) { order[2 * n_order] = prev & 0xFF; order[2 * n_order + 1] = prev >> 8; n_order++; }
                order[2 * n_order] = v & 0xFF; order[2 * n_order + 1] = v >> 8; n_order++;
            }
            n_order = __shfl_sync(FULLMASK, n_order, 0);
            // replay the winning rollout to recover its move list
            uint32_t sim = NONE_U32;
#pragma unroll
            for (int j = 0; j < EJ; ++j) {
                const uint32_t cand = __shfl_sync(FULLMASK, st[j].first_solved, first_child & 31);
                sim = ((first_child >> 5) == j) ? cand : sim;
            }
            bool solved;
            const uint32_t tag = ((uint32_t)depth << 16) | (uint32_t)first_child;
            const int steps = rollout<RHO, W, EJ, true>(s, d.rows, n_alive - 2, p.seed, d.stream, tag,
                                                        sim, solved, order + 2 * n_order);
            if (lane == 0) {
                p.n_order[prob] = n_order + steps;
                p.status[prob] = solved ? ST_SOLVED : ST_DEAD;   // solved is always true here
            }
            return solved ? ST_SOLVED : ST_DEAD;
        }

        // ---- commit the best child (MCTS.py:105-119) -----------------------------
        compact_entries_adv(s, cbuf);
        n_alive -= 2;
        if (lane == 0) {
            if (prev) { order[2 * n_order] = prev & 0xFF; order[2 * n_order + 1] = prev >> 8; n_order++; }
            p.n_order[prob] = n_order;
            p.prev_tile[prob] = (uint16_t)v;
            p.path[(size_t)prob * (p.es / 2) + depth] = best_entry;
            p.depth[prob] = depth + 1;
        }
        store_board(s, p.occ + (size_t)prob * p.rs * p.ws, d.rows, d.cols, p.ws);
    }
    store_entries(s, p.tiles + (size_t)prob * p.es * 2, d.n_entries);

    // ---- expand: LFB, legal entries, children of the next depth ----------------
    int status = ST_ACTIVE;
    int r = -1, c = -1, run = 0, k = 0;
    uint32_t legal[EJ];
#pragma unroll
    for (int j = 0; j < EJ; ++j) legal[j] = 0u;
    if (n_alive == 0) {
        status = ST_SOLVED;                                  // MCTS.py:121-123
    } else if (!find_lfb(s, r, c, run)) {
        status = ST_DEAD;                                    // full board, tiles left
    } else {
        k = legal_masks(s, d.rows, r, run, legal);
        if (k == 0) status = ST_DEAD;                        // MCTS.py:101-103
    }
    unsigned base = 0;
    if (lane == 0) {
        p.status[prob] = status;
        p.lfb[prob] = make_int4(r, c, run, n_alive);
        uint4 lg = make_uint4(legal[0], EJ > 1 ? legal[EJ > 1 ? 1 : 0] : 0u,
                              EJ > 2 ? legal[EJ > 2 ? 2 : 0] : 0u, EJ > 3 ? legal[EJ > 3 ? 3 : 0] : 0u);
        p.legal[prob] = lg;
        if (status != ST_SOLVED && n_alive > 0 && r >= 0)
            atomicAdd(p.n_tiles_placed + prob, (unsigned long long)n_alive);   // one attempt per entry (:60)
        if (k > 0) {
            if (!RESIDENT) base = atomicAdd(&p.child_count[next_slot * FS_CLASSES + cls], (unsigned)k);
            atomicAdd(p.rollouts_executed + prob, (unsigned long long)k * p.n_local);
        }
    }
    if (k > 0) {
        if (!RESIDENT) {
            base = __shfl_sync(FULLMASK, base, 0);
#pragma unroll
            for (int j = 0; j < EJ; ++j) {
                if ((legal[j] >> lane) & 1u)
                    p.child_list[cls_list_offset + base + __popc(legal[j] & lt)] =
                        (uint32_t)prob | ((uint32_t)(lane + 32 * j) << 24);
                base += __popc(legal[j]);
            }
        }
        if (p.use_lanes) write_lane_tables<CG>(p, prob, d, s, n_alive, placed_r, placed_c, placed_h, placed_w);
        if (RESIDENT) publish_units<EJ>(p, q, slots, prob, legal, k, cbuf, n_warps);
    }
    return status;
}

template <int RHO, int W, int EJ, bool FIRST>
__global__ void __launch_bounds__(ADV_WARPS * 32)
fs_advance_kernel(SearchParams p, int cls, int cls_list_offset,
                  const int* __restrict__ cls_problems, int n_cls_problems, int next_slot,
                  const ChildStats* __restrict__ gathered)
{
    __shared__ uint16_t cbuf[ADV_WARPS][32 * EJ];
    const int warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * ADV_WARPS + warp;
    if (idx >= n_cls_problems) return;
    advance_problem<RHO, W, EJ, FIRST, false>(p, cls_problems[idx], cls, cls_list_offset, next_slot,
                                              gathered, cbuf[warp], nullptr, nullptr, 0u);
}

// Out of line on purpose: the resident kernel is compiled under the register cap of its rollout
// loop (resident_blocks), and the advance -- which runs once per problem and depth -- keeps its
// own, larger working set (spilling if it must) without costing the loop a register.
template <int RHO, int W, int EJ>
__device__ __noinline__ int advance_resident(const SearchParams& p, int prob, uint16_t* cbuf,
                                             UnitQueue* q, UnitSlot* slots, unsigned n_warps)
{
    return advance_problem<RHO, W, EJ, false, true>(p, prob, 0, 0, 0, nullptr, cbuf, q, slots, n_warps);
}

// resident blocks per SM = what the rollout loop alone (fs_rollout_lanes_kernel) reaches
constexpr int resident_blocks(int nb, int w, int ew)
{
    const int wi = w == 1 ? 0 : (w == 2 ? 1 : 2), ei = ew == 4 ? 1 : 0;
    constexpr int tab[4][3][2] = {{{9, 8}, {8, 7}, {5, 5}},
                                  {{9, 8}, {7, 7}, {5, 4}},
                                  {{9, 7}, {7, 6}, {4, 4}},
                                  {{9, 6}, {7, 5}, {4, 4}}};
    return tab[nb - 5][wi][ei];
}

// ------------------------------------------------------- resident search kernel
// The whole flat search of one search class in ONE launch.  Persistent warps take units from
// the class queue and play them (lanes_play_unit); the warp that completes the last unit of a
// problem's depth advances that problem itself (advance_problem) and publishes the units of
// its next depth.  Problems therefore move through their depths independently: no per-depth
// launches, no barrier between depths, no tail per depth.
//   * visibility: records / state / tables written by one warp are read by another inside
//     the same launch, so those reads go through L2 (ld.global.cg) and every hand-over is
//     fence -> atomic / volatile flag -> fence.
//   * leaving: a warp that is about to wait while more warps are already waiting than the
//     problems in flight can feed leaves the kernel, so the SM slots go to the next class's
//     kernel queued on another stream; everyone leaves when no problem is in flight.
//   * watchdog: a warp that waits ~seconds for its slot raises `abort`; the host reports it.
template <int NB, int RHO, int W, int EJ>
__global__ void __launch_bounds__(128, resident_blocks(NB, W, EJ))
fs_search_resident_kernel(SearchParams p, UnitQueue* q, UnitSlot* slots,
                          const int* __restrict__ cls_problems, int n_cls_problems)
{
    constexpr int EW = EJ;
    static_assert(EJ == 2 || EJ == 4, "entry words of the lane kernel");
    __shared__ LaneTables<NB, W, EW> s_tb[4];
    __shared__ uint8_t s_sel8[256 * 8];
    __shared__ uint16_t s_cbuf[4][32 * EJ];
    fill_select_table(s_sel8);
    __syncthreads();
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const unsigned n_warps = gridDim.x * 4u;

    // ---- prologue: publish depth 0 (state and tables were written by bp_search_reset) -----
    for (;;) {
        unsigned idx = 0;
        if (lane == 0) idx = atomicAdd(&q->init_next, 1u);
        idx = __shfl_sync(FULLMASK, idx, 0);
        if (idx >= (unsigned)n_cls_problems) break;
        const int prob = cls_problems[idx];
        const uint4 lg = p.legal[prob];
        const uint32_t lgw[4] = {lg.x, lg.y, lg.z, lg.w};
        uint32_t legal[EJ];
        int k = 0;
#pragma unroll
        for (int j = 0; j < EJ; ++j) { legal[j] = lgw[j]; k += __popc(legal[j]); }
        if (p.status[prob] == ST_ACTIVE && k > 0)
            publish_units<EJ>(p, q, slots, prob, legal, k, s_cbuf[warp], n_warps);
        else if (lane == 0)
            atomicSub(&q->in_flight, 1);
        if (lane == 0) { __threadfence(); atomicAdd(&q->init_done, 1u); }
    }
    if (lane == 0) {
        unsigned spins = 0;
        while (ld_volatile(&q->init_done) < (unsigned)n_cls_problems && !ld_volatile(&q->abort)) {
            __nanosleep(1000);
            if (++spins > (1u << 22)) st_volatile(&q->abort, 1);
        }
    }
    __syncwarp();

    int cached_key = -1;
    bool waited = true;          // lane 0: the previous ticket was not ready when taken
    const long long c_begin = clock64();
    long long c_wait = 0, c_play = 0, c_adv = 0;
    unsigned n_played = 0, n_adv = 0;
    for (;;) {
        // ---- take a ticket, or leave ----------------------------------------------------
        // (the leave rule reads the queue's hot lines, so only a warp that just had to wait
        // for work evaluates it)
        unsigned long long ticket = 0;
        int leave = 0;
        if (lane == 0) {
            if (waited) {
                const int in_flight = ld_volatile(&q->in_flight);
                if (in_flight <= 0 || ld_volatile(&q->abort)) {
                    leave = 1;
                } else {
                    const long long waiters =
                        (long long)(ld_volatile(&q->head) - ld_volatile(&q->tail));
                    const long long feed = 2ll * in_flight * max(ld_volatile(&q->est_units), 1) + 4;
                    if (waiters >= feed && !(p.res_flags & 2)) leave = 1;
                }
            }
            if (!leave) ticket = atomicAdd(&q->head, 1ull);
        }
        leave = __shfl_sync(FULLMASK, leave, 0);
        if (leave) break;
        ticket = __shfl_sync(FULLMASK, ticket, 0);

        // ---- wait for the slot ------------------------------------------------------------
        const long long c0 = clock64();
        UnitSlot* sl = slots + (ticket & q->cap_mask);
        const uint32_t want = (uint32_t)(ticket + 1);
        uint4 unit = make_uint4(0u, 0u, 0u, 0u);
        if (lane == 0) {
            // poll this ticket's own slot; the shared flags only every 16th time (thousands of
            // idle warps hammering one line would starve the atomics of the working ones)
            unsigned spins = 0, ns = 64;
            waited = false;
            for (;;) {
                if (ld_volatile(&sl->seq) == want) {
                    __threadfence();
                    unit.y = ld_volatile(&sl->child);
                    unit.z = ld_volatile(&sl->chunk_begin);
                    unit.w = ld_volatile(&sl->chunk_count);
                    __threadfence();
                    if (ld_volatile(&sl->seq) != want) st_volatile(&q->abort, 2);   // ring overrun
                    unit.x = 1u;
                    break;
                }
                if ((spins & 15u) == 15u &&
                    (ld_volatile(&q->in_flight) <= 0 || ld_volatile(&q->abort))) break;
                waited = true;
                __nanosleep(ns);
                if (ns < 4096) ns *= 2;
                if (++spins > (1u << 21)) { st_volatile(&q->abort, 1); break; }
            }
        }
        unit.x = __shfl_sync(FULLMASK, unit.x, 0);
        if (!unit.x) break;
        unit.y = __shfl_sync(FULLMASK, unit.y, 0);
        unit.z = __shfl_sync(FULLMASK, unit.z, 0);
        unit.w = __shfl_sync(FULLMASK, unit.w, 0);
        const int prob = (int)(unit.y & 0xFFFFFFu), entry = (int)(unit.y >> 24);
        const long long c1 = clock64();
        c_wait += c1 - c0;

        lanes_play_unit<true>(p, s_tb[warp], s_sel8, prob, entry, (int)unit.z, (int)(unit.z + unit.w),
                              cached_key);

        // ---- the last unit of this problem's depth advances it ----------------------------
        int last = 0;
        if (lane == 0) {
            __threadfence();                                   // records before the count
            last = atomicSub(p.pending + prob, 1) == 1;
        }
        last = __shfl_sync(FULLMASK, last, 0);
        const long long c2 = clock64();
        c_play += c2 - c1;
        n_played++;
        if (last) {
            __threadfence();
            const int status = advance_resident<RHO, W, EJ>(p, prob, s_cbuf[warp], q, slots, n_warps);
            if (status != ST_ACTIVE && lane == 0) {
                __threadfence();
                atomicSub(&q->in_flight, 1);
            }
            c_adv += clock64() - c2;
            n_adv++;
        }
    }
    if (lane == 0) {
        atomicAdd(&q->cyc_total, (unsigned long long)(clock64() - c_begin));
        atomicAdd(&q->cyc_wait, (unsigned long long)c_wait);
        atomicAdd(&q->cyc_play, (unsigned long long)c_play);
        atomicAdd(&q->cyc_advance, (unsigned long long)c_adv);
        atomicAdd(&q->n_units_played, (unsigned long long)n_played);
        atomicAdd(&q->n_advances, (unsigned long long)n_adv);
    }
}

}  // namespace bp

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

// resident-kernel instantiation of a resident class index: ((nb - 5) * 3 + w index) * 2 + (EW == 4)
constexpr int RS_CLASSES = 24;
static int resident_class_of(int scls)
{
    const int cls = scls % N_CLASSES, nb = 5 + (scls / N_CLASSES) % 4;
    const int ej = cls % 3, w = (cls / 3) % 3;
    return ((nb - 5) * 3 + w) * 2 + (ej == 2 ? 1 : 0);
}
// f(IC<NB>, IC<RHO>, IC<W>, IC<EJ>)
template <typename F>
static void dispatch_resident(int rcls, F&& f)
{
    const int ew = rcls % 2, w = (rcls / 2) % 3, nb = 5 + rcls / 6;
    auto with_ew = [&](auto NBc, auto R, auto Wc) {
        if (ew == 0) f(NBc, R, Wc, IC<2>{}); else f(NBc, R, Wc, IC<4>{});
    };
    auto with_w = [&](auto NBc, auto R) {
        switch (w) {
            case 0: with_ew(NBc, R, IC<1>{}); break;
            case 1: with_ew(NBc, R, IC<2>{}); break;
            default: with_ew(NBc, R, IC<4>{}); break;
        }
    };
    switch (nb) {
        case 5: with_w(IC<5>{}, IC<1>{}); break;
        case 6: with_w(IC<6>{}, IC<2>{}); break;
        case 7: with_w(IC<7>{}, IC<4>{}); break;
        default: with_w(IC<8>{}, IC<4>{}); break;
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
    int cls_count[FS_CLASSES] = {0};
    int cls_prob_off[FS_CLASSES] = {0};          // offset into d_cls_problems
    int cls_list_off[FS_CLASSES] = {0};          // offset into child_list
    int grid_rollout[FS_CLASSES] = {0};
    int grid_lanes[FS_CLASSES] = {0};
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
    bool queues_ready = false;                  // resident queues initialised since the last reset
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
    // resident search (one launch per resident class): queues, rings, problem lists
    int mode = BP_SEARCH_AUTO;
    bool ran_resident = false;
    std::vector<int> rclasses;                  // resident classes present
    int rcls_count[RS_CLASSES] = {0};
    int rcls_prob_off[RS_CLASSES] = {0};
    size_t rcls_slot_off[RS_CLASSES] = {0};
    int grid_resident[RS_CLASSES] = {0};
    std::vector<UnitQueue> h_queues;            // initial queue headers
    int* d_rcls_problems = nullptr;
    UnitQueue* d_queues = nullptr;              // [RS_CLASSES]
    UnitQueue* d_queues_init = nullptr;
    UnitSlot* d_slots = nullptr;
    size_t n_slots = 0;
    std::vector<cudaStream_t> rcls_stream;
    std::vector<cudaEvent_t> rcls_done;
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
    std::vector<std::vector<int>> by_class(FS_CLASSES);
    for (int i = 0; i < n_problems; ++i) {
        ProblemDesc& d = s->h_desc[i];
        d.rows = rows[i]; d.cols = cols[i]; d.n_entries = n_entries[i];
        const int nb = rows[i] <= 31 ? 5 : (rows[i] <= 63 ? 6 : (rows[i] <= 127 ? 7 : 8));
        const int three_words = (cols[i] > 64 && cols[i] <= 96) ? 4 : 0;
        d.cls = (nb - 5 + three_words) * N_CLASSES + shape_class(rows[i], cols[i], n_entries[i]);
        d.stream = streams ? streams[i] : (uint32_t)i;
        d.pad0 = d.pad1 = d.pad2 = 0;
        by_class[d.cls].push_back(i);
    }

    std::vector<int> cls_problems;
    int list_off = 0;
    for (int c = 0; c < FS_CLASSES; ++c) {
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
    for (int i = 0; i < n_problems && lanes_ok; ++i) {       // equal tiles adjacent in the list?
        const uint8_t* t = tiles + (size_t)i * entry_stride * 2;
        std::vector<int> closed;
        int prev = -1;
        for (int e = 0; e < n_entries[i]; ++e) {
            const int v = t[2 * e] | (t[2 * e + 1] << 8);
            if (v == 0 || v == prev) continue;
            if (std::find(closed.begin(), closed.end(), v) != closed.end()) { lanes_ok = false; break; }
            if (prev >= 0) closed.push_back(prev);
            prev = v;
        }
    }
    p.use_lanes = lanes_ok ? 1 : 0;
    if (const char* env = getenv("BP_RESIDENT_FLAGS")) p.res_flags = atoi(env);

    // resident classes: problem lists, ring capacities (every unit that can be outstanding)
    std::vector<int> rcls_problems;
    s->h_queues.assign(RS_CLASSES, UnitQueue{});
    if (p.use_lanes) {
        std::vector<std::vector<int>> by_rcls(RS_CLASSES);
        for (int i = 0; i < n_problems; ++i) by_rcls[resident_class_of(s->h_desc[i].cls)].push_back(i);
        const size_t upc = (size_t)std::min(p.n_chunks, MAX_UNITS_PER_CHILD);
        for (int c = 0; c < RS_CLASSES; ++c) {
            s->rcls_count[c] = (int)by_rcls[c].size();
            s->rcls_prob_off[c] = (int)rcls_problems.size();
            s->rcls_slot_off[c] = s->n_slots;
            if (!s->rcls_count[c]) continue;
            s->rclasses.push_back(c);
            rcls_problems.insert(rcls_problems.end(), by_rcls[c].begin(), by_rcls[c].end());
            size_t need = 0;
            for (int i : by_rcls[c]) need += (size_t)s->h_desc[i].n_entries * upc;
            size_t cap = 1024;
            while (cap < need) cap <<= 1;
            s->n_slots += cap;
            UnitQueue& q = s->h_queues[c];
            q.in_flight = s->rcls_count[c];
            q.cap_mask = (unsigned)(cap - 1);
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
        if (p.use_lanes) {
            TRY(dev_alloc(s, &s->d_rcls_problems, rcls_problems.size()));
            TRY(dev_alloc(s, &s->d_queues_init, (size_t)RS_CLASSES));
        }
        s->up_end = mark();
        if (boards) TRY(dev_alloc(s, &s->d_init_occ, P * p.rs * p.ws));
        s->zero_begin = mark();
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
        s->counters_bytes = (size_t)(MAX_DEPTHS + 1) * FS_CLASSES * sizeof(unsigned);
        TRY(dev_alloc(s, &p.child_count, (size_t)(MAX_DEPTHS + 1) * FS_CLASSES));
        TRY(dev_alloc(s, &p.task_counter, (size_t)(MAX_DEPTHS + 1) * FS_CLASSES));
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
            TRY(dev_alloc(s, &s->d_queues, (size_t)RS_CLASSES));
            TRY(dev_alloc(s, &s->d_slots, s->n_slots));
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
    if (p.use_lanes) {
        put(s->d_rcls_problems, rcls_problems.data(), rcls_problems.size() * sizeof(int));
        put(s->d_queues_init, s->h_queues.data(), RS_CLASSES * sizeof(UnitQueue));
    }
    BP_CUDA(cudaMemcpyAsync(s->up_begin, s->h_upload.data(), s->h_upload.size(), cudaMemcpyHostToDevice,
                            ctx->stream));
    if (boards)
        BP_CUDA(cudaMemcpyAsync(s->d_init_occ, boards, P * p.rs * p.ws * 4, cudaMemcpyHostToDevice,
                                ctx->stream));
    BP_CUDA(cudaStreamSynchronize(ctx->stream));

    // persistent grid per class: every SM filled to the occupancy limit of that kernel
    for (int c : s->classes) {
        if (!ctx->occ_rollout[c]) {
            int per_sm = 1;
            dispatch_class(c % N_CLASSES, [&](auto R, auto Wc, auto E) {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                    &per_sm, fs_rollout_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>,
                    128, 0);
            });
            ctx->occ_rollout[c] = std::max(per_sm, 1);
        }
        s->grid_rollout[c] = ctx->sm_count * ctx->occ_rollout[c];
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
            s->grid_lanes[c] = ctx->sm_count * ctx->occ_lanes[c];
        }
    }
    for (int c : s->rclasses) {
        if (!ctx->occ_resident[c]) {
            int per_sm = 1;
            dispatch_resident(c, [&](auto NBc, auto R, auto Wc, auto E) {
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                    &per_sm,
                    fs_search_resident_kernel<decltype(NBc)::value, decltype(R)::value, decltype(Wc)::value,
                                              decltype(E)::value>,
                    128, 0);
            });
            ctx->occ_resident[c] = std::max(per_sm, 1);
        }
        s->grid_resident[c] = ctx->sm_count * ctx->occ_resident[c];
    }
    // side streams and fork / join events come from the context's pool
    if (!ctx->fork_event) BP_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    s->fork_ev = ctx->fork_event;
    for (size_t k = 0; k < s->rclasses.size(); ++k) {
        cudaStream_t st = side_stream(ctx, k);
        cudaEvent_t e = side_event(ctx, k);
        if (!st || !e) { set_error("cannot create a side stream / event"); bp_search_destroy(s); return BP_ERR_CUDA; }
        s->rcls_stream.push_back(st);
        s->rcls_done.push_back(e);
    }
    for (size_t k = 0; k < s->classes.size(); ++k) {
        cudaStream_t st = side_stream(ctx, k);
        cudaEvent_t e = side_event(ctx, k);
        if (!st || !e) { set_error("cannot create a side stream / event"); bp_search_destroy(s); return BP_ERR_CUDA; }
        s->cls_stream.push_back(st);
        s->cls_done.push_back(e);
        const int c = s->classes[k];
        int md = 0;
        for (int i = 0; i < n_problems; ++i)
            if (s->h_desc[i].cls == c) md = std::max(md, s->h_desc[i].n_entries / 2);
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
    BP_REQUIRE(sim_begin >= 0 && n_local >= 0 && sim_begin + n_local <= s->p.n_sim, "sim range");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    BP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    s->p.rank = rank;
    s->p.n_ranks = n_ranks;
    s->p.sim_begin = sim_begin;
    s->p.n_local = n_local;
    s->p.n_chunks = std::max(1, (n_local + CHUNK - 1) / CHUNK);
    s->ready = false;
    return BP_OK;    // the record buffer was sized for all n_sim rollouts
}

}  // extern "C"

template <bool FIRST>
static void launch_advance_class(bp_search* s, int c, int next_slot, const ChildStats* gathered,
                                 cudaStream_t st)
{
    const int n = s->cls_count[c];
    const int grid = (n + ADV_WARPS - 1) / ADV_WARPS;
    dispatch_class(c % N_CLASSES, [&](auto R, auto Wc, auto E) {
        fs_advance_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value, FIRST>
            <<<grid, ADV_WARPS * 32, 0, st>>>(s->p, c, s->cls_list_off[c],
                                              s->d_cls_problems + s->cls_prob_off[c], n, next_slot,
                                              gathered);
    });
    s->ctx->launches++;
}

template <bool FIRST>
static void launch_advance(bp_search* s, int next_slot, const ChildStats* gathered)
{
    for (int c : s->classes) launch_advance_class<FIRST>(s, c, next_slot, gathered, s->ctx->stream);
}

// `overlapped`: other classes' launches run beside this one on their own streams.  Measured on
// problems_g (4 classes, 1000 problems, N = 1000; profiles/r1): with >= 8 units per resident warp
// a whole search takes 13.12 ms, with 4 / 2 / 1 it takes 12.61 / 12.27 / 12.00 ms -- larger units
// amortise the per-unit table fill and queue atomic, and the longer tail of a launch is filled
// by the other classes' kernels.  A class running alone has nobody to fill its tail: there the
// sum of the rollout kernels is 13.52 / 13.53 / 14.04 / 14.88 ms, so it keeps small units.
static void launch_rollouts_class(bp_search* s, int c, int depth, cudaStream_t st, bool overlapped)
{
    int upw = overlapped ? 1 : 8;
    if (const char* env = getenv("BP_UNITS_PER_WARP")) upw = std::max(1, atoi(env));
    if (s->p.use_lanes)
        dispatch_lanes(c, [&](auto NBc, auto Wc, auto E) {
            fs_rollout_lanes_kernel<decltype(NBc)::value, decltype(Wc)::value, decltype(E)::value>
                <<<s->grid_lanes[c], 128, 0, st>>>(s->p, c, s->cls_list_off[c], depth, upw);
        });
    else
        dispatch_class(c % N_CLASSES, [&](auto R, auto Wc, auto E) {
            fs_rollout_kernel<decltype(R)::value, decltype(Wc)::value, decltype(E)::value>
                <<<s->grid_rollout[c], 128, 0, st>>>(s->p, c, s->cls_list_off[c], depth);
        });
    s->ctx->launches++;
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
    s->queues_ready = false;
    launch_advance<true>(s, 0, nullptr);
    BP_CUDA(cudaGetLastError());
    s->ran_resident = false;
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
        s->ev.resize((size_t)3 * std::max(s->max_depth + 1, RS_CLASSES));
        for (cudaEvent_t& e : s->ev) BP_CUDA(cudaEventCreate(&e));
    }
    return BP_OK;
}

int bp_search_set_mode(bp_search* s, int mode)
{
    BP_REQUIRE(s != nullptr, "search is null");
    BP_REQUIRE(mode == BP_SEARCH_AUTO || mode == BP_SEARCH_STEPPED || mode == BP_SEARCH_RESIDENT, "mode");
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
        const int ej = (c % N_CLASSES) % 3 == 0 ? 1 : ((c % N_CLASSES) % 3 == 1 ? 2 : 4);
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
    *resident = s->ran_resident ? 1 : 0;
    return BP_OK;
}

int bp_search_resident_stats(bp_search* s, double* out8)
{
    BP_REQUIRE(s && out8, "null argument");
    BP_CUDA(cudaSetDevice(s->ctx->device));
    BP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    for (int i = 0; i < 8; ++i) out8[i] = 0.0;
    if (!s->ran_resident) return BP_OK;
    std::vector<UnitQueue> queues(RS_CLASSES);
    BP_CUDA(cudaMemcpy(queues.data(), s->d_queues, RS_CLASSES * sizeof(UnitQueue), cudaMemcpyDeviceToHost));
    for (int c : s->rclasses) {
        const UnitQueue& q = queues[c];
        out8[0] += (double)q.cyc_total;
        out8[1] += (double)q.cyc_wait;
        out8[2] += (double)q.cyc_play;
        out8[3] += (double)q.cyc_advance;
        out8[4] += (double)q.n_units_played;
        out8[5] += (double)q.n_advances;
        out8[6] += (double)q.head;
        out8[7] += (double)q.tail;
    }
    return BP_OK;
}

int bp_search_max_depth(bp_search* s, int* max_depth)
{
    BP_REQUIRE(s && max_depth, "null argument");
    *max_depth = s->max_depth;
    return BP_OK;
}

int bp_search_run(bp_search* s)
{
    BP_REQUIRE(s != nullptr, "search is null");
    if (s->p.n_ranks != 1) {
        set_error("bp_search_run is single-rank; use the step API with a partition");
        return BP_ERR_STATE;
    }
    // AUTO = the faster form on this hardware, measured: launch-per-depth
    int mode = s->mode;
    if (const char* env = getenv("BP_SEARCH")) {
        if (strcmp(env, "resident") == 0) mode = BP_SEARCH_RESIDENT;
        else if (strcmp(env, "stepped") == 0) mode = BP_SEARCH_STEPPED;
    }
    if (mode == BP_SEARCH_RESIDENT && !s->p.use_lanes) {
        set_error("the resident search kernel needs the lane kernel (empty initial boards, equal "
                  "tiles adjacent in every list; BP_ROLLOUT != warp)");
        return BP_ERR_STATE;
    }
    if (mode == BP_SEARCH_RESIDENT) {
        // resident search: one launch per resident class, each on its own stream (the block
        // scheduler hands the SM slots of warps that leave one class's kernel to the next);
        // when profiling, serially on the caller's stream with an event pair around each
        if (!s->ready || s->cur_depth != 0) {
            set_error("bp_search_run: call bp_search_reset first");
            return BP_ERR_STATE;
        }
        BP_CUDA(cudaSetDevice(s->ctx->device));
        if (!s->queues_ready) {
            BP_CUDA(cudaMemcpyAsync(s->d_queues, s->d_queues_init, RS_CLASSES * sizeof(UnitQueue),
                                    cudaMemcpyDeviceToDevice, s->ctx->stream));
            BP_CUDA(cudaMemsetAsync(s->d_slots, 0, s->n_slots * sizeof(UnitSlot), s->ctx->stream));
            s->queues_ready = true;
        }
        const bool fork = !s->profiling && s->rclasses.size() > 1;
        if (fork) {
            BP_CUDA(cudaEventRecord(s->fork_ev, s->ctx->stream));
            for (size_t k = 0; k < s->rclasses.size(); ++k)
                BP_CUDA(cudaStreamWaitEvent(s->rcls_stream[k], s->fork_ev, 0));
        }
        for (size_t k = 0; k < s->rclasses.size(); ++k) {
            const int c = s->rclasses[k];
            cudaStream_t st = fork ? s->rcls_stream[k] : s->ctx->stream;
            if (s->profiling) BP_CUDA(cudaEventRecord(s->ev[3 * k], st));
            dispatch_resident(c, [&](auto NBc, auto R, auto Wc, auto E) {
                fs_search_resident_kernel<decltype(NBc)::value, decltype(R)::value, decltype(Wc)::value,
                                          decltype(E)::value>
                    <<<s->grid_resident[c], 128, 0, st>>>(s->p, s->d_queues + c,
                                                          s->d_slots + s->rcls_slot_off[c],
                                                          s->d_rcls_problems + s->rcls_prob_off[c],
                                                          s->rcls_count[c]);
            });
            s->ctx->launches++;
            if (s->profiling) {
                BP_CUDA(cudaEventRecord(s->ev[3 * k + 1], st));
                BP_CUDA(cudaEventRecord(s->ev[3 * k + 2], st));
            }
        }
        BP_CUDA(cudaGetLastError());
        if (fork) {
            for (size_t k = 0; k < s->rclasses.size(); ++k) {
                BP_CUDA(cudaEventRecord(s->rcls_done[k], s->rcls_stream[k]));
                BP_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->rcls_done[k], 0));
            }
        }
        if (s->profiling) s->ev_depths = (int)s->rclasses.size();
        s->ran_resident = true;
        s->cur_depth = s->max_depth;
        return BP_OK;
    }
    if (s->profiling || s->classes.size() == 1 || getenv("BP_SERIAL_CLASSES")) {
        s->in_profiled_run = s->profiling;
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
    // fork: every class runs its whole depth loop on its own stream, then join.  Launches are
    // issued depth-major (all classes' depth 0, then depth 1, ...) so that every chain starts
    // at once instead of waiting for the host to issue the chains before it.
    BP_CUDA(cudaSetDevice(s->ctx->device));
    BP_CUDA(cudaEventRecord(s->fork_ev, s->ctx->stream));
    for (size_t k = 0; k < s->classes.size(); ++k)
        BP_CUDA(cudaStreamWaitEvent(s->cls_stream[k], s->fork_ev, 0));
    for (int d = 0; d < s->max_depth; ++d) {
        for (size_t k = 0; k < s->classes.size(); ++k) {
            if (d >= s->cls_depth[k]) continue;
            launch_rollouts_class(s, s->classes[k], d, s->cls_stream[k], true);
            launch_advance_class<false>(s, s->classes[k], d + 1, nullptr, s->cls_stream[k]);
        }
    }
    BP_CUDA(cudaGetLastError());
    for (size_t k = 0; k < s->classes.size(); ++k) {
        BP_CUDA(cudaEventRecord(s->cls_done[k], s->cls_stream[k]));
        BP_CUDA(cudaStreamWaitEvent(s->ctx->stream, s->cls_done[k], 0));
    }
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
    std::vector<UnitQueue> queues;
    if (s->ran_resident) {
        queues.resize(RS_CLASSES);
        BP_CUDA(cudaMemcpyAsync(queues.data(), s->d_queues, RS_CLASSES * sizeof(UnitQueue),
                                cudaMemcpyDeviceToHost, st));
    }
    BP_CUDA(cudaStreamSynchronize(st));
    for (int c : s->rclasses) {
        if (!s->ran_resident) break;
        if (queues[c].abort || queues[c].in_flight != 0) {
            std::vector<int> pending(P);
            cudaMemcpy(pending.data(), p.pending, P * 4, cudaMemcpyDeviceToHost);
            int stuck = -1, n_stuck = 0;
            for (size_t i = 0; i < P; ++i)
                if (status[i] == ST_ACTIVE && resident_class_of(s->h_desc[i].cls) == c) {
                    if (stuck < 0) stuck = (int)i;
                    n_stuck++;
                }
            set_error("resident search kernel of class %d gave up (abort=%d, %d problems in flight): %s; "
                      "head=%llu tail=%llu est_units=%d init_done=%u; %d active, first: problem %d "
                      "depth %d pending %d",
                      c, queues[c].abort, queues[c].in_flight,
                      queues[c].abort == 2 ? "unit ring overrun" : "a warp waited too long for work",
                      queues[c].head, queues[c].tail, queues[c].est_units, queues[c].init_done,
                      n_stuck, stuck, stuck >= 0 ? depth[stuck] : -1, stuck >= 0 ? pending[stuck] : -1);
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
