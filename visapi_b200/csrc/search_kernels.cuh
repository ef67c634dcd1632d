// Flat search kernels with one launch per depth: the two rollout kernels (warp-per-board,
// lane-per-rollout), the per-problem advance (score children, commit, expand, lane tables) and
// the root-parallel reduce.
#pragma once
#include "search_queue.cuh"
#include "search_types.cuh"

namespace bp {

// Debug build only (-DBP_TIMELINE, tools/timeline.py): every warp of a launch-per-depth kernel
// stamps %globaltimer when it starts and when it leaves; per launch the host reads first start,
// last end, the sum of warp lifetimes and the warp count -- which kernels overlap, how full the
// GPU is in each, where a chain waits.  Not compiled into the shipped library.
#ifdef BP_TIMELINE
static __device__ unsigned long long* g_timeline;      // [launch id][4]
struct TimelineScope {
    int id;
    unsigned long long t0;
    __device__ __forceinline__ explicit TimelineScope(int launch_id) : id(launch_id), t0(global_timer_ns()) {}
    __device__ __forceinline__ ~TimelineScope()
    {
        if ((threadIdx.x & 31) == 0 && g_timeline && id >= 0) {
            const unsigned long long t1 = global_timer_ns();
            atomicMax(&g_timeline[4 * id + 0], ~t0);
            atomicMax(&g_timeline[4 * id + 1], t1);
            atomicAdd(&g_timeline[4 * id + 2], t1 - t0);
            atomicAdd(&g_timeline[4 * id + 3], 1ull);
        }
    }
};
#define BP_TIMELINE_SCOPE(launch_id) TimelineScope timeline_scope_(launch_id)
#else
#define BP_TIMELINE_SCOPE(launch_id)
#endif

// ------------------------------------------------------------------ rollout kernel
// Persistent warps pull (child, chunk) tasks; a task builds the child board from its
// parent (one placement) and plays CHUNK rollouts from it.
template <int RHO, int W, int EJ>
__global__ void __launch_bounds__(128)
fs_rollout_kernel(SearchParams p, int cls, int cls_list_offset, int depth_slot)
{
    const int lane = lane_id();
    const unsigned n_children = p.child_count[depth_slot * FS_CHAINS + cls];
    const unsigned long long n_tasks = (unsigned long long)n_children * p.n_chunks;
    unsigned* counter = &p.task_counter[depth_slot * FS_CHAINS + cls];
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
fs_rollout_lanes_kernel(SearchParams p, int cls, int cls_list_offset, int depth_slot, int half_units_per_warp,
                        int n_warps_full)
{
    BP_TIMELINE_SCOPE((depth_slot * FS_CHAINS + cls) * 2);
    __shared__ LaneTables<NB, W, EW> s_tb[4];
    __shared__ uint8_t s_sel8[256 * 8];
    fill_select_table(s_sel8);
    __syncthreads();
    // (everything above overlaps the advance launch before this one; its child list is read below)
    grid_dependency_wait();
    const unsigned n_children = p.child_count[depth_slot * FS_CHAINS + cls];
    // a block that is not needed (the blocks before it hold a warp for every 32-rollout chunk there
    // is) leaves at once: the grid is sized for the children the depth CAN have
    if ((unsigned long long)blockIdx.x * 4ull >= (unsigned long long)n_children * p.n_chunks) return;
    grid_launch_dependents();              // the advance launch may take its (few) places now
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    // Work unit = `group` consecutive 32-rollout chunks of ONE child: the problem's tables (shared
    // memory) and the child board (registers) are set up once per unit.  Units stay small
    // when there is little work (load balance) and grow up to a whole child when there is
    // plenty (>= half_units_per_warp / 2 units per resident warp; launch_rollouts_class chooses).
    const unsigned long long n_tasks = (unsigned long long)n_children * p.n_chunks;
    // (the unit size follows the warps of a FULL grid; the launch itself may hold fewer blocks when
    // the host knows that there cannot be more units than that)
    const unsigned long long n_warps = (unsigned long long)n_warps_full;
    const int group = (int)min((unsigned long long)p.n_chunks,
                               max(1ull, 2ull * n_tasks / (n_warps * (unsigned long long)half_units_per_warp)));
    const int groups_per_child = (p.n_chunks + group - 1) / group;
    const unsigned long long n_units = (unsigned long long)n_children * groups_per_child;
    unsigned* counter = &p.task_counter[depth_slot * FS_CHAINS + cls];
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
// CG: the buffer was filled by peer-memory stores during this launch (resident search)
template <bool CG>
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
        const ChildStats* src = gathered + ((size_t)g * p.n_problems + prob) * p.es + entry;
        ChildStats r;
        if (CG) {
            const int2 a = __ldcg(reinterpret_cast<const int2*>(src));
            r.max_steps = a.x;
            r.first_solved = (uint32_t)a.y;
            r.sum_steps = __ldcg(&src->sum_steps);
            r.prefix_steps = __ldcg(&src->prefix_steps);
        } else {
            r = *src;
        }
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

// ------------------------------------------------------------------ advance
// One warp, one problem.  FIRST: set up depth 0.  Otherwise: score the children of the
// current depth (lane = child), stop at the first child with a solved rollout, else commit
// the first maximum; then expand the new state (LFB, legal entries, lane tables) and hand the
// children of the next depth to the rollout kernels (appended to the class's child list).
// (The resident kernel advances problems in the skyline state instead: advance_skyline,
// search_resident.cuh.)  Returns the problem's status.
template <int RHO, int W, int EJ, bool FIRST>
__device__ __forceinline__ int advance_problem(const SearchParams& p, int prob, int cls,
                                               int cls_list_offset, int next_slot,
                                               const ChildStats* __restrict__ gathered,
                                               uint16_t* cbuf)
{
    constexpr bool CG = false;
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
                st[j] = gathered ? stats_from_gathered<CG>(p, gathered, prob, e)
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
            base = atomicAdd(&p.child_count[next_slot * FS_CHAINS + cls], (unsigned)k);
            atomicAdd(p.rollouts_executed + prob, (unsigned long long)k * p.n_local);
        }
    }
    if (k > 0) {
        base = __shfl_sync(FULLMASK, base, 0);
#pragma unroll
        for (int j = 0; j < EJ; ++j) {
            if ((legal[j] >> lane) & 1u)
                p.child_list[cls_list_offset + base + __popc(legal[j] & lt)] =
                    (uint32_t)prob | ((uint32_t)(lane + 32 * j) << 24);
            base += __popc(legal[j]);
        }
        if (p.use_lanes) write_lane_tables<CG>(p, prob, d, s, n_alive, placed_r, placed_c, placed_h, placed_w);
        if (FIRST && p.use_lanes && lane == 0) {               // skyline state of the resident search
            p.alive[prob] = make_uint4(word_range(0, n_alive, 0), word_range(0, n_alive, 1),
                                       word_range(0, n_alive, 2), word_range(0, n_alive, 3));
        }
    }
    return status;
}

template <int RHO, int W, int EJ, bool FIRST>
__global__ void __launch_bounds__(ADV_WARPS * 32)
fs_advance_kernel(SearchParams p, int cls, int cls_list_offset,
                  const int* __restrict__ cls_problems, int n_cls_problems, int next_slot,
                  const ChildStats* __restrict__ gathered)
{
    BP_TIMELINE_SCOPE(FIRST ? -1 : ((next_slot - 1) * FS_CHAINS + cls) * 2 + 1);
    // !FIRST: the next depth's rollout launch may bring its blocks in (they fill their tables and
    // wait for this launch to complete); this launch waits for the rollouts it scores.
    // FIRST (bp_search_reset): the classes' launches are independent of each other, so each lets the
    // next one start at once and they run side by side on the one stream; each waits for its
    // predecessor only before it ENDS, so that the last one's completion still means all are done.
    grid_launch_dependents();
    if (!FIRST) grid_dependency_wait();
    __shared__ uint16_t cbuf[ADV_WARPS][32 * EJ];
    const int warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * ADV_WARPS + warp;
    if (idx < n_cls_problems)
        advance_problem<RHO, W, EJ, FIRST>(p, cls_problems[idx], cls, cls_list_offset, next_slot, gathered,
                                           cbuf[warp]);
    if (FIRST) grid_dependency_wait();
}

}  // namespace bp
