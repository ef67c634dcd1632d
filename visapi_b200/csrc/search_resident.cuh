// The resident form of the flat search: ONE launch runs every depth of every problem of a
// search, all shape classes together.
//
// Persistent warps take work units (a few 32-rollout chunks of one child) from the device queue
// (search_queue.cuh) and play them with the lane-per-rollout playout of the unit's shape class;
// the warp that completes the last unit of a problem's depth advances that problem itself
// (advance_problem: score the children, commit, expand, lane tables) and publishes the units of
// its next depth.  Problems therefore move through their depths independently: no launch, no
// barrier and no tail per depth -- what a depth costs beyond its rollouts is one hand-over
// (fence, ticket, one 64-bit slot) per unit and one advance per problem.
//   * visibility: records / state / tables written by one warp are read by another inside the
//     same launch, so those reads go through L2 (ld.global.cg) and every hand-over is
//     fence -> slot store / atomic -> acquire load.
//   * root-parallel (n_ranks > 1): before the advance the warp stores this rank's per-child
//     statistics into EVERY rank's gathered buffer over NVLink peer memory, raises a flag there
//     and waits for the other ranks' flags in its own memory (exchange_child_stats) -- the
//     collective of binpack/MCTS.py:126-149 split over GPUs is part of the kernel, there is no
//     host step and no NCCL call per move.
//   * leaving: a warp whose ticket is not served leaves when no problem is in flight; the last
//     warp out sets tail = head.  A warp that waits ~4 s raises `abort`; the host reports it.
#pragma once
#include "search_kernels.cuh"

namespace bp {

// =============================================================== skyline search state
// Inside the resident kernel a problem's state is (planes, alive, lfb, legal, depth): the
// bit-sliced skyline and the set of live entries of the INITIAL tile list.  The tile list is never
// compacted, so the per-problem tables (wmask, hmask, tile records, pair masks) are written once
// by the reset and stay valid for every depth -- a warp that has a problem's tables in shared
// memory keeps them across depths, and the advance of a depth is a few table lookups instead of
// a rebuild.  Equal entries are adjacent and leave in rotation pairs of equal rank, so "the
// first equal entry" of the reference (solution_checker.py:319-332) and "this entry" leave the
// same tile sequence; indices the host sees (path, child_scores, Philox tag) are ranks among
// the live entries, i.e. positions in the reference's compacted list.

// One 16-byte record per unit: x = first solved rollout (global index, NONE_U32 = none),
// y / z = low words of the placements of all rollouts / of the rollouts up to the first solved
// one, w = max placements | high 12 bits of y << 8 | high 12 bits of z << 20.
__device__ __forceinline__ uint4 pack_unit_rec(const ChildStats& u)
{
    const unsigned long long a = (unsigned long long)u.sum_steps, b = (unsigned long long)u.prefix_steps;
    return make_uint4(u.first_solved, (uint32_t)a, (uint32_t)b,
                      (uint32_t)u.max_steps | ((uint32_t)(a >> 32) << 8) | ((uint32_t)(b >> 32) << 20));
}
// merge record r (the next unit of the child, ascending rollouts) into st
__device__ __forceinline__ void merge_unit_rec(ChildStats& st, const uint4& r)
{
    const long long sum = (long long)(((unsigned long long)((r.w >> 8) & 0xFFFu) << 32) | r.y);
    const long long pre = (long long)(((unsigned long long)(r.w >> 20) << 32) | r.z);
    st.max_steps = max(st.max_steps, (int)(r.w & 0xFFu));
    st.sum_steps += sum;
    if (st.first_solved == NONE_U32) {
        st.prefix_steps += pre;                     // == sum when the unit has no solved rollout
        st.first_solved = r.x;
    }
}
// warp sum of values below 2^52 with four REDUX instead of ten 64-bit shuffle steps
__device__ __forceinline__ long long warp_sum_redux(long long v)
{
    const uint32_t lo = (uint32_t)v & 0x3FFFFFFu, hi = (uint32_t)((unsigned long long)v >> 26);
    return ((long long)__reduce_add_sync(FULLMASK, hi) << 26) + (long long)__reduce_add_sync(FULLMASK, lo);
}

// rank of entry e among the live entries = its index in the reference's current tile list
template <int EW>
__device__ __forceinline__ int live_rank(const uint32_t (&alive)[EW], int e)
{
    int r = 0;
#pragma unroll
    for (int q = 0; q < EW; ++q) {
        const int lo = e - 32 * q;
        const uint32_t below = lo <= 0 ? 0u : (lo >= 32 ? FULLMASK : ((1u << lo) - 1u));
        r += __popc(alive[q] & below);
    }
    return r;
}

// first set bit of a multi-word mask and the run of set bits that starts there
template <int W>
__device__ __forceinline__ void first_run(const uint32_t (&cand)[W], int& col, int& run)
{
    col = 32 * W;
    run = 0;
    bool found = false, extending = false;
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const uint32_t xq = cand[q];
        if (!found) {
            if (xq) {
                found = true;
                const int c0 = ctz32(xq);
                col = 32 * q + c0;
                const int ones = ctz32(~(xq >> c0));          // 32 when every bit from c0 up is set
                run = min(ones, 32 - c0);
                extending = (c0 + run == 32);
            }
        } else if (extending) {
            const int ones = ctz32(~xq);
            run += ones;
            extending = ones == 32;
        }
    }
}

// What a playout needs of the search, copied once per block into shared memory: the playout
// functions are out of line, and a reference to the kernel's parameter block would force the
// whole block into local memory and every pointer in it through generic addressing.
struct PlayParams {
    const ProblemDesc* desc;
    const ProbState* pstate;
    const uint32_t* planes;
    const uint32_t* wmask;
    const uint32_t* hmask;
    const uint4* pairs;
    const uint4* tile;
    const uint8_t* ulist;
    uint4* urec;
    uint4* cacc;
    int es, n_chunks, n_local, sim_begin, urec_stride;
    uint32_t seed;
};

// One work unit in the skyline state: unit `uidx` of problem `prob` = chunks [g * group, ...) of
// the child with rank uidx / units_per_child among the legal entries.  Writes ONE record for the
// unit (the merge of its chunks, same rule as the merge over ranks).
template <int NB, int W, int EW>
__device__ __forceinline__ void sky_play_unit(const PlayParams& p, LaneTables<NB, W, EW>& tb,
                                              const uint8_t* sel8, int prob, int uidx, int& cached_prob)
{
    constexpr int WROWS = LaneTables<NB, W, EW>::WROWS, HROWS = LaneTables<NB, W, EW>::HROWS,
                  NE = LaneTables<NB, W, EW>::NE;
    const int lane = lane_id();
    const ProblemDesc d = G(p.desc)[prob];
    // state written inside this launch: through L2 (three words of the problem's state record)
    const uint4* sp = reinterpret_cast<const uint4*>(G(p.pstate) + prob);
    const uint4 q0 = __ldcg(sp), q1 = __ldcg(sp + 1), al = __ldcg(sp + 3);
    const int depth = (int)q0.x;
    const int4 lfb = make_int4((int)q0.z, (int)q0.w, (int)q1.x, (int)q0.y);   // (hmin, col, run, live entries)
    const int2 ug = make_int2((int)q1.z, (int)q1.w);      // (chunks per unit, units per child)
    const int ci = (int)udiv_small((unsigned)uidx, (unsigned)ug.y), g = uidx - ci * ug.y;
    const int entry = (int)__ldcg(G(p.ulist) + (size_t)prob * p.es + ci);
    LaneBoard<NB, W, EW> s;
    const uint32_t* gp = G(p.planes) + (size_t)prob * LANES_PLANE_STRIDE;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int q = 0; q < W; ++q) s.pl[b][q] = __ldcg(gp + b * 4 + q);

    if (prob != cached_prob) {                        // the problem's tables: static since the reset
        __syncwarp();
        const uint32_t* gw = G(p.wmask) + (size_t)prob * LANES_TABLE_ROWS * LANES_MASK_WORDS;
        const uint32_t* gh = G(p.hmask) + (size_t)prob * LANES_TABLE_ROWS * LANES_MASK_WORDS;
        for (int i = lane; i < WROWS * EW; i += 32) {
            const int xx = i / EW, q = i % EW;
            tb.wmask[i] = (xx <= d.cols) ? gw[xx * LANES_MASK_WORDS + q] : 0u;
        }
        for (int i = lane; i < HROWS * EW; i += 32) {
            const int y = i / EW - ((1 << NB) - 1 - d.rows), q = i % EW;
            tb.hmask[i] = (y >= 0 && y <= d.rows) ? gh[y * LANES_MASK_WORDS + q] : 0u;
        }
        const uint4* gg = G(p.pairs) + (size_t)prob * p.es;
        const uint4* gt = G(p.tile) + (size_t)prob * p.es;
        for (int i = lane; i < NE; i += 32) {
            const bool in = i < d.n_entries;
            const uint4 pr = in ? gg[i] : make_uint4(0u, 0u, 0u, 0u);
            const uint32_t pair[4] = {pr.x, pr.y, pr.z, pr.w};
            tb.entries[i].set(in ? gt[i] : make_uint4(0u, 0u, 0u, 0u), pair);
        }
        __syncwarp();
        cached_prob = prob;
    }

    // the child: parent + entry placed at the parent's first free cell, pair removed
    const uint32_t alw[4] = {al.x, al.y, al.z, al.w};
#pragma unroll
    for (int q = 0; q < EW; ++q) s.alive[q] = alw[q];
    const uint32_t tag = ((uint32_t)depth << 16) | (uint32_t)live_rank<EW>(s.alive, entry);
    const uint4 et = G(p.tile)[(size_t)prob * p.es + entry];
    uint32_t foot[W];
#pragma unroll
    for (int q = 0; q < W; ++q) foot[q] = word_range(lfb.y, lfb.y + (int)et.y, q);
    lanes_place(s, foot, lfb.x, lfb.x + (int)et.x);
#pragma unroll
    for (int q = 0; q < EW; ++q) s.alive[q] &= ~tb.entries[entry].pair(q);
    const int n_alive = lfb.w - 2;

    ChildStats u;
    u.max_steps = 0;
    u.first_solved = NONE_U32;
    u.sum_steps = 0;
    u.prefix_steps = 0;
    const int chunk_begin = g * ug.x, chunk_end = min(chunk_begin + ug.x, p.n_chunks);
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
        u.max_steps = max(u.max_steps, mx);
        u.sum_steps += sum;
        if (u.first_solved == NONE_U32) {
            if (solved_mask) {
                u.first_solved = (uint32_t)(p.sim_begin + first + first_lane);
                u.prefix_steps += __reduce_add_sync(FULLMASK, (lane <= first_lane) ? my_steps : 0);
            } else {
                u.prefix_steps += sum;
            }
        }
    }
    if (lane == 0) {
        G(p.urec)[(size_t)prob * p.urec_stride + uidx] = pack_unit_rec(u);
        // ... and folded into the child's accumulator (ResidentParams::cacc)
        uint4* acc = G(p.cacc) + (size_t)prob * p.es + ci;
        atomicMax(&acc->x, (uint32_t)u.max_steps);
        if (u.first_solved != NONE_U32) atomicMin(&acc->y, u.first_solved);
        atomicAdd(reinterpret_cast<unsigned long long*>(&acc->z), (unsigned long long)u.sum_steps);
    }
}

// ---- root-parallel: per-child statistics of this rank to every rank, then wait for all -----
// One warp, after the last unit of (prob, depth) of THIS rank has been played; `cs[j]` = this
// rank's statistics of the child of rank lane + 32 j, replaced by the merge over all ranks.
// Buffers are indexed by (run & 1, depth & 1): a rank can only be two exchanges ahead of a peer
// that has not read yet when it has itself passed the exchange in between, which needs that
// peer's arrival there.  Flags only grow (run << 8 | depth + 1), so they are never cleared.
// Out of line and on plain arguments: it is the rare path (partitioned searches only) and must
// not cost the common advance code size or registers.
struct ExchangeArgs {
    const ResidentParams::PeerTable* peers;
    int* abort_flag;
    int rank, n_ranks, n_problems, es, run_epoch;
};
template <int NJ>
static __device__ __noinline__ bool exchange_child_stats(ExchangeArgs a, int prob, int depth, int k,
                                                         ChildStats (&cs)[NJ])
{
    const int lane = lane_id();
    const ResidentParams::PeerTable* peers = G(a.peers);
    const int buf = ((a.run_epoch & 1) << 1) | (depth & 1);
    const size_t buf_off = (size_t)buf * a.n_ranks * a.n_problems * a.es;
    const size_t mine = buf_off + ((size_t)a.rank * a.n_problems + prob) * a.es;
    for (int r = 0; r < a.n_ranks; ++r) {
        ChildStats* dst = G(peers->gathered[r]) + mine;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
            if (lane + 32 * j < k) dst[lane + 32 * j] = cs[j];
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");           // every lane: its stores before the flag
    __syncwarp();
    const int stamp = (a.run_epoch << 8) | (depth + 1);
    if (lane < a.n_ranks)
        st_release_sys_s32(G(peers->arrived[lane]) + (size_t)prob * RS_MAX_RANKS + a.rank, stamp);
    bool ok = true;
    if (lane < a.n_ranks) {
        const int* flag = G(peers->arrived[a.rank]) + (size_t)prob * RS_MAX_RANKS + lane;
        const unsigned long long t_begin = global_timer_ns();
        unsigned ns = 32;
        while (ld_relaxed_sys_s32(flag) < stamp) {
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
            if (ld_volatile(G(a.abort_flag)) || global_timer_ns() - t_begin > 4000000000ull) { ok = false; break; }
        }
    }
    ok = __all_sync(FULLMASK, ok);
    __syncwarp();
    if (!ok) return false;
    // merge over the ranks (ascending contiguous rollout ranges), read through L2
    const ChildStats* gathered = G(peers->gathered[a.rank]) + buf_off;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int ci = lane + 32 * j;
        if (ci >= k) continue;
        ChildStats m;
        m.max_steps = 0; m.first_solved = NONE_U32; m.sum_steps = 0; m.prefix_steps = 0;
        for (int g = 0; g < a.n_ranks; ++g) {
            const ChildStats* src = gathered + ((size_t)g * a.n_problems + prob) * a.es + ci;
            const int2 w0 = __ldcg(reinterpret_cast<const int2*>(src));
            const long long sum = __ldcg(&src->sum_steps), pre = __ldcg(&src->prefix_steps);
            m.max_steps = max(m.max_steps, w0.x);
            m.sum_steps += sum;
            if (m.first_solved == NONE_U32) {
                m.prefix_steps += pre;                 // == sum when the rank has no solved rollout
                m.first_solved = (uint32_t)w0.y;
            }
        }
        cs[j] = m;
    }
    return true;
}

// =========================================================== advance (server warps)
// The advance of a problem runs on dedicated warps, not on the warp that played the last unit:
// that warp would execute ~1500 instructions of cold code under the register cap and the calling
// convention of the playout functions, among 31 warps that keep the integer pipes busy --
// measured 60 000 cycles per advance, on every problem's critical path.  The server warps are the
// blocks on the first n_server_sms SMs to receive a block of the launch (elected by arrival at
// kernel start, see the kernel; the grid is one wave, so every SM gets its share of blocks):
// those SMs run nothing else, so the advance code stays in their instruction caches and issues
// without competition.  The
// code is the same for every shape class: the board is the eight bit planes of four words held ONE WORD PER
// LANE (lane = plane * 4 + word), children are one per lane, and the width / height table rows
// come straight from the problem's global tables.

// leftmost column of minimum height of the board held one plane word per lane; returns the
// minimum as the planes store it, inv = 255 - hmin; cand = the columns of that height.
// The plane words go through 128 bytes of the warp's shared memory (one store, broadcast
// loads) and the narrowing is branch-free: shuffles inside data-dependent control flow compile
// to convergence sequences that cost more than the search itself.
__device__ __forceinline__ int board_min(uint32_t pw, int wpr, uint32_t (&cand)[4], uint32_t* scratch)
{
    __syncwarp();
    scratch[lane_id()] = pw;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) cand[q] = q < wpr ? FULLMASK : 0u;
    int inv = 0;
#pragma unroll
    for (int b = 7; b >= 0; --b) {
        const uint4 pl = *reinterpret_cast<const uint4*>(scratch + b * 4);
        const uint32_t z0 = cand[0] & pl.x, z1 = cand[1] & pl.y, z2 = cand[2] & pl.z, z3 = cand[3] & pl.w;
        const bool any = (z0 | z1 | z2 | z3) != 0u;
        cand[0] = any ? z0 : cand[0];
        cand[1] = any ? z1 : cand[1];
        cand[2] = any ? z2 : cand[2];
        cand[3] = any ? z3 : cand[3];
        inv |= any ? (1 << b) : 0;
    }
    return inv;
}

// raise the columns [col, col + w) from inverted height `inv` to inv - h (lane = plane * 4 + word)
__device__ __forceinline__ uint32_t board_place(uint32_t pw, int inv, int h, int col, int w)
{
    const int b = lane_id() >> 2, q = lane_id() & 3;
    const uint32_t diff = (uint32_t)(inv ^ (inv - h));
    return ((diff >> b) & 1u) ? (pw ^ word_range(col, col + w, q)) : pw;
}

// the winning rollout once more, recording its moves (MCTS.py:82-93 splices them into the
// solution order).  Every lane computes the same thing; lane 0 writes.  Returns the steps.
// gw / gh / gt / gpr: the problem's width and height table rows, tile records and pair masks
static __device__ __noinline__ int server_replay(const uint4* gw, const uint4* gh, const uint4* gt,
                                                 const uint4* gpr, int rows, int cols, uint32_t seed,
                                                 uint32_t stream, uint32_t pw, uint4 alive4, int n_alive,
                                                 uint32_t tag, uint32_t sim, uint8_t* moves, bool& solved,
                                                 uint32_t* scratch)
{
    gw = G(gw); gh = G(gh); gt = G(gt); gpr = G(gpr); moves = G(moves);
    const int wpr = (cols + 31) >> 5;
    uint32_t alive[4] = {alive4.x, alive4.y, alive4.z, alive4.w};
    int steps = 0;
    solved = false;
    for (uint32_t block = 0;; ++block) {
        const Philox4 rnd = rollout_stream_block(seed, stream, tag, sim, block);
#pragma unroll 1
        for (int w4 = 0; w4 < 4; ++w4) {
            if (n_alive == 0) { solved = true; return steps; }           // MCTS.py:167-169
            uint32_t cand[4];
            const int inv = board_min(pw, wpr, cand, scratch);
            int col, run;
            first_run<4>(cand, col, run);
            const uint4 wm = gw[run], hm = gh[rows - (255 - inv)];
            const uint32_t legal[4] = {alive[0] & wm.x & hm.x, alive[1] & wm.y & hm.y,
                                       alive[2] & wm.z & hm.z, alive[3] & wm.w & hm.w};
            const int k = __popc(legal[0]) + __popc(legal[1]) + __popc(legal[2]) + __popc(legal[3]);
            if (k == 0) return steps;                                     // MCTS.py:171-172
            const uint32_t u = w4 == 0 ? rnd.x : (w4 == 1 ? rnd.y : (w4 == 2 ? rnd.z : rnd.w));
            int rem = (int)__umulhi(u, (uint32_t)k);                     // MCTS.py:174
            int e = -1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = __popc(legal[q]);
                if (e < 0) {
                    if (rem < c) {
                        uint32_t m = legal[q];
                        for (; rem > 0; --rem) m &= m - 1u;
                        e = 32 * q + __ffs(m) - 1;
                    } else {
                        rem -= c;
                    }
                }
            }
            const uint4 t = gt[e];                                        // (h, w, ...)
            const uint4 pr = gpr[e];
            pw = board_place(pw, inv, (int)t.x, col, (int)t.y);
            alive[0] &= ~pr.x; alive[1] &= ~pr.y; alive[2] &= ~pr.z; alive[3] &= ~pr.w;
            if (lane_id() == 0) { moves[2 * steps] = (uint8_t)t.x; moves[2 * steps + 1] = (uint8_t)t.y; }
            n_alive -= 2;
            steps += 1;
        }
    }
}

// the arrays the host reads, once, when a problem is finished
__device__ __forceinline__ void finish_problem(const SearchParams& p, int prob, const ProbState& st)
{
    if (lane_id() == 0) {
        store_state(p.pstate + prob, st);
        p.status[prob] = st.status;
        p.depth[prob] = st.depth;
        p.n_order[prob] = st.n_order;
        p.n_tiles_placed[prob] = st.n_tiles_placed;
        p.rollouts[prob] = st.rollouts;
        p.rollout_placements[prob] = st.rollout_placements;
        p.rollouts_executed[prob] = st.rollouts_executed;
        p.placements_executed[prob] = st.placements_executed;
    }
}

// ---- advance one problem by one depth in the skyline state ----------------------------------
// One server warp: merge the unit records per child (lane = child rank among the legal entries),
// stop at the first child with a solved rollout, else commit the first maximum (MCTS.py:95-119,
// get_max_index :20-29); place the tile on the skyline, drop its pair from the live set, find the
// next first free cell and the legal entries by table lookup, store the state and publish the
// units of the next depth.  The state is in registers, the same in every lane.
// phase[] receives the cycles of (merge, commit + expand, publish).
template <int NJ>
__device__ __forceinline__ int server_advance_core(const SearchParams& p, const ResidentParams& x, int prob,
                                                   UnitSink sink, long long& c_exchange,
                                                   long long (&phase)[3], uint32_t* scratch)
{
    const int lane = lane_id();
    const long long t_begin = clock64();
    ProbState st;
    load_state_cg(st, p.pstate + prob);
    const ProblemDesc d = p.desc[prob];
    uint32_t pw = __ldcg(p.planes + (size_t)prob * LANES_PLANE_STRIDE + lane);
    const int k0 = st.k, depth = st.depth, gpc = st.gpc;

    // ---- this rank's statistics per child: merge of the child's unit records -------------
    // (every lane also fetches the tile and pair mask of its own child, so that the winner's are
    // one shuffle away instead of one more trip to L2 after the decision)
    ChildStats cs[NJ];
    int ent[NJ];
    uint32_t c_tile[NJ];
    uint4 c_pair[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int ci = lane + 32 * j;
        cs[j].max_steps = 0; cs[j].first_solved = NONE_U32; cs[j].sum_steps = 0; cs[j].prefix_steps = 0;
        ent[j] = 0;
        c_tile[j] = 0u;
        c_pair[j] = make_uint4(0u, 0u, 0u, 0u);
        if (ci < k0) {
            ent[j] = (int)__ldcg(x.ulist + (size_t)prob * p.es + ci);
            const uint4 t4 = p.tile[(size_t)prob * p.es + ent[j]];
            c_tile[j] = t4.x | (t4.y << 8);
            c_pair[j] = reinterpret_cast<const uint4*>(p.pairs)[(size_t)prob * p.es + ent[j]];
            // the child's statistics: one accumulator record (every unit folded itself in); the
            // placements up to a solved rollout need the unit records in order -- only for a child
            // that has one, i.e. at most once per problem
            const uint4 a = __ldcg(x.cacc + (size_t)prob * p.es + ci);
            cs[j].max_steps = (int)a.x;
            cs[j].first_solved = a.y;
            cs[j].sum_steps = (long long)((unsigned long long)a.z | ((unsigned long long)a.w << 32));
            cs[j].prefix_steps = cs[j].sum_steps;
            if (a.y != NONE_U32) {
                const uint4* rec = x.urec + (size_t)prob * x.urec_stride + (size_t)ci * gpc;
                ChildStats t;
                t.max_steps = 0; t.first_solved = NONE_U32; t.sum_steps = 0; t.prefix_steps = 0;
                for (int g = 0; g < gpc; ++g) merge_unit_rec(t, __ldcg(rec + g));
                cs[j].prefix_steps = t.prefix_steps;
            }
        }
    }
    if (p.n_ranks > 1) {
        const long long c0 = clock64();
        ExchangeArgs ea;
        ea.peers = x.peers; ea.abort_flag = &x.q->abort; ea.rank = p.rank; ea.n_ranks = p.n_ranks;
        ea.n_problems = p.n_problems; ea.es = p.es; ea.run_epoch = x.run_epoch;
        ChildStats tmp[NJ];                                  // (its address is taken: keep cs in registers)
#pragma unroll
        for (int j = 0; j < NJ; ++j) tmp[j] = cs[j];
        const bool ok = exchange_child_stats<NJ>(ea, prob, depth, k0, tmp);
#pragma unroll
        for (int j = 0; j < NJ; ++j) cs[j] = tmp[j];
        c_exchange += clock64() - c0;
        if (!ok) {
            if (lane == 0) st_volatile(&x.q->abort, 1);
            return ST_DEAD;
        }
    }
    const long long t_merged = clock64();

    // ---- what the reference would have counted (MCTS.py:60,76,134-140) -------------------
    int first_ci = -1;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (32 * j < k0) {                                      // warp-uniform
            const uint32_t m = __ballot_sync(FULLMASK, lane + 32 * j < k0 && cs[j].first_solved != NONE_U32);
            if (first_ci < 0 && m) first_ci = 32 * j + __ffs(m) - 1;
        }
    }
    long long roll = 0, plc = 0, plc_exec = 0;
    int best_score = -1, best_ci = -1;
    int* scores = p.child_scores + ((size_t)prob * (p.es / 2) + depth) * p.es;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        if (32 * j < k0) {                                      // warp-uniform
            const int ci = lane + 32 * j;
            const bool is_child = ci < k0;
            int score = -1;
            if (is_child) {
                plc_exec += cs[j].sum_steps;
                if (first_ci < 0 || ci < first_ci) {
                    roll += p.n_sim;
                    plc += cs[j].sum_steps;
                    score = p.strategy == BP_MAX_DEPTH ? cs[j].max_steps : (int)cs[j].sum_steps;
                } else if (ci == first_ci) {
                    roll += (long long)cs[j].first_solved + 1;
                    plc += cs[j].prefix_steps;
                    score = -2;
                }
                scores[live_rank<4>(st.alive, ent[j])] = score;
            }
            const int key_score = (is_child && score >= 0) ? score : -1;
            const int mx = __reduce_max_sync(FULLMASK, key_score);
            if (mx > best_score) {                               // strict: earlier j wins ties
                const uint32_t who = __ballot_sync(FULLMASK, key_score == mx);
                best_score = mx;
                best_ci = 32 * j + __ffs(who) - 1;
            }
        }
    }
    roll = warp_sum_redux(roll);
    plc = warp_sum_redux(plc);
    st.placements_executed += (unsigned long long)warp_sum_redux(plc_exec);
    st.rollouts += (unsigned long long)roll;
    st.rollout_placements += (unsigned long long)plc;
    st.n_tiles_placed += (unsigned long long)plc;

    const int take_ci = first_ci >= 0 ? first_ci : best_ci;
    int take = ent[0];
    uint32_t v = c_tile[0];
    uint4 pr = c_pair[0];
#pragma unroll
    for (int j = 1; j < NJ; ++j) {
        const bool sel = (take_ci >> 5) == j;
        take = sel ? ent[j] : take;
        v = sel ? c_tile[j] : v;
        pr = sel ? c_pair[j] : pr;
    }
    take = __shfl_sync(FULLMASK, take, take_ci & 31);
    v = __shfl_sync(FULLMASK, v, take_ci & 31);
    pr.x = __shfl_sync(FULLMASK, pr.x, take_ci & 31);
    pr.y = __shfl_sync(FULLMASK, pr.y, take_ci & 31);
    pr.z = __shfl_sync(FULLMASK, pr.z, take_ci & 31);
    pr.w = __shfl_sync(FULLMASK, pr.w, take_ci & 31);
    const int take_rank = live_rank<4>(st.alive, take);
    const int th = (int)(v & 0xFFu), tw = (int)(v >> 8);
    uint8_t* order = p.order + (size_t)prob * (p.es / 2 + 2) * 2;
    const int prev = st.prev_tile;

    pw = board_place(pw, 255 - st.hmin, th, st.col, tw);
    st.alive[0] &= ~pr.x; st.alive[1] &= ~pr.y; st.alive[2] &= ~pr.z; st.alive[3] &= ~pr.w;

    if (first_ci >= 0) {
        // ---- solved inside a rollout (MCTS.py:77-93) -----------------------------------
        // entries after the solving child were never attempted (:58-60)
        st.n_tiles_placed -= (unsigned long long)(st.n_alive - take_rank - 1);
        if (lane == 0) {
            int n = st.n_order;
            if (prev) { order[2 * n] = prev & 0xFF; order[2 * n + 1] = prev >> 8; n++; }
            order[2 * n] = v & 0xFF; order[2 * n + 1] = v >> 8;
        }
        st.n_order += (prev ? 2 : 1);
        uint32_t sim = NONE_U32;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const uint32_t cand = __shfl_sync(FULLMASK, cs[j].first_solved, first_ci & 31);
            sim = ((first_ci >> 5) == j) ? cand : sim;
        }
        bool solved;
        const uint32_t tag = ((uint32_t)depth << 16) | (uint32_t)take_rank;
        const int steps = server_replay(
            reinterpret_cast<const uint4*>(p.wmask) + (size_t)prob * LANES_TABLE_ROWS,
            reinterpret_cast<const uint4*>(p.hmask) + (size_t)prob * LANES_TABLE_ROWS,
            p.tile + (size_t)prob * p.es, reinterpret_cast<const uint4*>(p.pairs) + (size_t)prob * p.es,
            d.rows, d.cols, p.seed, d.stream, pw,
            make_uint4(st.alive[0], st.alive[1], st.alive[2], st.alive[3]), st.n_alive - 2, tag, sim,
            order + 2 * st.n_order, solved, scratch);
        st.n_order += steps;
        st.status = solved ? ST_SOLVED : ST_DEAD;                // solved is always true here
        finish_problem(p, prob, st);
        phase[0] += t_merged - t_begin;
        phase[1] += clock64() - t_merged;
        return st.status;
    }

    // ---- commit the best child (MCTS.py:105-119), expand the new state -----------------
    if (lane == 0) {
        if (prev) { order[2 * st.n_order] = prev & 0xFF; order[2 * st.n_order + 1] = prev >> 8; }
        p.path[(size_t)prob * (p.es / 2) + depth] = take_rank;
    }
    st.n_order += prev ? 1 : 0;
    st.prev_tile = (int)v;
    st.depth = depth + 1;
    st.n_alive -= 2;
    st.status = ST_ACTIVE;
    st.hmin = -1; st.col = -1; st.run = 0; st.k = 0;
    st.legal[0] = st.legal[1] = st.legal[2] = st.legal[3] = 0u;
    if (st.n_alive == 0) {
        st.status = ST_SOLVED;                               // MCTS.py:121-123
    } else {
        uint32_t cand[4];
        const int inv = board_min(pw, (d.cols + 31) >> 5, cand, scratch);
        if (255 - inv >= d.rows) {
            st.status = ST_DEAD;                             // full board, tiles left
        } else {
            st.hmin = 255 - inv;
            first_run<4>(cand, st.col, st.run);
            const uint4 wm = (reinterpret_cast<const uint4*>(p.wmask) + (size_t)prob * LANES_TABLE_ROWS)[st.run];
            const uint4 hm = (reinterpret_cast<const uint4*>(p.hmask) + (size_t)prob * LANES_TABLE_ROWS)[d.rows - st.hmin];
            st.legal[0] = st.alive[0] & wm.x & hm.x;
            st.legal[1] = st.alive[1] & wm.y & hm.y;
            st.legal[2] = st.alive[2] & wm.z & hm.z;
            st.legal[3] = st.alive[3] & wm.w & hm.w;
            st.k = __popc(st.legal[0]) + __popc(st.legal[1]) + __popc(st.legal[2]) + __popc(st.legal[3]);
            if (st.k == 0) st.status = ST_DEAD;              // MCTS.py:101-103
            st.n_tiles_placed += (unsigned long long)st.n_alive;          // one attempt per entry (:60)
            st.rollouts_executed += (unsigned long long)st.k * p.n_local;
        }
    }
    const long long t_committed = clock64();
    if (st.status == ST_ACTIVE) {
        p.planes[(size_t)prob * LANES_PLANE_STRIDE + lane] = pw;
        if (sink.in_flight < 0)                                  // resident: the problems still searching
            sink.in_flight = __shfl_sync(FULLMASK, ld_volatile(&x.q->in_flight), 0);
        publish_units(p, x, prob, st, sink);
    } else {
        finish_problem(p, prob, st);
    }
    phase[0] += t_merged - t_begin;
    phase[1] += t_committed - t_merged;
    phase[2] += clock64() - t_committed;
    return st.status;
}

// Out of line as well: every class's playout gets its own register allocation (the loop of
// lanes_rollout compiles as it does in the per-class kernel) instead of one allocation over
// all the classes' loops inlined into the unit loop, which cost 14 % more instructions.
template <int NB, int W, int EW>
__device__ __noinline__ void play_unit_resident(const PlayParams* pp, unsigned char* tb,
                                                const uint8_t* sel8, int prob, int uidx,
                                                int& cached_prob)
{
    sky_play_unit<NB, W, EW>(*pp, *reinterpret_cast<LaneTables<NB, W, EW>*>(tb), sel8, prob, uidx,
                             cached_prob);
}

// shared memory of one warp: the lane tables of the largest class the kernel variant plays
template <int MAXW, int MAXEW>
struct alignas(16) ResidentWarpSmem {
    unsigned char tb[sizeof(LaneTables<8, MAXW, MAXEW>)];
};

// resident blocks per SM a variant is compiled for (register cap 64 / 80 / 96 per thread)
#ifndef BP_RESIDENT_BLOCKS_22
#define BP_RESIDENT_BLOCKS_22 8
#endif
constexpr int resident_blocks(int maxw, int maxew) { return (maxw <= 2 && maxew <= 2) ? BP_RESIDENT_BLOCKS_22 : ((maxw <= 2 || maxew <= 2) ? 6 : 5); }

// ------------------------------------------------------- resident search kernel
// MAXW / MAXEW: the widest plane (words) and entry mask (words) among the problems of the
// search; the host launches the smallest variant that covers the batch, because the variant's
// register cap (and with it the resident warps per SM) follows its widest playout.
// A problem's class: NB = 5 / 6 / 8 height planes (rows <= 31 / 63 / 128), W = 1..4 words per
// plane, EW = 2 / 4 words of entry mask; the advance runs on the warp-per-board state
// (RHO rows per lane = 1 / 2 / 4 by plane count, W rounded up to 1 / 2 / 4).
template <int MAXW, int MAXEW>
__global__ void __launch_bounds__(128, resident_blocks(MAXW, MAXEW))
fs_search_resident_kernel(SearchParams p, ResidentParams x)
{
    __shared__ ResidentWarpSmem<MAXW, MAXEW> s_warp[4];
    __shared__ uint8_t s_sel8[256 * 8];
    __shared__ PlayParams s_pp;
    fill_select_table(s_sel8);
    if (threadIdx.x == 0) {
        s_pp.desc = p.desc; s_pp.pstate = p.pstate;
        s_pp.planes = p.planes; s_pp.wmask = p.wmask; s_pp.hmask = p.hmask;
        s_pp.pairs = reinterpret_cast<const uint4*>(p.pairs); s_pp.tile = p.tile;
        s_pp.ulist = x.ulist; s_pp.urec = x.urec; s_pp.cacc = x.cacc;
        s_pp.es = p.es; s_pp.n_chunks = p.n_chunks; s_pp.n_local = p.n_local;
        s_pp.sim_begin = p.sim_begin; s_pp.urec_stride = x.urec_stride; s_pp.seed = p.seed;
    }
    __syncthreads();
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    UnitQueue* q = x.q;

    const long long c_begin = clock64();
    long long c_wait = 0, c_play = 0, c_adv = 0, c_exch = 0;
    long long phase[3] = {0, 0, 0};
    unsigned n_played = 0, n_adv = 0, n_polls = 0;
    const unsigned n_roll_warps = (unsigned)x.n_roll_warps;
    // longest sleep between two polls of a rollout warp that finds no unit: the warps of the
    // first blocks poll often (they take the units of a thin depth at once), the later blocks'
    // warps sleep longer, so that a few hundred pollers, not several thousand, hammer the queue
    // heads while little work is in flight
    const unsigned poll_cap = x.poll_tier > 0
        ? (unsigned)x.poll_ns << min(4u, (blockIdx.x * 4u + (threadIdx.x >> 5)) / (unsigned)x.poll_tier)
        : (unsigned)x.poll_ns;
    // Role of this block: the first n_server_sms SMs that receive a block of the launch become
    // server SMs (every block placed there advances problems), elected by arrival so that the
    // election does not depend on how %smid is numbered or on what else runs on the device.
    __shared__ int s_is_server, s_home;
    __shared__ unsigned s_ring_off[RS_MAX_QUEUES], s_ring_mask[RS_MAX_QUEUES];
    if (threadIdx.x == 0) {
        // (constant indices: a run-time index into the parameter block would move it to local memory)
#pragma unroll
        for (int k = 0; k < RS_MAX_QUEUES; ++k) { s_ring_off[k] = x.ring_off[k]; s_ring_mask[k] = x.ring_mask[k]; }
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned rank = 0xFFFFFFFEu;                           // an SM id beyond the table plays
        if (smid < 256u) {
            unsigned* slot = &q->sm_rank[smid];
            if (atomicCAS(slot, 0u, 0xFFFFFFFFu) == 0u) {
                rank = atomicAdd(&q->sm_order, 1u) + 1u;
                st_volatile(slot, rank);
            } else {
                while ((rank = ld_volatile(slot)) == 0xFFFFFFFFu) __nanosleep(20);
            }
        }
        s_is_server = rank <= (unsigned)x.n_server_sms ? 1 : 0;
        s_home = rank < 256u ? (int)x.home_tab[rank] : 0;       // the playout this SM keeps in its caches
    }
    __syncthreads();
    const int home = s_home, n_queues = x.n_queues;

    // take a ticket from `head_ctr` and wait for its slot in `ring`; false = the search is over
    auto take = [&](unsigned long long* head_ctr, const unsigned long long* ring, unsigned mask,
                    uint32_t& payload) -> bool {
        int got = 0;
        payload = 0;
        if (lane == 0) {
            const unsigned long long ticket = atomicAdd(head_ctr, 1ull);
            const unsigned long long* sl = ring + (ticket & mask);
            const uint32_t want = (uint32_t)(ticket + 1ull);
            unsigned ns = 32;
            unsigned long long t_begin = 0;
            for (;;) {
                const unsigned long long v = ld_relaxed_u64(sl);
                if ((uint32_t)v == want) { payload = (uint32_t)(v >> 32); got = 1; break; }
                if (ld_volatile(&q->in_flight) <= 0 || ld_volatile(&q->abort)) break;
                ++n_polls;
                __nanosleep(ns);
                if (ns < 256) ns *= 2;
                else {
                    const unsigned long long now = global_timer_ns();
                    if (t_begin == 0) t_begin = now;
                    else if (now - t_begin > 4000000000ull) { st_volatile(&q->abort, 1); break; }
                }
            }
        }
        got = __shfl_sync(FULLMASK, got, 0);
        payload = __shfl_sync(FULLMASK, payload, 0);
        __syncwarp();
        return got != 0;
    };

    // rollout warps: a unit from the home queue while it has any, else from the next queue that
    // has; a warp only takes a ticket where it has just seen head < tail (a lost race leaves it
    // with a ticket for the next unit published there, which it waits for)
    auto take_unit = [&](uint32_t& payload) -> bool {
        int got = 0;
        payload = 0;
        if (lane == 0) {
            unsigned ns = 32;
            unsigned long long t_begin = 0;
            for (;;) {
                int k = -1;
                if (n_queues == 1) {
                    k = 0;
                } else {
                    for (int i = 0; i < n_queues; ++i) {
                        const int c = home + i < n_queues ? home + i : home + i - n_queues;
                        if (ld_relaxed_u64(&q->cq[c].tail) > ld_relaxed_u64(&q->cq[c].head)) { k = c; break; }
                    }
                }
                if (k >= 0) {
                    const unsigned long long ticket = atomicAdd(&q->cq[k].head, 1ull);
                    const unsigned long long* sl = x.slots + s_ring_off[k] + (ticket & s_ring_mask[k]);
                    const uint32_t want = (uint32_t)(ticket + 1ull);
                    for (;;) {
                        const unsigned long long v = ld_relaxed_u64(sl);
                        if ((uint32_t)v == want) { payload = (uint32_t)(v >> 32); got = 1; break; }
                        if (ld_volatile(&q->in_flight) <= 0 || ld_volatile(&q->abort)) break;
                        ++n_polls;
                        __nanosleep(ns);
                        if (ns < poll_cap) ns *= 2;
                        else {
                            const unsigned long long now = global_timer_ns();
                            if (t_begin == 0) t_begin = now;
                            else if (now - t_begin > 4000000000ull) { st_volatile(&q->abort, 1); break; }
                        }
                    }
                    break;
                }
                if (ld_volatile(&q->in_flight) <= 0 || ld_volatile(&q->abort)) break;
                ++n_polls;
                __nanosleep(ns);
                if (ns < poll_cap) ns *= 2;
                else {
                    const unsigned long long now = global_timer_ns();
                    if (t_begin == 0) t_begin = now;
                    else if (now - t_begin > 4000000000ull) { st_volatile(&q->abort, 1); break; }
                }
            }
        }
        got = __shfl_sync(FULLMASK, got, 0);
        payload = __shfl_sync(FULLMASK, payload, 0);
        __syncwarp();
        return got != 0;
    };

    if (s_is_server) {
        // ---- server warps: advance the problems whose depth has been played --------------
        for (;;) {
            const long long c0 = clock64();
            uint32_t prob;
            if (!take(&q->done_head, x.done_slots, x.done_mask, prob)) break;
            const long long c1 = clock64();
            c_wait += c1 - c0;
            uint32_t* scratch = reinterpret_cast<uint32_t*>(s_warp[warp].tb);
            const int n_children = __ldcg(&p.pstate[prob].k);
            // the common case -- at most 32 legal children, one per lane -- is its own compact
            // straight line (the server SMs' instruction caches hold little more than it); more
            // children (the first moves of a large tile list) take the general form
            int status;
            const int qk = p.desc[prob].pad0;                   // the unit queue of the problem's class
            const UnitSink sink = {&q->cq[qk].tail, x.slots + s_ring_off[qk], s_ring_mask[qk], -1,
                                   n_roll_warps, nullptr, nullptr};
            if (n_children <= 32) status = server_advance_core<1>(p, x, (int)prob, sink, c_exch, phase, scratch);
            else status = server_advance_core<4>(p, x, (int)prob, sink, c_exch, phase, scratch);
            if (status != ST_ACTIVE && lane == 0) atomicSub(&q->in_flight, 1);
            c_adv += clock64() - c1;
            n_adv++;
        }
    } else {
        // ---- rollout warps: play work units ------------------------------------------------
        int cached_key = -1;
        for (;;) {
            const long long c0 = clock64();
            uint32_t payload;
            if (!take_unit(payload)) break;
            const int prob = (int)(payload & ((1u << RS_PROBLEM_BITS) - 1u));
            const int uidx = (int)(payload >> RS_PROBLEM_BITS);
            const ProblemDesc d = p.desc[prob];
            const long long c1 = clock64();
            c_wait += c1 - c0;

            // the playout of the problem's class (classes outside the variant do not exist)
            const int nb = d.rows <= 31 ? 5 : (d.rows <= 63 ? 6 : 8);
            const int w = (d.cols + 31) >> 5;
            const int ew = d.n_entries <= 64 ? 2 : 4;
            unsigned char* tb = s_warp[warp].tb;
#define BP_PLAY(NBv, Wv, EWv) play_unit_resident<NBv, Wv, EWv>(&s_pp, tb, s_sel8, prob, uidx, cached_key)
#define BP_PLAY_NB(Wv, EWv)                                                                        \
            do {                                                                                   \
                if (nb == 5) BP_PLAY(5, Wv, EWv);                                                  \
                else if (nb == 6) BP_PLAY(6, Wv, EWv);                                             \
                else BP_PLAY(8, Wv, EWv);                                                          \
            } while (0)
#define BP_PLAY_W(EWv)                                                                             \
            do {                                                                                   \
                if (w == 1) BP_PLAY_NB(1, EWv);                                                    \
                else if (w == 2) BP_PLAY_NB(2, EWv);                                               \
                else if (MAXW > 2 && w == 3) BP_PLAY_NB((MAXW > 2 ? 3 : 1), EWv);                  \
                else if (MAXW > 2) BP_PLAY_NB((MAXW > 2 ? 4 : 1), EWv);                            \
            } while (0)
            if (ew == 2) BP_PLAY_W(2);
            else if (MAXEW > 2) BP_PLAY_W((MAXEW > 2 ? 4 : 2));
#undef BP_PLAY_W
#undef BP_PLAY_NB
#undef BP_PLAY

            // the last unit of this problem's depth hands the problem to the server warps
            if (lane == 0 && atom_add_release_s32(p.pending + prob, -1) == 1) {   // the unit's record first
                const unsigned long long t = atomicAdd(&q->done_tail, 1ull);
                st_release_u64(x.done_slots + (t & x.done_mask),
                               ((unsigned long long)(uint32_t)prob << 32) | (unsigned long long)(uint32_t)(t + 1ull));
            }
            c_play += clock64() - c1;
            n_played++;
        }
    }
    if (lane == 0) {
        atomicAdd(&q->cyc_total, (unsigned long long)(clock64() - c_begin));
        atomicAdd(&q->cyc_wait, (unsigned long long)c_wait);
        atomicAdd(&q->cyc_play, (unsigned long long)c_play);
        atomicAdd(&q->cyc_advance, (unsigned long long)(c_adv - c_exch));
        atomicAdd(&q->cyc_exchange, (unsigned long long)c_exch);
        atomicAdd(&q->n_units_played, (unsigned long long)n_played);
        atomicAdd(&q->n_advances, (unsigned long long)n_adv);
        atomicAdd(&q->n_polls, (unsigned long long)n_polls);
        atomicAdd(&q->cyc_adv_merge, (unsigned long long)phase[0]);
        atomicAdd(&q->cyc_adv_commit, (unsigned long long)phase[1]);
        atomicAdd(&q->cyc_adv_publish, (unsigned long long)phase[2]);
        // the last warp out leaves empty queues behind: tickets handed out but never served
        // are forgotten, the next launch publishes from `head` on
        if (atomicAdd(&q->exited, 1u) == gridDim.x * 4u - 1u) {
            for (int k = 0; k < RS_MAX_QUEUES; ++k) {
                const unsigned long long h = ld_volatile(&q->cq[k].head), t = ld_volatile(&q->cq[k].tail);
                const unsigned long long m = h > t ? h : t;
                st_volatile(&q->cq[k].head, m);
                st_volatile(&q->cq[k].tail, m);
            }
            const unsigned long long dh = ld_volatile(&q->done_head), dt = ld_volatile(&q->done_tail);
            const unsigned long long dm = dh > dt ? dh : dt;
            st_volatile(&q->done_head, dm);
            st_volatile(&q->done_tail, dm);
        }
    }
}

// ---- seed: the state record of every problem from what the reset left, depth 0 published ---
// one warp per problem (of `problems[0 .. n)`, or all of them); the state arrays, legal masks and
// lane tables were written by the reset
template <int UNUSED>                              // (a template so that the header owns one definition)
__global__ void __launch_bounds__(ADV_WARPS * 32)
fs_resident_seed_kernel(SearchParams p, ResidentParams x, UnitSink sink, const int* __restrict__ problems, int n)
{
    const int idx = blockIdx.x * ADV_WARPS + (threadIdx.x >> 5);
    if (idx >= n) return;
    const int prob = problems ? problems[idx] : idx;
    if (p.status[prob] != ST_ACTIVE) return;       // solved / dead before the first move: arrays are final
    const uint4 lg = p.legal[prob], al = p.alive[prob];
    const int4 lfb = p.lfb[prob];
    ProbState st;
    st.depth = 0; st.n_alive = lfb.w; st.hmin = lfb.x; st.col = lfb.y;
    st.run = lfb.z; st.group = 0; st.gpc = 0;
    st.legal[0] = lg.x; st.legal[1] = lg.y; st.legal[2] = lg.z; st.legal[3] = lg.w;
    st.alive[0] = al.x; st.alive[1] = al.y; st.alive[2] = al.z; st.alive[3] = al.w;
    st.k = __popc(lg.x) + __popc(lg.y) + __popc(lg.z) + __popc(lg.w);
    st.n_order = p.n_order[prob]; st.prev_tile = p.prev_tile[prob]; st.status = ST_ACTIVE; st.pad0 = 0;
    st.n_tiles_placed = p.n_tiles_placed[prob];
    st.rollouts = p.rollouts[prob];
    st.rollout_placements = p.rollout_placements[prob];
    st.rollouts_executed = p.rollouts_executed[prob];
    st.placements_executed = p.placements_executed[prob];
    st.pad1 = 0;
    if (st.k == 0) return;
    if (sink.slots == nullptr) {                   // resident search: the unit queue of the problem's class
        if (lane_id() == 0) atomicAdd(&x.q->in_flight, 1);
        const int qk = p.desc[prob].pad0;
        sink.tail = &x.q->cq[qk].tail;
        sink.slots = x.slots + x.ring_off[qk];
        sink.mask = x.ring_mask[qk];
    }
    publish_units(p, x, prob, st, sink);
}

// =================================================== launch-per-depth form on the skyline state
// The same playout and the same advance as the resident kernel, with launches as the barriers:
// per depth and chain one rollout launch over the chain's unit list and one advance launch (one
// warp per problem).  The tables are written once by the reset and the advance is the compact
// skyline advance, so a depth's advance launch is a few microseconds instead of the ~19 of the
// warp-per-board advance that rebuilt every table at every depth.
template <int NB, int W, int EW>
__global__ void __launch_bounds__(128)
fs_rollout_sky_kernel(SearchParams p, ResidentParams x, const unsigned long long* units,
                      const unsigned long long* n_units_ptr, unsigned* counter)
{
    BP_TIMELINE_SCOPE((int)(counter - p.task_counter) * 2);
    __shared__ LaneTables<NB, W, EW> s_tb[4];
    __shared__ uint8_t s_sel8[256 * 8];
    fill_select_table(s_sel8);
    __syncthreads();
    grid_dependency_wait();                // (the table fill overlaps the advance launch before this one)
    // a block that is not needed (the blocks before it hold a warp for every unit) leaves at once:
    // the grid is sized for the units the depth CAN have, and late in a search most problems are
    // finished
    // (list and count are written by the launch before this one, which a programmatic dependent
    // launch overlaps: read through L2, never by the non-coherent path a const __restrict__ load takes)
    const unsigned long long n_units = __ldcg(n_units_ptr);
    if ((unsigned long long)blockIdx.x * 4ull >= n_units) return;
    grid_launch_dependents();
    PlayParams pp;
    pp.desc = p.desc; pp.pstate = p.pstate; pp.planes = p.planes; pp.wmask = p.wmask; pp.hmask = p.hmask;
    pp.pairs = reinterpret_cast<const uint4*>(p.pairs); pp.tile = p.tile; pp.ulist = x.ulist; pp.urec = x.urec; pp.cacc = x.cacc;
    pp.es = p.es; pp.n_chunks = p.n_chunks; pp.n_local = p.n_local; pp.sim_begin = p.sim_begin;
    pp.urec_stride = x.urec_stride; pp.seed = p.seed;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    int cached_prob = -1;
    for (;;) {
        unsigned unit = 0;
        if (lane == 0) unit = atomicAdd(counter, 1u);
        unit = __shfl_sync(FULLMASK, unit, 0);
        if (unit >= n_units) break;
        const uint32_t payload = (uint32_t)(__ldcg(units + unit) >> 32);
        sky_play_unit<NB, W, EW>(pp, s_tb[warp], s_sel8, (int)(payload & ((1u << RS_PROBLEM_BITS) - 1u)),
                                 (int)(payload >> RS_PROBLEM_BITS), cached_prob);
    }
}

// HANDOVER (hybrid form): the last launch-per-depth advance of a chain.  The problems that go on
// are counted in the queue header and their units go to the unit queue of their class instead of
// the chain's list: the resident kernel, launched when every chain has handed over, plays the
// remaining depths.  `n_total` = problems of the whole search (unit size: the chain's share of the
// active problems is taken for the whole batch's).
template <int HANDOVER>
__global__ void __launch_bounds__(ADV_WARPS * 32)
fs_advance_sky_kernel(SearchParams p, ResidentParams x, UnitSink sink, const int* __restrict__ problems, int n,
                      int n_total)
{
    BP_TIMELINE_SCOPE(((int)(sink.active_next - p.child_count) - FS_CHAINS) * 2 + 1);
    grid_launch_dependents();
    grid_dependency_wait();
    __shared__ __align__(16) uint32_t s_scratch[ADV_WARPS][32];
    const int warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * ADV_WARPS + warp;
    if (idx >= n) return;
    const int prob = problems[idx];
    const uint4 q1 = reinterpret_cast<const uint4*>(p.pstate + prob)[1];
    const uint4 q4 = reinterpret_cast<const uint4*>(p.pstate + prob)[4];
    if ((int)q4.z != ST_ACTIVE || p.status[prob] != ST_ACTIVE) return;
    if (HANDOVER) {
        const int qk = p.desc[prob].pad0;
        sink.tail = &x.q->cq[qk].tail;
        sink.slots = x.slots + x.ring_off[qk];
        sink.mask = x.ring_mask[qk];
        sink.in_flight = max(1, (int)(((long long)*sink.active_now * n_total) / n));
        sink.active_now = nullptr;
    }
    long long c_exch = 0, phase[3] = {0, 0, 0};
    int status;
    if ((int)q1.y <= 32) status = server_advance_core<1>(p, x, prob, sink, c_exch, phase, s_scratch[warp]);
    else status = server_advance_core<4>(p, x, prob, sink, c_exch, phase, s_scratch[warp]);
    if (HANDOVER && status == ST_ACTIVE && lane_id() == 0) atomicAdd(&x.q->in_flight, 1);
}

}  // namespace bp
