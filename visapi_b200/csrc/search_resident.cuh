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

// ---- root-parallel: per-child statistics of this rank to every rank, then wait for all -----
// One warp, after the last unit of (prob, depth) of THIS rank has been played.  Buffers are
// indexed by (run & 1, depth & 1): a rank can only be two exchanges ahead of a peer that has
// not read yet when it has itself passed the exchange in between, which needs that peer's
// arrival there.  Flags only grow (run << 8 | depth + 1), so they are never cleared.
template <int EJ>
__device__ __forceinline__ const ChildStats* exchange_child_stats(const SearchParams& p,
                                                                   const ResidentParams& x, int prob,
                                                                   int depth, bool& ok)
{
    const int lane = lane_id();
    const uint4 lg = __ldcg(p.legal + prob);
    const uint32_t legal[4] = {lg.x, lg.y, lg.z, lg.w};
    const int buf = ((x.run_epoch & 1) << 1) | (depth & 1);
    const size_t buf_off = (size_t)buf * p.n_ranks * p.n_problems * p.es;
    const size_t mine = buf_off + ((size_t)p.rank * p.n_problems + prob) * p.es;
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        if ((legal[j] >> lane) & 1u) {
            const int e = lane + 32 * j;
            const ChildStats st = stats_from_records<true>(p, prob, e);
            for (int r = 0; r < p.n_ranks; ++r) x.gathered[r][mine + e] = st;
        }
    }
    __threadfence_system();
    __syncwarp();
    const int stamp = (x.run_epoch << 8) | (depth + 1);
    if (lane < p.n_ranks) st_release_sys_s32(x.arrived[lane] + (size_t)prob * RS_MAX_RANKS + p.rank, stamp);
    ok = true;
    if (lane < p.n_ranks) {
        const int* flag = x.arrived[p.rank] + (size_t)prob * RS_MAX_RANKS + lane;
        const unsigned long long t_begin = global_timer_ns();
        unsigned ns = 32;
        while (ld_acquire_sys_s32(flag) < stamp) {
            __nanosleep(ns);
            if (ns < 1024) ns *= 2;
            if (ld_volatile(&x.q->abort) || global_timer_ns() - t_begin > 4000000000ull) { ok = false; break; }
        }
    }
    ok = __all_sync(FULLMASK, ok);
    __syncwarp();
    return x.gathered[p.rank] + buf_off;
}

// Out of line on purpose: the resident kernel is compiled under the register cap of its rollout
// loops, and the advance -- which runs once per problem and depth -- keeps its own, larger
// working set (spilling if it must) without costing the loops a register.
template <int RHO, int W, int EJ>
__device__ __noinline__ int advance_resident(const SearchParams& p, const ResidentParams& x, int prob,
                                             uint16_t* cbuf, unsigned n_warps, long long& c_exchange)
{
    const ChildStats* gathered = nullptr;
    if (p.n_ranks > 1) {
        const long long c0 = clock64();
        bool ok;
        gathered = exchange_child_stats<EJ>(p, x, prob, __ldcg(p.depth + prob), ok);
        c_exchange += clock64() - c0;
        if (!ok) {
            if (lane_id() == 0) st_volatile(&x.q->abort, 1);
            return ST_DEAD;
        }
    }
    return advance_problem<RHO, W, EJ, false, true>(p, prob, 0, 0, 0, gathered, cbuf, &x, n_warps);
}

// Out of line as well: every class's playout gets its own register allocation (the loop of
// lanes_rollout compiles as it does in the per-class kernel) instead of one allocation over
// all the classes' loops inlined into the unit loop, which cost 14 % more instructions.
template <int NB, int W, int EW>
__device__ __noinline__ void play_unit_resident(const SearchParams& p, unsigned char* tb,
                                                const uint8_t* sel8, int prob, int entry,
                                                int chunk_begin, int chunk_end, int& cached_key)
{
    lanes_play_unit<true, NB, W, EW>(p, *reinterpret_cast<LaneTables<NB, W, EW>*>(tb), sel8, prob, entry,
                                     chunk_begin, chunk_end, cached_key);
}

// shared memory of one warp: the lane tables of the largest class the kernel variant plays
template <int MAXW, int MAXEW>
struct alignas(16) ResidentWarpSmem {
    unsigned char tb[sizeof(LaneTables<8, MAXW, MAXEW>)];
    uint16_t cbuf[32 * MAXEW];
};

// resident blocks per SM a variant is compiled for (register cap 64 / 80 / 96 per thread)
constexpr int resident_blocks(int maxw, int maxew) { return (maxw <= 2 && maxew <= 2) ? 8 : ((maxw <= 2 || maxew <= 2) ? 6 : 5); }

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
    fill_select_table(s_sel8);
    __syncthreads();
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const unsigned n_warps = gridDim.x * 4u;
    UnitQueue* q = x.q;

    int cached_key = -1;
    const long long c_begin = clock64();
    long long c_wait = 0, c_play = 0, c_adv = 0, c_exch = 0;
    unsigned n_played = 0, n_adv = 0, n_polls = 0;
    for (;;) {
        // ---- take a ticket and wait for its slot ------------------------------------------
        const long long c0 = clock64();
        uint32_t payload = 0;
        int got = 0;
        if (lane == 0) {
            const unsigned long long ticket = atomicAdd(&q->head, 1ull);
            const unsigned long long* sl = x.slots + (ticket & x.cap_mask);
            const uint32_t want = (uint32_t)(ticket + 1ull);
            unsigned ns = 32;
            unsigned long long t_begin = 0;
            for (;;) {
                const unsigned long long v = ld_acquire_u64(sl);
                if ((uint32_t)v == want) { payload = (uint32_t)(v >> 32); got = 1; break; }
                if (ld_volatile(&q->in_flight) <= 0 || ld_volatile(&q->abort)) break;
                ++n_polls;
                __nanosleep(ns);
                if (ns < 512) ns *= 2;
                else {
                    const unsigned long long now = global_timer_ns();
                    if (t_begin == 0) t_begin = now;
                    else if (now - t_begin > 4000000000ull) { st_volatile(&q->abort, 1); break; }
                }
            }
        }
        got = __shfl_sync(FULLMASK, got, 0);
        if (!got) break;
        payload = __shfl_sync(FULLMASK, payload, 0);
        __syncwarp();
        const int prob = (int)(payload & ((1u << RS_PROBLEM_BITS) - 1u));
        const int uidx = (int)(payload >> RS_PROBLEM_BITS);
        const int2 ug = __ldcg(x.ugeom + prob);                    // (chunks per unit, units per child)
        const int ci = uidx / ug.y, g = uidx - ci * ug.y;
        const int entry = (int)__ldcg(x.ulist + (size_t)prob * p.es + ci);
        const int chunk_begin = g * ug.x, chunk_end = min(chunk_begin + ug.x, p.n_chunks);
        const ProblemDesc d = p.desc[prob];
        const long long c1 = clock64();
        c_wait += c1 - c0;

        // ---- play it with the playout of the problem's class ------------------------------
        const int nb = d.rows <= 31 ? 5 : (d.rows <= 63 ? 6 : 8);
        const int w = (d.cols + 31) >> 5;
        const int ew = d.n_entries <= 64 ? 2 : 4;
#define BP_PLAY(NBv, Wv, EWv)                                                                      \
        play_unit_resident<NBv, Wv, EWv>(p, s_warp[warp].tb, s_sel8, prob, entry, chunk_begin,     \
                                         chunk_end, cached_key)
#define BP_PLAY_NB(Wv, EWv)                                                                        \
        do {                                                                                       \
            if (nb == 5) BP_PLAY(5, Wv, EWv);                                                      \
            else if (nb == 6) BP_PLAY(6, Wv, EWv);                                                 \
            else BP_PLAY(8, Wv, EWv);                                                              \
        } while (0)
#define BP_PLAY_W(EWv)                                                                             \
        do {                                                                                       \
            if (w == 1) BP_PLAY_NB(1, EWv);                                                        \
            else if (w == 2) BP_PLAY_NB(2, EWv);                                                   \
            else if (MAXW > 2 && w == 3) BP_PLAY_NB((MAXW > 2 ? 3 : 1), EWv);                      \
            else if (MAXW > 2) BP_PLAY_NB((MAXW > 2 ? 4 : 1), EWv);                                \
        } while (0)
        if (ew == 2) BP_PLAY_W(2);
        else if (MAXEW > 2) BP_PLAY_W((MAXEW > 2 ? 4 : 2));
#undef BP_PLAY_W
#undef BP_PLAY_NB
#undef BP_PLAY

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
            const int rho = nb == 5 ? 1 : (nb == 6 ? 2 : 4);
            int status = ST_DEAD;
            uint16_t* cbuf = s_warp[warp].cbuf;
#define BP_ADV(Rv, Wv, Ev) status = advance_resident<Rv, Wv, Ev>(p, x, prob, cbuf, n_warps, c_exch)
#define BP_ADV_R(Wv, Ev)                                                                           \
            do {                                                                                   \
                if (rho == 1) BP_ADV(1, Wv, Ev);                                                   \
                else if (rho == 2) BP_ADV(2, Wv, Ev);                                              \
                else BP_ADV(4, Wv, Ev);                                                            \
            } while (0)
#define BP_ADV_W(Ev)                                                                               \
            do {                                                                                   \
                if (w == 1) BP_ADV_R(1, Ev);                                                       \
                else if (w == 2) BP_ADV_R(2, Ev);                                                  \
                else if (MAXW > 2) BP_ADV_R((MAXW > 2 ? 4 : 1), Ev);                               \
            } while (0)
            if (ew == 2) BP_ADV_W(2);
            else if (MAXEW > 2) BP_ADV_W((MAXEW > 2 ? 4 : 2));
#undef BP_ADV_W
#undef BP_ADV_R
#undef BP_ADV
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
        atomicAdd(&q->cyc_advance, (unsigned long long)(c_adv - c_exch));
        atomicAdd(&q->cyc_exchange, (unsigned long long)c_exch);
        atomicAdd(&q->n_units_played, (unsigned long long)n_played);
        atomicAdd(&q->n_advances, (unsigned long long)n_adv);
        atomicAdd(&q->n_polls, (unsigned long long)n_polls);
        // the last warp out leaves an empty queue behind: tickets handed out but never served
        // are forgotten, the next launch publishes from `head` on
        __threadfence();
        if (atomicAdd(&q->exited, 1u) == n_warps - 1u) {
            const unsigned long long h = ld_volatile(&q->head), t = ld_volatile(&q->tail);
            const unsigned long long m = h > t ? h : t;
            st_volatile(&q->head, m);
            st_volatile(&q->tail, m);
        }
    }
}

// ---- seed: publish depth 0 of every problem (after a reset that ran in stepped form) -------
// one warp per problem; the state, legal masks and lane tables were written by the reset
template <int EJ>
__global__ void __launch_bounds__(ADV_WARPS * 32)
fs_resident_seed_kernel(SearchParams p, ResidentParams x, unsigned n_warps)
{
    const int prob = blockIdx.x * ADV_WARPS + (threadIdx.x >> 5);
    if (prob >= p.n_problems) return;
    if (p.status[prob] != ST_ACTIVE) return;
    const uint4 lg = p.legal[prob];
    const uint32_t lgw[4] = {lg.x, lg.y, lg.z, lg.w};
    uint32_t legal[EJ];
    int k = 0;
#pragma unroll
    for (int j = 0; j < EJ; ++j) { legal[j] = lgw[j]; k += __popc(legal[j]); }
    if (k == 0) return;
    if (lane_id() == 0) atomicAdd(&x.q->in_flight, 1);
    publish_units<EJ>(p, x, prob, legal, k, x.seed_in_flight, n_warps);
}

}  // namespace bp
