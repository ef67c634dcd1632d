// The resident form of the flat search: one launch per shape class runs every depth.
#pragma once
#include "search_kernels.cuh"

namespace bp {

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
                    if (waiters >= feed) leave = 1;
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
