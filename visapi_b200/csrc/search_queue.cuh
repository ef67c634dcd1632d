// Work queue of the resident search kernel (search_resident.cuh): one queue per search, shared by
// every shape class; header, unit slots and the producer side (publish_units).
#pragma once
#include "search_types.cuh"

namespace bp {

// A unit = `group` consecutive 32-rollout chunks of one child of one problem.  The slot is ONE
// 64-bit word -- low half: (uint32_t)(ticket + 1), high half: problem | unit index << 20 -- so a
// unit is published by a single store and taken by a single load: nothing can be torn.  The
// unit index counts (child rank among the legal entries) * units_per_child + unit of the child;
// the geometry of the problem's current depth is in ugeom[problem], the legal entries in
// ulist[problem][rank].
//
// Producers (the warp that advanced a problem) reserve a ticket range with one atomicAdd on
// `tail` and store the slots; consumers take a ticket from `head` and poll their own slot.
// The ring holds every unit that can be outstanding at once (gpc_max bounds the units of a
// child), and a unit is outstanding until it has been PLAYED, so slot (t + capacity) cannot be
// published before the consumer of ticket t is done with its slot.
// Tickets run on across searches: the last warp to leave a launch sets tail = head, so the next
// launch starts with an empty queue and no slot has to be cleared.
constexpr int RS_PROBLEM_BITS = 20;              // problems per resident search < 2^20
constexpr int RS_MAX_UNITS = 1 << 12;            // units of one problem's depth <= 4096
constexpr int RS_MAX_RANKS = 8;                  // root-parallel ranks (one NVSwitch box)

struct UnitQueue {
    // every hot word on its own 128-byte line
    unsigned long long head;     // tickets handed to consumers
    unsigned pad_h[30];
    unsigned long long tail;     // tickets reserved by producers
    unsigned pad_t[30];
    // ---- zeroed by every reset from here on -------------------------------------------------
    int in_flight;               // problems still searching (read by waiting warps)
    int abort;                   // 1: a warp waited too long (watchdog), 2: unit limit exceeded
    unsigned pad_f[30];
    unsigned exited;             // warps that have left the launch
    unsigned pad_e[31];
    // where the warps' time went (SM clock cycles summed over warps, added when a warp leaves)
    unsigned long long cyc_total, cyc_wait, cyc_play, cyc_advance, cyc_exchange;
    unsigned long long n_units_played, n_advances, n_polls, n_units_published;
    unsigned pad_c[14];
};
static_assert(sizeof(UnitQueue) == 640, "queue header layout");
constexpr size_t RS_QUEUE_RESET_OFFSET = 256;     // bytes of the header a reset keeps (head, tail)

// Kernel-parameter block of the resident search (passed by value beside SearchParams).
struct ResidentParams {
    UnitQueue* q;
    unsigned long long* slots;
    unsigned cap_mask;           // ring capacity - 1
    int gpc_max;                 // most units one child may be split into
    int units_per_warp_x4;       // target units in flight per resident warp, times 4
    int seed_in_flight;          // problems assumed in flight while depth 0 is being published
    uint8_t* ulist;              // [P][ES] legal entries of the current depth, in table order
    int2* ugeom;                 // [P] (chunks per unit, units per child) of the current depth
    // root-parallel exchange over peer memory (n_ranks > 1): every rank's buffers, own included
    ChildStats* gathered[RS_MAX_RANKS];   // [4][n_ranks][P][ES], buffer = (run & 1) * 2 + (depth & 1)
    int* arrived[RS_MAX_RANKS];           // [P][RS_MAX_RANKS] last (run << 8 | depth + 1) each rank reached
    int run_epoch;
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* ptr)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* ptr, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys_s32(const int* ptr)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_s32(int* ptr, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- publish the units of a problem's new depth ---------------------------------------------
// Called by a whole warp after the problem's state, legal mask and lane tables are written.
// `in_flight`: problems that share the machine (decides the unit size: whole children while
// there is plenty of work, single chunks when one problem has the GPU to itself).
template <int EJ>
__device__ __forceinline__ void publish_units(const SearchParams& p, const ResidentParams& x, int prob,
                                              const uint32_t (&legal)[EJ], int k, int in_flight,
                                              unsigned n_warps)
{
    const int lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    uint8_t* ul = x.ulist + (size_t)prob * p.es;
    int base = 0;
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        if ((legal[j] >> lane) & 1u) ul[base + __popc(legal[j] & lt)] = (uint8_t)(lane + 32 * j);
        base += __popc(legal[j]);
    }
    // units per child: enough units that the problems in flight offer units_per_warp units to
    // every resident warp, no more than gpc_max (ring capacity) and RS_MAX_UNITS per problem
    const long long want_units = ((long long)n_warps * x.units_per_warp_x4) / (4ll * max(in_flight, 1));
    int gpc = (int)min((long long)p.n_chunks, (want_units + k - 1) / k);
    gpc = max(1, min(gpc, min(x.gpc_max, RS_MAX_UNITS / k)));
    const int group = (p.n_chunks + gpc - 1) / gpc;
    gpc = (p.n_chunks + group - 1) / group;                        // equal units, none empty
    const int n_units = k * gpc;
    if (lane == 0) {
        st_volatile(p.pending + prob, n_units);
        x.ugeom[prob] = make_int2(group, gpc);
    }
    // state, tables, pending and geometry before any slot: every lane fences its own writes,
    // the warp barrier orders them before the slot stores of all lanes
    __threadfence();
    __syncwarp();
    unsigned long long t0 = 0;
    if (lane == 0) {
        t0 = atomicAdd(&x.q->tail, (unsigned long long)n_units);
        atomicAdd(&x.q->n_units_published, (unsigned long long)n_units);
    }
    t0 = __shfl_sync(FULLMASK, t0, 0);
    for (int i = lane; i < n_units; i += 32) {
        const unsigned long long t = t0 + (unsigned long long)i;
        const uint32_t payload = (uint32_t)prob | ((uint32_t)i << RS_PROBLEM_BITS);
        st_relaxed_u64(x.slots + (t & x.cap_mask),
                       ((unsigned long long)payload << 32) | (unsigned long long)(uint32_t)(t + 1ull));
    }
    __syncwarp();
}

}  // namespace bp
