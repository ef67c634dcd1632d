// Work queue of the resident search kernel (search_resident.cuh): one queue per search, shared by
// every shape class; header, unit slots and the producer side (publish_units).
#pragma once
#include "search_types.cuh"

namespace bp {

// A unit = `group` consecutive 32-rollout chunks of one child of one problem.  The slot is ONE
// 64-bit word -- low half: (uint32_t)(ticket + 1), high half: problem | unit index << 20 -- so a
// unit is published by a single store and taken by a single load: nothing can be torn.  The
// unit index counts (child rank among the legal entries) * units_per_child + unit of the child;
// the geometry of the problem's current depth is in its ProbState (group, gpc), the legal
// entries in ulist[problem][rank].
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

constexpr int RS_MAX_QUEUES = 8;     // unit queues of a search: one per shape class (the last one shared)

// one unit queue: tickets handed to consumers / reserved by producers, each on its own line
struct ClassQueue {
    unsigned long long head;
    unsigned pad_h[30];
    unsigned long long tail;
    unsigned pad_t[30];
};

struct UnitQueue {
    // every hot word on its own 128-byte line.  One unit queue per shape class: an SM has a HOME
    // class and takes its units from that queue while there are any, so that its instruction caches
    // hold one class's playout instead of all of them (four mixed playouts cost the resident
    // kernel 16 % -- `no_instruction` was its top stall -- against 4.6 % with one class); a warp
    // that finds its home queue empty takes from the next one.
    ClassQueue cq[RS_MAX_QUEUES];
    unsigned long long done_head;   // the same pair for the queue of problems waiting for their advance
    unsigned pad_dh[30];
    unsigned long long done_tail;
    unsigned pad_dt[30];
    // ---- zeroed by every reset from here on -------------------------------------------------
    int in_flight;               // problems still searching (read by waiting warps)
    int abort;                   // 1: a warp waited too long (watchdog), 2: unit limit exceeded
    unsigned pad_f[30];
    unsigned exited;             // warps that have left the launch
    unsigned sm_order;           // SMs that have received their first block of the launch
    unsigned pad_e[30];
    // arrival rank (1-based) of every SM of the launch, by %smid; 0 = no block yet, ~0 = being
    // ranked.  The SMs ranked 1 .. n_server_sms advance problems, the others play.
    unsigned sm_rank[256];
    // where the warps' time went (SM clock cycles summed over warps, added when a warp leaves)
    unsigned long long cyc_total, cyc_wait, cyc_play, cyc_advance, cyc_exchange;
    unsigned long long n_units_played, n_advances, n_polls, n_units_published;
    unsigned long long cyc_adv_merge, cyc_adv_commit, cyc_adv_publish;   // phases of the advance
    unsigned pad_c[8];
};
static_assert(sizeof(UnitQueue) == 256 * RS_MAX_QUEUES + 256 + 384 + 1024, "queue header layout");
constexpr size_t RS_QUEUE_RESET_OFFSET = 256 * RS_MAX_QUEUES + 256;   // bytes of the header a reset keeps (the ticket counters)

// Kernel-parameter block of the resident search (passed by value beside SearchParams).
struct ResidentParams {
    UnitQueue* q;
    unsigned long long* slots;   // the rings of all unit queues (and, launch-per-depth, the chains' unit lists)
    int n_queues;
    unsigned ring_off[RS_MAX_QUEUES];    // first slot of queue k's ring
    unsigned ring_mask[RS_MAX_QUEUES];   // its capacity - 1
    const uint8_t* home_tab;     // [256] home queue of the SM with arrival rank r
    int gpc_max;                 // most units one child may be split into
    int units_per_warp_x4;       // target units in flight per resident warp, times 4
    int seed_in_flight;          // problems assumed in flight while depth 0 is being published
    uint8_t* ulist;              // [P][ES] legal entries of the current depth, in table order
    uint4* urec;                 // [P][urec_stride] one 16-byte record per unit of the current depth
    int urec_stride;
    // [P][ES] per child (rank among the legal entries) of the current depth: x = max placements,
    // y = first solved rollout (NONE_U32), z / w = placements of all rollouts (64 bits).  Every
    // unit folds its statistics in with three reductions at L2 (max, min and sum do not depend on
    // the order), so the advance reads ONE record per child instead of merging the child's unit
    // records in order; only a child with a solved rollout still needs them (its prefix count).
    uint4* cacc;
    // problems whose depth has been played, waiting for a server warp (slot = problem << 32 | ticket + 1)
    unsigned long long* done_slots;
    unsigned done_mask;
    int n_server_sms;            // the first n_server_sms SMs to receive a block advance problems, the others play
    int n_roll_warps;            // warps that play (sizes the work units)
    int poll_ns, poll_tier;      // rollout warps' poll back-off: cap (ns), warps per tier (0: one tier)
    // root-parallel exchange over peer memory (n_ranks > 1): every rank's buffers, own included.
    // The pointers live in a small device table, not in the parameter block: indexing a parameter
    // array with a run-time rank would move the whole block into local memory.
    struct PeerTable {
        ChildStats* gathered[RS_MAX_RANKS];   // [4][n_ranks][P][ES], buffer = (run & 1) * 2 + (depth & 1)
        int* arrived[RS_MAX_RANKS];           // [P][RS_MAX_RANKS] last (run << 8 | depth + 1) each rank reached
    };
    const PeerTable* peers;
    int run_epoch;
};

// Hand-over protocol of the resident kernel.  Everything one warp writes for another warp of the
// same launch is READ THROUGH L2 (ld.global.cg, atomics, strong loads), never from an L1 line, and
// every reader gets its addresses from the value it waited for (slot payload, atomic result), so
// a release on the writer's side is all the ordering that is needed:
//   writer:  plain stores ... st.release.gpu / atom.release.gpu   (MEMBAR.ALL.GPU + the store)
//   reader:  ld.relaxed.gpu poll / atomic result -> dependent ld.global.cg loads
// ld.acquire / __threadfence would add an L1 invalidate of the whole SM (CCTL.IVALL) to every
// poll and every unit, and __threadfence is the sequentially consistent MEMBAR.SC on top.
// Inside the out-of-line device functions the search's pointers come out of a parameter block
// in local memory, so the compiler no longer knows that they point to global memory and emits
// generic loads / stores and a shared-or-global switch around every atomic.  G() restores that.
template <typename T>
__device__ __forceinline__ T* G(T* ptr)
{
    __builtin_assume(__isGlobal(ptr));
    return ptr;
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* ptr)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* ptr, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(unsigned long long* ptr, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ int atom_add_release_s32(int* ptr, int v)
{
    int old;
    asm volatile("atom.release.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(ptr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ int ld_relaxed_sys_s32(const int* ptr)
{
    int v;
    asm volatile("ld.relaxed.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
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

// the whole state record as eight 16-byte stores / loads (every lane of a warp may load it: the
// address is uniform, the loads are broadcasts and nothing has to be shuffled afterwards)
__device__ __forceinline__ void store_state(ProbState* dst, const ProbState& st)
{
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4((uint32_t)st.depth, (uint32_t)st.n_alive, (uint32_t)st.hmin, (uint32_t)st.col);
    d[1] = make_uint4((uint32_t)st.run, (uint32_t)st.k, (uint32_t)st.group, (uint32_t)st.gpc);
    d[2] = make_uint4(st.legal[0], st.legal[1], st.legal[2], st.legal[3]);
    d[3] = make_uint4(st.alive[0], st.alive[1], st.alive[2], st.alive[3]);
    d[4] = make_uint4((uint32_t)st.n_order, (uint32_t)st.prev_tile, (uint32_t)st.status, 0u);
    d[5] = make_uint4((uint32_t)st.n_tiles_placed, (uint32_t)(st.n_tiles_placed >> 32),
                      (uint32_t)st.rollouts, (uint32_t)(st.rollouts >> 32));
    d[6] = make_uint4((uint32_t)st.rollout_placements, (uint32_t)(st.rollout_placements >> 32),
                      (uint32_t)st.rollouts_executed, (uint32_t)(st.rollouts_executed >> 32));
    d[7] = make_uint4((uint32_t)st.placements_executed, (uint32_t)(st.placements_executed >> 32), 0u, 0u);
}
__device__ __forceinline__ void load_state_cg(ProbState& st, const ProbState* src)
{
    const uint4* s = reinterpret_cast<const uint4*>(src);
    const uint4 q0 = __ldcg(s), q1 = __ldcg(s + 1), q2 = __ldcg(s + 2), q3 = __ldcg(s + 3);
    const uint4 q4 = __ldcg(s + 4), q5 = __ldcg(s + 5), q6 = __ldcg(s + 6), q7 = __ldcg(s + 7);
    st.depth = (int)q0.x; st.n_alive = (int)q0.y; st.hmin = (int)q0.z; st.col = (int)q0.w;
    st.run = (int)q1.x; st.k = (int)q1.y; st.group = (int)q1.z; st.gpc = (int)q1.w;
    st.legal[0] = q2.x; st.legal[1] = q2.y; st.legal[2] = q2.z; st.legal[3] = q2.w;
    st.alive[0] = q3.x; st.alive[1] = q3.y; st.alive[2] = q3.z; st.alive[3] = q3.w;
    st.n_order = (int)q4.x; st.prev_tile = (int)q4.y; st.status = (int)q4.z; st.pad0 = 0;
    st.n_tiles_placed = ((unsigned long long)q5.y << 32) | q5.x;
    st.rollouts = ((unsigned long long)q5.w << 32) | q5.z;
    st.rollout_placements = ((unsigned long long)q6.y << 32) | q6.x;
    st.rollouts_executed = ((unsigned long long)q6.w << 32) | q6.z;
    st.placements_executed = ((unsigned long long)q7.y << 32) | q7.x;
    st.pad1 = 0;
}

// a / b for a < 2^23, b >= 1: float quotient with a one-step fix-up (an integer division is ~20
// dependent instructions, and the advance does three of them on every problem's critical path)
__device__ __forceinline__ unsigned udiv_small(unsigned a, unsigned b)
{
    unsigned q = (unsigned)__fdividef((float)a, (float)b);
    if (q * b > a) --q;
    else if ((q + 1u) * b <= a) ++q;
    return q;
}

// Where the units of a problem's new depth go: the resident kernel's ring (tickets, mask =
// capacity - 1) or, in the launch-per-depth form, the plain unit list of the problem's chain for
// the next depth (counter = units of that launch, mask = all ones).  `in_flight` / `n_warps` size
// the units.
struct UnitSink {
    unsigned long long* tail;
    unsigned long long* slots;
    unsigned mask;
    int in_flight;
    unsigned n_warps;
    // launch-per-depth form: problems of the chain that entered the depth being advanced (read,
    // complete when the advance launch starts) and that enter the next one (counted here)
    const unsigned* active_now;
    unsigned* active_next;
};

// ---- store a problem's new state and publish the units of its depth --------------------------
// Called by a whole warp with the state in registers (the same values in every lane); the lane
// tables and planes are already written.  `in_flight`: problems that share the machine (decides
// the unit size: whole children while there is plenty of work, single chunks when one problem
// has the GPU to itself).
__device__ __forceinline__ void publish_units(const SearchParams& p, const ResidentParams& x, int prob,
                                              ProbState& st, const UnitSink& sink)
{
    const int in_flight = sink.active_now ? (int)*sink.active_now : sink.in_flight;
    const unsigned n_warps = sink.n_warps;
    if (sink.active_next && lane_id() == 0) atomicAdd(sink.active_next, 1u);
    const int lane = lane_id();
    const uint32_t lt = (1u << lane) - 1u;
    uint8_t* ul = x.ulist + (size_t)prob * p.es;
    int base = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (st.legal[j]) {                                       // warp-uniform
            if ((st.legal[j] >> lane) & 1u) ul[base + __popc(st.legal[j] & lt)] = (uint8_t)(lane + 32 * j);
            base += __popc(st.legal[j]);
        }
    }
    for (int ci = lane; ci < st.k; ci += 32)                     // the children's accumulators, empty
        x.cacc[(size_t)prob * p.es + ci] = make_uint4(0u, NONE_U32, 0u, 0u);
    // units per child: enough units that the problems in flight offer units_per_warp units to
    // every resident warp, no more than gpc_max (ring capacity) and RS_MAX_UNITS per problem
    const unsigned uk = (unsigned)st.k, nc = (unsigned)p.n_chunks;
    const unsigned want_units = (unsigned)__fdividef((float)(n_warps * (unsigned)x.units_per_warp_x4),
                                                     (float)(4 * max(in_flight, 1)));
    unsigned ugpc = min(nc, udiv_small(min(want_units, 1u << 20) + uk - 1u, uk));
    ugpc = max(1u, min(ugpc, min((unsigned)x.gpc_max, udiv_small((unsigned)RS_MAX_UNITS, uk))));
    st.group = (int)udiv_small(nc + ugpc - 1u, ugpc);
    st.gpc = (int)udiv_small(nc + (unsigned)st.group - 1u, (unsigned)st.group);   // equal units, none empty
    const int n_units = st.k * st.gpc;
    if (lane == 0) {
        store_state(p.pstate + prob, st);
        st_volatile(p.pending + prob, n_units);
    }
    // state and pending before any slot: the warp barrier orders every lane's writes before the
    // slot stores of all lanes; the first store of every lane is the release (one MEMBAR for the
    // warp), the later ones are issued after it and cannot overtake writes that are visible
    __syncwarp();
    unsigned long long t0 = 0;
    if (lane == 0) {
        t0 = atomicAdd(sink.tail, (unsigned long long)n_units);
        atomicAdd(&x.q->n_units_published, (unsigned long long)n_units);
    }
    t0 = __shfl_sync(FULLMASK, t0, 0);
    for (int i = lane; i < n_units; i += 32) {
        const unsigned long long t = t0 + (unsigned long long)i;
        const uint32_t payload = (uint32_t)prob | ((uint32_t)i << RS_PROBLEM_BITS);
        const unsigned long long word =
            ((unsigned long long)payload << 32) | (unsigned long long)(uint32_t)(t + 1ull);
        if (i < 32) st_release_u64(sink.slots + (t & sink.mask), word);
        else st_relaxed_u64(sink.slots + (t & sink.mask), word);
    }
    __syncwarp();
}

}  // namespace bp
