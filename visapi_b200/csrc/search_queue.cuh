// Work queue of the resident search kernel: header, slots, and the producer side
// (publish_units).  The consumer side is the kernel itself (search_resident.cuh).
#pragma once
#include "search_types.cuh"

namespace bp {

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

}  // namespace bp
