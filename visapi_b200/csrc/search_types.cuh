// Flat search: constants, HBM records and the parameter block shared by the search kernels
// (search_kernels.cuh, search_resident.cuh) and the host side (flat_search.cu).
#pragma once
#include <cstdint>
#include <cstring>

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
// A class's problems may be dealt to several CHAINS (chain = class + FS_CLASSES * k): every chain
// has its own child list, counters and launch sequence on its own stream, so that a small batch
// still has enough independent launch sequences in flight to fill each other's ramps and tails.
constexpr int FS_MAX_SUBCHAINS = 4;
constexpr int FS_CHAINS = FS_CLASSES * FS_MAX_SUBCHAINS;

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

// Everything the resident kernel changes in a problem from depth to depth, in ONE 128-byte
// record: a playout reads three 16-byte words of it, the advance reads and writes it whole, and
// the counters need no atomics (only the advancing warp touches them).  The arrays the host
// reads (status, depth, n_order, counters) are written once, when the problem is finished.
struct __align__(16) ProbState {
    int depth, n_alive, hmin, col;                 // word 0: depth; live entries; first free cell
    int run, k, group, gpc;                        // word 1: free run; legal children; chunks per unit, units per child
    uint32_t legal[4];                             // word 2: legal entries of the initial tile list
    uint32_t alive[4];                             // word 3: live entries of the initial tile list
    int n_order, prev_tile, status, pad0;          // word 4
    unsigned long long n_tiles_placed, rollouts;                   // word 5
    unsigned long long rollout_placements, rollouts_executed;      // word 6
    unsigned long long placements_executed, pad1;                  // word 7
};
static_assert(sizeof(ProbState) == 128, "one cache line");

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
    unsigned* child_count;   // [MAX_DEPTHS + 1][FS_CHAINS]
    unsigned* task_counter;  // [MAX_DEPTHS + 1][FS_CHAINS]
    // lane-per-rollout tables of the current state of each problem (lanes.cuh)
    int use_lanes;
    uint32_t* planes;        // [P][8][4] bit-sliced column heights, inverted
    uint32_t* pairs;         // [P][ES][4] bits of the two entries that leave when e is played
    uint4* tile;             // [P][ES] (h, w, width mask (1 << w) - 1 as lo / hi word)
    uint32_t* wmask;         // [P][129][4] entries with w <= x
    uint32_t* hmask;         // [P][129][4] entries with h <= y
    // resident (single-launch) search: units of the current depth of each problem still running,
    // and the live entries of the initial (never compacted) tile list
    int* pending;            // [P]
    uint4* alive;            // [P] (as the reset leaves it; the resident kernel continues in pstate)
    ProbState* pstate;       // [P]
};

// Programmatic dependent launch (the launch-per-depth chains, flat_search.cu): a kernel launched with
// the attribute may start while the kernel before it in the stream is still running.  It must not
// touch what that kernel writes before grid_dependency_wait() (returns when the previous kernel
// has completed and its writes are visible; a no-op in a plain launch); the previous kernel lets
// it start with grid_launch_dependents() (else: when its blocks exit).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

}  // namespace bp
