// Warp-per-board state and transition primitives (sm_100a).
//
// One warp owns one board.  Lane L holds rows L, L+32, ... (RHO of them), each as
// W 32-bit occupancy words (bit c%32 of word c/32 = column c; padding columns and
// padding rows are stored as occupied), and entries L, L+32, ... (EJ of them) of
// the ordered tile table, packed h | w << 8 with 0 = dead/absent.  Everything a
// rollout touches lives in registers; the only cross-lane traffic is ballots and
// one or two shuffles per step.
//
// Semantics follow the reference (paths relative to /root/reference):
//   find_lfb        engine/solution_checker.py:82-92  (row-major first free cell)
//                   + :290-305 (first occupied column to its right -> free run)
//   legal_masks     :307-317  (w <= run and lfb_row + h <= rows, table order)
//   place           :97-132   (fill [r, r+h) x [c, c+w); only the first footprint
//                   row is tested by the caller via `run`, exactly as the reference)
//   kill_pair       :319-332  (first occurrence of the tile, then of its rotation)
//   rollout         binpack/MCTS.py:152-201 with the Philox stream of philox.cuh
#pragma once
#include <cstdint>

#include "philox.cuh"

namespace bp {

constexpr uint32_t FULLMASK = 0xFFFFFFFFu;
constexpr int MAX_ROWS = 128;
constexpr int MAX_COLS = 128;
constexpr int MAX_ENTRIES = 128;

template <int RHO, int W, int EJ>
struct WarpState {
    uint32_t occ[RHO][W];
    uint32_t ent[EJ];
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int ctz32(uint32_t x) { return __clz(__brev(x)); }  // 32 for x == 0

// bits [lo, hi) of a 32-bit word, 0 <= lo < hi <= 32
__device__ __forceinline__ uint32_t bit_range(int lo, int hi)
{
    return (FULLMASK >> (32 - (hi - lo))) << lo;
}

// ---------------------------------------------------------------- load / store
// Host/HBM layout of a board: row-major words, `wpr` words per row, bit set = occupied.
// Padding (cols..32*W-1, rows..32*RHO-1) is filled in here.
template <int RHO, int W, int EJ>
__device__ __forceinline__ void load_board(WarpState<RHO, W, EJ>& s, const uint32_t* __restrict__ occ,
                                           int rows, int cols, int wpr)
{
    const int lane = lane_id();
#pragma unroll
    for (int k = 0; k < RHO; ++k) {
        const int row = lane + 32 * k;
#pragma unroll
        for (int q = 0; q < W; ++q) {
            uint32_t v = FULLMASK;
            if (row < rows && q < wpr) {
                v = occ ? occ[row * wpr + q] : 0u;
                const int live = cols - 32 * q;          // columns of this word that exist
                if (live < 32) v |= (live <= 0) ? FULLMASK : (FULLMASK << live);
            }
            s.occ[k][q] = v;
        }
    }
}

template <int RHO, int W, int EJ>
__device__ __forceinline__ void store_board(const WarpState<RHO, W, EJ>& s, uint32_t* __restrict__ occ,
                                            int rows, int cols, int wpr)
{
    const int lane = lane_id();
#pragma unroll
    for (int k = 0; k < RHO; ++k) {
        const int row = lane + 32 * k;
#pragma unroll
        for (int q = 0; q < W; ++q) {
            if (row < rows && q < wpr) {
                uint32_t v = s.occ[k][q];
                const int live = cols - 32 * q;
                if (live < 32) v &= ~(FULLMASK << live);  // strip the padding columns
                occ[row * wpr + q] = v;
            }
        }
    }
}

// tile table: uint8 (h, w) pairs in list order; (0, 0) = absent
template <int RHO, int W, int EJ>
__device__ __forceinline__ void load_entries(WarpState<RHO, W, EJ>& s, const uint8_t* __restrict__ tiles,
                                             int n_entries)
{
    const int lane = lane_id();
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(tiles);
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const int e = lane + 32 * j;
        s.ent[j] = (e < n_entries) ? (uint32_t)t16[e] : 0u;
    }
}

template <int RHO, int W, int EJ>
__device__ __forceinline__ void store_entries(const WarpState<RHO, W, EJ>& s, uint8_t* __restrict__ tiles,
                                              int n_entries)
{
    const int lane = lane_id();
    uint16_t* t16 = reinterpret_cast<uint16_t*>(tiles);
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const int e = lane + 32 * j;
        if (e < n_entries) t16[e] = (uint16_t)s.ent[j];
    }
}

template <int RHO, int W, int EJ>
__device__ __forceinline__ int count_alive(const WarpState<RHO, W, EJ>& s)
{
    int n = 0;
#pragma unroll
    for (int j = 0; j < EJ; ++j) n += __popc(__ballot_sync(FULLMASK, s.ent[j] != 0u));
    return n;
}

// ------------------------------------------------------------------------ LFB
// Row-major first free cell (r, c) and the number of free cells from it to the
// right in its row.  Returns false when the board is full.
template <int RHO, int W, int EJ>
__device__ __forceinline__ bool find_lfb(const WarpState<RHO, W, EJ>& s, int& r, int& c, int& run)
{
    int ksel = -1;
    uint32_t bsel = 0;
#pragma unroll
    for (int k = 0; k < RHO; ++k) {
        if (ksel < 0) {                                     // warp-uniform
            bool open = false;
#pragma unroll
            for (int q = 0; q < W; ++q) open |= (s.occ[k][q] != FULLMASK);
            const uint32_t b = __ballot_sync(FULLMASK, open);
            if (b) { ksel = k; bsel = b; }
        }
    }
    if (ksel < 0) return false;
    const int src = __ffs(bsel) - 1;
    // every lane scans its own row of slot ksel; only lane `src` is read back
    uint32_t x[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
        x[q] = s.occ[0][q];
#pragma unroll
        for (int k = 1; k < RHO; ++k) x[q] = (ksel == k) ? s.occ[k][q] : x[q];
    }
    int col, len;
    if (W == 1) {
        col = ctz32(~x[0]);
        const uint32_t t = x[0] >> (col & 31);
        len = min(ctz32(t), 32 - col);
    } else if (W == 2) {
        const unsigned long long x64 = (unsigned long long)x[0] | ((unsigned long long)x[1] << 32);
        col = __ffsll((long long)~x64) - 1;
        const unsigned long long t = x64 >> (col & 63);
        len = t ? (__ffsll((long long)t) - 1) : (64 - col);
    } else {
        col = 32 * W;
        len = 0;
        bool found = false, extending = false;
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const uint32_t xq = x[q];
            if (!found) {
                if (xq != FULLMASK) {
                    found = true;
                    const int c0 = ctz32(~xq);
                    col = 32 * q + c0;
                    const uint32_t t = xq >> c0;
                    if (t) len = ctz32(t); else { len = 32 - c0; extending = true; }
                }
            } else if (extending) {
                if (xq) { len += ctz32(xq); extending = false; } else len += 32;
            }
        }
    }
    const int packed = __shfl_sync(FULLMASK, col | (len << 8), src);
    r = 32 * ksel + src;
    c = packed & 0xFF;
    run = packed >> 8;
    return true;
}

// ---------------------------------------------------------------- legal moves
template <int RHO, int W, int EJ>
__device__ __forceinline__ int legal_masks(const WarpState<RHO, W, EJ>& s, int rows, int r, int run,
                                           uint32_t (&legal)[EJ])
{
    int k = 0;
    const uint32_t room_h = (uint32_t)(rows - r);
#pragma unroll
    for (int j = 0; j < EJ; ++j) {
        const uint32_t e = s.ent[j];
        const bool ok = (e != 0u) && ((e >> 8) <= (uint32_t)run) && ((e & 0xFFu) <= room_h);
        legal[j] = __ballot_sync(FULLMASK, ok);
        k += __popc(legal[j]);
    }
    return k;
}

// the idx-th legal entry in table order (entry = lane + 32 j): returns its packed value
template <int RHO, int W, int EJ>
__device__ __forceinline__ uint32_t select_legal(const WarpState<RHO, W, EJ>& s,
                                                 const uint32_t (&legal)[EJ], int idx, int& entry)
{
    const uint32_t lt = (1u << lane_id()) - 1u;
    int jsel = 0;
    uint32_t msel = legal[0];
    uint32_t esel = s.ent[0];
#pragma unroll
    for (int j = 1; j < EJ; ++j) {
        const int cnt = __popc(legal[j - 1]);
        // once idx falls inside word j-1 it stays selected
        const bool past = (jsel == j - 1) && (idx >= cnt);
        if (past) { idx -= cnt; jsel = j; msel = legal[j]; esel = s.ent[j]; }
    }
    const bool hit = ((msel >> lane_id()) & 1u) && (__popc(msel & lt) == idx);
    const int src = __ffs(__ballot_sync(FULLMASK, hit)) - 1;
    entry = src + 32 * jsel;
    return __shfl_sync(FULLMASK, esel, src);
}

// ---------------------------------------------------------------------- place
template <int RHO, int W, int EJ>
__device__ __forceinline__ void place(WarpState<RHO, W, EJ>& s, int r, int c, int h, int w)
{
    uint32_t m[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const int lo = max(c - 32 * q, 0);
        const int hi = min(c + w - 32 * q, 32);
        m[q] = (hi > lo) ? bit_range(lo, hi) : 0u;
    }
    const int lane = lane_id();
#pragma unroll
    for (int k = 0; k < RHO; ++k) {
        const bool in = (uint32_t)(lane + 32 * k - r) < (uint32_t)h;
#pragma unroll
        for (int q = 0; q < W; ++q) s.occ[k][q] |= in ? m[q] : 0u;
    }
}

// ------------------------------------------------------------------ kill pair
// Remove the lowest-index live entry equal to v, then the lowest-index live entry
// equal to its rotation (for a square: the two lowest).  Returns false when either
// is missing (the reference raises ValueError, solution_checker.py:325-329).
template <int RHO, int W, int EJ>
__device__ __forceinline__ bool kill_pair(WarpState<RHO, W, EJ>& s, uint32_t v)
{
    const uint32_t rot = ((v & 0xFFu) << 8) | (v >> 8);
    const int lane = lane_id();
    bool ok = true;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const uint32_t want = pass == 0 ? v : rot;
        bool done = false;
#pragma unroll
        for (int j = 0; j < EJ; ++j) {
            const uint32_t m = __ballot_sync(FULLMASK, s.ent[j] == want);
            if (!done && m) {                              // warp-uniform
                if (lane == __ffs(m) - 1) s.ent[j] = 0u;
                done = true;
            }
        }
        ok = ok && done;
    }
    return ok;
}

// -------------------------------------------------------------------- rollout
// One uniform-random playout from `s` (consumed).  n_alive = live entries in s.
// Returns the number of placements made (== the depth the reference returns for
// an unsolved rollout); solved = every tile was used.  When `moves` is non-null
// lane 0 records the (h, w) of every placement.
template <int RHO, int W, int EJ, bool RECORD>
__device__ __forceinline__ int rollout(WarpState<RHO, W, EJ> s, int rows, int n_alive, uint32_t seed,
                                       uint32_t stream, uint32_t tag, uint32_t sim, bool& solved,
                                       uint8_t* moves)
{
    const int lane = lane_id();
    const Philox4 rnd = rollout_stream_block(seed, stream, tag, sim, (uint32_t)lane);
    int steps = 0;
    solved = false;
    for (;;) {
        if (n_alive == 0) { solved = true; break; }        // MCTS.py:167-169
        int r, c, run;
        if (!find_lfb(s, r, c, run)) break;                // full board, tiles left
        uint32_t legal[EJ];
        const int k = legal_masks(s, rows, r, run, legal);
        if (k == 0) break;                                 // :171-172
        const int w4 = steps & 3;
        const uint32_t mine = w4 == 0 ? rnd.x : (w4 == 1 ? rnd.y : (w4 == 2 ? rnd.z : rnd.w));
        const uint32_t u = __shfl_sync(FULLMASK, mine, (steps >> 2) & 31);
        const int idx = (int)__umulhi(u, (uint32_t)k);     // :174
        int entry;
        const uint32_t v = select_legal(s, legal, idx, entry);
        const int h = (int)(v & 0xFFu), w = (int)(v >> 8);
        place(s, r, c, h, w);                              // :175-178
        kill_pair(s, v);                                   // :194
        if (RECORD) {
            if (lane == 0) { moves[2 * steps] = (uint8_t)h; moves[2 * steps + 1] = (uint8_t)w; }
        }
        n_alive -= 2;
        steps += 1;                                        // :179, :200
    }
    return steps;
}

}  // namespace bp
