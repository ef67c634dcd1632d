// Lane-per-rollout playout: 32 independent rollouts of one child advance in lock step,
// one per lane, every lane holding a whole board in registers.
//
// Why a second mapping.  The warp-per-board rollout (board.cuh) spends a full warp
// instruction on work that is one scalar operation per board: ncu on
// fs_rollout_kernel<1,1,2> shows 137 warp instructions per placement with the ALU pipe 85 %
// and the XU pipe (POPC/FLO/BREV) 75 % busy (profiles/r1).  Here each warp instruction
// advances 32 boards, and a placement costs ~6 warp instructions.
//
// Board = skyline.  Every tile is placed with its top-left corner on the row-major first
// free cell, so each column's occupied cells are a prefix from row 0 (the reference relies on
// this: engine/solution_checker.py:107).  A board is therefore height[c] per column, stored
// BIT-SLICED: plane b, word q holds bit b of the heights of columns 32q..32q+31.  With that
//   LFB            = leftmost column of minimum height: NB masked steps from the top plane
//                    down (solution_checker.py:82-92);
//   free run       = the run of minimum-height columns starting there (:290-305):
//                    cand & ~(cand + colbit);
//   place (h, w)   = set the heights of the footprint columns to hmin + h: one 3-input
//                    logic op per plane, no loop over the tile's cells (:123);
//   legal entries  = alive & wmask[run] & hmask[rows - hmin], two table rows prepared once per
//                    depth by the advance kernel (:307-317);
//   pair removal   = clear the lowest alive bit of the tile's group of equal entries and of
//                    its rotation's group (:319-332).
// Padding columns carry the maximum height 2^NB - 1 >= rows, so they are never free.
//
// The random stream, the tie rules and the results are exactly those of board.cuh::rollout;
// tests/test_gpu_parity.py checks both against the oracle and the reference-made fixtures.
#pragma once
#include <cstdint>

#include "board.cuh"

namespace bp {

// Per-problem tables prepared once per depth by the advance kernel (write_lane_tables):
//   planes[8][4]        bit-sliced column heights of the current board
//   wmask[x][4], hmask[y][4]   live entries with w <= x / h <= y   (x, y = 0..128)
//   pairs[e][4]         the two entries that leave the list when entry e is played: e itself
//                       and its partner, the entry of the rotated tile with the same rank
//                       inside its group of equal entries (for a square: the neighbour in its
//                       own group).  The reference removes the FIRST equal entry and the FIRST
//                       rotated entry (solution_checker.py:319-332); with equal entries adjacent
//                       in the list (checked on the host) removing any member of a group leaves
//                       the same sequence of tiles, so the rollouts are identical and the removal
//                       is two AND-NOTs instead of two find-lowest-bit chains.
//   tile[e]             (h, w, width mask (1 << w) - 1 as two words); in shared memory (h, mask)
//                       for one-word boards and all four words otherwise
constexpr int LANES_TABLE_ROWS = 129;
constexpr int LANES_MASK_WORDS = 4;
constexpr int LANES_PLANE_STRIDE = 32;
constexpr int LANES_PAIR_WORDS = 4;

// The value 1, read from constant memory so that ptxas cannot fold it: `x * FMA_ONE + y` stays an
// IMAD (FMA pipe) where a plain move / add / select would be scheduled on the ALU pipe that
// binds this kernel (LOP3, SEL, ISETP, SHF issue at half rate on sm_100).
static __constant__ uint32_t FMA_ONE = 1u;

// Shared-memory record of one entry: its tile and its pair mask behind ONE address, so a step
// computes one table address and issues one or two vector loads.  Built from the global
// tile record (h, w, width mask lo, hi) and the entry's pair-mask words.
template <int W, int EW>
struct __align__(16) LaneEntry {                  // 32 bytes
    uint4 t;                                      // h, w, width mask lo / hi
    uint32_t pr[4];                               // pair mask
    __device__ __forceinline__ void set(const uint4& tile, const uint32_t (&pair)[4])
    {
        t = tile;
#pragma unroll
        for (int q = 0; q < 4; ++q) pr[q] = pair[q];
    }
    __device__ __forceinline__ uint32_t h() const { return t.x; }
    __device__ __forceinline__ uint32_t w() const { return t.y; }
    __device__ __forceinline__ uint32_t wm_lo() const { return t.z; }
    __device__ __forceinline__ uint32_t wm_hi() const { return t.w; }
    __device__ __forceinline__ uint32_t pair(int q) const { return pr[q]; }
};
template <>
struct __align__(16) LaneEntry<1, 2> {            // 16 bytes: one-word board, <= 64 entries
    uint4 v;                                      // h, width mask, pair mask lo / hi
    __device__ __forceinline__ void set(const uint4& tile, const uint32_t (&pair)[4])
    {
        v = make_uint4(tile.x, tile.z, pair[0], pair[1]);
    }
    __device__ __forceinline__ uint32_t h() const { return v.x; }
    __device__ __forceinline__ uint32_t w() const { return 0u; }
    __device__ __forceinline__ uint32_t wm_lo() const { return v.y; }
    __device__ __forceinline__ uint32_t wm_hi() const { return 0u; }
    __device__ __forceinline__ uint32_t pair(int q) const { return q == 0 ? v.z : v.w; }
};

template <int NB, int W, int EW>
struct LaneBoard {
    uint32_t pl[NB][W];    // bit-sliced column heights, stored INVERTED (bit set = height bit clear)
    uint32_t alive[EW];    // live entries of the (compacted) tile list
};

// bits [lo, hi) of word q of a multi-word mask
__device__ __forceinline__ uint32_t word_range(int lo, int hi, int q)
{
    const int a = max(lo - 32 * q, 0), b = min(hi - 32 * q, 32);
    return (b > a) ? bit_range(a, b) : 0u;
}

// Position of the k-th (0-based) set bit of m, k < popc(m), through a 2 KB shared-memory
// table: sel8[b * 8 + j] = position of the j-th set bit of byte b.  The kernel is bound by the
// ALU pipe (LOP3 / SEL / ISETP / SHF issue at half rate), so the byte search avoids it:
//   c0..c2  = popcounts of the low 1, 2, 3 bytes (shifts as IMAD.SHL on the FMA pipe, POPC on XU)
//   a       = 0x808080 + (k-c0) + (k-c1) << 8 + (k-c2) << 16   (four IMADs; every byte stays in
//             [96, 159], so there is no borrow between bytes): bit 7 of byte j  <=>  k >= c_j
//   byte    = popc(a & 0x808080)            -- the byte of m that holds the bit
//   a2      = a << 8 | 0x80 + k             -- byte j = 0x80 + k - (popcount below byte j)
// and two byte permutes (PRMT) pull byte `byte` out of m and of a2: 3 ALU-pipe instructions
// instead of the 17 of a compare/select chain.
// `kp` = 0x80 + k, `base` is added to the result.
__device__ __forceinline__ int select_bit32_lut(uint32_t m, uint32_t kp, uint32_t base,
                                                const uint8_t* __restrict__ sel8)
{
    const uint32_t c0 = __popc(m << 24), c1 = __popc(m << 16), c2 = __popc(m << 8);
    uint32_t a = kp * 0x010101u;
    a -= c0;
    a -= c1 * 256u;
    a -= c2 * 65536u;
    const uint32_t byte = __popc(a & 0x808080u);
    const uint32_t a2 = a * 256u + kp;
    uint32_t sel;                                        // result byte 0 = source byte, rest 0
    asm("mad.lo.u32 %0, %1, %2, 0x4440;" : "=r"(sel) : "r"(byte), "r"(FMA_ONE));
    uint32_t bits, rank;                                 // rank = 0x80 + rank inside the byte
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(bits) : "r"(m), "r"(sel));
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(rank) : "r"(a2), "r"(sel));
    return (int)(8u * byte + base) + (int)sel8[bits * 8u + rank - 0x80u];
}

// fill sel8[256 * 8] (call with all threads of the block, then __syncthreads)
__device__ __forceinline__ void fill_select_table(uint8_t* sel8)
{
    for (int b = threadIdx.x; b < 256; b += blockDim.x) {
        int k = 0;
#pragma unroll
        for (int bit = 0; bit < 8; ++bit)
            if ((b >> bit) & 1) sel8[b * 8 + k++] = (uint8_t)bit;
        for (; k < 8; ++k) sel8[b * 8 + k] = 0;
    }
}

// raise the footprint columns from height `from` (all of them) to `to`: flip plane b under
// the footprint wherever bit b differs (a flip is the same on inverted planes)
template <int NB, int W, int EW>
__device__ __forceinline__ void lanes_place(LaneBoard<NB, W, EW>& s, const uint32_t (&foot)[W], int from, int to)
{
    const uint32_t diff = (uint32_t)(from ^ to);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int q = 0; q < W; ++q) {
            // one test + one predicated XOR per plane word
            asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                "and.b32 t, %2, %3;\n\t"
                "setp.ne.u32 p, t, 0;\n\t"
                "@p xor.b32 %0, %0, %1;\n\t}"
                : "+r"(s.pl[b][q]) : "r"(foot[q]), "r"(diff), "r"(1u << b));
        }
    }
}

// Leftmost column of minimum height: cand = columns of minimum height; returns that height in
// the planes' own inverted form, inv = (2^NB - 1) - hmin, which is what the callers need (the
// hmask rows are laid out by inv, and a placement flips the bits where inv and inv - h differ).
// Per plane: one LOP3 with predicate output on the ALU pipe (the binding pipe); the
// conditional narrowing and the height bit are predicated multiply-adds, which issue on the
// FMA pipe (FMA_ONE is opaque to ptxas so they are not folded back into SEL / IADD3).
template <int NB, int W, int EW>
__device__ __forceinline__ int lanes_min(const LaneBoard<NB, W, EW>& s, uint32_t (&cand)[W])
{
    const uint32_t one = FMA_ONE;
    // top plane: every column is still a candidate
    uint32_t top = 0u;
#pragma unroll
    for (int q = 0; q < W; ++q) top |= s.pl[NB - 1][q];
    const bool top_set = top != 0u;
#pragma unroll
    for (int q = 0; q < W; ++q) cand[q] = top_set ? s.pl[NB - 1][q] : 0xFFFFFFFFu;
    int inv = top_set ? (1 << (NB - 1)) : 0;
#pragma unroll
    for (int b = NB - 2; b >= 0; --b) {
        if (W == 1) {
            asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                "and.b32 t, %0, %2;\n\t"
                "setp.ne.u32 p, t, 0;\n\t"
                "@p mad.lo.u32 %0, t, %3, 0;\n\t"
                "@p mad.lo.u32 %1, %3, %4, %1;\n\t}"
                : "+r"(cand[0]), "+r"(inv) : "r"(s.pl[b][0]), "r"(one), "r"(1u << b));
        } else {
            // any = OR of the narrowed words; each word is replaced by a predicated FMA-pipe move
            uint32_t z[W];
            uint32_t any = 0u;
#pragma unroll
            for (int q = 0; q < W; ++q) { z[q] = cand[q] & s.pl[b][q]; any |= z[q]; }
            asm("{\n\t.reg .pred p;\n\t"
                "setp.ne.u32 p, %3, 0;\n\t"
                "@p mad.lo.u32 %0, %2, %4, 0;\n\t"
                "@p mad.lo.u32 %1, %4, %5, %1;\n\t}"
                : "+r"(cand[0]), "+r"(inv) : "r"(z[0]), "r"(any), "r"(one), "r"(1u << b));
#pragma unroll
            for (int q = 1; q < W; ++q)
                asm("{\n\t.reg .pred p;\n\t"
                    "setp.ne.u32 p, %2, 0;\n\t"
                    "@p mad.lo.u32 %0, %1, %3, 0;\n\t}"
                    : "+r"(cand[q]) : "r"(z[q]), "r"(any), "r"(one));
        }
    }
    return inv;
}

// One rollout per lane, all 32 lanes of the warp in lock step (same child, sims sim0 + lane).
// `root` is the child state (identical in every lane).  Tables live in this warp's shared
// memory: wmask[x * EW + q], hmask[inv * EW + q] (row inv = entries that fit above a column of
// inverted height inv), entries[e].
// Returns the number of placements of this lane's rollout; solved = all tiles used.
template <int NB, int W, int EW>
__device__ __forceinline__ int lanes_rollout(LaneBoard<NB, W, EW> s, int rows, int n_alive,
                                             const uint32_t* __restrict__ wmask,
                                             const uint32_t* __restrict__ hmask,
                                             const LaneEntry<W, EW>* __restrict__ entries,
                                             const uint8_t* __restrict__ sel8, uint32_t seed,
                                             uint32_t stream, uint32_t tag, uint32_t sim, int active,
                                             int& solved)
{
    // Every placement removes two entries, so the step count and the verdict follow from the
    // number of live entries left: no counters in the loop.
    const int n_alive0 = n_alive;
    const int playing = active;
    if (n_alive == 0) active = 0;                                // MCTS.py:167-169: nothing to place
    // one lock-step iteration; `u` is this lane's draw for the step
    // The body is branch-free: a lane whose rollout is over keeps stepping a dead state in lock
    // step (the warp would issue the instructions for the other lanes anyway) and only its
    // live-entry count is frozen.  With no legal entry the select reads a fixed in-range entry
    // (legal mask 0, idx 0), so every table address stays valid.
    auto iterate = [&](uint32_t u) {
            {
                {
                    uint32_t cand[W];
                    const int inv = lanes_min(s, cand);          // (2^NB - 1) - minimum height
                    // first candidate column and the run of candidates that starts there
                    uint32_t foot[W];
                    uint32_t colbit1 = 0u;
                    unsigned long long colbit64 = 0ull;
                    bool in_lo = true;                            // W > 2: colbit64 is in the low half
                    int col = 0, run = 0;
                    if (W == 1) {
                        colbit1 = cand[0] & (0u - cand[0]);
                        const uint32_t runmask = cand[0] & ~(cand[0] + colbit1);
                        run = __popc(runmask);
                    } else if (W == 2) {
                        const unsigned long long c64 =
                            (unsigned long long)cand[0] | ((unsigned long long)cand[W > 1 ? 1 : 0] << 32);
                        colbit64 = c64 & (0ull - c64);
                        const unsigned long long runmask = c64 & ~(c64 + colbit64);
                        run = __popcll(runmask);
                    } else {
                        // W = 3 or 4: two 64-bit halves; the run of the low half continues into
                        // the high half when it reaches bit 63
                        const unsigned long long c_lo =
                            (unsigned long long)cand[0] | ((unsigned long long)cand[W > 1 ? 1 : 0] << 32);
                        const unsigned long long c_hi =
                            (unsigned long long)cand[W > 2 ? 2 : 0] |
                            (W > 3 ? ((unsigned long long)cand[W > 3 ? 3 : 0] << 32) : 0ull);
                        in_lo = c_lo != 0ull;
                        const unsigned long long x = in_lo ? c_lo : c_hi;
                        const unsigned long long cb = x & (0ull - x);
                        const unsigned long long runmask = x & ~(x + cb);
                        colbit64 = cb;
                        col = __popcll(cb - 1ull) + (in_lo ? 0 : 64);
                        run = __popcll(runmask);
                        if (in_lo && (runmask >> 63)) run += __popcll(c_hi & ~(c_hi + 1ull));
                    }
                    uint32_t legal[EW], cnt[EW];
                    int k = 0;
#pragma unroll
                    for (int q = 0; q < EW; ++q) {
                        legal[q] = s.alive[q] & wmask[run * EW + q] & hmask[inv * EW + q];
                        cnt[q] = (uint32_t)__popc(legal[q]);
                        k += (int)cnt[q];
                    }
                                    active = min(active, k);                     // 0 / 1 flag; k == 0: MCTS.py:171-172
                    {
                        const uint32_t idx = __umulhi(u, (uint32_t)k);          // MCTS.py:174
                        // word of the legal mask that holds the idx-th entry; kp = 0x80 + rank
                        // inside that word.  Predicated FMA-pipe moves instead of selects.
                        uint32_t kp, m = legal[0], base = 0u;
                        asm("mad.lo.u32 %0, %1, %2, 0x80;" : "=r"(kp) : "r"(idx), "r"(FMA_ONE));
                        if (EW == 2) {
                            asm("{\n\t.reg .pred p;\n\t"
                                "setp.ge.u32 p, %3, %4;\n\t"
                                "@p mad.lo.u32 %0, %5, %6, 0;\n\t"
                                "@p sub.u32 %1, %1, %4;\n\t"
                                "@p mad.lo.u32 %2, %6, 32, 0;\n\t}"
                                : "+r"(m), "+r"(kp), "+r"(base)
                                : "r"(idx), "r"(cnt[0]), "r"(legal[EW > 1 ? 1 : 0]), "r"(FMA_ONE));
                        } else {
                            uint32_t rem = idx;
#pragma unroll
                            for (int q = 1; q < EW; ++q) {
                                const bool past = (base == 32u * (q - 1)) && (rem >= cnt[q - 1]);
                                rem -= past ? cnt[q - 1] : 0u;
                                kp -= past ? cnt[q - 1] : 0u;
                                m = past ? legal[q] : m;
                                base = past ? 32u * q : base;
                            }
                        }
                        const int e = select_bit32_lut(m, kp, base, sel8);
                        const LaneEntry<W, EW> t = entries[e];
                        const int th = (int)t.h();
                        if (W == 1) {
                            foot[0] = colbit1 * t.wm_lo();           // width mask shifted to col
                        } else if (W == 2) {
                            // 64-bit product colbit * width mask: three IMADs, no shifts
                            const unsigned long long f =
                                colbit64 * ((unsigned long long)t.wm_lo() |
                                            ((unsigned long long)t.wm_hi() << 32));
                            foot[0] = (uint32_t)f;
                            foot[W > 1 ? 1 : 0] = (uint32_t)(f >> 32);
                        } else {
                            const int tw = (int)t.w();
                            if (tw <= 64) {
                                // footprint = width mask * (one-hot column bit).  The column bit has
                                // ONE non-zero 32-bit limb pw, so the product is (mask * pw) -- two
                                // wide multiplies whose halves do not overlap -- moved up by the
                                // limb's index (0..3) with selects
                                const uint32_t pw_hi = (uint32_t)(colbit64 >> 32);
                                const uint32_t pw = (uint32_t)colbit64 | pw_hi;
                                const unsigned long long p0 = (unsigned long long)t.wm_lo() * pw;
                                const unsigned long long p1 = (unsigned long long)t.wm_hi() * pw;
                                const uint32_t q0 = (uint32_t)p0;
                                const uint32_t q1 = (uint32_t)(p0 >> 32) | (uint32_t)p1;
                                const uint32_t q2 = (uint32_t)(p1 >> 32);
                                const bool up1 = pw_hi != 0u;
                                const uint32_t r0 = up1 ? 0u : q0, r1 = up1 ? q0 : q1;
                                const uint32_t r2 = up1 ? q1 : q2, r3 = up1 ? q2 : 0u;
                                foot[0] = in_lo ? r0 : 0u;
                                foot[W > 1 ? 1 : 0] = in_lo ? r1 : 0u;
                                foot[W > 2 ? 2 : 0] = in_lo ? r2 : r0;
                                if (W > 3) foot[W > 3 ? 3 : 0] = in_lo ? r3 : r1;
                            } else {                              // a tile wider than 64 columns
#pragma unroll
                                for (int q = 0; q < W; ++q) foot[q] = word_range(col, col + tw, q);
                            }
                        }
                        lanes_place(s, foot, inv, inv - th);                    // MCTS.py:175-178
#pragma unroll
                        for (int q = 0; q < EW; ++q) s.alive[q] &= ~t.pair(q);       // MCTS.py:194
                        n_alive -= 2 * active;
                        active = min(active, n_alive);           // MCTS.py:167-169 of the next step
                    }
                }
            }
    };
    // four iterations per Philox block, unrolled so that the word of the block is static
    for (uint32_t block = 0;; ++block) {
        const Philox4 rnd = rollout_stream_block(seed, stream, tag, sim, block);
        iterate(rnd.x);
        iterate(rnd.y);
        if (!__any_sync(FULLMASK, active != 0)) break;   // (an extra step on dead lanes is harmless)
        iterate(rnd.z);
        iterate(rnd.w);
        if (!__any_sync(FULLMASK, active != 0)) break;
    }
    solved = (playing && n_alive == 0) ? 1 : 0;
    return (n_alive0 - n_alive) >> 1;
}

}  // namespace bp
