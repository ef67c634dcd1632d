// Philox4x32-10 on the device; same function as oracle/philox.py (the stream spec).
//
// One evaluation per rollout: lane L of the warp computes block L of the
// rollout's stream, counter = (sim, tag, L, 0), key = (seed, stream).  Draw s of
// the rollout is word (s & 3) of lane (s >> 2), fetched with one shuffle, so a
// 128-draw budget costs ~70 instructions per rollout instead of per four draws.
#pragma once
#include <cstdint>

namespace bp {

struct Philox4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3, uint32_t k0, uint32_t k1)
{
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// the four words this lane contributes to rollout (seed, stream, tag, sim)
__device__ __forceinline__ Philox4 rollout_stream_block(uint32_t seed, uint32_t stream,
                                                        uint32_t tag, uint32_t sim,
                                                        uint32_t block)
{
    return philox4x32_10(sim, tag, block, 0u, seed, stream);
}

}  // namespace bp
