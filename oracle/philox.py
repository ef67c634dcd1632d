"""Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3").

TEST INFRASTRUCTURE (see oracle/__init__.py).  This file is the *specification* of
the random stream that replaces CPython's ``random.randint`` at the reference's
single hot-path call site (binpack/MCTS.py:174).  The CUDA kernels
(visapi_b200/csrc/philox.cuh) and the C oracle (oracle/rectpack_oracle.c)
implement the same function; tests check all three against the Random123
known-answer vectors.

Stream layout used by every rollout (one "simulation" of binpack/MCTS.py:152-201):

    key     = (seed, stream)          seed: run seed (123, run_mcts.py:187)
                                      stream: problem / root index
    counter = (sim, tag, block, 0)    sim:   simulation index n of MCTS.py:134
                                      tag:   (depth << 16) | entry for the flat
                                             search, caller-chosen for raw rollouts
                                      block: draw_index >> 2
    draw s of the rollout = philox(counter, key)[s & 3]
    index among K legal entries = (draw * K) >> 32
"""

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """ctr: 4 uint32, key: 2 uint32 -> tuple of 4 uint32."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = (
            ((p1 >> 32) ^ c1 ^ k0) & MASK,
            p1 & MASK,
            ((p0 >> 32) ^ c3 ^ k1) & MASK,
            p0 & MASK,
        )
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def flat_tag(depth, entry):
    """Counter word 1 for the flat search: which child of which depth."""
    return ((depth & 0xFFFF) << 16) | (entry & 0xFFFF)


class RolloutStream(object):
    """The draws of one rollout: draw(K) -> uniform index in [0, K)."""

    __slots__ = ("key", "sim", "tag", "s", "_block", "_words")

    def __init__(self, seed, stream, tag, sim):
        self.key = (seed & MASK, stream & MASK)
        self.sim = sim & MASK
        self.tag = tag & MASK
        self.s = 0
        self._block = -1
        self._words = None

    def next_u32(self):
        b = self.s >> 2
        if b != self._block:
            self._words = philox4x32_10((self.sim, self.tag, b, 0), self.key)
            self._block = b
        u = self._words[self.s & 3]
        self.s += 1
        return u

    def draw(self, k):
        return (self.next_u32() * k) >> 32
