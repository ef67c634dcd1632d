"""visapi_b200 — B200-native rollout engine for perfect rectangle packing.

Drop-in for the MCTS hot path of igorpejic/visapi: the same ``SolutionChecker`` /
``State`` / ``CustomMCTS`` / ``MCTS`` / ``BinPackGame`` surface, executed by
hand-written sm_100a CUDA kernels behind the C ABI of include/rectpack_b200.h.
There is no CPU fallback: without the built library and a CUDA device the engine
raises ``EngineUnavailable``.
"""
from ._native import (Engine, EngineError, EngineUnavailable, Search,  # noqa: F401
                      load_library, pack_occupancy, pack_tiles, unpack_occupancy)

__all__ = ["Engine", "EngineError", "EngineUnavailable", "Search", "load_library",
           "pack_occupancy", "pack_tiles", "unpack_occupancy"]
