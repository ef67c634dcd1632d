"""Host-side replay of a reported solution (the gate every benchmark result passes).

A solution is the reference's ``solution_tiles_order``: tiles in the order they are
placed, each with its top-left corner on the row-major first free cell
(engine/solution_checker.py:82-132 of the reference).  The replay checks, on a plain
cell grid, that every tile fits inside the board on free cells only, that the tiles
used are exactly the instance's tiles (either orientation), and that the board ends
full.  It is independent of the CUDA kernels and of oracle/.
"""
import numpy as np


def replay_solution(rows, cols, base_tiles, order):
    """Returns (ok, reason).  base_tiles: the n tiles (one orientation each)."""
    grid = np.zeros((rows, cols), dtype=np.int32)
    remaining = {}
    for t in base_tiles:
        key = (min(int(t[0]), int(t[1])), max(int(t[0]), int(t[1])))
        remaining[key] = remaining.get(key, 0) + 1
    if len(order) != len(base_tiles):
        return False, "order has %d tiles, instance has %d" % (len(order), len(base_tiles))
    for k, t in enumerate(order):
        h, w = int(t[0]), int(t[1])
        free = np.flatnonzero(grid.ravel() == 0)
        if free.size == 0:
            return False, "board full before tile %d" % k
        r, c = divmod(int(free[0]), cols)
        if r + h > rows or c + w > cols:
            return False, "tile %d (%d,%d) leaves the board at (%d,%d)" % (k, h, w, r, c)
        if grid[r:r + h, c:c + w].any():
            return False, "tile %d (%d,%d) overlaps at (%d,%d)" % (k, h, w, r, c)
        key = (min(h, w), max(h, w))
        if remaining.get(key, 0) == 0:
            return False, "tile %d (%d,%d) is not in the instance" % (k, h, w)
        remaining[key] -= 1
        grid[r:r + h, c:c + w] = k + 1
    if (grid == 0).any():
        return False, "board not full"
    return True, "ok"
