"""Edge cases of the CUDA path against the C oracle: extreme shapes, every shape class,
unsorted and duplicate-heavy tile lists, non-empty initial boards, both rollout kernels."""
import os
import random

import numpy as np
import pytest

from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import visapi_b200
    e = visapi_b200.Engine(0)
    yield e
    e.close()


def gen(n, rows, cols, seed, order=True):
    from visapi_b200.data_generator import DataGenerator
    random.seed(seed)
    tiles, board = DataGenerator().gen_tiles_and_board(n, rows, cols, seed=seed, order_tiles=True)
    tiles = [(int(t[0]), int(t[1])) for t in tiles]
    if not order:
        random.Random(seed).shuffle(tiles)
    return tiles, board


SHAPES = [
    (1, 1, 1), (1, 5, 3), (2, 1, 9), (3, 2, 2),            # degenerate
    (6, 31, 32), (6, 32, 31), (7, 33, 33), (10, 63, 64),    # class boundaries
    (10, 64, 65), (12, 65, 20), (9, 20, 97), (16, 127, 128),
    (20, 128, 128), (40, 100, 100), (64, 128, 128),         # maximum board and tile count
    (33, 60, 60), (50, 25, 25),                             # > 64 entries
]


def _run_both_kernels(eng, rows, cols, tiles, n_sim, strategy, stream, boards=None):
    import visapi_b200
    from visapi_b200 import pack_tiles
    out = []
    # lane kernel with one launch pair per depth (the default), lane kernel inside the resident
    # search kernel, warp-per-board kernel
    for impl, mode in (("lanes", "stepped"), ("lanes", "resident"), ("warp", "stepped")):
        os.environ["BP_ROLLOUT"] = impl
        os.environ["BP_SEARCH"] = mode
        try:
            out.append(eng.flat_search([rows], [cols], pack_tiles([tiles], len(tiles) + len(tiles) % 2),
                                       [len(tiles)], n_sim, strategy, 7, [stream], boards=boards,
                                       want_scores=True))
        except visapi_b200.EngineError as e:
            # the resident kernel only exists for the lane kernel's inputs; it must say so
            if not (mode == "resident" and "needs the lane kernel" in str(e)):
                raise
        finally:
            os.environ.pop("BP_ROLLOUT", None)
            os.environ.pop("BP_SEARCH", None)
    return out


def _check(res, want, i=0):
    assert int(res["depth"][i]) == want["depth"]
    assert bool(res["solution_found"][i]) == want["solution_found"]
    assert int(res["n_tiles_placed"][i]) == want["n_tiles_placed"]
    assert int(res["rollouts"][i]) == want["rollouts"]
    assert int(res["rollout_placements"][i]) == want["rollout_placements"]
    assert res["order"][i, :res["n_order"][i]].tolist() == want["solution_tiles_order"]
    assert [int(v) for v in res["path"][i] if v >= 0] == want["path"]


@pytest.mark.parametrize("n,rows,cols", SHAPES)
def test_flat_search_shapes(eng, n, rows, cols):
    tiles, board = gen(n, rows, cols, seed=n * 1000 + rows)
    n_sim = 37 if n <= 20 else 5
    for strategy in (0, 1):
        want = c_oracle.flat_search(tiles, board, n_sim, "max_depth" if strategy == 0 else "avg_depth",
                                    7, 3)
        for res in _run_both_kernels(eng, rows, cols, tiles, n_sim, strategy, 3):
            _check(res, want)


def test_flat_search_unsorted_lists(eng):
    """equal tiles not adjacent in the list: first-occurrence removal order matters"""
    for seed in range(6):
        tiles, board = gen(14, 17, 23, seed=seed, order=False)
        want = c_oracle.flat_search(tiles, board, 21, "max_depth", 7, seed)
        for res in _run_both_kernels(eng, 17, 23, tiles, 21, 0, seed):
            _check(res, want)


def test_flat_search_many_duplicates(eng):
    tiles = sorted([(1, 1)] * 30 + [(2, 1), (1, 2)] * 5, key=lambda x: (x[1], x[0]))
    board = np.zeros((5, 5))               # 15 + 10 = 25 cells
    want = c_oracle.flat_search(tiles, board, 9, "max_depth", 7, 0)
    for res in _run_both_kernels(eng, 5, 5, tiles, 9, 0, 0):
        _check(res, want)


def test_flat_search_from_nonempty_board(eng):
    """an initial board that is not a skyline: the warp-per-board kernel is selected"""
    from visapi_b200 import pack_occupancy
    tiles, _ = gen(8, 9, 11, seed=5)
    board = np.zeros((11, 12))
    board[:, 11] = 1                        # a wall on the right
    board[0:2, :] = 1                       # two filled rows on top
    board[10, 3] = 1                        # a hole-making cell: free cells above an occupied one
    want = c_oracle.flat_search(tiles, board, 15, "max_depth", 7, 1)
    occ = pack_occupancy(board[None])
    for res in _run_both_kernels(eng, 11, 12, tiles, 15, 0, 1, boards=occ):
        _check(res, want)
    s = eng.search([11], [12], __import__("visapi_b200").pack_tiles([tiles]), [len(tiles)], 15, 0, 7, [1],
                   boards=occ)
    assert s.info()["lane_kernel"] is False
    s.close()


def test_mixed_classes_one_batch(eng):
    """problems of many shape classes and tile counts in a single batch"""
    from visapi_b200 import pack_tiles
    cases = [gen(n, r, c, seed=n + r + c) + (r, c) for (n, r, c) in SHAPES if n <= 20 and n > 1]
    es = max(len(t[0]) for t in cases)
    res = eng.flat_search([c[2] for c in cases], [c[3] for c in cases],
                          pack_tiles([c[0] for c in cases], es), [len(c[0]) for c in cases], 11, 0, 7,
                          list(range(len(cases))), want_scores=True)
    for i, (tiles, board, r, c) in enumerate(cases):
        _check(res, c_oracle.flat_search(tiles, board, 11, "max_depth", 7, i), i)


def test_rollouts_api_shapes(eng):
    from visapi_b200 import pack_occupancy, pack_tiles
    for (n, rows, cols) in [(20, 128, 128), (64, 128, 128), (12, 65, 20), (9, 20, 97), (1, 1, 1)]:
        tiles, board = gen(n, rows, cols, seed=rows + cols, order=False)
        steps, solved, moves = eng.rollouts(pack_occupancy(board[None]), cols,
                                            pack_tiles([tiles], len(tiles)), 40, seed=11, streams=[5],
                                            tags=[9], want_moves=True)
        for s in range(0, 40, 3):
            o = c_oracle.rollout(board, tiles, 11, 5, 9, s)
            assert bool(solved[0, s]) == o["solved"] and int(steps[0, s]) == o["placements"]
            assert moves[0, 0 + s, :steps[0, s]].tolist() == o["moves"]


def test_empty_batches_and_zero_entries(eng):
    from visapi_b200 import pack_occupancy
    occ = pack_occupancy(np.zeros((0, 4, 4)))
    assert eng.lfb(occ.reshape(0, 4, 1), 4).shape == (0, 3)
    full = pack_occupancy(np.ones((2, 3, 3)))
    assert eng.lfb(full, 3).tolist() == [[-1, -1, 0]] * 2
    mask, lfb = eng.valid_masks(full, 3, np.zeros((2, 4, 2), np.uint8))
    assert not mask.any()
    assert eng.game_ended(full, 3, np.zeros((2, 4, 2), np.uint8)).tolist() == [1.0, 1.0]
