"""Literal known-answer vectors held by the reference's own unit tests
(engine/test_solution_checker.py, cited per block).  Data only."""

# :124-158  get_next_lfb_on_grid
LFB = [
    ([[1, 1], [0, 0]], (1, 0)),
    ([[1, 1, 0], [0, 0, 0]], (0, 2)),
    ([[1, 1, 1], [1, 1, 1]], None),
    ([[1, 1, 1], [1, 1, 1], [0, 1, 0], [0, 0, 0]], (2, 0)),
    ([[1, 1, 1, 1], [1, 1, 1, 1], [0, 0, 0, 1]], (2, 0)),
]

# :160-217  get_next_turn (board, tiles, tile) -> new board; copy vs destroy vs only-success
NEXT_TURN_BOARD = [[1, 1, 1, 1], [1, 1, 1, 0], [0, 0, 0, 0]]
NEXT_TURN_TILES = [(2, 1), (1, 2)]
NEXT_TURN_TILE = [2, 1]
NEXT_TURN_AFTER = [[1, 1, 1, 1], [1, 1, 1, 1], [0, 0, 0, 1]]

# :219-302  get_valid_next_moves (board, unsorted tile list -> sorted by (x[1], x[0]), expected)
B1 = [[1, 1, 1, 1], [1, 1, 1, 0], [0, 0, 0, 0]]
B3 = [[1, 1, 1, 1], [1, 0, 0, 1], [1, 0, 0, 1], [1, 0, 0, 1]]
B4 = [[1, 1, 1, 1], [1, 0, 1, 0], [1, 0, 1, 0], [1, 0, 1, 0]]
B5 = [[1, 1, 1, 1], [1, 0, 0, 0], [1, 0, 0, 0], [1, 0, 0, 0]]
VALID_MOVES = [
    (B1, [(2, 1), (1, 2)], False, [(2, 1)]),
    (B1, [(2, 1), (1, 2), (3, 3), (4, 4), (1, 1)], True, [(1, 1), (2, 1)]),
    (B3, [(2, 1), (1, 2), (3, 3), (3, 2), (1, 1), (2, 3)], True,
     [(1, 1), (2, 1), (1, 2), (3, 2)]),
    (B3, [(2, 1), (1, 2), (3, 3), (3, 2), (1, 1), (6, 1)], True,
     [(1, 1), (2, 1), (1, 2), (3, 2)]),
    (B4, [(2, 1), (1, 2), (3, 3), (3, 2), (1, 1), (6, 1), (3, 1)], True,
     [(1, 1), (2, 1), (3, 1)]),
    (B5, [(2, 3), (2, 4), (3, 3)], True, [(2, 3), (3, 3)]),
]

# :305-332  eliminate_pair_tiles (list, tile, expected | ValueError)
ELIMINATE = [
    ([(1, 1), (2, 1), (1, 1), (1, 2)], (1, 1), [(2, 1), (1, 2)]),
    ([(8, 1), (2, 1), (1, 1), (1, 8)], (1, 8), [(2, 1), (1, 1)]),
    ([(1, 1), (2, 1), (1, 1), (1, 1), (2, 1), (1, 1)], (1, 1),
     [(2, 1), (1, 1), (2, 1), (1, 1)]),
    ([(1, 8), (2, 1), (1, 1), (1, 1), (2, 1), (1, 1)], (1, 8), ValueError),
]

# :335-372  get_possible_tile_actions_given_grid / get_valid_tile_actions_indexes_given_grid
ACTION_GRID = [[1, 1, 1, 0, 0], [1, 1, 1, 0, 0], [1, 1, 1, 0, 0]]
POSSIBLE_TILES = [[1, 2], [2, 1], [4, 5], [1, 1], [1, 1], [3, 2], [3, 3], [2, 3]]
POSSIBLE_EXPECTED = [[1, 2], [2, 1], [1, 1], [1, 1], [3, 2]]
MASK_TILES = [[1, 2], [2, 1], [0, 0], [4, 5], [1, 1], [1, 1], [3, 2], [3, 3], [2, 3]]
MASK_EXPECTED = [1, 1, 0, 0, 1, 1, 1, 0, 0]
PAD_CASE = ([[1, 2], [2, 1], [4, 5]], 2, [[1, 2], [2, 1], [4, 5], [0, 0], [0, 0]])


def sort_tiles(tiles):
    return sorted(tiles, key=lambda x: (x[1], x[0]))
