"""Drop-in ``SolutionChecker`` namespace (reference: engine/solution_checker.py).

Same static-method names, arguments, return shapes and error behaviour as the
reference, so CustomMCTS-style callers and the reference's own unit tests run against
it unchanged.  The scans and fit tests (first free cell, free run, legal-move masks)
run on the device through the C ABI (bp_lfb / bp_valid_masks); painting a numpy board
with ``val`` and editing Python tile lists stay on the host, where those objects live.
Single-state calls are here for compatibility; the batched forms at the bottom are
the fast path.
"""
import numpy as np

from . import runtime
from ._native import pack_occupancy, pack_tiles

ALL_TILES_USED = 'ALL_TILES_USED'
TILE_CANNOT_BE_PLACED = 'TILE_CANNOT_BE_PLACED'
NO_NEXT_POSITION_TILES_UNUSED = 'NO_NEXT_POSITION_TILES_UNUSED'

GLOBAL_OCCUPIED_VAL = 1


def get_cols(board):
    return board.shape[1]


def get_rows(board):
    return board.shape[0]


def _free_grid(grid):
    """get_next_lfb_on_grid looks for cells equal to 0 (solution_checker.py:86)."""
    return pack_occupancy(np.asarray(grid)[None], colorful=True)


def _tile_table(tiles):
    """(h, w) lists -> uint8 table; entries that do not fit a byte can never be placed
    on a <= 128 x 128 board and are sent as 255 x 255."""
    a = np.asarray([(int(t[0]), int(t[1])) for t in tiles], dtype=np.int64).reshape(-1, 2)
    a = np.where((a < 0) | (a > 255), 255, a)
    return pack_tiles([a.tolist()])


class SolutionChecker(object):
    """Static methods only on the hot path; the constructor keeps the reference's fields."""

    def __init__(self, n, cols, rows):
        self.n = n
        self.cols = cols
        self.rows = rows
        self.grid = np.array([[0 for _ in range(self.cols)] for _ in range(self.rows)])

    # ---- a2 -------------------------------------------------------------------------
    @staticmethod
    def get_next_lfb_on_grid(grid):
        """Row-major first cell equal to 0 -> (row, col), or None (:82-92)."""
        grid = np.asarray(grid)
        r, c, _ = runtime.get_engine().lfb(_free_grid(grid), grid.shape[1])[0]
        if r < 0:
            return None
        return (int(r), int(c))

    def get_next_lfb(self):
        return SolutionChecker.get_next_lfb_on_grid(self.grid)

    # ---- a3 -------------------------------------------------------------------------
    @staticmethod
    def get_next_occupied_col(board, next_position, colorful_states=False):
        """First occupied column at/after next_position in its row, else n_cols (:290-305)."""
        board = np.asarray(board)
        r, c = int(next_position[0]), int(next_position[1])
        row = np.array(board[r:r + 1], dtype=np.float64)
        occ = (row != 0) if colorful_states else (row == GLOBAL_OCCUPIED_VAL)
        occ = occ.astype(np.uint8)
        occ[0, :c] = 1                                   # only the tail counts
        _, lc, run = runtime.get_engine().lfb(pack_occupancy(occ[None]), board.shape[1])[0]
        if lc != c:                                      # the cell itself is occupied
            return c
        return int(c + run)

    # ---- a4 -------------------------------------------------------------------------
    @staticmethod
    def get_valid_next_moves(state, tiles, val=1, colorful_states=False):
        """Entries that fit at the LFB, in list order, duplicates kept (:307-317)."""
        board = np.asarray(state.board)
        if len(tiles) == 0:
            if SolutionChecker.get_next_lfb_on_grid(board) is None:
                raise TypeError("'NoneType' object is not subscriptable")
            return []
        # the LFB is found on cells == 0; the run ends at the first occupied cell
        free = _free_grid(board)
        eng = runtime.get_engine()
        r, c, _ = eng.lfb(free, board.shape[1])[0]
        if r < 0:
            raise TypeError("'NoneType' object is not subscriptable")
        if colorful_states:
            mask, _ = eng.valid_masks(free, board.shape[1], _tile_table(tiles))
            flags = mask[0]
        else:
            nxt = SolutionChecker.get_next_occupied_col(board, (r, c), colorful_states=False)
            room = nxt - c
            flags = [t[1] <= room and t[0] + r <= board.shape[0] for t in tiles]
        return [t for t, f in zip(tiles, flags) if f]

    # ---- a5 -------------------------------------------------------------------------
    @staticmethod
    def place_element_on_grid_given_grid(_bin, position, val, grid, cols, rows,
                                         get_only_success=False, colorful_states=False):
        """Bounds, first-footprint-row occupancy test, in-place fill with val (:97-132)."""
        if position[1] + _bin[1] > cols:
            return False, None
        if position[0] + _bin[0] > rows:
            return False, None
        r, c = int(position[0]), int(position[1])
        h, w = int(_bin[0]), int(_bin[1])
        nxt = SolutionChecker.get_next_occupied_col(grid, (r, c), colorful_states=colorful_states)
        if nxt - c < w:
            return False, None
        if get_only_success:
            return True, None
        grid[r: r + h, c: c + w] = val
        return True, grid

    def place_element_on_grid(self, _bin, position, val, cols, rows):
        return SolutionChecker.place_element_on_grid_given_grid(
            _bin, position, val, self.grid, self.cols, self.rows)

    # ---- a6 -------------------------------------------------------------------------
    @staticmethod
    def get_next_turn(state, tile, val=1, get_only_success=False, destroy_state=False,
                      colorful_states=False):
        """(True, board) | (sentinel, None) (:261-287)."""
        next_position = SolutionChecker.get_next_lfb_on_grid(state.board)
        if not next_position and len(state.tiles) == 0:
            return ALL_TILES_USED, None
        elif not next_position:
            return NO_NEXT_POSITION_TILES_UNUSED, None
        board = state.board if destroy_state else np.copy(state.board)
        success, new_board = SolutionChecker.place_element_on_grid_given_grid(
            tile, next_position, val, board, get_cols(state.board), get_rows(state.board),
            get_only_success=get_only_success, colorful_states=colorful_states)
        if not success:
            return TILE_CANNOT_BE_PLACED, None
        return True, new_board

    # ---- a7 -------------------------------------------------------------------------
    @staticmethod
    def eliminate_pair_tiles(tiles, tile_to_remove):
        """Remove the first ``tile`` and then the first rotated ``tile`` (:319-332);
        ValueError when either is absent, as list.index raises in the reference."""
        index = tiles.index(tile_to_remove)
        new_tiles = tiles[:index] + tiles[index + 1:]
        rotated_index = new_tiles.index((tile_to_remove[1], tile_to_remove[0]))
        return new_tiles[:rotated_index] + new_tiles[rotated_index + 1:]

    # ---- a8 -------------------------------------------------------------------------
    @staticmethod
    def get_valid_tile_actions_indexes_given_grid(grid, tiles):
        """0/1 per entry of the padded list; [0, 0] pads are 0 (:359-381, val=1 rule)."""
        grid = np.asarray(grid)
        eng = runtime.get_engine()
        r, c, _ = eng.lfb(_free_grid(grid), grid.shape[1])[0]
        if r < 0:
            raise TypeError("'NoneType' object is not subscriptable")
        # occupied iff == 1, but the LFB is the first cell == 0: mark every cell before the
        # LFB as occupied so that the device LFB of the "== 1" board is the same cell
        occ = (grid == GLOBAL_OCCUPIED_VAL).astype(np.uint8)
        flat = occ.reshape(-1)
        flat[: r * grid.shape[1] + c] = 1
        mask, _ = eng.valid_masks(pack_occupancy(occ[None]), grid.shape[1], _tile_table(tiles))
        return [int(v) for v in mask[0]]

    @staticmethod
    def get_possible_tile_actions_given_grid(grid, tiles, pad_with_zeros=False):
        """Tiles that can be placed at the LFB (:335-357)."""
        n = len(tiles)
        mask = SolutionChecker.get_valid_tile_actions_indexes_given_grid(grid, tiles)
        # unlike the index mask, a [0, 0] entry counts as placeable here (:345-352)
        new_tiles = [t for t, m in zip(tiles, mask) if m or (t[0] == 0 and t[1] == 0)]
        if pad_with_zeros:
            new_tiles = SolutionChecker.pad_tiles_with_zero_scalars(new_tiles, n - len(new_tiles))
        return new_tiles

    # ---- list helpers (host) ----------------------------------------------------------
    @staticmethod
    def pad_tiles_with_zero_scalars(tiles, n_zero_tiles_to_add):
        new_tiles = tiles[:]
        for _ in range(n_zero_tiles_to_add):
            new_tiles.append([0, 0])
        return new_tiles

    @staticmethod
    def tiles_to_np_array(tiles):
        return np.array([np.array(x) for x in tiles])

    @staticmethod
    def np_array_to_tiles(tiles):
        return [list(x) for x in tiles]

    @staticmethod
    def get_n_nonplaced_tiles(_tiles_ints):
        tiles = SolutionChecker.np_array_to_tiles(_tiles_ints)
        return len([x for x in tiles if list(x) != [0, 0]])

    @staticmethod
    def get_tiles_with_orientation(tiles):
        if isinstance(tiles, np.ndarray):
            tiles = tiles.tolist()
        out = tiles[:]
        for tile in tiles:
            out.append((tile[1], tile[0]))
        return out

    # ---- batched forms: n states of one shape in one launch --------------------------
    @staticmethod
    def batch_valid_masks(grids, tile_lists, colorful_states=True):
        """grids [n, rows, cols], tile_lists n lists -> (mask uint8 [n, E], lfb int32 [n, 3])."""
        grids = np.asarray(grids)
        return runtime.get_engine().valid_masks(pack_occupancy(grids, colorful_states),
                                                grids.shape[2], pack_tiles(tile_lists))

    @staticmethod
    def batch_next_states(grids, tile_lists, actions, compact=True, colorful_states=True):
        """Place entry actions[i] at the LFB of state i; returns (occupancy words, tile
        tables, verdict codes, lfb); see bp_transitions in include/rectpack_b200.h."""
        grids = np.asarray(grids)
        return runtime.get_engine().transitions(pack_occupancy(grids, colorful_states),
                                                grids.shape[2], pack_tiles(tile_lists), actions,
                                                compact=compact)
