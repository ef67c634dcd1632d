"""Drop-in game adapter (reference: alpha/Game.py, alpha/binpack/BinPackGame.py,
alpha/binpack/BinPackLogic.py).

A board is the reference's tuple ``(grid, tiles)``: a float H x W grid whose occupied
cells are 1, and the padded tile list of 2n ``[h, w]`` rows with ``[0, 0]`` pads at the
end.  Valid-move masks, the add-tile transition (place at the first free cell, remove
the tile and its rotation, compact, pad) and the terminal test run on the device
(bp_valid_masks / bp_transitions / bp_game_ended); the ``*_batch`` methods take many
boards per launch.
"""
import numpy as np

from . import runtime
from ._native import PLACED, pack_occupancy, pack_tiles, unpack_occupancy
from .data_generator import DataGenerator
from .solution_checker import SolutionChecker

ORIENTATIONS = 2


class Game(object):
    """Interface of alpha-zero-general's Game (alpha/Game.py:1-113)."""

    def getInitBoard(self):
        raise NotImplementedError

    def getBoardSize(self):
        raise NotImplementedError

    def getActionSize(self):
        raise NotImplementedError

    def getNextState(self, board, player, action):
        raise NotImplementedError

    def getValidMoves(self, board, player):
        raise NotImplementedError

    def getGameEnded(self, board, player):
        raise NotImplementedError

    def getCanonicalForm(self, board, player):
        raise NotImplementedError

    def getSymmetries(self, board, pi):
        raise NotImplementedError

    def stringRepresentation(self, board):
        raise NotImplementedError


def _tiles_array(tiles):
    return np.asarray([[int(t[0]), int(t[1])] for t in tiles], dtype=np.int64).reshape(-1, 2)


class Board(object):
    """BinPack board (BinPackLogic.py:14-117)."""

    ORIENTATIONS = 2

    def __init__(self, height, width, tiles, n_tiles, state=None, visualization_state=None,
                 device=None):
        self.height, self.width = height, width
        self.tiles = tiles
        self.orientations = ORIENTATIONS
        self.n_tiles = n_tiles
        self.device = device
        if state is None:
            self.state = (np.zeros([height, width]), self.tiles)
            self.vis_state = np.zeros([height, width])
        else:
            self.state = (state, self.tiles)
            self.vis_state = visualization_state
        assert self.state[0].shape == (self.height, self.width)

    def _packed(self):
        occ = pack_occupancy(np.asarray(self.state[0])[None], colorful=True)
        return occ, pack_tiles([_tiles_array(self.tiles).tolist()], ORIENTATIONS * self.n_tiles)

    def add_tile(self, position, player):
        """Place tiles[position] at the first free cell with value 1, drop the tile and its
        rotation, pad [0, 0] at the end (:46-69).  On failure the grid becomes None."""
        occ, table = self._packed()
        eng = runtime.get_engine(self.device)
        out_occ, out_tiles, verdict, lfb = eng.transitions(occ, self.width, table, [int(position)],
                                                           compact=True)
        if lfb[0][0] < 0:
            raise TypeError("'NoneType' object is not subscriptable")   # full board, as :48-52
        if not np.any(_tiles_array(self.tiles)[int(position)]):
            # a [0, 0] pad "fits" anywhere and removes two pads that are padded back (:49-63)
            self.tiles = [tuple(t) if t[0] or t[1] else [0, 0]
                          for t in _tiles_array(self.tiles).tolist()]
            self.state = (self.state[0], self.tiles)
            self.vis_state = self.state[0]
            return self.state, self.vis_state
        if verdict[0] == PLACED:
            grid = self.state[0]
            new_occ = unpack_occupancy(out_occ, self.width)[0]
            grid[(new_occ == 1) & (grid == 0)] = 1                       # in place, like :123
            self.tiles = [tuple(t) if t[0] or t[1] else [0, 0] for t in out_tiles[0].tolist()]
        else:
            grid = None
        self.state = (grid, self.tiles)
        self.vis_state = self.state[0]
        return self.state, self.vis_state

    def get_valid_moves(self):
        """0/1 per entry of the padded tile list (:72-82)."""
        return np.array(SolutionChecker.get_valid_tile_actions_indexes_given_grid(
            self.state[0], [list(t) for t in _tiles_array(self.tiles).tolist()]))

    def all_tiles_placed(self):
        return not _tiles_array(self.tiles).any()

    def get_win_state(self):
        """False while moves remain; 1 when the board is full; else filled / total (:90-101)."""
        occ, table = self._packed()
        v = float(runtime.get_engine(self.device).game_ended(occ, self.width, table)[0])
        if v != 0.0:
            return 1 if v == 1.0 else v
        if np.asarray(self.state[0]).any():
            return False
        # empty grid: an ended game is worth filled / total = 0.0
        if self.all_tiles_placed() or not np.any(self.get_valid_moves()):
            return 0.0
        return False

    def with_state(self, state, vis_state=None):
        """Copy of the board holding ``state`` = (grid, tiles) (:104-110)."""
        if state is None:
            state = self.state
        return Board(self.height, self.width, np.copy(state[1]), self.n_tiles, np.copy(state[0]),
                     vis_state, device=self.device)

    def __str__(self):
        out = ''
        for part in self.state:
            for row in part:
                out += str(row)
        return out


class BinPackGame(Game):
    """alpha/binpack/BinPackGame.py:11-88."""

    ORIENTATIONS = 2

    def __init__(self, height=None, width=None, n_tiles=None, device=None):
        self.device = device
        self.initialize(height, width, n_tiles)

    def initialize(self, height, width, n_tiles):
        self.height, self.width, self.n_tiles = height, width, n_tiles
        self.tiles, _ = DataGenerator().gen_instance(n_tiles, width, height)
        both = SolutionChecker.get_tiles_with_orientation(self.tiles)
        self._base_board = Board(height, width, both, n_tiles, device=self.device)

    def getInitBoard(self):
        self.initialize(self.height, self.width, self.n_tiles)
        return self._base_board.state, self._base_board.vis_state

    def getBoardSize(self):
        return (self.height, self.width, self.n_tiles * ORIENTATIONS + 1)

    def getActionSize(self):
        return self.n_tiles * self.ORIENTATIONS

    def getNextState(self, board, player, action, vis_state=None):
        """-> ((grid, tiles), player, vis_state); the input board is not modified (:51-55)."""
        b = self._base_board.with_state(state=board, vis_state=vis_state)
        b.add_tile(action, player)
        return b.state, player, b.vis_state

    def getValidMoves(self, board, player):
        return self._base_board.with_state(state=board).get_valid_moves()

    def getGameEnded(self, board, player):
        winstate = self._base_board.with_state(state=board).get_win_state()
        if winstate is False:
            return 0
        return winstate

    def getCanonicalForm(self, state, player):
        return state

    def getSymmetries(self, board, pi):
        return [(board, pi)]

    def stringRepresentation(self, board):
        b = self._base_board.with_state(state=board)
        return str(b), b

    # ---- batched forms -----------------------------------------------------------------
    def pack_boards(self, boards):
        grids = np.stack([np.asarray(b[0]) for b in boards])
        tiles = pack_tiles([_tiles_array(b[1]).tolist() for b in boards], self.getActionSize())
        return pack_occupancy(grids, colorful=True), tiles

    def getValidMoves_batch(self, boards, player=1):
        occ, tiles = self.pack_boards(boards)
        mask, _ = runtime.get_engine(self.device).valid_masks(occ, self.width, tiles)
        return mask

    def getGameEnded_batch(self, boards, player=1):
        occ, tiles = self.pack_boards(boards)
        return runtime.get_engine(self.device).game_ended(occ, self.width, tiles)

    def getNextState_batch(self, boards, player, actions):
        occ, tiles = self.pack_boards(boards)
        out_occ, out_tiles, verdict, _ = runtime.get_engine(self.device).transitions(
            occ, self.width, tiles, actions, compact=True)
        grids = unpack_occupancy(out_occ, self.width).astype(np.float64)
        return [((grids[i], out_tiles[i].astype(np.int64)) if verdict[i] == PLACED
                 else (None, boards[i][1])) for i in range(len(boards))], verdict


def display(board):
    print(" -----------------------")
    print(' '.join(map(str, range(len(board[0])))))
    print(board)
    print(" -----------------------")
