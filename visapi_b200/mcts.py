"""Drop-in search drivers (reference: binpack/MCTS.py).

``CustomMCTS`` — the flat greedy Monte-Carlo search of binpack/MCTS.py:31-201.  The
whole ``predict`` loop (every depth, every child, N rollouts per child, first-max
commit, early exit at the first solved rollout) runs on the device in one
``bp_flat_search`` call; the host only rebuilds the reference's ``State`` tree from the
per-depth child scores the device returns.  ``predict_many`` runs any number of
independent problems in one batch (the fast path).

Randomness: the reference draws from CPython's Mersenne Twister seeded with 123 per
problem (run_mcts.py:187).  Here every rollout (depth d, entry e, simulation n) owns
the Philox4x32-10 stream key=(seed, stream), counter=(n, d<<16|e, block, 0), so results
do not depend on how rollouts are scheduled or on the number of GPUs.

``MCTS`` (PUCT, binpack/MCTS.py:204-335) lives in visapi_b200/puct.py.
"""
import math

import numpy as np

from . import runtime
from ._native import AVG_DEPTH, MAX_DEPTH, pack_occupancy, pack_tiles
from .solution_checker import ALL_TILES_USED, SolutionChecker, get_cols, get_rows
from .state import State

ORIENTATIONS = 2
EPS = 0.1


def get_max_index(_list):
    """Index of the first element with the strictly greatest score; None entries skipped
    (binpack/MCTS.py:20-29)."""
    max_index, max_value = 0, -math.inf
    for i, el in enumerate(_list):
        if not el:
            continue
        if el.score > max_value:
            max_value, max_index = el.score, i
    return max_index


def _as_tile_tuples(tiles):
    return [(int(t[0]), int(t[1])) for t in tiles]


def _first_free(board):
    free = np.flatnonzero(np.asarray(board).ravel() == 0)
    if free.size == 0:
        return None
    return divmod(int(free[0]), board.shape[1])


class CustomMCTS(object):
    def __init__(self, tiles, board, strategy='max_depth', seed=123, stream=0, device=None):
        self.initial_tiles, self.initial_board = tiles, board
        self.state = State(self.initial_board, self.initial_tiles)
        self.solution_checker = SolutionChecker(len(tiles), get_rows(board), get_cols(board))
        self.strategy = strategy
        self.n_tiles_placed = 0
        self.solution_tiles_order = []
        self.colorful_states = True
        self.seed, self.stream, self.device = int(seed), int(stream), device
        self.val = 1
        # device-side counters of the last predict()
        self.rollouts = 0
        self.rollout_placements = 0
        self.rollouts_executed = 0
        self._sim_tag = 0

    # ------------------------------------------------------------------ a12 predict
    def predict(self, temp=1, N=3000):
        """-> (root State, depth, solution_found); sets n_tiles_placed and
        solution_tiles_order as the reference does WITH ITS OCCUPANCY BUG FIXED: a cell is occupied
        iff it is non-zero.  The reference's predict() paints committed tiles with 1, 2, 3, ... but
        tests placements with colorful_states=False ("== 1" only, binpack/MCTS.py:59 ->
        engine/solution_checker.py:117-119), so from depth 2 on the as-shipped code can lay a tile
        over cells painted >= 2; depth, score and n_tiles_placed then differ from this class on
        such instances (tests/golden/as_shipped.json)."""
        board = np.asarray(self.state.board)
        tiles = _as_tile_tuples(self.state.tiles)
        eng = runtime.get_engine(self.device)
        strat = MAX_DEPTH if self.strategy == 'max_depth' else AVG_DEPTH
        res = eng.flat_search(
            [board.shape[0]], [board.shape[1]], pack_tiles([tiles], len(tiles) + len(tiles) % 2),
            [len(tiles)], int(N), strat, self.seed, [self.stream],
            # an empty board needs no upload and lets the engine take its lane-per-rollout kernels
            boards=pack_occupancy(board[None], colorful=True) if board.any() else None,
            want_scores=True)
        depth = int(res["depth"][0])
        found = bool(res["solution_found"][0])
        self.n_tiles_placed = int(res["n_tiles_placed"][0])
        self.rollouts = int(res["rollouts"][0])
        self.rollout_placements = int(res["rollout_placements"][0])
        self.rollouts_executed = int(res["rollouts_executed"][0])
        order = res["order"][0, :int(res["n_order"][0])].tolist()
        self.solution_tiles_order = [(int(t[0]), int(t[1])) for t in order]
        self._rebuild_tree(res["child_scores"][0], res["path"][0], depth, found, int(N))
        return self.state, depth, found

    def _rebuild_tree(self, child_scores, path, depth, found, n_sim):
        """Children of every visited depth with their rollout scores; the committed
        child of each depth leads to the next one (binpack/MCTS.py:58-119)."""
        state = self.state
        val = 1
        for d in range(child_scores.shape[0]):
            tiles = state.tiles
            scores = child_scores[d][:len(tiles)]
            if not len(tiles) or not (scores != -1).any():
                break
            pos = _first_free(state.board)
            by_entry = {}
            for i, tile in enumerate(tiles):
                sc = int(scores[i])
                if sc == -1:
                    continue
                new_board = np.copy(state.board)
                new_board[pos[0]: pos[0] + tile[0], pos[1]: pos[1] + tile[1]] = val
                child = State(board=new_board,
                              tiles=SolutionChecker.eliminate_pair_tiles(tiles, tile), parent=state)
                state.children.append(child)
                by_entry[i] = child
                if sc == -2:                       # a rollout from this child used every tile
                    child.score = len(tiles) / ORIENTATIONS - 1
                    self._attach_solution_chain(child, tile, val)
                else:
                    child.score = sc if self.strategy == 'max_depth' else np.float64(sc) / n_sim
                    child.tile_placed = tile
            if d >= depth or int(path[d]) not in by_entry:
                break
            state = by_entry[int(path[d])]
            val += 1
        self.val = val

    def _attach_solution_chain(self, child, tile, val):
        """The solved rollout as a chain of single children (MCTS.py:82-87,137,195-199)."""
        order = self.solution_tiles_order
        # order = [committed tiles ..., tile, rollout moves ...]; the rollout part is what
        # remains after this child's own tile
        n_rollout = len(child.tiles) // ORIENTATIONS
        moves = order[len(order) - n_rollout:] if n_rollout else []
        node = child.copy()
        child.children = [node]
        for mv in moves:
            val += 1
            # the rollout plays with destroy_state=True (MCTS.py:175-178): the state it moves FROM
            # is painted in place, and its successor starts as a copy of that board -- so in the
            # rendered chain every node already shows the tile that leads to its child
            pos = _first_free(node.board)
            node.board[pos[0]: pos[0] + mv[0], pos[1]: pos[1] + mv[1]] = val
            nxt = State(board=node.board, tiles=SolutionChecker.eliminate_pair_tiles(node.tiles, mv),
                        parent=node)
            node.score = len(node.tiles) / ORIENTATIONS - 1
            node.children.append(nxt)
            nxt.score = -1
            node = nxt

    # ------------------------------------------------------------- a10 / a11 rollouts
    def perform_simulations(self, state, N=3000, record_simulations=False):
        """N rollouts from ``state`` on the device; stops (logically) at the first solved
        one -> (ALL_TILES_USED | max | average depth, move list) (MCTS.py:126-149)."""
        tiles = _as_tile_tuples(state.tiles)
        if len(tiles) == 0:
            return ALL_TILES_USED, []
        eng = runtime.get_engine(self.device)
        board = np.asarray(state.board)
        occ = pack_occupancy(board[None], colorful=True)
        table = pack_tiles([tiles], len(tiles) + len(tiles) % 2)
        tag = self._sim_tag
        self._sim_tag += 1
        steps, solved, _ = eng.rollouts(occ, board.shape[1], table, int(N), seed=self.seed,
                                        streams=[self.stream], tags=[tag])
        steps, solved = steps[0].astype(np.int64), solved[0]
        hit = np.flatnonzero(solved)
        last = int(hit[0]) if hit.size else int(N) - 1
        self.n_tiles_placed += int(steps[:last + 1].sum())
        _, _, moves = eng.rollouts(occ, board.shape[1], table, 1, seed=self.seed,
                                   streams=[self.stream], tags=[tag], sim_begin=last,
                                   want_moves=True)
        order = [(int(t[0]), int(t[1])) for t in moves[0, 0, :int(steps[last])].tolist()]
        if hit.size:
            return ALL_TILES_USED, order
        if self.strategy == 'max_depth':
            return np.max(steps), order
        return np.average(steps), order

    def perform_simulation(self, state):
        """One rollout -> (depth | ALL_TILES_USED, root state, move list) (MCTS.py:152-201)."""
        tiles = _as_tile_tuples(state.tiles)
        if len(tiles) == 0:
            return ALL_TILES_USED, state, []
        eng = runtime.get_engine(self.device)
        board = np.asarray(state.board)
        tag = self._sim_tag
        self._sim_tag += 1
        steps, solved, moves = eng.rollouts(
            pack_occupancy(board[None], colorful=True), board.shape[1],
            pack_tiles([tiles], len(tiles) + len(tiles) % 2), 1, seed=self.seed,
            streams=[self.stream], tags=[tag], want_moves=True)
        n = int(steps[0, 0])
        self.n_tiles_placed += n
        order = [(int(t[0]), int(t[1])) for t in moves[0, 0, :n].tolist()]
        return (ALL_TILES_USED if solved[0, 0] else n), state, order

    # ------------------------------------------------------------------ batched form
    @staticmethod
    def predict_many(problems, N=1000, strategy='max_depth', seed=123, streams=None, device=None,
                     want_scores=False):
        """problems: list of (tiles, board) pairs (boards must be empty) or of
        visapi_b200.datasets.Problem.  One device batch; returns a list of dicts with the
        reference's per-problem outputs."""
        from . import datasets
        eng = runtime.get_engine(device)
        rows, cols, n_entries, tables = [], [], [], []
        for p in problems:
            if isinstance(p, datasets.Problem):
                ent, r, c = p.entries(), p.rows, p.cols
            else:
                ent, r, c = _as_tile_tuples(p[0]), p[1].shape[0], p[1].shape[1]
                if np.asarray(p[1]).any():
                    raise ValueError("predict_many takes empty boards; use CustomMCTS for others")
            rows.append(r)
            cols.append(c)
            n_entries.append(len(ent))
            tables.append(ent)
        es = max(n_entries) + max(n_entries) % 2
        if streams is None:
            streams = np.arange(len(problems), dtype=np.uint32)
        res = eng.flat_search(rows, cols, pack_tiles(tables, es), n_entries, int(N),
                              MAX_DEPTH if strategy == 'max_depth' else AVG_DEPTH, seed, streams,
                              want_scores=want_scores)
        out = []
        for i in range(len(problems)):
            n = n_entries[i] // ORIENTATIONS
            found = bool(res["solution_found"][i])
            out.append({
                "depth": int(res["depth"][i]), "solution_found": found,
                "score": 0 if found else n - int(res["depth"][i]),
                "n_tiles_placed": int(res["n_tiles_placed"][i]),
                "solution_tiles_order": [(int(t[0]), int(t[1])) for t in
                                         res["order"][i, :int(res["n_order"][i])].tolist()],
                "rollouts": int(res["rollouts"][i]),
                "rollout_placements": int(res["rollout_placements"][i]),
                "rollouts_executed": int(res["rollouts_executed"][i]),
                "path": [int(v) for v in res["path"][i] if v >= 0],
                "child_scores": res["child_scores"][i] if want_scores else None,
            })
        return out
