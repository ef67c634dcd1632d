"""numpy restatement of the reference's Python hot path.  TEST INFRASTRUCTURE.

See oracle/__init__.py for who may import this.  Each function cites the reference
lines it follows (paths relative to /root/reference).  The restatement keeps the
reference's data types on purpose (float64 numpy boards, Python lists of (h, w)
tuples, whole-board scans and copies) so that timing it is a fair stand-in for the
reference's own CPU implementation; ``oracle/rectpack_oracle.c`` is the fast one.

Occupancy: ``colorful=True`` means "occupied iff non-zero" (the rollout path,
binpack/MCTS.py:42,170,176-177, and the canonical semantics of SURVEY.md fact 2);
``colorful=False`` means "occupied iff == 1" (solution_checker.py:117-119).

Parity status: PINNED by tests/test_oracle_golden.py (reference unit-test vectors +
fixtures produced by the reference itself, tests/golden/make_golden.py).
"""
import math

import numpy as np

from . import philox

ALL_TILES_USED = "ALL_TILES_USED"
TILE_CANNOT_BE_PLACED = "TILE_CANNOT_BE_PLACED"
NO_NEXT_POSITION_TILES_UNUSED = "NO_NEXT_POSITION_TILES_UNUSED"
OCCUPIED = 1            # GLOBAL_OCCUPIED_VAL, solution_checker.py:20
EPS = 0.1               # binpack/MCTS.py:6


# --------------------------------------------------------------------------- a1
def find_first(needle, haystack):
    """engine/search.f90:1-16 — first 0-based k with int32(haystack[k]) == needle, else -1."""
    hay = np.asarray(haystack).astype(np.int32, copy=False).ravel()
    hits = np.flatnonzero(hay == np.int32(needle))
    return int(hits[0]) if hits.size else -1


# --------------------------------------------------------------------------- a2
def next_lfb(grid):
    """solution_checker.py:82-92 — row-major first cell equal to 0, or None."""
    k = find_first(0, grid.ravel())
    if k < 0:
        return None
    return (k // grid.shape[1], k % grid.shape[1])


# --------------------------------------------------------------------------- a3
def next_occupied_col(grid, pos, colorful=True):
    """solution_checker.py:290-305 — first occupied column at/after ``pos`` in its row."""
    tail = grid[pos[0], pos[1]:]
    if colorful:
        nz = np.flatnonzero(tail != 0)
        return grid.shape[1] if nz.size == 0 else pos[1] + int(nz[0])
    k = find_first(OCCUPIED, tail)
    return grid.shape[1] if k < 0 else pos[1] + k


# --------------------------------------------------------------------------- a4
def valid_next_moves(grid, tiles, colorful=True):
    """solution_checker.py:307-317 — entries that fit at the LFB, list order, duplicates kept.

    Raises TypeError on a full board (LFB is None), as the reference does (:310-312).
    """
    lfb = next_lfb(grid)
    room = next_occupied_col(grid, lfb, colorful) - lfb[1]
    rows = grid.shape[0]
    return [t for t in tiles if t[1] <= room and t[0] + lfb[0] <= rows]


# --------------------------------------------------------------------------- a5
def place(tile, pos, val, grid, get_only_success=False, colorful=True):
    """solution_checker.py:97-132 — bounds, first-footprint-row occupancy test, in-place fill."""
    rows, cols = grid.shape
    if pos[1] + tile[1] > cols or pos[0] + tile[0] > rows:
        return False, None
    first_row = grid[pos[0], pos[1]: pos[1] + tile[1]]
    blocked = np.any(first_row) if colorful else (OCCUPIED in first_row)
    if blocked:
        return False, None
    if get_only_success:
        return True, None
    grid[pos[0]: pos[0] + tile[0], pos[1]: pos[1] + tile[1]] = val
    return True, grid


# --------------------------------------------------------------------------- a6
def next_turn(grid, tiles, tile, val=1, get_only_success=False, destroy=False, colorful=True):
    """solution_checker.py:261-287 — (True | sentinel, board | None)."""
    pos = next_lfb(grid)
    if pos is None:
        if len(tiles) == 0:
            return ALL_TILES_USED, None
        return NO_NEXT_POSITION_TILES_UNUSED, None
    target = grid if destroy else np.copy(grid)
    ok, out = place(tile, pos, val, target, get_only_success=get_only_success, colorful=colorful)
    if not ok:
        return TILE_CANNOT_BE_PLACED, None
    return True, out


# --------------------------------------------------------------------------- a7
def eliminate_pair(tiles, tile):
    """solution_checker.py:319-332 — drop first ``tile`` then first rotated ``tile``."""
    i = tiles.index(tile)                       # ValueError if absent
    rest = tiles[:i] + tiles[i + 1:]
    j = rest.index((tile[1], tile[0]))          # ValueError if the pair is missing
    return rest[:j] + rest[j + 1:]


# --------------------------------------------------------------------------- a8
def valid_action_mask(grid, tiles):
    """solution_checker.py:359-381 — 0/1 per entry of the padded tile list (val=1 semantics)."""
    lfb = next_lfb(grid)
    out = []
    for t in tiles:
        if t[0] == 0 and t[1] == 0:
            out.append(0)
            continue
        ok, _ = place(t, lfb, 1, grid, get_only_success=True, colorful=False)
        out.append(1 if ok else 0)
    return out


def possible_tile_actions(grid, tiles):
    """solution_checker.py:335-357 (pad_with_zeros=False)."""
    lfb = next_lfb(grid)
    return [t for t in tiles
            if place(t, lfb, 1, grid, get_only_success=True, colorful=False)[0]]


def tiles_with_orientation(tiles):
    """solution_checker.py:411-418 — [t1..tn, rot(t1)..rot(tn)], unsorted."""
    base = [tuple(int(v) for v in t) for t in tiles]
    return base + [(t[1], t[0]) for t in base]


def sorted_both_orientations(tiles):
    """data_generator.py:86-98 — both orientations, sorted by (x[1], x[0])."""
    out = []
    for t in tiles:
        out.append((int(t[0]), int(t[1])))
        out.append((int(t[1]), int(t[0])))
    return sorted(out, key=lambda x: (x[1], x[0]))


# -------------------------------------------------------------------------- a10
def rollout(grid, tiles, stream, val=1):
    """binpack/MCTS.py:152-201 — one uniform-random playout; mutates ``grid``.

    ``stream.draw(K)`` replaces random.randint(0, K-1) (:174); exactly one draw per
    step, also when K == 1.  Returns (solved, depth, placements, moves).
    """
    moves = []
    depth = 0
    placements = 0
    tiles = list(tiles)
    if len(tiles) == 0:
        return True, 0, 0, moves
    while True:
        val += 1
        if len(tiles) == 0:
            return True, depth, placements, moves
        legal = valid_next_moves(grid, tiles, colorful=True)
        if not legal:
            return False, depth, placements, moves
        pick = legal[stream.draw(len(legal))]
        verdict, _ = next_turn(grid, tiles, pick, val, destroy=True, colorful=True)
        placements += 1
        moves.append(pick)
        if verdict is not True:                  # unreachable on area-exact instances
            return False, depth, placements, moves
        tiles = eliminate_pair(tiles, pick)
        grid = np.copy(grid)                     # State(board=...) copies, state.py:66
        depth += 1


# -------------------------------------------------------------------------- a11
def simulations(grid, tiles, n_sim, strategy, seed, stream_id, tag, val=1, sim_begin=0):
    """binpack/MCTS.py:126-149 — N rollouts, stop at the first solved one.

    Returns (score | ALL_TILES_USED, moves_of_last_rollout, rollouts_run, placements).
    """
    depths = []
    placements = 0
    moves = []
    for n in range(sim_begin, sim_begin + n_sim):
        rs = philox.RolloutStream(seed, stream_id, tag, n)
        solved, depth, pl, moves = rollout(np.copy(grid), tiles, rs, val)
        placements += pl
        if solved:
            return ALL_TILES_USED, moves, n - sim_begin + 1, placements
        depths.append(depth)
    if strategy == "max_depth":
        score = np.max(np.array(depths))
    else:
        score = np.average(np.array(depths))
    return score, moves, n_sim, placements


# -------------------------------------------------------------------------- a12
def flat_search(tiles, board, n_sim, strategy="max_depth", seed=123, stream_id=0):
    """binpack/MCTS.py:44-123 (+ get_max_index :20-29, run_mcts.py:221-224).

    Occupancy is "non-zero" throughout.  Returns the reference's
    outputs plus per-depth child scores.
    """
    tiles = [tuple(int(v) for v in t) for t in tiles]
    n = len(tiles) // 2
    grid = np.array(board, dtype=np.float64)
    out = {"depth": 0, "solution_found": False, "n_tiles_placed": 0,
           "solution_tiles_order": [], "children": [], "rollouts": 0,
           "rollout_placements": 0, "path": []}
    order = out["solution_tiles_order"]
    prev_tile = None                 # state.tile_placed of the current state
    depth = 0
    val = 1
    while len(tiles):
        scores = []                  # one per entry; None when it cannot be placed
        children = []
        placed_any = False
        for i, tile in enumerate(tiles):
            verdict, child = next_turn(grid, tiles, tile, val, destroy=False, colorful=True)
            out["n_tiles_placed"] += 1
            if verdict == ALL_TILES_USED:
                out["solution_found"] = True
                out["depth"] = depth
                return _finish(out, n)
            if verdict == TILE_CANNOT_BE_PLACED:
                scores.append(None)
                children.append(None)
                continue
            placed_any = True
            rest = eliminate_pair(tiles, tile)
            score, sim_moves, ran, pl = simulations(
                child, rest, n_sim, strategy, seed, stream_id,
                philox.flat_tag(depth, i), val)
            out["rollouts"] += ran
            out["rollout_placements"] += pl
            out["n_tiles_placed"] += pl
            if score == ALL_TILES_USED:
                out["children"].append([depth, i, "S"])
                out["solution_found"] = True
                out["depth"] = depth
                if prev_tile:
                    order.extend([prev_tile, tile] + sim_moves)
                else:
                    order.extend([tile] + sim_moves)
                return _finish(out, n)
            out["children"].append([depth, i, float(score)])
            scores.append(score)
            children.append((child, rest, tile))
        if not placed_any:
            out["depth"] = depth
            return _finish(out, n)
        val += 1
        depth += 1
        best, best_score = 0, -math.inf          # get_max_index: strict >, first wins
        for i, s in enumerate(scores):
            if s is None:
                continue
            if s > best_score:
                best_score, best = s, i
        if prev_tile:
            order.append(prev_tile)
        grid, tiles, prev_tile = children[best]
        out["path"].append(best)
    out["solution_found"] = True
    out["depth"] = depth
    return _finish(out, n)


def _finish(out, n):
    out["score"] = 0 if out["solution_found"] else n - out["depth"]
    out["solution_tiles_order"] = [[int(t[0]), int(t[1])] for t in out["solution_tiles_order"]]
    return out


# -------------------------------------------------------------------------- a13
class Board(object):
    """alpha/binpack/BinPackLogic.py:14-117 restated over (grid, padded tile list)."""

    def __init__(self, height, width, tiles, n_tiles, grid=None):
        self.height, self.width, self.n_tiles = height, width, n_tiles
        self.tiles = [[int(t[0]), int(t[1])] for t in tiles]
        self.grid = np.zeros([height, width]) if grid is None else grid

    def with_state(self, state):
        """:104-110 — copies grid and tiles."""
        return Board(self.height, self.width, np.copy(state[1]).tolist(), self.n_tiles,
                     np.copy(state[0]))

    def add_tile(self, action):
        """:46-69 — place tiles[action] at the LFB with val 1, remove the pair, pad [0,0]."""
        lfb = next_lfb(self.grid)
        ok, grid = place(self.tiles[action], lfb, 1, self.grid, colorful=False)
        if ok:
            tl = [tuple(t) for t in self.tiles]
            tl = eliminate_pair(tl, tuple(self.tiles[action]))
            tl = [list(t) for t in tl]
            tl += [[0, 0]] * (2 * self.n_tiles - len(tl))
            self.tiles = tl
        self.grid = grid
        return (self.grid, self.tiles)

    def valid_moves(self):
        """:72-82."""
        return np.array(valid_action_mask(self.grid, self.tiles))

    def win_state(self):
        """:90-101 — False while running, 1 when full, else filled/total."""
        done = all(t[0] == 0 and t[1] == 0 for t in self.tiles)
        if done or not np.any(self.valid_moves()):
            filled = np.sum(self.grid)
            total = self.grid.shape[0] * self.grid.shape[1]
            return 1 if filled == total else filled / total
        return False

    def key(self):
        """:112-117 — concatenated str(row) of grid rows then tile rows."""
        s = ""
        for row in self.grid:
            s += str(row)
        for row in np.array(self.tiles):
            s += str(row)
        return s


# ---------------------------------------------------------------------- a15/a16
class PuctTree(object):
    """binpack/MCTS.py:204-335 restated: dict tree keyed by Board.key().

    ``predict(grid, tiles) -> (pi, v)`` stands in for nnet.predict.  Single-player:
    the value is not negated on the way up (:335).
    """

    def __init__(self, base_board, predict, num_sims, cpuct):
        self.base, self.predict = base_board, predict
        self.num_sims, self.cpuct = num_sims, cpuct
        self.Qsa, self.Nsa, self.Ns, self.Ps, self.Es, self.Vs = {}, {}, {}, {}, {}, {}
        self.leaf_evals = 0

    def action_size(self):
        return 2 * self.base.n_tiles

    def game_ended(self, b):
        w = b.win_state()
        return 0 if w is False else w             # BinPackGame.py:61-68

    def search(self, state):
        b = self.base.with_state(state)
        s = b.key()
        if s not in self.Es:
            self.Es[s] = self.game_ended(b)
        if self.Es[s] != 0:
            return self.Es[s]
        if s not in self.Ps:
            pi, v = self.predict(state[0], state[1])
            self.leaf_evals += 1
            valids = b.valid_moves()
            p = np.asarray(pi, dtype=np.float64) * valids
            tot = np.sum(p)
            if tot > 0:
                p = p / tot
            else:
                p = p + valids
                p = p / np.sum(p)
            self.Ps[s], self.Vs[s], self.Ns[s] = p, valids, 0
            return v
        valids = self.Vs[s]
        best, best_a = -float("inf"), -1
        for a in range(self.action_size()):
            if valids[a]:
                if (s, a) in self.Qsa:
                    u = self.Qsa[(s, a)] + self.cpuct * self.Ps[s][a] * \
                        math.sqrt(self.Ns[s]) / (1 + self.Nsa[(s, a)])
                else:
                    u = self.cpuct * self.Ps[s][a] * math.sqrt(self.Ns[s] + EPS) + 100
                if u > best:
                    best, best_a = u, a
        a = best_a
        nb = self.base.with_state(state)
        nxt = nb.add_tile(a)
        v = self.search(nxt)
        if (s, a) in self.Qsa:
            self.Qsa[(s, a)] = (self.Nsa[(s, a)] * self.Qsa[(s, a)] + v) / (self.Nsa[(s, a)] + 1)
            self.Nsa[(s, a)] += 1
        else:
            self.Qsa[(s, a)] = v
            self.Nsa[(s, a)] = 1
        self.Ns[s] += 1
        return v

    def root_stats(self, state):
        """Visit counts Nsa and value sums Wsa = Qsa * Nsa of the root's actions: what the ranks
        of a root-parallel PUCT search sum before :221-244 forms the policy."""
        s = self.base.with_state(state).key()
        counts = [self.Nsa.get((s, a), 0) for a in range(self.action_size())]
        wsa = [self.Qsa.get((s, a), 0.0) * self.Nsa.get((s, a), 0) for a in range(self.action_size())]
        return counts, wsa

    def action_prob(self, state, temp=1):
        """:221-244."""
        for _ in range(self.num_sims):
            self.search(state)
        s = self.base.with_state(state).key()
        counts = [self.Nsa.get((s, a), 0) for a in range(self.action_size())]
        if temp == 0:
            best = int(np.argmax(counts))
            probs = [0] * len(counts)
            probs[best] = 1
            return probs
        counts = [x ** (1.0 / temp) for x in counts]
        tot = float(sum(counts))
        return [x / tot for x in counts]
