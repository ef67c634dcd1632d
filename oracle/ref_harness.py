"""Run the UNMODIFIED reference hot path from /root/reference, DB-free.

TEST INFRASTRUCTURE, build container only (the GPU box has no /root/reference).
Used by tests/golden/make_golden.py to write the golden fixtures and by the
``-m "not gpu"`` tests that are skipped when the reference tree is absent.

No reference file is edited.  What the harness does:
  * puts two stub modules in front of the reference on sys.path:
    ``matplotlib`` (plot imports at solution_checker.py:2, data_generator.py:1,7-8)
    and ``search`` (the f2py module of engine/search.f90, solution_checker.py:10);
  * occupancy fix: forces ``colorful_states=True`` in
    SolutionChecker.place_element_on_grid_given_grid (solution_checker.py:98), i.e.
    "a cell is occupied iff it is non-zero".  As shipped, CustomMCTS.predict
    (binpack/MCTS.py:59) calls get_next_turn with the default colorful_states=False
    while painting tiles with val = 1, 2, 3 ..., so later tiles can be laid over
    earlier ones;
  * RNG injection: replaces the ``random`` name inside binpack.MCTS (single call
    site, MCTS.py:174) with the Philox rollout stream of oracle/philox.py;
  * silences the print() calls inside the hot loops;
  * records, per depth, the score of every child (for the golden fixtures).
"""
import contextlib
import io
import os
import sys

from . import philox

REFERENCE_ROOT = os.environ.get("VISAPI_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")

_loaded = None


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "binpack", "MCTS.py"))


def load(occupancy_fix=True):
    """Import the reference modules; returns a namespace object."""
    global _loaded
    if _loaded is not None:
        _loaded.set_occupancy_fix(occupancy_fix)
        return _loaded
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    # final order: shims, engine/, reference root (so ``binpack`` is binpack/, not
    # alpha/binpack/), alpha/ (top-level ``Game``), as visapi/settings.py:31-33 does
    for p in (os.path.join(REFERENCE_ROOT, "alpha"), REFERENCE_ROOT,
              os.path.join(REFERENCE_ROOT, "engine"), _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    # a stale 'search'/'matplotlib' from elsewhere must not win
    for name in ("search",):
        sys.modules.pop(name, None)
    with contextlib.redirect_stdout(io.StringIO()):
        import solution_checker as sc_mod
        import state as state_mod
        import data_generator as dg_mod
        import binpack.MCTS as mcts_mod
        import binpack as _bp  # noqa: F401
        from alpha.binpack import BinPackGame as game_mod  # type: ignore
        from alpha.binpack import BinPackLogic as logic_mod  # type: ignore
    _loaded = _Ref(sc_mod, state_mod, dg_mod, mcts_mod, game_mod, logic_mod)
    _loaded.set_occupancy_fix(occupancy_fix)
    return _loaded


class _PhiloxRandom(object):
    """Stands in for the ``random`` module inside binpack.MCTS."""

    def __init__(self):
        self.stream = None

    def randint(self, a, b):
        return a + self.stream.draw(b - a + 1)


class _Ref(object):
    def __init__(self, sc_mod, state_mod, dg_mod, mcts_mod, game_mod, logic_mod):
        self.sc_mod = sc_mod
        self.SolutionChecker = sc_mod.SolutionChecker
        self.State = state_mod.State
        self.DataGenerator = dg_mod.DataGenerator
        self.mcts_mod = mcts_mod
        self.CustomMCTS = mcts_mod.CustomMCTS
        self.MCTS = mcts_mod.MCTS
        self.BinPackGame = game_mod.BinPackGame
        self.Board = logic_mod.Board
        self.ALL_TILES_USED = sc_mod.ALL_TILES_USED
        self.TILE_CANNOT_BE_PLACED = sc_mod.TILE_CANNOT_BE_PLACED
        self.NO_NEXT_POSITION_TILES_UNUSED = sc_mod.NO_NEXT_POSITION_TILES_UNUSED
        self._orig_place = sc_mod.SolutionChecker.__dict__[
            "place_element_on_grid_given_grid"].__func__
        self._orig_random = mcts_mod.random

    # ---- (i) occupancy fix -------------------------------------------------
    def set_occupancy_fix(self, on):
        orig = self._orig_place
        if on:
            def place(_bin, position, val, grid, cols, rows,
                      get_only_success=False, colorful_states=False):
                return orig(_bin, position, val, grid, cols, rows,
                            get_only_success=get_only_success, colorful_states=True)
            self.SolutionChecker.place_element_on_grid_given_grid = staticmethod(place)
        else:
            self.SolutionChecker.place_element_on_grid_given_grid = staticmethod(orig)

    # ---- flat search under the Philox stream -------------------------------
    def run_flat_search(self, tiles, board, n_sim, strategy="max_depth", seed=123,
                        stream=0, use_philox=True, py_seed=123):
        """CustomMCTS(tiles, board, strategy).predict(N=n_sim), as run_one_simulation
        does (run_mcts.py:187,217-224), with every rollout draw taken from Philox.

        Returns a dict with the reference's outputs plus per-depth child scores and
        rollout/placement counters.
        """
        import random as pyrandom
        ref = self
        mcts_mod = self.mcts_mod
        sc = self.SolutionChecker
        tiles = [tuple(int(v) for v in t) for t in tiles]
        n_entries0 = len(tiles)
        log = {"children": [], "rollouts": 0, "rollout_placements": 0}
        ctx = {"entry": -1, "state_id": None, "depth": 0, "sim": 0}
        shim = _PhiloxRandom()

        orig_next_turn = sc.__dict__["get_next_turn"].__func__
        orig_sims = mcts_mod.CustomMCTS.perform_simulations
        orig_sim = mcts_mod.CustomMCTS.perform_simulation

        def next_turn(state, tile, val=1, get_only_success=False, destroy_state=False,
                      colorful_states=False):
            if not destroy_state:           # only predict() calls it this way (MCTS.py:59)
                if ctx["state_id"] is state:
                    ctx["entry"] += 1
                else:
                    ctx["state_id"] = state
                    ctx["entry"] = 0
            return orig_next_turn(state, tile, val, get_only_success=get_only_success,
                                  destroy_state=destroy_state,
                                  colorful_states=colorful_states)

        def perform_simulations(self_, state, N=3000, record_simulations=False):
            ctx["depth"] = (n_entries0 - len(state.tiles)) // 2 - 1
            ctx["sim"] = 0
            res = orig_sims(self_, state, N=N, record_simulations=record_simulations)
            score = res[0]
            log["children"].append([ctx["depth"], ctx["entry"],
                                    "S" if score == ref.ALL_TILES_USED else float(score)])
            return res

        def perform_simulation(self_, state):
            shim.stream = philox.RolloutStream(
                seed, stream, philox.flat_tag(ctx["depth"], ctx["entry"]), ctx["sim"])
            ctx["sim"] += 1
            before = self_.n_tiles_placed
            res = orig_sim(self_, state)
            log["rollouts"] += 1
            log["rollout_placements"] += self_.n_tiles_placed - before
            return res

        sc.get_next_turn = staticmethod(next_turn)
        mcts_mod.CustomMCTS.perform_simulations = perform_simulations
        mcts_mod.CustomMCTS.perform_simulation = perform_simulation
        if use_philox:
            mcts_mod.random = shim
        else:
            pyrandom.seed(py_seed)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                m = mcts_mod.CustomMCTS(tiles, board, strategy=strategy)
                root, depth, found = m.predict(N=n_sim)
        finally:
            sc.get_next_turn = staticmethod(orig_next_turn)
            mcts_mod.CustomMCTS.perform_simulations = orig_sims
            mcts_mod.CustomMCTS.perform_simulation = orig_sim
            mcts_mod.random = self._orig_random
        n = n_entries0 // 2
        return {
            "depth": int(depth),
            "solution_found": bool(found),
            "score": 0 if found else n - int(depth),          # run_mcts.py:221-224
            "n_tiles_placed": int(m.n_tiles_placed),
            "solution_tiles_order": [[int(t[0]), int(t[1])] for t in m.solution_tiles_order],
            "children": log["children"],
            "rollouts": log["rollouts"],
            "rollout_placements": log["rollout_placements"],
            "root": root,
        }

    # ---- one rollout from an arbitrary state --------------------------------
    def run_rollout(self, tiles, board, seed, stream, tag, sim, val=1):
        """CustomMCTS.perform_simulation (MCTS.py:152-201) from (board, tiles)."""
        mcts_mod = self.mcts_mod
        shim = _PhiloxRandom()
        shim.stream = philox.RolloutStream(seed, stream, tag, sim)
        tiles = [tuple(int(v) for v in t) for t in tiles]
        mcts_mod.random = shim
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                m = mcts_mod.CustomMCTS(tiles, board)
                m.val = val
                st = self.State(board, tiles)
                depth, _root, order = m.perform_simulation(st)
        finally:
            mcts_mod.random = self._orig_random
        solved = depth == self.ALL_TILES_USED
        return {
            "solved": bool(solved),
            "depth": -1 if solved else int(depth),
            "placements": int(m.n_tiles_placed),
            "moves": [[int(t[0]), int(t[1])] for t in order],
            "final_board": (st.board != 0).astype("uint8") if False else None,
        }


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
