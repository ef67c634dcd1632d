"""Pins both CPU oracles (numpy and C restatements) to the reference:
  * the literal vectors of engine/test_solution_checker.py (tests/reference_vectors.py);
  * fixtures produced by the reference itself (tests/golden/make_golden.py);
  * Random123 known-answer vectors for the shared Philox stream.
CPU only.
"""
import numpy as np
import pytest

from oracle import c_oracle, philox, py_oracle
from tests import reference_vectors as rv
from tests.conftest import board_from_hex, hex_rows_of

ORACLES = [pytest.param(py_oracle, id="py"), pytest.param(c_oracle, id="c")]

# Random123 kat_vectors, philox4x32 10 rounds
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers():
    for ctr, key, want in PHILOX_KAT:
        assert philox.philox4x32_10(ctr, key) == want
        assert c_oracle.philox(ctr, key) == want


def test_philox_py_c_agree_random():
    rng = np.random.default_rng(0)
    for _ in range(200):
        ctr = tuple(int(v) for v in rng.integers(0, 2**32, 4))
        key = tuple(int(v) for v in rng.integers(0, 2**32, 2))
        assert philox.philox4x32_10(ctr, key) == c_oracle.philox(ctr, key)


# ------------------------------------------------------------ reference unit vectors
@pytest.mark.parametrize("orc", ORACLES)
def test_ref_lfb_vectors(orc):
    for board, want in rv.LFB:
        assert orc.next_lfb(np.array(board)) == want


@pytest.mark.parametrize("orc", ORACLES)
def test_ref_valid_next_moves_vectors(orc):
    for board, tiles, do_sort, want in rv.VALID_MOVES:
        tl = rv.sort_tiles(tiles) if do_sort else tiles
        # the reference tests call it with the default colorful_states=False
        assert orc.valid_next_moves(np.array(board), tl, colorful=False) == want
        assert orc.valid_next_moves(np.array(board), tl, colorful=True) == want


@pytest.mark.parametrize("orc", ORACLES)
def test_ref_eliminate_pair_vectors(orc):
    for tiles, tile, want in rv.ELIMINATE:
        if want is ValueError:
            with pytest.raises(ValueError):
                orc.eliminate_pair(list(tiles), tile)
        else:
            assert orc.eliminate_pair(list(tiles), tile) == want


def test_ref_next_turn_vectors_py():
    board = np.array(rv.NEXT_TURN_BOARD)
    keep = board.copy()
    ok, nb = py_oracle.next_turn(board, rv.NEXT_TURN_TILES, rv.NEXT_TURN_TILE, val=1,
                                 destroy=False, colorful=False)
    assert ok is True
    np.testing.assert_array_equal(nb, np.array(rv.NEXT_TURN_AFTER))
    np.testing.assert_array_equal(board, keep)
    ok, nb = py_oracle.next_turn(board, rv.NEXT_TURN_TILES, rv.NEXT_TURN_TILE, val=1,
                                 destroy=True, colorful=False)
    assert ok is True and nb is board
    board = np.array(rv.NEXT_TURN_BOARD)
    ok, nb = py_oracle.next_turn(board, rv.NEXT_TURN_TILES, rv.NEXT_TURN_TILE, val=1,
                                 get_only_success=True, destroy=True, colorful=False)
    assert ok is True and nb is None


def test_ref_next_turn_vectors_c():
    v, nb = c_oracle.next_turn(np.array(rv.NEXT_TURN_BOARD), 2, rv.NEXT_TURN_TILE, val=1,
                               colorful=False)
    assert v == c_oracle.PLACED
    np.testing.assert_array_equal(nb, np.array(rv.NEXT_TURN_AFTER))


@pytest.mark.parametrize("orc", ORACLES)
def test_ref_action_mask_vectors(orc):
    assert orc.valid_action_mask(np.array(rv.ACTION_GRID), rv.MASK_TILES) == rv.MASK_EXPECTED


def test_ref_possible_actions_py():
    got = py_oracle.possible_tile_actions(np.array(rv.ACTION_GRID), rv.POSSIBLE_TILES)
    assert got == rv.POSSIBLE_EXPECTED


def test_find_first():
    hay = np.array([3.0, 0.0, 1.0, 0.0])
    for orc in (py_oracle, c_oracle):
        assert orc.find_first(0, hay) == 1
        assert orc.find_first(1, hay) == 2
        assert orc.find_first(7, hay) == -1
        assert orc.find_first(0, np.array([])) == -1


# ------------------------------------------------------------ reference-generated fixtures
@pytest.mark.parametrize("orc", ORACLES)
def test_transitions_golden(orc, golden):
    sentinel = {py_oracle.ALL_TILES_USED: c_oracle.ALL_TILES_USED,
                py_oracle.TILE_CANNOT_BE_PLACED: c_oracle.TILE_CANNOT_BE_PLACED,
                py_oracle.NO_NEXT_POSITION_TILES_UNUSED: c_oracle.NO_NEXT_POSITION_TILES_UNUSED}
    walks = golden("transitions.json")
    if orc is py_oracle:
        walks = walks[::3]           # the numpy restatement is slow; every third walk
    for w in walks:
        rows, cols = w["rows"], w["cols"]
        grid = np.zeros((rows, cols))
        tiles = [tuple(t) for t in w["tiles"]]
        for st in w["steps"]:
            lfb = orc.next_lfb(grid)
            if st["lfb"] is None:
                assert lfb is None
                if orc is py_oracle:
                    v, _ = orc.next_turn(grid, tiles, tiles[0] if tiles else (1, 1))
                    assert v == st["verdict"]
                else:
                    v, _ = orc.next_turn(grid, len(tiles), tiles[0] if tiles else (1, 1))
                    assert v == sentinel[st["verdict"]]
                break
            assert list(lfb) == st["lfb"]
            assert orc.next_occupied_col(grid, lfb, True) == st["next_occupied_col"]
            legal = orc.valid_next_moves(grid, tiles, True)
            assert legal == [t for t, f in zip(tiles, st["legal"]) if f]
            assert orc.valid_action_mask((grid != 0).astype(np.float64), tiles) == st["legal"]
            if "bad_tile" in st:
                if orc is py_oracle:
                    v, nb = orc.next_turn(grid, tiles, tuple(st["bad_tile"]))
                    assert v == st["bad_verdict"] and nb is None
                else:
                    v, _ = orc.next_turn(grid, len(tiles), tuple(st["bad_tile"]))
                    assert v == sentinel[st["bad_verdict"]]
            if "pick" not in st:
                assert not legal
                break
            tile = legal[st["pick"]]
            assert list(tile) == st["tile"]
            if orc is py_oracle:
                v, grid = orc.next_turn(grid, tiles, tile, val=2)
                assert v is True
            else:
                v, grid = orc.next_turn(grid, len(tiles), tile, val=2)
                assert v == c_oracle.PLACED
            assert hex_rows_of(grid) == st["board_after"]
            tiles = orc.eliminate_pair(tiles, tile)


def test_rollouts_golden(golden):
    for case in golden("rollouts.json"):
        rows, cols = case["rows"], case["cols"]
        tiles = [tuple(t) for t in case["tiles"]]
        grid = board_from_hex(rows, cols, case.get("board"))
        c = c_oracle.rollout(grid, tiles, case["seed"], case["stream"], case["tag"], case["sim"])
        assert c["solved"] == case["solved"]
        assert c["depth"] == case["depth"]
        assert c["placements"] == case["placements"]
        assert c["moves"] == case["moves"]
        if case["sim"] == 0:
            rs = philox.RolloutStream(case["seed"], case["stream"], case["tag"], case["sim"])
            solved, depth, pl, moves = py_oracle.rollout(grid.copy(), tiles, rs)
            assert solved == case["solved"] and pl == case["placements"]
            assert (-1 if solved else depth) == case["depth"]
            assert [list(m) for m in moves] == case["moves"]


FLAT_KEYS = ["depth", "solution_found", "score", "n_tiles_placed", "solution_tiles_order",
             "children", "rollouts", "rollout_placements"]


def test_flat_search_golden_c(golden):
    for case in golden("flat_search.json"):
        board = np.zeros((case["rows"], case["cols"]))
        got = c_oracle.flat_search(case["tiles"], board, case["n_sim"], case["strategy"],
                                   case["seed"], case["stream"])
        for k in FLAT_KEYS:
            assert got[k] == case[k], (case["name"], case["strategy"], k)


def test_flat_search_golden_py(golden):
    cases = [c for c in golden("flat_search.json") if c["n_tiles_placed"] < 16000][:6]
    assert cases
    for case in cases:
        board = np.zeros((case["rows"], case["cols"]))
        got = py_oracle.flat_search(case["tiles"], board, case["n_sim"], case["strategy"],
                                    case["seed"], case["stream"])
        for k in FLAT_KEYS:
            assert got[k] == case[k], (case["name"], case["strategy"], k)


def test_flat_search_batch_matches_single(golden):
    cases = [c for c in golden("flat_search.json") if len(c["tiles"]) == 40][:8]
    E = 40
    rows = [c["rows"] for c in cases]
    cols = [c["cols"] for c in cases]
    tiles = np.array([c["tiles"] for c in cases], np.int32)
    streams = [c["stream"] for c in cases]
    got = c_oracle.flat_search_batch(rows, cols, tiles, [E] * len(cases), 5, "max_depth", 123,
                                     streams, n_threads=2)
    for c, g in zip(cases, got):
        one = c_oracle.flat_search(c["tiles"], np.zeros((c["rows"], c["cols"])), 5, "max_depth",
                                   123, c["stream"])
        for k in ("depth", "solution_found", "n_tiles_placed", "solution_tiles_order", "rollouts"):
            assert g[k] == one[k]


@pytest.mark.parametrize("impl", ["py", "c"])
def test_alpha_game_golden(impl, golden):
    for game in golden("alpha_game.json"):
        h, w, n = game["height"], game["width"], game["n_tiles"]
        assert game["action_size"] == 2 * n and game["board_size"] == [h, w, 2 * n + 1]
        grid = np.zeros((h, w))
        tiles = [list(t) for t in game["init_tiles"]]
        for st in game["steps"]:
            assert tiles == st["tiles"]
            assert hex_rows_of(grid) == st["board"]
            if impl == "py":
                b = py_oracle.Board(h, w, tiles, n, grid.copy())
                valid = b.valid_moves().tolist()
                ws = b.win_state()
            else:
                ws = c_oracle.board_win_state(grid, tiles)
                full = not (grid == 0).any()
                valid = [0] * len(tiles) if full else c_oracle.valid_action_mask(grid, tiles)
            ended = 0 if ws is False else ws
            assert ended == st["ended"]
            if st["ended"] == 0 or impl == "py":
                assert valid == st["valid"]
            if "action" not in st:
                break
            if impl == "py":
                b.add_tile(st["action"])
                grid, tiles = b.grid, [list(t) for t in b.tiles]
            else:
                rc, g2, t2 = c_oracle.board_add_tile(grid, tiles, st["action"])
                assert rc == 1
                grid, tiles = g2.astype(np.float64), t2.tolist()


def _stub_predict(action_size):
    def predict(grid, tiles):
        filled = int(np.count_nonzero(np.asarray(grid)))
        cells = int(np.asarray(grid).size)
        pi = np.array([((a * 2654435761 + filled * 40503 + 12345) % 1000 + 1)
                       for a in range(action_size)], dtype=np.float64)
        return pi / pi.sum(), 0.5 * filled / cells
    return predict


def test_puct_golden_py(golden):
    """a15/a16 had no reference test; pinned here against the reference run under a stub net."""
    for game in golden("puct.json"):
        h, w, n = game["height"], game["width"], game["n_tiles"]
        base = py_oracle.Board(h, w, game["init_tiles"], n)
        for mv in game["moves"]:
            state = (board_from_hex(h, w, mv["board"]), np.array(mv["tiles"]))
            tree = py_oracle.PuctTree(base, _stub_predict(2 * n), game["num_sims"], game["cpuct"])
            probs = tree.action_prob(state, temp=1)
            assert probs == mv["probs"]
            assert tree.leaf_evals == mv["leaf_evals"]
            assert len(tree.Ps) == mv["n_nodes"]
            tree0 = py_oracle.PuctTree(base, _stub_predict(2 * n), game["num_sims"],
                                       game["cpuct"])
            assert tree0.action_prob(state, temp=0) == mv["probs_temp0"]


def test_published_aggregates_present(golden):
    pub = golden("published.json")
    assert pub["g"]["max_depth/1000"]["n"] == 1000
    assert pub["g"]["max_depth/1000"]["solved"] == 403
    assert pub["g"]["avg_depth/1000"]["solved"] == 461
    assert pub["b"]["avg_depth/5000"]["solved"] == 55


def test_as_shipped_reference_versus_occupancy_fix(golden):
    """parity is pinned to the reference WITH its occupancy bug fixed (cells painted
    2, 3, ... count as occupied inside predict()).  tests/golden/as_shipped.json holds the truly
    unmodified reference next to the fixed one on the same instances and Philox streams: the
    oracle must reproduce the FIXED runs exactly, and the fixture measures the divergence (3 of 24
    small instances, always the as-shipped code laying a tile over painted cells and "solving")."""
    from oracle import c_oracle
    data = golden("as_shipped.json")
    differ = 0
    for inst in data["instances"]:
        tiles = [tuple(t) for t in inst["tiles"]]
        board = np.zeros((inst["rows"], inst["cols"]))
        got = c_oracle.flat_search(tiles, board, inst["n_sim"], "max_depth", 123, inst["stream"])
        for k in ("depth", "solution_found", "n_tiles_placed", "rollouts"):
            assert got[k] == inst["fixed"][k], (inst["stream"], k)
        same = all(inst["fixed"][k] == inst["as_shipped"][k] for k in inst["fixed"])
        assert same == inst["same"]
        differ += 0 if same else 1
        if not same:   # the as-shipped run never does worse: overlapping placements only add "fits"
            assert inst["as_shipped"]["depth"] >= inst["fixed"]["depth"] or inst["as_shipped"]["solution_found"]
    assert differ == data["n_differ"] and 0 < differ < len(data["instances"]) // 2


def test_coach_episode_fixture_against_oracle(golden):
    """alpha/Coach.py:25-63 as run by the reference (tests/golden/coach_episode.json) against the
    oracle's PUCT tree and Board: same number of moves, same policies bit for bit, same reward,
    same number of net evaluations."""
    from oracle import py_oracle

    def stub(grid, _tiles, size):
        filled, cells = int(np.count_nonzero(np.asarray(grid))), int(np.asarray(grid).size)
        pi = np.array([((a * 2654435761 + filled * 40503 + 12345) % 1000 + 1) for a in range(size)],
                      dtype=np.float64)
        return pi / pi.sum(), 0.5 * filled / cells
    for case in golden("coach_episode.json"):
        h, w, n = case["height"], case["width"], case["n_tiles"]
        base = py_oracle.Board(h, w, case["init_tiles"], n)
        tree = py_oracle.PuctTree(base, lambda g, t: stub(g, t, 2 * n), case["num_sims"], case["cpuct"])
        np.random.seed(case["np_seed"])
        state = (np.zeros((h, w)), np.array(case["init_tiles"]))
        step, stream = 0, []
        while True:
            step += 1
            pi = tree.action_prob(state, temp=int(step < case["tempThreshold"]))
            stream.append((state, pi))
            action = np.random.choice(len(pi), p=pi)
            nxt = base.with_state(state).add_tile(action)
            state = (nxt[0], np.array(nxt[1]))
            reward = tree.game_ended(base.with_state(state))
            if reward != 0:
                break
        assert len(stream) == len(case["examples"]) and tree.leaf_evals == case["leaf_evals"]
        for (st, pi), want in zip(stream, case["examples"]):
            assert [float(p) for p in pi] == want["pi"]
            assert [[int(t[0]), int(t[1])] for t in st[1]] == want["tiles"]
            assert float(reward) == want["reward"]
