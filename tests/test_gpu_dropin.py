"""The reference-facing Python surface (SolutionChecker, State, CustomMCTS, BinPackGame,
MCTS) against the reference's own unit-test vectors and the reference-made fixtures.
Everything here runs the CUDA path: marked gpu."""
import os

import numpy as np
import pytest

from tests import reference_vectors as rv
from tests.conftest import board_from_hex, hex_rows_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def SC():
    from visapi_b200.solution_checker import SolutionChecker
    return SolutionChecker


def _state(board, tiles):
    from visapi_b200.state import State
    return State(np.array(board), tiles)


# ---- the reference's engine/test_solution_checker.py cases, same calls ----------------
def test_get_next_lfb_on_grid(SC):
    for board, want in rv.LFB:
        assert SC.get_next_lfb_on_grid(np.array(board)) == want


def test_get_next_turn_copy_destroy_only_success(SC):
    st = _state(rv.NEXT_TURN_BOARD, rv.NEXT_TURN_TILES)
    success, new_board = SC.get_next_turn(st, rv.NEXT_TURN_TILE, val=1, get_only_success=False,
                                          destroy_state=False)
    assert success is True
    np.testing.assert_array_equal(new_board, np.array(rv.NEXT_TURN_AFTER))
    np.testing.assert_array_equal(st.board, np.array(rv.NEXT_TURN_BOARD))
    success, new_board = SC.get_next_turn(st, rv.NEXT_TURN_TILE, val=1, destroy_state=True)
    assert success is True
    np.testing.assert_array_equal(st.board, new_board)
    st = _state(rv.NEXT_TURN_BOARD, rv.NEXT_TURN_TILES)
    success, new_board = SC.get_next_turn(st, rv.NEXT_TURN_TILE, val=1, get_only_success=True,
                                          destroy_state=True)
    assert success is True and new_board is None


def test_get_next_turn_sentinels(SC):
    from visapi_b200 import solution_checker as m
    full = np.ones((2, 2))
    assert SC.get_next_turn(_state(full, []), (1, 1))[0] == m.ALL_TILES_USED
    assert SC.get_next_turn(_state(full, [(1, 1), (1, 1)]), (1, 1))[0] == m.NO_NEXT_POSITION_TILES_UNUSED
    st = _state(rv.NEXT_TURN_BOARD, rv.NEXT_TURN_TILES)
    assert SC.get_next_turn(st, (1, 2)) == (m.TILE_CANNOT_BE_PLACED, None)
    assert SC.get_next_turn(st, (3, 1)) == (m.TILE_CANNOT_BE_PLACED, None)


def test_get_valid_next_moves(SC):
    for board, tiles, do_sort, want in rv.VALID_MOVES:
        tl = rv.sort_tiles(tiles) if do_sort else tiles
        st = _state(board, tl)
        assert SC.get_valid_next_moves(st, tl) == want
        assert SC.get_valid_next_moves(st, tl, colorful_states=True) == want
    with pytest.raises(TypeError):
        SC.get_valid_next_moves(_state(np.ones((2, 2)), [(1, 1)]), [(1, 1)])


def test_colorful_vs_plain_occupancy(SC):
    """a cell painted 2 is free for the '== 1' rule and occupied for the colorful rule"""
    board = np.array([[1, 1, 1, 1], [0, 0, 2, 0], [0, 0, 0, 0]])
    tiles = [(1, 1), (1, 2), (1, 3), (1, 4)]
    st = _state(board, tiles)
    assert SC.get_valid_next_moves(st, tiles, colorful_states=True) == [(1, 1), (1, 2)]
    assert SC.get_valid_next_moves(st, tiles, colorful_states=False) == tiles
    assert SC.get_next_occupied_col(board, (1, 0), colorful_states=True) == 2
    assert SC.get_next_occupied_col(board, (1, 0), colorful_states=False) == 4
    ok, _ = SC.place_element_on_grid_given_grid((1, 3), (1, 0), 5, board.copy(), 4, 3)
    assert ok is True
    ok, g = SC.place_element_on_grid_given_grid((1, 3), (1, 0), 5, board.copy(), 4, 3,
                                                colorful_states=True)
    assert ok is False and g is None


def test_eliminate_pair_tiles(SC):
    for tiles, tile, want in rv.ELIMINATE:
        if want is ValueError:
            with pytest.raises(ValueError):
                SC.eliminate_pair_tiles(list(tiles), tile_to_remove=tile)
        else:
            assert SC.eliminate_pair_tiles(list(tiles), tile_to_remove=tile) == want


def test_possible_actions_and_mask(SC):
    grid = np.array(rv.ACTION_GRID)
    assert SC.get_possible_tile_actions_given_grid(grid, rv.POSSIBLE_TILES) == rv.POSSIBLE_EXPECTED
    assert SC.get_valid_tile_actions_indexes_given_grid(grid, rv.MASK_TILES) == rv.MASK_EXPECTED
    tiles, n, want = rv.PAD_CASE
    assert SC.pad_tiles_with_zero_scalars(tiles, n) == want


def test_find_first_module():
    from visapi_b200 import search
    assert search.find_first(0, np.array([1.0, 1.0, 0.0, 0.0])) == 2
    assert search.find_first(1, np.zeros(5)) == -1


# ---- CustomMCTS ------------------------------------------------------------------------
def test_custom_mcts_predict_golden(golden):
    from visapi_b200.mcts import CustomMCTS
    from visapi_b200.verify import replay_solution
    cases = golden("flat_search.json")
    picked = [c for c in cases if c["solution_found"]][:3] + cases[:3] + \
             [c for c in cases if len(c["tiles"]) > 60][:1]
    for c in picked:
        board = np.zeros((c["rows"], c["cols"]))
        tiles = [tuple(t) for t in c["tiles"]]
        m = CustomMCTS(tiles, board, strategy=c["strategy"], seed=c["seed"], stream=c["stream"])
        root, depth, found = m.predict(N=c["n_sim"])
        assert depth == c["depth"] and found == c["solution_found"]
        assert m.n_tiles_placed == c["n_tiles_placed"]
        assert [list(t) for t in m.solution_tiles_order] == c["solution_tiles_order"]
        assert root is m.state and not board.any()
        # the tree: one child per evaluated entry of every depth, with the golden scores
        node, want = root, {}
        for d, e, sc in c["children"]:
            want.setdefault(d, []).append(sc)
        for d in sorted(want):
            got = [ch.score for ch in node.children]
            exp = [(len(node.tiles) / 2 - 1) if s == "S" else s for s in want[d]]
            assert got == exp, (c["name"], d)
            if d >= depth:
                break
            # the committed child is the first one with the greatest score (get_max_index)
            scores = [ch.score for ch in node.children]
            node = node.children[scores.index(max(scores))]
        if found:
            base = [tuple(t) for t in c["tiles"]]
            half = []
            pool = list(base)
            while pool:
                t = pool.pop(0)
                pool.remove((t[1], t[0]))
                half.append(t)
            ok, why = replay_solution(c["rows"], c["cols"], half, m.solution_tiles_order)
            assert ok, why


def test_custom_mcts_rollout_api(golden):
    from oracle import c_oracle
    from visapi_b200.mcts import CustomMCTS
    from visapi_b200.state import State
    c = golden("flat_search.json")[0]
    tiles = [tuple(t) for t in c["tiles"]]
    board = np.zeros((c["rows"], c["cols"]))
    m = CustomMCTS(tiles, board, seed=5, stream=9)
    depth, root, order = m.perform_simulation(State(board, tiles))
    want = c_oracle.rollout(board, tiles, 5, 9, 0, 0)
    assert (depth == "ALL_TILES_USED") == want["solved"]
    assert [list(t) for t in order] == want["moves"]
    assert m.n_tiles_placed == want["placements"]
    score, order = m.perform_simulations(State(board, tiles), N=40)
    depths = [c_oracle.rollout(board, tiles, 5, 9, 1, s) for s in range(40)]
    if not any(d["solved"] for d in depths):
        assert score == max(d["steps"] for d in depths)


def test_predict_many_matches_single(golden):
    from visapi_b200.mcts import CustomMCTS
    cases = [c for c in golden("flat_search.json") if c["n_sim"] == 8 and c["strategy"] == "max_depth"
             and len(c["tiles"]) == 40]
    probs = [([tuple(t) for t in c["tiles"]], np.zeros((c["rows"], c["cols"]))) for c in cases]
    out = CustomMCTS.predict_many(probs, N=8, streams=[c["stream"] for c in cases])
    for c, o in zip(cases, out):
        assert o["depth"] == c["depth"] and o["n_tiles_placed"] == c["n_tiles_placed"]
        assert [list(t) for t in o["solution_tiles_order"]] == c["solution_tiles_order"]


# ---- alpha path ------------------------------------------------------------------------
def test_binpack_game_golden(golden):
    from visapi_b200.game import BinPackGame
    for g in golden("alpha_game.json"):
        h, w, n = g["height"], g["width"], g["n_tiles"]
        game = BinPackGame(h, w, n)
        assert game.getActionSize() == g["action_size"]
        assert list(game.getBoardSize()) == g["board_size"]
        board = (np.zeros([h, w]), np.array(g["init_tiles"]))
        for st in g["steps"]:
            assert hex_rows_of(board[0]) == st["board"]
            assert [[int(t[0]), int(t[1])] for t in board[1]] == st["tiles"]
            ended = game.getGameEnded(board, 1)
            assert ended == st["ended"]
            if st["ended"] == 0:
                assert game.getValidMoves(board, 1).tolist() == st["valid"]
            if "action" not in st:
                break
            keep = board[0].copy()
            nxt, player, _ = game.getNextState(board, 1, st["action"])
            np.testing.assert_array_equal(board[0], keep)       # input untouched
            board = (nxt[0], np.array([[int(t[0]), int(t[1])] for t in nxt[1]]))


def test_binpack_game_batched(golden):
    from visapi_b200.game import BinPackGame
    g = golden("alpha_game.json")[1]
    h, w, n = g["height"], g["width"], g["n_tiles"]
    game = BinPackGame(h, w, n)
    boards, valid, ended, actions, after = [], [], [], [], []
    steps = [s for s in g["steps"] if "action" in s]
    for s in steps:
        boards.append((board_from_hex(h, w, s["board"]), np.array(s["tiles"])))
        valid.append(s["valid"])
        actions.append(s["action"])
    assert game.getValidMoves_batch(boards).tolist() == valid
    assert not game.getGameEnded_batch(boards).any()
    nxt, verdict = game.getNextState_batch(boards, 1, actions)
    assert (verdict == 1).all()
    for k, s in enumerate(steps[:-1]):
        assert hex_rows_of(nxt[k][0]) == g["steps"][k + 1]["board"]
        assert nxt[k][1].tolist() == g["steps"][k + 1]["tiles"]


def _stub_predict(action_size, calls):
    class Net(object):
        def predict(self, grid):
            calls.append(1)
            filled = int(np.count_nonzero(np.asarray(grid)))
            cells = int(np.asarray(grid).size)
            pi = np.array([((a * 2654435761 + filled * 40503 + 12345) % 1000 + 1)
                           for a in range(action_size)], dtype=np.float64)
            return pi / pi.sum(), 0.5 * filled / cells
    return Net()


def test_puct_mcts_golden(golden):
    """MCTS.getActionProb bit-exact against the reference run under the same stub net."""
    from visapi_b200.game import BinPackGame
    from visapi_b200.puct import MCTS
    for g in golden("puct.json"):
        h, w, n = g["height"], g["width"], g["n_tiles"]
        game = BinPackGame(h, w, n)
        for mv in g["moves"]:
            board = (board_from_hex(h, w, mv["board"]), np.array(mv["tiles"]))
            calls = []
            m = MCTS(game, _stub_predict(2 * n, calls), {"numMCTSSims": g["num_sims"], "cpuct": g["cpuct"]})
            probs = m.getActionProb(board, temp=1)
            assert probs == mv["probs"]
            assert len(calls) == mv["leaf_evals"]
            m0 = MCTS(game, _stub_predict(2 * n, []), {"numMCTSSims": g["num_sims"], "cpuct": g["cpuct"]})
            assert m0.getActionProb(board, temp=0) == mv["probs_temp0"]


def test_batched_mcts_equals_single(golden):
    from visapi_b200.puct import BatchedMCTS
    g = golden("puct.json")[1]
    h, w, n = g["height"], g["width"], g["n_tiles"]
    moves = g["moves"]
    B = len(moves)

    def predict_batch(grids):
        pis, vs = [], []
        for grid in grids:
            filled = int(np.count_nonzero(grid))
            pi = np.array([((a * 2654435761 + filled * 40503 + 12345) % 1000 + 1)
                           for a in range(2 * n)], dtype=np.float64)
            pis.append(pi / pi.sum())
            vs.append(0.5 * filled / grid.size)
        return np.array(pis), np.array(vs)

    bm = BatchedMCTS(h, w, 2 * n, predict_batch, g["num_sims"], g["cpuct"], B)
    bm.set_roots(np.stack([board_from_hex(h, w, mv["board"]) for mv in moves]),
                 np.array([mv["tiles"] for mv in moves]))
    bm.simulate()
    probs = bm.action_probs(temp=1)
    for k, mv in enumerate(moves):
        assert probs[k].tolist() == mv["probs"]
    assert bm.leaf_batches <= g["num_sims"]


def test_coach_episode_and_batched_selfplay(tmp_path):
    """f2: an episode through the reference-shaped Coach and 8 lock-step games through
    BatchedSelfPlay, both on the device tree with a PyTorch net; every example is labelled
    with its episode's reward and the policy targets are distributions over valid moves."""
    from visapi_b200.coach import BatchedSelfPlay, Coach
    from visapi_b200.game import BinPackGame
    from visapi_b200.nnet import TorchBinPackNet
    np.random.seed(0)
    game = BinPackGame(8, 8, 6)
    net = TorchBinPackNet(game, epochs=1)
    args = {"numMCTSSims": 12, "cpuct": 1.0, "tempThreshold": 3, "numIters": 1, "numEps": 2,
            "checkpoint": str(tmp_path)}
    coach = Coach(game, net, args)
    ex = coach.executeEpisode()
    assert 1 <= len(ex) <= 6
    for grid, tiles, pi, r in ex:
        assert grid.shape == (8, 8) and tiles.shape == (12, 2)
        assert abs(sum(pi) - 1.0) < 1e-9 and 0 < r <= 1
        valid = game.getValidMoves((grid, tiles), 1)
        assert all(p == 0 for p, v in zip(pi, valid) if not v)
    losses = coach.learn()
    assert len(losses) == 1 and np.isfinite(losses[0])
    assert (tmp_path / "best.pth.tar").exists()
    sp = BatchedSelfPlay(8, 8, 6, net, args, n_games=8, seed=1)
    examples, reward, searches = sp.play()
    assert len(reward) == 8 and ((reward > 0) & (reward <= 1)).all()
    assert len(examples) >= 8 and searches >= 8 * 12
    assert sp.leaf_batches <= searches            # leaves of all games share a batch
    for grid, tiles, pi, r in examples:
        assert abs(pi.sum() - 1.0) < 1e-9
    assert np.isfinite(net.train(examples))


def test_run_mcts_command(tmp_path, capsys):
    """manage.py run_mcts, DB-free: generated instances and --from_file, results CSV in the
    reference's csv_results schema, every reported solution verified by the command itself."""
    import csv
    from visapi_b200 import run_mcts
    out = tmp_path / "results.csv"
    run_mcts.main(["20", "11", "11", "--n_sim", "50", "--count", "3", "--both_strategies",
                   "--out", str(out), "--quiet"])
    run_mcts.main(["20", "0", "0", "--from_file", "--set", "g", "--limit", "16", "--n_sim", "100",
                   "--avg_depth", "--out", str(out), "--quiet"])
    run_mcts.main(["8", "6", "7", "--n_sim", "20", "--count", "2", "--quiet",
                   "--dump_tree", str(tmp_path / "trees")])
    import json
    trees = sorted((tmp_path / "trees").glob("*_tree.json"))
    assert len(trees) == 2
    t = json.load(open(trees[0]))
    assert t["tree"]["children"] and "board" in t["tree"] and "solution_tiles_order" in t
    text = capsys.readouterr().out
    assert "solved (every solution replayed and verified)" in text
    rows = list(csv.DictReader(open(out), delimiter=";"))
    assert len(rows) == 3 * 2 + 16
    assert set(rows[0]) == {"rows", "cols", "tiles", "n_tiles", "strategy", "n_simulations", "score",
                            "solution_found", "n_tiles_placed", "created_on"}
    assert all(r["n_tiles"] == "20" for r in rows)
    assert {r["strategy"] for r in rows} == {"max_depth", "avg_depth"}
    assert all(int(r["n_tiles_placed"]) > 0 for r in rows)


def test_device_leaf_path_matches_host_path():
    """BatchedMCTS.simulate_on_device (no host copies; masking and normalising in the backup
    kernel) visits the same nodes as the host path when the net output is the same."""
    import torch
    from visapi_b200.puct import BatchedMCTS
    h, w, n, B, sims = 10, 10, 8, 6, 40
    from visapi_b200.game import BinPackGame
    np.random.seed(3)
    games = [BinPackGame(h, w, n) for _ in range(B)]
    grids = np.zeros((B, h, w))
    tiles = np.array([[[int(t[0]), int(t[1])] for t in g._base_board.tiles] for g in games])

    def policy_of(filled):               # depends only on the number of filled cells
        a = torch.arange(2 * n, dtype=torch.float64)
        return ((a * 7 + filled * 3) % 11 + 1)

    def predict_batch(gr):
        pis, vs = [], []
        for g in gr:
            f = float(np.count_nonzero(g))
            p = policy_of(f).numpy()
            pis.append(p / p.sum())
            vs.append(0.25 + 0.5 * f / g.size)
        return np.array(pis), np.array(vs)

    def predict_dev(gr, tl):
        f = gr.sum(dim=(1, 2)).to(torch.float64)
        a = torch.arange(2 * n, dtype=torch.float64, device=gr.device)
        p = ((a[None, :] * 7 + f[:, None] * 3) % 11 + 1)
        p = p / p.sum(dim=1, keepdim=True)
        return p.to(torch.float32), (0.25 + 0.5 * f / (h * w)).to(torch.float32)

    host = BatchedMCTS(h, w, 2 * n, predict_batch, sims, 1.0, B)
    host.set_roots(grids, tiles)
    host.simulate()
    dev = BatchedMCTS(h, w, 2 * n, None, sims, 1.0, B)
    dev.set_roots(grids, tiles)
    dev.simulate_on_device(predict_dev)
    ch, cd = host._tree.root_counts(), dev._tree.root_counts()
    assert (ch[0].sum(axis=1) == cd[0].sum(axis=1)).all()
    # float32 policies differ from float64 ones in the last bits: visit counts may differ by
    # a few visits when two UCB values are nearly tied, never by much
    assert np.abs(ch[0] - cd[0]).max() <= 3
    assert (cd[1] > 1).all()


def _exact_nets(h, w, n):
    """Host and device nets whose outputs are exactly representable in float32 (small integers
    as unnormalised policy, eighths as value): masking and normalising them gives the same
    float64 priors on both paths, so the two searches must agree visit for visit."""
    import torch

    def predict_batch(gr):
        pis, vs = [], []
        for g in gr:
            f = int(np.count_nonzero(g))
            pis.append(np.array([float((a * 7 + f * 3) % 11 + 1) for a in range(2 * n)]))
            vs.append((f % 8) / 8.0)
        return np.array(pis), np.array(vs)

    def predict_dev(gr, tl):
        f = gr.sum(dim=(1, 2)).to(torch.int64)
        a = torch.arange(2 * n, dtype=torch.int64, device=gr.device)
        p = ((a[None, :] * 7 + f[:, None] * 3) % 11 + 1).to(torch.float32)
        return p, ((f % 8).to(torch.float32) / 8.0)
    return predict_batch, predict_dev


def _random_roots(h, w, n, B, seed):
    from visapi_b200.game import BinPackGame
    np.random.seed(seed)
    games = [BinPackGame(h, w, n) for _ in range(B)]
    tiles = np.array([[[int(t[0]), int(t[1])] for t in g._base_board.tiles] for g in games])
    return np.zeros((B, h, w)), tiles


def test_device_leaf_path_is_exact_with_float32_exact_net():
    """The device-leaf path (the fast one) against the host-leaf path, visit for visit."""
    from visapi_b200.puct import BatchedMCTS
    h, w, n, B, sims = 10, 10, 8, 6, 60
    grids, tiles = _random_roots(h, w, n, B, 3)
    predict_batch, predict_dev = _exact_nets(h, w, n)
    host = BatchedMCTS(h, w, 2 * n, predict_batch, sims, 1.0, B)
    host.set_roots(grids, tiles)
    host.simulate()
    dev = BatchedMCTS(h, w, 2 * n, None, sims, 1.0, B)
    dev.set_roots(grids, tiles)
    dev.simulate_on_device(predict_dev)
    ch, cd = host._tree.root_counts(), dev._tree.root_counts()
    assert np.array_equal(ch[0], cd[0]) and np.array_equal(ch[1], cd[1])
    assert np.array_equal(host.action_probs(1), dev.action_probs(1))


def test_simulate_on_device_under_a_side_stream():
    """the engine kernels and the torch ops between them share ONE stream even when
    the caller runs under torch.cuda.stream(...); same visits as on the default stream."""
    import torch
    from visapi_b200.puct import BatchedMCTS
    h, w, n, B, sims = 10, 10, 8, 16, 40
    grids, tiles = _random_roots(h, w, n, B, 5)
    _, predict_dev = _exact_nets(h, w, n)
    ref = BatchedMCTS(h, w, 2 * n, None, sims, 1.0, B)
    ref.set_roots(grids, tiles)
    ref.simulate_on_device(predict_dev)
    want = ref._tree.root_counts()[0]
    side = torch.cuda.Stream()
    for _ in range(3):
        t = BatchedMCTS(h, w, 2 * n, None, sims, 1.0, B)
        t.set_roots(grids, tiles)
        with torch.cuda.stream(side):
            t.simulate_on_device(predict_dev)
        side.synchronize()
        assert np.array_equal(t._tree.root_counts()[0], want)
    assert ref._tree.engine.stream != side.cuda_stream          # the binding was undone


def test_puct_root_parallel_sum_rule():
    """independent trees per 'rank' from the same roots, root Nsa and Wsa summed
    before the policy is formed.  Three ranks are emulated by three searches with different
    numbers of simulations (so their counts differ); with no process group the all-reduce is the
    identity, so the merge is checked against the sum of the per-rank statistics by hand and
    against probs_from_counts / getActionProb's rule."""
    import torch
    from visapi_b200.parallel import _DevicePointer, allreduce_root_stats, probs_from_counts
    from visapi_b200.puct import BatchedMCTS
    h, w, n, B = 10, 10, 8, 5
    grids, tiles = _random_roots(h, w, n, B, 11)
    _, predict_dev = _exact_nets(h, w, n)
    counts_sum = np.zeros((B, 2 * n), np.int64)
    wsa_sum = np.zeros((B, 2 * n))
    per_rank = []
    for sims in (20, 35, 50):
        t = BatchedMCTS(h, w, 2 * n, None, sims, 1.0, B)
        t.set_roots(grids, tiles)
        t.simulate_on_device(predict_dev)
        probs, q = t.action_probs_root_parallel(temp=1)          # one rank: identity merge
        assert np.array_equal(probs, t.action_probs(1))
        c_ptr, w_ptr = t._tree.root_stats_dev()
        c = torch.as_tensor(_DevicePointer(c_ptr, B * 2 * n * 4), device="cuda").view(torch.int32).view(B, -1)
        wv = torch.as_tensor(_DevicePointer(w_ptr, B * 2 * n * 8), device="cuda").view(torch.float64).view(B, -1)
        c, wv = c.cpu().numpy().copy(), wv.cpu().numpy().copy()
        assert np.array_equal(c, t._tree.root_counts()[0]) and c.sum(axis=1).max() <= sims
        assert np.allclose(np.where(c > 0, wv / np.maximum(c, 1), 0.0), q)
        per_rank.append(c)
        counts_sum += c
        wsa_sum += wv
    assert not np.array_equal(per_rank[0] * 3, counts_sum)       # the ranks really differ
    merged_c, merged_w, merged_q = allreduce_root_stats(torch.from_numpy(counts_sum),
                                                       torch.from_numpy(wsa_sum))
    probs = probs_from_counts(merged_c.numpy(), 1)
    assert np.allclose(probs.sum(axis=1), 1.0)
    assert np.array_equal(probs, counts_sum / counts_sum.sum(axis=1, keepdims=True))
    assert np.array_equal(probs_from_counts(counts_sum, 0).argmax(axis=1), counts_sum.argmax(axis=1))
    assert np.allclose(merged_q.numpy() * counts_sum, wsa_sum)


def _game_with_tiles(h, w, n, init_tiles, device=None, np_seed=None):
    """BinPackGame whose next getInitBoard() hands out ``init_tiles`` (the reference's instance)
    and then seeds np.random as the fixture generator did at that point."""
    from visapi_b200.game import BinPackGame, Board
    np.random.seed(0)
    game = BinPackGame(h, w, n, device=device)
    queue = [list(map(list, t)) for t in (init_tiles if isinstance(init_tiles[0][0], list) else [init_tiles])]

    def get_init_board(_g=game, _q=queue):
        tiles = _q.pop(0)
        _g._base_board = Board(h, w, [tuple(t) for t in tiles], n, device=_g.device)
        if np_seed is not None:
            np.random.seed(np_seed)
        return _g._base_board.state, _g._base_board.vis_state
    game.getInitBoard = get_init_board
    return game


def test_coach_episode_matches_reference(golden):
    """alpha/Coach.py:25-63 run by the reference under the stub net and a seeded np.random
    (tests/golden/coach_episode.json): the same example stream, bit for bit."""
    from visapi_b200.coach import Coach

    for case in golden("coach_episode.json"):
        h, w, n = case["height"], case["width"], case["n_tiles"]
        stub = _stub_predict(2 * n, [])

        class Net(object):
            calls = 0

            def predict(self, board):
                Net.calls += 1
                return stub.predict(board[0] if isinstance(board, (tuple, list)) else board)
        args = {"numMCTSSims": case["num_sims"], "cpuct": case["cpuct"], "tempThreshold": case["tempThreshold"]}
        game = _game_with_tiles(h, w, n, case["init_tiles"], np_seed=case["np_seed"])
        coach = Coach(game, Net(), args)
        examples = coach.executeEpisode()
        assert len(examples) == len(case["examples"])
        assert Net.calls == case["leaf_evals"]
        for got, want in zip(examples, case["examples"]):
            assert hex_rows_of(got[0]) == want["board"]
            assert [[int(t[0]), int(t[1])] for t in got[1]] == want["tiles"]
            assert [float(p) for p in got[2]] == want["pi"]
            assert float(got[3]) == want["reward"]


def test_arena_matches_reference(golden):
    """alpha/Arena.py:26-123 with two scripted players on the reference's instances."""
    from visapi_b200.arena import Arena
    for case in golden("arena.json"):
        h, w, n = case["height"], case["width"], case["n_tiles"]
        game = _game_with_tiles(h, w, n, case["init_tiles"])

        def first_valid(board, _g=game):
            return int(np.argmax(_g.getValidMoves(board, 1)))

        def last_valid(board, _g=game):
            v = np.asarray(_g.getValidMoves(board, 1))
            return int(len(v) - 1 - np.argmax(v[::-1]))
        one, two, draws = Arena(first_valid, last_valid, game).playGames(case["games"])
        assert (one, two, draws) == (case["one_won"], case["two_won"], case["draws"])


def _strip(node):
    counter = [0]

    def walk(n):
        me = counter[0]
        counter[0] += 1
        score = n["score"]
        return {"name": me, "tiles": [[int(t[0]), int(t[1])] for t in n["tiles"]],
                "board": hex_rows_of(np.asarray(n["board"])),
                "tile_placed": None if n["tile_placed"] is None else [int(v) for v in n["tile_placed"]],
                "score": float(score) if isinstance(score, (float, np.floating, np.integer)) else score,
                "children": [walk(c) for c in n["children"]]}
    return walk(node)


def test_tree_export_matches_reference(golden):
    """f3: State.render_to_json() / render_children() of the tree CustomMCTS.predict builds,
    against the reference's own rendering of its tree (engine/state.py:16-61,88-136,
    run_mcts.py:227-239) — node for node: tiles, occupancy, tile_placed, score, children."""
    import re
    from visapi_b200.mcts import CustomMCTS
    for case in golden("tree_render.json"):
        tiles = [tuple(t) for t in case["tiles"]]
        m = CustomMCTS(tiles, np.zeros((case["rows"], case["cols"])), strategy=case["strategy"],
                       seed=123, stream=case["stream"])
        root, depth, found = m.predict(N=case["n_sim"])
        assert (depth, found) == (case["depth"], case["solution_found"])
        nested, index = root.render_children(only_ids=False)

        def labels(d):
            out = []
            for k, v in d.items():
                k = re.sub(r"^\(\d+\) ", "", k)
                k = k.split(" Sim. depth:")[0] + " Sim. depth:" + k.split(" Sim. depth:")[1].split(" ")[0]
                out.append([k, labels(v)])
            return out
        assert labels(nested) == case["labels_before_centralize"]
        ids, _ = root.render_children(only_ids=True)

        def shape(d):
            return [shape(v) for v in d.values()]
        assert shape(ids) == case["ids_shape_before_centralize"]
        root.centralize_node_with_children()
        got = _strip(root.render_to_json())
        want = case["tree"]

        def drop(n):
            n = dict(n)
            n.pop("board_max", None)
            n["score"] = float(n["score"]) if isinstance(n["score"], (int, float)) else n["score"]
            n["children"] = [drop(c) for c in n["children"]]
            return n
        assert drop(got) == drop(want)


def test_ascii_tree_dump(tmp_path):
    """run_mcts.py:268-279: the search tree as the asciitree LeftAligned text next to the JSON."""
    from visapi_b200 import run_mcts
    out = tmp_path / "trees"
    run_mcts.main(["6", "8", "8", "--n_sim", "5", "--count", "1", "--quiet", "--dump_tree", str(out)])
    names = sorted(os.listdir(out))
    assert any(x.endswith("_tree.json") for x in names) and any(x.endswith(".txt") for x in names)
    text = open(os.path.join(out, [x for x in names if x.endswith(".txt")][0])).read()
    assert text.startswith("(") and "\n +-- (" in text and "Remaining tiles" in text
