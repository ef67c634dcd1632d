"""Host-side logic of the product that needs no GPU."""
import json
import os
import random

import numpy as np
import pytest

from tests.conftest import board_from_hex


def test_problem_sets_and_batches():
    from visapi_b200 import datasets
    g, b = datasets.load_problem_set("g"), datasets.load_problem_set("b")
    assert len(g) == 1000 and len(b) == 1000
    assert all(len(p.tiles) == 20 for p in g + b)
    assert all(sum(h * w for h, w in p.tiles) == p.rows * p.cols for p in g + b)   # perfect packings
    assert max(p.cols for p in g) <= 40 and max(p.cols for p in b) <= 114
    batch = datasets.batch_arrays(g[:5])
    assert batch["tiles"].shape == (5, 40, 2) and batch["n_entries"].tolist() == [40] * 5
    ent = g[0].entries()
    assert ent == sorted(ent, key=lambda x: (x[1], x[0])) and len(ent) == 40
    assert batch["tiles"][0].tolist() == [list(t) for t in ent]


def test_csv_dialects_and_result_writer(tmp_path):
    from visapi_b200 import datasets
    g = datasets.load_problem_set("g")[:3]
    p1 = tmp_path / "g.csv"
    with open(p1, "w") as f:
        f.write('"rows";"cols";"tiles";"n_tiles"\n')
        for p in g:
            f.write('"%d";"%d";"%s";"20"\n' % (p.rows, p.cols, json.dumps([list(t) for t in p.tiles])))
    p2 = tmp_path / "braam.csv"
    with open(p2, "w") as f:
        f.write('"job_id","puzzle_id","num_tiles","board_width","board_height","tiles","start","end","their_id"\n')
        for p in g:
            tiles = json.dumps([{"X": t[0], "Y": t[1]} for t in p.tiles]).replace('"', '""')
            f.write('"1","1","20","%d","%d","%s","","","7"\n' % (p.rows, p.cols, tiles))
    for path in (p1, p2):
        got = datasets.read_problems_csv(str(path))
        assert [(q.rows, q.cols, q.tiles) for q in got] == [(q.rows, q.cols, q.tiles) for q in g]
    out = tmp_path / "res.csv"
    row = {"rows": 3, "cols": 4, "tiles": "[[1, 2]]", "n_tiles": 1, "strategy": "max_depth",
           "n_simulations": 10, "score": 0.0, "solution_found": True, "n_tiles_placed": 5,
           "created_on": "now"}
    datasets.write_results_csv(str(out), [row])
    datasets.write_results_csv(str(out), [row], append=True)
    lines = open(out).read().splitlines()
    assert lines[0].startswith('"rows";"cols";"tiles";"n_tiles";"strategy";"n_simulations";"score"')
    assert len(lines) == 3


def test_generator_reproduces_reference_instances(golden):
    """tile lists recorded from the reference's generator (tests/golden/flat_search.json)"""
    from visapi_b200.data_generator import DataGenerator
    seen = 0
    for c in golden("flat_search.json"):
        if not c["name"].startswith("gen"):
            continue
        n, shape, seed = c["name"][3:].split("_")
        r, cc = shape.split("x")
        random.seed(int(seed[1:]))
        tiles, board = DataGenerator().gen_tiles_and_board(int(n), int(r), int(cc), order_tiles=True)
        assert [[int(t[0]), int(t[1])] for t in tiles] == c["tiles"]
        assert board.shape == (c["rows"], c["cols"]) and not board.any()
        seen += 1
    assert seen >= 5
    assert DataGenerator()._split_bin([5, 4, (1, 1)], 0, 2) == [[2, 4, (1, 1)], [3, 4, (3, 1)]]
    assert DataGenerator()._split_bin([5, 4, (1, 1)], 1, 3) == [[5, 3, (1, 1)], [5, 1, (1, 4)]]
    assert DataGenerator.x_y_str_to_bins_format('[{"X":4,"Y":1},{"X":14,"Y":1}]') == \
        [[4, 1, (0, 0)], [14, 1, (0, 0)]]


def test_replay_solution(golden):
    from visapi_b200.verify import replay_solution
    solved = [c for c in golden("flat_search.json") if c["solution_found"]]
    assert solved
    for c in solved:
        pool, base = [tuple(t) for t in c["tiles"]], []
        while pool:
            t = pool.pop(0)
            pool.remove((t[1], t[0]))
            base.append(t)
        order = c["solution_tiles_order"]
        assert replay_solution(c["rows"], c["cols"], base, order) == (True, "ok")
        assert not replay_solution(c["rows"], c["cols"], base, order[:-1])[0]
        bad = [list(t) for t in order]
        bad[0] = [bad[0][0] + 1, bad[0][1]]
        assert not replay_solution(c["rows"], c["cols"], base, bad)[0]
        if len(order) > 2 and order[0] != order[1]:
            swapped = [order[1], order[0]] + order[2:]
            # a different order may or may not pack; it must never be accepted with overlaps
            ok, why = replay_solution(c["rows"], c["cols"], base, swapped)
            assert ok or why != "ok"


def test_pack_unpack_roundtrip():
    from visapi_b200 import pack_occupancy, pack_tiles, unpack_occupancy
    rng = np.random.default_rng(0)
    for cols in (1, 31, 32, 33, 64, 100, 128):
        g = (rng.random((3, 7, cols)) < 0.4).astype(np.float64) * rng.integers(1, 4, (3, 7, cols))
        occ = pack_occupancy(g, colorful=True)
        assert occ.shape == (3, 7, (cols + 31) // 32) and occ.dtype == np.uint32
        np.testing.assert_array_equal(unpack_occupancy(occ, cols), (g != 0).astype(np.uint8))
        np.testing.assert_array_equal(unpack_occupancy(pack_occupancy(g, colorful=False), cols),
                                      (g == 1).astype(np.uint8))
    assert (pack_occupancy(np.array([[1, 0, 1]]))[0, 0, 0]) == 0b101
    t = pack_tiles([[(1, 2), (3, 4)], [(5, 6)]])
    assert t.tolist() == [[[1, 2], [3, 4]], [[5, 6], [0, 0]]]
    with pytest.raises(ValueError):
        pack_tiles([[(300, 1)]])


def test_partition_helpers():
    from visapi_b200 import parallel
    for n_sim in (1, 7, 100, 1000, 5000):
        for world in (1, 2, 3, 4, 8):
            parts = [parallel.split_sims(n_sim, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(p[1] for p in parts) == n_sim
            for a, b in zip(parts, parts[1:]):
                assert a[0] + a[1] == b[0]
    idx = [parallel.shard_indices(10, r, 4) for r in range(4)]
    assert sorted(i for part in idx for i in part) == list(range(10))
    assert parallel.CHILD_STATS_DTYPE.itemsize == 24


def test_merge_child_stats_rule():
    from visapi_b200 import parallel
    rng = np.random.default_rng(1)
    n_sim, world, n_children = 103, 4, 9
    steps = rng.integers(0, 20, (n_children, n_sim))
    solved = rng.random((n_children, n_sim)) < 0.01
    solved[0] = False
    per_rank = np.zeros((world, n_children), parallel.CHILD_STATS_DTYPE)
    for r in range(world):
        b, n = parallel.split_sims(n_sim, r, world)
        for c in range(n_children):
            hit = np.flatnonzero(solved[c, b:b + n])
            per_rank[r, c]["max_steps"] = steps[c, b:b + n].max()
            per_rank[r, c]["sum_steps"] = steps[c, b:b + n].sum()
            if hit.size:
                per_rank[r, c]["first_solved"] = b + hit[0]
                per_rank[r, c]["prefix_steps"] = steps[c, b:b + hit[0] + 1].sum()
            else:
                per_rank[r, c]["first_solved"] = parallel.NONE_U32
                per_rank[r, c]["prefix_steps"] = steps[c, b:b + n].sum()
    merged = parallel.merge_child_stats(per_rank)
    for c in range(n_children):
        hit = np.flatnonzero(solved[c])
        assert merged[c]["max_steps"] == steps[c].max()
        assert merged[c]["sum_steps"] == steps[c].sum()
        if hit.size:
            assert merged[c]["first_solved"] == hit[0]
            assert merged[c]["prefix_steps"] == steps[c, :hit[0] + 1].sum()
        else:
            assert merged[c]["first_solved"] == parallel.NONE_U32
            assert merged[c]["prefix_steps"] == steps[c].sum()


def test_state_and_helpers():
    from visapi_b200.mcts import get_max_index
    from visapi_b200.solution_checker import SolutionChecker
    from visapi_b200.state import State
    board = np.zeros((2, 3))
    s = State(board, [(1, 1), (1, 1)])
    board[0, 0] = 5
    assert not s.board.any()                      # State copies its board
    c = State(s.board, [], parent=s)
    c.score, c.tile_placed = 3, (1, 1)
    s.children.append(c)
    assert s.copy().parent is None and s.child_with_biggest_score() is c
    js = s.render_to_json()
    assert js["children"][0]["score"] == 3 and js["children"][0]["tile_placed"] == (1, 1)
    tree, nodes = s.render_children(only_ids=True)
    assert list(tree.keys()) == [s.uuid_str()] and set(nodes) == {s.uuid_str(), c.uuid_str()}
    assert "Remaining tiles: 1.0" in str(s)

    class E(object):
        def __init__(self, score):
            self.score = score
    assert get_max_index([None, E(2), E(5), E(5), None]) == 2
    assert get_max_index([None, None]) == 0
    assert SolutionChecker.get_tiles_with_orientation([[1, 2], [3, 4]]) == [[1, 2], [3, 4], (2, 1), (4, 3)]
    assert SolutionChecker.get_n_nonplaced_tiles(np.array([[1, 2], [0, 0]])) == 1


def test_cli_instance_loading():
    from visapi_b200 import run_mcts
    import argparse
    parser = argparse.ArgumentParser()
    run_mcts.add_arguments(parser)
    opts = vars(parser.parse_args(["20", "11", "11", "--from_file", "--set", "b", "--limit", "7",
                                   "--n_sim", "50", "--avg_depth", "--device", "cuda:1"]))
    assert opts["n_sim"] == 50 and opts["avg_depth"] and opts["device"] == "cuda:1"
    probs = run_mcts.load_instances(opts)
    assert len(probs) == 7 and all(len(p.tiles) == 20 for p in probs)
    opts = vars(parser.parse_args(["8", "6", "7", "--count", "3"]))
    probs = run_mcts.load_instances(opts)
    assert len(probs) == 3 and all((p.rows, p.cols, len(p.tiles)) == (6, 7, 8) for p in probs)
    assert all(sum(h * w for h, w in p.tiles) == 42 for p in probs)


def test_cli_reports_instances_outside_the_engine_limits(tmp_path, capsys, monkeypatch):
    """A 129-column instance (puzzles_n20_with_solutions_2.csv has one, 129 x 67) is reported and
    skipped by the command before any device call; with nothing left the command returns."""
    from visapi_b200 import datasets, run_mcts
    big = datasets.Problem(67, 129, [(67, 129)], "too-wide")
    monkeypatch.setattr(run_mcts, "load_instances", lambda options: [big])
    monkeypatch.setattr(run_mcts.runtime, "set_default_device", lambda device: None)
    out = run_mcts.run_mcts({"device": "cuda:0", "n_tiles": 1, "both_strategies": False, "avg_depth": False,
                             "n_sim": 5, "seed": 123, "quiet": True, "out": None, "dump_tree": None})
    assert out == []
    text = capsys.readouterr().out
    assert "skipped 1 instance" in text and "67x129" in text


def test_ascii_tree_is_asciitree_left_aligned():
    """run_mcts.py:268-279 writes asciitree.LeftAligned()(tree); asciitree (0.3.3, not installed
    here) documents its output with this example, which the restatement must reproduce."""
    from collections import OrderedDict as OD
    from visapi_b200.run_mcts import ascii_tree
    tree = {"asciitree": OD([("sometimes", OD([("you", {})])),
                             ("just", OD([("want", OD([("to", {}), ("draw", {})]))])),
                             ("trees", {}),
                             ("in", OD([("your", OD([("terminal", {})]))]))])}
    assert ascii_tree(tree) == "\n".join([
        "asciitree",
        " +-- sometimes",
        " |   +-- you",
        " +-- just",
        " |   +-- want",
        " |       +-- to",
        " |       +-- draw",
        " +-- trees",
        " +-- in",
        "     +-- your",
        "         +-- terminal"])
