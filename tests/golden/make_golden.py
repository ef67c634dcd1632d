#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference through oracle/ref_harness.py (matplotlib + search stubs,
occupancy fix, Philox stream injected at binpack/MCTS.py:174) and records:

  transitions.json   random legal-move walks: per step LFB, free run, legal flags,
                     verdict and the next board, from SolutionChecker itself
  rollouts.json      CustomMCTS.perform_simulation outcomes under the Philox stream
  flat_search.json   CustomMCTS.predict outputs (depth, found, n_tiles_placed,
                     solution_tiles_order, every child's score)
  alpha_game.json    Board/BinPackGame walks: valid masks, add_tile, win state
  puct.json          MCTS.getActionProb with a deterministic stub net
  visapi_b200/data/problems_*.npz   the reference's problem sets, compact binary form
  published.json     aggregates of csv_results/results_{g,b}.csv

Nothing here is read at test time from /root/reference; the fixtures are committed.
"""
import csv
import hashlib
import json
import os
import random
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_harness  # noqa: E402

REF = ref_harness.REFERENCE_ROOT


def dump(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"))
    print("wrote %s (%d bytes)" % (name, os.path.getsize(path)))


def occ_rows(grid):
    """rows as lists of occupied column indices packed into Python ints (bit c = col c)."""
    out = []
    for row in np.asarray(grid):
        v = 0
        for c, x in enumerate(row):
            if x != 0:
                v |= 1 << c
        out.append(v)
    return out


def both_sorted(tiles):
    out = []
    for t in tiles:
        out.append((int(t[0]), int(t[1])))
        out.append((int(t[1]), int(t[0])))
    return sorted(out, key=lambda x: (x[1], x[0]))


# ------------------------------------------------------------------ datasets
def load_g(path):
    probs = []
    with open(path) as f:
        for row in csv.DictReader(f, delimiter=";"):
            probs.append((int(row["rows"]), int(row["cols"]), json.loads(row["tiles"]),
                          row.get("external_id", "")))
    return probs


def load_braam(path):
    probs = []
    with open(path) as f:
        for row in csv.DictReader(f, delimiter=","):
            tiles = [[t["X"], t["Y"]] for t in json.loads(row["tiles"])]
            probs.append((int(row["board_width"]), int(row["board_height"]), tiles,
                          row.get("their_id", row.get("problem_id", ""))))
    return probs


def write_npz(name, probs):
    n = len(probs)
    nt = max(len(p[2]) for p in probs)
    rows = np.array([p[0] for p in probs], np.int16)
    cols = np.array([p[1] for p in probs], np.int16)
    tiles = np.zeros((n, nt, 2), np.uint8)
    n_tiles = np.zeros(n, np.int16)
    for i, p in enumerate(probs):
        n_tiles[i] = len(p[2])
        tiles[i, :len(p[2])] = np.array(p[2], np.uint8)
    ids = np.array([str(p[3]) for p in probs])
    path = os.path.join(ROOT, "visapi_b200", "data", name)
    np.savez_compressed(path, rows=rows, cols=cols, tiles=tiles, n_tiles=n_tiles, ids=ids)
    print("wrote %s (%d bytes)" % (name, os.path.getsize(path)))


def datasets():
    g = load_g(os.path.join(REF, "problems", "problems_g.csv"))
    b = load_g(os.path.join(REF, "problems", "problems_b.csv"))
    bb = load_braam(os.path.join(REF, "problems", "problems_b_braam.csv"))
    assert len(g) == 1000 and len(b) == 1000 and len(bb) == 1000
    # rows := board_width, cols := board_height reproduces problems_b.csv
    for x, y in zip(b, bb):
        assert x[0] == y[0] and x[1] == y[1] and x[2] == y[2], (x, y)
    write_npz("problems_g.npz", g)
    write_npz("problems_b.npz", b)
    return g, b


def published():
    out = {}
    for name in ("g", "b"):
        agg = {}
        with open(os.path.join(REF, "csv_results", "results_%s.csv" % name)) as f:
            for row in csv.DictReader(f, delimiter=";"):
                k = "%s/%s" % (row["strategy"], row["n_simulations"])
                a = agg.setdefault(k, {"n": 0, "solved": 0, "score": 0.0, "placed": 0})
                a["n"] += 1
                a["solved"] += 1 if row["solution_found"] == "True" else 0
                a["score"] += float(row["score"])
                a["placed"] += int(row["n_tiles_placed"])
        out[name] = {k: {"n": a["n"], "solved": a["solved"],
                         "mean_score": a["score"] / a["n"],
                         "mean_n_tiles_placed": a["placed"] / a["n"]}
                     for k, a in sorted(agg.items())}
    dump("published.json", out)


# ------------------------------------------------------------------ instances
def sample_instances(ref, g, b):
    """(name, rows, cols, sorted both-orientation tile list)"""
    inst = []
    for i in (0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987):
        inst.append(("g%d" % i, g[i][0], g[i][1], both_sorted(g[i][2])))
    for i in (0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987):
        inst.append(("b%d" % i, b[i][0], b[i][1], both_sorted(b[i][2])))
    # widest / tallest members of each set (multi-word rows, > 32 rows)
    for name, ps in (("g", g), ("b", b)):
        wide = max(range(len(ps)), key=lambda k: ps[k][1])
        tall = max(range(len(ps)), key=lambda k: ps[k][0])
        for k in (wide, tall):
            inst.append(("%s%d" % (name, k), ps[k][0], ps[k][1], both_sorted(ps[k][2])))
    dg = ref.DataGenerator()
    for (n, r, c, seed) in ((20, 11, 11, 0), (20, 11, 11, 1), (8, 6, 7, 2), (50, 25, 25, 3),
                            (50, 40, 40, 4), (30, 33, 64, 5), (12, 70, 20, 6)):
        random.seed(seed)
        with ref_harness.quiet():
            tiles, board = dg.gen_tiles_and_board(n, r, c, order_tiles=True)
        inst.append(("gen%d_%dx%d_s%d" % (n, r, c, seed), board.shape[0], board.shape[1],
                     [(int(t[0]), int(t[1])) for t in tiles]))
    return inst


# ------------------------------------------------------------------ transitions
def transitions(ref, inst):
    SC = ref.SolutionChecker
    out = []
    for name, rows, cols, tiles in inst:
        for walk in range(3):
            rng = random.Random(zlib.crc32(('%s/%d' % (name, walk)).encode()))
            board = np.zeros((rows, cols))
            tl = list(tiles)
            steps = []
            val = 1
            while True:
                state = ref.State(board, tl)
                lfb = SC.get_next_lfb_on_grid(state.board)
                if lfb is None:
                    verdict, _ = SC.get_next_turn(state, tl[0] if tl else (1, 1), val)
                    steps.append({"lfb": None, "verdict": verdict})
                    break
                nxt = SC.get_next_occupied_col(state.board, lfb, colorful_states=True)
                legal = SC.get_valid_next_moves(state, tl, colorful_states=True)
                flags = [1 if (t[1] <= nxt - lfb[1] and t[0] + lfb[0] <= rows) else 0 for t in tl]
                assert [t for t, f in zip(tl, flags) if f] == legal
                # a8 mask on the same (0/1) occupancy
                mask = SC.get_valid_tile_actions_indexes_given_grid(
                    (state.board != 0).astype(float), [list(t) for t in tl])
                assert mask == flags
                step = {"lfb": [int(lfb[0]), int(lfb[1])], "next_occupied_col": int(nxt),
                        "legal": flags, "n_tiles": len(tl)}
                # a tile that cannot be placed, when there is one
                bad = [t for t, f in zip(tl, flags) if not f]
                if bad:
                    v, nb = SC.get_next_turn(state, bad[0], val, colorful_states=True)
                    assert nb is None
                    step["bad_tile"] = [int(bad[0][0]), int(bad[0][1])]
                    step["bad_verdict"] = v
                if not legal:
                    steps.append(step)
                    break
                pick = rng.randrange(len(legal))
                tile = legal[pick]
                val += 1
                verdict, nb = SC.get_next_turn(state, tile, val, destroy_state=False,
                                               colorful_states=True)
                assert verdict is True
                step["pick"] = pick
                step["tile"] = [int(tile[0]), int(tile[1])]
                step["board_after"] = [hex(v) for v in occ_rows(nb)]
                steps.append(step)
                board = nb
                tl = SC.eliminate_pair_tiles(tl, tile)
            out.append({"name": name, "walk": walk, "rows": rows, "cols": cols,
                        "tiles": [list(t) for t in tiles], "steps": steps})
    dump("transitions.json", out)


# ------------------------------------------------------------------ rollouts
def rollouts(ref, inst):
    out = []
    for k, (name, rows, cols, tiles) in enumerate(inst):
        for sim in range(6):
            board = np.zeros((rows, cols))
            tag = (k * 7 + sim) & 0xFFFFFFFF
            r = ref.run_rollout(tiles, board, seed=123, stream=k, tag=tag, sim=sim)
            out.append({"name": name, "rows": rows, "cols": cols,
                        "tiles": [list(t) for t in tiles], "seed": 123, "stream": k,
                        "tag": tag, "sim": sim, "solved": r["solved"], "depth": r["depth"],
                        "placements": r["placements"], "moves": r["moves"]})
    # rollouts from mid-game roots: play a few reference moves first
    SC = ref.SolutionChecker
    for k, (name, rows, cols, tiles) in enumerate(inst[:12]):
        rng = random.Random(k)
        board = np.zeros((rows, cols))
        tl = list(tiles)
        for _ in range(4):
            st = ref.State(board, tl)
            legal = SC.get_valid_next_moves(st, tl, colorful_states=True)
            if not legal:
                break
            t = legal[rng.randrange(len(legal))]
            _, board = SC.get_next_turn(st, t, 1, colorful_states=True)
            tl = SC.eliminate_pair_tiles(tl, t)
        for sim in range(4):
            r = ref.run_rollout(tl, np.copy(board), seed=7, stream=1000 + k, tag=99, sim=sim)
            out.append({"name": name + "_mid", "rows": rows, "cols": cols,
                        "tiles": [list(t) for t in tl], "board": [hex(v) for v in occ_rows(board)],
                        "seed": 7, "stream": 1000 + k, "tag": 99, "sim": sim,
                        "solved": r["solved"], "depth": r["depth"],
                        "placements": r["placements"], "moves": r["moves"]})
    dump("rollouts.json", out)


# ------------------------------------------------------------------ flat search
def flat_search(ref, inst):
    out = []
    cases = []
    for k, (name, rows, cols, tiles) in enumerate(inst):
        if len(tiles) > 60:
            cases.append((k, name, rows, cols, tiles, 3, "max_depth"))
            continue
        n_sim = 5 + (k % 4) * 3
        cases.append((k, name, rows, cols, tiles, n_sim, "max_depth"))
        if k % 2 == 0:
            cases.append((k, name, rows, cols, tiles, n_sim, "avg_depth"))
    # two larger-N cases (chunk boundaries: N > 32 and not a multiple of 32)
    cases.append((100, inst[0][0], inst[0][1], inst[0][2], inst[0][3], 70, "max_depth"))
    cases.append((101, inst[16][0], inst[16][1], inst[16][2], inst[16][3], 45, "avg_depth"))
    for (k, name, rows, cols, tiles, n_sim, strategy) in cases:
        board = np.zeros((rows, cols))
        r = ref.run_flat_search(tiles, board, n_sim, strategy, seed=123, stream=k)
        out.append({"name": name, "rows": rows, "cols": cols, "tiles": [list(t) for t in tiles],
                    "n_sim": n_sim, "strategy": strategy, "seed": 123, "stream": k,
                    "depth": r["depth"], "solution_found": r["solution_found"],
                    "score": r["score"], "n_tiles_placed": r["n_tiles_placed"],
                    "solution_tiles_order": r["solution_tiles_order"],
                    "children": r["children"], "rollouts": r["rollouts"],
                    "rollout_placements": r["rollout_placements"]})
        print("  flat %-22s N=%-3d %-9s depth=%2d found=%d placed=%d" % (
            name, n_sim, strategy, r["depth"], r["solution_found"], r["n_tiles_placed"]))
    dump("flat_search.json", out)


# ------------------------------------------------------------------ alpha path
def stub_predict(grid, tiles, action_size):
    """Deterministic stand-in for nnet.predict: depends on the board only through
    integer features so every implementation computes bit-identical float64 values."""
    filled = int(np.count_nonzero(np.asarray(grid)))
    cells = int(np.asarray(grid).size)
    pi = np.array([((a * 2654435761 + filled * 40503 + 12345) % 1000 + 1) for a in
                   range(action_size)], dtype=np.float64)
    pi = pi / pi.sum()
    v = 0.5 * filled / cells
    return pi, v


def alpha_game(ref):
    out = []
    for (h, w, n, seed) in ((8, 8, 6, 0), (20, 20, 15, 1), (11, 11, 20, 2), (25, 25, 50, 3),
                            (12, 40, 10, 4)):
        np.random.seed(seed)
        random.seed(seed)
        with ref_harness.quiet():
            game = ref.BinPackGame(h, w, n)
        base = game._base_board
        init_tiles = [list(int(v) for v in t) for t in base.tiles]
        for walk in range(3):
            rng = random.Random(seed * 10 + walk)
            b = base.with_state(state=(np.zeros([h, w]), np.array(init_tiles)))
            steps = []
            while True:
                state = (b.state[0], b.state[1])
                bb = base.with_state(state=state)
                valid = bb.get_valid_moves().tolist()
                ended = game.getGameEnded(state, 1)
                step = {"valid": valid, "ended": ended,
                        "tiles": [[int(t[0]), int(t[1])] for t in b.state[1]],
                        "board": [hex(v) for v in occ_rows(b.state[0])]}
                if ended != 0:
                    steps.append(step)
                    break
                choices = [i for i, v in enumerate(valid) if v]
                a = choices[rng.randrange(len(choices))]
                step["action"] = a
                steps.append(step)
                nb = base.with_state(state=state)
                nb.add_tile(a, 1)
                b = nb
            out.append({"height": h, "width": w, "n_tiles": n, "init_tiles": init_tiles,
                        "walk": walk, "steps": steps,
                        "action_size": game.getActionSize(),
                        "board_size": list(game.getBoardSize())})
    dump("alpha_game.json", out)


def puct(ref):
    out = []

    class Args(dict):
        __getattr__ = dict.__getitem__

    for (h, w, n, seed, sims, cpuct) in ((8, 8, 6, 0, 40, 1.0), (20, 20, 15, 1, 50, 1.0),
                                         (11, 11, 20, 2, 60, 1.5), (10, 10, 8, 3, 120, 1.0)):
        np.random.seed(seed)
        random.seed(seed)
        with ref_harness.quiet():
            game = ref.BinPackGame(h, w, n)
        base = game._base_board
        init_tiles = [list(int(v) for v in t) for t in base.tiles]

        # numpy >= 1.24: BinPackGame.getNextState (BinPackGame.py:51-55) fails on
        # np.copy of the ragged (grid, tiles) tuple; call the same Board methods
        # it calls, without that copy.
        def get_next_state(board, player, action, vis_state=None, _g=game):
            b = _g._base_board.with_state(state=board, vis_state=vis_state)
            b.add_tile(action, player)
            return b.state, player, b.vis_state
        game.getNextState = get_next_state

        class Net(object):
            def __init__(self):
                self.calls = 0

            def predict(self, grid):
                self.calls += 1
                return stub_predict(grid, None, 2 * n)

        net = Net()
        args = Args(numMCTSSims=sims, cpuct=cpuct)
        state = (np.zeros([h, w]), np.array(init_tiles))
        moves = []
        with ref_harness.quiet():
            for move in range(3):
                m = ref.MCTS(game, net, args)       # fresh tree per move (deterministic)
                calls0 = net.calls
                probs = m.getActionProb(state, temp=1)
                evals = net.calls - calls0
                probs0 = ref.MCTS(game, net, args).getActionProb(state, temp=0)
                ended = game.getGameEnded(state, 1)
                moves.append({"tiles": [[int(t[0]), int(t[1])] for t in state[1]],
                              "board": [hex(v) for v in occ_rows(state[0])],
                              "probs": [float(p) for p in probs],
                              "probs_temp0": [int(p) for p in probs0],
                              "leaf_evals": evals,
                              "n_nodes": len(m.Ps)})
                a = int(np.argmax(probs))
                nxt, _, _ = game.getNextState(state, 1, a)
                if nxt[0] is None or game.getGameEnded(nxt, 1) != 0:
                    break
                state = (nxt[0], np.array(nxt[1]))
        out.append({"height": h, "width": w, "n_tiles": n, "init_tiles": init_tiles,
                    "num_sims": sims, "cpuct": cpuct, "moves": moves})
    dump("puct.json", out)


def main():
    ref = ref_harness.load(occupancy_fix=True)
    g, b = datasets()
    published()
    inst = sample_instances(ref, g, b)
    transitions(ref, inst)
    rollouts(ref, inst)
    flat_search(ref, inst)
    # the alpha path runs as shipped (all tiles painted 1: both occupancy rules agree)
    alpha_game(ref)
    puct(ref)
    sha = hashlib.sha256()
    for name in sorted(os.listdir(HERE)):
        if name.endswith(".json") or name.endswith(".npz"):
            sha.update(open(os.path.join(HERE, name), "rb").read())
    print("fixtures sha256", sha.hexdigest())


if __name__ == "__main__":
    main()
