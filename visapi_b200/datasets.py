"""Problem sets and instance I/O for the rollout engine (host side).

Bundled sets (visapi_b200/data/*.npz) are the reference's problems/problems_g.csv and
problems/problems_b.csv (1000 twenty-tile instances each) converted by
tests/golden/make_golden.py; ``read_problems_csv`` reads the reference's own CSV
dialects directly (SURVEY.md section 8d "Concrete inputs"):

  problems_g.csv / problems_b.csv   ';'-separated  rows;cols;tiles;n_tiles[;external_id]
                                    tiles = [[a, b], ...]
  problems_*_braam.csv              ','-separated  board_width,board_height,tiles,...
                                    tiles = [{"X": a, "Y": b}, ...]; rows := board_width,
                                    cols := board_height
  puzzles_n20_with_solutions_2.csv  ','-separated  id,...,board_width,board_height,tiles,
                                    tiles_placed  (engine/data_generator.py:117-135;
                                    run_mcts.py:56-61 uses rows := board_height,
                                    cols := board_width)
"""
import csv
import json
import os

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class Problem(object):
    """One instance: board shape and the n base tiles (one orientation each)."""

    __slots__ = ("rows", "cols", "tiles", "ident")

    def __init__(self, rows, cols, tiles, ident=""):
        self.rows, self.cols = int(rows), int(cols)
        self.tiles = [(int(t[0]), int(t[1])) for t in tiles]
        self.ident = str(ident)

    def entries(self):
        """Both orientations of every tile sorted by (x[1], x[0]) — the list
        CustomMCTS receives (engine/data_generator.py:86-98)."""
        return both_orientations_sorted(self.tiles)

    def board(self):
        return np.zeros((self.rows, self.cols))


def both_orientations_sorted(tiles):
    out = []
    for t in tiles:
        out.append((int(t[0]), int(t[1])))
        out.append((int(t[1]), int(t[0])))
    return sorted(out, key=lambda x: (x[1], x[0]))


def load_problem_set(name):
    """name: 'g' (Guillotinable) or 'b' (Braam); returns a list of Problem."""
    path = os.path.join(DATA_DIR, "problems_%s.npz" % name)
    z = np.load(path)
    out = []
    for i in range(len(z["rows"])):
        n = int(z["n_tiles"][i])
        out.append(Problem(z["rows"][i], z["cols"][i], z["tiles"][i, :n].tolist(), z["ids"][i]))
    return out


def read_problems_csv(path):
    """Reads any of the reference's three CSV dialects; returns a list of Problem."""
    with open(path, newline="") as f:
        head = f.readline()
    delim = ";" if head.count(";") > head.count(",") else ","
    out = []
    with open(path, newline="") as f:
        for row in csv.DictReader(f, delimiter=delim, quotechar='"'):
            raw = json.loads(row["tiles"])
            if raw and isinstance(raw[0], dict):
                tiles = [(t["X"], t["Y"]) for t in raw]
            else:
                tiles = [(t[0], t[1]) for t in raw]
            if "rows" in row:
                rows, cols = int(row["rows"]), int(row["cols"])
                ident = row.get("external_id", "")
            elif "tiles_placed" in row:                 # puzzles_n20_with_solutions_2.csv
                rows, cols = int(row["board_height"]), int(row["board_width"])
                ident = row.get("id", "")
            else:                                       # *_braam.csv
                rows, cols = int(row["board_width"]), int(row["board_height"])
                ident = row.get("their_id", row.get("problem_id", ""))
            out.append(Problem(rows, cols, tiles, ident))
    return out


def batch_arrays(problems, entry_stride=None):
    """Pack a list of Problem for Engine.search / Engine.flat_search."""
    ent = [p.entries() for p in problems]
    es = entry_stride or max(len(e) for e in ent)
    es += es % 2
    tiles = np.zeros((len(problems), es, 2), np.uint8)
    for i, e in enumerate(ent):
        tiles[i, :len(e)] = np.asarray(e, np.uint8)
    return {
        "rows": np.array([p.rows for p in problems], np.int32),
        "cols": np.array([p.cols for p in problems], np.int32),
        "n_entries": np.array([len(e) for e in ent], np.int32),
        "tiles": tiles,
    }


RESULT_FIELDS = ["rows", "cols", "tiles", "n_tiles", "strategy", "n_simulations", "score",
                 "solution_found", "n_tiles_placed", "created_on"]


def write_results_csv(path, rows, append=False):
    """Rows in the schema of the reference's csv_results/results_*.csv (';'-separated,
    all fields quoted) so outputs diff directly against the published files."""
    mode = "a" if append and os.path.exists(path) else "w"
    with open(path, mode, newline="") as f:
        w = csv.writer(f, delimiter=";", quotechar='"', quoting=csv.QUOTE_ALL)
        if mode == "w":
            w.writerow(RESULT_FIELDS)
        for r in rows:
            w.writerow([r.get(k, "") for k in RESULT_FIELDS])
