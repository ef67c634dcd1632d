"""DB-free ``run_mcts`` command (reference: binpack/management/commands/run_mcts.py:39-323).

    python manage.py run_mcts n_tiles rows cols [--n_sim N] [--from_file] [--avg_depth]
                              [--profile] [--device cuda:0] ...

Same positional arguments and flags as the reference command, with three differences
the survey calls for: the arguments are honoured (the reference
overwrites rows/cols/n_sim), ``--from_file`` reads the problem sets (bundled, or any of
the reference's CSV dialects via ``--problems``) instead of a file in the cwd, and no
Django/PostgreSQL is needed — one CSV row per problem in the schema of the reference's
csv_results/results_*.csv is appended to ``--out``.  ``--device`` selects the GPU; there
is no CPU execution path.  Under torchrun the problems are sharded over the ranks.
"""
import argparse
import datetime
import json
import os
import random
import sys
import time

import numpy as np

from . import datasets, runtime
from .data_generator import DataGenerator
from .mcts import CustomMCTS
from .verify import replay_solution

ORIENTATIONS = 2


MAX_SIDE = 128          # BP_MAX_ROWS / BP_MAX_COLS of include/rectpack_b200.h
MAX_ENTRIES = 128       # BP_MAX_ENTRIES


def add_arguments(parser):
    parser.add_argument("n_tiles", type=int, help="number of tiles")
    parser.add_argument("rows", type=int, help="number of rows")
    parser.add_argument("cols", type=int, help="number of cols")
    parser.add_argument("--n_sim", type=int, default=10, help="number of simulations from each action")
    parser.add_argument("--from_file", action="store_true", help="use instances from file")
    parser.add_argument("--avg_depth", action="store_true", help="avg_depth")
    parser.add_argument("--profile", action="store_true", help="Profile for performance bottlenecks")
    parser.add_argument("--device", default="cuda:0", help="CUDA device (cuda, cuda:N); no CPU path")
    parser.add_argument("--set", default="g", choices=["g", "b"], help="bundled problem set")
    parser.add_argument("--problems", default=None, help="CSV in one of the reference's dialects")
    parser.add_argument("--limit", type=int, default=0, help="use only the first K instances")
    parser.add_argument("--count", type=int, default=1, help="generated instances when not --from_file")
    parser.add_argument("--seed", type=int, default=123, help="Philox seed (reference: random.seed(123))")
    parser.add_argument("--both_strategies", action="store_true", help="run max_depth and avg_depth")
    parser.add_argument("--out", default=None, help="results CSV (reference results_*.csv schema)")
    parser.add_argument("--dump_tree", default=None,
                        help="directory for the per-problem search tree as JSON (State.render_to_json), "
                             "as the reference stores in Result.result_tree")
    parser.add_argument("--quiet", action="store_true")


def load_instances(options):
    """-> list of datasets.Problem"""
    n = options["n_tiles"]
    if options["from_file"]:
        probs = (datasets.read_problems_csv(options["problems"]) if options["problems"]
                 else datasets.load_problem_set(options["set"]))
        probs = [p for p in probs if len(p.tiles) == n]
    else:
        probs = []
        for k in range(options["count"]):
            random.seed(k)
            tiles, board = DataGenerator().gen_tiles_and_board(
                n, options["rows"], options["cols"], order_tiles=True)
            # entries hold both orientations; keep one of each pair as the base tiles
            pool, base = [tuple(int(v) for v in t) for t in tiles], []
            while pool:
                t = pool.pop(0)
                pool.remove((t[1], t[0]))
                base.append(t)
            probs.append(datasets.Problem(board.shape[0], board.shape[1], base, "generated-%d" % k))
    if options["limit"]:
        probs = probs[:options["limit"]]
    return probs


def run_mcts(options):
    import torch  # noqa: F401  (device plumbing only)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = options["device"]
    if world > 1:
        import torch.distributed as dist
        device = "cuda:%d" % local_rank
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
    runtime.set_default_device(device)

    probs = load_instances(options)
    # The engine holds a board as <= 128 rows of <= 128 columns (four 32-bit words per row or bit
    # plane) and <= 128 tile entries.  Larger instances are reported and skipped, never silently
    # altered: 1 of the 1615 instances of puzzles_n20_with_solutions_2.csv (129 x 67) is outside.
    too_big = [p for p in probs if p.rows > MAX_SIDE or p.cols > MAX_SIDE or 2 * len(p.tiles) > MAX_ENTRIES]
    if too_big:
        probs = [p for p in probs if p not in too_big]
        if rank == 0:
            print("skipped %d instance(s) outside the engine's limits (rows, cols <= %d; tiles <= %d): %s"
                  % (len(too_big), MAX_SIDE, MAX_ENTRIES // 2,
                     ", ".join("%dx%d/%d tiles" % (p.rows, p.cols, len(p.tiles)) for p in too_big[:8])))
    if not probs:
        print("no instance with %d tiles" % options["n_tiles"])
        return []
    strategies = (["max_depth", "avg_depth"] if options["both_strategies"]
                  else ["avg_depth" if options["avg_depth"] else "max_depth"])
    mine = list(range(rank, len(probs), world))
    rows_out = []
    t0 = time.time()
    for strategy in strategies:
        res = CustomMCTS.predict_many([probs[i] for i in mine], N=options["n_sim"], strategy=strategy,
                                      seed=options["seed"], streams=np.asarray(mine, dtype=np.uint32),
                                      device=device)
        for i, r in zip(mine, res):
            p = probs[i]
            if r["solution_found"]:
                ok, why = replay_solution(p.rows, p.cols, p.tiles, r["solution_tiles_order"])
                if not ok:
                    raise RuntimeError("problem %d: reported solution is invalid: %s" % (i, why))
            rows_out.append({
                "index": i, "rows": p.rows, "cols": p.cols,
                "tiles": json.dumps([list(t) for t in p.tiles]), "n_tiles": len(p.tiles),
                "strategy": strategy, "n_simulations": options["n_sim"], "score": float(r["score"]),
                "solution_found": r["solution_found"], "n_tiles_placed": r["n_tiles_placed"],
                "created_on": datetime.datetime.now().isoformat(sep=" "),
                "solution_tiles_order": r["solution_tiles_order"], "rollouts": r["rollouts"],
                "depth": r["depth"],
            })
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(rows_out, gathered, dst=0)
        if rank == 0:
            rows_out = sorted((r for part in gathered for r in part),
                              key=lambda r: (r["strategy"], r["index"]))
        dist.destroy_process_group()
        if rank != 0:
            return []
    dt = time.time() - t0
    if not options["quiet"]:
        for r in rows_out:
            print("problem %4d  %3dx%-3d %-9s N=%d  depth=%2d score=%.0f found=%s n_tiles_placed=%d" % (
                r["index"], r["rows"], r["cols"], r["strategy"], r["n_simulations"], r["depth"],
                r["score"], r["solution_found"], r["n_tiles_placed"]))
            if r["solution_found"]:
                print("  Final order of tiles: %s" % (r["solution_tiles_order"],))
    for strategy in strategies:
        sel = [r for r in rows_out if r["strategy"] == strategy]
        solved = sum(1 for r in sel if r["solution_found"])
        rollouts = sum(r["rollouts"] for r in sel)
        print("%s: %d/%d solved (every solution replayed and verified), mean score %.3f, "
              "mean n_tiles_placed %.0f, %d rollouts" % (
                  strategy, solved, len(sel), sum(r["score"] for r in sel) / len(sel),
                  sum(r["n_tiles_placed"] for r in sel) / len(sel), rollouts))
    print("wall %.3f s on %d GPU(s), device %s" % (dt, world, device))
    if options["out"]:
        datasets.write_results_csv(options["out"], rows_out, append=True)
    if options.get("dump_tree"):
        dump_trees(options, probs, strategies, device)
    return rows_out


class _NpEncoder(json.JSONEncoder):
    def default(self, obj):
        if isinstance(obj, np.integer):
            return int(obj)
        if isinstance(obj, np.floating):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return super(_NpEncoder, self).default(obj)


def ascii_tree(nested):
    """The text the reference writes next to every result (run_mcts.py:268-279:
    ``asciitree.LeftAligned()(tree)`` on ``State.render_children()``), restated from asciitree
    0.3.3's LeftAligned drawer with its default BoxStyle (ASCII glyphs, horizontal length 2,
    indent 1, one space before the label): children hang under their parent as `` +-- label``,
    open levels continue with `` |   ``."""
    (label, children), = nested.items()

    def render(node_label, node_children):
        lines = [str(node_label)]
        items = list(node_children.items())
        for i, (lab, sub) in enumerate(items):
            last = i == len(items) - 1
            child = render(lab, sub)
            lines.append("+-- " + child[0])
            lines.extend((("    " if last else "|   ") + line) for line in child[1:])
        return lines
    out = render(label, children)
    return "\n".join([out[0]] + [" " + line for line in out[1:]])


def dump_trees(options, probs, strategies, device):
    """Per (problem, strategy): the tree CustomMCTS.predict builds as JSON (every evaluated child
    of every depth with its score and board, run_mcts.py:227-239) and as the ascii tree text
    (run_mcts.py:268-279)."""
    os.makedirs(options["dump_tree"], exist_ok=True)
    for i, p in enumerate(probs):
        for strategy in strategies:
            m = CustomMCTS(p.entries(), p.board(), strategy=strategy, seed=options["seed"], stream=i,
                           device=device)
            root, depth, found = m.predict(N=options["n_sim"])
            text = ascii_tree(root.render_children(only_ids=False)[0])
            with open(os.path.join(options["dump_tree"], "%d_%d_%d_%s_%d_%d.txt" % (
                    len(p.tiles), p.cols, p.rows, strategy, options["n_sim"], i)), "w") as f:
                f.write(text)
            root.centralize_node_with_children()
            name = "%d_%d_%d_%s_%d_%d_tree.json" % (len(p.tiles), p.cols, p.rows, strategy,
                                                    options["n_sim"], i)
            with open(os.path.join(options["dump_tree"], name), "w") as f:
                json.dump({"depth": depth, "solution_found": found,
                           "n_tiles_placed": m.n_tiles_placed,
                           "solution_tiles_order": m.solution_tiles_order,
                           "tree": root.render_to_json()}, f, cls=_NpEncoder)


def main(argv=None):
    parser = argparse.ArgumentParser(prog="manage.py run_mcts", description="Run mcts")
    add_arguments(parser)
    options = vars(parser.parse_args(argv))
    if options["profile"]:
        import cProfile
        import io
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        pr.runcall(run_mcts, options)
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(pstats.SortKey.CUMULATIVE).print_stats(40)
        print(s.getvalue())
        pr.dump_stats("profile.out")
    else:
        run_mcts(options)


if __name__ == "__main__":
    main()
