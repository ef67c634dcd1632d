#!/usr/bin/env python
"""Solve rate, mean score and mean n_tiles_placed of the full G and B sets at every N and
strategy the reference published (csv_results/results_{g,b}.csv), next to the published
aggregates (tests/golden/published.json).  Prints a JSON table; takes a few seconds on a B200."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from visapi_b200 import Engine, datasets
    from visapi_b200.verify import replay_solution
    pub = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "published.json")))
    eng = Engine(0)
    out = []
    for set_name in ("g", "b"):
        probs = datasets.load_problem_set(set_name)
        batch = datasets.batch_arrays(probs)
        for strategy in ("max_depth", "avg_depth"):
            for n_sim in (100, 200, 500, 1000, 2000, 5000):
                t0 = time.time()
                res = eng.flat_search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"],
                                      n_sim, 0 if strategy == "max_depth" else 1, 123)
                dt = time.time() - t0
                bad = 0
                for i, p in enumerate(probs):
                    if res["solution_found"][i]:
                        ok, _ = replay_solution(p.rows, p.cols, p.tiles,
                                                res["order"][i, :res["n_order"][i]].tolist())
                        bad += 0 if ok else 1
                pb = pub[set_name]["%s/%d" % (strategy, n_sim)]
                out.append({"set": set_name, "strategy": strategy, "n_sim": n_sim,
                            "solved": int(res["solution_found"].sum()), "published_solved": pb["solved"],
                            "invalid_solutions": bad,
                            "mean_score": float(res["score"].mean()), "published_mean_score": pb["mean_score"],
                            "mean_n_tiles_placed": float(res["n_tiles_placed"].mean()),
                            "published_mean_n_tiles_placed": pb["mean_n_tiles_placed"],
                            "rollouts": int(res["rollouts"].sum()), "wall_s": dt})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
