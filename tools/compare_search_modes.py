"""The launch-per-depth search against the resident (single-launch) search kernel on several
workloads: wall time per search, equality of the results, and where the resident kernel's warps
spent their time (profiles/r1/resident_vs_stepped.txt).

    python tools/compare_search_modes.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(eng, probs, n_sim, label, reps=3):
    from visapi_b200 import datasets
    batch = datasets.batch_arrays(probs)
    s = eng.search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"], n_sim, 0, 123)
    out = None
    try:
        for mode in ("stepped", "resident"):
            s.set_mode(mode)
            for _ in range(reps):
                s.reset()
                eng.synchronize()
                t0 = time.perf_counter()
                s.run()
                eng.synchronize()
                dt = time.perf_counter() - t0
            res = s.results()
            if out is None:
                out = res
            same = all(np.array_equal(res[k], out[k]) for k in ("depth", "solution_found", "rollouts", "order"))
            print("%-28s %-8s %8.2f ms  solved %d  same=%s" %
                  (label, mode, dt * 1e3, int(res["solution_found"].sum()), same), flush=True)
            if mode == "resident":
                st = s.resident_stats()
                print("      wait %.3f play %.3f advance %.3f other %.3f  units %d advances %d polls %d" %
                      (st["wait_frac"], st["play_frac"], st["advance_frac"], st["other_frac"],
                       st["units_played"], st["advances"], st["polls"]), flush=True)
    except Exception as e:   # noqa: BLE001
        print("%-28s FAILED: %s" % (label, e), flush=True)
    s.close()


def main():
    import visapi_b200
    from visapi_b200 import datasets
    eng = visapi_b200.Engine(0)
    g = datasets.load_problem_set("g")
    small = [p for p in g if p.rows <= 31 and p.cols <= 32]
    print("class (5,1): %d problems" % len(small), flush=True)
    run(eng, small, 1000, "all of class 5,1")
    run(eng, g[:1000], 1000, "g[0:1000]")
    run(eng, g[:1000], 100, "g[0:1000] N=100")
    run(eng, datasets.load_problem_set("b")[:1000], 1000, "b[0:1000]")
    run(eng, g[:125], 1000, "g[0:125] (1/8 shard)")
    run(eng, g[:1], 1000, "g[0:1]")
    eng.close()


if __name__ == "__main__":
    main()
