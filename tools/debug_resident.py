"""Debug driver of the resident search kernel: runs problem subsets under knob combinations.

    python tools/debug_resident.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(eng, probs, n_sim, flags, label, reps=2):
    from visapi_b200 import datasets
    os.environ["BP_RESIDENT_FLAGS"] = str(flags)
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
            print("%-28s flags=%d %-8s %8.2f ms  solved %d  same=%s" %
                  (label, flags, mode, dt * 1e3, int(res["solution_found"].sum()), same), flush=True)
            if mode == "resident":
                st = s.resident_stats()
                print("      wait %.3f play %.3f advance %.3f other %.3f  units %d advances %d tickets %d" %
                      (st["wait_frac"], st["play_frac"], st["advance_frac"], st["other_frac"],
                       st["units_played"], st["advances"], st["tickets"]), flush=True)
    except Exception as e:   # noqa: BLE001
        print("%-28s flags=%d FAILED: %s" % (label, flags, e), flush=True)
    s.close()


def main():
    import visapi_b200
    from visapi_b200 import datasets
    eng = visapi_b200.Engine(0)
    g = datasets.load_problem_set("g")
    small = [p for p in g if p.rows <= 31 and p.cols <= 32]
    print("class (5,1): %d problems" % len(small), flush=True)
    for flags in (1, 0):
        run(eng, small, 1000, flags, "all of class 5,1", reps=3)
        run(eng, g[:1000], 1000, flags, "g[0:1000]", reps=3)
    run(eng, g[:1000], 100, 0, "g[0:1000] N=100", reps=3)
    run(eng, g[:125], 1000, 0, "g[0:125] (1/8 shard)", reps=3)
    run(eng, g[:1], 1000, 0, "g[0:1]", reps=3)
    eng.close()


if __name__ == "__main__":
    main()
