#!/usr/bin/env python
"""One flat search over the problems of ONE shape class, for ncu.

    python tools/profile_rollout.py --cls 1,1,2 --n-problems 200 --n-sim 1000

Prints a JSON line with the placements / rollouts the device executed, so that
ncu's smsp__inst_executed.sum over the fs_rollout_kernel launches of the same command
gives warp instructions per placement.  Exactly one step, no
warm-up: every fs_rollout_kernel launch in the ncu log belongs to that step.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def cls_of(p):
    def rc(v):
        return 1 if v <= 1 else (2 if v <= 2 else 4)
    return "%d,%d,%d" % (rc((p.rows + 31) // 32), rc((p.cols + 31) // 32), rc((2 * len(p.tiles) + 31) // 32))


def lanes_cls_of(p):
    nb = 5 if p.rows <= 31 else (6 if p.rows <= 63 else (7 if p.rows <= 127 else 8))
    w = 1 if p.cols <= 32 else (2 if p.cols <= 64 else (3 if p.cols <= 96 else 4))
    ew = 2 if 2 * len(p.tiles) <= 64 else 4
    return "%d,%d,%d" % (nb, w, ew)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="g")
    ap.add_argument("--cls", default="all", help="warp class RHO,W,EJ or 'all'")
    ap.add_argument("--lcls", default="all", help="lane class NB,W,EW or 'all'")
    ap.add_argument("--n-problems", type=int, default=200)
    ap.add_argument("--n-sim", type=int, default=1000)
    ap.add_argument("--strategy", default="max_depth")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--mode", default="auto", choices=["auto", "stepped", "resident"])
    a = ap.parse_args()
    from visapi_b200 import Engine, datasets
    probs = datasets.load_problem_set(a.set)
    if a.cls != "all":
        probs = [p for p in probs if cls_of(p) == a.cls]
    if a.lcls != "all":
        probs = [p for p in probs if lanes_cls_of(p) == a.lcls]
    probs = probs[:a.n_problems]
    batch = datasets.batch_arrays(probs)
    eng = Engine(0)
    s = eng.search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"], a.n_sim,
                   0 if a.strategy == "max_depth" else 1, 123, np.arange(len(probs), dtype=np.uint32))
    s.set_mode(a.mode)
    s.set_profiling(True)
    for _ in range(a.steps):
        s.reset()
        s.run()
    rollout_ms, advance_ms, n_depths = s.profile()
    r = s.results()
    print(json.dumps({"cls": a.cls, "lcls": a.lcls, "problems": len(probs), "n_sim": a.n_sim,
                      "placements_executed": int(r["placements_executed"].sum()),
                      "rollouts_executed": int(r["rollouts_executed"].sum()),
                      "rollouts": int(r["rollouts"].sum()), "solved": int(r["solution_found"].sum()),
                      "mode": a.mode, "resident": s.last_run_resident(), "rollout_ms": rollout_ms, "advance_ms": advance_ms, "depths": n_depths}))


if __name__ == "__main__":
    main()
