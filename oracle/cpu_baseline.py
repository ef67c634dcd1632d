"""CPU baseline: the oracle timed on the host cores.  TEST/BENCH INFRASTRUCTURE.

Only bench.py's ``cpu_baseline`` leg and ``--impl reference`` arm run this (as a
subprocess, so that no CUDA context is forked):

    python -m oracle.cpu_baseline --set g --n-sim 1000 --strategy max_depth \
        --budget 20 --procs 0 --c-problems 16

Two measurements of the same workload (flat search over the problem list, in order):
  python port  oracle/py_oracle.py — same data types and per-step costs as the
               reference's Python (float64 numpy boards, list scans, board copies);
               one single-threaded process per host core, problems dealt round-robin,
               each process stops when the wall-clock budget expires (a bounded
               sample: the rollouts finished by then are what is counted).
  C port       oracle/rectpack_oracle.c — scalar C restatement, one pthread per core,
               on the first ``--c-problems`` problems.
Prints one JSON object.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class _Budget(Exception):
    pass


def _load(set_name):
    z = np.load(os.path.join(ROOT, "visapi_b200", "data", "problems_%s.npz" % set_name))
    probs = []
    for i in range(len(z["rows"])):
        n = int(z["n_tiles"][i])
        probs.append((int(z["rows"][i]), int(z["cols"][i]), z["tiles"][i, :n].tolist()))
    return probs


def _py_worker(args):
    worker, n_workers, set_name, n_sim, strategy, seed, budget = args
    from oracle import py_oracle
    probs = _load(set_name)
    count = {"rollouts": 0, "placements": 0, "problems": 0}
    deadline = time.time() + budget
    real_rollout = py_oracle.rollout

    def counted_rollout(grid, tiles, stream, val=1):
        if time.time() > deadline:
            raise _Budget()
        res = real_rollout(grid, tiles, stream, val)
        count["rollouts"] += 1
        count["placements"] += res[2]
        return res

    py_oracle.rollout = counted_rollout
    t0 = time.time()
    try:
        for i in range(worker, len(probs), n_workers):
            rows, cols, tiles = probs[i]
            py_oracle.flat_search(py_oracle.sorted_both_orientations(tiles), np.zeros((rows, cols)),
                                  n_sim, strategy, seed, i)
            count["problems"] += 1
    except _Budget:
        pass
    finally:
        py_oracle.rollout = real_rollout
    count["seconds"] = time.time() - t0
    return count


def python_port(set_name, n_sim, strategy, seed, budget, procs):
    procs = procs or (os.cpu_count() or 1)
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        t0 = time.time()
        parts = pool.map(_py_worker, [(w, procs, set_name, n_sim, strategy, seed, budget)
                                      for w in range(procs)])
        wall = time.time() - t0
    rollouts = sum(p["rollouts"] for p in parts)
    placements = sum(p["placements"] for p in parts)
    busy = max(p["seconds"] for p in parts)
    return {"rollouts": rollouts, "placements": placements, "seconds": busy, "wall": wall,
            "cores": procs, "rollouts_per_s": rollouts / busy, "placements_per_s": placements / busy,
            "problems_finished": sum(p["problems"] for p in parts)}


def c_port(set_name, n_sim, strategy, seed, n_problems, threads):
    from oracle import c_oracle, py_oracle
    probs = _load(set_name)[:n_problems]
    E = max(2 * len(p[2]) for p in probs)
    tiles = np.zeros((len(probs), E, 2), np.int32)
    n_entries = []
    for i, (_, _, tl) in enumerate(probs):
        ent = py_oracle.sorted_both_orientations(tl)
        tiles[i, :len(ent)] = np.asarray(ent, np.int32)
        n_entries.append(len(ent))
    threads = threads or c_oracle.num_threads()
    t0 = time.time()
    res = c_oracle.flat_search_batch([p[0] for p in probs], [p[1] for p in probs], tiles, n_entries,
                                     n_sim, strategy, seed, list(range(len(probs))), threads)
    dt = time.time() - t0
    rollouts = sum(r["rollouts"] for r in res)
    placements = sum(r["rollout_placements"] for r in res)
    return {"rollouts": rollouts, "placements": placements, "seconds": dt, "cores": threads,
            "problems": len(probs), "rollouts_per_s": rollouts / dt,
            "placements_per_s": placements / dt,
            "solved": sum(1 for r in res if r["solution_found"])}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--set", default="g")
    ap.add_argument("--n-sim", type=int, default=1000)
    ap.add_argument("--strategy", default="max_depth")
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--budget", type=float, default=20.0, help="seconds per python-port sample")
    ap.add_argument("--procs", type=int, default=0, help="0 = all host cores")
    ap.add_argument("--c-problems", type=int, default=16, help="0 = skip the C port")
    ap.add_argument("--skip-python", action="store_true")
    a = ap.parse_args()
    out = {"cpu_model": cpu_model(), "host_cores": os.cpu_count()}
    if not a.skip_python:
        out["python_port"] = python_port(a.set, a.n_sim, a.strategy, a.seed, a.budget, a.procs)
    if a.c_problems > 0:
        out["c_port"] = c_port(a.set, a.n_sim, a.strategy, a.seed, a.c_problems, a.procs)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
