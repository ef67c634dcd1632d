#!/usr/bin/env python
"""Config 4: 50 tiles on 25x25 and 40x40 boards from the generator, seeds
0..7, N = 5000 rollouts per child, root-parallel: every rank plays N/G rollouts of every
child; the per-child statistics are merged once per committed move, inside the resident kernels
over NVLink peer memory (--exchange resident) or by an NCCL all-gather between launches
(--exchange nccl).  Times reset + search of a search object created once.  Run alone or under
torchrun; prints one JSON line (rank 0) with the time per instance and the proof that
the result does not depend on the number of ranks (a digest of all outputs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29541 tools/bench_root_parallel.py
"""
import hashlib
import json
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def instance(n, rows, cols, seed):
    from visapi_b200 import datasets
    from visapi_b200.data_generator import DataGenerator
    random.seed(seed)
    tiles, board = DataGenerator().gen_tiles_and_board(n, rows, cols, order_tiles=True)
    pool, base = [tuple(int(v) for v in t) for t in tiles], []
    while pool:
        t = pool.pop(0)
        pool.remove((t[1], t[0]))
        base.append(t)
    return datasets.Problem(board.shape[0], board.shape[1], base, "gen50_%dx%d_s%d" % (rows, cols, seed))


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-sim", type=int, default=5000)
    ap.add_argument("--seeds", type=int, default=8)
    ap.add_argument("--exchange", default="resident", choices=["resident", "nccl"])
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    from visapi_b200 import Engine, datasets, parallel
    from visapi_b200.verify import replay_solution
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local, stream=torch.cuda.current_stream().cuda_stream)
    n_sim = a.n_sim
    out = {"world": world, "n_sim": n_sim, "exchange": a.exchange, "boards": {}}
    digest = hashlib.sha256()
    for (rows, cols) in ((25, 25), (40, 40)):
        times, depths, found, rollouts = [], [], 0, 0
        for seed in range(a.seeds):
            p = instance(50, rows, cols, seed)
            batch = datasets.batch_arrays([p])
            rp = parallel.RootParallelSearch(eng, batch, n_sim, 0, 123, np.array([seed], np.uint32),
                                             exchange=a.exchange)
            best = None
            for rep in range(a.reps + 1):                          # the first run warms up
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rp.run()                                          # reset + the whole search
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rep > 0:
                    best = float(t.item()) if best is None else min(best, float(t.item()))
            r = rp.results()
            rp.close()
            times.append(best)
            depths.append(int(r["depth"][0]))
            rollouts += int(r["rollouts"][0])
            if r["solution_found"][0]:
                found += 1
                ok, why = replay_solution(p.rows, p.cols, p.tiles, r["order"][0, :r["n_order"][0]].tolist())
                assert ok, why
            for k in ("depth", "solution_found", "n_tiles_placed", "rollouts", "order", "path"):
                digest.update(np.ascontiguousarray(r[k]).tobytes())
        out["boards"]["%dx%d" % (rows, cols)] = {
            "ms_per_instance": float(np.mean(times)), "ms_min": min(times), "ms_max": max(times),
            "mean_depth": float(np.mean(depths)), "solved": found,
            "rollouts_per_s": rollouts / (sum(times) / 1e3)}
    out["result_digest"] = digest.hexdigest()[:16]
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
