#!/usr/bin/env python
"""Run under torchrun on N >= 2 GPUs: root-parallel and instance-parallel results must be
bit-identical to the single-rank search (every rollout owns its Philox stream).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/check_multi_gpu.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from visapi_b200 import Engine, datasets, parallel
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(local, stream=torch.cuda.current_stream().cuda_stream)
    probs = datasets.load_problem_set("g")[:24] + datasets.load_problem_set("b")[:8]
    batch = datasets.batch_arrays(probs)
    out = {}
    for strategy in (0, 1):
        for n_sim in (100, 333):
            single = eng.flat_search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"],
                                     n_sim, strategy, 123, np.arange(len(probs), dtype=np.uint32),
                                     want_scores=True)
            for exchange in ("resident", "nccl"):     # in-kernel peer-memory exchange / NCCL all-gather
                rp = parallel.root_parallel_search(eng, batch, n_sim, strategy, 123,
                                                   np.arange(len(probs), dtype=np.uint32), want_scores=True,
                                                   exchange=exchange)
                for k in ("depth", "solution_found", "n_tiles_placed", "rollouts", "rollout_placements",
                          "n_order", "order", "path", "child_scores"):
                    assert np.array_equal(single[k], rp[k]), (exchange, strategy, n_sim, k)
            # instance-parallel: my shard, compared with the same rows of the full run
            mine = parallel.shard_indices(len(probs), rank, world)
            sub = datasets.batch_arrays([probs[i] for i in mine], entry_stride=batch["tiles"].shape[1])
            part = eng.flat_search(sub["rows"], sub["cols"], sub["tiles"], sub["n_entries"], n_sim,
                                   strategy, 123, np.asarray(mine, dtype=np.uint32))
            for k in ("depth", "solution_found", "n_tiles_placed", "rollouts", "order"):
                assert np.array_equal(part[k], single[k][mine]), (strategy, n_sim, k)
            out["%d/%d" % (strategy, n_sim)] = int(single["solution_found"].sum())
    # timing of the root-parallel exchange on one hard instance (C4-like: 50 tiles, 40 x 40)
    import random
    from visapi_b200.data_generator import DataGenerator
    random.seed(4)
    tiles, board = DataGenerator().gen_tiles_and_board(50, 40, 40, order_tiles=True)
    pool, base = [tuple(int(v) for v in t) for t in tiles], []
    while pool:
        t = pool.pop(0)
        pool.remove((t[1], t[0]))
        base.append(t)
    hard = datasets.batch_arrays([datasets.Problem(40, 40, base)])
    for n_sim in (5000,):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = parallel.root_parallel_search(eng, hard, n_sim, 0, 123, np.zeros(1, np.uint32))
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["root_parallel_50tiles_40x40_N%d_ms" % n_sim] = float(t.item())
        out["root_parallel_rollouts"] = int(r["rollouts"][0])
        out["root_parallel_depth"] = int(r["depth"][0])
    if rank == 0:
        print(json.dumps({"ok": True, "world": world, **out}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
