#!/usr/bin/env python
"""Config 5: BinPackGame(20, 20, 15), numMCTSSims = 50, cpuct = 1 (alpha/main.py:16-37)
with a fixed-weights PyTorch net: searches/s and leaf-evaluation batch size of the lock-step
batched self-play, and of the one-tree-at-a-time MCTS for comparison."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=256)
    ap.add_argument("--sims", type=int, default=50)
    ap.add_argument("--host-leaves", action="store_true", help="evaluate leaves through host copies")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from visapi_b200 import runtime
    runtime.set_default_device(local)
    from visapi_b200.coach import BatchedSelfPlay, Coach
    from visapi_b200.game import BinPackGame
    from visapi_b200.nnet import TorchBinPackNet
    np.random.seed(rank)
    game = BinPackGame(20, 20, 15, device=local)
    net = TorchBinPackNet(game, device="cuda:%d" % local)
    args = {"numMCTSSims": a.sims, "cpuct": 1.0, "tempThreshold": 15}
    sp = BatchedSelfPlay(20, 20, 15, net, args, n_games=a.games, device=local, seed=rank,
                         device_leaves=not a.host_leaves)
    sp.play()                                            # warm-up (cudnn autotune, allocations)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    examples, reward, searches = sp.play()
    torch.cuda.synchronize()
    dt = time.time() - t0
    if world > 1:                                        # replicas: sum the work, max the time
        agg = torch.tensor([float(searches), float(len(examples))], device="cuda")
        tt = torch.tensor([dt], device="cuda")
        dist.all_reduce(agg)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        searches_all, dt_all = float(agg[0].item()), float(tt.item())
    else:
        searches_all, dt_all = float(searches), dt
    evals, batches = sp.search.leaf_evals, sp.search.leaf_batches
    coach = Coach(game, net, args, device=local)
    t1 = time.time()
    n_single = 0
    for _ in range(2):
        ex = coach.executeEpisode()
        n_single += len(ex) * a.sims
    dt1 = time.time() - t1
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    print(json.dumps({
        "config": "BinPackGame(20,20,15), numMCTSSims=%d, cpuct=1" % a.sims, "games": a.games,
        "n_gpus": world, "aggregate_searches_per_s": searches_all / dt_all,
        "leaf_path": "host" if a.host_leaves else "device",
        "batched_searches_per_s": searches / dt, "batched_seconds": dt,
        "examples": len(examples), "mean_reward": float(reward.mean()),
        "mean_leaf_batch": evals / max(batches, 1), "leaf_batches": batches,
        "single_tree_searches_per_s": n_single / dt1}))


if __name__ == "__main__":
    main()
