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
    a = ap.parse_args()
    import torch
    from visapi_b200.coach import BatchedSelfPlay, Coach
    from visapi_b200.game import BinPackGame
    from visapi_b200.nnet import TorchBinPackNet
    np.random.seed(0)
    game = BinPackGame(20, 20, 15)
    net = TorchBinPackNet(game)
    args = {"numMCTSSims": a.sims, "cpuct": 1.0, "tempThreshold": 15}
    sp = BatchedSelfPlay(20, 20, 15, net, args, n_games=a.games, seed=0)
    sp.play()                                            # warm-up (cudnn autotune, allocations)
    torch.cuda.synchronize()
    t0 = time.time()
    examples, reward, searches = sp.play()
    torch.cuda.synchronize()
    dt = time.time() - t0
    evals, batches = sp.search.leaf_evals, sp.search.leaf_batches
    coach = Coach(game, net, args)
    t1 = time.time()
    n_single = 0
    for _ in range(2):
        ex = coach.executeEpisode()
        n_single += len(ex) * a.sims
    dt1 = time.time() - t1
    print(json.dumps({
        "config": "BinPackGame(20,20,15), numMCTSSims=%d, cpuct=1" % a.sims, "games": a.games,
        "batched_searches_per_s": searches / dt, "batched_seconds": dt,
        "examples": len(examples), "mean_reward": float(reward.mean()),
        "mean_leaf_batch": evals / max(batches, 1), "leaf_batches": batches,
        "single_tree_searches_per_s": n_single / dt1}))


if __name__ == "__main__":
    main()
