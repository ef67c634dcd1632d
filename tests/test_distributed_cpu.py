"""world_size-2 gloo runs on CPU of the multi-rank host logic: the sim-range partition, the
all-gather of per-child statistics and the merge rule of the root-parallel search, with the C
oracle playing the rollouts; and the gather of instance-parallel result rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, n_sim, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import c_oracle
    from visapi_b200 import parallel
    rows, cols = case["rows"], case["cols"]
    tiles = [tuple(t) for t in case["tiles"]]
    board = np.zeros((rows, cols))
    legal = c_oracle.valid_next_moves(board, tiles)
    children = [i for i, t in enumerate(tiles) if t in legal]
    begin, n_local = parallel.split_sims(n_sim, rank, world)
    local = np.zeros(len(children), parallel.CHILD_STATS_DTYPE)
    for k, e in enumerate(children):
        v, child = c_oracle.next_turn(board, len(tiles), tiles[e], val=1)
        rest = c_oracle.eliminate_pair(tiles, tiles[e])
        local[k]["first_solved"] = parallel.NONE_U32
        for s in range(begin, begin + n_local):
            r = c_oracle.rollout(child, rest, 123, 7, e, s)
            local[k]["max_steps"] = max(local[k]["max_steps"], r["steps"])
            local[k]["sum_steps"] += r["steps"]
            if local[k]["first_solved"] == parallel.NONE_U32:
                local[k]["prefix_steps"] += r["steps"]
                if r["solved"]:
                    local[k]["first_solved"] = s
    mine = torch.from_numpy(local.view(np.uint8).copy())
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    per_rank = np.stack([g.numpy().view(parallel.CHILD_STATS_DTYPE) for g in gathered])
    merged = parallel.merge_child_stats(per_rank)
    # every rank takes the same first-max decision
    best = int(np.argmax(merged["max_steps"]))
    rows_out = [{"index": i, "rank": rank} for i in parallel.shard_indices(11, rank, world)]
    parts = [None] * world if rank == 0 else None
    dist.gather_object(rows_out, parts, dst=0)
    if rank == 0:
        got = sorted(r["index"] for part in parts for r in part)
        assert got == list(range(11))
    np.save(os.path.join(out_dir, "merged_%d.npy" % rank), merged.view(np.uint8))
    np.save(os.path.join(out_dir, "best_%d.npy" % rank), np.array([best]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_root_parallel_merge_world2(golden, tmp_path):
    from oracle import c_oracle
    from visapi_b200 import parallel
    case = [c for c in golden("flat_search.json") if c["name"] == "gen20_11x11_s0"][0]
    n_sim, world = 37, 2
    mp.spawn(_worker, args=(world, _free_port(), case, n_sim, str(tmp_path)), nprocs=world, join=True)
    merged = [np.load(tmp_path / ("merged_%d.npy" % r)).view(parallel.CHILD_STATS_DTYPE) for r in range(world)]
    np.testing.assert_array_equal(merged[0], merged[1])
    assert np.load(tmp_path / "best_0.npy")[0] == np.load(tmp_path / "best_1.npy")[0]
    # equals what one rank computes over all n_sim rollouts
    rows, cols = case["rows"], case["cols"]
    tiles = [tuple(t) for t in case["tiles"]]
    board = np.zeros((rows, cols))
    legal = c_oracle.valid_next_moves(board, tiles)
    children = [i for i, t in enumerate(tiles) if t in legal]
    for k, e in enumerate(children):
        _, child = c_oracle.next_turn(board, len(tiles), tiles[e], val=1)
        rest = c_oracle.eliminate_pair(tiles, tiles[e])
        steps = [c_oracle.rollout(child, rest, 123, 7, e, s) for s in range(n_sim)]
        assert merged[0][k]["max_steps"] == max(r["steps"] for r in steps)
        assert merged[0][k]["sum_steps"] == sum(r["steps"] for r in steps)
        hit = [s for s, r in enumerate(steps) if r["solved"]]
        if hit:
            assert merged[0][k]["first_solved"] == hit[0]
            assert merged[0][k]["prefix_steps"] == sum(r["steps"] for r in steps[:hit[0] + 1])
        else:
            assert merged[0][k]["first_solved"] == parallel.NONE_U32


def _puct_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from oracle import py_oracle
    from visapi_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank searches its OWN tree from the same root (CPU oracle tree, deterministic stub
    # net); a different number of simulations per rank stands in for different random streams
    h, w, n = 8, 8, 6
    tiles = np.array([[2, 2], [4, 2], [2, 4], [4, 4], [2, 8], [8, 2], [2, 2], [2, 4], [4, 2], [4, 4], [8, 2], [2, 8]])

    def predict(grid, _tiles):
        filled = int(np.count_nonzero(np.asarray(grid)))
        pi = np.array([float((a * 7 + filled * 3) % 11 + 1) for a in range(2 * n)])
        return pi / pi.sum(), (filled % 8) / 8.0
    tree = py_oracle.PuctTree(py_oracle.Board(h, w, tiles.tolist(), n), predict, 0, 1.0)
    root = (np.zeros((h, w)), tiles.copy())
    for _ in range(30 + 17 * rank):
        tree.search(root)
    counts, wsa = tree.root_stats(root)
    c = torch.tensor(counts, dtype=torch.int32).view(1, -1)
    wv = torch.tensor(wsa, dtype=torch.float64).view(1, -1)
    np.save(os.path.join(out_dir, "own_%d.npy" % rank), np.stack([c.numpy()[0].astype(np.float64), wv.numpy()[0]]))
    mc, mw, mq = parallel.allreduce_root_stats(c, wv)
    probs = parallel.probs_from_counts(mc.numpy(), 1)
    np.save(os.path.join(out_dir, "merged_puct_%d.npy" % rank),
            np.stack([mc.numpy()[0].astype(np.float64), mw.numpy()[0], mq.numpy()[0], probs[0]]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_puct_root_parallel_allreduce_world2(tmp_path):
    """SURVEY 8e row 3 on two gloo ranks: root Nsa / Wsa summed over ranks, policy from the sum."""
    world = 2
    mp.spawn(_puct_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    own = [np.load(tmp_path / ("own_%d.npy" % r)) for r in range(world)]
    merged = [np.load(tmp_path / ("merged_puct_%d.npy" % r)) for r in range(world)]
    np.testing.assert_array_equal(merged[0], merged[1])                  # every rank holds the same merge
    np.testing.assert_array_equal(merged[0][0], own[0][0] + own[1][0])   # counts: sum rule
    np.testing.assert_allclose(merged[0][1], own[0][1] + own[1][1])      # value sums
    assert not np.array_equal(own[0][0], own[1][0])                      # the ranks' trees differ
    c = merged[0][0]
    np.testing.assert_allclose(merged[0][3], c / c.sum())
    np.testing.assert_allclose(merged[0][2] * c, merged[0][1])
