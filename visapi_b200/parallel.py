"""Multi-GPU drivers: one process per GPU (torch.distributed), two partitionings.

* instance-parallel — independent problems are dealt to ranks (problem i -> rank
  i mod N); no collective on the data path, results are gathered on rank 0 at the end.
* root-parallel — one (or a few) hard instances: every rank plays a contiguous slice
  of the N rollouts of EVERY child; once per committed move the per-child statistics
  (max depth, depth sum, first solved rollout, placements before it: 24 bytes per
  child) of all ranks are merged and every rank takes the same first-max decision.
  Because each rollout owns a counter-based Philox stream, the result is bit-identical
  to the single-GPU search for any number of ranks.  Two forms of the exchange:
  ``root_parallel_search`` (default, "resident"): ONE kernel launch per rank plays the whole
  search; the warp that finishes a problem's depth stores its rank's statistics into every
  rank's buffer over NVLink peer memory (CUDA IPC mappings), raises a flag and waits for
  the others' flags inside the kernel — no host step, no NCCL call per move.
  ``root_parallel_search(..., exchange="nccl")``: launch-per-depth kernels with one
  ``all_gather_into_tensor`` per move (the round-1 form, kept as the cross-check).
"""
import numpy as np

CHILD_STATS_DTYPE = np.dtype([("max_steps", "<i4"), ("first_solved", "<u4"),
                              ("sum_steps", "<i8"), ("prefix_steps", "<i8")])
NONE_U32 = 0xFFFFFFFF


def shard_indices(n_items, rank, world):
    """Problem i goes to rank i mod world."""
    return list(range(rank, n_items, world))


def split_sims(n_sim, rank, world):
    """Contiguous ascending slices of the n_sim rollouts: (sim_begin, n_local)."""
    base, rem = divmod(n_sim, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def merge_child_stats(per_rank):
    """Host statement of the merge rule the commit kernel applies to the gathered
    statistics (visapi_b200/csrc/search_kernels.cuh: stats_from_gathered).  per_rank: array
    [world, ...] of CHILD_STATS_DTYPE, rank-major, ranks holding ascending sim ranges."""
    per_rank = np.asarray(per_rank)
    out = np.zeros(per_rank.shape[1:], CHILD_STATS_DTYPE)
    out["first_solved"] = NONE_U32
    for g in range(per_rank.shape[0]):
        r = per_rank[g]
        out["max_steps"] = np.maximum(out["max_steps"], r["max_steps"])
        out["sum_steps"] += r["sum_steps"]
        open_ = out["first_solved"] == NONE_U32
        out["prefix_steps"] = np.where(open_, out["prefix_steps"] + r["prefix_steps"],
                                       out["prefix_steps"])
        out["first_solved"] = np.where(open_, r["first_solved"], out["first_solved"])
    return out


def allreduce_root_stats(counts, wsa, group=None):
    """Root-parallel PUCT: every rank searched its own trees from the same
    roots; the root visit counts Nsa and value sums Wsa = Qsa * Nsa of all ranks are summed in
    place (torch tensors on any device) before the policy is formed from the counts
    (binpack/MCTS.py:221-244).  Returns (counts, wsa, q) with q = wsa / counts (0 where unvisited)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(wsa, op=dist.ReduceOp.SUM, group=group)
    q = torch.where(counts > 0, wsa / counts.clamp(min=1).to(wsa.dtype), torch.zeros_like(wsa))
    return counts, wsa, q


def probs_from_counts(counts, temp=1):
    """MCTS.getActionProb's policy (binpack/MCTS.py:236-244) from (summed) visit counts [B, E]."""
    counts = np.asarray(counts, dtype=np.float64)
    if temp == 0:
        probs = np.zeros_like(counts)
        probs[np.arange(counts.shape[0]), counts.argmax(axis=1)] = 1
        return probs
    c = counts ** (1. / temp)
    total = c.sum(axis=1, keepdims=True)
    return np.divide(c, total, out=np.zeros_like(c), where=total > 0)


class _DevicePointer(object):
    """Exposes a raw device pointer through __cuda_array_interface__ (for torch.as_tensor)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}


class RootParallelSearch(object):
    """A root-parallel flat search kept across runs.  Every rank must construct it, ``run()`` it
    and ``close()`` it together, on the CUDA stream the engine was given.

    exchange="resident": partition + peer-memory exchange buffers (CUDA IPC handles swapped once
    over torch.distributed); ``run()`` = reset + ONE resident kernel launch per rank, the
    per-move merge happens inside the kernels over NVLink.
    exchange="nccl": launch-per-depth kernels with one ``all_gather_into_tensor`` per move."""

    def __init__(self, engine, batch, n_sim, strategy, seed, streams=None, group=None,
                 exchange="resident"):
        import torch
        import torch.distributed as dist

        if exchange not in ("resident", "nccl"):
            raise ValueError("exchange must be 'resident' or 'nccl'")
        self.exchange = exchange
        self.group = group
        self.dist = dist if dist.is_initialized() else None
        world = dist.get_world_size(group) if self.dist else 1
        rank = dist.get_rank(group) if self.dist else 0
        self.world, self.rank = world, rank
        self.search = engine.search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"],
                                    n_sim, strategy, seed, streams)
        begin, n_local = split_sims(n_sim, rank, world)
        if exchange == "resident":
            self.search.set_mode("resident")
            if world > 1:
                self.search.set_partition(rank, world, begin, n_local)
                handle, _ = self.search.exchange_create()
                mine = torch.from_numpy(handle).cuda()
                every = torch.empty(world * 64, dtype=torch.uint8, device=mine.device)
                dist.all_gather_into_tensor(every, mine, group=group)
                self.search.exchange_connect(ipc_handles=every.cpu().numpy().reshape(world, 64))
                dist.barrier(group=group)
        else:
            self.search.set_mode("stepped")
            self.search.set_partition(rank, world, begin, n_local)
            ptr, nbytes = self.search.stats_buffer()
            self._local = torch.as_tensor(_DevicePointer(ptr, nbytes), device="cuda")
            self._gathered = torch.empty(world * nbytes, dtype=torch.uint8, device=self._local.device)

    def run(self):
        s = self.search
        s.reset()
        if self.exchange == "resident":
            s.run()
            return
        for _ in range(s.max_depth):
            s.step_rollouts()
            s.step_reduce()
            if self.world > 1:
                self.dist.all_gather_into_tensor(self._gathered, self._local, group=self.group)
            else:
                self._gathered.copy_(self._local)
            s.step_commit(self._gathered.data_ptr())

    def results(self, want_scores=False):
        return self.search.results(want_scores=want_scores)

    def close(self):
        if self.search is not None:
            if self.dist and self.world > 1:
                import torch
                torch.cuda.synchronize()
                self.dist.barrier(group=self.group)      # nobody stores into a freed buffer
            self.search.close()
            self.search = None


def root_parallel_search(engine, batch, n_sim, strategy, seed, streams=None, group=None,
                         want_scores=False, profile=None, exchange="resident"):
    """Flat search of ``batch`` (datasets.batch_arrays) with the rollouts of every child split
    over the ranks of ``group``.  Must be called by every rank, on the CUDA stream the engine
    was given (torch's current stream).  Returns the same dict as Engine.flat_search."""
    rp = RootParallelSearch(engine, batch, n_sim, strategy, seed, streams, group, exchange)
    try:
        rp.run()
        return rp.results(want_scores=want_scores)
    finally:
        rp.close()
