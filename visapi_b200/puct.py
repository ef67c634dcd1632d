"""Drop-in PUCT ``MCTS`` (reference: binpack/MCTS.py:204-335, alpha/MCTS.py:302-434).

The tree lives on the device (visapi_b200/csrc/puct.cu): struct-of-arrays visit /
value / prior tables in HBM, select and backup as kernels, nodes keyed by the exact
(board, tile list) state exactly like the reference's string keys.  The neural net
stays where the reference has it — ``nnet.predict`` is called by the host on every
leaf, and its policy is masked with the valid moves and renormalised next to the call
(MCTS.py:281-292).  ``BatchedMCTS`` advances many games in lock step and evaluates all
their leaves as one batch.
"""
import numpy as np

from . import runtime
from ._native import Tree, pack_occupancy, pack_tiles, unpack_occupancy

EPS = 0.1


def _mask_and_normalise_rows(pi, valids):
    """Row-wise form of _mask_and_normalise for a batch of leaves (same float64 operations
    per row: multiply, sum along the row, divide)."""
    p = np.asarray(pi, dtype=np.float64) * valids
    total = p.sum(axis=1, keepdims=True)
    masked = total[:, 0] <= 0
    if masked.any():
        p[masked] = p[masked] + valids[masked]
        total[masked] = p[masked].sum(axis=1, keepdims=True)
    return p / total


def _mask_and_normalise(pi, valids):
    """Ps = Ps * valids; renormalise; uniform over valid moves when everything is masked
    (binpack/MCTS.py:281-292)."""
    p = np.asarray(pi, dtype=np.float64) * valids
    total = np.sum(p)
    if total > 0:
        p = p / total
    else:
        p = p + valids
        p = p / np.sum(p)
    return p


class MCTS(object):
    """One tree.  ``args`` needs numMCTSSims and cpuct; ``args.get('maxNodes')`` bounds the
    nodes kept across calls (default: enough for a whole episode)."""

    predict_on_grid_only = True        # binpack/MCTS.py:277 passes canonicalBoard[0]

    def __init__(self, game, nnet, args, device=None):
        self.game, self.nnet, self.args = game, nnet, args
        self.device = device
        self.n_actions = game.getActionSize()
        sims = int(args['numMCTSSims'] if isinstance(args, dict) else args.numMCTSSims)
        cpuct = float(args['cpuct'] if isinstance(args, dict) else args.cpuct)
        max_nodes = None
        if isinstance(args, dict):
            max_nodes = args.get('maxNodes')
        if not max_nodes:
            max_nodes = (sims + 2) * (self.n_actions // 2 + 2) * 2
        self.num_sims = sims
        h, w = game.getBoardSize()[0], game.getBoardSize()[1]
        self.height, self.width = h, w
        self._tree = Tree(runtime.get_engine(device), 1, h, w, self.n_actions, max_nodes, cpuct)
        self._root_key = None
        self.leaf_evals = 0

    def _pack(self, board):
        occ = pack_occupancy(np.asarray(board[0])[None], colorful=True)
        tiles = pack_tiles([[(int(t[0]), int(t[1])) for t in board[1]]], self.n_actions)
        return occ, tiles

    def _set_root(self, board):
        occ, tiles = self._pack(board)
        key = (occ.tobytes(), tiles.tobytes())
        if key != self._root_key:
            self._tree.set_root(occ, tiles)
            self._root_key = key

    def search(self, canonicalBoard):
        """One simulation from ``canonicalBoard``; returns the value propagated to it."""
        self._set_root(canonicalBoard)
        status, occ, tiles, valid = self._tree.select()
        priors = np.zeros((1, self.n_actions))
        values = np.zeros(1)
        if status[0] == 0:
            grid = unpack_occupancy(occ, self.width)[0].astype(np.float64)
            if self.predict_on_grid_only:
                pi, v = self.nnet.predict(grid)
            else:
                pi, v = self.nnet.predict((grid, tiles[0].astype(np.int64)))
            self.leaf_evals += 1
            priors[0] = _mask_and_normalise(pi, valid[0].astype(np.float64))
            values[0] = v
        return float(self._tree.backup(priors, values)[0])

    def getActionProb(self, canonicalBoard, temp=1):
        """numMCTSSims simulations, then the visit-count policy (MCTS.py:221-244)."""
        for _ in range(self.num_sims):
            self.search(canonicalBoard)
        self._set_root(canonicalBoard)
        counts = [int(c) for c in self._tree.root_counts()[0][0]]
        if temp == 0:
            best = int(np.argmax(counts))
            probs = [0] * len(counts)
            probs[best] = 1
            return probs
        counts = [x ** (1. / temp) for x in counts]
        total = float(sum(counts))
        return [x / total for x in counts]

    def n_nodes(self):
        return int(self._tree.root_counts()[1][0])

    def reset(self):
        self._tree.reset()
        self._root_key = None


class BatchedMCTS(object):
    """B games of one shape searched in lock step; ``predict_batch(grids [B', H, W]) ->
    (pi [B', 2n], v [B'])`` is called once per simulation on the leaves of all trees."""

    def __init__(self, height, width, n_actions, predict_batch, num_sims, cpuct, n_trees,
                 max_nodes=None, device=None, with_tiles=False):
        self.height, self.width, self.n_actions = height, width, n_actions
        self.predict_batch = predict_batch
        self.num_sims, self.B = int(num_sims), int(n_trees)
        if not max_nodes:
            max_nodes = (self.num_sims + 2) * (n_actions // 2 + 2) * 2
        self._tree = Tree(runtime.get_engine(device), self.B, height, width, n_actions, max_nodes,
                          cpuct)
        self.leaf_evals = 0
        self.leaf_batches = 0
        self.with_tiles = with_tiles       # predict_batch(grids, tiles) instead of (grids)

    def set_roots(self, grids, tile_tables):
        """grids [B, H, W] (non-zero = occupied); tile_tables [B, 2n, 2]."""
        self._tree.set_root(pack_occupancy(np.asarray(grids), colorful=True),
                            pack_tiles(np.asarray(tile_tables, dtype=np.uint8)))

    def simulate(self, n=None):
        for _ in range(self.num_sims if n is None else n):
            status, occ, tiles, valid = self._tree.select()
            priors = np.zeros((self.B, self.n_actions))
            values = np.zeros(self.B)
            need = np.flatnonzero(status == 0)
            if need.size:
                grids = unpack_occupancy(occ[need], self.width).astype(np.float32)
                if self.with_tiles:
                    pi, v = self.predict_batch(grids, tiles[need].astype(np.float32))
                else:
                    pi, v = self.predict_batch(grids)
                priors[need] = _mask_and_normalise_rows(pi, valid[need].astype(np.float64))
                values[need] = np.asarray(v, dtype=np.float64).reshape(-1)
                self.leaf_evals += int(need.size)
                self.leaf_batches += 1
            self._tree.backup(priors, values, want_values=False)

    # ---- device-resident leaf path ------------------------------------------------------
    def simulate_on_device(self, predict_dev, n=None):
        """Same lock-step simulations with no host copy in the loop: the leaf boards stay in
        device buffers, ``predict_dev(grids [B, H, W] float32 cuda tensor, tiles [B, 2n, 2]
        float32 cuda tensor) -> (policy [B, 2n] float32, value [B] float32)`` runs on the same
        CUDA stream, and masking / normalising the policy happens in the backup kernel."""
        import torch
        from .parallel import _DevicePointer
        B, H, W, E = self.B, self.height, self.width, self.n_actions
        wpr = (W + 31) // 32
        shifts = None
        # The select / backup kernels and the torch ops between them (bit unpacking, the net) must
        # run on ONE stream: the engine is bound to torch's current stream for the duration, so the
        # loop is also correct under torch.cuda.stream(...) and the temporaries the net returns are
        # not recycled by the caching allocator before the backup kernel has read them.
        with self._tree.engine.on_stream(torch.cuda.current_stream(
                torch.device("cuda", self._tree.engine.device)).cuda_stream):
            self._simulate_on_device(predict_dev, n, torch, _DevicePointer, B, H, W, E, wpr, shifts)

    def _simulate_on_device(self, predict_dev, n, torch, _DevicePointer, B, H, W, E, wpr, shifts):
        for _ in range(self.num_sims if n is None else n):
            occ_p, tiles_p, _valid_p, _status_p = self._tree.select_dev()
            if shifts is None:
                dev = torch.device("cuda", self._tree.engine.device)
                occ_t = torch.as_tensor(_DevicePointer(occ_p, B * H * wpr * 4), device=dev)
                occ_t = occ_t.view(torch.int32).view(B, H, wpr, 1)
                tiles_t = torch.as_tensor(_DevicePointer(tiles_p, B * E * 2), device=dev).view(B, E, 2)
                shifts = torch.arange(32, device=dev, dtype=torch.int32).view(1, 1, 1, 32)
            grids = ((occ_t >> shifts) & 1).reshape(B, H, wpr * 32)[:, :, :W].to(torch.float32)
            policy, value = predict_dev(grids, tiles_t.to(torch.float32))
            policy = policy.to(torch.float32).contiguous()
            value = value.to(torch.float32).contiguous()
            self._tree.backup_dev(policy.data_ptr(), value.data_ptr())
            self.leaf_batches += 1
            self.leaf_evals += B
        self._tree.check()

    def action_probs(self, temp=1):
        counts = self._tree.root_counts()[0].astype(np.float64)
        if temp == 0:
            probs = np.zeros_like(counts)
            probs[np.arange(self.B), counts.argmax(axis=1)] = 1
            return probs
        c = counts ** (1. / temp)
        total = c.sum(axis=1, keepdims=True)
        return np.divide(c, total, out=np.zeros_like(c), where=total > 0)   # finished games: zeros

    def root_counts_dev(self):
        return self._tree.root_counts_dev()

    def action_probs_root_parallel(self, temp=1, group=None):
        """Root-parallel form of ``action_probs``: the root Nsa and Wsa of every rank's trees are
        all-reduced (SUM, NCCL over NVLink, on the device buffers) and the policy is formed from
        the summed counts.  Every rank must call it, on the engine's stream; without an
        initialised process group it equals ``action_probs``.  -> (probs [B, E], q [B, E])."""
        import torch
        from .parallel import _DevicePointer, allreduce_root_stats, probs_from_counts
        B, E = self.B, self.n_actions
        dev = torch.device("cuda", self._tree.engine.device)
        with self._tree.engine.on_stream(torch.cuda.current_stream(dev).cuda_stream):
            c_ptr, w_ptr = self._tree.root_stats_dev()
            counts = torch.as_tensor(_DevicePointer(c_ptr, B * E * 4), device=dev).view(torch.int32).view(B, E)
            wsa = torch.as_tensor(_DevicePointer(w_ptr, B * E * 8), device=dev).view(torch.float64).view(B, E)
            counts, wsa, q = allreduce_root_stats(counts, wsa, group)
            return probs_from_counts(counts.cpu().numpy(), temp), q.cpu().numpy()

    def reset(self):
        self._tree.reset()
