"""Self-play (reference: alpha/Coach.py:11-162).

``Coach`` keeps the reference's episode loop — getActionProb with temperature 1 for the
first ``tempThreshold`` moves and 0 afterwards, sample the action from pi, store
(grid, tiles, pi) and label every stored position with the episode's final reward — on
top of the device PUCT search.  ``BatchedSelfPlay`` plays many episodes at once: one tree
per game on the device, all trees stepped in lock step, the leaves of all games evaluated
by the net as one batch per simulation.  With several ranks every rank plays its own games
(replicas, no collective on the search path) and the examples are gathered.
"""
import os
import pickle
from collections import deque
from random import shuffle

import numpy as np

from .game import BinPackGame
from .puct import MCTS, BatchedMCTS


def _arg(args, name, default=None):
    if isinstance(args, dict):
        return args.get(name, default)
    return getattr(args, name, default)


class Coach(object):
    def __init__(self, game, nnet, args, device=None):
        self.game, self.nnet, self.args, self.device = game, nnet, args, device
        self.arena_results = []
        self.mcts = MCTS(self.game, self.nnet, self.args, device=device)
        self.trainExamplesHistory = []
        self.skipFirstSelfPlay = False

    def executeEpisode(self):
        """One episode -> [(grid, tiles, pi, reward), ...] (Coach.py:25-63)."""
        trainExamples = []
        board, _ = self.game.getInitBoard()
        self.mcts = MCTS(self.game, self.nnet, self.args, device=self.device)
        self.mcts.predict_on_grid_only = False
        episodeStep = 0
        while True:
            episodeStep += 1
            canonicalBoard = self.game.getCanonicalForm(board, 1)
            temp = int(episodeStep < _arg(self.args, 'tempThreshold', 15))
            pi = self.mcts.getActionProb(canonicalBoard, temp=temp)
            for b, p in self.game.getSymmetries(canonicalBoard, pi):
                trainExamples.append([np.copy(b[0]), np.array([[int(t[0]), int(t[1])] for t in b[1]]), p])
            action = np.random.choice(len(pi), p=pi)
            board, _, _ = self.game.getNextState(board, 1, action)
            r = self.game.getGameEnded(board, 1)
            if r != 0:
                return [(x[0], x[1], x[2], r) for x in trainExamples]

    def learn(self):
        """numIters x (numEps episodes of self-play, then nnet.train) (Coach.py:65-134).  Like the
        reference, the previous and the new net are pitted in the Arena for ``arenaCompare`` games
        (:116-118) and the new model is accepted whatever the outcome (`if False:` at :122);
        ``self.arena_results`` keeps the (prev wins, new wins, draws) of every iteration."""
        self.arena_results = []
        losses = []
        for i in range(1, _arg(self.args, 'numIters', 1) + 1):
            if not self.skipFirstSelfPlay or i > 1:
                it = deque([], maxlen=_arg(self.args, 'maxlenOfQueue', 200000))
                for _ in range(_arg(self.args, 'numEps', 1)):
                    it += self.executeEpisode()
                self.trainExamplesHistory.append(it)
            if len(self.trainExamplesHistory) > _arg(self.args, 'numItersForTrainExamplesHistory', 20):
                self.trainExamplesHistory.pop(0)
            ckpt = _arg(self.args, 'checkpoint')
            if ckpt:
                self.saveTrainExamples(i - 1)
            examples = [e for part in self.trainExamplesHistory for e in part]
            shuffle(examples)
            previous = self._snapshot_net()
            losses.append(self.nnet.train(examples))
            n_arena = int(_arg(self.args, 'arenaCompare', 0) or 0)
            if n_arena >= 2 and previous is not None:
                from .arena import Arena, mcts_player
                pmcts = MCTS(self.game, previous, self.args, device=self.device)
                nmcts = MCTS(self.game, self.nnet, self.args, device=self.device)
                for m in (pmcts, nmcts):
                    m.predict_on_grid_only = False
                arena = Arena(mcts_player(pmcts), mcts_player(nmcts), self.game)
                self.arena_results.append(arena.playGames(n_arena))
            if ckpt:
                self.nnet.save_checkpoint(folder=ckpt, filename=self.getCheckpointFile(i))
                self.nnet.save_checkpoint(folder=ckpt, filename='best.pth.tar')
        return losses

    def _snapshot_net(self):
        """A frozen copy of the current net (the reference's pnet, Coach.py:19,109-111)."""
        import copy
        try:
            return copy.deepcopy(self.nnet)
        except Exception:
            return None

    def getCheckpointFile(self, iteration):
        return 'checkpoint_' + str(iteration) + '.pth.tar'

    def saveTrainExamples(self, iteration):
        folder = _arg(self.args, 'checkpoint')
        os.makedirs(folder, exist_ok=True)
        with open(os.path.join(folder, self.getCheckpointFile(iteration) + ".examples"), "wb+") as f:
            pickle.dump(self.trainExamplesHistory, f)

    def loadTrainExamples(self, path):
        with open(path, "rb") as f:
            self.trainExamplesHistory = pickle.load(f)
        self.skipFirstSelfPlay = True


class BatchedSelfPlay(object):
    """n_games episodes of BinPackGame(height, width, n_tiles) in lock step on one GPU."""

    def __init__(self, height, width, n_tiles, nnet, args, n_games, device=None, seed=0,
                 device_leaves=True):
        self.height, self.width, self.n_tiles = height, width, n_tiles
        self.nnet, self.args, self.B = nnet, args, int(n_games)
        self.game = BinPackGame(height, width, n_tiles, device=device)
        self.rng = np.random.default_rng(seed)
        self.num_sims = int(_arg(args, 'numMCTSSims', 50))
        self.search = BatchedMCTS(
            height, width, 2 * n_tiles, self._predict, self.num_sims, float(_arg(args, 'cpuct', 1.0)),
            self.B, device=device, with_tiles=True)
        # leaves evaluated without leaving the device when the net offers predict_batch_dev
        self.device_leaves = device_leaves and hasattr(nnet, "predict_batch_dev")
        self.leaf_batches = 0

    def _predict(self, grids, tiles):
        return self.nnet.predict_batch(grids, tiles)

    def play(self):
        """-> (examples, rewards [B], searches); examples as Coach.executeEpisode makes them:
        (grid, tiles, pi, reward of that game).  All per-move bookkeeping is array-wise: the
        B boards live in two numpy arrays, the move of every game is sampled in one shot, the
        transitions and terminal tests are one batched device call each."""
        from . import runtime
        from ._native import PLACED, pack_occupancy, unpack_occupancy
        eng = runtime.get_engine(self.game.device)
        B, H, W, E = self.B, self.height, self.width, 2 * self.n_tiles
        grids = np.zeros((B, H, W), np.uint8)
        tiles = np.zeros((B, E, 2), np.uint8)
        for i in range(B):
            self.game.getInitBoard()                       # a fresh random instance per game
            tiles[i] = np.asarray(self.game._base_board.tiles, dtype=np.uint8)
        self.search.reset()
        done = np.zeros(B, bool)
        reward = np.zeros(B)
        ex_game, ex_grid, ex_tiles, ex_pi = [], [], [], []
        step = 0
        searches = 0
        temp_moves = int(_arg(self.args, 'tempThreshold', 15))
        while not done.all():
            step += 1
            self.search.set_roots(grids, tiles)
            if self.device_leaves:
                self.search.simulate_on_device(self.nnet.predict_batch_dev)
            else:
                self.search.simulate()
            live = np.flatnonzero(~done)
            searches += self.num_sims * live.size
            probs = self.search.action_probs(temp=int(step < temp_moves))[live]
            ex_game.append(live)
            ex_grid.append(grids[live].astype(np.float32))
            ex_tiles.append(tiles[live].astype(np.int64))
            ex_pi.append(probs)
            # one draw per live game: first action whose cumulative probability exceeds u
            cdf = np.cumsum(probs, axis=1)
            u = self.rng.random(live.size) * cdf[:, -1]
            actions = np.minimum((cdf <= u[:, None]).sum(axis=1), E - 1).astype(np.int32)
            zero = probs[np.arange(live.size), actions] <= 0       # guard against rounding at the edge
            if zero.any():
                actions[zero] = probs[zero].argmax(axis=1)
            out_occ, out_tiles, verdict, _ = eng.transitions(pack_occupancy(grids[live]), W,
                                                             tiles[live], actions, compact=True)
            assert (verdict == PLACED).all()
            grids[live] = unpack_occupancy(out_occ, W)
            tiles[live] = out_tiles
            ended = eng.game_ended(pack_occupancy(grids[live]), W, tiles[live])
            fin = ended != 0
            done[live[fin]] = True
            reward[live[fin]] = ended[fin]
        self.leaf_batches = self.search.leaf_batches
        examples = []
        for games, g, t, p in zip(ex_game, ex_grid, ex_tiles, ex_pi):
            for k, i in enumerate(games):
                examples.append((g[k], t[k], p[k], reward[i]))
        return examples, reward, searches


def gather_examples(examples):
    """All ranks' examples on every rank (torch.distributed all_gather_object); no-op alone."""
    try:
        import torch.distributed as dist
    except ImportError:
        return examples
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return examples
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, examples)
    return [e for part in parts for e in part]
