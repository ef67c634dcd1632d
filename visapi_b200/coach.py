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
        """numIters x (numEps episodes of self-play, then nnet.train) (Coach.py:65-134).  The
        reference accepts every new model (`if False:` at :122), so no arena is played."""
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
            losses.append(self.nnet.train(examples))
            if ckpt:
                self.nnet.save_checkpoint(folder=ckpt, filename=self.getCheckpointFile(i))
                self.nnet.save_checkpoint(folder=ckpt, filename='best.pth.tar')
        return losses

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
        """-> (examples, rewards [B], searches) ; examples as Coach.executeEpisode makes them."""
        boards = []
        for _ in range(self.B):
            self.game.getInitBoard()                       # a fresh random instance per game
            boards.append((np.zeros([self.height, self.width]),
                           np.array([[int(t[0]), int(t[1])] for t in self.game._base_board.tiles])))
        self.search.reset()
        done = np.zeros(self.B, bool)
        reward = np.zeros(self.B)
        pending = [[] for _ in range(self.B)]
        step = 0
        searches = 0
        while not done.all():
            step += 1
            self.search.set_roots(np.stack([b[0] for b in boards]), np.stack([b[1] for b in boards]))
            if self.device_leaves:
                self.search.simulate_on_device(self.nnet.predict_batch_dev)
            else:
                self.search.simulate()
            searches += self.num_sims * int((~done).sum())
            temp = int(step < int(_arg(self.args, 'tempThreshold', 15)))
            probs = self.search.action_probs(temp=temp)
            actions = np.zeros(self.B, np.int32)
            for i in range(self.B):
                if done[i]:
                    continue
                pending[i].append((boards[i][0].copy(), boards[i][1].copy(), probs[i].copy()))
                actions[i] = self.rng.choice(len(probs[i]), p=probs[i])
            live = [i for i in range(self.B) if not done[i]]
            nxt, verdict = self.game.getNextState_batch([boards[i] for i in live], 1, actions[live])
            for k, i in enumerate(live):
                boards[i] = (nxt[k][0], np.asarray(nxt[k][1]))
            ended = self.game.getGameEnded_batch([boards[i] for i in live])
            for k, i in enumerate(live):
                if ended[k] != 0:
                    done[i] = True
                    reward[i] = ended[k]
        self.leaf_batches = self.search.leaf_batches
        examples = [(g, t, p, reward[i]) for i in range(self.B) for (g, t, p) in pending[i]]
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
