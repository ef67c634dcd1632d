"""Neural-net leaf evaluator (reference interface: alpha/NeuralNet.py:1-50).

``NeuralNet`` is the reference's abstract interface (train / predict / save_checkpoint /
load_checkpoint).  ``TorchBinPackNet`` is a PyTorch implementation of it for the bin
packing game — a small two-tower net in the spirit of the reference's Keras model
(alpha/binpack/keras/ScalarKerasBinpackNNet.py:105-273: convolutions over the board,
dense layers over the tile list, policy over the 2n actions and a scalar value).  It is
the one place PyTorch does compute in this repository; the tree search itself is in
csrc/puct.cu.  ``predict_batch`` evaluates the leaves of many trees in one forward pass.
"""
import os

import numpy as np


class NeuralNet(object):
    def __init__(self, game):
        pass

    def train(self, examples):
        raise NotImplementedError

    def predict(self, board):
        raise NotImplementedError

    def save_checkpoint(self, folder, filename):
        raise NotImplementedError

    def load_checkpoint(self, folder, filename):
        raise NotImplementedError


def _build_model(height, width, n_actions, channels=32, hidden=128):
    import torch.nn as nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.board = nn.Sequential(
                nn.Conv2d(1, channels, 3, padding=1), nn.ReLU(),
                nn.Conv2d(channels, channels, 3, padding=1), nn.ReLU(),
                nn.Conv2d(channels, channels, 3, padding=1), nn.ReLU(),
                nn.AdaptiveMaxPool2d(1), nn.Flatten(), nn.Linear(channels, hidden), nn.ReLU())
            self.tiles = nn.Sequential(
                nn.Flatten(), nn.Linear(2 * n_actions, hidden), nn.ReLU(),
                nn.Linear(hidden, hidden), nn.ReLU())
            self.trunk = nn.Sequential(nn.Linear(2 * hidden, hidden), nn.ReLU())
            self.pi = nn.Linear(hidden, n_actions)
            self.v = nn.Linear(hidden, 1)

        def forward(self, grid, tiles):
            x = self.trunk(__import__("torch").cat([self.board(grid), self.tiles(tiles)], dim=1))
            return self.pi(x), __import__("torch").sigmoid(self.v(x)).squeeze(1)

    return Net()


class TorchBinPackNet(NeuralNet):
    """predict(board) -> (pi [2n], v) with board = (grid, tiles) or just the grid
    (binpack/MCTS.py:277 passes the grid only; then the tile tower sees zeros)."""

    def __init__(self, game, device=None, lr=1e-3, epochs=5, batch_size=64, seed=0,
                 inference_dtype="bfloat16"):
        import torch
        self.torch = torch
        self.height, self.width = game.getBoardSize()[0], game.getBoardSize()[1]
        self.n_actions = game.getActionSize()
        self.scale = float(max(self.height, self.width))
        self.device = torch.device(device if device is not None else
                                   ("cuda" if torch.cuda.is_available() else "cpu"))
        torch.manual_seed(seed)
        self.model = _build_model(self.height, self.width, self.n_actions).to(self.device)
        if self.device.type == "cuda":
            self.model = self.model.to(memory_format=torch.channels_last)
        self.opt = torch.optim.Adam(self.model.parameters(), lr=lr)
        self.epochs, self.batch_size = epochs, batch_size
        # leaf evaluation in self-play runs under autocast (tensor cores via cuDNN/cuBLAS); the
        # policy softmax and the value are returned in float32
        self.inference_dtype = getattr(torch, inference_dtype) if inference_dtype else None

    # ---- inference -------------------------------------------------------------------
    def _tensors(self, grids, tiles):
        torch = self.torch
        g = torch.as_tensor(np.asarray(grids, dtype=np.float32) != 0, dtype=torch.float32,
                            device=self.device).reshape(-1, 1, self.height, self.width)
        if tiles is None:
            t = torch.zeros((g.shape[0], self.n_actions, 2), device=self.device)
        else:
            t = torch.as_tensor(np.asarray(tiles, dtype=np.float32), device=self.device)
            t = t.reshape(-1, self.n_actions, 2) / self.scale
        return g, t

    def predict_batch(self, grids, tiles=None):
        """grids [B, H, W], tiles [B, 2n, 2] -> (pi [B, 2n] float64, v [B] float64)"""
        torch = self.torch
        self.model.eval()
        with torch.no_grad():
            g, t = self._tensors(grids, tiles)
            logits, v = self.model(g, t)
            pi = torch.softmax(logits.double(), dim=1)
        return pi.cpu().numpy(), v.double().cpu().numpy()

    def predict_batch_dev(self, grids, tiles):
        """cuda tensors in (grids [B, H, W] 0/1 float32, tiles [B, 2n, 2] float32), cuda tensors
        out (softmax policy [B, 2n] float32, value [B] float32); nothing leaves the device."""
        torch = self.torch
        self.model.eval()
        g = grids.reshape(-1, 1, self.height, self.width).contiguous(memory_format=torch.channels_last)
        t = tiles.reshape(-1, self.n_actions, 2) / self.scale
        with torch.no_grad():
            if self.inference_dtype is not None and g.is_cuda:
                with torch.autocast("cuda", dtype=self.inference_dtype):
                    logits, v = self.model(g, t)
            else:
                logits, v = self.model(g, t)
            return torch.softmax(logits.float(), dim=1), v.float()

    def predict(self, board):
        if isinstance(board, tuple):
            pi, v = self.predict_batch(np.asarray(board[0])[None], np.asarray(board[1])[None])
        else:
            pi, v = self.predict_batch(np.asarray(board)[None], None)
        return pi[0], float(v[0])

    # ---- training --------------------------------------------------------------------
    def train(self, examples):
        """examples: (grid, tiles, pi, v) tuples as produced by Coach.executeEpisode."""
        torch = self.torch
        if not examples:
            return 0.0
        grids = np.stack([np.asarray(e[0], dtype=np.float32) for e in examples])
        tiles = np.stack([np.asarray(e[1], dtype=np.float32).reshape(self.n_actions, 2) for e in examples])
        pis = torch.as_tensor(np.stack([np.asarray(e[2], dtype=np.float32) for e in examples]),
                              device=self.device)
        vs = torch.as_tensor(np.asarray([float(e[3]) for e in examples], dtype=np.float32),
                             device=self.device)
        g, t = self._tensors(grids, tiles)
        self.model.train()
        n = g.shape[0]
        last = 0.0
        gen = torch.Generator(device="cpu").manual_seed(0)
        for _ in range(self.epochs):
            perm = torch.randperm(n, generator=gen).to(self.device)
            for i in range(0, n, self.batch_size):
                idx = perm[i:i + self.batch_size]
                logits, v = self.model(g[idx], t[idx])
                loss_pi = -(pis[idx] * torch.log_softmax(logits, dim=1)).sum(dim=1).mean()
                loss_v = ((v - vs[idx]) ** 2).mean()
                loss = loss_pi + loss_v
                self.opt.zero_grad()
                loss.backward()
                self.opt.step()
                last = float(loss.item())
        return last

    def save_checkpoint(self, folder='checkpoint', filename='checkpoint.pth.tar'):
        os.makedirs(folder, exist_ok=True)
        self.torch.save({"state_dict": self.model.state_dict()}, os.path.join(folder, filename))

    def load_checkpoint(self, folder='checkpoint', filename='checkpoint.pth.tar'):
        ckpt = self.torch.load(os.path.join(folder, filename), map_location=self.device)
        self.model.load_state_dict(ckpt["state_dict"])
