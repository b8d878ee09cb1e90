"""The value tower shared by every game: stem conv3x3 -> N residual blocks -> GAP -> Linear -> tanh.

Architecture and parameter names follow the reference's models/chess_value/network.py:9-45 so
that state_dicts (and whole-module pickles saved by scripts/train.py:143) interchange:
`stem.0/1`, `res.<i>.seq.0/1/3/4`, `head.2`.  Only the input plane count differs per game.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.utils.data import Dataset


class ResidualBlock(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.seq = nn.Sequential(
            nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c), nn.ReLU(inplace=True),
            nn.Conv2d(c, c, 3, padding=1, bias=False), nn.BatchNorm2d(c))
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(x + self.seq(x))


class ValueTower(nn.Module):
    in_planes = 17

    def __init__(self, channels: int = 128, blocks: int = 8, in_planes: int | None = None):
        super().__init__()
        if in_planes is not None:
            self.in_planes = in_planes
        self.stem = nn.Sequential(nn.Conv2d(self.in_planes, channels, 3, padding=1, bias=False),
                                  nn.BatchNorm2d(channels), nn.ReLU(inplace=True))
        self.res = nn.Sequential(*[ResidualBlock(channels) for _ in range(blocks)])
        self.head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(channels, 1), nn.Tanh())

    def forward(self, x):
        return self.head(self.res(self.stem(x)))


class ValueNetDataset(Dataset):
    """(states float32[N,C,H,W], values float32[N]) -> tensors; reference network.py:47-56."""

    def __init__(self, states: np.ndarray, values: np.ndarray):
        self.states = torch.from_numpy(np.asarray(states)).float()
        self.values = torch.from_numpy(np.asarray(values)).float()

    def __len__(self):
        return self.states.shape[0]

    def __getitem__(self, i):
        return self.states[i], self.values[i]


def safe_globals(*extra):
    """Classes a whole-module checkpoint needs under torch.load(weights_only=True); network.py:58-71."""
    torch.serialization.add_safe_globals([
        ResidualBlock, ValueTower, nn.Conv2d, nn.BatchNorm2d, nn.ReLU, nn.AdaptiveAvgPool2d, nn.Linear, nn.Tanh,
        nn.Sequential, nn.Flatten, *extra])


def train(model, dataloader, epochs: int = 10, lr: float = 1e-3, device=None, grad_sync=None):
    """Adam + MSE loop with the reference's signature and return value (network.py:75-101):
    the mean over epochs of the per-sample average loss.  `grad_sync`, if given, is called between
    backward() and step() -- the insertion point of the NCCL gradient all-reduce."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    model.to(device)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    mse = nn.MSELoss()
    total = 0.0
    for epoch in range(1, epochs + 1):
        model.train()
        seen = 0.0
        for states, targets in dataloader:
            states = states.to(device)
            targets = targets.to(device).unsqueeze(1)
            opt.zero_grad()
            loss = mse(model(states), targets)
            loss.backward()
            if grad_sync is not None:
                grad_sync(model)
            opt.step()
            seen += loss.item() * states.size(0)
        epoch_loss = seen / len(dataloader.dataset)
        total += epoch_loss
        print(f"Epoch {epoch}/{epochs} — Loss: {epoch_loss:.4f}")
    return total / epochs
