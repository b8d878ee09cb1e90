"""The value tower shared by every game: stem conv3x3 -> N residual blocks -> GAP -> Linear -> tanh.

Sub-module names (`stem.0/1`, `res.<i>.seq.0/1/3/4`, `head.2`) and their creation order follow the reference's
models/chess_value/network.py:9-45, so state_dicts and the whole-module pickles written by scripts/train.py:143
interchange, and the same seed builds the same weights (tests/test_network_golden.py relies on it).  Only the
input plane count differs per game.  This module is the fp32 definition: checkpoints, training, the reference
forward of the parity tests.  Inference during search runs in csrc/tower.cuh from its BatchNorm-folded weights.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

import numpy as np
import torch
from torch import nn
from torch.utils.data import Dataset

KERNEL = dict(kernel_size=3, padding=1, bias=False)


def _conv_bn(cin: int, cout: int) -> list:
    return [nn.Conv2d(cin, cout, **KERNEL), nn.BatchNorm2d(cout)]


class ResidualBlock(nn.Module):
    """x -> relu(x + bn(conv(relu(bn(conv(x))))))"""

    def __init__(self, c: int):
        super().__init__()
        first, second = _conv_bn(c, c), _conv_bn(c, c)
        self.seq = nn.Sequential(*first, nn.ReLU(inplace=True), *second)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu(self.seq(x) + x)


class ValueTower(nn.Module):
    in_planes = 17      # chess; subclasses override (Connect Four: 2)

    def __init__(self, channels: int = 128, blocks: int = 8, in_planes: Optional[int] = None):
        super().__init__()
        self.in_planes = self.in_planes if in_planes is None else in_planes
        self.stem = nn.Sequential(*_conv_bn(self.in_planes, channels), nn.ReLU(inplace=True))
        self.res = nn.Sequential(*[ResidualBlock(channels) for _ in range(blocks)])
        self.head = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(channels, 1), nn.Tanh())

    def forward(self, planes):
        return self.head(self.res(self.stem(planes)))


class ValueNetDataset(Dataset):
    """positions float32[N,C,H,W] with their targets float32[N] (the arrays Engine.get_dataset returns)"""

    def __init__(self, states: np.ndarray, values: np.ndarray):
        self.states = torch.as_tensor(np.asarray(states), dtype=torch.float32)
        self.values = torch.as_tensor(np.asarray(values), dtype=torch.float32)
        assert len(self.states) == len(self.values)

    def __len__(self) -> int:
        return int(self.values.shape[0])

    def __getitem__(self, i):
        return self.states[i], self.values[i]


def safe_globals(*extra, reference_module: Optional[str] = None) -> None:
    """Register what torch.load(weights_only=True) must be allowed to unpickle for a whole-module checkpoint
    (the reference does the same so that the safe loader works, network.py:59-72).  `reference_module`: the path
    the REFERENCE pickles these classes under (e.g. "models.chess_value.network"), so checkpoints written by the
    reference load here too; the names resolve to this package's classes through the repo-root import aliases."""
    layers = (nn.Sequential, nn.Conv2d, nn.BatchNorm2d, nn.ReLU, nn.AdaptiveAvgPool2d, nn.Flatten, nn.Linear, nn.Tanh)
    allowed = [ValueTower, ResidualBlock, *layers, *extra]
    if reference_module:
        allowed += [(ResidualBlock, f"{reference_module}.ResidualBlock")]
        allowed += [(cls, f"{reference_module}.{cls.__name__}") for cls in extra]
    torch.serialization.add_safe_globals(allowed)


def _one_epoch(model, batches: Iterable, optimiser, device, grad_sync: Optional[Callable]) -> float:
    """sum over the epoch of (batch loss x batch size)"""
    weighted = 0.0
    for planes, target in batches:
        planes, target = planes.to(device), target.to(device).unsqueeze(1)
        optimiser.zero_grad()
        loss = nn.functional.mse_loss(model(planes), target)
        loss.backward()
        if grad_sync is not None:      # multi-GPU: average gradients over ranks (one NCCL all-reduce)
            grad_sync(model)
        optimiser.step()
        weighted += loss.item() * planes.shape[0]
    return weighted


def train(model, dataloader, epochs: int = 10, lr: float = 1e-3, device=None, grad_sync=None):
    """Adam on the mean squared error, `epochs` passes over `dataloader`; returns the mean over epochs of the
    per-sample loss (what the reference's train() returns, network.py:75-101).  `grad_sync(model)`, if given,
    runs between backward() and step() -- the place where the reference would need its gradient all-reduce."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    model.to(device)
    optimiser = torch.optim.Adam(model.parameters(), lr=lr)
    history = []
    for epoch in range(epochs):
        model.train()
        history.append(_one_epoch(model, dataloader, optimiser, device, grad_sync) / len(dataloader.dataset))
        print(f"Epoch {epoch + 1}/{epochs} — Loss: {history[-1]:.4f}")
    return sum(history) / epochs
