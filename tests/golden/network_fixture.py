"""Shared between make_golden_network.py (runs the reference) and the tests (run ours): how the fixture's
network is built.  The reference's ValueNetwork (models/chess_value/network.py:24-45) and this repo's module
create the same parameters in the same order, so the same seed gives the same weights in both."""
import torch

SEED = 20241018


def build(cls):
    torch.manual_seed(SEED)
    model = cls().eval()
    g = torch.Generator().manual_seed(SEED + 1)
    for m in model.modules():       # running statistics and affine terms away from their defaults
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.weight.shape, generator=g) * 0.4 + 0.8)
            m.bias.data.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    return model
