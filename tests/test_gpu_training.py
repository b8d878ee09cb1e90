"""GPU: the training step (zeroclone_b200/training.py) -- the CUDA-graph step must train exactly like the eager one."""
import numpy as np
import pytest
import torch

from zeroclone_b200.training import train_epochs

pytestmark = pytest.mark.gpu


def _data(n):
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(n, 2, 6, 7, generator=g) < 0.3).float().numpy()
    y = (torch.rand(n, generator=g) * 2 - 1).numpy().astype(np.float32)
    return x, y


def test_graph_step_matches_eager_step():
    """ONE training step, eager vs CUDA graphs, from the same weights on the same batch.  Adam's first update is
    lr * sign(gradient): the two runs must move (almost) every parameter the same way -- cuDNN's backward accumulates with
    atomics, so a gradient within rounding of zero may flip -- and BatchNorm's statistics, a pure forward quantity, must agree
    closely.  The warm-up steps the capture recipe needs must leave no trace."""
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    x, y = _data(64)
    dev = torch.device("cuda")
    torch.manual_seed(0)
    init = ValueNetwork(blocks=2)
    p_init = torch.cat([p.detach().float().view(-1) for p in init.parameters()])
    out = {}
    for graphs in (False, True):
        torch.manual_seed(0)
        model = ValueNetwork(blocks=2)
        st = train_epochs(model, x, y, epochs=1, lr=1e-3, batch_size=64, device=dev, verbose=False, use_graphs=graphs)
        assert st["cuda_graphs"] == graphs and st["steps"] == 1
        out[graphs] = (st["loss"], torch.cat([p.detach().float().cpu().view(-1) for p in model.parameters()]) - p_init,
                       torch.cat([b.detach().float().cpu().view(-1) for b in model.buffers() if b.dtype.is_floating_point]))
        assert all(p.grad is None for p in model.parameters())
    (l0, d0, b0), (l1, d1, b1) = out[False], out[True]
    assert abs(l0 - l1) < 1e-5 * max(1.0, abs(l0)), (l0, l1)                  # the loss of the first batch: forward only
    for d in (d0, d1):                                                        # exactly one Adam step happened: nothing moved by more than lr
        assert d.abs().max().item() < 1.02e-3 and (d.abs() > 0.9e-3).float().mean().item() > 0.5
    moved = (d0.abs() > 0.5e-3) & (d1.abs() > 0.5e-3)
    assert moved.float().mean().item() > 0.5 and (torch.sign(d0[moved]) == torch.sign(d1[moved])).float().mean().item() > 0.99
    assert ((b0 - b1).abs() / (b0.abs() + 1.0)).max().item() < 1e-4


def test_training_reduces_the_loss_and_refreshes_the_tower():
    from zeroclone_b200.value_functions import Value
    torch.manual_seed(0)
    value = Value("network_latest", model_type="connect4_value")
    x, y = _data(512)
    ev = value.evaluator()
    planes = torch.from_numpy(x[:64]).to("cuda", ev.dtype)
    before = ev(planes).clone()
    first = train_epochs(value.model, x, y, epochs=1, lr=1e-3, batch_size=128, device=torch.device("cuda"), verbose=False)
    last = train_epochs(value.model, x, y, epochs=3, lr=1e-3, batch_size=128, device=torch.device("cuda"), verbose=False)
    assert last["loss"] < first["loss"]
    value.model.eval()
    value.refresh()
    after = ev(planes)
    assert not torch.equal(before, after)
    with torch.no_grad():
        ref = value.model.float()(torch.from_numpy(x[:64]).cuda()).view(-1)
    # the refreshed tower IS the trained network (bf16 operands on Connect Four: ~0.5 % of the value per leaf, and a trained
    # network's values are larger than a random one's; fp16 -- `value: {dtype: fp16}` -- is 8x tighter)
    assert (after - ref).abs().max().item() < 1e-2
    ev16 = type(ev)(value.model.cpu(), "cuda", torch.float16)
    assert (ev16(torch.from_numpy(x[:64]).to("cuda", torch.float16)) - ref).abs().max().item() < 1.5e-3
