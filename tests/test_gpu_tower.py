"""Fused value-tower kernel (csrc/tower.cuh, zc_tower_*) against a plain PyTorch fp32 forward of the
same network (models/chess_value/network.py:24-45 of the reference).  Tolerance 1e-3 on the tanh
output, the bound BASELINE.json's north_star states for the neural evaluator (bf16 vs fp32)."""
import numpy as np
import pytest
import torch

from zeroclone_b200.evaluator import NetEvaluator, TorchTowerEvaluator

pytestmark = pytest.mark.gpu
TOL = 1e-3


def make_model(game, perturb_bn, seed=0):
    if game == "c4":
        from zeroclone_b200.models.connect4_value.network import ValueNetwork
        shape = (2, 6, 7)
    else:
        from zeroclone_b200.models.chess_value.network import ValueNetwork
        shape = (17, 8, 8)
    torch.manual_seed(seed)
    model = ValueNetwork().eval()
    if perturb_bn:   # running statistics and affine terms away from their defaults so the fold is exercised
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.8, 1.2)
                m.bias.data.normal_(0, 0.1)
    return model, shape


def random_planes(n, shape, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(n, *shape, generator=g) < 0.3).float()
    if shape[0] == 2:
        x[:, 1] *= 1 - x[:, 0]      # a cell holds one disc
    return x


# operand formats of the fused kernel and the per-leaf bound each must hold against fp32: bf16 (8 significand bits)
# the north-star 1e-3; fp16 (11 bits, the default and the reference's own GPU dtype) five times tighter
DTYPES = [(torch.float16, 2e-4), (torch.bfloat16, TOL)]


@pytest.mark.parametrize("game", ["c4", "chess"])
@pytest.mark.parametrize("perturb_bn", [False, True])
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_tower_matches_fp32_reference(game, perturb_bn, dtype, tol):
    model, shape = make_model(game, perturb_bn)
    n = 3000
    x = random_planes(n, shape, 1)
    with torch.no_grad():
        ref = model(x).view(-1)
    ev = NetEvaluator(model, "cuda", dtype)
    got = ev(x.to("cuda", dtype)).cpu()
    err = (got - ref).abs().max().item()
    print(f"{game} perturb_bn={perturb_bn} {dtype}: per-leaf |fused - fp32| max {err:.2e}")
    assert err < tol, f"|fused {dtype} - fp32| max {err}"
    assert ev.launches == 1


@pytest.mark.parametrize("game", ["c4", "chess"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tower_ragged_batches_and_determinism(game, dtype):
    """every leaf is evaluated independently of its neighbours in the tile and of the batch size:
    the same position must give the same bits at any offset, for any n (incl. n not a multiple of the
    boards per tile, n smaller than one tile, n = 0)"""
    model, shape = make_model(game, True, seed=3)
    x = random_planes(1500, shape, 2).to("cuda", dtype)
    ev = NetEvaluator(model, "cuda", dtype)
    full = ev(x).cpu()
    again = ev(x).cpu()
    assert torch.equal(full, again), "not deterministic"
    assert ev(x[:0]).numel() == 0
    for n in (1, 2, 3, 4, 5, 7, 127, 128, 129, 887):
        part = ev(x[:n].contiguous()).cpu()
        assert torch.equal(part, full[:n]), n
    off = ev(x[301:1207].contiguous()).cpu()
    assert torch.equal(off, full[301:1207])
    # a leaf does not see its tile neighbours: same leaf surrounded by different boards
    y = x.clone()
    y[1::2] = 0
    alt = ev(y).cpu()
    assert torch.equal(alt[0::2], full[0::2])


@pytest.mark.parametrize("game", ["c4", "chess"])
@pytest.mark.parametrize("dtype,tol", DTYPES)
def test_tower_agrees_with_torch_16bit_path(game, dtype, tol):
    model, shape = make_model(game, False)
    x = random_planes(2048, shape, 5).to("cuda", dtype)
    a = NetEvaluator(model, "cuda", dtype)(x).cpu()
    b = TorchTowerEvaluator(model, "cuda", dtype)(x).cpu()
    assert (a - b).abs().max().item() < 2 * tol


def test_tower_full_batch_size():
    """131072 leaves = one search batch of BASELINE configs[1] (4096 trees x 32): spot-check against fp32"""
    model, shape = make_model("c4", False)
    n = 131072
    x = random_planes(n, shape, 7)
    ev = NetEvaluator(model, "cuda")
    got = ev(x.to("cuda", ev.dtype)).cpu()
    idx = torch.randint(0, n, (2048,), generator=torch.Generator().manual_seed(0))
    idx[-1] = n - 1
    with torch.no_grad():
        ref = model(x[idx]).view(-1)
    assert (got[idx] - ref).abs().max().item() < TOL
    assert torch.isfinite(got).all()
    # the CTA-pair / two-tiles-in-flight pipeline must not depend on timing: bit-identical reruns, and the
    # same bits as evaluating a slice on its own (different CTA, tile slot and pair rank for every leaf)
    xd = x.to("cuda", ev.dtype)
    for _ in range(3):
        assert torch.equal(ev(xd).cpu(), got)
    assert torch.equal(ev(xd[70001:70001 + 5000].contiguous()).cpu(), got[70001:70001 + 5000])


def test_tower_rejects_what_it_cannot_compute():
    model, shape = make_model("c4", False)
    ev = NetEvaluator(model, "cuda")
    with pytest.raises(ValueError):
        ev(torch.zeros(4, *shape, device="cuda", dtype=torch.float32))
    with pytest.raises(ValueError):
        ev(torch.zeros(4, *shape, dtype=ev.dtype))                                        # host tensor
    with pytest.raises(ValueError):
        ev(torch.zeros(4, *shape, device="cuda", dtype=torch.float16))                    # planes in the other 16-bit format
    with pytest.raises(ValueError):
        NetEvaluator(model, "cuda", torch.float32)
    assert ev.dtype == torch.bfloat16 and NetEvaluator(model, "cuda", torch.float16).dtype == torch.float16      # Connect Four default: bf16


def test_tower_soak_random_batches_bit_identical():
    """a few thousand launches with random sizes / offsets, other kernels in between: a leaf's value never
    depends on the batch it is evaluated in (tools/tower_soak.py runs the long version)"""
    import time
    model, shape = make_model("c4", True, seed=9)
    ev = NetEvaluator(model, "cuda")
    n_all = 20000
    x = random_planes(n_all, shape, 11).to("cuda", ev.dtype)
    ref = ev(x).clone()
    g = torch.Generator().manual_seed(1)
    busy = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
    t0, launches = time.time(), 0
    while time.time() - t0 < 3.0:
        n = int(torch.randint(1, 40 if launches % 3 == 0 else 3000, (1,), generator=g))
        off = int(torch.randint(0, n_all - n, (1,), generator=g))
        if launches % 5 == 0:
            busy.fill_(launches & 255)
        assert torch.equal(ev(x[off:off + n].contiguous()), ref[off:off + n]), (launches, n, off)
        launches += 1
    assert launches > 500


@pytest.mark.parametrize("blocks", [1, 3])
def test_tower_with_fewer_blocks(blocks):
    """the kernel takes the number of residual blocks from the model (<= 8, the reference's depth)"""
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    torch.manual_seed(4)
    model = ValueNetwork(blocks=blocks).eval()
    x = random_planes(700, (2, 6, 7), 3)
    with torch.no_grad():
        ref = model(x).view(-1)
    ev = NetEvaluator(model, "cuda")
    got = ev(x.to("cuda", ev.dtype)).cpu()
    assert (got - ref).abs().max().item() < TOL
    with pytest.raises(Exception):
        NetEvaluator(ValueNetwork(blocks=9).eval(), "cuda")
