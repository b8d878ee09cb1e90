"""GPU: split-phase (external evaluator) search for Connect Four.

1. machinery: with an evaluator that is an exact (dyadic) function of the packed planes the
   select/backprop kernels must reproduce the oracle bit for bit;
2. neural evaluator: bf16 forward on the GPU vs the fp32 CPU forward driving the oracle --
   per-leaf values and root values within 1e-3 (BASELINE.json north_star tolerance).
"""
import numpy as np
import pytest
import torch

from oracle import zc_oracle as zo
from test_gpu_c4_search import random_roots, roots_array
from zeroclone_b200 import _ffi
from zeroclone_b200.evaluator import NetEvaluator
from zeroclone_b200.models.connect4_value.network import ValueNetwork
from zeroclone_b200.search import TreeSearch

pytestmark = pytest.mark.gpu
TOL = 1e-3   # north_star: root values within 1e-3, bf16 forward vs fp32 reference
W = torch.tensor([1, 2, 3, 4, 3, 2, 1], dtype=torch.float32)


class DyadicEvaluator:
    """v = sum_cells w[col] * (own - opp) / 64, computed from the planes on the device."""

    def __init__(self, dtype):
        self.dtype = dtype

    def __call__(self, planes, out):
        w = W.to(planes.device).view(1, 1, 7)
        out.copy_(((planes[:, 0].float() - planes[:, 1].float()) * w).sum(dim=(1, 2)) / 64.0)
        return out


def oracle_states_to_planes(states_u8):
    """zo_c4_state bytes (42 chars + pad + turn) -> float32 planes [n,2,6,7] (c4_backend.py:52-61)."""
    n = states_u8.shape[0]
    cells = states_u8[:, :42].reshape(n, 6, 7)
    turn = states_u8[:, 44].astype(np.int64)
    cur = np.where(turn == 0, ord('X'), ord('O'))[:, None, None]
    opp = np.where(turn == 0, ord('O'), ord('X'))[:, None, None]
    return np.stack([(cells == cur), (cells == opp)], axis=1).astype(np.float32)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32, torch.float16])
@pytest.mark.parametrize("sims,batch,policy", [(800, 32, "first"), (203, 7, "last")])
def test_split_phase_machinery_bit_exact(dtype, sims, batch, policy):
    n = 48
    packed, states = random_roots(n, seed=11)
    ts = TreeSearch(_ffi.GAME_C4, n, sims)
    ts.set_roots(roots_array(packed))
    pol = {"first": _ffi.POLICY_FIRST, "last": _ffi.POLICY_LAST}[policy]
    ts.run_network(DyadicEvaluator(dtype), sims, 1.4, batch, pol)
    out, hashes = ts.results(), ts.tree_hash()

    def ext(states_u8):
        p = oracle_states_to_planes(states_u8)
        return ((p[:, 0] - p[:, 1]) * W.numpy().reshape(1, 1, 7)).sum(axis=(1, 2)) / 64.0

    for i in range(n):
        o = zo.search(zo.GAME_C4, states[i], sims, 1.4, batch, zo.EVAL_EXTERNAL,
                      {"first": zo.POLICY_FIRST, "last": zo.POLICY_LAST}[policy], external=ext)
        k = o.n_moves
        assert out["visits"][i][:k].tolist() == o.Na, i
        assert out["value_sums"][i][:k].tolist() == o.Wa, i
        assert int(out["result"][i]["best"]) == o.best
        assert int(hashes[i]) == o.tree_hash, i


def test_network_leaf_values_and_root_values_within_tolerance():
    torch.manual_seed(0)
    model = ValueNetwork().eval()
    ev = NetEvaluator(model, "cuda")
    n, sims = 24, 256
    packed, states = random_roots(n, seed=3)

    # per-leaf: bf16 GPU vs fp32 CPU on the same positions
    seen = []

    def ext(states_u8):
        p = oracle_states_to_planes(states_u8)
        seen.append(p)
        with torch.no_grad():
            return model(torch.from_numpy(p)).view(-1).double().numpy()

    oracle_root = []
    for i in range(n):
        o = zo.search(zo.GAME_C4, states[i], sims, 1.4, 32, zo.EVAL_EXTERNAL, zo.POLICY_FIRST, external=ext)
        oracle_root.append(sum(o.Wa) / max(1, sum(o.Na)))
    planes = torch.from_numpy(np.concatenate(seen)[:4096])
    with torch.no_grad():
        ref = model(planes).view(-1)
    got = ev(planes.to("cuda", ev.dtype)).cpu()
    leaf_err = (got - ref).abs().max().item()
    assert leaf_err < TOL, f"per-leaf |{ev.dtype} - fp32| max {leaf_err}"

    ts = TreeSearch(_ffi.GAME_C4, n, sims)
    ts.set_roots(roots_array(packed))
    ts.run_network(ev, sims, 1.4, 32, _ffi.POLICY_FIRST)
    out = ts.results()
    root = out["value_sums"].sum(axis=1) / np.maximum(1, out["visits"].sum(axis=1))
    err = np.abs(root - np.array(oracle_root)).max()
    assert err < TOL, f"root value |gpu - oracle| max {err}"
