"""The value network against vectors produced by the UNMODIFIED reference (tests/golden/make_golden_network.py:
the reference's models/chess_value/network.py in fp32 on positions encoded by the reference's state_to_tensor).
CPU: this repo's PyTorch module reproduces them exactly -> it is the same network.  GPU: the fused tower kernel
agrees within the 1e-3 BASELINE.json allows for the bf16 forward."""
import gzip
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REPO

sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
import network_fixture  # noqa: E402


def fixture():
    with gzip.open(os.path.join(REPO, "tests", "golden", "chess_network.json.gz"), "rt") as fh:
        d = json.load(fh)
    x = np.stack([np.unpackbits(np.frombuffer(bytes.fromhex(h), dtype=np.uint8))[:17 * 64].reshape(17, 8, 8) for h in d["planes_bits_hex"]])
    return torch.from_numpy(x.astype(np.float32)), torch.tensor(d["values_fp32"], dtype=torch.float64)


def test_module_reproduces_reference_network_outputs():
    from zeroclone_b200.models.chess_value.network import ValueNetwork
    x, want = fixture()
    model = network_fixture.build(ValueNetwork)
    with torch.no_grad():
        got = model(x).view(-1).double()
    assert (got - want).abs().max().item() < 1e-6
    # the planes came from the reference's state_to_tensor: plane 12 is all-ones or all-zeros (side to move)
    assert set(x[:, 12].reshape(len(x), -1).mean(1).tolist()) <= {0.0, 1.0}


@pytest.mark.gpu
def test_fused_tower_matches_reference_network_outputs():
    from zeroclone_b200.evaluator import NetEvaluator
    from zeroclone_b200.models.chess_value.network import ValueNetwork
    x, want = fixture()
    ev = NetEvaluator(network_fixture.build(ValueNetwork), "cuda")
    got = ev(x.to("cuda", ev.dtype)).cpu().double()
    assert (got - want).abs().max().item() < 1e-3
