#!/usr/bin/env python
"""Golden vectors for the value network, produced by the UNMODIFIED reference in the dev container:
models/chess_value/network.py (imported from /root/reference) evaluated in fp32 on positions reached by
random playouts of the reference's chess backend (oracle/_ref, built from the reference's C++ by
oracle/Makefile), encoded by the reference's own state_to_tensor.  Writes tests/golden/chess_network.json.gz.
The GPU box only reads the JSON."""
import gzip
import importlib.util
import json
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("ZC_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REPO, "oracle", "_ref"))
sys.path.insert(0, HERE)
import chess_backend as ref_chess  # noqa: E402  reference C++ backend
import network_fixture  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_network", os.path.join(REF, "models", "chess_value", "network.py"))
ref_network = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_network)

model = network_fixture.build(ref_network.ValueNetwork)
rng = random.Random(7)
planes = []
for game in range(12):
    s = ref_chess.create_init_state()
    for ply in range(60):
        moves = list(ref_chess.get_legal_moves(s))
        if not moves:
            break
        s = ref_chess.play_move(s, rng.choice(moves))
        if ply % 6 == 5:
            planes.append(np.asarray(ref_chess.state_to_tensor(s), dtype=np.float32))
x = np.stack(planes)
assert set(np.unique(x)) <= {0.0, 1.0}
with torch.no_grad():
    y = model(torch.from_numpy(x)).view(-1).double().tolist()
bits = [np.packbits(p.reshape(-1).astype(np.uint8)).tobytes().hex() for p in x]
out = {"seed": network_fixture.SEED, "shape": [17, 8, 8], "planes_bits_hex": bits, "values_fp32": y,
       "source": "reference models/chess_value/network.py + chess_backend.state_to_tensor, torch %s" % torch.__version__}
path = os.path.join(HERE, "chess_network.json.gz")
with gzip.open(path, "wt") as fh:
    json.dump(out, fh)
print("wrote", path, len(bits), "positions; values", min(y), max(y))
