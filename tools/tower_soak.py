"""Development aid: soak the fused tower kernel -- thousands of launches with random batch sizes and offsets,
every result compared bit-for-bit with one reference evaluation of the whole set (a leaf's value does not
depend on its batch).  Catches synchronisation bugs that only show under timing variation."""
import sys
import time

import torch

from zeroclone_b200.evaluator import NetEvaluator

game = sys.argv[1] if len(sys.argv) > 1 else "c4"
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 20.0
if game == "c4":
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    shape = (2, 6, 7)
else:
    from zeroclone_b200.models.chess_value.network import ValueNetwork
    shape = (17, 8, 8)
torch.manual_seed(0)
ev = NetEvaluator(ValueNetwork().eval(), "cuda")
N = 40000
x = (torch.rand(N, *shape) < 0.3).to("cuda", ev.dtype).contiguous()
ref = ev(x).clone()
g = torch.Generator().manual_seed(1)
t0, launches, leaves = time.time(), 0, 0
busy = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
while time.time() - t0 < seconds:
    n = int(torch.randint(1, 6000, (1,), generator=g))
    if launches % 7 == 0:
        n = int(torch.randint(1, 40, (1,), generator=g))
    off = int(torch.randint(0, N - n, (1,), generator=g))
    if launches % 5 == 0:
        busy.fill_(launches & 255)            # perturb timing: another kernel right before
    got = ev(x[off:off + n].contiguous())
    if not torch.equal(got, ref[off:off + n]):
        bad = (got != ref[off:off + n]).nonzero().view(-1)
        print(f"MISMATCH launch {launches} n={n} off={off} first bad {bad[:5].tolist()} of {bad.numel()}")
        sys.exit(1)
    launches += 1
    leaves += n
print(f"soak ok: {launches} launches, {leaves} leaves, all bit-identical to the reference evaluation")
