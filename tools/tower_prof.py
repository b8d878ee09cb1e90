"""Development aid: run only the fused tower forward a few times (ncu target)."""
import sys
import torch
from zeroclone_b200.evaluator import FusedTowerEvaluator

game = sys.argv[1] if len(sys.argv) > 1 else "c4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
if game == "c4":
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    shape = (2, 6, 7)
else:
    from zeroclone_b200.models.chess_value.network import ValueNetwork
    shape = (17, 8, 8)
torch.manual_seed(0)
model = ValueNetwork().eval()
ev = FusedTowerEvaluator(model, "cuda", {"f16": torch.float16, "bf16": torch.bfloat16}.get(sys.argv[4]) if len(sys.argv) > 4 else None)
xd = (torch.rand(B, *shape) < 0.3).to("cuda", ev.dtype).contiguous()
out = torch.empty(B, dtype=torch.float32, device="cuda")
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
for _ in range(3 if reps == 1 else 20):
    ev(xd, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    ev(xd, out=out)
b.record()
torch.cuda.synchronize()
print(f"{a.elapsed_time(b) / reps:.3f} ms", float(out.abs().mean()))
