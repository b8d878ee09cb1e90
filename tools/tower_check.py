"""Development aid: fused tower kernel vs the PyTorch paths (accuracy + speed) on the GPU."""
import sys
import torch
from zeroclone_b200.evaluator import FusedTowerEvaluator, TorchTowerEvaluator, tower_flops_per_leaf

game = sys.argv[1] if len(sys.argv) > 1 else "c4"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
if game == "c4":
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    shape, fl = (2, 6, 7), tower_flops_per_leaf(2, 6, 7)
else:
    from zeroclone_b200.models.chess_value.network import ValueNetwork
    shape, fl = (17, 8, 8), tower_flops_per_leaf(17, 8, 8)
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = ValueNetwork().eval()
# non-trivial BN statistics so folding is exercised
for m in model.modules():
    if isinstance(m, torch.nn.BatchNorm2d):
        m.running_mean.normal_(0, 0.1)
        m.running_var.uniform_(0.5, 1.5)
        m.weight.data.uniform_(0.8, 1.2)
        m.bias.data.normal_(0, 0.1)
x = (torch.rand(B, *shape) < 0.3).float()
nref = min(B, 4096)
with torch.no_grad():
    ref = model(x[:nref]).view(-1)
fused = FusedTowerEvaluator(model, "cuda")
xd = x.to("cuda", fused.dtype).contiguous()
cud = TorchTowerEvaluator(model, "cuda", fused.dtype)
for n in (1, 2, 3, 4, 7, 300, nref):
    got = fused(xd[:n]).cpu()
    torch.cuda.synchronize()
    err = (got - ref[:n]).abs().max().item()
    print(f"n={n}: fused vs fp32 max err {err:.3e}", flush=True)
got = fused(xd[:nref]).cpu()
oth = cud(xd[:nref]).cpu()
print(f"fused vs fp32: max {(got - ref).abs().max():.3e} mean {(got - ref).abs().mean():.3e}")
print(f"cudnn vs fp32: max {(oth - ref).abs().max():.3e} mean {(oth - ref).abs().mean():.3e}")
print(f"fused vs cudnn: max {(got - oth).abs().max():.3e}")
print("ref sample", ref[:4].tolist(), "fused", got[:4].tolist())
for name, ev in (("fused", fused), ("cudnn", cud)):
    for _ in range(2):
        ev(xd)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        ev(xd)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"{name}: {ms:.2f} ms per {B} leaves  {B / ms * 1e3:.3e} leaves/s  {B * fl / ms / 1e9:.1f} TFLOP/s", flush=True)
