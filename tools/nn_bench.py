"""Ad-hoc: accuracy + speed of the value-net forward variants on the GPU (development aid)."""
import sys, time
import torch
from zeroclone_b200.evaluator import TorchTowerEvaluator as NetEvaluator, tower_flops_per_leaf
from zeroclone_b200.models.connect4_value.network import ValueNetwork

torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = ValueNetwork().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
x = (torch.rand(B, 2, 6, 7) < 0.3).float()
x[:, 1] *= (1 - x[:, 0])
with torch.no_grad():
    ref = model(x[:8192]).view(-1)
fl = tower_flops_per_leaf(2, 6, 7)
for dtype in (torch.bfloat16, torch.float16):
    for rfp32 in (False,):
        for chunk in (8192, 32768, 131072):
            ev = NetEvaluator(model, "cuda", dtype, chunk=chunk)
            xd = x.to("cuda", dtype)
            err = (ev(xd[:8192]).cpu() - ref).abs()
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
            print(f"{dtype} res_fp32={rfp32} chunk={chunk}: {ms:.2f} ms  {B / ms * 1e3:.3e} leaves/s  {B * fl / ms / 1e9:.1f} TFLOP/s  "
                  f"err max {err.max():.2e} mean {err.mean():.2e}")
