"""Probe cuDNN fused conv(+add)+bias+relu ops on B200 for the value tower (development aid)."""
import sys
import torch
import torch.nn.functional as F
from zeroclone_b200.evaluator import TorchTowerEvaluator as NetEvaluator, tower_flops_per_leaf, _fold
from zeroclone_b200.models.connect4_value.network import ValueNetwork

torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = ValueNetwork().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
x = (torch.rand(B, 2, 6, 7) < 0.3).float()
with torch.no_grad():
    ref = model(x[:8192]).view(-1)
fl = tower_flops_per_leaf(2, 6, 7)


class Fused(NetEvaluator):
    @torch.inference_mode()
    def _forward_chunk(self, planes):
        x = planes.contiguous(memory_format=torch.channels_last)
        s, p, d = [1, 1], [1, 1], [1, 1]
        x = torch.cudnn_convolution_relu(x, self.stem[0], self.stem[1], s, p, d, 1)
        for w1, b1, w2, b2 in self.blocks:
            y = torch.cudnn_convolution_relu(x, w1, b1, s, p, d, 1)
            x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, s, p, d, 1)
        pooled = x.float().mean(dim=(2, 3))
        return torch.tanh(torch.addmm(self.lin_b, pooled, self.lin_w)).view(-1)


for dtype in (torch.bfloat16, torch.float16):
    for cls in (NetEvaluator, Fused):
        for chunk in (16384, 32768, 131072):
            try:
                ev = cls(model, "cuda", dtype, chunk=chunk)
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
                print(f"{cls.__name__} {dtype} chunk={chunk}: {ms:.2f} ms {B / ms * 1e3:.3e} leaves/s {B * fl / ms / 1e9:.1f} TFLOP/s err max {err.max():.2e}", flush=True)
            except Exception as e:
                print(cls.__name__, dtype, chunk, "FAILED", repr(e)[:300], flush=True)
