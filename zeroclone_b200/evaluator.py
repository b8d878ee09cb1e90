"""Batched neural leaf evaluation: the `network_latest` / `network_at_path` modes of the
reference's Value class (engine/value_functions.py:61-129) without the worker thread and the
per-state queues -- every tree's pending leaves are already one contiguous device batch, so the
evaluation is ONE forward per search batch.

`NetEvaluator` (= FusedTowerEvaluator) is the product path: the whole tower as one hand-written
sm_100a kernel (csrc/tower.cuh) behind the C-ABI (zc_tower_*).  `TorchTowerEvaluator` is the same
network through PyTorch/cuDNN; nothing in the package calls it -- it exists for A/B timing (tools/) and
as a second opinion in tests.  BatchNorm (eval mode, network.py:15,18,30) is folded into
the convolution weights in fp32 before the cast to the compute dtype in both.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn.functional as F
from torch import nn


def _fold(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    w = conv.weight.detach().float()
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    if conv.bias is not None:
        b = b + conv.bias.detach().float() * scale
    return w * scale.view(-1, 1, 1, 1), b


class TorchTowerEvaluator:
    """PyTorch/cuDNN reference of the fused kernel.  values = tanh(head(res(stem(planes))))  for planes[B, C, H, W] in `dtype`; returns float32[B]."""

    def __init__(self, model: nn.Module, device="cuda", dtype: torch.dtype = torch.bfloat16, chunk: int = 131072):
        model = model.eval()
        self.device = torch.device(device)
        self.dtype = dtype
        self.chunk = chunk
        self.in_planes = model.stem[0].in_channels
        cl = torch.channels_last

        def prep(w, b):
            return (w.to(self.device, dtype).contiguous(memory_format=cl), b.to(self.device, dtype))

        self.stem = prep(*_fold(model.stem[0], model.stem[1]))
        self.blocks: List[tuple] = []
        for blk in model.res:
            self.blocks.append(prep(*_fold(blk.seq[0], blk.seq[1])) + prep(*_fold(blk.seq[3], blk.seq[4])))
        lin = model.head[2]
        self.lin_w = lin.weight.detach().float().to(self.device).t().contiguous()   # [C, 1]
        self.lin_b = lin.bias.detach().float().to(self.device)
        self.flops_per_leaf = None

    @torch.inference_mode()
    def _forward_chunk(self, planes: torch.Tensor) -> torch.Tensor:
        # cuDNN runtime-fused conv + bias (+ residual) + ReLU: one kernel per convolution, no separate
        # elementwise passes over the 1.4 GB activation tensor (they were 2/3 of the forward time,
        # profiles/r1_launches_c4_value_net_summary.csv).
        one = [1, 1]
        x = planes.contiguous(memory_format=torch.channels_last)
        x = torch.cudnn_convolution_relu(x, self.stem[0], self.stem[1], one, one, one, 1)
        for w1, b1, w2, b2 in self.blocks:
            y = torch.cudnn_convolution_relu(x, w1, b1, one, one, one, 1)
            x = torch.cudnn_convolution_add_relu(y, w2, x, 1.0, b2, one, one, one, 1)     # relu(conv(y) + x + b)
        pooled = x.float().mean(dim=(2, 3))
        return torch.tanh(torch.addmm(self.lin_b, pooled, self.lin_w)).view(-1)

    @torch.inference_mode()
    def forward_unfused(self, planes: torch.Tensor) -> torch.Tensor:
        """The same network as separate conv / add / relu calls (kept for A/B timing)."""
        x = planes.contiguous(memory_format=torch.channels_last)
        x = F.relu_(F.conv2d(x, self.stem[0], self.stem[1], padding=1))
        for w1, b1, w2, b2 in self.blocks:
            y = F.relu_(F.conv2d(x, w1, b1, padding=1))
            y = F.conv2d(y, w2, b2, padding=1)
            x = F.relu_(y.add_(x))
        pooled = x.float().mean(dim=(2, 3))
        return torch.tanh(torch.addmm(self.lin_b, pooled, self.lin_w)).view(-1)

    @torch.inference_mode()
    def __call__(self, planes: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        n = planes.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=planes.device)
        for s in range(0, n, self.chunk):
            out[s:s + self.chunk] = self._forward_chunk(planes[s:s + self.chunk])
        return out


class FusedTowerEvaluator:
    """The same network as ONE hand-written sm_100a kernel (csrc/tower.cuh, zc_tower_* in
    include/zc_b200.h): activations of a leaf never leave shared memory between the stem and the head.
    Call: evaluator(planes[B,C,H,W] contiguous in `self.dtype`, out=float32[B]).

    dtype: the tensor-core operand format, torch.bfloat16 or torch.float16 (same MMA rate; measured on a power-capped
    B200 fp16 runs ~3 % slower: wider multipliers draw more power).  Default (None) is the cheapest format that meets
    the north-star tolerance -- root values of an 800-simulation search within 1e-3 of the fp32 reference -- for the
    game: bf16 on Connect Four (worst root of 128: 1.5e-4), fp16 on chess (bf16: 3.3e-3 at the worst root of 32,
    fp16 passes; fp16 is also what the reference itself evaluates in on a GPU, value_functions.py:6).
    tests/test_gpu_parity_bench_sets.py holds both measurements."""

    DEFAULT_DTYPE = {2: torch.bfloat16, 17: torch.float16}      # by input planes: Connect Four, chess

    def __init__(self, model: nn.Module, device="cuda", dtype: torch.dtype | None = None):
        import ctypes as C

        from . import _ffi

        if dtype is None:
            dtype = self.DEFAULT_DTYPE.get(model.stem[0].in_channels, torch.float16)
        if dtype not in (torch.bfloat16, torch.float16):
            raise ValueError("the fused tower computes in fp16 or bf16 operands (fp32 accumulation)")
        self.dtype = dtype
        plane = _ffi.PLANE_F16 if dtype == torch.float16 else _ffi.PLANE_BF16
        model = model.eval()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FusedTowerEvaluator needs a CUDA device: libzc_b200 has no CPU path")
        self.in_planes = model.stem[0].in_channels
        game = {2: _ffi.GAME_C4, 17: _ffi.GAME_CHESS}.get(self.in_planes)
        if game is None or model.stem[0].out_channels != 128:
            raise ValueError("fused tower supports the 128-channel tower on Connect Four (2 planes) or chess (17 planes)")
        self.n_blocks = len(model.res)
        conv_w, conv_b, head_w, head_b = self._folded(model)
        self._h = C.c_void_p()
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._index = index
        _ffi.check(_ffi.lib().zc_tower_create(game, index, self.n_blocks, plane, conv_w.ctypes.data_as(C.c_void_p),
                                              conv_b.ctypes.data_as(C.c_void_p), head_w.ctypes.data_as(C.c_void_p),
                                              head_b, C.byref(self._h)))

    @staticmethod
    def _folded(model: nn.Module):
        """BatchNorm (eval mode) folded into every convolution in fp32: the plain float arrays zc_tower_create takes"""
        import numpy as np

        ws, bs = [], []
        w, b = _fold(model.stem[0], model.stem[1])
        ws.append(w.reshape(-1))
        bs.append(b)
        for blk in model.res:
            for conv, bn in ((blk.seq[0], blk.seq[1]), (blk.seq[3], blk.seq[4])):
                w, b = _fold(conv, bn)
                ws.append(w.reshape(-1))
                bs.append(b)
        conv_w = np.ascontiguousarray(torch.cat(ws).cpu().numpy(), dtype=np.float32)
        conv_b = np.ascontiguousarray(torch.cat(bs).cpu().numpy(), dtype=np.float32)
        lin = model.head[2]
        head_w = np.ascontiguousarray(lin.weight.detach().float().view(-1).cpu().numpy(), dtype=np.float32)
        return conv_w, conv_b, head_w, float(lin.bias.detach().float().item())

    def update_weights(self, model: nn.Module) -> None:
        """Freshly trained weights into the resident tower (zc_tower_update_weights): what the training loop does
        between cycles instead of rebuilding the evaluator (scripts/train.py:143-146 + value_functions.py:104-112)."""
        import ctypes as C

        from . import _ffi

        if len(model.res) != self.n_blocks or model.stem[0].in_channels != self.in_planes:
            raise ValueError("update_weights: the model's architecture differs from the tower's")
        conv_w, conv_b, head_w, head_b = self._folded(model)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _ffi.check(_ffi.lib().zc_tower_update_weights(self._h, conv_w.ctypes.data_as(C.c_void_p), conv_b.ctypes.data_as(C.c_void_p),
                                                      head_w.ctypes.data_as(C.c_void_p), head_b, C.c_void_p(stream)))

    def close(self) -> None:
        from . import _ffi

        if getattr(self, "_h", None):
            _ffi.lib().zc_tower_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fault(self) -> int:
        """0, or 0x80000000 | wait tag if the kernel gave up on one of its bounded barrier waits (a protocol bug, never expected);
        readable even after the CUDA context has been poisoned by the trap (the word lives in pinned host memory).  Clears it."""
        from . import _ffi

        return int(_ffi.lib().zc_tower_fault(self._h)) if getattr(self, "_h", None) else 0

    @property
    def launches(self) -> int:
        from . import _ffi

        return int(_ffi.lib().zc_tower_launches(self._h))

    def __call__(self, planes: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        import ctypes as C

        from . import _ffi

        if planes.dtype != self.dtype or not planes.is_contiguous() or planes.device.type != "cuda":
            raise ValueError(f"planes must be a contiguous {self.dtype} CUDA tensor [B, C, H, W]")
        n = planes.shape[0]
        if out is None:
            out = torch.empty(n, dtype=torch.float32, device=planes.device)
        stream = torch.cuda.current_stream(planes.device).cuda_stream
        _ffi.check(_ffi.lib().zc_tower_forward(self._h, C.c_void_p(planes.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                               C.c_void_p(stream)))
        return out


NetEvaluator = FusedTowerEvaluator


def tower_flops_per_leaf(in_planes: int, h: int, w: int, channels: int = 128, blocks: int = 8) -> float:
    """2*MACs of one forward: stem + 2*blocks convs + head."""
    cells = h * w
    return 2.0 * (cells * in_planes * channels * 9 + 2 * blocks * cells * channels * channels * 9 + channels)
