"""Host-side mirror of the reference network interface (ai/model.py:27-77).

`ChessNet` keeps the reference's contract — `forward(x[B,12,8,8]) -> (policy[B,4096] logits, value[B,1] tanh)`
and a `state_dict()` with the reference's 104 keys (conv1.*, bn1.*, conv2.*, bn2.*, res_blocks.{i}.{conv1,bn1,
conv2,bn2}.*, policy_conv/bn/fc.*, value_conv/bn/fc1/fc2.*), so reference checkpoints load unchanged
(ai/model_utils.py:16-20) — but `forward` runs the hand-written sm_100a kernels of libkv_b200.so
(TMA + tcgen05 implicit-GEMM tower, fused stem and heads) in bf16 with fp32 accumulation, eval-mode BatchNorm.
It is an inference module: there is no autograd through it and no CPU path.

`fp32_reference_forward` is plain PyTorch fp32 of the same graph; only tests and bench baselines call it.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N


class _Block(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(ch, ch, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(ch)
        self.conv2 = nn.Conv2d(ch, ch, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(ch)


class ChessNet(nn.Module):
    """Parameter container with the reference layout + B200 forward.  The reference architecture is the
    default (stem 256, tower 512, 5 blocks); `ChessNet(stem=256, tower=256, blocks=20, conv2=False)` is the
    20x256 tower of BASELINE config 5 (same heads and I/O contract)."""

    def __init__(self, verbose: bool = False, stem: int = 256, tower: int = 512, blocks: int = 5, conv2: bool = True,
                 max_batch: int = 4096):
        super().__init__()
        self.verbose = verbose
        self.arch = (stem, tower, blocks, bool(conv2))
        self.max_batch = max_batch
        self.conv1 = nn.Conv2d(12, stem, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(stem)
        if conv2:
            self.conv2 = nn.Conv2d(stem, tower, 3, padding=1)
            self.bn2 = nn.BatchNorm2d(tower)
        self.res_blocks = nn.ModuleList([_Block(tower) for _ in range(blocks)])
        self.policy_conv = nn.Conv2d(tower, 2, 1)
        self.policy_bn = nn.BatchNorm2d(2)
        self.policy_fc = nn.Linear(128, 4096)
        self.value_conv = nn.Conv2d(tower, 1, 1)
        self.value_bn = nn.BatchNorm2d(1)
        self.value_fc1 = nn.Linear(64, 512)
        self.value_fc2 = nn.Linear(512, 1)
        self._engine = None
        self._dirty = True

    # ---- weights -----------------------------------------------------------------------------------
    def weight_blob(self) -> torch.Tensor:
        """fp32 state_dict tensors concatenated in key order (num_batches_tracked skipped): kv_net_load's input."""
        parts = [v.detach().reshape(-1).to(torch.float32).cpu() for k, v in self.state_dict().items()
                 if not k.endswith("num_batches_tracked")]
        return torch.cat(parts).contiguous()

    def load_state_dict(self, *a, **kw):
        self._dirty = True
        return super().load_state_dict(*a, **kw)

    def mark_weights_changed(self):
        self._dirty = True

    def attach(self, engine, max_batch: int | None = None):
        """Bind to an Engine (one kv_ctx per GPU) and upload + fold the weights."""
        self._engine = engine
        if max_batch:
            self.max_batch = max_batch
        if getattr(engine, "net_geometry", None) != (self.arch, self.max_batch):
            engine.net_create(*self.arch, self.max_batch)        # (re)allocates activations / weight buffers
            engine.net_geometry = (self.arch, self.max_batch)
        self.sync_weights()
        return self

    def sync_weights(self):
        """Hand the current parameters to the attached engine.  Parameters that live on the engine's GPU go device to
        device: concatenated straight into the engine's fp32 staging blob (the buffer an NCCL broadcast would target),
        then folded / converted on the device (kv_net_commit_weights) — no host round trip.  This is the weight bridge
        between the trainer's fp32 master weights and the self-play engine (ai/model_utils.py:10-29 role)."""
        eng = self._engine
        p = next(self.parameters())
        if p.is_cuda and p.device == eng.device:
            parts = [v.detach().reshape(-1).to(torch.float32) for k, v in self.state_dict().items()
                     if not k.endswith("num_batches_tracked")]
            torch.cat(parts, out=eng.net_blob_tensor())
            eng.net_commit()
        else:
            eng.net_load(self.weight_blob())
        self._dirty = False

    def _ensure(self, device):
        if self._engine is None:
            from .engine import Engine
            self.attach(Engine(device))
        elif self._dirty:
            self.sync_weights()
        return self._engine

    # ---- forward (B200 kernels) ------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        if self.training:
            raise N.KVError("knightvision_b200.ChessNet.forward is inference only (call .eval(); BatchNorm is folded)")
        if not x.is_cuda:
            if not torch.cuda.is_available():
                raise N.KVError("knightvision_b200.ChessNet needs a CUDA device (no CPU fallback)")
            x = x.cuda()
        eng = self._ensure(x.device)
        x = x.to(torch.float32).contiguous()
        pol, val = [], []
        for i in range(0, x.shape[0], self.max_batch):
            p, v = eng.net_forward_planes(x[i:i + self.max_batch])
            pol.append(p)
            val.append(v)
        return torch.cat(pol), torch.cat(val).unsqueeze(1)

    @torch.no_grad()
    def forward_lines(self, lines: torch.Tensor):
        """Same, from device board lines (the fused encode+stem path the self-play engine uses)."""
        eng = self._ensure(lines.device)
        p, v = eng.net_forward(lines)
        return p, v.unsqueeze(1)


def fp32_reference_forward(net: ChessNet, x: torch.Tensor):
    """Plain PyTorch fp32, eval-mode BN: the graph of ai/model.py:51-77.  Test/baseline use only.
    True fp32 on a GPU as well: TF32 is switched off for the duration of the call (cuDNN convolutions and cuBLAS
    matmuls would otherwise run with 10-bit mantissas, which is not a reference for a 2e-2 tolerance)."""
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return _fp32_graph(net, x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def _fp32_graph(net: ChessNet, x: torch.Tensor):
    def bn(m, t):
        return F.batch_norm(t, m.running_mean, m.running_var, m.weight, m.bias, False, 0.0, m.eps)
    x = x.to(next(net.parameters()).device, torch.float32)
    h = F.relu(bn(net.bn1, net.conv1(x)))
    if net.arch[3]:
        h = F.relu(bn(net.bn2, net.conv2(h)))
    for b in net.res_blocks:
        t = F.relu(bn(b.bn1, b.conv1(h)))
        h = F.relu(bn(b.bn2, b.conv2(t)) + h)
    p = F.relu(bn(net.policy_bn, net.policy_conv(h))).flatten(1)
    p = net.policy_fc(p)
    v = F.relu(bn(net.value_bn, net.value_conv(h))).flatten(1)
    v = torch.tanh(net.value_fc2(F.relu(net.value_fc1(v))))
    return p, v
