"""Training-side 3x3 convolution of the policy/value tower as an autograd function over the sm_100a kernels.

The reference trains `ai/model.py` under autocast with cuDNN convolutions (scripts/train.py:158-181).  Here the tower
convolutions (99.8 % of the FLOPs) run on the hand-written tcgen05 kernels in all three directions:

    forward   y  = conv(x, W) + b        kv_conv3x3_fprop  (the self-play tower kernel on caller-owned tensors)
    dgrad     dx = conv(dy, W')          the same kernel on the flip-transposed weights (kv_conv3x3_pack)
    wgrad     dW = x (*) dy              kv_conv3x3_wgrad  (MN-major tcgen05 operands straight from NHWC, split-K)

Activations are bf16 in NHWC memory (torch channels_last), accumulation is fp32, the parameter and its gradient stay
fp32 — the bf16 analogue of the reference's fp16 autocast.  Train-mode BatchNorm + ReLU (+ the residual add) is one
fused forward and one fused backward operator over the same NHWC bf16 tensors (csrc/kv_bn.cu: HBM-bound passes,
deterministic two-stage reductions).  The 12-channel stem convolution, the heads, the loss and the optimizer stay
ordinary PyTorch (0.2 % of the FLOPs).
There is no CPU path: the function raises without the CUDA library.
"""
from __future__ import annotations

import torch


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """Logical NCHW tensor -> contiguous bf16 [N,8,8,C] view (copying only if it is not channels_last bf16 already)."""
    t = t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    v = t.permute(0, 2, 3, 1)
    return v if v.is_contiguous() else v.contiguous()


class _Conv3x3B200(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, eng):
        xh = _nhwc(x)
        y = eng.conv3x3_fprop(xh, eng.conv3x3_pack(weight), bias.detach().float().contiguous() if bias is not None else None)
        ctx.save_for_backward(xh, weight)
        ctx.eng = eng
        ctx.has_bias = bias is not None
        return y.permute(0, 3, 1, 2)          # logical NCHW, channels_last memory

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        xh, weight = ctx.saved_tensors
        eng = ctx.eng
        gy = _nhwc(grad_out)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = eng.conv3x3_fprop(gy, eng.conv3x3_pack(weight, flip_transpose=True)).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dw = eng.conv3x3_wgrad(xh, gy).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = eng.channel_sum(gy)
        return dx, dw, db, None


def conv3x3_b200(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, eng) -> torch.Tensor:
    """3x3 / padding 1 convolution of a [N,C,8,8] CUDA tensor through the B200 kernels, differentiable in x, weight, bias."""
    if not x.is_cuda:
        raise RuntimeError("conv3x3_b200 needs CUDA tensors (knightvision_b200 has no CPU fallback)")
    if tuple(x.shape[2:]) != (8, 8) or tuple(weight.shape[2:]) != (3, 3):
        raise ValueError("conv3x3_b200: expects [N,C,8,8] inputs and [Cout,Cin,3,3] weights")
    return _Conv3x3B200.apply(x, weight, bias, eng)


class _BNReLUB200(torch.autograd.Function):
    """Train-mode BatchNorm2d + ReLU (+ residual add), csrc/kv_bn.cu.  Inputs/outputs are logical NCHW tensors in
    channels_last bf16 memory; gamma/beta and their gradients are fp32."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, z, gamma, beta, residual, eng, running_mean, running_var, momentum, eps, relu):
        zh = _nhwc(z)
        rh = _nhwc(residual) if residual is not None else None
        g = gamma.detach().float().contiguous()
        y, mean, rstd = eng.bn_relu_fwd(zh, g, beta.detach().float().contiguous(), running_mean, running_var, momentum, eps,
                                        residual=rh, relu=relu)
        ctx.save_for_backward(zh, y, g, mean, rstd)
        ctx.eng, ctx.relu, ctx.has_res = eng, relu, residual is not None
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        zh, y, g, mean, rstd = ctx.saved_tensors
        dz, dres, dgamma, dbeta = ctx.eng.bn_relu_bwd(_nhwc(grad_out), y, zh, g, mean, rstd, relu=ctx.relu,
                                                      want_dres=ctx.has_res and ctx.needs_input_grad[3])
        return (dz.permute(0, 3, 1, 2), dgamma, dbeta, dres.permute(0, 3, 1, 2) if dres is not None else None,
                None, None, None, None, None, None)


def bn_relu_b200(z: torch.Tensor, bn: torch.nn.BatchNorm2d, eng, residual: torch.Tensor | None = None, relu: bool = True):
    """relu(bn(z) [+ residual]) with batch statistics, updating bn.running_mean / running_var / num_batches_tracked as
    torch.nn.BatchNorm2d does in training mode."""
    if not bn.training or not bn.track_running_stats or bn.momentum is None:
        raise RuntimeError("bn_relu_b200 implements training-mode BatchNorm2d with a fixed momentum")
    with torch.no_grad():
        bn.num_batches_tracked += 1
    return _BNReLUB200.apply(z, bn.weight, bn.bias, residual, eng, bn.running_mean, bn.running_var, float(bn.momentum),
                             float(bn.eps), relu)


class _ResidualBlockB200(torch.autograd.Function):
    """ai/model.py:19-25 as ONE autograd node: relu(bn2(conv2(relu(bn1(conv1(x))))) + x) in training mode.

    Same kernels as conv3x3_b200 / bn_relu_b200 chained, but the backward knows the block's shape: the gradient of the
    skip connection (the ReLU-masked incoming gradient) goes into the `residual` input of the last dgrad launch, so
    dx = dgrad(dz1) + g is produced by the convolution epilogue instead of a separate elementwise pass, and the
    block is 1 autograd node instead of 4."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, w1, b1, g1, be1, w2, b2, g2, be2, eng, bn1, bn2):
        xh = _nhwc(x)
        f32 = lambda t: t.detach().float().contiguous()
        z1 = eng.conv3x3_fprop(xh, eng.conv3x3_pack(w1), f32(b1))
        t, m1, r1 = eng.bn_relu_fwd(z1, f32(g1), f32(be1), bn1.running_mean, bn1.running_var, float(bn1.momentum),
                                    float(bn1.eps))
        z2 = eng.conv3x3_fprop(t, eng.conv3x3_pack(w2), f32(b2))
        y, m2, r2 = eng.bn_relu_fwd(z2, f32(g2), f32(be2), bn2.running_mean, bn2.running_var, float(bn2.momentum),
                                    float(bn2.eps), residual=xh)
        ctx.save_for_backward(xh, z1, t, z2, y, w1, w2, f32(g1), f32(g2), m1, r1, m2, r2)
        ctx.eng = eng
        return y.permute(0, 3, 1, 2)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        xh, z1, t, z2, y, w1, w2, g1, g2, m1, r1, m2, r2 = ctx.saved_tensors
        eng = ctx.eng
        dy = _nhwc(grad_out)
        dz2, g, dg2, dbe2 = eng.bn_relu_bwd(dy, y, z2, g2, m2, r2, want_dres=True)
        dw2 = eng.conv3x3_wgrad(t, dz2)
        db2 = eng.channel_sum(dz2)
        dt = eng.conv3x3_fprop(dz2, eng.conv3x3_pack(w2, flip_transpose=True))
        dz1, _, dg1, dbe1 = eng.bn_relu_bwd(dt, t, z1, g1, m1, r1)
        dw1 = eng.conv3x3_wgrad(xh, dz1)
        db1 = eng.channel_sum(dz1)
        dx = eng.conv3x3_fprop(dz1, eng.conv3x3_pack(w1, flip_transpose=True), residual=g)     # + skip gradient, fused
        return (dx.permute(0, 3, 1, 2), dw1.to(w1.dtype), db1, dg1, dbe1, dw2.to(w2.dtype), db2, dg2, dbe2, None, None, None)


def residual_block_b200(x: torch.Tensor, block, eng) -> torch.Tensor:
    """One ResidualBlock (conv1, bn1, conv2, bn2 modules with biases) in training mode through the B200 kernels."""
    for bn in (block.bn1, block.bn2):
        if not bn.training or not bn.track_running_stats or bn.momentum is None:
            raise RuntimeError("residual_block_b200 implements training-mode BatchNorm2d with a fixed momentum")
    with torch.no_grad():
        block.bn1.num_batches_tracked += 1
        block.bn2.num_batches_tracked += 1
    return _ResidualBlockB200.apply(x, block.conv1.weight, block.conv1.bias, block.bn1.weight, block.bn1.bias,
                                    block.conv2.weight, block.conv2.bias, block.bn2.weight, block.bn2.bias, eng,
                                    block.bn1, block.bn2)


def bn_supported(C: int) -> bool:
    return C in (64, 128, 256, 512, 1024)


def supported(cin: int, cout: int) -> bool:
    """Channel counts the three kernels accept (the tower layers: 256 / 512 channels)."""
    return cin % 256 == 0 and cout % 256 == 0 and cin <= 512 and cout <= 512
