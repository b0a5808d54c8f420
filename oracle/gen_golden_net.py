#!/usr/bin/env python
"""Generate tests/golden/net_wide.npz: true-fp32 CPU outputs of the network on 256 positions.

  ref_*    the UNMODIFIED reference ChessNet (ai/model.py:27-77, 5 x 512), torch CPU fp32, for the two weight recipes
           of oracle/gen_golden.py (`init`: torch.manual_seed(0) default init; `bnrand`: + randomised BN statistics);
  t20_*    the 20-block x 256-channel tower of BASELINE config 5.  The reference class cannot express that shape
           (ai/model.py:34-40 hard-codes 2 stem convs + 5 x 512), so this is the builder's container
           (knightvision_b200.model.ChessNet(stem=256, tower=256, blocks=20, conv2=False)) run through the same graph
           in plain torch CPU fp32 (`fp32_reference_forward`) — no cuDNN, no TF32.
A full [256, 4096] policy tensor per recipe would be 4 MB; the fixture keeps, per position, 256 seeded random policy
columns, the 8 largest logits (index + value) and the value.   Run:  python oracle/gen_golden_net.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_golden as G  # noqa: E402

N_POS, N_COLS, N_T20 = 256, 256, 128


def bn_randomise(net):
    g = torch.Generator().manual_seed(1)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.05)
            m.running_var.copy_(1.0 + 0.2 * torch.rand(m.running_var.shape, generator=g))
            m.weight.data.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
            m.bias.data.copy_(0.05 * torch.randn(m.bias.shape, generator=g))


def pack(pol, val, cols):
    pol = pol.numpy()
    top = np.argsort(-pol, axis=1)[:, :8]
    return dict(cols=np.take_along_axis(pol, cols, axis=1), top_idx=top.astype(np.int32),
                top_val=np.take_along_axis(pol, top, axis=1), value=val.numpy())


def main():
    torch.set_num_threads(os.cpu_count())
    refai, RefNet = G.import_ai()
    play = np.load(os.path.join(G.GOLD, "playouts.npz"))["line_in"]
    syn = np.load(os.path.join(G.GOLD, "synthetic.npz"))["line_in"]
    rng = np.random.default_rng(2026)
    lines = np.concatenate([play[rng.permutation(len(play))[:N_POS - 48]], syn[0::2][rng.permutation(1500)[:48]]])
    planes = np.stack([refai.encode_board(G.L.unpack_fields(l)["board"]) for l in lines]).astype(np.float32)
    cols = np.stack([np.sort(rng.permutation(4096)[:N_COLS]) for _ in range(N_POS)]).astype(np.int64)
    x = torch.from_numpy(planes)
    out = dict(lines=lines, cols=cols.astype(np.int32))
    for variant in ("init", "bnrand"):
        torch.manual_seed(0)
        net = RefNet().eval()
        if variant == "bnrand":
            bn_randomise(net)
        with torch.no_grad():
            pol, val = net(x)
        for k, v in pack(pol, val, cols).items():
            out[f"ref_{variant}_{k}"] = v
        print("reference", variant, "value range", float(val.min()), float(val.max()), flush=True)
    # 20 x 256 tower, builder container, plain torch CPU fp32
    from knightvision_b200.model import ChessNet, fp32_reference_forward
    torch.manual_seed(2)
    net = ChessNet(stem=256, tower=256, blocks=20, conv2=False).eval()
    bn_randomise(net)
    with torch.no_grad():
        pol, val = fp32_reference_forward(net, x[:N_T20])
    for k, v in pack(pol, val, cols[:N_T20]).items():
        out[f"t20_{k}"] = v
    np.savez_compressed(os.path.join(G.GOLD, "net_wide.npz"), **out)
    print("net_wide.npz written", os.path.getsize(os.path.join(G.GOLD, "net_wide.npz")) // 1024, "KB")


if __name__ == "__main__":
    main()
